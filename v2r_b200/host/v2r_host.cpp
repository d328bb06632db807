// v2r_host.cpp -- tools:: and idw:: of the C++ host (see v2r_host.hpp for the reference citations).
#include "v2r_host.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace v2r {

static void check(int rc, const char *what) {
    if (rc != V2R_IDW_OK) throw Error(std::string(what) + ": " + v2r_idw_last_error());
}

namespace tools {

// Page-locked when a CUDA driver is present (async DMA); plain aligned memory otherwise, so that the readers
// and `--plan` work on a box without a GPU.  This is memory only: every solve still needs the device.
static bool g_pageable_only = false;
static void *host_alloc(size_t bytes, bool &is_pinned) {
    void *p = nullptr;
    if (!g_pageable_only && v2r_idw_alloc_pinned(bytes, &p) == V2R_IDW_OK) { is_pinned = true; return p; }
    g_pageable_only = true;
    is_pinned = false;
    p = aligned_alloc(64, (bytes + 63) / 64 * 64);
    if (!p) throw Error("out of host memory");
    return p;
}
static void host_free(void *p, bool is_pinned) {
    if (!p) return;
    if (is_pinned) v2r_idw_free_pinned(p);
    else free(p);
}

PointSet::PointSet(int64_t capacity, bool with_keys) : keys_(with_keys) { reserve(capacity > 0 ? capacity : 1); }
PointSet::PointSet(PointSet &&o) noexcept { *this = std::move(o); }
PointSet &PointSet::operator=(PointSet &&o) noexcept {
    if (this != &o) {
        release();
        x = o.x; y = o.y; w = o.w; key_r = o.key_r; key_c = o.key_c; n_ = o.n_; cap_ = o.cap_; keys_ = o.keys_; pinned_ = o.pinned_;
        o.x = o.y = o.w = nullptr; o.key_r = o.key_c = nullptr; o.n_ = o.cap_ = 0;
    }
    return *this;
}
PointSet::~PointSet() { release(); }
void PointSet::release() {
    for (void *p : {(void *)x, (void *)y, (void *)w, (void *)key_r, (void *)key_c}) host_free(p, pinned_);
    x = y = w = nullptr;
    key_r = key_c = nullptr;
    n_ = cap_ = 0;
}
void PointSet::reserve(int64_t cap) {
    if (cap <= cap_) return;
    bool now_pinned = pinned_;
    auto grow = [&](auto *&arr) {
        using T = std::remove_reference_t<decltype(*arr)>;
        bool pin = false;
        T *np = (T *)host_alloc((size_t)cap * sizeof(T), pin);
        if (arr) { memcpy(np, arr, (size_t)n_ * sizeof(T)); host_free(arr, pinned_); }
        arr = np;
        now_pinned = pin;
    };
    grow(x); grow(y); grow(w);
    if (keys_) { grow(key_r); grow(key_c); }
    pinned_ = now_pinned;
    cap_ = cap;
}
void PointSet::push_back(double px, double py, double pw) {
    if (n_ == cap_) reserve(cap_ * 2 > 1024 ? cap_ * 2 : 1024);
    x[n_] = px; y[n_] = py; w[n_] = pw;
    ++n_;
}

void GetDimensions(const Info &xi, const Info &yi, int64_t &rows, int64_t &cols) {
    check(v2r_idw_get_dimensions(xi.Min, xi.Max, xi.Step, yi.Min, yi.Max, yi.Step, &rows, &cols), "GetDimensions");
}
OrderedPair PointToPair(const Point &p, const Info &xi, const Info &yi) {
    OrderedPair k;
    check(v2r_idw_point_to_pair(p.X, p.Y, xi.Min, xi.Step, yi.Min, yi.Step, &k.R, &k.C), "PointToPair");
    return k;
}
PointSet MakeCoordSpace(const PointSet &in, const Info &xi, const Info &yi) {
    PointSet out(in.size(), true);
    int64_t n = 0;
    check(v2r_idw_make_coord_space(in.size(), in.x, in.y, in.w, xi.Min, xi.Step, yi.Min, yi.Step, out.x, out.y, out.w,
                                   out.key_r, out.key_c, &n), "MakeCoordSpace");
    out.n_ = n;
    return out;
}
std::vector<double> ExpRange(double es, double ee, double ei) {
    int64_t cnt = 0;
    std::vector<double> v(4096);
    check(v2r_idw_exp_range(es, ee, ei, v.data(), (int64_t)v.size(), &cnt), "expRange");
    if (cnt > (int64_t)v.size()) throw Error("exponent range has more than 4096 steps");
    v.resize((size_t)cnt);
    return v;
}

}  // namespace tools

namespace idw {

static std::mutex g_mu;
static v2r_idw_ctx *g_ctx = nullptr;
static int g_ctx_gpus = 0;

v2r_idw_ctx *Context(const Options &opt) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx && g_ctx_gpus != opt.n_gpus) { v2r_idw_destroy(g_ctx); g_ctx = nullptr; }
    if (!g_ctx) {
        check(v2r_idw_init(opt.n_gpus, &g_ctx), "v2r_idw_init");
        g_ctx_gpus = opt.n_gpus;
    }
    return g_ctx;
}

static bool ends_with(const std::string &s, const char *suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

static void check_outfile(const std::string &f) {  // features/idw/idw.go:19-25
    if (!(ends_with(f, ".asc") || ends_with(f, ".xlsx") || ends_with(f, ".tiff") || ends_with(f, ".tif")))
        throw Error("outfile is not supported. Use one of these files: tiff (.tif/.tiff), ascii (.asc), or excel (.xlsx).");
    if (ends_with(f, ".xlsx")) throw Error("excel output is outside the GPU path (cmd/idw.go:155 always writes .tiff)");
}

static std::string fmt_g(double v) {
    char b[64];
    snprintf(b, sizeof b, "%g", v);
    return b;
}
// Go's %v of a float64 (features/idw/idw.go:63 "pow%v"): the shortest decimal that reads back as the same double,
// in %g style -- 0.30000000000000004 stays 0.30000000000000004, 2 prints as 2
static std::string fmt_v(double v) {
    char b[64];
    for (int prec = 1; prec <= 17; ++prec) {
        snprintf(b, sizeof b, "%.*g", prec, v);
        if (strtod(b, nullptr) == v || v != v) break;
    }
    return b;
}

// one streaming session: at most kMaxSweep exponents (include/v2r_idw.h)
static void SweepSession(v2r_idw_ctx *ctx, const tools::PointSet &data, const std::string *outfiles, const tools::Info &xInfo,
                         const tools::Info &yInfo, int64_t rows, int64_t cols, const std::string &proj, const double *pows, int E,
                         int64_t bandRows, int64_t chunkC, const Options &opt);

void SweepSolve(const tools::PointSet &data, const std::vector<std::string> &outfiles, const tools::Info &xInfo,
                const tools::Info &yInfo, const std::string &proj, const std::vector<double> &pows, int64_t bandRows,
                const Channel &channel, const Options &opt, int64_t chunkC) {
    const auto start = std::chrono::steady_clock::now();
    if (outfiles.size() != pows.size() || pows.empty()) throw Error("one outfile per exponent is required");
    for (const auto &f : outfiles) check_outfile(f);
    if (data.size() > 0 && !data.key_r) throw Error("points are not keyed: run MakeCoordSpace first");
    int64_t rows, cols;
    tools::GetDimensions(xInfo, yInfo, rows, cols);
    if (rows < 0 || cols < 0) throw Error("negative grid dimensions");
    v2r_idw_ctx *ctx = Context(opt);
    constexpr size_t kMaxSweep = 4096;  // exponents per library call; a longer sweep runs as several sessions
    for (size_t e0 = 0; e0 < pows.size(); e0 += kMaxSweep) {
        const size_t e1 = std::min(pows.size(), e0 + kMaxSweep);
        SweepSession(ctx, data, outfiles.data() + e0, xInfo, yInfo, rows, cols, proj, pows.data() + e0, (int)(e1 - e0), bandRows, chunkC, opt);
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    if (channel)
        for (double p : pows)
            channel("pow" + fmt_v(p) + " [" + std::to_string(rows) + "X" + std::to_string(cols) + "] completed in " + fmt_g(secs) + "s");
}

static void SweepSession(v2r_idw_ctx *ctx, const tools::PointSet &data, const std::string *outfiles, const tools::Info &xInfo,
                         const tools::Info &yInfo, int64_t rows, int64_t cols, const std::string &proj, const double *pows, int E,
                         int64_t bandRows, int64_t chunkC, const Options &opt) {
    v2r_idw_grid grid{xInfo.Min, xInfo.Step, yInfo.Min, yInfo.Step, rows, cols};
    bool asc = false;
    for (int e = 0; e < E; ++e) asc |= ends_with(outfiles[e], ".asc");
    int64_t band = bandRows > 0 ? bandRows : rows;
    if (!asc && bandRows <= 0) {  // ~1 GiB of raster per band unless the caller fixed the height
        int64_t cap = (int64_t(1) << 27) / (std::max<int64_t>(cols, 1) * E);
        cap = std::max<int64_t>(32, cap / 32 * 32);
        band = std::min(band, cap);
    }
    if (asc) band = rows;  // AAIGrid cannot be updated in place: whole raster in one write
    band = std::max<int64_t>(band, 1);
    void *outp = nullptr;
    check(v2r_idw_alloc_pinned((size_t)band * (size_t)std::max<int64_t>(cols, 1) * E * sizeof(double), &outp), "alloc band");
    struct Free { void *p; ~Free() { v2r_idw_free_pinned(p); } } guard{outp};
    double *out = (double *)outp;

    check(v2r_idw_stream_begin(ctx, data.x, data.y, data.w, data.key_r, data.key_c, data.size(), &grid, pows, E,
                               opt.fp32 ? V2R_IDW_FP32 : V2R_IDW_FP64, band), "v2r_idw_stream_begin");
    struct End { v2r_idw_ctx *c; ~End() { v2r_idw_stream_end(c); } } end_guard{ctx};
    processing::GDalInfo gd = processing::CreateGDalInfo(xInfo.Min, yInfo.Min, xInfo.Step, yInfo.Step, 7, proj);
    std::vector<char> created(E, 0);
    std::vector<double> window;
    if (rows == 0 || cols == 0)
        for (int e = 0; e < E; ++e) processing::WriteGDAL(out, gd, outfiles[e], {0, 0}, {rows, cols}, {0, 0}, true);
    // progress like ChunkSolve's "~10%, i / total" lines (features/idw/idw_chunk.go:77,92-96), per streamed band
    const int64_t total_bands = std::max<int64_t>(1, (rows + band - 1) / band);
    const int64_t progress = std::max<int64_t>(1, total_bands / 10);
    int64_t done = 0;
    for (;;) {
        int64_t r0 = 0, nr = 0;
        check(v2r_idw_stream_next(ctx, out, &r0, &nr), "v2r_idw_stream_next");
        if (nr == 0) break;
        if (opt.progress && bandRows > 0 && ++done % progress == 0)
            opt.progress("~" + std::to_string(std::min<int64_t>(100, 100 * done / total_bands)) + "%, " + std::to_string(done) + " / " +
                         std::to_string(total_bands));
        const int64_t plane = nr * cols;
        // windows of chunkC columns at (r0, c0), like the reference's per-chunk WriteGDAL (features/idw/idw_chunk.go:69-90,
        // tools/processing/io_gdal.go:72-73); a full-width band is written as one window
        const int64_t wcols = (chunkC > 0 && chunkC < cols) ? chunkC : cols;
        for (int e = 0; e < E; ++e) {  // the next band is already computing while these windows are written
            const double *flat = out + (int64_t)e * plane;
            for (int64_t c0 = 0; c0 < cols; c0 += wcols) {
                const int64_t wc = std::min(wcols, cols - c0);
                const double *buf = flat;
                if (wc != cols) {  // WriteGDAL takes a contiguous window
                    window.resize((size_t)(nr * wc));
                    for (int64_t r = 0; r < nr; ++r) memcpy(&window[(size_t)(r * wc)], flat + r * cols + c0, (size_t)wc * sizeof(double));
                    buf = window.data();
                }
                processing::WriteGDAL(buf, gd, outfiles[e], {r0, c0}, {rows, cols}, {nr, wc}, !created[e]);
                created[e] = 1;
            }
        }
    }
}

void FullSolve(const tools::PointSet &data, const std::string &outfile, const tools::Info &xInfo, const tools::Info &yInfo,
               const std::string &proj, double pow, const Channel &channel, const Options &opt) {
    SweepSolve(data, {outfile}, xInfo, yInfo, proj, {pow}, 0, channel, opt);
}

void ChunkSolve(const tools::PointSet &data, const std::string &outfile, const tools::Info &xInfo, const tools::Info &yInfo,
                int64_t chunkR, int64_t chunkC, const std::string &proj, double pow, const Channel &channel,
                const Options &opt) {
    int64_t rows, cols;
    tools::GetDimensions(xInfo, yInfo, rows, cols);
    if (chunkR > rows || chunkC > cols) { chunkR = rows / 4; chunkC = cols / 4; }  // features/idw/idw_chunk.go:48-53
    if (chunkR <= 0 || chunkC <= 0) throw Error("chunk size is not usable for this grid");
    // On the GPU a chunk is a write granule: bands of chunkR rows stream out while the next band computes and are
    // written as windows of chunkC columns, the reference's chunk-write protocol.
    SweepSolve(data, {outfile}, xInfo, yInfo, proj, {pow}, chunkR, channel, opt, chunkC);
}

}  // namespace idw
}  // namespace v2r
