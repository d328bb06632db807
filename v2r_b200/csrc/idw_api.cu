// idw_api.cu -- host side of libv2r_idw.so: contexts, host helpers that define the inputs of the
// hot path, and the host-buffer solve entry points (single process driving 1..8 GPUs).
// See include/v2r_idw.h for the contract of every function.
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <nccl.h>  // types and prototypes only; the library is dlopen'ed when n_gpus > 1

#include "idw_internal.h"

namespace v2r {

// ------------------------------------------------------------------------------------------
// error reporting: thread-local message, integer status across the ABI
// ------------------------------------------------------------------------------------------
static thread_local char t_err[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
    return code;
}
const char *last_error() { return t_err; }

#define CUDA_TRY(expr)                                                                                \
    do {                                                                                              \
        cudaError_t ce__ = (expr);                                                                    \
        if (ce__ != cudaSuccess) {                                                                    \
            cudaGetLastError();                                                                       \
            return set_error(ce__ == cudaErrorMemoryAllocation ? V2R_IDW_ENOMEM : V2R_IDW_ECUDA,      \
                             "%s: %s", #expr, cudaGetErrorString(ce__));                              \
        }                                                                                             \
    } while (0)

// ------------------------------------------------------------------------------------------
// NCCL, loaded lazily (single-GPU contexts never touch it)
// ------------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return V2R_IDW_OK;
    const char *names[] = {getenv("V2R_IDW_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return set_error(V2R_IDW_ENCCL, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define LOAD(sym)                                                                      \
    g_nccl.sym = reinterpret_cast<decltype(g_nccl.sym)>(dlsym(h, "nccl" #sym));        \
    if (!g_nccl.sym) return set_error(V2R_IDW_ENCCL, "NCCL symbol nccl" #sym " missing")
    LOAD(CommInitAll);
    LOAD(CommDestroy);
    LOAD(GroupStart);
    LOAD(GroupEnd);
    LOAD(Broadcast);
    LOAD(Send);
    LOAD(Recv);
    LOAD(GetErrorString);
#undef LOAD
    g_nccl.handle = h;
    return V2R_IDW_OK;
}

#define NCCL_TRY(expr)                                                                            \
    do {                                                                                          \
        ncclResult_t nr__ = (expr);                                                               \
        if (nr__ != ncclSuccess)                                                                  \
            return set_error(V2R_IDW_ENCCL, "%s: %s", #expr, g_nccl.GetErrorString(nr__));        \
    } while (0)

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct DeviceState {
    int device = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    double *d_pts = nullptr;  // one block: [x | y | z | key_r | key_c], stride pts_stride elements
    double *d_x = nullptr, *d_y = nullptr, *d_z = nullptr;  // views into d_pts
    long long *d_kr = nullptr, *d_kc = nullptr;
    size_t pts_cap = 0, pts_stride = 0;  // elements per array
    double *d_band[2] = {nullptr, nullptr};
    size_t band_cap[2] = {0, 0};  // doubles
    cudaEvent_t k_stop[2] = {nullptr, nullptr};    // kernels of the round that used buffer b have finished
    cudaEvent_t drained[2] = {nullptr, nullptr};   // buffer b has been copied out / sent: it may be overwritten
    bool drained_valid[2] = {false, false};
    cudaEvent_t k_first = nullptr, k_last = nullptr, c_first = nullptr, c_last = nullptr, up_start = nullptr, up_stop = nullptr;
    cudaEvent_t h2d_done = nullptr;  // the caller's host arrays have been read completely
    bool k_any = false, c_any = false;
    ncclComm_t comm = nullptr;
    long long band_row0 = 0, band_rows = 0;  // rows owned in the current session
};

struct Session {
    bool active = false;
    v2r_idw_grid grid{};
    long long row0 = 0, nrows = 0;  // requested row range
    std::vector<double> exps;
    int precision = 0;
    long long sub_rows = 0;   // largest number of rows per device per round (a round = one kernel launch per device)
    std::vector<long long> round_off;  // round k covers rows [round_off[k], round_off[k+1]) of every device's band
    long long slice_rows = 0; // rows handed to the caller per _next call (<= sub_rows): the write granule
    long long slice_off = 0;  // rows of the current (round, device) band already delivered
    long long rounds = 0;
    long long cur_round = 0;
    long long launched = 0;   // rounds whose kernels are queued
    int cur_dev = 0;
    long long n = 0;
};

}  // namespace v2r

struct v2r_idw_ctx {
    std::mutex mu;
    std::vector<v2r::DeviceState> dev;
    double *d_land[2] = {nullptr, nullptr};  // on device 0: landing buffers for peer bands (NCCL gather path)
    size_t land_cap[2] = {0, 0};
    int land_next = 0;
    v2r::Session ses;
    double h2d_ms = 0, kernel_ms = 0, d2h_ms = 0;
    char err[512] = "";
};

namespace v2r {

static void free_device_state(DeviceState &d) {
    cudaSetDevice(d.device);
    if (d.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d.comm);
    for (void *p : {(void *)d.d_pts, (void *)d.d_band[0], (void *)d.d_band[1]})
        if (p) cudaFree(p);
    for (cudaEvent_t e : {d.k_stop[0], d.k_stop[1], d.drained[0], d.drained[1], d.k_first, d.k_last, d.c_first, d.c_last, d.up_start, d.up_stop, d.h2d_done})
        if (e) cudaEventDestroy(e);
    if (d.compute) cudaStreamDestroy(d.compute);
    if (d.copy) cudaStreamDestroy(d.copy);
}

static int init_device_state(DeviceState &d, int device) {
    d.device = device;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaStreamCreateWithFlags(&d.compute, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&d.k_stop[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&d.drained[i], cudaEventDisableTiming));
    }
    for (cudaEvent_t *e : {&d.k_first, &d.k_last, &d.c_first, &d.c_last, &d.up_start, &d.up_stop}) CUDA_TRY(cudaEventCreate(e));
    CUDA_TRY(cudaEventCreateWithFlags(&d.h2d_done, cudaEventDisableTiming));
    return V2R_IDW_OK;
}

static int create_ctx(const std::vector<int> &devices, v2r_idw_ctx **out) {
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(V2R_IDW_ECUDA, "no CUDA device available (%s); this library has no CPU fallback",
                         ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
    }
    for (int dv : devices)
        if (dv < 0 || dv >= count) return set_error(V2R_IDW_EINVAL, "device %d not present (%d visible)", dv, count);
    auto *ctx = new (std::nothrow) v2r_idw_ctx();
    if (!ctx) return set_error(V2R_IDW_ENOMEM, "out of host memory");
    ctx->dev.resize(devices.size());
    for (size_t i = 0; i < devices.size(); ++i) {
        int rc = init_device_state(ctx->dev[i], devices[i]);
        if (rc) {
            for (auto &d : ctx->dev) free_device_state(d);
            delete ctx;
            return rc;
        }
    }
    if (devices.size() > 1) {
        int rc = load_nccl();
        if (rc == V2R_IDW_OK) {
            std::vector<ncclComm_t> comms(devices.size());
            ncclResult_t nr = g_nccl.CommInitAll(comms.data(), (int)devices.size(), devices.data());
            if (nr != ncclSuccess) rc = set_error(V2R_IDW_ENCCL, "ncclCommInitAll: %s", g_nccl.GetErrorString(nr));
            else
                for (size_t i = 0; i < devices.size(); ++i) ctx->dev[i].comm = comms[i];
        }
        if (rc) {
            for (auto &d : ctx->dev) free_device_state(d);
            delete ctx;
            return rc;
        }
    }
    *out = ctx;
    return V2R_IDW_OK;
}

template <typename T>
static int ensure(T *&p, size_t &cap, size_t need) {
    if (need <= cap && p) return V2R_IDW_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t alloc = need + (need >> 3) + 64;
    cudaError_t ce = cudaMalloc((void **)&p, alloc * sizeof(T));
    if (ce != cudaSuccess) {
        cudaGetLastError();
        return set_error(V2R_IDW_ENOMEM, "cudaMalloc(%zu bytes): %s", alloc * sizeof(T), cudaGetErrorString(ce));
    }
    cap = alloc;
    return V2R_IDW_OK;
}

// The five point arrays live in ONE device block [x | y | z | key_r | key_c] with a common stride (all 8-byte
// elements), so replicating the point set is literally one NCCL broadcast.
static int ensure_points(DeviceState &d, size_t n) {
    const size_t stride = ((std::max<size_t>(n, 1) + 1) / 2) * 2;  // even: every array stays 16-byte aligned (TMA)
    if (stride > d.pts_cap || !d.d_pts) {
        CUDA_TRY(cudaSetDevice(d.device));
        if (d.d_pts) cudaFree(d.d_pts);
        d.d_pts = nullptr;
        d.pts_cap = 0;
        const size_t cap = ((stride + (stride >> 3) + 1023) / 1024) * 1024;
        CUDA_TRY(cudaMalloc((void **)&d.d_pts, cap * 5 * 8));
        d.pts_cap = cap;
    }
    d.pts_stride = stride;
    d.d_x = d.d_pts;
    d.d_y = d.d_pts + stride;
    d.d_z = d.d_pts + 2 * stride;
    d.d_kr = reinterpret_cast<long long *>(d.d_pts + 3 * stride);
    d.d_kc = reinterpret_cast<long long *>(d.d_pts + 4 * stride);
    return V2R_IDW_OK;
}

// Upload the points to device 0 and replicate them to every other device with ONE NCCL broadcast over NVLink
// (north_star: "the point set is replicated with one NCCL broadcast"): 5 * stride * 8 bytes, c4: 39 MB.
// Returns only after the host arrays have been read completely (include/v2r_idw.h: no caller pointer is kept).
static int upload_points(v2r_idw_ctx *ctx, const double *x, const double *y, const double *z, const int64_t *kr,
                         const int64_t *kc, long long n) {
    for (auto &d : ctx->dev) {
        int rc = ensure_points(d, (size_t)std::max<long long>(n, 1));
        if (rc) return rc;
    }
    DeviceState &d0 = ctx->dev[0];
    CUDA_TRY(cudaSetDevice(d0.device));
    CUDA_TRY(cudaEventRecord(d0.up_start, d0.compute));
    if (n > 0) {
        CUDA_TRY(cudaMemcpyAsync(d0.d_x, x, n * 8, cudaMemcpyHostToDevice, d0.compute));
        CUDA_TRY(cudaMemcpyAsync(d0.d_y, y, n * 8, cudaMemcpyHostToDevice, d0.compute));
        CUDA_TRY(cudaMemcpyAsync(d0.d_z, z, n * 8, cudaMemcpyHostToDevice, d0.compute));
        CUDA_TRY(cudaMemcpyAsync(d0.d_kr, kr, n * 8, cudaMemcpyHostToDevice, d0.compute));
        CUDA_TRY(cudaMemcpyAsync(d0.d_kc, kc, n * 8, cudaMemcpyHostToDevice, d0.compute));
    }
    CUDA_TRY(cudaEventRecord(d0.h2d_done, d0.compute));
    if (ctx->dev.size() > 1 && n > 0) {
        NCCL_TRY(g_nccl.GroupStart());  // one collective; the group only spans this process's communicators
        for (auto &d : ctx->dev)
            NCCL_TRY(g_nccl.Broadcast(d.d_pts, d.d_pts, 5 * d.pts_stride, ncclDouble, 0, d.comm, d.compute));
        NCCL_TRY(g_nccl.GroupEnd());
    }
    CUDA_TRY(cudaSetDevice(d0.device));
    CUDA_TRY(cudaEventRecord(d0.up_stop, d0.compute));
    CUDA_TRY(cudaEventSynchronize(d0.h2d_done));  // pinned host arrays make the copies truly asynchronous: wait for the DMA
    return V2R_IDW_OK;
}

static long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

// rows of device g's sub-band in round k, clipped to its band
static void sub_band(const Session &s, const DeviceState &d, long long k, long long *r0, long long *nr) {
    long long a = d.band_row0 + std::min(s.round_off[k], d.band_rows);
    long long e = d.band_row0 + std::min(s.round_off[k + 1], d.band_rows);
    *r0 = a;
    *nr = e > a ? e - a : 0;
}

// Queues the kernels of round k on every device.  Buffer k & 1 is reused every second round: the compute stream
// waits (on the device, not the host) until the previous contents have been copied out.
static int launch_round(v2r_idw_ctx *ctx, long long k) {
    Session &s = ctx->ses;
    const int buf = (int)(k & 1);
    for (auto &d : ctx->dev) {
        long long r0, nr;
        sub_band(s, d, k, &r0, &nr);
        if (nr <= 0) continue;
        CUDA_TRY(cudaSetDevice(d.device));
        if (d.drained_valid[buf]) CUDA_TRY(cudaStreamWaitEvent(d.compute, d.drained[buf], 0));
        if (!d.k_any) { CUDA_TRY(cudaEventRecord(d.k_first, d.compute)); d.k_any = true; }
        int rc;
        if (s.precision == V2R_IDW_FP32)
            rc = launch_rows_fp32(d.device, d.compute, d.d_x, d.d_y, d.d_z, d.d_kr, d.d_kc, s.n, s.grid, r0, nr, s.exps.data(), (int)s.exps.size(), d.d_band[buf]);
        else
            rc = launch_rows_fp64(d.device, d.compute, d.d_x, d.d_y, d.d_z, d.d_kr, d.d_kc, s.n, s.grid, r0, nr, s.exps.data(), (int)s.exps.size(), d.d_band[buf]);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(d.k_stop[buf], d.compute));
        CUDA_TRY(cudaEventRecord(d.k_last, d.compute));
    }
    s.launched = k + 1;
    return V2R_IDW_OK;
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
static bool force_nccl_gather() {
    static const bool v = getenv("V2R_IDW_GATHER") && !strcmp(getenv("V2R_IDW_GATHER"), "nccl");
    return v;
}

static int begin_session(v2r_idw_ctx *ctx, const double *x, const double *y, const double *z, const int64_t *kr,
                         const int64_t *kc, long long n, const v2r_idw_grid *grid, long long row0, long long nrows,
                         const double *exps, int n_exps, int precision, long long band_rows) {
    if (ctx->ses.active) return set_error(V2R_IDW_ESTATE, "a streaming session is already open on this context");
    int rc = validate_geometry(n, grid, row0, nrows, n_exps);
    if (rc) return rc;
    if (n > 0 && (!x || !y || !z || !kr || !kc)) return set_error(V2R_IDW_EINVAL, "point or key arrays are NULL");
    if (precision != V2R_IDW_FP64 && precision != V2R_IDW_FP32) return set_error(V2R_IDW_EINVAL, "unknown precision %d", precision);
    std::vector<Pass> passes;
    if ((rc = plan_sweep(exps, n_exps, passes))) return rc;

    Session &s = ctx->ses;
    s = Session{};
    s.grid = *grid;
    s.row0 = row0;
    s.nrows = nrows;
    s.exps.assign(exps, exps + n_exps);
    s.precision = precision;
    s.n = n;
    ctx->h2d_ms = ctx->kernel_ms = ctx->d2h_ms = 0;
    ctx->land_next = 0;

    const long long G = (long long)ctx->dev.size();
    // contiguous row band per device, multiple of 32 rows so CTA tiles stay whole
    long long per = ceil_div(nrows, G);
    per = ceil_div(per, 32) * 32;
    for (long long g = 0; g < G; ++g) {
        DeviceState &d = ctx->dev[g];
        long long a = std::min(row0 + g * per, row0 + nrows), e = std::min(a + per, row0 + nrows);
        d.band_row0 = a;
        d.band_rows = e - a;
        d.k_any = d.c_any = false;
        d.drained_valid[0] = d.drained_valid[1] = false;
    }
    // Launch granule (rows per device per round).  Cutting a device's band into R rounds lets the copy-out of round k
    // overlap the kernels of round k+1, so only bytes/R of copy is exposed at the end -- but every launch ends with a
    // tail of about 1/16 of a tile time (idw_kernels.cuh), and a tile costs 4096 cells x n points.  R minimises
    // R * t_tail + bytes / (R * bw): c2 (100 k points, 134 MB per device) -> 2 rounds; the c4 slab (967 k points:
    // 14 ms of tail per launch against 13 ms of copy) -> 1 round.  Each round holds at most ~1 GiB of output.  When the
    // caller asks for small bands (the reference's default chunk is 200 rows) the granule is a multiple of the
    // requested height.
    const long long cols1 = std::max<long long>(1, grid->cols);
    const long long row_bytes = cols1 * 8 * n_exps;
    const long long cap_rows = std::max<long long>(32, ((1LL << 30) / row_bytes) / 32 * 32);
    const double t_tail = 4096.0 * (double)std::max<long long>(n, 1) * 148.0 / estimate_pair_rate(exps, n_exps, precision) / 16.0;
    const double t_copy = (double)per * (double)row_bytes / 20e9;  // one device's band over its PCIe link
    long long want_rounds = (long long)std::llround(std::sqrt(t_copy / std::max(t_tail, 1e-6)));
    want_rounds = std::max<long long>(1, std::min<long long>(want_rounds, 8));
    // Rounds: want_rounds - 1 equal granules followed by a SHORT last round (about a quarter of a granule): the last
    // round's copy is the only one that no kernel overlaps.  Granules are multiples of `unit` rows (32 = whole CTA tiles,
    // or the caller's band height).
    const long long unit = band_rows > 0 ? band_rows : 32;
    auto round_up = [&](long long v) { return std::max(unit, ceil_div(v, unit) * unit); };
    const long long cap_units = std::max(unit, cap_rows / unit * unit);
    s.round_off.clear();
    if (per > 0) {
        long long last = 0;
        if (want_rounds >= 2) {
            last = round_up(per / (4 * want_rounds));
            if (last >= per) last = 0;
        }
        const long long body = per - last;
        const long long granule = std::min(cap_units, round_up(ceil_div(body, std::max<long long>(1, want_rounds - (last ? 1 : 0)))));
        for (long long a0 = 0; a0 < body; a0 += granule) s.round_off.push_back(a0);
        if (last) s.round_off.push_back(body);
        s.round_off.push_back(per);
    }
    s.rounds = s.round_off.empty() ? 0 : (long long)s.round_off.size() - 1;
    s.sub_rows = 1;
    for (long long k = 0; k < s.rounds; ++k) s.sub_rows = std::max(s.sub_rows, s.round_off[k + 1] - s.round_off[k]);
    s.slice_rows = band_rows > 0 ? std::min<long long>(band_rows, s.sub_rows) : s.sub_rows;
    s.slice_off = 0;

    const size_t band_doubles = (size_t)s.sub_rows * (size_t)cols1 * (size_t)n_exps;
    for (auto &d : ctx->dev) {
        CUDA_TRY(cudaSetDevice(d.device));
        for (int i = 0; i < 2; ++i)
            if ((rc = ensure(d.d_band[i], d.band_cap[i], band_doubles))) return rc;
    }
    if ((rc = upload_points(ctx, x, y, z, kr, kc, n))) return rc;
    if (s.rounds > 0 && nrows > 0 && grid->cols > 0)
        if ((rc = launch_round(ctx, 0))) return rc;
    s.active = true;
    return V2R_IDW_OK;
}

// Queues the copy-out of rows [off, off + snr) of device g's band of round k (exponent planes at
// dst + e * dst_stride).  Direct path: the device copies from its own band buffer (one PCIe link per GPU; used when
// `dst` is page-locked, and always for device 0 and single-GPU contexts).  Gather path (north_star's shape, taken when
// `dst` is pageable or V2R_IDW_GATHER=nccl): the band travels to device 0 over NVLink (ncclSend / ncclRecv into one
// of two landing buffers) once per band, and is copied out from there.  Returns the stream the copy was queued on.
static int queue_copy_out(v2r_idw_ctx *ctx, int g, long long k, long long off, long long snr, long long nr, double *dst,
                          long long dst_stride, bool direct, bool first_slice, bool last_slice, cudaStream_t *queued_on) {
    Session &s = ctx->ses;
    DeviceState &d = ctx->dev[g];
    DeviceState &d0 = ctx->dev[0];
    const int buf = (int)(k & 1);
    const int E = (int)s.exps.size();
    const long long plane_src = nr * s.grid.cols, plane_dst = snr * s.grid.cols;
    const double *src = d.d_band[buf];
    DeviceState *cd = &d;  // device whose copy stream does the D2H
    if (g != 0 && !direct) {
        if (first_slice) {
            const size_t need = (size_t)plane_src * E;
            const int lb = ctx->land_next;
            CUDA_TRY(cudaSetDevice(d0.device));
            int rc = ensure(ctx->d_land[lb], ctx->land_cap[lb], need);  // (a reallocation synchronises the device: first use only)
            if (rc) return rc;
            CUDA_TRY(cudaSetDevice(d.device));
            CUDA_TRY(cudaStreamWaitEvent(d.copy, d.k_stop[buf], 0));
            NCCL_TRY(g_nccl.GroupStart());
            NCCL_TRY(g_nccl.Send(d.d_band[buf], need, ncclDouble, 0, d.comm, d.copy));
            NCCL_TRY(g_nccl.Recv(ctx->d_land[lb], need, ncclDouble, g, d0.comm, d0.copy));
            NCCL_TRY(g_nccl.GroupEnd());
            CUDA_TRY(cudaEventRecord(d.drained[buf], d.copy));  // the band has left this device
            d.drained_valid[buf] = true;
            ctx->land_next ^= 1;
        }
        src = ctx->d_land[ctx->land_next ^ 1];
        cd = &d0;
    }
    CUDA_TRY(cudaSetDevice(cd->device));
    if (cd == &d && first_slice) CUDA_TRY(cudaStreamWaitEvent(d.copy, d.k_stop[buf], 0));
    if (!cd->c_any) { CUDA_TRY(cudaEventRecord(cd->c_first, cd->copy)); cd->c_any = true; }
    for (int e = 0; e < E; ++e)
        CUDA_TRY(cudaMemcpyAsync(dst + (long long)e * dst_stride, src + (long long)e * plane_src + off * s.grid.cols,
                                 (size_t)plane_dst * 8, cudaMemcpyDeviceToHost, cd->copy));
    CUDA_TRY(cudaEventRecord(cd->c_last, cd->copy));
    if (cd == &d && last_slice) {
        CUDA_TRY(cudaEventRecord(d.drained[buf], d.copy));
        d.drained_valid[buf] = true;
    }
    *queued_on = cd->copy;
    return V2R_IDW_OK;
}

// The whole row range into one host buffer: every round of every device is queued up front -- kernels, the
// device-side waits for buffer reuse, the copies -- and the host only waits at the end (end_session).
static int run_all(v2r_idw_ctx *ctx, double *out, long long plane_stride) {
    Session &s = ctx->ses;
    if (s.grid.cols == 0) return V2R_IDW_OK;
    const int G = (int)ctx->dev.size();
    const bool direct = G == 1 || (is_pinned(out) && !force_nccl_gather());
    for (long long k = 0; k < s.rounds; ++k) {
        if (k + 1 < s.rounds && s.launched < k + 2) {
            int rc = launch_round(ctx, k + 1);  // waits on `drained` of round k-1, queued during iteration k-1
            if (rc) return rc;
        }
        for (int g = 0; g < G; ++g) {
            long long r0, nr;
            sub_band(s, ctx->dev[g], k, &r0, &nr);
            if (nr <= 0) continue;
            cudaStream_t q;
            int rc = queue_copy_out(ctx, g, k, 0, nr, nr, out + (r0 - s.row0) * s.grid.cols, plane_stride, direct, true, true, &q);
            if (rc) return rc;
        }
    }
    s.cur_round = s.rounds;
    return V2R_IDW_OK;
}

// Delivers the next slice (<= slice_rows rows) of the next (round, device) band into out (planes of the slice's
// own height, packed).
static int next_band(v2r_idw_ctx *ctx, double *out, long long *o_row0, long long *o_nrows) {
    Session &s = ctx->ses;
    if (!s.active) return set_error(V2R_IDW_ESTATE, "no streaming session is open");
    const int G = (int)ctx->dev.size();
    for (;;) {
        if (s.cur_round >= s.rounds || s.grid.cols == 0) {
            *o_row0 = s.row0 + s.nrows;
            *o_nrows = 0;
            return V2R_IDW_OK;
        }
        const bool band_start = s.slice_off == 0;
        DeviceState &d = ctx->dev[s.cur_dev];
        const int g = s.cur_dev;
        long long r0, nr;
        sub_band(s, d, s.cur_round, &r0, &nr);
        if (nr <= 0) {
            if (++s.cur_dev == G) { s.cur_dev = 0; ++s.cur_round; }
            continue;
        }
        const long long sr0 = r0 + s.slice_off, snr = std::min(s.slice_rows, nr - s.slice_off);
        const bool last_slice = s.slice_off + snr >= nr;
        const bool direct = G == 1 || (is_pinned(out) && !force_nccl_gather());
        cudaStream_t q;
        int rc = queue_copy_out(ctx, g, s.cur_round, s.slice_off, snr, nr, out, snr * s.grid.cols, direct, band_start, last_slice, &q);
        if (rc) return rc;
        // the next round computes while this one is copied out: its kernels wait (on the device) for the buffers
        // of round cur_round - 1, all of which were drained by earlier calls
        if (band_start && g == 0 && s.cur_round + 1 < s.rounds && s.launched < s.cur_round + 2)
            if ((rc = launch_round(ctx, s.cur_round + 1))) return rc;
        CUDA_TRY(cudaStreamSynchronize(q));
        s.slice_off += snr;
        if (last_slice) {
            s.slice_off = 0;
            if (++s.cur_dev == G) { s.cur_dev = 0; ++s.cur_round; }
        }
        *o_row0 = sr0;
        *o_nrows = snr;
        return V2R_IDW_OK;
    }
}

static int end_session(v2r_idw_ctx *ctx) {
    Session &s = ctx->ses;
    if (!s.active) return set_error(V2R_IDW_ESTATE, "no streaming session is open");
    int rc = V2R_IDW_OK;
    for (auto &d : ctx->dev) {
        cudaSetDevice(d.device);
        cudaError_t ce = cudaStreamSynchronize(d.compute);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(d.copy);
        if (ce != cudaSuccess && rc == V2R_IDW_OK) rc = set_error(V2R_IDW_ECUDA, "stream sync: %s", cudaGetErrorString(ce));
        float ms = 0;
        // span first kernel -> last kernel on this device (includes waits for a free band buffer, if any)
        if (d.k_any && cudaEventElapsedTime(&ms, d.k_first, d.k_last) == cudaSuccess) ctx->kernel_ms = std::max(ctx->kernel_ms, (double)ms);
        // copy-out that was NOT hidden behind kernels: end of this device's last kernel -> end of its last copy
        // (copies of earlier rounds overlap the kernels of later ones)
        if (d.k_any && d.c_any && cudaEventElapsedTime(&ms, d.k_last, d.c_last) == cudaSuccess) ctx->d2h_ms = std::max(ctx->d2h_ms, (double)std::max(ms, 0.0f));
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->dev[0].up_start, ctx->dev[0].up_stop) == cudaSuccess) ctx->h2d_ms = ms;
    cudaGetLastError();
    s.active = false;
    return rc;
}

// keeps the message of a failed call with the context, so a caller that fetches it from another OS thread
// (goroutines migrate between two cgo calls) still gets the right text
static int remember(v2r_idw_ctx *ctx, int rc) {
    if (rc != V2R_IDW_OK && ctx) {
        strncpy(ctx->err, last_error(), sizeof ctx->err - 1);
        ctx->err[sizeof ctx->err - 1] = 0;
    }
    return rc;
}

}  // namespace v2r

using namespace v2r;

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int v2r_idw_abi_version(void) { return V2R_IDW_ABI_VERSION; }
const char *v2r_idw_last_error(void) { return last_error(); }

int v2r_idw_device_count(int *count) {
    if (!count) return set_error(V2R_IDW_EINVAL, "count is NULL");
    cudaError_t ce = cudaGetDeviceCount(count);
    if (ce != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return set_error(V2R_IDW_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(ce));
    }
    return V2R_IDW_OK;
}

int v2r_idw_init(int n_gpus, v2r_idw_ctx **out) {
    if (!out) return set_error(V2R_IDW_EINVAL, "out is NULL");
    *out = nullptr;
    if (n_gpus != 1 && n_gpus != 2 && n_gpus != 4 && n_gpus != 8) return set_error(V2R_IDW_EINVAL, "n_gpus must be 1, 2, 4 or 8 (got %d)", n_gpus);
    return guarded([&] {
        std::vector<int> devs(n_gpus);
        for (int i = 0; i < n_gpus; ++i) devs[i] = i;
        return create_ctx(devs, out);
    });
}

int v2r_idw_init_device(int device, v2r_idw_ctx **out) {
    if (!out) return set_error(V2R_IDW_EINVAL, "out is NULL");
    *out = nullptr;
    return guarded([&] { return create_ctx({device}, out); });
}

void v2r_idw_destroy(v2r_idw_ctx *ctx) {
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        cudaSetDevice(ctx->dev[0].device);
        for (double *p : ctx->d_land)
            if (p) cudaFree(p);
        for (auto &d : ctx->dev) free_device_state(d);
    }
    delete ctx;
}

int v2r_idw_alloc_pinned(size_t bytes, void **ptr) {
    if (!ptr) return set_error(V2R_IDW_EINVAL, "ptr is NULL");
    *ptr = nullptr;
    cudaError_t ce = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        return set_error(ce == cudaErrorMemoryAllocation ? V2R_IDW_ENOMEM : V2R_IDW_ECUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(ce));
    }
    return V2R_IDW_OK;
}

int v2r_idw_free_pinned(void *ptr) {
    if (!ptr) return V2R_IDW_OK;
    cudaError_t ce = cudaFreeHost(ptr);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        return set_error(V2R_IDW_ECUDA, "cudaFreeHost: %s", cudaGetErrorString(ce));
    }
    return V2R_IDW_OK;
}

// ---- host helpers ---------------------------------------------------------------------------
// Go's float64 -> int conversion on amd64: truncation; NaN / out of range -> INT64_MIN.
static int64_t go_int(double v) {
    if (!(v > -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)v;
}

int v2r_idw_get_dimensions(double x_min, double x_max, double x_step, double y_min, double y_max, double y_step,
                           int64_t *rows, int64_t *cols) {
    if (!rows || !cols) return set_error(V2R_IDW_EINVAL, "rows/cols is NULL");
    *cols = go_int((1 + x_max - x_min) / x_step);
    *rows = go_int((1 + y_max - y_min) / y_step);
    return V2R_IDW_OK;
}

int v2r_idw_point_to_pair(double x, double y, double x_min, double x_step, double y_min, double y_step, int64_t *r,
                          int64_t *c) {
    if (!r || !c) return set_error(V2R_IDW_EINVAL, "r/c is NULL");
    *r = go_int((y - y_min) / y_step);
    *c = go_int((x - x_min) / x_step);
    return V2R_IDW_OK;
}

int v2r_idw_exp_range(double es, double ee, double ei, double *out, int64_t max_out, int64_t *count) {
    if (!count) return set_error(V2R_IDW_EINVAL, "count is NULL");
    *count = 0;
    if (!(ei > 0)) return set_error(V2R_IDW_EINVAL, "exponential increment (%g) must be > 0", ei);
    if (!(ee > es)) return set_error(V2R_IDW_EINVAL, "exponent range: [%g, %g) has no iterations", es, ee);
    int64_t n = 0;
    for (double e = es; e < ee; e += ei) {
        if (out && n < max_out) out[n] = e;
        if (++n > (1 << 20)) return set_error(V2R_IDW_EINVAL, "exponent range has more than 2^20 steps");
    }
    *count = n;
    return V2R_IDW_OK;
}

// tools/utils.go:97-119 with the arrival order fixed to input order.  The Go map becomes a flat
// open-addressing table (linear probing, load <= 1/2) from key to output slot: 2 M points take ~0.1 s.
int v2r_idw_make_coord_space(int64_t n, const double *x, const double *y, const double *z, double x_min,
                             double x_step, double y_min, double y_step, double *ox, double *oy, double *oz,
                             int64_t *okr, int64_t *okc, int64_t *n_out) {
    if (!n_out) return set_error(V2R_IDW_EINVAL, "n_out is NULL");
    *n_out = 0;
    if (n < 0) return set_error(V2R_IDW_EINVAL, "n < 0");
    if (n > 0 && (!x || !y || !z || !ox || !oy || !oz || !okr || !okc)) return set_error(V2R_IDW_EINVAL, "NULL array");
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2 + 1) cap <<= 1;
    std::vector<int64_t> slot;
    try {
        slot.assign(cap, -1);
    } catch (const std::bad_alloc &) {
        return set_error(V2R_IDW_ENOMEM, "out of host memory in make_coord_space");
    }
    const uint64_t mask = cap - 1;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = go_int((y[i] - y_min) / y_step), c = go_int((x[i] - x_min) / x_step);
        uint64_t h = (uint64_t)r * 0x9E3779B97F4A7C15ull + (uint64_t)c * 0xC2B2AE3D27D4EB4Full;
        h ^= h >> 29;
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 32;
        uint64_t j = h & mask;
        while (slot[j] >= 0 && !(okr[slot[j]] == r && okc[slot[j]] == c)) j = (j + 1) & mask;
        double w = z[i];
        int64_t s = slot[j];
        if (s >= 0) {
            w = (w + oz[s]) / 2;  // tools/utils.go:111: newElev := (p.Weight + elev.Weight) / 2
        } else {
            s = slot[j] = cnt++;
            okr[s] = r;
            okc[s] = c;
        }
        ox[s] = x[i];  // position of the last arrival is kept (tools/utils.go:115)
        oy[s] = y[i];
        oz[s] = w;
    }
    *n_out = cnt;
    return V2R_IDW_OK;
}

// ---- solves -----------------------------------------------------------------------------------
int v2r_idw_solve_rows(v2r_idw_ctx *ctx, const double *x, const double *y, const double *z, const int64_t *key_r,
                       const int64_t *key_c, int64_t n, const v2r_idw_grid *grid, int64_t row0, int64_t nrows,
                       const double *exps, int n_exps, int precision, double *out) {
    if (!ctx) return set_error(V2R_IDW_EINVAL, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (grid && nrows > 0 && grid->cols > 0 && !out) return set_error(V2R_IDW_EINVAL, "out is NULL");
    return remember(ctx, guarded([&] {
        int rc = begin_session(ctx, x, y, z, key_r, key_c, n, grid, row0, nrows, exps, n_exps, precision, 0);
        if (rc) return rc;
        rc = run_all(ctx, out, (long long)nrows * grid->cols);
        int rc2 = end_session(ctx);
        return rc ? rc : rc2;
    }));
}

int v2r_idw_solve(v2r_idw_ctx *ctx, const double *x, const double *y, const double *z, const int64_t *key_r,
                  const int64_t *key_c, int64_t n, const v2r_idw_grid *grid, const double *exps, int n_exps,
                  int precision, double *out) {
    if (!grid) return set_error(V2R_IDW_EINVAL, "grid is NULL");
    return v2r_idw_solve_rows(ctx, x, y, z, key_r, key_c, n, grid, 0, grid->rows, exps, n_exps, precision, out);
}

int v2r_idw_stream_begin(v2r_idw_ctx *ctx, const double *x, const double *y, const double *z, const int64_t *key_r,
                         const int64_t *key_c, int64_t n, const v2r_idw_grid *grid, const double *exps, int n_exps,
                         int precision, int64_t band_rows) {
    if (!ctx) return set_error(V2R_IDW_EINVAL, "ctx is NULL");
    if (!grid) return set_error(V2R_IDW_EINVAL, "grid is NULL");
    if (band_rows < 0) return set_error(V2R_IDW_EINVAL, "band_rows < 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return remember(ctx, guarded([&] { return begin_session(ctx, x, y, z, key_r, key_c, n, grid, 0, grid->rows, exps, n_exps, precision, band_rows); }));
}

int v2r_idw_stream_next(v2r_idw_ctx *ctx, double *out, int64_t *row0, int64_t *nrows) {
    if (!ctx || !row0 || !nrows) return set_error(V2R_IDW_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!out && ctx->ses.active && ctx->ses.cur_round < ctx->ses.rounds && ctx->ses.grid.cols > 0)
        return set_error(V2R_IDW_EINVAL, "out is NULL");
    long long r0 = 0, nr = 0;
    int rc = remember(ctx, guarded([&] { return next_band(ctx, out, &r0, &nr); }));
    *row0 = r0;
    *nrows = nr;
    return rc;
}

int v2r_idw_stream_end(v2r_idw_ctx *ctx) {
    if (!ctx) return set_error(V2R_IDW_EINVAL, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return remember(ctx, end_session(ctx));
}

const char *v2r_idw_ctx_last_error(v2r_idw_ctx *ctx) {
    if (!ctx) return "";
    std::lock_guard<std::mutex> lk(ctx->mu);
    return ctx->err;
}

int v2r_idw_last_timing(v2r_idw_ctx *ctx, double *h2d_ms, double *kernel_ms, double *d2h_ms) {
    if (!ctx) return set_error(V2R_IDW_EINVAL, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (h2d_ms) *h2d_ms = ctx->h2d_ms;
    if (kernel_ms) *kernel_ms = ctx->kernel_ms;
    if (d2h_ms) *d2h_ms = ctx->d2h_ms;
    return V2R_IDW_OK;
}

}  // extern "C"
