// idw_launch.cu -- exponent classification, kernel dispatch and the device-level C-ABI entry
// (v2r_idw_launch_rows, v2r_idw_describe, v2r_idw_measure_peaks).  See include/v2r_idw.h.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "idw_internal.h"
#include "idw_kernels.cuh"

namespace v2r {

static const double kLgHost[3] = {1.4426950408889634074, -0.72134758493810647712, 0.48089839855788806054};  // = kLg (idw_math.cuh)

// ------------------------------------------------------------------------------------------
// Sweep classification.  The host accumulates exponents in FP64
// (cmd/idw.go:153), so class membership is decided on the exact double that arrives.
// ------------------------------------------------------------------------------------------
static bool is_int(double v) { return std::floor(v) == v; }

int plan_sweep(const double *exps, int n_exps, std::vector<Pass> &passes) {
    passes.clear();
    if (!exps || n_exps < 1 || n_exps > kMaxSweep) return set_error(V2R_IDW_EINVAL, "n_exps must be in [1, %d]", kMaxSweep);
    for (int e = 0; e < n_exps; ++e)
        if (!std::isfinite(exps[e])) return set_error(V2R_IDW_EINVAL, "exponent %d is not finite", e);

    bool half = true, whole = true, even = true;
    for (int e = 0; e < n_exps; ++e) {
        const double p = exps[e];
        if (!(p > 0.0) || !is_int(2.0 * p) || p > 32.0) half = false;
        if (!is_int(p)) whole = false;
        if (!is_int(p / 2.0)) even = false;
    }
    if (!half) {
        constexpr int kMaxGenE = 9;  // the 10-exponent general build spills registers: 10 exponents run as 5 + 5
        for (int e0 = 0; e0 < n_exps;) {
            const int left = n_exps - e0;
            const int chunks = (left + kMaxGenE - 1) / kMaxGenE;
            const int len = (left + chunks - 1) / chunks;
            Pass ps{};
            ps.mode = MODE_GEN;
            ps.first = e0;
            ps.count = len;
            ps.k0 = GEN_SINGLE;  // every exponent in (0, 1.99]; else GEN_SAFE (0, 3.75]; else GEN_ANY (see GenKind)
            for (int i = 0; i < len; ++i) {
                const double p = exps[e0 + i];
                ps.cexp[i] = -p / 2.0;
                if (!(p > 0.0 && p <= 3.75)) ps.k0 = GEN_ANY;
                else if (p > 1.99 && ps.k0 == GEN_SINGLE) ps.k0 = GEN_SAFE;
            }
            passes.push_back(ps);
            e0 += len;
        }
        return V2R_IDW_OK;
    }
    const int mode = even ? MODE_RCP : (whole ? MODE_RSQ : MODE_QRT);
    const double unit = even ? 2.0 : (whole ? 1.0 : 0.5);
    std::vector<int> k(n_exps);
    for (int e = 0; e < n_exps; ++e) k[e] = (int)std::llround(exps[e] / unit);
    bool ap = true;
    const int dk = n_exps > 1 ? k[1] - k[0] : 0;
    if (n_exps > 1 && dk < 1) ap = false;
    for (int e = 2; e < n_exps && ap; ++e)
        if (k[e] - k[e - 1] != dk) ap = false;
    if (ap) {
        for (int e0 = 0; e0 < n_exps;) {
            const int left = n_exps - e0;
            const int chunks = (left + kMaxE - 1) / kMaxE;
            const int len = (left + chunks - 1) / chunks;
            Pass ps{};
            ps.mode = mode;
            ps.first = e0;
            ps.count = len;
            ps.k0 = k[e0];
            ps.dk = len > 1 ? dk : 0;
            passes.push_back(ps);
            e0 += len;
        }
    } else {  // arbitrary list: one pass per exponent, each in its own best class
        for (int e = 0; e < n_exps; ++e) {
            std::vector<Pass> one;
            int rc = plan_sweep(exps + e, 1, one);
            if (rc) return rc;
            one[0].first = e;
            passes.push_back(one[0]);
        }
    }
    return V2R_IDW_OK;
}

// ------------------------------------------------------------------------------------------
// Kernel table
// ------------------------------------------------------------------------------------------
struct Variant {
    int mode, E, k0, dk;  // k0 == 0: runtime ladder (general class: k0 = GenKind)
    int cr, cc, threads;
    void (*fn)(const KParams);
    const char *name;
    bool paired;
    const char *tag;  // non-null: tuning build, selected only by V2R_IDW_TUNE=<tag>
    int smem;         // dynamic shared memory per CTA
    int opt;          // bit 0 = paired reciprocal, bit 1 = dy^2 staged in shared memory
};

// OPT of the kernels: bit 0 = paired reciprocal, bit 1 = dy^2 staged in shared memory (idw_kernels.cuh)
#define V2R_KERNEL(MODE, E, CR, CC, K0, DK, THREADS, MINB, OPT, PAIRED, TAG)                                              \
    Variant { MODE, E, K0, DK, CR, CC, THREADS, idw_fp64_kernel<MODE, E, CR, CC, K0, DK, THREADS, MINB, OPT>, #MODE, PAIRED, TAG, \
              smem_bytes(Fp64BodyFor<MODE, E, CR, CC, K0, DK, THREADS, MINB, OPT>::BODY_BYTES, 2 * (E) * (CR) * (CC), THREADS), OPT }
#define V2R_VARIANT(MODE, E, CR, CC, K0, DK, THREADS, MINB) V2R_KERNEL(MODE, E, CR, CC, K0, DK, THREADS, MINB, kStage, false, nullptr)
#define V2R_TUNE(TAG, MODE, E, CR, CC, K0, DK, THREADS, MINB) V2R_KERNEL(MODE, E, CR, CC, K0, DK, THREADS, MINB, kStage, false, TAG)
#define V2R_UNSTAGED(TAG, MODE, E, CR, CC, K0, DK, THREADS, MINB) V2R_KERNEL(MODE, E, CR, CC, K0, DK, THREADS, MINB, 0, false, TAG)
#define V2R_PAIRED(TAG, CR, CC, THREADS, MINB) V2R_KERNEL(MODE_RCP, 1, CR, CC, 1, 0, THREADS, MINB, 1, true, TAG)
#define V2R_PAIRED_YS(TAG, CR, CC, THREADS, MINB) V2R_KERNEL(MODE_RCP, 1, CR, CC, 1, 0, THREADS, MINB, 3, true, TAG)
constexpr int kStage = 2;

// Cell block per thread shrinks as E grows (2*E*CR*CC FP64 accumulators live in registers).
// (the single-exponent runtime-ladder build stays unstaged: with the staging pointers it spills at 128 registers)
#define V2R_ROW(MODE)                                                                                  \
    V2R_UNSTAGED(nullptr, MODE, 1, 2, 4, 0, 0, 256, 2), V2R_VARIANT(MODE, 2, 2, 4, 0, 0, 256, 1),      \
        V2R_VARIANT(MODE, 3, 2, 2, 0, 0, 256, 2), V2R_VARIANT(MODE, 4, 2, 2, 0, 0, 256, 2),            \
        V2R_VARIANT(MODE, 5, 2, 2, 0, 0, 256, 1), V2R_VARIANT(MODE, 6, 2, 2, 0, 0, 256, 1),            \
        V2R_VARIANT(MODE, 7, 2, 2, 0, 0, 256, 1), V2R_VARIANT(MODE, 8, 2, 2, 0, 0, 256, 1),            \
        V2R_VARIANT(MODE, 9, 2, 2, 0, 0, 256, 1), V2R_VARIANT(MODE, 10, 2, 2, 0, 0, 256, 1)
#define V2R_GEN_ROW(KIND)                                                                                    \
    V2R_VARIANT(MODE_GEN, 1, 2, 4, KIND, 0, 256, 2), V2R_VARIANT(MODE_GEN, 2, 2, 2, KIND, 0, 256, 2),        \
        V2R_VARIANT(MODE_GEN, 3, 2, 2, KIND, 0, 256, 2), V2R_VARIANT(MODE_GEN, 4, 2, 2, KIND, 0, 256, 2)
#define V2R_GEN_ROW_HI(KIND)                                                                                 \
    V2R_VARIANT(MODE_GEN, 5, 2, 2, KIND, 0, 256, 1), V2R_VARIANT(MODE_GEN, 6, 2, 2, KIND, 0, 256, 1),        \
        V2R_VARIANT(MODE_GEN, 7, 2, 2, KIND, 0, 256, 1), V2R_VARIANT(MODE_GEN, 8, 2, 2, KIND, 0, 256, 1),    \
        V2R_VARIANT(MODE_GEN, 9, 2, 2, KIND, 0, 256, 1)

static const Variant kVariants[] = {
    // tuning builds (V2R_IDW_TUNE=<tag>), kept to document what was measured; never picked by default
    V2R_TUNE("g2x2", MODE_GEN, 1, 2, 2, GEN_SINGLE, 0, 256, 2), V2R_TUNE("g4x4", MODE_GEN, 1, 4, 4, GEN_SINGLE, 0, 256, 1),
    V2R_TUNE("g4x2", MODE_GEN, 1, 4, 2, GEN_SINGLE, 0, 256, 2),
    V2R_PAIRED("c4x6", 4, 6, 256, 1), V2R_PAIRED_YS("ys4x6", 4, 6, 256, 1),
    V2R_KERNEL(MODE_RCP, 1, 4, 4, 1, 0, 256, 1, 7, true, "quad"),  // p = 2, FOUR points per reciprocal (half the MUFU + seed moves)
    V2R_PAIRED("nys", 4, 4, 256, 1),                // p = 2, paired, every lane computes its own dy^2 (round 1's loop)
    V2R_UNSTAGED("nys", MODE_RSQ, 1, 2, 4, 1, 0, 256, 2), V2R_UNSTAGED("nys", MODE_QRT, 1, 2, 4, 3, 0, 256, 2),
    V2R_UNSTAGED("nys", MODE_QRT, 9, 2, 2, 1, 1, 256, 1), V2R_UNSTAGED("nys", MODE_QRT, 4, 2, 2, 2, 1, 256, 2),
    V2R_UNSTAGED("nys", MODE_GEN, 1, 2, 4, GEN_SINGLE, 0, 256, 2), V2R_UNSTAGED("nys", MODE_GEN, 4, 2, 2, GEN_SINGLE, 0, 256, 2),
    // compile-time ladders for the headline shapes (looked up first)
    V2R_PAIRED_YS(nullptr, 4, 4, 256, 1),           // p = 2, paired, dy^2 staged in shared memory once per CTA (c2, c4)
    V2R_VARIANT(MODE_RCP, 1, 4, 4, 1, 0, 256, 1),   // p = 2, one reciprocal per pair (V2R_IDW_FP64_PAIRED=0)
    // rsqrt / root4 bases have longer dependent chains: 2x4 cells at two CTAs per SM measured +8 % over 4x4 at one
    V2R_VARIANT(MODE_RSQ, 1, 2, 4, 1, 0, 256, 2),   // p = 1
    V2R_VARIANT(MODE_RSQ, 1, 2, 4, 3, 0, 256, 2),   // p = 3
    V2R_VARIANT(MODE_QRT, 1, 2, 4, 1, 0, 256, 2),   // p = 0.5
    V2R_VARIANT(MODE_QRT, 1, 2, 4, 3, 0, 256, 2),   // p = 1.5 (the CLI default --es 1.5)
    V2R_VARIANT(MODE_QRT, 9, 2, 2, 1, 1, 256, 1),   // p = 0.5 .. 4.5 step 0.5  (config c3)
    V2R_VARIANT(MODE_QRT, 4, 2, 2, 2, 1, 256, 2),   // p = 1, 1.5, 2, 2.5       (config c5)
    // runtime ladders
    V2R_ROW(MODE_RCP), V2R_ROW(MODE_RSQ), V2R_ROW(MODE_QRT),
    // general exponents: all in (0, 1.99] (2^n by an integer add), all in (0, 3.75] (no specials in the loop), any value
    V2R_GEN_ROW(GEN_SINGLE),
    V2R_GEN_ROW(GEN_SAFE), V2R_GEN_ROW_HI(GEN_SAFE),
    V2R_GEN_ROW(GEN_ANY), V2R_GEN_ROW_HI(GEN_ANY),
};

static const Variant *pick_variant(const Pass &ps) {
    static const bool allow_paired = !(getenv("V2R_IDW_FP64_PAIRED") && atoi(getenv("V2R_IDW_FP64_PAIRED")) == 0);
    static const char *tune = getenv("V2R_IDW_TUNE");
    const Variant *fallback = nullptr;
    if (tune && *tune)
        for (const Variant &v : kVariants)
            if (v.tag && !strcmp(v.tag, tune) && v.mode == ps.mode && v.E == ps.count && v.k0 == ps.k0 && v.dk == ps.dk) return &v;
    for (const Variant &v : kVariants) {
        if (v.tag) continue;
        if (v.paired && !allow_paired) continue;
        if (v.mode != ps.mode || v.E != ps.count) continue;
        if (v.k0 == ps.k0 && v.dk == ps.dk) return &v;
        if (ps.mode == MODE_GEN) {
            if (ps.k0 == GEN_SINGLE && v.k0 == GEN_SAFE && !fallback) fallback = &v;  // no integer-add build for this E
            continue;
        }
        if (v.k0 == 0 && !fallback) fallback = &v;
    }
    return fallback;
}

static const char *mode_name(int m) {
    switch (m) {
        case MODE_RCP: return "rcp";
        case MODE_RSQ: return "rsqrt";
        case MODE_QRT: return "root4";
        default: return "general";
    }
}

// FP64-pipe instruction model per cell-point pair.
static double model_ops(const Pass &ps, const Variant &v) {
    const int E = ps.count;
    double per_point = 2.0 * (v.cr + v.cc) / double(v.cr * v.cc);  // dx, dx^2, dy, dy^2 amortised
    if (v.opt & 2) per_point = 2.0 / v.cr;                          // dy, dy^2 come from shared memory (once per CTA)
    double ops = per_point + 1.0 /* d2 */ + 2.0 * E /* fma + add */;
    auto ladder = [](int k) {  // multiplies of square-and-multiply
        int m = 0, hi = 0;
        for (int b = 0; b < 31; ++b)
            if (k >> b & 1) { hi = b; ++m; }
        return hi + m - 1;
    };
    if (ps.mode == MODE_GEN) {
        const double ex = v.k0 == GEN_SINGLE ? 7.0 : 8.0;  // tab_exp2: 3 rounding + 3 series + 1 or 2 scaling
        if (E == 1 && v.k0 != GEN_ANY) ops += 6.0 + ex;     // folded log2 (5 + the exponent's magic add)
        else ops += 7.0 + (1.0 + ex) * E;                   // log2 once, c*L + exp2 per exponent
    } else {
        ops += ps.mode == MODE_RCP ? 3 : (ps.mode == MODE_RSQ ? 5 : 7);
        if (v.paired) ops -= 0.5;  // 11 instead of 12 instructions per two pairs
        ops += ladder(ps.k0);
        if (E > 1) ops += ladder(ps.dk) + (E - 1);
    }
    return ops;
}

// Rough throughput of one GPU on a sweep (cell-point pairs per second), from the instruction model: used by the
// contexts to decide into how many rounds a device's band is cut (idw_api.cu).  Measured: 2.6e12 at 6 slots (p = 2),
// 0.45e12 at 36 (c3 sweep), 0.75e12 at 17.5 (general), 5.6e12 for the packed FP32 kernel.
double estimate_pair_rate(const double *exps, int n_exps, int precision) {
    std::vector<Pass> passes;
    if (plan_sweep(exps, n_exps, passes) != V2R_IDW_OK) return 1e12;
    double slots = 0;
    for (const Pass &ps : passes) {
        const Variant *v = pick_variant(ps);
        slots += v ? model_ops(ps, *v) + (ps.mode == MODE_GEN ? 4.0 * ps.count : 0.0) : 8.0;
    }
    const double fp64 = 2.6e12 * 6.0 / std::max(slots, 1.0);
    return precision == V2R_IDW_FP32 ? 2.0 * fp64 : fp64;
}

static std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }
void count_launch() { g_launches++; }

// ------------------------------------------------------------------------------------------
// Per-device facts: SM count, table image of the general class, occupancy of each kernel
// ------------------------------------------------------------------------------------------
struct DeviceFacts {
    int sms = 0;
    GenTables *tables = nullptr;
    std::map<const void *, int> ctas_per_sm;  // kernel -> resident CTAs per SM with its shared memory
    bool pool_ready = false;
};
static std::mutex g_facts_mu;
static std::map<int, DeviceFacts> g_facts;

// resident CTAs of `fn` on the whole device (0 on error, message set)
int device_slots(int device, const void *fn, int threads, int smem, const GenTables **tables) {
    std::lock_guard<std::mutex> lk(g_facts_mu);
    DeviceFacts &f = g_facts[device];
    if (!f.sms) cudaDeviceGetAttribute(&f.sms, cudaDevAttrMultiProcessorCount, device);
    if (!f.pool_ready) {  // keep freed scratch in the stream-ordered pool instead of returning it to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        f.pool_ready = true;
    }
    if (tables) {
        if (!f.tables) {
            static GenTables host;  // identical for every device
            static bool built = false;
            if (!built) { build_gen_tables(host); built = true; }
            if (cudaMalloc((void **)&f.tables, sizeof(GenTables)) != cudaSuccess ||
                cudaMemcpy(f.tables, &host, sizeof(GenTables), cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaGetLastError();
                f.tables = nullptr;
                set_error(V2R_IDW_ENOMEM, "general-class tables: allocation / upload failed");
                return 0;
            }
        }
        *tables = f.tables;
    }
    auto it = f.ctas_per_sm.find(fn);
    if (it == f.ctas_per_sm.end()) {
        cudaError_t ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int occ = 0;
        if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, (size_t)smem);
        if (ce != cudaSuccess || occ < 1) {
            cudaGetLastError();
            set_error(V2R_IDW_ECUDA, "kernel needs %d bytes of shared memory: %s", smem, ce == cudaSuccess ? "does not fit an SM" : cudaGetErrorString(ce));
            return 0;
        }
        it = f.ctas_per_sm.emplace(fn, occ).first;
    }
    return f.sms * it->second;
}

// Work decomposition of one launch (idw_kernels.cuh): whole tiles for the full waves, (tile, segment) units for the
// tiles of the last partial wave.  `segs` / `seg_tiles` depend on n only, so a cell's value does not depend on the cut.
void plan_grid(KParams &kp, int tile_r, int tile_c, int slots) {
    static const bool tail_split = !(getenv("V2R_IDW_TAIL") && atoi(getenv("V2R_IDW_TAIL")) == 0);
    const long long tiles_x = (kp.cols + tile_c - 1) / tile_c, tiles_y = (kp.nrows + tile_r - 1) / tile_r;
    kp.tiles_x = (int)tiles_x;
    kp.n_tiles = tiles_x * tiles_y;
    const long long ntp = (kp.n + kTilePoints - 1) / kTilePoints;
    long long K = std::max<long long>(1, std::min<long long>(ntp, kMaxSegs));
    const long long seg_tiles = std::max<long long>(1, (ntp + K - 1) / K);
    K = std::max<long long>(1, (ntp + seg_tiles - 1) / seg_tiles);
    kp.segs = (int)K;
    kp.seg_tiles = (int)seg_tiles;
    const long long R = kp.n_tiles % slots;
    if (K == 1 || R == 0 || !tail_split) {
        kp.n_main = kp.n_tiles;
        kp.n_tail = 0;
    } else {
        kp.n_main = kp.n_tiles - R;
        kp.n_tail = (int)std::min<long long>(slots, R * K);
    }
}

// Stream-ordered scratch for the tail tiles: partials, then the arrival counters and prefix lengths (2 R words, zeroed),
// then one flag word.
int alloc_scratch(cudaStream_t stream, KParams &kp, int nacc, int threads, void **block, unsigned **flag) {
    const long long R = kp.n_tiles - kp.n_main;
    const size_t partial = (size_t)R * kp.segs * nacc * threads * sizeof(double);
    const size_t words = 2 * (size_t)R + 4;
    cudaError_t ce = cudaMallocAsync(block, partial + words * sizeof(unsigned), stream);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        return set_error(V2R_IDW_ENOMEM, "cudaMallocAsync(%zu bytes of tail scratch): %s", partial + words * 4, cudaGetErrorString(ce));
    }
    kp.scratch = reinterpret_cast<double *>(*block);
    kp.counters = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(*block) + partial);
    if (flag) *flag = kp.counters + 2 * R;
    ce = cudaMemsetAsync(kp.counters, 0, words * sizeof(unsigned), stream);
    if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "cudaMemsetAsync: %s", cudaGetErrorString(ce));
    return V2R_IDW_OK;
}

// Exact-hit rule (features/idw/idw_utils.go:13-16): every keyed cell returns its point's weight,
// for every exponent.  Runs after the main kernel on the same stream.
__global__ void idw_patch_exact_kernel(const double *__restrict__ w, const long long *__restrict__ kr,
                                       const long long *__restrict__ kc, long long n, long long cols,
                                       long long row0, long long nrows, double *__restrict__ out,
                                       long long plane, int n_exps) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long r = kr[i], c = kc[i];
    if (r < row0 || r >= row0 + nrows || c < 0 || c >= cols) return;
    double v = w[i];
    for (int e = 0; e < n_exps; ++e) out[(long long)e * plane + (r - row0) * cols + c] = v;
}

int launch_patch_exact(cudaStream_t stream, const double *d_z, const long long *d_kr, const long long *d_kc, long long n,
                       const v2r_idw_grid &g, long long row0, long long nrows, double *d_out, int n_exps) {
    if (n <= 0 || !d_kr || !d_kc) return V2R_IDW_OK;
    const int tb = 256;
    idw_patch_exact_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, stream>>>(d_z, d_kr, d_kc, n, g.cols, row0, nrows, d_out,
                                                                           nrows * g.cols, n_exps);
    g_launches++;
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "patch kernel launch: %s", cudaGetErrorString(ce));
    return V2R_IDW_OK;
}

// Kernels without in-loop special cases (general class with positive exponents, packed FP32) rely on this pass for
// the reference's d == 0 rule on non-key cells (features/idw/idw_utils.go:19-25 with tools/utils.go:166-168:
// Pow(0, -p) = +Inf, Inf/Inf = NaN): a cell whose squared distance to some point -- evaluated exactly like the
// kernels and the reference do, point minus node, each square rounded on its own -- is zero or subnormal becomes NaN
// in the planes [e0, e0 + ne).  A NaN coordinate makes every weight NaN in the reference: it raises *flag and
// idw_nanfill_kernel then fills those planes.  The exact-hit patch runs afterwards and overrides both.
__global__ void idw_poison_kernel(const double *__restrict__ x, const double *__restrict__ y, long long n, double x_min,
                                  double x_step, double y_min, double y_step, long long cols, long long row0, long long nrows,
                                  double *__restrict__ out, long long plane, int ne, unsigned *flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = x[i], py = y[i];
    if (px != px || py != py) { if (flag) *flag = 1u; return; }
    const double qc = rint((px - x_min) / x_step), qr = rint((py - y_min) / y_step);
    if (!(qc >= -2.0 && qc <= (double)cols + 1.0 && qr >= (double)row0 - 2.0 && qr <= (double)(row0 + nrows) + 1.0)) return;
    const long long c0 = (long long)qc, r0 = (long long)qr;
    for (long long r = r0 - 1; r <= r0 + 1; ++r) {
        if (r < row0 || r >= row0 + nrows) continue;
        const double dy = __dsub_rn(py, __dadd_rn(y_min, __dmul_rn((double)r, y_step)));
        const double dy2 = __dmul_rn(dy, dy);
        for (long long c = c0 - 1; c <= c0 + 1; ++c) {
            if (c < 0 || c >= cols) continue;
            const double dx = __dsub_rn(px, __dadd_rn(x_min, __dmul_rn((double)c, x_step)));
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), dy2);
            if (d2 < 2.2250738585072014e-308)
                for (int e = 0; e < ne; ++e) out[(long long)e * plane + (r - row0) * cols + c] = __longlong_as_double(0x7ff8000000000000LL);
        }
    }
}
__global__ void idw_nanfill_kernel(double *__restrict__ out, long long count, const unsigned *flag) {
    if (*flag == 0u) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        out[i] = __longlong_as_double(0x7ff8000000000000LL);
}

int launch_poison(cudaStream_t stream, const double *d_x, const double *d_y, long long n, const v2r_idw_grid &g, long long row0,
                  long long nrows, double *d_out, int ne, unsigned *flag) {
    if (n <= 0) return V2R_IDW_OK;
    const int tb = 256;
    const long long plane = nrows * g.cols;
    idw_poison_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, stream>>>(d_x, d_y, n, g.x_min, g.x_step, g.y_min, g.y_step, g.cols, row0,
                                                                      nrows, d_out, plane, ne, flag);
    g_launches++;
    if (flag) {
        const long long count = plane * ne;
        idw_nanfill_kernel<<<(unsigned)std::min<long long>(4096, (count + tb - 1) / tb), tb, 0, stream>>>(d_out, count, flag);
        g_launches++;
    }
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "poison kernel launch: %s", cudaGetErrorString(ce));
    return V2R_IDW_OK;
}

void fill_params(KParams &kp, const double *d_x, const double *d_y, const double *d_z, long long n, const v2r_idw_grid &g,
                 long long row0, long long nrows) {
    kp.x = d_x; kp.y = d_y; kp.w = d_z; kp.n = n;
    kp.x_min = g.x_min; kp.x_step = g.x_step; kp.y_min = g.y_min; kp.y_step = g.y_step;
    kp.inv_x_step = 1.0 / g.x_step; kp.inv_y_step = 1.0 / g.y_step;
    kp.cols = g.cols; kp.row0 = row0; kp.nrows = nrows;
    kp.plane = nrows * g.cols;
}

int launch_rows_fp64(int device, cudaStream_t stream, const double *d_x, const double *d_y, const double *d_z,
                     const long long *d_kr, const long long *d_kc, long long n, const v2r_idw_grid &g,
                     long long row0, long long nrows, const double *exps, int n_exps, double *d_out) {
    std::vector<Pass> passes;
    int rc = plan_sweep(exps, n_exps, passes);
    if (rc) return rc;
    const long long plane = nrows * g.cols;
    for (const Pass &ps : passes) {
        const Variant *v = pick_variant(ps);
        if (!v) return set_error(V2R_IDW_EINVAL, "no kernel variant for mode %d E %d", ps.mode, ps.count);
        KParams kp{};
        fill_params(kp, d_x, d_y, d_z, n, g, row0, nrows);
        kp.out = d_out + (long long)ps.first * plane;
        kp.k0 = ps.k0 > 0 ? ps.k0 : 1;
        kp.dk = ps.dk > 0 ? ps.dk : 1;
        for (int i = 0; i < ps.count; ++i) kp.cexp[i] = kExpScale * ps.cexp[i];  // exact scaling by a power of two
        for (int i = 0; i < 3; ++i) kp.lg_s[i] = kp.cexp[0] * kLgHost[i];
        const GenTables *tables = nullptr;
        const int slots = device_slots(device, (const void *)v->fn, v->threads, v->smem, ps.mode == MODE_GEN ? &tables : nullptr);
        if (!slots) return V2R_IDW_ECUDA;
        kp.tables = tables;
        plan_grid(kp, (v->threads / 32) * v->cr, 32 * v->cc, slots);
        const long long grid = kp.n_main + kp.n_tail;
        if (grid > 2147483647LL || kp.n_tiles * kp.segs > 2147483647LL)
            return set_error(V2R_IDW_EINVAL, "band of %lld rows x %lld cols exceeds one launch; use smaller bands", nrows, (long long)g.cols);
        const bool poison = ps.mode == MODE_GEN && v->k0 != GEN_ANY;
        void *block = nullptr;
        unsigned *flag = nullptr;
        if (kp.n_tail > 0 || poison) {
            if ((rc = alloc_scratch(stream, kp, 2 * v->E * v->cr * v->cc, v->threads, &block, &flag))) return rc;
        }
        v->fn<<<(unsigned)grid, v->threads, v->smem, stream>>>(kp);
        g_launches++;
        cudaError_t ce = cudaGetLastError();
        if (ce == cudaSuccess && poison) {
            if ((rc = launch_poison(stream, d_x, d_y, n, g, row0, nrows, kp.out, ps.count, flag))) ce = cudaErrorUnknown;
        }
        if (block) cudaFreeAsync(block, stream);
        if (rc) return rc;
        if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "idw kernel launch: %s", cudaGetErrorString(ce));
    }
    return launch_patch_exact(stream, d_z, d_kr, d_kc, n, g, row0, nrows, d_out, n_exps);
}

int validate_geometry(long long n, const v2r_idw_grid *g, long long row0, long long nrows, int n_exps) {
    if (!g) return set_error(V2R_IDW_EINVAL, "grid is NULL");
    if (n < 0) return set_error(V2R_IDW_EINVAL, "n < 0");
    if (g->rows < 0 || g->cols < 0) return set_error(V2R_IDW_EINVAL, "negative grid dimensions");
    if (row0 < 0 || nrows < 0 || row0 + nrows > g->rows)
        return set_error(V2R_IDW_EINVAL, "rows [%lld, %lld) outside the grid of %lld rows", row0, row0 + nrows, (long long)g->rows);
    if (n_exps < 1 || n_exps > kMaxSweep) return set_error(V2R_IDW_EINVAL, "n_exps must be in [1, %d]", kMaxSweep);
    return V2R_IDW_OK;
}

int validate_launch(const void *d_x, const void *d_y, const void *d_z, long long n, const v2r_idw_grid *g,
                    long long row0, long long nrows, int n_exps, const void *d_out) {
    int rc = validate_geometry(n, g, row0, nrows, n_exps);
    if (rc) return rc;
    if (n > 0 && (!d_x || !d_y || !d_z)) return set_error(V2R_IDW_EINVAL, "point arrays are NULL");
    if (nrows > 0 && g->cols > 0 && !d_out) return set_error(V2R_IDW_EINVAL, "out is NULL");
    if (((uintptr_t)d_x | (uintptr_t)d_y | (uintptr_t)d_z) & 15)
        return set_error(V2R_IDW_EINVAL, "device point arrays must be 16-byte aligned (TMA bulk copies)");
    return V2R_IDW_OK;
}

}  // namespace v2r

using namespace v2r;

extern "C" int v2r_idw_launch_rows(int device, void *stream, const double *d_x, const double *d_y,
                                   const double *d_z, const int64_t *d_key_r, const int64_t *d_key_c,
                                   int64_t n, const v2r_idw_grid *grid, int64_t row0, int64_t nrows,
                                   const double *exps, int n_exps, int precision, double *d_out) {
    int rc = validate_launch(d_x, d_y, d_z, n, grid, row0, nrows, n_exps, d_out);
    if (rc) return rc;
    if (precision != V2R_IDW_FP64 && precision != V2R_IDW_FP32) return set_error(V2R_IDW_EINVAL, "unknown precision %d", precision);
    if (nrows == 0 || grid->cols == 0) return V2R_IDW_OK;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(ce));
    return guarded([&] {
        if (precision == V2R_IDW_FP32)
            return launch_rows_fp32(device, (cudaStream_t)stream, d_x, d_y, d_z, (const long long *)d_key_r, (const long long *)d_key_c, n,
                                    *grid, row0, nrows, exps, n_exps, d_out);
        return launch_rows_fp64(device, (cudaStream_t)stream, d_x, d_y, d_z, (const long long *)d_key_r, (const long long *)d_key_c, n,
                                *grid, row0, nrows, exps, n_exps, d_out);
    });
}

extern "C" int v2r_idw_describe(const double *exps, int n_exps, int precision, char *buf, size_t buf_len,
                                double *fp64_ops_per_pair) {
  return guarded([&] {
    std::vector<Pass> passes;
    int rc = plan_sweep(exps, n_exps, passes);
    if (rc) return rc;
    std::string s;
    double ops = 0;
    if (precision == V2R_IDW_FP32) {
        s = describe_fp32(passes);
    } else {
        for (const Pass &ps : passes) {
            const Variant *v = pick_variant(ps);
            if (!v) return set_error(V2R_IDW_EINVAL, "no kernel variant");
            char tmp[160];
            snprintf(tmp, sizeof tmp, "%sfp64/%s E=%d cells=%dx%d k0=%d dk=%d%s", s.empty() ? "" : " + ", mode_name(ps.mode),
                     ps.count, v->cr, v->cc, ps.k0, ps.dk,
                     v->tag ? v->tag : (v->paired ? " (paired reciprocal)" : (ps.mode == MODE_GEN ? (v->k0 == GEN_SINGLE ? " (0 < p <= 1.99)" : (v->k0 == GEN_SAFE ? " (0 < p <= 3.75)" : " (any p)")) : (v->k0 ? " (static ladder)" : ""))));
            s += tmp;
            ops += model_ops(ps, *v);
        }
    }
    if (buf && buf_len) {
        strncpy(buf, s.c_str(), buf_len - 1);
        buf[buf_len - 1] = 0;
    }
    if (fp64_ops_per_pair) *fp64_ops_per_pair = ops;
    return (int)V2R_IDW_OK;
  });
}

// ------------------------------------------------------------------------------------------
// Roofline denominators: issue-rate microbenchmarks (MEASURED_PEAKS.json has no FP64 entry).
// ------------------------------------------------------------------------------------------
namespace v2r {
template <int WHICH>
__global__ void __launch_bounds__(256) peak_kernel(double *out, int iters) {
    const double s = 1.0 + 1e-9 * threadIdx.x;
    if constexpr (WHICH == 0) {  // DFMA
        double a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = s + j;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], 0.999999, 1e-7);
        double t = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += a[j];
        if (t == 12345.678) out[0] = t;
    } else if constexpr (WHICH == 1) {  // MUFU.RCP64H
        double a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = s + j;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = rcp_seed(a[j]);
        double t = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += a[j];
        if (t == 12345.678) out[0] = t;
    } else if constexpr (WHICH == 2) {  // MUFU.RCP (fp32)
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = (float)s + j;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float r;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[j]));
                a[j] = r + 0.5f;  // the FADD (FMA pipe) stops the compiler folding rcp(rcp(x)) -> x
            }
        float t = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += a[j];
        if (t == 12345.678f) out[0] = t;
    } else if constexpr (WHICH == 4) {  // DFMA with three distinct register-pair sources (what the IDW kernels issue)
        double a[8], b[8], c[8];
        const double rnd = 1e-12 * (double)((clock64() + threadIdx.x) & 1023);  // unknown at compile time: operands must live in registers
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] = s + j; b[j] = 0.999999 + rnd * (j + 1); c[j] = 1.000001 - rnd * (2 * j + 1); }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(b[j], c[(j + 1) & 7], a[j]);
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = fma(a[j], c[j], b[(j + 3) & 7]);
        }
        double t = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += a[j] + b[j];
        if (t == 12345.678) out[0] = t;
    } else if constexpr (WHICH == 5) {  // DADD, two register sources
        double a[8], b[8];
        const double rnd = 1e-12 * (double)((clock64() + threadIdx.x) & 1023);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] = s + j; b[j] = rnd * (j + 1); }
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] + b[(j + (i & 1)) & 7];
        double t = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += a[j];
        if (t == 12345.678) out[0] = t;
    } else {  // FFMA
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = (float)s + j;
        const float m = 0.999999f + 1e-9f * threadIdx.x, c = 1e-7f * threadIdx.x;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], m, c);
        float t = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += a[j];
        if (t == 12345.678f) out[0] = t;
    }
}

template <int WHICH>
static int run_peak(double ms_target, int sms, double *ginst) {
    double *d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return set_error(V2R_IDW_ENOMEM, "cudaMalloc");
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = sms * 8;
    int iters = 2000;
    float ms = 0;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        peak_kernel<WHICH><<<blocks, 256>>>(d, iters);
        g_launches++;
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) {
            cudaFree(d);
            return set_error(V2R_IDW_ECUDA, "peak kernel: %s", cudaGetErrorString(cudaGetLastError()));
        }
        cudaEventElapsedTime(&ms, a, b);
        double rate = (double)blocks * 256 * iters * (WHICH == 4 ? 16 : 8) / (ms * 1e-3) / 1e9;
        if (rep > 0 && rate > best) best = rate;
        if (ms < ms_target) iters = (int)(iters * (ms_target / (ms > 0.01f ? ms : 0.01f)));
        if (iters > 4000000) iters = 4000000;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *ginst = best;
    return V2R_IDW_OK;
}
}  // namespace v2r

extern "C" int v2r_idw_measure_peaks(int device, double ms, double *dfma, double *mufu64, double *mufu32, double *ffma) {
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(ce));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (ms <= 0) ms = 50;
    double v = 0;
    int rc;
    if (dfma) { if ((rc = run_peak<0>(ms, sms, &v))) return rc; *dfma = v; }
    if (mufu64) { if ((rc = run_peak<1>(ms, sms, &v))) return rc; *mufu64 = v; }
    if (mufu32) { if ((rc = run_peak<2>(ms, sms, &v))) return rc; *mufu32 = v; }
    if (ffma) { if ((rc = run_peak<3>(ms, sms, &v))) return rc; *ffma = v; }
    return V2R_IDW_OK;
}

extern "C" int v2r_idw_measure_peak(int device, double ms, int which, double *ginst_s) {
    if (!ginst_s) return set_error(V2R_IDW_EINVAL, "ginst_s is NULL");
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(ce));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (ms <= 0) ms = 50;
    switch (which) {
        case V2R_IDW_PEAK_DFMA: return run_peak<0>(ms, sms, ginst_s);
        case V2R_IDW_PEAK_MUFU_RCP64H: return run_peak<1>(ms, sms, ginst_s);
        case V2R_IDW_PEAK_MUFU_RCP_F32: return run_peak<2>(ms, sms, ginst_s);
        case V2R_IDW_PEAK_FFMA: return run_peak<3>(ms, sms, ginst_s);
        case V2R_IDW_PEAK_DFMA_3SRC: return run_peak<4>(ms, sms, ginst_s);
        case V2R_IDW_PEAK_DADD: return run_peak<5>(ms, sms, ginst_s);
        default: return set_error(V2R_IDW_EINVAL, "unknown microbenchmark %d", which);
    }
}

extern "C" int64_t v2r_idw_launch_count(void) { return launch_count(); }
