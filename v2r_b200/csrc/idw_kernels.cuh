// idw_kernels.cuh -- sm_100a kernels of the IDW gridding hot path (FP64 classes) and the tile driver they share with
// the FP32 fast path (idw_kernels_fp32.cuh).
//
// Replaces the per-cell loop of the reference (features/idw/idw_utils.go:8-27 calculateIDW called
// from features/idw/idw.go:34-40 and features/idw/idw_chunk.go:27-41):
//     z(r,c) = sum_i w_i * d_i^-p / sum_i d_i^-p ,  d_i = |cell(r,c) - point_i|
// for every cell of a row band, for E exponents in ONE pass over the points.
//
// Shape of the computation:
//   * a CTA owns a tile of (WARPS*CR) rows x (32*CC) columns; a thread owns a CR x CC register block
//     (lanes run along columns so the raster stores coalesce);
//   * the point stream (SoA x[], y[], w[]) is brought into shared memory by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), kStages deep; every thread reads the same point at the
//     same time, i.e. shared-memory broadcasts;
//   * per point and thread: CR squared dy and CC squared dx (reference: Pow(dx,2), Pow(dy,2),
//     tools/utils.go:157 -- both exactly rounded products), then per pair ONE add d2 = dx2 + dy2
//     (bit-identical to the reference's `total`), the power d2^(-p/2), one FMA and one ADD;
//   * d^-p never goes through sqrt+pow: it is rcp / rsqrt / fourth-root seeds from MUFU.RCP64H /
//     MUFU.RSQ64H refined by one cubic step in the FP64 pipe, or exp2(c*log2 d2) for general p.
//
// Work decomposition (the "tail" of a launch):
//   * the points are cut into `segs` segments (a function of n only); a cell's sums are DEFINED as
//     ((S_0 + S_1) + S_2) + ... over the per-segment sums, so the value of a cell does not depend on
//     how the launch was cut into tiles, bands or devices (band / multi-GPU results are bit-identical);
//   * CTAs [0, n_main) own one whole tile each: per-segment sums are folded into a running total that
//     lives in shared memory ("stash"), touched once per segment;
//   * the tiles that would form the last, partial wave are shared by the CTAs [n_main, n_main + n_tail):
//     each takes a contiguous run of (tile, segment) units, writes per-segment partials to a global
//     scratch area, and the CTA that delivers a tile's last segment adds them up in segment order and
//     writes the cells.  The idle tail of a launch shrinks from one tile time to one segment time.
//
// No tensor cores: this is not a contraction (every pair needs its own reciprocal power).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "idw_math.cuh"

namespace v2r {

enum Mode : int {
    MODE_RCP = 0,  // all p are positive even integers : base b = 1/d2,        h = b^(p/2)
    MODE_RSQ = 1,  // all p are positive integers      : base b = d2^(-1/2),   h = b^p
    MODE_QRT = 2,  // all p are positive multiples of .5: base b = d2^(-1/4),  h = b^(2p)
    MODE_GEN = 3   // anything else (incl. p <= 0)      : h = exp2(c * log2 d2), c = -p/2 ; p == 0 -> h = 1
};
// MODE_GEN flavours (template parameter K0 of the kernel)
enum GenKind : int {
    GEN_ANY = 0,     // any exponents: saturating log2, clamped exp2 (p <= 0, large p, d2 == 0 handled in the loop)
    GEN_SAFE = 1,    // all exponents in (0, 3.75]: no special cases in the loop (cells with d2 ~ 0 are poisoned afterwards)
    GEN_SINGLE = 2   // all exponents in (0, 1.99]: additionally 2^n is one integer add (result always normal)
};

constexpr int kMaxE = 10;        // exponents per pass (register budget); longer sweeps run several passes
constexpr int kTilePoints = 512; // points per shared-memory stage
constexpr int kStages = 3;
constexpr int kMaxSegs = 16;     // point segments per cell sum

struct KParams {
    const double *x, *y, *w;  // device SoA, 16-byte aligned
    long long n;              // number of points
    double x_min, x_step, y_min, y_step;
    long long cols;           // grid columns
    long long row0, nrows;    // band [row0, row0 + nrows)
    double *out;              // E planes of nrows*cols
    long long plane;          // nrows * cols
    int k0, dk;               // integer power ladder: h_e = b^(k0 + e*dk) (runtime variant)
    double cexp[kMaxE];       // MODE_GEN: 512 * (-p_e/2): the exponent multiplier, pre-scaled for tab_exp2
    double lg_s[3];           // MODE_GEN, E == 1: cexp[0] * kLg[] (the multiplier folded into the log series)
    // work decomposition (plan_grid, idw_launch.cu)
    int tiles_x;              // CTA tiles per row of tiles
    long long n_tiles;        // tiles_x * tile rows
    long long n_main;         // CTAs [0, n_main) own one whole tile each
    int n_tail;               // CTAs [n_main, n_main + n_tail) share tiles [n_main, n_tiles) by point segments
    int segs;                 // point segments (function of n only)
    int seg_tiles;            // point tiles per segment
    double *scratch;          // tail partials: [tile - n_main][segment][accumulator][thread]
    unsigned *counters;       // tail tiles: [0, R) segments delivered so far, [R, 2R) length of the prefix run (zero at launch)
    const GenTables *tables;  // MODE_GEN: table image in global memory (built once per device)
    double inv_x_step, inv_y_step;  // tile guard of the paired kernels (approximate reciprocals are enough)
};

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------- shared memory layout
// [ stages: kStages x {x, y, w} x kTilePoints doubles | mbarriers + flags (128 B) | body-specific area | stash ]
constexpr int kStageBytes = kStages * 3 * kTilePoints * (int)sizeof(double);
constexpr int kCtrlBytes = 128;
struct Ctrl {
    uint64_t full[kStages];
    int finisher;     // tail tiles: this CTA delivered the last segment
    int guard[2];     // paired kernels: point tile `it` is unsafe for pairing (slot it & 1)
};
static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block");

__host__ __device__ constexpr int smem_bytes(int body_bytes, int nacc, int threads) {
    return kStageBytes + kCtrlBytes + body_bytes + nacc * threads * (int)sizeof(double);
}

// ---------------------------------------------------------------------------- tile driver
// Streams the point tiles of segments [s0, s1) of one cell tile through the TMA pipeline and hands them to Body:
//   static constexpr int THREADS, NACC, BODY_BYTES;   double acc[NACC];
//   void setup(p, tile_row, tile_col)          cell coordinates of this thread for tile (tile_row, tile_col)
//   void points(p, sx, sy, sw, cnt, ctrl, it)  accumulate one point tile into acc[] (may __syncthreads: every thread calls it)
//   void store(p)                              out = num / den for the thread's cells
// WHOLE = all segments of the tile (the CTAs [0, n_main), one call each: the lean instantiation, whose loop state
// is two counters -- the unrolled point loops of the bodies run at the register limit, and every extra live value
// costs the scheduler operand-reuse slots); !WHOLE = a run of segments of a tail tile (the general instantiation).
// `it` counts the point tiles this CTA has consumed: stage = it % kStages, mbarrier parity = (it / kStages) & 1.
template <bool WHOLE, class Body>
__device__ __forceinline__ unsigned run_piece(const KParams &p, Body &body, unsigned char *smem, int tile, int s0, int s1, unsigned it) {
    constexpr int T = kTilePoints, S = kStages, THREADS = Body::THREADS, NACC = Body::NACC;
    double(*sm)[3][T] = reinterpret_cast<double(*)[3][T]>(smem);
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(smem + kStageBytes);
    double *stash = reinterpret_cast<double *>(smem + kStageBytes + kCtrlBytes + Body::BODY_BYTES);
    const int tid = threadIdx.x;
    const int K = p.segs;
    const int ntp = (int)((p.n + T - 1) / T);  // point tiles
    const int nfull = (int)(p.n / T);          // tiles fetched by TMA; a ragged last tile is loaded by the threads
    const int pt_end = min(s1 * p.seg_tiles, ntp);
    const int tma_end = min(pt_end, nfull);
    const bool whole = WHOLE || (s0 == 0 && s1 == K);

    auto issue = [&](int t, int st) {
        mbar_expect_tx(&ctrl->full[st], 3u * T * (uint32_t)sizeof(double));
        tma_load_1d(&sm[st][0][0], p.x + (long long)t * T, T * (uint32_t)sizeof(double), &ctrl->full[st]);
        tma_load_1d(&sm[st][1][0], p.y + (long long)t * T, T * (uint32_t)sizeof(double), &ctrl->full[st]);
        tma_load_1d(&sm[st][2][0], p.w + (long long)t * T, T * (uint32_t)sizeof(double), &ctrl->full[st]);
    };

    body.setup(p, tile / p.tiles_x, tile % p.tiles_x);
#pragma unroll
    for (int i = 0; i < NACC; ++i) body.acc[i] = 0.0;
    int pt = s0 * p.seg_tiles;
    if (tid == 0)
        for (int i = 0; i < S - 1; ++i)
            if (pt + i < tma_end) issue(pt + i, (int)((it + i) % S));

    for (int seg = s0; seg < s1; ++seg) {
        const int seg_end = min((seg + 1) * p.seg_tiles, ntp);
        for (; pt < seg_end; ++pt, ++it) {
            const int st = (int)(it % S);
            // stage (it-1)%S was released by the __syncthreads that ended the previous iteration
            if (tid == 0 && pt + S - 1 < tma_end) issue(pt + S - 1, (int)((it + S - 1) % S));
            int cnt = T;
            if (pt >= nfull) {  // ragged last tile: ordinary loads, then complete the stage's phase by hand
                const long long base = (long long)pt * T;
                cnt = (int)(p.n - base);
                for (int i = tid; i < cnt; i += THREADS) {
                    sm[st][0][i] = p.x[base + i];
                    sm[st][1][i] = p.y[base + i];
                    sm[st][2][i] = p.w[base + i];
                }
                __syncthreads();
                if (tid == 0) mbar_expect_tx(&ctrl->full[st], 0u);
            }
            mbar_wait(&ctrl->full[st], (it / S) & 1u);
            body.points(p, sm[st][0], sm[st][1], sm[st][2], cnt, ctrl, it);
            __syncthreads();
        }
        // end of segment `seg`: fold the segment sums into the running total -- whole tiles and the PREFIX run of a tail
        // tile (s0 == 0: its sums are a prefix of the defined order, so they may be folded on the spot) use the
        // shared-memory stash, the last segment of the run stays in registers -- or hand them over one by one (the
        // other runs of a tail tile: global scratch)
        if (whole || s0 == 0) {
            if (seg + 1 < s1) {
                if (seg == 0) {
#pragma unroll
                    for (int i = 0; i < NACC; ++i) stash[i * THREADS + tid] = body.acc[i];
                } else {
#pragma unroll
                    for (int i = 0; i < NACC; ++i) stash[i * THREADS + tid] += body.acc[i];
                }
#pragma unroll
                for (int i = 0; i < NACC; ++i) body.acc[i] = 0.0;
            }
        } else {
            double *dst = p.scratch + (((long long)(tile - (int)p.n_main) * K + seg) * NACC) * THREADS + tid;
#pragma unroll
            for (int i = 0; i < NACC; ++i) __stcg(dst + (long long)i * THREADS, body.acc[i]);
#pragma unroll
            for (int i = 0; i < NACC; ++i) body.acc[i] = 0.0;
        }
    }
    if (whole || s0 == 0) {
        if (s1 - s0 > 1) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) body.acc[i] = stash[i * THREADS + tid] + body.acc[i];
        }
    }
    if (whole) {
        body.store(p);
    } else {
        const long long slot = tile - (int)p.n_main;
        const long long tail_tiles = p.n_tiles - p.n_main;
        if (s0 == 0) {  // the prefix total S_0 + ... + S_{s1-1} goes into the scratch slot of its last segment
            double *dst = p.scratch + ((slot * K + (s1 - 1)) * NACC) * THREADS + tid;
#pragma unroll
            for (int i = 0; i < NACC; ++i) __stcg(dst + (long long)i * THREADS, body.acc[i]);
            if (tid == 0) p.counters[tail_tiles + slot] = (unsigned)s1;  // prefix length, read by the finisher
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned old = atomicAdd(&p.counters[slot], (unsigned)(s1 - s0));
            ctrl->finisher = (old + (unsigned)(s1 - s0) == (unsigned)K);
        }
        __syncthreads();
        if (ctrl->finisher) {  // every segment of this tile is in scratch: prefix total, then the rest in segment order
            __threadfence();
            const int L = (int)__ldcg(&p.counters[tail_tiles + slot]);
            const double *src = p.scratch + ((slot * K + (L - 1)) * NACC) * THREADS + tid;
#pragma unroll
            for (int i = 0; i < NACC; ++i) body.acc[i] = __ldcg(src + (long long)i * THREADS);
            for (int s = L; s < K; ++s) {  // NACC independent loads in flight per step
                src += (long long)NACC * THREADS;
#pragma unroll
                for (int i = 0; i < NACC; ++i) body.acc[i] += __ldcg(src + (long long)i * THREADS);
            }
            body.store(p);
        }
        __syncthreads();
    }
    return it;
}

template <class Body>
__device__ __forceinline__ void tile_driver(const KParams &p, Body &body, unsigned char *smem) {
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(smem + kStageBytes);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&ctrl->full[s], 1);
        ctrl->finisher = 0;
        ctrl->guard[0] = ctrl->guard[1] = 0;
        mbar_fence_init();
    }
    __syncthreads();  // also publishes whatever Body's constructor wrote to its shared area
    const int K = p.segs;
    if ((long long)blockIdx.x < p.n_main) {
        run_piece<true>(p, body, smem, (int)blockIdx.x, 0, K, 0u);
    } else {
        // tail CTA: a contiguous run of (tile, segment) units, unit = tile * K + segment
        const long long j = (long long)blockIdx.x - p.n_main;
        const long long U = (p.n_tiles - p.n_main) * K;
        int u = (int)(p.n_main * K + j * U / p.n_tail);
        const int u1 = (int)(p.n_main * K + (j + 1) * U / p.n_tail);
        unsigned it = 0;
        while (u < u1) {
            const int tile = u / K, s0 = u - tile * K;
            const int s1 = (u1 - u < K - s0) ? s0 + (u1 - u) : K;
            it = run_piece<false>(p, body, smem, tile, s0, s1, it);
            u += s1 - s0;
        }
    }
}

// ---------------------------------------------------------------------------- powers
// b^K as a balanced product b^(K/2) * b^(K - K/2): depth log2 K instead of K.  In a sweep every exponent calls
// this with its own K; the common sub-powers are merged by the compiler (pure FP multiplies of identical
// operands), so consecutive powers 1..9 still cost 8 multiplies but the dependent chain is 4 deep, not 8.
template <int K>
__device__ __forceinline__ double ipow_c(double b) {
    static_assert(K >= 1, "power must be positive");
    if constexpr (K == 1) {
        return b;
    } else {
        return ipow_c<K / 2>(b) * ipow_c<K - K / 2>(b);
    }
}
// runtime (warp-uniform) power, k >= 1
__device__ __forceinline__ double ipow_rt(double b, int k) {
    double r = b;
    bool have = (k & 1);
    k >>= 1;
    while (k) {
        b *= b;
        if (k & 1) { r = have ? r * b : b; have = true; }
        k >>= 1;
    }
    return r;
}

template <int MODE>
__device__ __forceinline__ double base_pow(double d2) {
    if constexpr (MODE == MODE_RCP) return pow_m1(d2);
    else if constexpr (MODE == MODE_RSQ) return pow_m12(d2);
    else return pow_m14(d2);
}

template <int E, int K0, int DK, int I>
__device__ __forceinline__ void static_powers(double b, double (&h)[E]) {
    if constexpr (I < E) {
        h[I] = ipow_c<K0 + I * DK>(b);
        static_powers<E, K0, DK, I + 1>(b, h);
    }
}

// h[e] = d2^(-p_e/2) for all exponents of the pass.
template <int MODE, int E, int K0, int DK>
__device__ __forceinline__ void weights(double d2, const KParams &p, const GenTables *gt, double (&h)[E]) {
    if constexpr (MODE == MODE_GEN) {
        constexpr int XM = K0 == GEN_SINGLE ? EXP_SINGLE : (K0 == GEN_SAFE ? EXP_BOUNDED : EXP_ANY);
        if constexpr (E == 1 && K0 != GEN_ANY) {
            // folded form: the CTA's log table and p.lg_s carry the factor 512 c, the log's last FMA delivers T
            h[0] = tab_exp2<XM>(log2_core(d2, gt->lg, p.lg_s[0], p.lg_s[1], p.lg_s[2], p.cexp[0]), gt->ex);
        } else {
            // In the `any` form L is finite (saturating log2), so p == 0 gives 0 * L == 0 and exp2(0) == 1 exactly:
            // Go's Pow(d, -0) == 1 even at d == 0 needs no special case; d == 0 with p > 0 -> +inf -> NaN cell,
            // p < 0 -> 0.
            const double L = (K0 == GEN_ANY) ? tab_log2_any(d2, gt->lg) : tab_log2_safe(d2, gt->lg);
#pragma unroll
            for (int e = 0; e < E; ++e) h[e] = tab_exp2<XM>(p.cexp[e] * L, gt->ex);
        }
    } else {
        double b = base_pow<MODE>(d2);
        if constexpr (K0 > 0) {
            static_powers<E, K0, (DK > 0 ? DK : 1), 0>(b, h);
        } else {
            h[0] = ipow_rt(b, p.k0);
            if constexpr (E > 1) {
                const double m = ipow_rt(b, p.dk);
#pragma unroll
                for (int e = 1; e < E; ++e) h[e] = h[e - 1] * m;
            }
        }
    }
}

// Tile guard of the paired kernels.  Pairing two points under one reciprocal needs the product of their squared
// distances to stay a normal double for EVERY cell of the CTA tile.  A point is unsafe for the tile when some cell
// is closer than `tiny` to it in both axes (or it coincides with a node: d2 == 0 must yield NaN through the
// one-reciprocal code) or farther than `far` in either axis; NaN coordinates are unsafe too.
// Cost matters (two points per thread and point tile): a dozen FP64 compares decide every point that is not within
// one cell of the tile's bounding box; only those few go through the exact nearest-node distances.
struct TileBox {
    double xlo, xhi, ylo, yhi;  // coordinates of the tile's extreme nodes
    double wx, wy;              // extent + two cells: a point with |x-xlo| + |x-xhi| <= wx lies within one cell of the tile
    double c0, r0;              // first column / row of the tile as doubles
};
template <int W, int H>
__device__ __forceinline__ TileBox tile_box(const KParams &p, long long tc0, long long tr0) {
    TileBox b;
    b.c0 = (double)tc0;
    b.r0 = (double)tr0;
    b.xlo = __dadd_rn(p.x_min, __dmul_rn(b.c0, p.x_step));
    b.xhi = __dadd_rn(p.x_min, __dmul_rn(b.c0 + (W - 1), p.x_step));
    b.ylo = __dadd_rn(p.y_min, __dmul_rn(b.r0, p.y_step));
    b.yhi = __dadd_rn(p.y_min, __dmul_rn(b.r0 + (H - 1), p.y_step));
    b.wx = fabs(b.xhi - b.xlo) + 2.0 * fabs(p.x_step);
    b.wy = fabs(b.yhi - b.ylo) + 2.0 * fabs(p.y_step);
    return b;
}
template <int W, int H>
__device__ __forceinline__ bool pairing_unsafe(double x, double y, const KParams &p, const TileBox &b, double tiny, double far) {
    const double ax0 = fabs(x - b.xlo), ax1 = fabs(x - b.xhi), ay0 = fabs(y - b.ylo), ay1 = fabs(y - b.yhi);
    if (!(ax0 <= far && ax1 <= far && ay0 <= far && ay1 <= far)) return true;  // too far, or NaN
    if (!(ax0 + ax1 <= b.wx && ay0 + ay1 <= b.wy)) return false;               // more than a cell away from every node
    // near the tile (rare): exact distance to the nearest node column / row; +-1 covers the rounding of the division
    auto nearest = [](double v, double lo, double vmin, double step, double inv_step, double first, int count) {
        int k = __double2int_rn((v - lo) * inv_step);
        k = min(max(k, 0), count - 1);
        double dmin = INFINITY;
#pragma unroll
        for (int d = -1; d <= 1; ++d) {
            const int kk = min(max(k + d, 0), count - 1);
            dmin = fmin(dmin, fabs(v - __dadd_rn(vmin, __dmul_rn(first + (double)kk, step))));
        }
        return dmin;
    };
    const double dx = nearest(x, b.xlo, p.x_min, p.x_step, p.inv_x_step, b.c0, W);
    const double dy = nearest(y, b.ylo, p.y_min, p.y_step, p.inv_y_step, b.r0, H);
    return !(dx >= tiny || dy >= tiny);
}

// ---------------------------------------------------------------------------- FP64 body
//
// PAIRED (p == 2 only): two points share ONE reciprocal, 1/a + 1/b = (a+b)/(ab) and
// w0/a + w1/b = (w0 b + w1 a)/(ab): 11 FP64 instructions per two pairs instead of 12, and half the MUFU
// traffic.  Needs a*b to stay a normal double; point tiles for which that cannot be guaranteed on this CTA's cells
// (pairing_unsafe) run through the one-reciprocal loop instead.
//
// YSP > 0 (staged dy^2): the squared row distances depend on (point, tile row) only, yet every one of the 32 lanes of a
// warp would compute them.  Instead the CTA computes them ONCE per chunk of YSP points into shared memory
// ([point][row of the tile]; same two instructions, so the values are bit-identical) and the point loops read
// their CR rows back with one broadcast load: 2/CC fewer FP64 instructions per pair (p = 2, 4x4 cells: 6.5 -> 6.0).

// points per staged chunk: the largest of 256 / 128 / 64 that still fits the shared-memory budget of MINB CTAs per SM
__host__ __device__ constexpr int ys_points(int mode, int E, int CR, int CC, int threads, int minb) {
    const int rows = (threads / 32) * CR;
    const int fixed = kStageBytes + kCtrlBytes + (mode == MODE_GEN ? (int)sizeof(GenTables) : 0) + 2 * E * CR * CC * threads * 8;
    const int budget = (227 * 1024) / minb - 1024;
    for (int pts = 256; pts >= 64; pts /= 2)
        if (fixed + pts * rows * 8 <= budget) return pts;
    return 0;
}

template <int MODE, int E, int CR, int CC, int K0, int DK, int THREADS_, bool PAIRED, int YSP, bool QUAD = false>
struct Fp64Body {
    static constexpr int THREADS = THREADS_;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int H = WARPS * CR;          // rows of the CTA tile
    static constexpr int NACC = 2 * E * CR * CC;  // acc[((e*CR + a)*CC + b)*2 + {0: num, 1: den}]
    static constexpr int TAB_BYTES = (MODE == MODE_GEN) ? (int)sizeof(GenTables) : 0;
    static constexpr int BODY_BYTES = TAB_BYTES + YSP * H * (int)sizeof(double);
    static_assert(YSP == 0 || (YSP % 4 == 0 && THREADS % H == 0 && CR % 2 == 0), "staging layout");
    static_assert(!QUAD || (PAIRED && YSP > 0), "four points per reciprocal: a flavour of the paired, staged body");

    double acc[NACC];
    double cx[CC], cy[CR];
    int tile_r, tile_c;         // tile indices of the CTA tile
    const GenTables *gt;
    double *ys;                 // staged squared row distances [point][tile row]

    __device__ __forceinline__ Fp64Body(const KParams &p, unsigned char *smem) : gt(nullptr), ys(nullptr) {
        if constexpr (YSP > 0) ys = reinterpret_cast<double *>(smem + kStageBytes + kCtrlBytes + TAB_BYTES);
        if constexpr (MODE == MODE_GEN) {
            // copy the table image global -> shared; single-exponent safe passes fold 512 c into log2_c
            GenTables *t = reinterpret_cast<GenTables *>(smem + kStageBytes + kCtrlBytes);
            constexpr bool FOLD = (E == 1 && K0 != GEN_ANY);
            for (int i = threadIdx.x; i < kLogTab; i += THREADS) {
                LogEntry e = p.tables->lg[i];
                if constexpr (FOLD) e.log2_c = p.cexp[0] * e.log2_c;
                t->lg[i] = e;
            }
            for (int j = threadIdx.x; j < kExpTab; j += THREADS)
                t->ex[j] = (K0 == GEN_SINGLE) ? exp_entry_single(p.tables->ex[j], j) : p.tables->ex[j];
            gt = t;
        }
    }

    __device__ __forceinline__ long long row_base(const KParams &p) const { return p.row0 + (long long)tile_r * H + (threadIdx.x >> 5) * CR; }
    __device__ __forceinline__ long long col_base() const { return (long long)tile_c * (32 * CC) + (threadIdx.x & 31); }

    __device__ __forceinline__ void setup(const KParams &p, int tr, int tc) {
        tile_r = tr;
        tile_c = tc;
        const long long c_base = col_base(), r_base = row_base(p);
        // RCToPoint (tools/utils.go:78-82): px = xMin + float64(c)*xStep, multiply then add, unfused.
#pragma unroll
        for (int b = 0; b < CC; ++b) cx[b] = __dadd_rn(p.x_min, __dmul_rn((double)(c_base + 32 * b), p.x_step));
#pragma unroll
        for (int a = 0; a < CR; ++a) cy[a] = __dadd_rn(p.y_min, __dmul_rn((double)(r_base + a), p.y_step));
    }

    // one cell-point pair per reciprocal power; yq = this warp's staged dy^2 of point i (YSP > 0) or unused
    __device__ __forceinline__ void single_points(const KParams &p, const double *sx, const double *sy, const double *sw, int i, int end,
                                                  const double *yq) {
#pragma unroll 1
        for (; i < end; ++i) {
            const double x = sx[i], w = sw[i];
            double dx2[CC], dy2[CR];
            // euclidDist(p, p0): point minus cell, each square rounded on its own (tools/utils.go:157)
#pragma unroll
            for (int b = 0; b < CC; ++b) { double d = __dsub_rn(x, cx[b]); dx2[b] = __dmul_rn(d, d); }
            if constexpr (YSP > 0) {
#pragma unroll
                for (int a = 0; a < CR; a += 2) {
                    const double2 u = *reinterpret_cast<const double2 *>(yq + a);
                    dy2[a] = u.x; dy2[a + 1] = u.y;
                }
                yq += H;
            } else {
                const double y = sy[i];
#pragma unroll
                for (int a = 0; a < CR; ++a) { double d = __dsub_rn(y, cy[a]); dy2[a] = __dmul_rn(d, d); }
            }
#pragma unroll
            for (int a = 0; a < CR; ++a)
#pragma unroll
                for (int b = 0; b < CC; ++b) {
                    const double d2 = __dadd_rn(dx2[b], dy2[a]);
                    double h[E];
                    weights<MODE, E, K0, DK>(d2, p, gt, h);
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        double &num = acc[((e * CR + a) * CC + b) * 2], &den = acc[((e * CR + a) * CC + b) * 2 + 1];
                        num = fma(w, h[e], num);
                        den += h[e];
                    }
                }
        }
    }

    // two points per reciprocal (p = 2); [i, end) holds an even number of points
    __device__ __forceinline__ void paired_points(const double *sx, const double *sy, const double *sw, int i, int end, const double *yq) {
#pragma unroll 1
        for (; i < end; i += 2) {
            const double2 x = *reinterpret_cast<const double2 *>(sx + i);
            const double2 w = *reinterpret_cast<const double2 *>(sw + i);
            double ax[CC], bx[CC], ay[CR], by[CR];
#pragma unroll
            for (int b = 0; b < CC; ++b) {
                double d0 = __dsub_rn(x.x, cx[b]), d1 = __dsub_rn(x.y, cx[b]);
                ax[b] = __dmul_rn(d0, d0);
                bx[b] = __dmul_rn(d1, d1);
            }
            if constexpr (YSP > 0) {
#pragma unroll
                for (int a = 0; a < CR; a += 2) {
                    const double2 u = *reinterpret_cast<const double2 *>(yq + a), v = *reinterpret_cast<const double2 *>(yq + H + a);
                    ay[a] = u.x; ay[a + 1] = u.y;
                    by[a] = v.x; by[a + 1] = v.y;
                }
                yq += 2 * H;
            } else {
                const double2 y = *reinterpret_cast<const double2 *>(sy + i);
#pragma unroll
                for (int a = 0; a < CR; ++a) {
                    double d0 = __dsub_rn(y.x, cy[a]), d1 = __dsub_rn(y.y, cy[a]);
                    ay[a] = __dmul_rn(d0, d0);
                    by[a] = __dmul_rn(d1, d1);
                }
            }
#pragma unroll
            for (int a = 0; a < CR; ++a)
#pragma unroll
                for (int b = 0; b < CC; ++b) {
                    const double A = __dadd_rn(ax[b], ay[a]), B = __dadd_rn(bx[b], by[a]);
                    const double r = pow_m1(A * B);
                    const double N = fma(w.x, B, w.y * A);
                    double &num = acc[(a * CC + b) * 2], &den = acc[(a * CC + b) * 2 + 1];
                    num = fma(N, r, num);
                    den = fma(A + B, r, den);
                }
        }
    }

    // four points per reciprocal (p = 2, staged dy^2 only; rcp_shared4, idw_math.cuh); [i, end) holds a multiple of four
    // points: the same 22 FP64 instructions per four pairs as two paired steps, half the MUFU.RCP64H and seed moves
    __device__ __forceinline__ void quad_points(const double *sx, const double *sw, int i, int end, const double *yq) {
#pragma unroll 1
        for (; i < end; i += 4) {
            const double2 x01 = *reinterpret_cast<const double2 *>(sx + i), x23 = *reinterpret_cast<const double2 *>(sx + i + 2);
            const double2 w01 = *reinterpret_cast<const double2 *>(sw + i), w23 = *reinterpret_cast<const double2 *>(sw + i + 2);
            double qx[4][CC];
#pragma unroll
            for (int b = 0; b < CC; ++b) {
                const double d0 = __dsub_rn(x01.x, cx[b]), d1 = __dsub_rn(x01.y, cx[b]);
                const double d2 = __dsub_rn(x23.x, cx[b]), d3 = __dsub_rn(x23.y, cx[b]);
                qx[0][b] = __dmul_rn(d0, d0);
                qx[1][b] = __dmul_rn(d1, d1);
                qx[2][b] = __dmul_rn(d2, d2);
                qx[3][b] = __dmul_rn(d3, d3);
            }
#pragma unroll
            for (int a = 0; a < CR; a += 2) {
                const double2 y0 = *reinterpret_cast<const double2 *>(yq + a), y1 = *reinterpret_cast<const double2 *>(yq + H + a);
                const double2 y2 = *reinterpret_cast<const double2 *>(yq + 2 * H + a), y3 = *reinterpret_cast<const double2 *>(yq + 3 * H + a);
#pragma unroll
                for (int aa = 0; aa < 2; ++aa)
#pragma unroll
                    for (int b = 0; b < CC; ++b) {
                        const double A = __dadd_rn(qx[0][b], aa ? y0.y : y0.x), B = __dadd_rn(qx[1][b], aa ? y1.y : y1.x);
                        const double Cc = __dadd_rn(qx[2][b], aa ? y2.y : y2.x), D = __dadd_rn(qx[3][b], aa ? y3.y : y3.x);
                        double SN, SD, r;
                        rcp_shared4(A, B, Cc, D, w01.x, w01.y, w23.x, w23.y, SN, SD, r);
                        double &num = acc[((a + aa) * CC + b) * 2], &den = acc[((a + aa) * CC + b) * 2 + 1];
                        num = fma(SN, r, num);
                        den = fma(SD, r, den);
                    }
            }
            yq += 4 * H;
        }
    }

    __device__ __forceinline__ void points(const KParams &p, const double *sx, const double *sy, const double *sw, int cnt,
                                           Ctrl *ctrl, unsigned it) {
        bool paired_ok = false;
        if constexpr (PAIRED) {
            static_assert(MODE == MODE_RCP && E == 1 && K0 == 1, "pairing is the p = 2 trick");
            const int slot = it & 1;
            bool unsafe = false;
            const TileBox box = tile_box<32 * CC, H>(p, (long long)tile_c * (32 * CC), p.row0 + (long long)tile_r * H);
            // two points: a b normal for d2 in (2^-500, 2^499); four points: ab cd and the cross products (a+b) cd,
            // (w0 b + w1 a) cd need d2 in (2^-240, 2^121) -- beyond 2^60 grid units a tile takes the one-reciprocal loop
            constexpr double kTiny = QUAD ? 0x1p-120 : 0x1p-250, kFar = QUAD ? 0x1p+60 : 0x1p+249;
            for (int k = threadIdx.x; k < cnt; k += THREADS)
                unsafe |= pairing_unsafe<32 * CC, H>(sx[k], sy[k], p, box, kTiny, kFar);
            if (unsafe) ctrl->guard[slot] = 1;
            __syncthreads();
            paired_ok = !ctrl->guard[slot];
            if (threadIdx.x == 0) ctrl->guard[slot ^ 1] = 0;  // nobody touches the other slot until the next tile
        }
        if constexpr (YSP > 0) {
            const int srow = threadIdx.x % H, sfirst = threadIdx.x / H;
            // this thread's row in the staging pass (RCToPoint, tools/utils.go:78-82)
            const double cyr = __dadd_rn(p.y_min, __dmul_rn((double)(p.row0 + (long long)tile_r * H + srow), p.y_step));
            const double *yq = ys + (threadIdx.x >> 5) * CR;
            for (int h0 = 0; h0 < cnt; h0 += YSP) {
                const int hn = min(YSP, cnt - h0);
                if (h0) __syncthreads();  // the previous chunk has been consumed
                for (int k = sfirst; k < hn; k += THREADS / H) {
                    const double d = __dsub_rn(sy[h0 + k], cyr);
                    ys[k * H + srow] = __dmul_rn(d, d);
                }
                __syncthreads();
                int i = h0;
                if constexpr (PAIRED) {
                    if (paired_ok) {
                        if constexpr (QUAD) {
                            quad_points(sx, sw, h0, h0 + (hn & ~3), yq);
                            i = h0 + (hn & ~3);
                        } else {
                            paired_points(sx, sy, sw, h0, h0 + (hn & ~1), yq);
                            i = h0 + (hn & ~1);
                        }
                    }
                }
                single_points(p, sx, sy, sw, i, h0 + hn, yq + (long long)(i - h0) * H);
            }
        } else {
            int i = 0;
            if constexpr (PAIRED) {
                if (paired_ok) {
                    paired_points(sx, sy, sw, 0, cnt & ~1, nullptr);
                    i = cnt & ~1;
                }
            }
            single_points(p, sx, sy, sw, i, cnt, nullptr);
        }
    }

    __device__ __forceinline__ void store(const KParams &p) {
        const long long row_end = p.row0 + p.nrows, r_base = row_base(p), c_base = col_base();
#pragma unroll
        for (int a = 0; a < CR; ++a) {
            const long long r = r_base + a;
            if (r >= row_end) continue;
#pragma unroll
            for (int b = 0; b < CC; ++b) {
                const long long c = c_base + 32 * b;
                if (c >= p.cols) continue;
#pragma unroll
                for (int e = 0; e < E; ++e)
                    p.out[(long long)e * p.plane + (r - p.row0) * p.cols + c] =
                        acc[((e * CR + a) * CC + b) * 2] / acc[((e * CR + a) * CC + b) * 2 + 1];
            }
        }
    }
};

// OPT: bit 0 = paired reciprocal (p = 2), bit 1 = dy^2 staged in shared memory, bit 2 = four points per reciprocal
template <int MODE, int E, int CR, int CC, int K0, int DK, int THREADS, int MINB, int OPT>
using Fp64BodyFor = Fp64Body<MODE, E, CR, CC, K0, DK, THREADS, (OPT & 1) != 0, (OPT & 2) ? ys_points(MODE, E, CR, CC, THREADS, MINB) : 0, (OPT & 4) != 0>;

template <int MODE, int E, int CR, int CC, int K0, int DK, int THREADS, int MINB, int OPT = 0>
__global__ void __launch_bounds__(THREADS, MINB) idw_fp64_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    Fp64BodyFor<MODE, E, CR, CC, K0, DK, THREADS, MINB, OPT> body(p, smem);
    tile_driver(p, body, smem);
}

}  // namespace v2r
