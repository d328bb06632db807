// idw_kernels.cuh -- sm_100a kernels of the IDW gridding hot path (FP64 classes).
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
//     (cp.async.bulk + mbarrier complete_tx), STAGES deep; every thread reads the same point at the
//     same time, i.e. shared-memory broadcasts;
//   * per point and thread: CR squared dy and CC squared dx (reference: Pow(dx,2), Pow(dy,2),
//     tools/utils.go:157 -- both exactly rounded products), then per pair ONE add d2 = dx2 + dy2
//     (bit-identical to the reference's `total`), the power d2^(-p/2), one FMA and one ADD;
//   * d^-p never goes through sqrt+pow: it is rcp / rsqrt / fourth-root seeds from MUFU.RCP64H /
//     MUFU.RSQ64H refined by one cubic step in the FP64 pipe, or exp2(c*log2 d2) for general p.
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

constexpr int kMaxE = 10;        // exponents per pass (register budget); longer sweeps run several passes
constexpr int kTilePoints = 512; // points per shared-memory stage
constexpr int kStages = 3;

struct KParams {
    const double *x, *y, *w;  // device SoA, 16-byte aligned
    long long n;              // number of points
    double x_min, x_step, y_min, y_step;
    long long cols;           // grid columns
    long long row0, nrows;    // band [row0, row0 + nrows)
    double *out;              // E planes of nrows*cols
    long long plane;          // nrows * cols
    int k0, dk;               // integer power ladder: h_e = b^(k0 + e*dk) (runtime variant)
    double cexp[kMaxE];       // MODE_GEN: 32 * (-p_e/2): the exponent multiplier, pre-scaled for tab_exp2_32
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
// Shared-memory tables of the general class (2 KB + 256 B), built once per CTA.
struct GenTables {
    LogEntry lg[kLogTab];
    double ex[kExpTab];
};

template <int MODE, int E, int K0, int DK>
__device__ __forceinline__ void weights(double d2, const KParams &p, const GenTables *gt, double (&h)[E]) {
    if constexpr (MODE == MODE_GEN) {
        // K0 == 1 marks a pass whose exponents all satisfy 0 < p <= 3.75 (decided on the host): d2 = 0 may then be
        // poisoned to NaN straight away (the cell is NaN in the reference too) and |c * L| < 2040 needs no clamp.
        constexpr bool SAFE = (K0 == 1);
        const double L = tab_log2<SAFE>(d2, gt->lg);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            // cexp holds 32 * (-p/2).  In the general form L is finite (saturating log2), so p == 0 gives 0 * L == 0
            // and exp2(0) == 1 exactly: Go's Pow(d, -0) == 1 even at d == 0 needs no special case; d == 0 with
            // p > 0 -> +inf -> NaN cell, p < 0 -> 0.
            h[e] = tab_exp2_32<SAFE>(p.cexp[e] * L, gt->ex);
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

// ---------------------------------------------------------------------------- main kernel
//
// PAIRED (p == 2 only): two points share ONE reciprocal, 1/a + 1/b = (a+b)/(ab) and
// w0/a + w1/b = (w0 b + w1 a)/(ab): 11 FP64 instructions per two pairs instead of 12, and half the MUFU
// traffic.  Needs a*b to stay a normal double (squared distances in (1e-154, 1e154)).
template <int MODE, int E, int CR, int CC, int K0, int DK, int THREADS, int MINB, int PAIRED = 0>
__global__ void __launch_bounds__(THREADS, MINB) idw_fp64_kernel(const __grid_constant__ KParams p) {
    static_assert(!PAIRED || (MODE == MODE_RCP && E == 1 && K0 == 1), "pairing is the p = 2 trick");
    constexpr int UNR = PAIRED > 1 ? PAIRED : 1;  // PAIRED = unroll factor of the two-point loop
    constexpr int WARPS = THREADS / 32;
    constexpr int T = kTilePoints;
    constexpr int S = kStages;
    __shared__ __align__(128) double sm[S][3][T];
    __shared__ __align__(8) uint64_t full[S];
    const GenTables *gt = nullptr;
    if constexpr (MODE == MODE_GEN) {  // visible to all threads after the __syncthreads that follows the mbarrier init
        __shared__ __align__(16) GenTables gen_tables;
        for (int i = threadIdx.x; i < kLogTab; i += THREADS) build_log_entry(i, gen_tables.lg[i]);
        for (int j = threadIdx.x; j < kExpTab; j += THREADS) gen_tables.ex[j] = build_exp_entry(j);
        gt = &gen_tables;
    }

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long c_base = (long long)blockIdx.x * (32 * CC) + lane;
    const long long r_base = p.row0 + ((long long)blockIdx.y * WARPS + warp) * CR;

    // RCToPoint (tools/utils.go:78-82): px = xMin + float64(c)*xStep, multiply then add, unfused.
    double cx[CC], cy[CR];
#pragma unroll
    for (int b = 0; b < CC; ++b) cx[b] = __dadd_rn(p.x_min, __dmul_rn((double)(c_base + 32 * b), p.x_step));
#pragma unroll
    for (int a = 0; a < CR; ++a) cy[a] = __dadd_rn(p.y_min, __dmul_rn((double)(r_base + a), p.y_step));

    double num[E][CR][CC], den[E][CR][CC];
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int a = 0; a < CR; ++a)
#pragma unroll
            for (int b = 0; b < CC; ++b) { num[e][a][b] = 0.0; den[e][a][b] = 0.0; }

    const long long ntiles = (p.n + T - 1) / T;
    const long long nfull = p.n / T;  // tiles fetched by TMA; a ragged last tile is loaded by the threads

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](long long t) {
        const int s = (int)(t % S);
        mbar_expect_tx(&full[s], 3u * T * (uint32_t)sizeof(double));
        tma_load_1d(&sm[s][0][0], p.x + t * T, T * (uint32_t)sizeof(double), &full[s]);
        tma_load_1d(&sm[s][1][0], p.y + t * T, T * (uint32_t)sizeof(double), &full[s]);
        tma_load_1d(&sm[s][2][0], p.w + t * T, T * (uint32_t)sizeof(double), &full[s]);
    };
    if (threadIdx.x == 0) {
        for (long long t = 0; t < S - 1 && t < nfull; ++t) issue(t);
    }

    for (long long t = 0; t < ntiles; ++t) {
        const int s = (int)(t % S);
        int cnt = T;
        // stage (t-1)%S was released by the __syncthreads that ended iteration t-1
        if (threadIdx.x == 0 && t + S - 1 < nfull) issue(t + S - 1);
        if (t < nfull) {
            mbar_wait(&full[s], (uint32_t)((t / S) & 1));
        } else {
            cnt = (int)(p.n - t * T);
            for (int i = threadIdx.x; i < cnt; i += THREADS) {
                sm[s][0][i] = p.x[t * T + i];
                sm[s][1][i] = p.y[t * T + i];
                sm[s][2][i] = p.w[t * T + i];
            }
            __syncthreads();
        }
        const double *sx = sm[s][0], *sy = sm[s][1], *sw = sm[s][2];
        int i = 0;
        if constexpr (PAIRED) {
#pragma unroll UNR
            for (; i + 1 < cnt; i += 2) {
                const double2 x = *reinterpret_cast<const double2 *>(sx + i);
                const double2 y = *reinterpret_cast<const double2 *>(sy + i);
                const double2 w = *reinterpret_cast<const double2 *>(sw + i);
                double ax[CC], bx[CC], ay[CR], by[CR];
#pragma unroll
                for (int b = 0; b < CC; ++b) {
                    double d0 = __dsub_rn(x.x, cx[b]), d1 = __dsub_rn(x.y, cx[b]);
                    ax[b] = __dmul_rn(d0, d0);
                    bx[b] = __dmul_rn(d1, d1);
                }
#pragma unroll
                for (int a = 0; a < CR; ++a) {
                    double d0 = __dsub_rn(y.x, cy[a]), d1 = __dsub_rn(y.y, cy[a]);
                    ay[a] = __dmul_rn(d0, d0);
                    by[a] = __dmul_rn(d1, d1);
                }
#pragma unroll
                for (int a = 0; a < CR; ++a)
#pragma unroll
                    for (int b = 0; b < CC; ++b) {
                        const double A = __dadd_rn(ax[b], ay[a]), B = __dadd_rn(bx[b], by[a]);
                        const double r = pow_m1(A * B);
                        const double N = fma(w.x, B, w.y * A);
                        num[0][a][b] = fma(N, r, num[0][a][b]);
                        den[0][a][b] = fma(A + B, r, den[0][a][b]);
                    }
            }
        }
#pragma unroll 1
        for (; i < cnt; ++i) {
            const double x = sx[i], y = sy[i], w = sw[i];
            double dx2[CC], dy2[CR];
            // euclidDist(p, p0): point minus cell, each square rounded on its own (tools/utils.go:157)
#pragma unroll
            for (int b = 0; b < CC; ++b) { double d = __dsub_rn(x, cx[b]); dx2[b] = __dmul_rn(d, d); }
#pragma unroll
            for (int a = 0; a < CR; ++a) { double d = __dsub_rn(y, cy[a]); dy2[a] = __dmul_rn(d, d); }
#pragma unroll
            for (int a = 0; a < CR; ++a)
#pragma unroll
                for (int b = 0; b < CC; ++b) {
                    const double d2 = __dadd_rn(dx2[b], dy2[a]);
                    double h[E];
                    weights<MODE, E, K0, DK>(d2, p, gt, h);
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        num[e][a][b] = fma(w, h[e], num[e][a][b]);
                        den[e][a][b] += h[e];
                    }
                }
        }
        __syncthreads();
    }

    const long long row_end = p.row0 + p.nrows;
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
                p.out[(long long)e * p.plane + (r - p.row0) * p.cols + c] = num[e][a][b] / den[e][a][b];
        }
    }
}

}  // namespace v2r
