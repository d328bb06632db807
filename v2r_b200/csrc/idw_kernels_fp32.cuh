// idw_kernels_fp32.cuh -- the optional FP32 fast path of the IDW kernels (precision = V2R_IDW_FP32).
//
// Same tiling and TMA point pipeline as the FP64 kernels; what changes:
//   * coordinates become CTA-local: every staged FP64 tile is re-expressed relative to the centre of
//     the CTA's cell tile (one FP64 subtract per point and CTA) and split into hi + lo floats, so
//     dx = (xh - cxh) + (xl - cxl) keeps ~2^-24 RELATIVE accuracy even at EPSG:2284 magnitudes
//     (1.2e7 ft, where a plain float has a 1 ft ulp);
//   * distances, powers and the per-tile partial sums are FP32; every 512-point tile the partial sums
//     are promoted into FP64 accumulators, so the error does not grow with the number of points;
//   * p = 2 shares ONE MUFU.RCP between two points: 1/a + 1/b = (a+b)/(ab), w0/a + w1/b = (w0 b + w1 a)/(ab)
//     -- this halves MUFU pressure and makes the FP32 FMA pipe the bound instead of the MUFU pipe;
//   * other exponents use MUFU.RCP / MUFU.RSQ ladders or ex2(c * lg2 d2).
// Stated accuracy: ~1e-5 relative (tests use 2e-5 on the scale sum|w|h / sum h); distances must stay
// inside (3e-10, 4e9) units for the paired p = 2 kernel (product of two squared distances in FP32).
#pragma once
#include "idw_kernels.cuh"

namespace v2r {

__device__ __forceinline__ float frcp(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
__device__ __forceinline__ float frsq(float a) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
__device__ __forceinline__ float flg2(float a) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
__device__ __forceinline__ float fex2(float a) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

__device__ __forceinline__ float fipow_rt(float b, int k) {
    float r = b;
    bool have = (k & 1);
    k >>= 1;
    while (k) {
        b *= b;
        if (k & 1) { r = have ? r * b : b; have = true; }
        k >>= 1;
    }
    return r;
}

template <int MODE, int E>
__device__ __forceinline__ void fweights(float d2, const KParams &p, float (&h)[E]) {
    if constexpr (MODE == MODE_GEN) {
        float L = flg2(d2);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            float c = (float)p.cexp[e];
            h[e] = (c == 0.0f) ? 1.0f : fex2(c * L);
        }
    } else {
        float b;
        if constexpr (MODE == MODE_RCP) b = frcp(d2);
        else if constexpr (MODE == MODE_RSQ) b = frsq(d2);
        else { float y = frsq(d2); b = y * frsq(y); }
        h[0] = fipow_rt(b, p.k0);
        if constexpr (E > 1) {
            float m = fipow_rt(b, p.dk);
#pragma unroll
            for (int e = 1; e < E; ++e) h[e] = h[e - 1] * m;
        }
    }
}

// PAIRED: p == 2, E == 1 only.
template <int MODE, int E, int CR, int CC, bool PAIRED, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) idw_fp32_kernel(const __grid_constant__ KParams p) {
    static_assert(!PAIRED || (MODE == MODE_RCP && E == 1), "pairing is the p = 2 trick");
    constexpr int WARPS = THREADS / 32;
    constexpr int T = kTilePoints;
    constexpr int S = kStages;
    __shared__ __align__(128) double sm[S][3][T];
    __shared__ __align__(16) float fx[2][T], fy[2][T];  // [0] = hi, [1] = lo
    __shared__ __align__(16) float fw[T];
    __shared__ __align__(8) uint64_t full[S];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long c0 = (long long)blockIdx.x * (32 * CC);
    const long long r0 = p.row0 + (long long)blockIdx.y * (WARPS * CR);
    const long long c_base = c0 + lane;
    const long long r_base = r0 + (long long)warp * CR;
    // CTA-local origin = centre node of the CTA tile, computed like RCToPoint (tools/utils.go:78-82)
    const double ox = __dadd_rn(p.x_min, __dmul_rn((double)(c0 + 16 * CC), p.x_step));
    const double oy = __dadd_rn(p.y_min, __dmul_rn((double)(r0 + (WARPS * CR) / 2), p.y_step));

    float cxh[CC], cxl[CC], cyh[CR], cyl[CR];
#pragma unroll
    for (int b = 0; b < CC; ++b) {
        double v = __dadd_rn(p.x_min, __dmul_rn((double)(c_base + 32 * b), p.x_step)) - ox;
        cxh[b] = (float)v;
        cxl[b] = (float)(v - (double)cxh[b]);
    }
#pragma unroll
    for (int a = 0; a < CR; ++a) {
        double v = __dadd_rn(p.y_min, __dmul_rn((double)(r_base + a), p.y_step)) - oy;
        cyh[a] = (float)v;
        cyl[a] = (float)(v - (double)cyh[a]);
    }

    double num[E][CR][CC], den[E][CR][CC];
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int a = 0; a < CR; ++a)
#pragma unroll
            for (int b = 0; b < CC; ++b) { num[e][a][b] = 0.0; den[e][a][b] = 0.0; }

    const long long ntiles = (p.n + T - 1) / T;
    const long long nfull = p.n / T;
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
    if (threadIdx.x == 0)
        for (long long t = 0; t < S - 1 && t < nfull; ++t) issue(t);

    for (long long t = 0; t < ntiles; ++t) {
        const int s = (int)(t % S);
        int cnt = T;
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
        // FP64 tile -> CTA-local hi/lo floats (the previous tile's readers left at the trailing barrier)
        for (int i = threadIdx.x; i < cnt; i += THREADS) {
            double vx = sm[s][0][i] - ox, vy = sm[s][1][i] - oy;
            float hx = (float)vx, hy = (float)vy;
            fx[0][i] = hx;
            fx[1][i] = (float)(vx - (double)hx);
            fy[0][i] = hy;
            fy[1][i] = (float)(vy - (double)hy);
            fw[i] = (float)sm[s][2][i];
        }
        __syncthreads();

        float pn[E][CR][CC], pd[E][CR][CC];
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int a = 0; a < CR; ++a)
#pragma unroll
                for (int b = 0; b < CC; ++b) { pn[e][a][b] = 0.f; pd[e][a][b] = 0.f; }

        int i = 0;
        if constexpr (PAIRED) {
#pragma unroll 1
            for (; i + 1 < cnt; i += 2) {
                const float2 xh = *reinterpret_cast<const float2 *>(&fx[0][i]);
                const float2 xl = *reinterpret_cast<const float2 *>(&fx[1][i]);
                const float2 yh = *reinterpret_cast<const float2 *>(&fy[0][i]);
                const float2 yl = *reinterpret_cast<const float2 *>(&fy[1][i]);
                const float2 w = *reinterpret_cast<const float2 *>(&fw[i]);
                float ax[CC], bx[CC], ay[CR], by[CR];
#pragma unroll
                for (int b = 0; b < CC; ++b) {
                    float d0 = (xh.x - cxh[b]) + (xl.x - cxl[b]);
                    float d1 = (xh.y - cxh[b]) + (xl.y - cxl[b]);
                    ax[b] = d0 * d0;
                    bx[b] = d1 * d1;
                }
#pragma unroll
                for (int a = 0; a < CR; ++a) {
                    float d0 = (yh.x - cyh[a]) + (yl.x - cyl[a]);
                    float d1 = (yh.y - cyh[a]) + (yl.y - cyl[a]);
                    ay[a] = d0 * d0;
                    by[a] = d1 * d1;
                }
#pragma unroll
                for (int a = 0; a < CR; ++a)
#pragma unroll
                    for (int b = 0; b < CC; ++b) {
                        const float A = ax[b] + ay[a], B = bx[b] + by[a];
                        const float r = frcp(A * B);
                        const float N = fmaf(w.x, B, w.y * A);
                        pn[0][a][b] = fmaf(N, r, pn[0][a][b]);
                        pd[0][a][b] = fmaf(A + B, r, pd[0][a][b]);
                    }
            }
        }
#pragma unroll 1
        for (; i < cnt; ++i) {
            const float xh = fx[0][i], xl = fx[1][i], yh = fy[0][i], yl = fy[1][i], w = fw[i];
            float dx2[CC], dy2[CR];
#pragma unroll
            for (int b = 0; b < CC; ++b) { float d = (xh - cxh[b]) + (xl - cxl[b]); dx2[b] = d * d; }
#pragma unroll
            for (int a = 0; a < CR; ++a) { float d = (yh - cyh[a]) + (yl - cyl[a]); dy2[a] = d * d; }
#pragma unroll
            for (int a = 0; a < CR; ++a)
#pragma unroll
                for (int b = 0; b < CC; ++b) {
                    float h[E];
                    fweights<MODE, E>(dx2[b] + dy2[a], p, h);
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        pn[e][a][b] = fmaf(w, h[e], pn[e][a][b]);
                        pd[e][a][b] += h[e];
                    }
                }
        }
        // promote the tile's partial sums
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int a = 0; a < CR; ++a)
#pragma unroll
                for (int b = 0; b < CC; ++b) { num[e][a][b] += (double)pn[e][a][b]; den[e][a][b] += (double)pd[e][a][b]; }
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
