// idw_kernels_fp32.cuh -- the optional FP32 fast path of the IDW kernels (precision = V2R_IDW_FP32).
//
// Same tile driver and TMA point pipeline as the FP64 kernels (idw_kernels.cuh); what changes:
//   * coordinates become CTA-local: every staged FP64 tile is re-expressed relative to the centre of the CTA's cell
//     tile (one FP64 subtract per point and CTA) and split into hi + lo floats, so dx keeps ~2^-24 RELATIVE accuracy
//     even at EPSG:2284 magnitudes (1.2e7 ft, where a plain float has a 1 ft ulp);
//   * distances, powers and the per-tile partial sums are FP32; every 512-point tile the partial sums are promoted
//     into FP64 accumulators, so the error does not grow with the number of points.
// p = 2 (the headline exponent) runs the PACKED body: sm_100's two-wide FP32 instructions (fma.rn.f32x2 / add / mul,
// SASS FFMA2 / FADD2 / FMUL2) over pairs of adjacent-in-register columns.  The packed forms do not raise the FP32
// pipe's flop rate (measured: FFMA2 issues at half the FFMA rate, profiles/r2_ubench.txt) -- they halve the ISSUE
// slots, and the round-1 kernel was issue-bound (7.5 instructions per pair).  To make every operand naturally
// two-wide:
//   - coordinates are measured in CELLS from the tile centre (x' = (x - ox) / x_step in FP64 per point and CTA), so
//     the cell offsets are small integers, exact in FP32, and need no lo part; y' is scaled back by y_step / x_step
//     when the steps differ.  Weights then carry a constant factor x_step^2 that cancels in num / den;
//   - per-point values are stored duplicated (v, v) in shared memory, so one LDS.64 delivers a two-wide operand;
//   - the squared ROW distances depend on (point, tile row) only: the CTA computes them once per chunk of points into
//     shared memory, duplicated, instead of every warp recomputing them with half-wasted two-wide instructions;
//   - two points still share ONE MUFU.RCP: 1/a + 1/b = (a+b)/(ab), w0/a + w1/b = (w0 b + w1 a)/(ab);
//     per two cells x two points: 8 packed instructions + 2 MUFU (round 1: 18 scalar + 2 MUFU).
//   A point tile that is not safe for pairing on this CTA's cells (a*b could leave the FP32 normal range: distances
//   below 2^-30 or above 2^29 cells) runs through the one-reciprocal packed loop.  Cells that coincide exactly with a
//   non-key point (reference: NaN) are poisoned afterwards by idw_poison_kernel, which evaluates the coincidence in
//   FP64 exactly as the reference does.
// Other exponents use the scalar body: MUFU.RCP / MUFU.RSQ ladders or ex2(c * lg2 d2).
// Stated accuracy: ~1e-5 relative (tests use 2e-5 on the scale sum|w|h / sum h; bench.py reports the observed maximum).
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

// ---------------------------------------------------------------------------- scalar body (all exponent classes)
template <int MODE, int E, int CR, int CC, int THREADS_>
struct Fp32Body {
    static constexpr int THREADS = THREADS_;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int NACC = 2 * E * CR * CC;
    static constexpr int T = kTilePoints;
    static constexpr int BODY_BYTES = 5 * T * (int)sizeof(float);  // fx hi/lo, fy hi/lo, fw

    double acc[NACC];
    float cxh[CC], cxl[CC], cyh[CR], cyl[CR];
    int tile_r, tile_c;
    double ox, oy;  // CTA-local origin = centre node of the CTA tile
    float *fxh, *fxl, *fyh, *fyl, *fw;

    __device__ __forceinline__ Fp32Body(const KParams &, unsigned char *smem) {
        float *f = reinterpret_cast<float *>(smem + kStageBytes + kCtrlBytes);
        fxh = f; fxl = f + T; fyh = f + 2 * T; fyl = f + 3 * T; fw = f + 4 * T;
    }

    __device__ __forceinline__ void setup(const KParams &p, int tr, int tc) {
        tile_r = tr;
        tile_c = tc;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const long long r0 = p.row0 + (long long)tile_r * (WARPS * CR), c0 = (long long)tile_c * (32 * CC);
        const long long c_base = c0 + lane, r_base = r0 + (long long)warp * CR;
        // origin computed like RCToPoint (tools/utils.go:78-82)
        ox = __dadd_rn(p.x_min, __dmul_rn((double)(c0 + 16 * CC), p.x_step));
        oy = __dadd_rn(p.y_min, __dmul_rn((double)(r0 + (WARPS * CR) / 2), p.y_step));
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
    }

    __device__ __forceinline__ void points(const KParams &p, const double *sx, const double *sy, const double *sw, int cnt,
                                           Ctrl *, unsigned) {
        // FP64 tile -> CTA-local hi/lo floats (the previous tile's readers left at the driver's trailing barrier)
        for (int i = threadIdx.x; i < cnt; i += THREADS) {
            double vx = sx[i] - ox, vy = sy[i] - oy;
            float hx = (float)vx, hy = (float)vy;
            fxh[i] = hx;
            fxl[i] = (float)(vx - (double)hx);
            fyh[i] = hy;
            fyl[i] = (float)(vy - (double)hy);
            fw[i] = (float)sw[i];
        }
        __syncthreads();
        float pn[E][CR][CC], pd[E][CR][CC];
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int a = 0; a < CR; ++a)
#pragma unroll
                for (int b = 0; b < CC; ++b) { pn[e][a][b] = 0.f; pd[e][a][b] = 0.f; }
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            const float xh = fxh[i], xl = fxl[i], yh = fyh[i], yl = fyl[i], w = fw[i];
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
                for (int b = 0; b < CC; ++b) {
                    acc[((e * CR + a) * CC + b) * 2] += (double)pn[e][a][b];
                    acc[((e * CR + a) * CC + b) * 2 + 1] += (double)pd[e][a][b];
                }
    }

    __device__ __forceinline__ void store(const KParams &p) {
        const long long row_end = p.row0 + p.nrows;
        const long long r_base = p.row0 + (long long)tile_r * (WARPS * CR) + (threadIdx.x >> 5) * CR;
        const long long c_base = (long long)tile_c * (32 * CC) + (threadIdx.x & 31);
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

template <int MODE, int E, int CR, int CC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) idw_fp32_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    Fp32Body<MODE, E, CR, CC, THREADS> body(p, smem);
    tile_driver(p, body, smem);
}

// ---------------------------------------------------------------------------- packed body (p = 2)
__device__ __forceinline__ float2 f2dup(float v) { return make_float2(v, v); }

// points per staged chunk of squared row distances ((v, v) float2 per point and tile row): fits MINB CTAs per SM
__host__ __device__ constexpr int ys32_points(int CR, int CC, int threads, int minb) {
    const int rows = (threads / 32) * CR;
    const int fixed = kStageBytes + kCtrlBytes + 5 * kTilePoints * 8 + 2 * CR * CC * threads * 8;
    const int budget = (227 * 1024) / minb - 1024;
    for (int pts = 256; pts >= 32; pts /= 2)
        if (fixed + pts * rows * 8 <= budget) return pts;
    return 0;
}

template <int CR, int CC, bool ISO, int THREADS_, int YSP>
struct Fp32PackedBody {
    static_assert(CC % 2 == 0, "columns are processed in pairs");
    static_assert(YSP > 0 && YSP % 2 == 0, "the squared row distances are staged in shared memory");
    static constexpr int THREADS = THREADS_;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int CP = CC / 2;                 // column pairs per thread
    static constexpr int NACC = 2 * CR * CC;          // acc[(a*CC + b)*2 + {0: num, 1: den}]
    static constexpr int T = kTilePoints;
    static constexpr int W = 32 * CC, H = WARPS * CR; // CTA tile in cells
    static_assert(THREADS % H == 0, "staging layout");
    // (v, v) pairs: x hi, x lo, w; y hi / lo as plain floats (only the staging pass reads them); staged dy^2 chunk
    static constexpr int BODY_BYTES = 3 * T * (int)sizeof(float2) + 2 * T * (int)sizeof(float) + YSP * H * (int)sizeof(float2);

    double acc[NACC];
    float2 nkx[CP];   // minus the column offsets (cells from the tile centre) of columns (2j, 2j+1)
    int tile_r, tile_c;
    double ox, oy;    // centre node of the CTA tile
    float2 *fxh, *fxl, *fw, *ys;
    float *fyh, *fyl;

    __device__ __forceinline__ Fp32PackedBody(const KParams &, unsigned char *smem) {
        float2 *f = reinterpret_cast<float2 *>(smem + kStageBytes + kCtrlBytes);
        fxh = f; fxl = f + T; fw = f + 2 * T;
        fyh = reinterpret_cast<float *>(f + 3 * T);
        fyl = fyh + T;
        ys = reinterpret_cast<float2 *>(fyl + T);
    }

    __device__ __forceinline__ void setup(const KParams &p, int tr, int tc) {
        tile_r = tr;
        tile_c = tc;
        const int lane = threadIdx.x & 31;
        const long long r0 = p.row0 + (long long)tile_r * H, c0 = (long long)tile_c * W;
        ox = __dadd_rn(p.x_min, __dmul_rn((double)(c0 + W / 2), p.x_step));
        oy = __dadd_rn(p.y_min, __dmul_rn((double)(r0 + H / 2), p.y_step));
#pragma unroll
        for (int j = 0; j < CP; ++j)
            nkx[j] = make_float2(-(float)(lane + 64 * j - W / 2), -(float)(lane + 64 * j + 32 - W / 2));
    }

    __device__ __forceinline__ void points(const KParams &p, const double *sx, const double *sy, const double *sw, int cnt,
                                           Ctrl *ctrl, unsigned it) {
        const int slot = it & 1;
        const double rho = ISO ? 1.0 : p.y_step * p.inv_x_step;  // y cells -> x cells
        // FP64 tile -> cell units from the tile centre, hi + lo floats (x and w stored duplicated); and the pairing guard
        bool unsafe = false;
        for (int i = threadIdx.x; i < cnt; i += THREADS) {
            const double vx = (sx[i] - ox) * p.inv_x_step, vy = (sy[i] - oy) * p.inv_y_step;
            const float hx = (float)vx, hy = (float)vy;
            fxh[i] = f2dup(hx);
            fxl[i] = f2dup((float)(vx - (double)hx));
            fyh[i] = hy;
            fyl[i] = (float)(vy - (double)hy);
            fw[i] = f2dup((float)sw[i]);
            // nearest / farthest cell of the tile along each axis (cells sit at the integers -W/2 .. W/2-1)
            const double kx = fmin(fmax(rint(vx), -(double)(W / 2)), (double)(W / 2 - 1));
            const double ky = fmin(fmax(rint(vy), -(double)(H / 2)), (double)(H / 2 - 1));
            const double nx = fabs(vx - kx), ny = fabs(vy - ky) * rho;
            const double fx = fmax(fabs(vx + W / 2), fabs(vx - (W / 2 - 1))), fy = fmax(fabs(vy + H / 2), fabs(vy - (H / 2 - 1))) * rho;
            const bool safe = (nx >= 0x1p-30 || ny >= 0x1p-30) && fx <= 0x1p+29 && fy <= 0x1p+29;
            unsafe |= !safe;
        }
        if (unsafe) ctrl->guard[slot] = 1;
        __syncthreads();
        const bool paired_ok = !ctrl->guard[slot];
        if (threadIdx.x == 0) ctrl->guard[slot ^ 1] = 0;

        float2 pn[CR][CP], pd[CR][CP];
#pragma unroll
        for (int a = 0; a < CR; ++a)
#pragma unroll
            for (int j = 0; j < CP; ++j) { pn[a][j] = make_float2(0.f, 0.f); pd[a][j] = make_float2(0.f, 0.f); }

        auto col_d = [&](float2 xh, float2 xl, int j) { return __fadd2_rn(__fadd2_rn(xh, nkx[j]), xl); };
        // staging pass: this thread's tile row (cells from the centre) and the squared row distance, duplicated
        const int srow = threadIdx.x % H, sfirst = threadIdx.x / H;
        const float nrow = -(float)(srow - H / 2), frho = (float)rho;
        const float2 *yq0 = ys + (threadIdx.x >> 5) * CR;

        for (int h0 = 0; h0 < cnt; h0 += YSP) {
            const int hn = min(YSP, cnt - h0);
            if (h0) __syncthreads();  // the previous chunk has been consumed
            for (int k = sfirst; k < hn; k += THREADS / H) {
                float d = (fyh[h0 + k] + nrow) + fyl[h0 + k];
                if constexpr (!ISO) d *= frho;
                ys[k * H + srow] = f2dup(d * d);
            }
            __syncthreads();
            const float2 *yq = yq0;
            int i = h0;
            if (paired_ok) {
#pragma unroll 1
                for (; i + 1 < h0 + hn; i += 2, yq += 2 * H) {
                    const float4 xh = *reinterpret_cast<const float4 *>(fxh + i), xl = *reinterpret_cast<const float4 *>(fxl + i);
                    const float4 w = *reinterpret_cast<const float4 *>(fw + i);
                    const float2 w0 = make_float2(w.x, w.y), w1 = make_float2(w.z, w.w);
                    float2 dx0[CP], dx1[CP];
#pragma unroll
                    for (int j = 0; j < CP; ++j) {
                        dx0[j] = col_d(make_float2(xh.x, xh.y), make_float2(xl.x, xl.y), j);
                        dx1[j] = col_d(make_float2(xh.z, xh.w), make_float2(xl.z, xl.w), j);
                    }
#pragma unroll
                    for (int a = 0; a < CR; ++a) {
                        const float2 ay = yq[a], by = yq[H + a];
#pragma unroll
                        for (int j = 0; j < CP; ++j) {
                            const float2 A = __ffma2_rn(dx0[j], dx0[j], ay), B = __ffma2_rn(dx1[j], dx1[j], by);
                            const float2 P = __fmul2_rn(A, B);
                            const float2 r = make_float2(frcp(P.x), frcp(P.y));
                            const float2 N = __ffma2_rn(w0, B, __fmul2_rn(w1, A));
                            pn[a][j] = __ffma2_rn(N, r, pn[a][j]);
                            pd[a][j] = __ffma2_rn(__fadd2_rn(A, B), r, pd[a][j]);
                        }
                    }
                }
            }
#pragma unroll 1
            for (; i < h0 + hn; ++i, yq += H) {  // one reciprocal per pair: odd tail, or a tile that is not safe for pairing
                const float2 xh = fxh[i], xl = fxl[i], w = fw[i];
#pragma unroll
                for (int j = 0; j < CP; ++j) {
                    const float2 dx = col_d(xh, xl, j);
#pragma unroll
                    for (int a = 0; a < CR; ++a) {
                        float2 A = __ffma2_rn(dx, dx, yq[a]);
                        // cell units lose offsets below ~1e-14 cells: a point that close (but not ON the node -- those
                        // cells are poisoned afterwards) must still dominate its cell with a finite weight; NaN passes
                        A.x = A.x < 0x1p-60f ? 0x1p-60f : A.x;
                        A.y = A.y < 0x1p-60f ? 0x1p-60f : A.y;
                        const float2 r = make_float2(frcp(A.x), frcp(A.y));
                        pn[a][j] = __ffma2_rn(w, r, pn[a][j]);
                        pd[a][j] = __fadd2_rn(pd[a][j], r);
                    }
                }
            }
        }
        // promote the tile's partial sums
#pragma unroll
        for (int a = 0; a < CR; ++a)
#pragma unroll
            for (int j = 0; j < CP; ++j) {
                acc[(a * CC + 2 * j) * 2] += (double)pn[a][j].x;
                acc[(a * CC + 2 * j) * 2 + 1] += (double)pd[a][j].x;
                acc[(a * CC + 2 * j + 1) * 2] += (double)pn[a][j].y;
                acc[(a * CC + 2 * j + 1) * 2 + 1] += (double)pd[a][j].y;
            }
    }

    __device__ __forceinline__ void store(const KParams &p) {
        const long long row_end = p.row0 + p.nrows;
        const long long r_base = p.row0 + (long long)tile_r * H + (threadIdx.x >> 5) * CR;
        const long long c_base = (long long)tile_c * W + (threadIdx.x & 31);
#pragma unroll
        for (int a = 0; a < CR; ++a) {
            const long long r = r_base + a;
            if (r >= row_end) continue;
#pragma unroll
            for (int b = 0; b < CC; ++b) {
                const long long c = c_base + 32 * b;
                if (c >= p.cols) continue;
                p.out[(r - p.row0) * p.cols + c] = acc[(a * CC + b) * 2] / acc[(a * CC + b) * 2 + 1];
            }
        }
    }
};

template <int CR, int CC, bool ISO, int THREADS, int MINB>
using Fp32PackedBodyFor = Fp32PackedBody<CR, CC, ISO, THREADS, ys32_points(CR, CC, THREADS, MINB)>;

template <int CR, int CC, bool ISO, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) idw_fp32_packed_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    Fp32PackedBodyFor<CR, CC, ISO, THREADS, MINB> body(p, smem);
    tile_driver(p, body, smem);
}

}  // namespace v2r
