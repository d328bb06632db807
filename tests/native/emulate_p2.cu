// Host-side restatement of the NUMERICAL SCHEME of the p = 2 FP64 kernels (v2r_b200/csrc/idw_kernels.cuh), built from the
// same __host__ __device__ arithmetic (idw_math.cuh; MUFU seeds emulated at their 20-bit worst case): per cell, the points
// are consumed in 512-point tiles grouped into K <= 16 segments (plan_grid, idw_launch.cu), inside a tile in groups of
// 4 / 2 / 1 points per reciprocal (Fp64Body::quad_points / paired_points / single_points), the segment sums are folded in
// segment order ((S0 + S1) + S2) + ..., and the cell is num / den.  tests/test_kernel_numerics_cpu.py compares the
// result with the CPU oracle at the parity bar -- GPU-free evidence that the scheme itself (not one lucky data set) meets
// 1e-12, for every grouping, on clustered points and signed weights.  This is test code: it is not the kernel.
//
// With a third argument the same scheme runs for any exponent p with one point per base power, through the class the
// dispatcher would pick (plan_sweep, idw_launch.cu): rcp (p even), rsqrt (p integer), root4 (2p integer), general
// (folded table log2 / exp2 for 0 < p <= 3.75, saturating forms otherwise; p == 0 gives h = 1).
//
//   emulate_p2 <in.bin> <out.bin> [p]
//   in : int64 n, m, group; double x_min, x_step, y_min, y_step; double x[n], y[n], w[n]; int64 r[m], c[m]
//   out: double z[m]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#include "idw_math.cuh"
using namespace v2r;

static const int kTilePoints = 512, kMaxSegs = 16, kChunk = 256;  // kTilePoints / kMaxSegs / ys_points() of the 4x4 p = 2 build

static double ipow_rt(double b, int k) {   // the kernels' runtime ladder (idw_kernels.cuh)
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

// h(d2) = d2^(-p/2) the way the class of p computes it
static std::function<double(double)> weight_of(double p) {
    const bool half = p > 0 && std::floor(2 * p) == 2 * p && p <= 32;
    if (half && std::floor(p / 2) == p / 2) { const int k = (int)(p / 2); return [k](double a) { return ipow_rt(pow_m1(a), k); }; }
    if (half && std::floor(p) == p) { const int k = (int)p; return [k](double a) { return ipow_rt(pow_m12(a), k); }; }
    if (half) { const int k = (int)(2 * p); return [k](double a) { return ipow_rt(pow_m14(a), k); }; }
    static GenTables gt, ft;
    static double exs[kExpTab];
    build_gen_tables(gt);
    const double cs = kExpScale * (-p / 2.0);
    if (p > 0 && p <= 3.75) {   // GEN_SINGLE / GEN_SAFE: the multiplier folded into the log table and series (Fp64Body constructor)
        ft = gt;
        for (int i = 0; i < kLogTab; ++i) ft.lg[i].log2_c = cs * gt.lg[i].log2_c;
        for (int j = 0; j < kExpTab; ++j) exs[j] = exp_entry_single(gt.ex[j], j);
        const double l0 = cs * kLg[0], l1 = cs * kLg[1], l2 = cs * kLg[2];
        const bool single = p <= 1.99;
        return [=](double a) {
            const double T = log2_core(a, ft.lg, l0, l1, l2, cs);
            return single ? tab_exp2<EXP_SINGLE>(T, exs) : tab_exp2<EXP_BOUNDED>(T, ft.ex);
        };
    }
    return [cs](double a) { return tab_exp2<EXP_ANY>(cs * tab_log2_any(a, gt.lg), gt.ex); };   // GEN_ANY
}

int main(int argc, char **argv) {
    if (argc != 3 && argc != 4) return 2;
    const bool any_p = argc == 4;
    std::function<double(double)> hfun = [](double a) { return pow_m1(a); };
    if (any_p) hfun = weight_of(atof(argv[3]));
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int64_t n, m, group;
    double g[4];
    if (fread(&n, 8, 1, f) != 1 || fread(&m, 8, 1, f) != 1 || fread(&group, 8, 1, f) != 1 || fread(g, 8, 4, f) != 4) return 2;
    std::vector<double> x(n), y(n), w(n), z(m);
    std::vector<int64_t> r(m), c(m);
    if (fread(x.data(), 8, n, f) != (size_t)n || fread(y.data(), 8, n, f) != (size_t)n || fread(w.data(), 8, n, f) != (size_t)n ||
        fread(r.data(), 8, m, f) != (size_t)m || fread(c.data(), 8, m, f) != (size_t)m) return 2;
    fclose(f);
    const double x_min = g[0], x_step = g[1], y_min = g[2], y_step = g[3];
    // plan_grid: segments are a function of n only
    const int64_t ntp = (n + kTilePoints - 1) / kTilePoints;
    int64_t K = ntp < kMaxSegs ? ntp : kMaxSegs;
    if (K < 1) K = 1;
    int64_t seg_tiles = (ntp + K - 1) / K;
    if (seg_tiles < 1) seg_tiles = 1;
    K = (ntp + seg_tiles - 1) / seg_tiles;
    if (K < 1) K = 1;
    for (int64_t q = 0; q < m; ++q) {
        // RCToPoint: multiply, then add, each rounded (tools/utils.go:78-82)
        volatile double cxm = (double)c[q] * x_step, cym = (double)r[q] * y_step;
        const double cx = x_min + cxm, cy = y_min + cym;
        double tnum = 0, tden = 0;
        for (int64_t seg = 0; seg < K; ++seg) {
            double num = 0, den = 0;
            const int64_t t_end = (seg + 1) * seg_tiles < ntp ? (seg + 1) * seg_tiles : ntp;
            for (int64_t t = seg * seg_tiles; t < t_end; ++t) {
                const int64_t base = t * kTilePoints, cnt = (n - base) < kTilePoints ? (n - base) : kTilePoints;
                for (int64_t h0 = 0; h0 < cnt; h0 += kChunk) {
                    const int64_t hn = (cnt - h0) < kChunk ? (cnt - h0) : kChunk;
                    auto d2 = [&](int64_t i) {
                        volatile double dx = x[base + i] - cx, dy = y[base + i] - cy;
                        volatile double dx2 = dx * dx, dy2 = dy * dy;   // each square rounded on its own (tools/utils.go:157)
                        return dx2 + dy2;
                    };
                    int64_t i = h0;
                    if (any_p) {
                    } else if (group == 4) {
                        for (; i + 4 <= h0 + (hn & ~3LL); i += 4) {
                            double sn, sd, rr;
                            rcp_shared4(d2(i), d2(i + 1), d2(i + 2), d2(i + 3), w[base + i], w[base + i + 1], w[base + i + 2], w[base + i + 3],
                                        sn, sd, rr);
                            num = fma(sn, rr, num);
                            den = fma(sd, rr, den);
                        }
                    } else if (group == 2) {
                        for (; i + 2 <= h0 + (hn & ~1LL); i += 2) {
                            const double A = d2(i), B = d2(i + 1);
                            const double rr = pow_m1(A * B);
                            const double N = fma(w[base + i], B, w[base + i + 1] * A);
                            num = fma(N, rr, num);
                            den = fma(A + B, rr, den);
                        }
                    }
                    for (; i < h0 + hn; ++i) {
                        const double h = hfun(d2(i));
                        num = fma(w[base + i], h, num);
                        den += h;
                    }
                }
            }
            if (seg == 0) { tnum = num; tden = den; }
            else { tnum += num; tden += den; }
        }
        z[q] = tnum / tden;
    }
    f = fopen(argv[2], "wb");
    if (!f || fwrite(z.data(), 8, m, f) != (size_t)m) return 2;
    fclose(f);
    return 0;
}
