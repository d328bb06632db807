// Host-side accuracy check of v2r_b200/csrc/idw_math.cuh (the table-driven log2/exp2 of the general class): compiled with nvcc as a
// plain host program by tests/test_math_cpu.py.  Prints "max_abs_log2 max_rel_exp2 max_rel_pow max_rel_pow_fold bad_specials".
#include <cmath>
#include <cstdio>
#include <random>

#include "idw_math.cuh"
using namespace v2r;

int main() {
    static GenTables gt;
    build_gen_tables(gt);
    static double exs[kExpTab];   // the single-factor kernels keep the exp table in this form
    for (int j = 0; j < kExpTab; ++j) exs[j] = exp_entry_single(gt.ex[j], j);
    auto flog2 = [&](double a) { return tab_log2_any(a, gt.lg); };
    auto fexp2 = [&](double t) { return tab_exp2<EXP_ANY>(kExpScale * t, gt.ex); };
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> U(0, 1);
    double maxl = 0, maxe = 0, maxh = 0, maxf = 0;
    for (long i = 0; i < 2000000; i++) {
        double a = std::exp2(U(g) * 200 - 100) * (1 + U(g));  // squared distances over 60 decades
        double el = fabs((double)(flog2(a) - log2l((long double)a)));
        if (el > maxl) maxl = el;
        double t = (U(g) * 2 - 1) * 60;
        long double et = exp2l((long double)t);
        double re = fabs((double)((fexp2(t) - et) / et));
        if (re > maxe) maxe = re;
        double c = -(U(g) * 5);  // c = -p/2 for p in (0, 10)
        long double ht = powl((long double)a, (long double)c);
        double rh = fabs((double)((tab_exp2<EXP_ANY>((kExpScale * c) * flog2(a), gt.ex) - ht) / ht));
        if (rh > maxh) maxh = rh;
    }
    int bad = 0;
    // the folded single-exponent form of the safe passes: T = 512 c log2(d2) straight from the log's last FMA, with the
    // table's log2_c and the series coefficients pre-multiplied by 512 c; single-factor scaling for |c| <= 0.995
    for (double pexp : {1.7, 0.3, 0.30000000000000004, 1.99, 0.1, 1.1, 2.2, 3.3, 3.75}) {
        const double cs = kExpScale * (-pexp / 2.0);
        static GenTables ft;
        ft = gt;
        for (int i = 0; i < kLogTab; ++i) ft.lg[i].log2_c = cs * gt.lg[i].log2_c;
        const double l0 = cs * kLg[0], l1 = cs * kLg[1], l2 = cs * kLg[2];
        for (long i = 0; i < 400000; i++) {
            double a = std::exp2(U(g) * 200 - 100) * (1 + U(g));
            if (i % 5 == 0) a = std::exp2(U(g) * 2040 - 1020) * (1 + U(g));  // the whole normal range
            const double T = log2_core(a, ft.lg, l0, l1, l2, cs);
            const double h = pexp <= 1.99 ? tab_exp2<EXP_SINGLE>(T, exs) : tab_exp2<EXP_BOUNDED>(T, ft.ex);
            long double ht = powl((long double)a, (long double)(-pexp / 2.0));
            if ((double)ht == 0.0 || std::isinf((double)ht)) { bad += !(h == (double)ht); continue; }
            if ((double)ht < 1e-290) continue;   // through the subnormals: only the two-factor form reaches here
            double rf = fabs((double)((h - ht) / ht));
            if (rf > maxf) maxf = rf;
            if (pexp <= 1.99) bad += !(h == tab_exp2<EXP_BOUNDED>(T, ft.ex));   // the integer add equals the two multiplies
        }
    }
    // saturating specials: log2(0) is a huge negative FINITE number, inf / NaN give NaN
    for (double a0 : {0.0, -0.0, 1e-320}) bad += !(flog2(a0) < -1e300 && std::isfinite(flog2(a0)));
    bad += !std::isnan(flog2(INFINITY)) + !std::isnan(flog2(NAN));
    bad += !(fabs(flog2(1.0)) < 6e-15) + !(fabs(flog2(2.0) - 1.0) < 6e-15) + !(fabs(flog2(0.5) + 1.0) < 6e-15);   // the series' design error
    const double ts[] = {0, -0.0, 1, -1, 1022, 1023, 1024, 1100, -1021, -3000, 3000, 2039.5, -2039.5, 2040, -2040, 2047.9, 1e300, -1e300};
    for (double t : ts) bad += !(fexp2(t) == exp2(t));
    bad += !(fabs(fexp2(1023.9) / exp2(1023.9) - 1) < 4e-15);
    for (double t : {-1022.5, -1040.0, -1074.0, -1075.0}) bad += !(fabs(fexp2(t) - exp2(t)) <= 1e-323);   // through the subnormals
    bad += !std::isnan(fexp2(NAN));
    // the compositions the kernel relies on: d2 = 0 with p > 0 -> +inf, p < 0 -> 0, p == 0 -> exactly 1
    bad += !(fexp2(-0.85 * flog2(0.0)) == INFINITY) + !(fexp2(0.5 * flog2(0.0)) == 0.0) + !(fexp2(0.0 * flog2(0.0)) == 1.0);
    bad += !(fexp2(-0.0 * flog2(123.0)) == 1.0) + !std::isnan(fexp2(-0.85 * flog2(NAN)));
    // safe forms agree bit for bit with the saturating ones on positive normal inputs and |t| < 2040
    for (int i = 0; i < 200000; i++) {
        double a = std::exp2(U(g) * 2040 - 1020) * (1 + U(g));
        double L = tab_log2_safe(a, gt.lg);
        bad += !(L == tab_log2_any(a, gt.lg));
        double T = -kExpScale * 1.875 * U(g) * L;
        bad += !(tab_exp2<EXP_BOUNDED>(T, gt.ex) == tab_exp2<EXP_ANY>(T, gt.ex));
    }
    // +inf distance (a point at infinity): weight ~ 0 relative to any finite point, never NaN, in the safe forms
    bad += !(tab_exp2<EXP_SINGLE>(kExpScale * -0.85 * tab_log2_safe(INFINITY, gt.lg), exs) < 1e-250);
    // base powers of the rcp / rsqrt / root4 classes: a ~2^-20 seed (emulated on the host by truncating operand and
    // result to their high words, the documented behaviour of MUFU.RCP64H / RSQ64H) plus ONE cubic step must reach
    // a few ulp; worst-case seeds are covered by sampling mantissas whose low words are all ones
    double max1 = 0, max12 = 0, max14 = 0;
    for (long i = 0; i < 2000000; i++) {
        double a = std::exp2(U(g) * 400 - 200) * (1 + U(g));
        if (i % 3 == 0) a = make_double(hi_word(a), (int)0xffffffff);   // largest truncation error of the seed operand
        long double t1 = 1.0L / a, t12 = 1.0L / sqrtl(a), t14 = 1.0L / sqrtl(sqrtl(a));
        max1 = fmax(max1, fabs((double)((pow_m1(a) - t1) / t1)));
        max12 = fmax(max12, fabs((double)((pow_m12(a) - t12) / t12)));
        max14 = fmax(max14, fabs((double)((pow_m14(a) - t14) / t14)));
    }
    bad += !(max1 < 4e-16) + !(max12 < 6e-16) + !(max14 < 8e-16);
    bad += !std::isnan(pow_m1(0.0)) + !std::isnan(pow_m12(0.0)) + !std::isnan(pow_m14(0.0));   // d = 0 off-key -> NaN cell
    // p = 2 with a shared reciprocal: two pairs (the formula of Fp64Body::paired_points) and four pairs (rcp_shared4)
    // against the four separate quotients in long double; error relative to sum |w_i| / d2_i (the parity scale).  The
    // squared distances cover the whole guarded domain of the four-point form, (2^-240, 2^121), with independent
    // magnitudes, equal magnitudes and the corners.
    double max2n = 0, max2d = 0, max4n = 0, max4d = 0;
    for (long i = 0; i < 1000000; i++) {
        double v[4], w[4];
        const double common = U(g) * 360 - 240;
        for (int k = 0; k < 4; ++k) {
            const double e2 = (i % 3 == 0) ? common + U(g) : U(g) * 361 - 240;   // similar magnitudes, or anything
            v[k] = std::exp2(e2 < 120 ? e2 : 120.0) * (1 + U(g));
            if (i % 7 == 0) v[k] = make_double(hi_word(v[k]), (int)0xffffffff);
            w[k] = (U(g) * 2 - 1) * std::exp2(U(g) * 40 - 20);
        }
        if (i == 0) for (int k = 0; k < 4; ++k) v[k] = 0x1p-240;
        if (i == 1) for (int k = 0; k < 4; ++k) v[k] = 0x1p+121;
        if (i == 2) { v[0] = v[1] = 0x1p-240; v[2] = v[3] = 0x1p+121; }
        long double td = 0, tn = 0, ts = 0;
        for (int k = 0; k < 4; ++k) { td += 1.0L / v[k]; tn += (long double)w[k] / v[k]; ts += fabsl((long double)w[k]) / v[k]; }
        double sn, sd, r;
        rcp_shared4(v[0], v[1], v[2], v[3], w[0], w[1], w[2], w[3], sn, sd, r);
        bad += !(std::isfinite(sn * r) && std::isfinite(sd * r) && sd * r > 0);
        max4d = fmax(max4d, fabs((double)((sd * r - td) / td)));
        max4n = fmax(max4n, fabs((double)((sn * r - tn) / ts)));
        // two pairs: (a+b) / (ab), (w0 b + w1 a) / (ab)
        const double r2 = pow_m1(v[0] * v[1]);
        const long double td2 = 1.0L / v[0] + 1.0L / v[1], tn2 = (long double)w[0] / v[0] + (long double)w[1] / v[1];
        const long double ts2 = fabsl((long double)w[0]) / v[0] + fabsl((long double)w[1]) / v[1];
        max2d = fmax(max2d, fabs((double)(((v[0] + v[1]) * r2 - td2) / td2)));
        max2n = fmax(max2n, fabs((double)((fma(w[0], v[1], w[1] * v[0]) * r2 - tn2) / ts2)));
    }
    bad += !(max2d < 8e-16) + !(max2n < 8e-16) + !(max4d < 1.2e-15) + !(max4n < 1.2e-15);
    fprintf(stderr, "shared reciprocal: max rel err 2 pairs den %.2e num %.2e, 4 pairs den %.2e num %.2e\n", max2d, max2n, max4d, max4n);
    fprintf(stderr, "base powers: max rel err rcp %.2e rsqrt %.2e root4 %.2e\n", max1, max12, max14);
    printf("%.3e %.3e %.3e %.3e %d\n", maxl, maxe, maxh, maxf, bad);
    return 0;
}
