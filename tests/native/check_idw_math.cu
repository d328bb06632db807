// Host-side accuracy check of v2r_b200/csrc/idw_math.cuh (the general-exponent class): compiled with nvcc as a
// plain host program by tests/test_math_cpu.py.  Prints "max_abs_log2 max_rel_exp2 max_rel_pow bad_specials".
#include <cmath>
#include <cstdio>
#include <random>

#include "idw_math.cuh"
using namespace v2r;

int main() {
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> U(0, 1);
    double maxl = 0, maxe = 0, maxh = 0;
    for (long i = 0; i < 2000000; i++) {
        double a = std::exp2(U(g) * 200 - 100) * (1 + U(g));  // squared distances over 60 decades
        double el = fabs((double)(fast_log2(a) - log2l((long double)a)));
        if (el > maxl) maxl = el;
        double t = (U(g) * 2 - 1) * 60;
        long double et = exp2l((long double)t);
        double re = fabs((double)((fast_exp2(t) - et) / et));
        if (re > maxe) maxe = re;
        double c = -(U(g) * 5);  // c = -p/2 for p in (0, 10)
        long double ht = powl((long double)a, (long double)c);
        double rh = fabs((double)((fast_exp2(c * fast_log2(a)) - ht) / ht));
        if (rh > maxh) maxh = rh;
    }
    int bad = 0;
    // saturating specials: log2(0) is a huge negative FINITE number, inf / NaN give NaN
    for (double a0 : {0.0, -0.0, 1e-320}) bad += !(fast_log2(a0) < -1e300 && std::isfinite(fast_log2(a0)));
    bad += !std::isnan(fast_log2(INFINITY)) + !std::isnan(fast_log2(NAN));
    bad += !(fast_log2(1.0) == 0.0) + !(fast_log2(2.0) == 1.0) + !(fast_log2(0.5) == -1.0);
    const double ts[] = {0, -0.0, 1, -1, 1022, 1023, 1023.9, 1024, 1100, -1021, -3000, 3000, 2039.5, -2039.5, 2040, -2040, 2047.9, 1e300, -1e300};
    for (double t : ts) bad += !(fast_exp2(t) == exp2(t));
    for (double t : {-1022.5, -1040.0, -1074.0, -1075.0}) bad += !(fabs(fast_exp2(t) - exp2(t)) <= 1e-323);   // through the subnormals (+-2 quanta)
    bad += !std::isnan(fast_exp2(NAN));
    // the compositions the kernel relies on: d2 = 0 with p > 0 -> +inf, p < 0 -> 0, p == 0 -> exactly 1
    bad += !(fast_exp2(-0.85 * fast_log2(0.0)) == INFINITY) + !(fast_exp2(0.5 * fast_log2(0.0)) == 0.0) + !(fast_exp2(0.0 * fast_log2(0.0)) == 1.0);
    bad += !(fast_exp2(-0.0 * fast_log2(123.0)) == 1.0) + !std::isnan(fast_exp2(-0.85 * fast_log2(NAN)));
    printf("%.3e %.3e %.3e %d\n", maxl, maxe, maxh, bad);
    return 0;
}
