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
    bad += !(fast_log2(0.0) == -INFINITY) + !(fast_log2(-0.0) == -INFINITY) + !(fast_log2(INFINITY) == INFINITY) + !std::isnan(fast_log2(NAN));
    bad += !(fast_log2(1.0) == 0.0) + !(fast_log2(2.0) == 1.0) + !(fast_log2(0.5) == -1.0) + !(fast_log2(1e-320) == -INFINITY);
    const double ts[] = {0, 1, -1, 1023, 1024, 1100, -1022, -1074, -1075, -3000, 3000, INFINITY, -INFINITY, 2039.5, -2039.5, 2040, -2040, 2047.9};
    for (double t : ts) bad += !(fast_exp2(t) == exp2(t));
    bad += !std::isnan(fast_exp2(NAN));
    printf("%.3e %.3e %.3e %d\n", maxl, maxe, maxh, bad);
    return 0;
}
