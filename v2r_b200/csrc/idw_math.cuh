// idw_math.cuh -- the arithmetic of d2^(-p/2): seed-plus-one-cubic-step base powers of the rcp / rsqrt / root4 classes
// (bottom of the file) and software log2 / exp2 for the `general` class (h = exp2(c * log2 d2)).
//
// The FP64 pipe is the bottleneck of every IDW kernel, so these are written for the fewest FP64-pipe
// instructions that still give ~1e-15: table-driven argument reduction (tables live in shared memory, built once
// per CTA), short series with coefficients in the constant bank, and all range reduction, special-case handling
// and 2^n scaling on the integer pipe (no FP64 compares, no branches).
//   tab_log2    : 8 FP64 + I2F + one 16-byte table load  (128 entries {1/c_i, -log2(1/c_i)}, degree-6 series)
//   tab_exp2_32 : 11 FP64 + one 8-byte table load        (32 entries 2^(j/32), degree-6 series)
// Both SATURATE instead of producing infinities from finite-range inputs (log2(0) = -2^1000-ish), which removes
// every special case from the kernel: c*L is never inf, and 0*L == 0 gives exp2(0) == 1 exactly (Go: Pow(d,-0) == 1
// even at d == 0).  Subnormal d2 is flushed to zero like rcp.approx.ftz does in the other classes.
// History (c2 band, p = 1.7, G evals/s): libdevice log2()/exp2() 218; table-free FDLIBM-style versions (16 + 17
// FP64 instructions) 302, then 367 once their coefficients sat in the constant bank and the per-cell branches
// were gone; the table-driven versions here need 19 instead of 33.
// __host__ versions exist only so tests/ can compile this header on the CPU and compare with libm.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace v2r {

#ifdef __CUDA_ARCH__
#define V2R_COEF __constant__
#else
#define V2R_COEF static const
#endif
// log2(1+r) = r * (kLg[0] + kLg[1] r + ... + kLg[5] r^5),  kLg[k] = log2(e) * (-1)^k / (k+1);  |r| <= 2^-8
V2R_COEF double kLg[6] = {1.4426950408889634074, -0.72134752044448170368, 0.48089834696298780245,
                          -0.36067376022224085184, 0.28853900817779268147, -0.24044917348149390123};
// 2^(g/32) = 1 + g * (kEx[0] + kEx[1] g + ... + kEx[5] g^5),  kEx[k] = (ln2/32)^(k+1) / (k+1)!;  |g| <= 1/2
V2R_COEF double kEx[6] = {0.021660849392498290919, 0.00023459619820224678939, 1.6938509724371820054e-6,
                          9.1725627018246432895e-9, 3.9737099845494161318e-11, 1.4345655584131935569e-13};
V2R_COEF double kMagic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer

struct LogEntry { double inv_c, log2_c; };  // 1/c_i rounded to double, and -log2 of that rounded value
constexpr int kLogTab = 128, kExpTab = 32;

__host__ __device__ __forceinline__ int hi_word(double a) {
#ifdef __CUDA_ARCH__
    return __double2hiint(a);
#else
    uint64_t u; memcpy(&u, &a, 8); return (int)(u >> 32);
#endif
}
__host__ __device__ __forceinline__ int lo_word(double a) {
#ifdef __CUDA_ARCH__
    return __double2loint(a);
#else
    uint64_t u; memcpy(&u, &a, 8); return (int)(u & 0xffffffffu);
#endif
}
__host__ __device__ __forceinline__ double make_double(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double a; memcpy(&a, &u, 8); return a;
#endif
}
__host__ __device__ __forceinline__ double int2double(int k) {
#ifdef __CUDA_ARCH__
    return __int2double_rn(k);
#else
    return (double)k;
#endif
}

// Entry i of the tables (the CTA builds them once, one entry per thread, before the point loop).
// c_i is the midpoint of the i-th 1/128 slice of [1,2); log2_c is taken from the ROUNDED reciprocal, so
// log2(m) = log2_c + log2(m * inv_c) holds exactly for the numbers actually used.
__host__ __device__ inline void build_log_entry(int i, LogEntry &e) {
    const double c = 1.0 + (i + 0.5) / kLogTab;
    e.inv_c = 1.0 / c;
    e.log2_c = -log2(e.inv_c);
}
__host__ __device__ inline double build_exp_entry(int j) { return exp2((double)j / kExpTab); }

// log2(a) for a >= 0, SATURATING: a == 0 (or subnormal / negative) -> about -2^1000 (finite), +inf / NaN -> NaN.
// POISON_ZERO (all exponents of the pass are > 0, where d2 = 0 makes the whole cell NaN anyway): zero, subnormal,
// inf and NaN all give NaN, found with ONE unsigned range test.
template <bool POISON_ZERO = false>
__host__ __device__ __forceinline__ double tab_log2(double a, const LogEntry *tab) {
    const int hi = hi_word(a);
    const int k = (hi >> 20) - 1023;
    const LogEntry t = tab[(hi >> 13) & (kLogTab - 1)];
    const double m = make_double((hi & 0x000fffff) | 0x3ff00000, lo_word(a));  // [1, 2)
    const double r = fma(m, t.inv_c, -1.0);                                    // |r| <= 2^-8
    double q = kLg[5];
    q = fma(q, r, kLg[4]);
    q = fma(q, r, kLg[3]);
    q = fma(q, r, kLg[2]);
    q = fma(q, r, kLg[1]);
    q = fma(q, r, kLg[0]);
    const double L = fma(q, r, int2double(k) + t.log2_c);
    int lhi = hi_word(L);
    if constexpr (POISON_ZERO) {
        if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) lhi = 0x7ff80000;  // not a positive normal number: NaN
    } else {
        if (hi < 0x00100000) lhi = (int)0xfe700000;   // zero / subnormal / negative: huge negative, finite
        if (hi >= 0x7ff00000) lhi = 0x7ff80000;       // +inf / NaN: NaN
    }
    return make_double(lhi, lo_word(L));
}

// 2^(T/32) for finite T (any magnitude) or NaN; the caller folds the factor 32 into its multiplier.
// Saturates: T/32 >= 1024 -> +inf, T/32 <= -1075 -> 0.
// BOUNDED: the caller guarantees |T/32| < 2040 or NaN (true when |p| <= 3.75: |log2 d2| <= 1075 for every positive
// normal d2, and poisoned logs are NaN), so the clamp is skipped.
template <bool BOUNDED = false>
__host__ __device__ __forceinline__ double tab_exp2_32(double T, const double *tab) {
    double Tc = T;
    if constexpr (!BOUNDED) {
        // clamp 65280 <= |T| < inf (i.e. |T/32| >= 2040) to +-65280 with one unsigned range test; NaN passes through
        const int hi = hi_word(T);
        const bool big = (unsigned)((hi & 0x7fffffff) - 0x40efe000) < (unsigned)(0x7ff00000 - 0x40efe000);
        Tc = make_double(big ? ((hi & (int)0x80000000) | 0x40efe000) : hi, big ? 0 : lo_word(T));
    }
    const double Tn = Tc + kMagic;
    const int N = lo_word(Tn);                // nearest integer to T (two's complement in the low word)
    const double g = Tc - (Tn - kMagic);      // in [-1/2, 1/2], exact
    double p = kEx[5];
    p = fma(p, g, kEx[4]);
    p = fma(p, g, kEx[3]);
    p = fma(p, g, kEx[2]);
    p = fma(p, g, kEx[1]);
    p = fma(p, g, kEx[0]);
    p = fma(p, g, 1.0);                       // 2^(g/32)
    const double e = tab[N & (kExpTab - 1)];  // 2^(j/32) in [1, 2)
    const int n = N >> 5;                     // floor(N / 32), |n| <= 2040
    const int n1 = n >> 1, n2 = n - n1;
    // 2^n in two factors, so overflow saturates to +inf and underflow rounds through the subnormals to 0 by the
    // multiplies themselves; the first factor rides on the table value's exponent field (an integer add)
    const double s1 = make_double(hi_word(e) + (n1 << 20), lo_word(e));
    return (p * s1) * make_double((n2 + 1023) << 20, 0);
}

// MUFU.RCP64H / MUFU.RSQ64H: ~2^-20 relative seeds computed from the high word of the operand.
// The host versions emulate the documented behaviour (operand and result truncated to their high words, i.e.
// 20 mantissa bits) so tests/ can check the refinement steps below against libm on the CPU.
__host__ __device__ __forceinline__ double rcp_seed(double a) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
#else
    return make_double(hi_word(1.0 / make_double(hi_word(a), 0)), 0);
#endif
}
__host__ __device__ __forceinline__ double rsq_seed(double a) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
#else
    return make_double(hi_word(1.0 / sqrt(make_double(hi_word(a), 0))), 0);
#endif
}

// ---------------------------------------------------------------------------- base powers
// 1/a: seed + one cubic step (e = 1 - a r0 ; r = r0 (1 + e + e^2)); truncation e^3 ~ 2^-60.
// 3 FP64-pipe instructions + 1 MUFU.  a = 0 -> NaN (reference: +Inf weight -> Inf/Inf = NaN cell).
__host__ __device__ __forceinline__ double pow_m1(double a) {
    double r0 = rcp_seed(a);
    double e = fma(-a, r0, 1.0);
    double t = fma(e, e, e);
    return fma(r0, t, r0);
}
// a^(-1/2): e = 1 - a y0^2 ; y = y0 (1 + e/2 + 3 e^2/8).  5 FP64 + 1 MUFU.
__host__ __device__ __forceinline__ double pow_m12(double a) {
    double y0 = rsq_seed(a);
    double t = a * y0;
    double e = fma(-t, y0, 1.0);
    double p = fma(e, 0.375, 0.5);
    double ye = y0 * e;
    return fma(ye, p, y0);
}
// a^(-1/4): q0 = rsq(a) * rsq(rsq(a)); e = 1 - a q0^4 ; q = q0 (1 + e/4 + 5 e^2/32).  7 FP64 + 2 MUFU.
__host__ __device__ __forceinline__ double pow_m14(double a) {
    double y0 = rsq_seed(a);   // a^-1/2 (low word 0, so the second seed sees it exactly)
    double z0 = rsq_seed(y0);  // a^+1/4
    double q0 = y0 * z0;       // a^-1/4, ~2^-20
    double q2 = q0 * q0;
    double t = a * q2;
    double e = fma(-t, q2, 1.0);
    double p = fma(e, 0.15625, 0.25);
    double qe = q0 * e;
    return fma(qe, p, q0);
}

}  // namespace v2r
