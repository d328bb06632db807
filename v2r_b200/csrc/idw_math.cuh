// idw_math.cuh -- the arithmetic of d2^(-p/2): seed-plus-one-cubic-step base powers of the rcp / rsqrt / root4 classes
// (bottom of the file) and software log2 / exp2 for the `general` class (h = exp2(c * log2 d2)).
//
// Measured on B200 (scripts/ubench.cu, profiles/r2_ubench.txt): the FP64 pipe issues 64 lanes/clk/SM, and so does the
// integer ALU pipe (LOP3 / SHF / IADD3) -- an integer instruction costs as much as a DFMA, I2F.F64 / F2I.F64 run at
// 16 lanes/clk/SM and stall the FP64 pipe for about one DFMA slot.  The general class is therefore written for the
// fewest FP64 *and* integer instructions per pair:
//   * log2: 1024-entry table {1/c_i, -log2(1/c_i)} (16 KB of shared memory per CTA, copied from a per-device global
//     image), |r| <= 2^-11, so a 3-coefficient series is enough (5e-15 absolute); the exponent is converted with the
//     2^52 magic-number add instead of I2F;
//   * exp2: 512-entry table 2^(j/512) (4 KB), 3-coefficient series (2.2e-15 relative); rounding to the nearest table
//     point with the 1.5*2^52 magic number;
//   * `fold` form for single-exponent passes: the multiplier c is folded into the log table and the series
//     coefficients when the CTA loads them, so T = 512 c log2(d2) comes out of the log's own final FMA;
//   * `single` scaling: when |c| <= 0.995 the result exponent stays normal for every positive normal d2, so 2^n is an
//     integer add on the table value's exponent field instead of two multiplies.
// FP64 instructions per pair: 6 (log2, +1 with the magic add) and 7 (exp2) against 8 + 11 in round 1.
// Specials: the `any` forms SATURATE instead of producing infinities from finite-range inputs (log2(0) ~ -2^1000),
// which removes every special case from the kernel: c*L is never inf, and 0*L == 0 gives exp2(0) == 1 exactly (Go:
// Pow(d,-0) == 1 even at d == 0).  The `safe` forms (all exponents > 0) have no special handling at all: d2 == 0 /
// subnormal cells are poisoned afterwards by idw_poison_kernel (idw_launch.cu).  Subnormal d2 counts as zero, like
// rcp.approx.ftz does in the other classes.
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
constexpr int kLogBits = 10, kExpBits = 9;
constexpr int kLogTab = 1 << kLogBits, kExpTab = 1 << kExpBits;
constexpr double kExpScale = (double)kExpTab;  // T = kExpScale * c * log2(d2)

// log2(1+r) = r * (kLg[0] + kLg[1] r + kLg[2] r^2), |r| <= 2^-11 (1 + 2^-20): interpolated at the Chebyshev nodes,
// maximum error 5.2e-15 in log2 units (scripts/fit_coefficients.py)
V2R_COEF double kLg[3] = {1.4426950408889634074, -0.72134758493810647712, 0.48089839855788806054};
// 2^(g/512) = 1 + g * (kEx[0] + kEx[1] g + kEx[2] g^2), |g| <= 1/2: maximum relative error 2.2e-15
V2R_COEF double kEx[3] = {0.0013538030870311431825, 9.1639142547048926913e-7, 4.1353784217323216102e-10};
V2R_COEF double kMagic = 6755399441055744.0;        // 1.5 * 2^52: adding it rounds to the nearest integer
V2R_COEF double kMagicExp = 4503599627371519.0;     // 2^52 + 1023: make_double(0x43300000, e) - this == e - 1023

struct LogEntry { double inv_c, log2_c; };  // 1/c rounded to double, and -log2 of that rounded value (or c times it: fold)
struct GenTables {
    LogEntry lg[kLogTab];
    double ex[kExpTab];
};

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

// Entry i of the tables.  c_i is the midpoint of the i-th 1/1024 slice of [1,2); log2_c is taken from the ROUNDED
// reciprocal, so log2(m) = log2_c + log2(m * inv_c) holds exactly for the numbers actually used.  Built once per
// device on the host (long double libm) and copied to global memory; every CTA copies the image to shared memory.
inline void build_gen_tables(GenTables &t) {
    for (int i = 0; i < kLogTab; ++i) {
        const long double c = 1.0L + (i + 0.5L) / kLogTab;
        t.lg[i].inv_c = (double)(1.0L / c);
        t.lg[i].log2_c = (double)(-log2l((long double)t.lg[i].inv_c));
    }
    for (int j = 0; j < kExpTab; ++j) t.ex[j] = (double)exp2l((long double)j / kExpTab);
}

// (hi & 0x000fffff) | 0x3ff00000.  With both constants as immediates the compiler needs two LOP3; with the constants in
// registers (hidden from constant propagation) it is ONE three-input LOP3 -- an integer instruction costs the loop as
// much as half an FP64 instruction.
__host__ __device__ __forceinline__ int mantissa_hi(int hi) {
#ifdef __CUDA_ARCH__
    int mask, one, r;
    asm("mov.b32 %0, 0x000fffff;" : "=r"(mask));
    asm("mov.b32 %0, 0x3ff00000;" : "=r"(one));
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(hi), "r"(mask), "r"(one));  // (a & b) | c
    return r;
#else
    return (hi & 0x000fffff) | 0x3ff00000;
#endif
}

// ---- log2 ------------------------------------------------------------------------------------------------
// Core: no special handling.  a must be a positive normal double for the result to mean anything.
// Returns  scale * log2(a)  where `scale` is 1 (plain table, lg = kLg) or 512 c (folded table: log2_c pre-multiplied,
// lg = 512 c kLg[], kscale = 512 c).  6 FP64-pipe instructions + 1 DADD for the exponent, 5 integer.
__host__ __device__ __forceinline__ double log2_core(double a, const LogEntry *tab, double lg0, double lg1, double lg2,
                                                     double kscale) {
    const int hi = hi_word(a);
    const LogEntry t = tab[(hi >> (20 - kLogBits)) & (kLogTab - 1)];
    const double m = make_double(mantissa_hi(hi), lo_word(a));                 // [1, 2)
    const double r = fma(m, t.inv_c, -1.0);                                    // |r| <= 2^-11
    const double kd = make_double(0x43300000, (int)((unsigned)hi >> 20)) - kMagicExp;  // unbiased exponent, exact
    double q = fma(lg2, r, lg1);
    q = fma(q, r, lg0);
    return fma(q, r, fma(kd, kscale, t.log2_c));
}

// log2(a), SATURATING: a == 0 (or subnormal / negative) -> about -2^1000 (finite), +inf / NaN -> NaN.
__host__ __device__ __forceinline__ double tab_log2_any(double a, const LogEntry *tab) {
    const double L = log2_core(a, tab, kLg[0], kLg[1], kLg[2], 1.0);
    const int hi = hi_word(a);
    int lhi = hi_word(L);
    if (hi < 0x00100000) lhi = (int)0xfe700000;   // zero / subnormal / negative: huge negative, finite
    if (hi >= 0x7ff00000) lhi = 0x7ff80000;       // +inf / NaN: NaN
    return make_double(lhi, lo_word(L));
}
// log2(a) for positive normal a (the `safe` passes): no special handling.
__host__ __device__ __forceinline__ double tab_log2_safe(double a, const LogEntry *tab) {
    return log2_core(a, tab, kLg[0], kLg[1], kLg[2], 1.0);
}

// Table value for the single-factor scaling: the high word of 2^(j/512) minus (j << 11), so that adding N * 2048
// (N = 512 n + j) to it yields hi + (n << 20) with ONE integer multiply-add and no masking of j.
__host__ __device__ __forceinline__ double exp_entry_single(double e, int j) {
    return make_double(hi_word(e) - (j << (20 - kExpBits)), lo_word(e));
}

// ---- exp2 ------------------------------------------------------------------------------------------------
// 2^(T/512).  Scaling of the result by 2^n:
//   EXP_ANY     any finite T or NaN: |T/512| >= 2040 is clamped, 2^n in two factors so overflow saturates to +inf and
//               underflow rounds through the subnormals to 0 by the multiplies themselves;
//   EXP_BOUNDED the caller guarantees |T/512| < 2040 (or NaN): no clamp, two factors;
//   EXP_SINGLE  the caller guarantees |T/512| <= 1021: n is added to the table value's exponent field by one integer
//               multiply-add; `tab` must hold exp_entry_single() values.
enum ExpMode : int { EXP_ANY = 0, EXP_BOUNDED = 1, EXP_SINGLE = 2 };

template <int XM>
__host__ __device__ __forceinline__ double tab_exp2(double T, const double *tab) {
    double Tc = T;
    if constexpr (XM == EXP_ANY) {
        // clamp 2040*512 <= |T| < inf to +-2040*512 with one unsigned range test; NaN passes through
        const int hi = hi_word(T);
        constexpr int kLim = 0x412fe000;  // high word of 2040 * 512 = 1044480.0
        const bool big = (unsigned)((hi & 0x7fffffff) - kLim) < (unsigned)(0x7ff00000 - kLim);
        Tc = make_double(big ? ((hi & (int)0x80000000) | kLim) : hi, big ? 0 : lo_word(T));
    }
    const double Tn = Tc + kMagic;
    const int N = lo_word(Tn);                // nearest integer to T (two's complement in the low word)
    const double g = Tc - (Tn - kMagic);      // in [-1/2, 1/2], exact
    double q = fma(kEx[2], g, kEx[1]);
    q = fma(q, g, kEx[0]);
    const double p = fma(q, g, 1.0);          // 2^(g/512)
    const double e = tab[N & (kExpTab - 1)];  // 2^(j/512) in [1, 2)
    if constexpr (XM == EXP_SINGLE) {
        // the table holds hi - (j << 11): hi' + N * 2048 = hi + (floor(N / 512) << 20)
        const double s = make_double(hi_word(e) + N * (1 << (20 - kExpBits)), lo_word(e));
        return p * s;
    } else {
        const int n = N >> kExpBits;          // floor(N / 512), |n| <= 2040
        const int n1 = n >> 1, n2 = n - n1;
        const double s1 = make_double(hi_word(e) + (n1 << 20), lo_word(e));
        return (p * s1) * make_double((n2 + 1023) << 20, 0);
    }
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
// FOUR pairs under one reciprocal (p = 2: h = 1/d2).  With a, b, c, d the four squared distances of a cell,
//   1/a + 1/b + 1/c + 1/d         = ((a+b) cd + (c+d) ab) / (ab cd)                      = sd * r
//   w0/a + w1/b + w2/c + w3/d     = ((w0 b + w1 a) cd + (w2 d + w3 c) ab) / (ab cd)      = sn * r ,   r = 1 / (ab cd)
// 20 FP64 instructions + the two accumulations = 22 per four pairs (as many as two two-point steps) and ONE MUFU.RCP64H.
// Needs ab cd and the cross products normal: every squared distance in (2^-240, 2^121) (the caller's tile guard).
__host__ __device__ __forceinline__ void rcp_shared4(double a, double b, double c, double d, double w0, double w1, double w2,
                                                     double w3, double &sn, double &sd, double &r) {
    const double p1 = a * b, p2 = c * d;
    r = pow_m1(p1 * p2);
    const double n1 = fma(w0, b, w1 * a), n2 = fma(w2, d, w3 * c);
    sn = fma(n1, p2, n2 * p1);
    sd = fma(a + b, p2, (c + d) * p1);
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
