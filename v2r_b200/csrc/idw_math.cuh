// idw_math.cuh -- software log2 / exp2 for the `general` exponent class (h = exp2(c * log2 d2)).
//
// The FP64 pipe is the bottleneck of every IDW kernel, so these are written for the fewest FP64-pipe
// instructions that still give ~2 ulp: all range reduction, special-case handling and scaling is done with
// integer instructions on the high/low words (ALU pipe, otherwise idle), and there are no FP64 compares.
//   fast_log2 : 16 FP64 + 1 MUFU.RCP64H   (FDLIBM e_log.c style: s = f/(2+f), 7-term series in s^2)
//   fast_exp2 : 17 FP64                   (round-to-int by magic add, exp(x/2) Taylor degree 10, one squaring,
//                                          2^n as two exact power-of-two factors so the result saturates)
// Both SATURATE instead of producing infinities from finite-range inputs (log2(0) = -2^1000), which removes
// every special case from the kernel.  With libdevice's log2()/exp2() the kernel was 1.9x slower.
// Subnormal d2 is flushed to zero like rcp.approx.ftz does in the other classes.
// __host__ versions exist only so tests/ can compile this header on the CPU and compare with libm.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace v2r {

// Polynomial coefficients live in the constant bank on the device: an FP64 instruction can take a c[bank][off]
// operand directly, whereas 64-bit literals were being rebuilt with two UMOVs per use (71 per loop iteration,
// which made the general-class kernel issue-bound).
#ifdef __CUDA_ARCH__
#define V2R_COEF __constant__
#else
#define V2R_COEF static const
#endif
V2R_COEF double kLg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
                          1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01};  // Lg1..Lg7, FDLIBM e_log.c
V2R_COEF double kEx[9] = {2.7557319223985893e-07, 2.7557319223985888e-06, 2.4801587301587302e-05, 1.9841269841269841e-04,
                          1.3888888888888889e-03, 8.3333333333333332e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};  // 1/10! .. 1/2!
V2R_COEF double kMisc[3] = {1.4426950408889634074, 6755399441055744.0, 0.34657359027997265471};  // log2(e), 1.5*2^52, ln(2)/2

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
__host__ __device__ __forceinline__ double rcp20(double a) {  // ~2^-20 reciprocal seed
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
#else
    return make_double(hi_word(1.0 / make_double(hi_word(a), 0)), 0);
#endif
}
__host__ __device__ __forceinline__ double int2double(int k) {
#ifdef __CUDA_ARCH__
    return __int2double_rn(k);
#else
    return (double)k;
#endif
}

// log2(a) for a >= 0, SATURATING: a == 0 (or subnormal, flushed like rcp.approx.ftz) -> -2^1000, +inf / NaN -> NaN.
// The result is therefore always finite or NaN, so c * L is never inf and 0 * L == 0: the caller needs no special
// case for p == 0 (Go: Pow(d, -0) == 1 even at d == 0) and fast_exp2 needs no inf handling.
// 16 FP64 + MUFU.RCP64H + I2F; specials cost two integer compares and two selects on the HIGH word only.
__host__ __device__ __forceinline__ double fast_log2(double a) {
    const int hi = hi_word(a);
    const int lo = lo_word(a);
    // m in [sqrt(1/2), sqrt(2)):  mantissa above 0x6a09e (sqrt2) folds into the lower binade
    int k = (hi >> 20) - 1023;
    int mhi = (hi & 0x000fffff) | 0x3ff00000;
    if (mhi >= 0x3ff6a09f) { mhi -= 0x00100000; k += 1; }
    const double m = make_double(mhi, lo);
    const double f = m - 1.0;                 // exact
    const double t = 2.0 + f;
    double r = rcp20(t);                      // 1/(2+f): seed + cubic step
    const double e = fma(-t, r, 1.0);
    const double e2 = fma(e, e, e);
    r = fma(r, e2, r);
    const double s = f * r;
    const double z = s * s;
    double p = kLg[6];                        // Lg7 .. Lg1
    p = fma(p, z, kLg[5]);
    p = fma(p, z, kLg[4]);
    p = fma(p, z, kLg[3]);
    p = fma(p, z, kLg[2]);
    p = fma(p, z, kLg[1]);
    p = fma(p, z, kLg[0]);
    const double q = p * z;                   // log(m) = 2s + s*q
    const double lm = fma(s, q, s + s);
    const double L = fma(lm, kMisc[0], int2double(k));
    int lhi = hi_word(L);
    if (hi < 0x00100000) lhi = (int)0xfe700000;   // zero / subnormal (/ negative): -2^1000-ish, finite
    if (hi >= 0x7ff00000) lhi = 0x7ff80000;       // +inf / NaN: NaN
    return make_double(lhi, lo_word(L));
}

// 2^t for finite t (any magnitude) or NaN.  Saturates: t >= 1024 -> +inf, t <= -1075 -> 0.  17 FP64.
__host__ __device__ __forceinline__ double fast_exp2(double t) {
    const int hi = hi_word(t);
    // clamp 2040 <= |t| < inf to +-2040 with one unsigned range test (NaN / inf fall outside and pass through)
    const bool big = (unsigned)((hi & 0x7fffffff) - 0x409fe000) < (unsigned)(0x7ff00000 - 0x409fe000);
    const double tc = make_double(big ? ((hi & (int)0x80000000) | 0x409fe000) : hi, big ? 0 : lo_word(t));
    const double magic = kMisc[1];            // 1.5 * 2^52: adding it rounds to the nearest integer
    const double tn = tc + magic;
    const int n = lo_word(tn);                // integer part (two's complement in the low word)
    const double f = tc - (tn - magic);       // in [-1/2, 1/2]
    const double x = f * kMisc[2];            // (ln 2)/2 * f : exp(x)^2 = 2^f, |x| <= 0.1733
    double p = kEx[0];                        // Taylor 1/10! .. 1/2!
    p = fma(p, x, kEx[1]);
    p = fma(p, x, kEx[2]);
    p = fma(p, x, kEx[3]);
    p = fma(p, x, kEx[4]);
    p = fma(p, x, kEx[5]);
    p = fma(p, x, kEx[6]);
    p = fma(p, x, kEx[7]);
    p = fma(p, x, kEx[8]);
    p = fma(p, x, 1.0);
    p = fma(p, x, 1.0);
    p = p * p;
    // 2^n in two factors (n in [-2040, 2040], each half keeps a normal exponent field): overflow saturates to
    // +inf, underflow rounds through the subnormals to 0, NaN stays NaN -- all by the multiplies themselves
    const int n1 = n >> 1, n2 = n - n1;
    return (p * make_double((n1 + 1023) << 20, 0)) * make_double((n2 + 1023) << 20, 0);
}

}  // namespace v2r
