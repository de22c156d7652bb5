// kem_math.cuh -- branch-free fp64 exp / reciprocal / division for the fused
// membrane kernel.
//
// Why not CUDA's libm exp() and operator/ ?  Both are <= 1 ulp and so are these,
// but the library versions carry a slow-path test (FSETP + BRA + BSSY/BSYNC, a
// CALL for denormal divisors) per call.  In the membrane kernel that costs three
// things the ncu capture of the first version showed (profiles/r1_hh_ideal_v0.md):
// 58 % of the instruction stream is non-FP64, the branches split the six
// independent exps of a Hodgkin-Huxley right-hand side into separate basic
// blocks (no interleaving -> "wait" stalls), and the 59 KB loop body misses the
// instruction cache ("no_instructions" stalls).  The versions below are
// straight-line: 16 FP64-pipe instructions per exp, 6 (+1 MUFU) per division,
// 3 (+1 MUFU) per reciprocal, 7 (+1 MUFU) per sqrt, ~27 per log.
//
// Accuracy (tests/test_kem_math.py, host build of this same header against
// long-double libm; tests/test_gpu_math.py on the device): exp < 1 ulp on
// [-708, 709]; rcp, div correctly rounded (0.5 ulp) in their domain.
// Domain notes, all outside anything a finite membrane state produces:
//   * exp flushes to 0 below x = -708.4 (no denormal results), returns +inf above
//     709.09 (libm: above 709.78); NaN propagates; exp(+-inf) is NaN.
//   * rcp/div assume a normal, non-zero divisor and quotient (|b| in [2^-1020, 2^1020]);
//     b = 0 gives NaN instead of +-inf.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define KEM_HD __host__ __device__ __forceinline__
#else
#define KEM_HD static inline
#endif

namespace kem {

KEM_HD double bits_to_double(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
#endif
}

KEM_HD uint64_t double_to_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, sizeof u);
    return u;
#endif
}

// ~20-bit reciprocal seed.  Device: MUFU.RCP64H (rcp.approx.ftz.f64: ignores the
// low 32 mantissa bits of the input, returns a zero low word; measured on B200:
// max |1 - b r| = 2^-19.94, tools/probes/probe_rcp_seed.cu).  Host: the same
// truncations around an exact division, so the host build exercises the same
// refinement from an equally coarse start.
KEM_HD double rcp_seed(double b)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return r;
#else
    const double bt = bits_to_double(double_to_bits(b) & 0xFFFFFFFF00000000ull);
    const double r = 1.0 / bt;
    return bits_to_double(double_to_bits(r) & 0xFFFFFFFF00000000ull);
#endif
}

// 1/b : one third-order step from the seed, r1 = r0 (1 + e + e^2), e = 1 - b r0.
// |e| <= 2^-19.9 leaves a truncation error of e^3 <= 2^-59.8 before the final rounding.
KEM_HD double rcp(double b)
{
    const double r = rcp_seed(b);
    const double e = fma(-b, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// a/b : reciprocal, then one residual correction of the quotient
KEM_HD double div(double a, double b)
{
    const double r = rcp(b);
    const double q = a * r;
    const double rem = fma(-b, q, a);
    return fma(rem, r, q);
}

// Constants of exp() in one table.  On the device it lives in constant memory so
// that ptxas loads it into uniform registers once (LDCU.128, two doubles per
// instruction) instead of re-materialising every 64-bit literal with a UMOV pair
// inside the sub-step loop (what the first fast build did: 34 UMOV per stage).
#define KEM_EXP_TABLE                                                                   \
    {0x1.71547652b82fep+0,  /* [0]  1/ln2 */                                            \
     0x1.8p+52 + 1023.0,    /* [1]  1.5*2^52 + bias: low word of the sum = k + 1023 */  \
     -0x1.62e42fefa39efp-1, /* [2]  -ln2 (high part) */                                 \
     -0x1.abc9e3b39803fp-56,/* [3]  -ln2 (low part)  */                                 \
     0x1.af38a9b0ec855p-26, /* [4]  c11 */                                              \
     0x1.289185613a3d6p-22, /* [5]  c10 */                                              \
     0x1.71de0dae63bb3p-19, /* [6]  c9  */                                              \
     0x1.a019b90d2ae7ap-16, /* [7]  c8  */                                              \
     0x1.a01a01a7c41d5p-13, /* [8]  c7  */                                              \
     0x1.6c16c1788bd90p-10, /* [9]  c6  */                                              \
     0x1.11111111109b3p-7,  /* [10] c5  */                                              \
     0x1.5555555553d63p-5,  /* [11] c4  */                                              \
     0x1.5555555555556p-3,  /* [12] c3  */                                              \
     0x1.0000000000001p-1}  /* [13] c2  */

#if defined(__CUDACC__)
__constant__ double KEM_EXP_C_DEV[14] = KEM_EXP_TABLE;
#endif
static const double KEM_EXP_C_HOST[14] = KEM_EXP_TABLE;

#if defined(__CUDA_ARCH__)
#define KEM_EXP_C KEM_EXP_C_DEV
#else
#define KEM_EXP_C KEM_EXP_C_HOST
#endif

// exp(x) = 2^k * p(r),  k = rint(x/ln2),  r = x - k ln2 in [-ln2/2, ln2/2],
// p = degree-11 polynomial (Chebyshev-node fit of (e^r-1-r)/r^2, c0 = c1 = 1;
// max relative error 1.6e-17 before rounding, tools/fit_exp_poly.py).
// 2^k is assembled in the integer pipe from the low word of the magic-number sum,
// clamped to the exponent field: k < -1022 gives 0, k > 1023 gives +inf.
KEM_HD double exp(double x)
{
    const double t = fma(x, KEM_EXP_C[0], KEM_EXP_C[1]);
    const double kd = t - KEM_EXP_C[1];
    double r = fma(kd, KEM_EXP_C[2], x);
    r = fma(kd, KEM_EXP_C[3], r);
    double p = KEM_EXP_C[4];
    p = fma(p, r, KEM_EXP_C[5]);
    p = fma(p, r, KEM_EXP_C[6]);
    p = fma(p, r, KEM_EXP_C[7]);
    p = fma(p, r, KEM_EXP_C[8]);
    p = fma(p, r, KEM_EXP_C[9]);
    p = fma(p, r, KEM_EXP_C[10]);
    p = fma(p, r, KEM_EXP_C[11]);
    p = fma(p, r, KEM_EXP_C[12]);
    p = fma(p, r, KEM_EXP_C[13]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int u = (int)(uint32_t)(double_to_bits(t) & 0xFFFFFFFFull);   // k + 1023
    u = u < 0 ? 0 : (u > 2047 ? 2047 : u);
    const double scale = bits_to_double((uint64_t)(uint32_t)u << 52);
    return p * scale;
}

// ~20-bit reciprocal-square-root seed (device: MUFU.RSQ64H; host: truncated 1/sqrt).
KEM_HD double rsqrt_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    const double xt = bits_to_double(double_to_bits(x) & 0xFFFFFFFF00000000ull);
    const double r = 1.0 / ::sqrt(xt);
    return bits_to_double(double_to_bits(r) & 0xFFFFFFFF00000000ull);
#endif
}

// sqrt(x), x > 0 normal: coupled Newton step on (g ~ sqrt x, h ~ 1/(2 sqrt x)) from the
// seed (2^-20 -> 2^-40), then one residual correction g += (x - g^2) h (-> rounding level).
// 7 FP64 instructions + 1 MUFU; sqrt(+-0) = +-0 is patched by an integer select (the
// straight-line path would give 0 * inf); negative x gives NaN, +inf gives NaN.
KEM_HD double sqrt(double x)
{
    const double y = rsqrt_seed(x);
    double g = x * y;
    double h = 0.5 * y;
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    const double res = fma(d, h, g);
    return ((double_to_bits(x) << 1) == 0) ? x : res;
}

// x^1.5 = x sqrt(x): two roundings, <= 1 ulp (CUDA's pow is specified to 2 ulp).
KEM_HD double pow15(double x) { return x * kem::sqrt(x); }

#define KEM_LOG_TABLE                                                                   \
    {0x1.5555555555558p-1, /* Lg1 */ 0x1.99999999949a8p-2, /* Lg2 */                    \
     0x1.2492492ef4288p-2, /* Lg3 */ 0x1.c71c619eb4eadp-3, /* Lg4 */                    \
     0x1.746310bc043a5p-3, /* Lg5 */ 0x1.39f28b0407abap-3, /* Lg6 */                    \
     0x1.2be91695763e8p-3, /* Lg7 */                                                    \
     0x1.62e42fee00000p-1, /* ln2 high: low 21 mantissa bits zero, k*ln2_hi exact */    \
     0x1.a39ef35793c76p-33 /* ln2 low */}

#if defined(__CUDACC__)
__constant__ double KEM_LOG_C_DEV[9] = KEM_LOG_TABLE;
#endif
static const double KEM_LOG_C_HOST[9] = KEM_LOG_TABLE;
#if defined(__CUDA_ARCH__)
#define KEM_LOG_C KEM_LOG_C_DEV
#else
#define KEM_LOG_C KEM_LOG_C_HOST
#endif

// log(x), x > 0 normal.  x = 2^k m, m in [sqrt(1/2), sqrt(2)); f = m - 1; s = f/(2+f);
// log(1+f) = f - hfsq + s (hfsq + R(s^2)), hfsq = f^2/2 (the classic fdlibm arrangement;
// coefficients from tools/fit_log_poly.py).  Straight-line; x <= 0, inf, NaN are patched
// at the end with selects: log(0) = -inf, log(x<0) = NaN, log(inf) = inf.
// Denormal x is treated as 0 (-inf); log(-0) = -inf.
KEM_HD double log(double x)
{
    const uint64_t bx = double_to_bits(x);
    int hx = (int)(uint32_t)(bx >> 32);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;          // mantissa above sqrt(2): halve it
    k += i >> 20;
    const uint64_t bm = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (bx & 0xFFFFFFFFull);
    const double f = bits_to_double(bm) - 1.0;
    const double dk = (double)k;
    const double s = f * kem::rcp(2.0 + f);
    const double z = s * s;
    const double w = z * z;
    double t1 = fma(w, KEM_LOG_C[5], KEM_LOG_C[3]);
    t1 = fma(w, t1, KEM_LOG_C[1]);
    t1 = w * t1;
    double t2 = fma(w, KEM_LOG_C[6], KEM_LOG_C[4]);
    t2 = fma(w, t2, KEM_LOG_C[2]);
    t2 = fma(w, t2, KEM_LOG_C[0]);
    const double R = fma(z, t2, t1);
    const double hfsq = 0.5 * f * f;
    // dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f)
    const double a = fma(s, hfsq + R, dk * KEM_LOG_C[8]);
    double res = fma(dk, KEM_LOG_C[7], -((hfsq - a) - f));
    // special operands, decided on the high word in the integer pipe:
    // x < 2^-1022 (zero, denormal, any negative) and exponent field all ones (inf, NaN)
    const int hx0 = (int)(uint32_t)(bx >> 32);
    const bool neg = hx0 < 0 && (bx << 1) != 0;        // log(-0) = -inf like log(+0)
    const double low = bits_to_double(neg ? 0x7FF8000000000000ull : 0xFFF0000000000000ull);
    res = (hx0 < 0x00100000) ? low : res;              // -> NaN for negatives, -inf for 0
    res = (hx0 >= 0x7ff00000) ? x : res;               // +inf -> +inf, NaN -> NaN
    return res;
}

}  // namespace kem
