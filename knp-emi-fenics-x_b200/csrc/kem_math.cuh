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
// straight-line: 16 FP64-pipe instructions per exp, 8 (+1 MUFU) per division.
//
// Accuracy (tests/test_kem_math.py, host build of this same header against
// long-double libm): exp <= 1 ulp on [-708, 709]; div, rcp <= 1 ulp.
// Domain notes, all outside anything a finite membrane state produces:
//   * exp saturates instead of overflowing: x > 709.78 gives ~2^1023..2^1024
//     (finite), x < -708.4 gives ~2^-1022; NaN propagates.
//   * rcp/div assume a normal, non-zero divisor (|b| in [2^-1020, 2^1020]);
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
// low 32 mantissa bits of the input, returns zero low word).  Host: the same
// truncations around an exact division, so the host build exercises the same
// Newton iterations from an equally coarse start.
KEM_HD double rcp_seed(double b)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return r;
#else
    const double bt = bits_to_double(double_to_bits(b) & 0xFFFFFFFF00000000ull);
    const double r = 1.0 / bt;
    return bits_to_double(double_to_bits(r) & 0xFFFFFFFFFFF00000ull & 0xFFFFFFFF00000000ull);
#endif
}

// 1/b : two Newton steps from the seed (2^-20 -> 2^-40 -> rounding level)
KEM_HD double rcp(double b)
{
    double r = rcp_seed(b);
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// a/b : reciprocal, then one residual correction of the quotient
KEM_HD double div(double a, double b)
{
    const double r = rcp(b);
    const double q = a * r;
    const double rem = fma(-b, q, a);
    return fma(rem, r, q);
}

// exp(x) = 2^k * p(r),  k = rint(x/ln2),  r = x - k ln2 in [-ln2/2, ln2/2],
// p = degree-11 polynomial (Chebyshev-node fit of (e^r-1-r)/r^2, c0 = c1 = 1;
// max relative error 1.6e-17 before rounding, tools/fit_exp_poly.py).
KEM_HD double exp(double x)
{
    const double L2E = 0x1.71547652b82fep+0;      // 1/ln2
    const double LN2_HI = 0x1.62e42fefa39efp-1;
    const double LN2_LO = 0x1.abc9e3b39803fp-56;
    const double MAGIC = 0x1.8p+52;               // 1.5 * 2^52: low word of the sum holds k
    const double t = fma(x, L2E, MAGIC);
    const double kd = t - MAGIC;
    double r = fma(kd, -LN2_HI, x);
    r = fma(kd, -LN2_LO, r);
    double p = 0x1.af38a9b0ec855p-26;
    p = fma(p, r, 0x1.289185613a3d6p-22);
    p = fma(p, r, 0x1.71de0dae63bb3p-19);
    p = fma(p, r, 0x1.a019b90d2ae7ap-16);
    p = fma(p, r, 0x1.a01a01a7c41d5p-13);
    p = fma(p, r, 0x1.6c16c1788bd90p-10);
    p = fma(p, r, 0x1.11111111109b3p-7);
    p = fma(p, r, 0x1.5555555553d63p-5);
    p = fma(p, r, 0x1.5555555555556p-3);
    p = fma(p, r, 0x1.0000000000001p-1);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int k = (int)(uint32_t)(double_to_bits(t) & 0xFFFFFFFFull);
    k = k < -1022 ? -1022 : (k > 1023 ? 1023 : k);
    const double scale = bits_to_double((uint64_t)(uint32_t)(k + 1023) << 52);
    return p * scale;
}

}  // namespace kem
