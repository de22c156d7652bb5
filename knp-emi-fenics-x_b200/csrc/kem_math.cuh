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
// straight-line: 10 FP64-pipe instructions per exp (+1 LDS), 6 (+1 MUFU) per division,
// 3 (+1 MUFU) per reciprocal, 7 (+1 MUFU) per sqrt, 12 per log (+1 LDS).
//
// Accuracy (tests/test_kem_math.py, host build of this same header against
// long-double libm; tests/test_gpu_math.py on the device): exp <= 1.06 ulp on
// [-708, 709], log <= 1.6 ulp; rcp 0.51 ulp, div correctly rounded (0.5 ulp) in their domain.
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

// ---- exp ------------------------------------------------------------------------------
// Table-assisted: x = (256 k + j) ln2/256 + r with |r| <= ln2/512 = 1.35e-3, so that
//     exp(x) = 2^k * T[j] * (1 + r + r^2 q(r)),   T[j] = 2^(j/256),   q of degree 2
// (tools/fit_exp_table.py: polynomial error 9.5e-18 = 0.09 ulp; T rounded to double).
// 10 FP64-pipe instructions (the degree-11 polynomial without a table took 16; this kernel is
// bound by the FP64 pipe, the table look-up is one LDS of the otherwise idle shared-memory
// pipe).  < 1 ulp: one rounding for T, one for T + T e, the scaling by 2^k is exact.
//
// The table lives in shared memory (per-thread index: constant memory would serialise):
// every kernel that evaluates kem::exp calls kem::load_tables<...>() and __syncthreads() first.
// Domain: |x| < 5e6 (the low word of the magic-number sum holds 256 k + j); results below
// 2^-1022 flush to 0, above 2^1024 give +inf, NaN propagates, exp(+-inf) is NaN.
#define KEM_EXP_TABLE_SIZE 256
#define KEM_EXP_TABLE_VALUES \
    0x1.0000000000000p+0, 0x1.00b1afa5abcbfp+0, 0x1.0163da9fb3335p+0, 0x1.02168143b0281p+0,    \
    0x1.02c9a3e778061p+0, 0x1.037d42e11bbccp+0, 0x1.04315e86e7f85p+0, 0x1.04e5f72f654b1p+0,    \
    0x1.059b0d3158574p+0, 0x1.0650a0e3c1f89p+0, 0x1.0706b29ddf6dep+0, 0x1.07bd42b72a836p+0,    \
    0x1.0874518759bc8p+0, 0x1.092bdf66607e0p+0, 0x1.09e3ecac6f383p+0, 0x1.0a9c79b1f3919p+0,    \
    0x1.0b5586cf9890fp+0, 0x1.0c0f145e46c85p+0, 0x1.0cc922b7247f7p+0, 0x1.0d83b23395decp+0,    \
    0x1.0e3ec32d3d1a2p+0, 0x1.0efa55fdfa9c5p+0, 0x1.0fb66affed31bp+0, 0x1.1073028d7233ep+0,    \
    0x1.11301d0125b51p+0, 0x1.11edbab5e2ab6p+0, 0x1.12abdc06c31ccp+0, 0x1.136a814f204abp+0,    \
    0x1.1429aaea92de0p+0, 0x1.14e95934f312ep+0, 0x1.15a98c8a58e51p+0, 0x1.166a45471c3c2p+0,    \
    0x1.172b83c7d517bp+0, 0x1.17ed48695bbc0p+0, 0x1.18af9388c8deap+0, 0x1.1972658375d2fp+0,    \
    0x1.1a35beb6fcb75p+0, 0x1.1af99f8138a1cp+0, 0x1.1bbe084045cd4p+0, 0x1.1c82f95281c6bp+0,    \
    0x1.1d4873168b9aap+0, 0x1.1e0e75eb44027p+0, 0x1.1ed5022fcd91dp+0, 0x1.1f9c18438ce4dp+0,    \
    0x1.2063b88628cd6p+0, 0x1.212be3578a819p+0, 0x1.21f49917ddc96p+0, 0x1.22bdda27912d1p+0,    \
    0x1.2387a6e756238p+0, 0x1.2451ffb82140ap+0, 0x1.251ce4fb2a63fp+0, 0x1.25e85711ece75p+0,    \
    0x1.26b4565e27cddp+0, 0x1.2780e341ddf29p+0, 0x1.284dfe1f56381p+0, 0x1.291ba7591bb70p+0,    \
    0x1.29e9df51fdee1p+0, 0x1.2ab8a66d10f13p+0, 0x1.2b87fd0dad990p+0, 0x1.2c57e39771b2fp+0,    \
    0x1.2d285a6e4030bp+0, 0x1.2df961f641589p+0, 0x1.2ecafa93e2f56p+0, 0x1.2f9d24abd886bp+0,    \
    0x1.306fe0a31b715p+0, 0x1.31432edeeb2fdp+0, 0x1.32170fc4cd831p+0, 0x1.32eb83ba8ea32p+0,    \
    0x1.33c08b26416ffp+0, 0x1.3496266e3fa2dp+0, 0x1.356c55f929ff1p+0, 0x1.36431a2de883bp+0,    \
    0x1.371a7373aa9cbp+0, 0x1.37f26231e754ap+0, 0x1.38cae6d05d866p+0, 0x1.39a401b7140efp+0,    \
    0x1.3a7db34e59ff7p+0, 0x1.3b57fbfec6cf4p+0, 0x1.3c32dc313a8e5p+0, 0x1.3d0e544ede173p+0,    \
    0x1.3dea64c123422p+0, 0x1.3ec70df1c5175p+0, 0x1.3fa4504ac801cp+0, 0x1.40822c367a024p+0,    \
    0x1.4160a21f72e2ap+0, 0x1.423fb2709468ap+0, 0x1.431f5d950a897p+0, 0x1.43ffa3f84b9d4p+0,    \
    0x1.44e086061892dp+0, 0x1.45c2042a7d232p+0, 0x1.46a41ed1d0057p+0, 0x1.4786d668b3237p+0,    \
    0x1.486a2b5c13cd0p+0, 0x1.494e1e192aed2p+0, 0x1.4a32af0d7d3dep+0, 0x1.4b17dea6db7d7p+0,    \
    0x1.4bfdad5362a27p+0, 0x1.4ce41b817c114p+0, 0x1.4dcb299fddd0dp+0, 0x1.4eb2d81d8abffp+0,    \
    0x1.4f9b2769d2ca7p+0, 0x1.508417f4531eep+0, 0x1.516daa2cf6642p+0, 0x1.5257de83f4eefp+0,    \
    0x1.5342b569d4f82p+0, 0x1.542e2f4f6ad27p+0, 0x1.551a4ca5d920fp+0, 0x1.56070dde910d2p+0,    \
    0x1.56f4736b527dap+0, 0x1.57e27dbe2c4cfp+0, 0x1.58d12d497c7fdp+0, 0x1.59c0827ff07ccp+0,    \
    0x1.5ab07dd485429p+0, 0x1.5ba11fba87a03p+0, 0x1.5c9268a5946b7p+0, 0x1.5d84590998b93p+0,    \
    0x1.5e76f15ad2148p+0, 0x1.5f6a320dceb71p+0, 0x1.605e1b976dc09p+0, 0x1.6152ae6cdf6f4p+0,    \
    0x1.6247eb03a5585p+0, 0x1.633dd1d1929fdp+0, 0x1.6434634ccc320p+0, 0x1.652b9febc8fb7p+0,    \
    0x1.6623882552225p+0, 0x1.671c1c70833f6p+0, 0x1.68155d44ca973p+0, 0x1.690f4b19e9538p+0,    \
    0x1.6a09e667f3bcdp+0, 0x1.6b052fa75173ep+0, 0x1.6c012750bdabfp+0, 0x1.6cfdcddd47645p+0,    \
    0x1.6dfb23c651a2fp+0, 0x1.6ef9298593ae5p+0, 0x1.6ff7df9519484p+0, 0x1.70f7466f42e87p+0,    \
    0x1.71f75e8ec5f74p+0, 0x1.72f8286ead08ap+0, 0x1.73f9a48a58174p+0, 0x1.74fbd35d7cbfdp+0,    \
    0x1.75feb564267c9p+0, 0x1.77024b1ab6e09p+0, 0x1.780694fde5d3fp+0, 0x1.790b938ac1cf6p+0,    \
    0x1.7a11473eb0187p+0, 0x1.7b17b0976cfdbp+0, 0x1.7c1ed0130c132p+0, 0x1.7d26a62ff86f0p+0,    \
    0x1.7e2f336cf4e62p+0, 0x1.7f3878491c491p+0, 0x1.80427543e1a12p+0, 0x1.814d2add106d9p+0,    \
    0x1.82589994cce13p+0, 0x1.8364c1eb941f7p+0, 0x1.8471a4623c7adp+0, 0x1.857f4179f5b21p+0,    \
    0x1.868d99b4492edp+0, 0x1.879cad931a436p+0, 0x1.88ac7d98a6699p+0, 0x1.89bd0a478580fp+0,    \
    0x1.8ace5422aa0dbp+0, 0x1.8be05bad61778p+0, 0x1.8cf3216b5448cp+0, 0x1.8e06a5e0866d9p+0,    \
    0x1.8f1ae99157736p+0, 0x1.902fed0282c8ap+0, 0x1.9145b0b91ffc6p+0, 0x1.925c353aa2fe2p+0,    \
    0x1.93737b0cdc5e5p+0, 0x1.948b82b5f98e5p+0, 0x1.95a44cbc8520fp+0, 0x1.96bdd9a7670b3p+0,    \
    0x1.97d829fde4e50p+0, 0x1.98f33e47a22a2p+0, 0x1.9a0f170ca07bap+0, 0x1.9b2bb4d53fe0dp+0,    \
    0x1.9c49182a3f090p+0, 0x1.9d674194bb8d5p+0, 0x1.9e86319e32323p+0, 0x1.9fa5e8d07f29ep+0,    \
    0x1.a0c667b5de565p+0, 0x1.a1e7aed8eb8bbp+0, 0x1.a309bec4a2d33p+0, 0x1.a42c980460ad8p+0,    \
    0x1.a5503b23e255dp+0, 0x1.a674a8af46052p+0, 0x1.a799e1330b358p+0, 0x1.a8bfe53c12e59p+0,    \
    0x1.a9e6b5579fdbfp+0, 0x1.ab0e521356ebap+0, 0x1.ac36bbfd3f37ap+0, 0x1.ad5ff3a3c2774p+0,    \
    0x1.ae89f995ad3adp+0, 0x1.afb4ce622f2ffp+0, 0x1.b0e07298db666p+0, 0x1.b20ce6c9a8952p+0,    \
    0x1.b33a2b84f15fbp+0, 0x1.b468415b749b1p+0, 0x1.b59728de5593ap+0, 0x1.b6c6e29f1c52ap+0,    \
    0x1.b7f76f2fb5e47p+0, 0x1.b928cf22749e4p+0, 0x1.ba5b030a1064ap+0, 0x1.bb8e0b79a6f1fp+0,    \
    0x1.bcc1e904bc1d2p+0, 0x1.bdf69c3f3a207p+0, 0x1.bf2c25bd71e09p+0, 0x1.c06286141b33dp+0,    \
    0x1.c199bdd85529cp+0, 0x1.c2d1cd9fa652cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c544778fafb22p+0,    \
    0x1.c67f12e57d14bp+0, 0x1.c7ba88988c933p+0, 0x1.c8f6d9406e7b5p+0, 0x1.ca3405751c4dbp+0,    \
    0x1.cb720dcef9069p+0, 0x1.ccb0f2e6d1675p+0, 0x1.cdf0b555dc3fap+0, 0x1.cf3155b5bab74p+0,    \
    0x1.d072d4a07897cp+0, 0x1.d1b532b08c968p+0, 0x1.d2f87080d89f2p+0, 0x1.d43c8eacaa1d6p+0,    \
    0x1.d5818dcfba487p+0, 0x1.d6c76e862e6d3p+0, 0x1.d80e316c98398p+0, 0x1.d955d71ff6075p+0,    \
    0x1.da9e603db3285p+0, 0x1.dbe7cd63a8315p+0, 0x1.dd321f301b460p+0, 0x1.de7d5641c0658p+0,    \
    0x1.dfc97337b9b5fp+0, 0x1.e11676b197d17p+0, 0x1.e264614f5a129p+0, 0x1.e3b333b16ee12p+0,    \
    0x1.e502ee78b3ff6p+0, 0x1.e653924676d76p+0, 0x1.e7a51fbc74c83p+0, 0x1.e8f7977cdb740p+0,    \
    0x1.ea4afa2a490dap+0, 0x1.eb9f4867cca6ep+0, 0x1.ecf482d8e67f1p+0, 0x1.ee4aaa2188510p+0,    \
    0x1.efa1bee615a27p+0, 0x1.f0f9c1cb6412ap+0, 0x1.f252b376bba97p+0, 0x1.f3ac948dd7274p+0,    \
    0x1.f50765b6e4540p+0, 0x1.f6632798844f8p+0, 0x1.f7bfdad9cbe14p+0, 0x1.f91d802243c89p+0,    \
    0x1.fa7c1819e90d8p+0, 0x1.fbdba3692d514p+0, 0x1.fd3c22b8f71f1p+0, 0x1.fe9d96b2a23d9p+0

#define KEM_EXP_CONSTS                                                                   \
    {0x1.71547652b82fep+8,   /* [0] 256/ln2 */                                           \
     0x1.8p+52,              /* [1] 1.5*2^52: low word of the sum = 256 k + j */         \
     -0x1.62e42fefa39efp-9,  /* [2] -ln2/256 (high part) */                              \
     -0x1.abc9e3b39803fp-64, /* [3] -ln2/256 (low part)  */                              \
     0x1.5555565c3ff25p-5,   /* [4] c4 */                                                \
     0x1.555556dfb5410p-3}   /* [5] c3  (c2 = 1/2 is an immediate) */

#if defined(__CUDACC__)
__constant__ double KEM_EXP_C_DEV[6] = KEM_EXP_CONSTS;
__device__ const double KEM_EXP_T_DEV[KEM_EXP_TABLE_SIZE] = {KEM_EXP_TABLE_VALUES};
__shared__ double kem_exp_tab_s[KEM_EXP_TABLE_SIZE];
__constant__ uint32_t KEM_EXP_BIAS_DEV = 0x3ff00000u;
#endif
static const double KEM_EXP_C_HOST[6] = KEM_EXP_CONSTS;
static const double KEM_EXP_T_HOST[KEM_EXP_TABLE_SIZE] = {KEM_EXP_TABLE_VALUES};

#if defined(__CUDA_ARCH__)
#define KEM_EXP_C KEM_EXP_C_DEV
#define KEM_EXP_T kem_exp_tab_s
#else
#define KEM_EXP_C KEM_EXP_C_HOST
#define KEM_EXP_T KEM_EXP_T_HOST
#endif

#if defined(KEM_EXP_NO_TABLE)
// Triage build (nvcc_flags=("-DKEM_EXP_NO_TABLE",)): the table-free exp of round 1, k = rint(x/ln2),
// degree-11 polynomial on [-ln2/2, ln2/2] (tools/fit_exp_poly.py), 16 FP64-pipe instructions, 0.94 ulp.
KEM_HD double exp(double x)
{
    const double t = fma(x, 0x1.71547652b82fep+0, 0x1.8p+52 + 1023.0);
    const double kd = t - (0x1.8p+52 + 1023.0);
    double r = fma(kd, -0x1.62e42fefa39efp-1, x);
    r = fma(kd, -0x1.abc9e3b39803fp-56, r);
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
    int u = (int)(uint32_t)(double_to_bits(t) & 0xFFFFFFFFull);   // k + 1023
    u = u < 0 ? 0 : (u > 2047 ? 2047 : u);
    return p * bits_to_double((uint64_t)(uint32_t)u << 52);
}
#else
KEM_HD double exp(double x)
{
    const double t = fma(x, KEM_EXP_C[0], KEM_EXP_C[1]);
    const double kd = t - KEM_EXP_C[1];
    double r = fma(kd, KEM_EXP_C[2], x);
    r = fma(kd, KEM_EXP_C[3], r);
    int m = (int)(uint32_t)(double_to_bits(t) & 0xFFFFFFFFull);         // 256 k + j, two's complement
    // clamp once, on 256 k + j: k = -1023 gives the scale 0 (flush), k = 1024 gives +inf
    m = m < -1023 * 256 ? -1023 * 256 : (m > 1024 * 256 ? 1024 * 256 : m);
    const double T = KEM_EXP_T[m & (KEM_EXP_TABLE_SIZE - 1)];
    // high word of 2^k: (k + 1023) << 20 = (256 k) * 4096 + (1023 << 20)
#if defined(__CUDA_ARCH__)
    // (one IMAD: the bias comes from constant memory so that ptxas cannot fold it into a
    // separate add before the shift)
    uint32_t hi;
    asm("mad.lo.u32 %0, %1, 4096, %2;" : "=r"(hi) : "r"((uint32_t)(m & ~(KEM_EXP_TABLE_SIZE - 1))), "r"(KEM_EXP_BIAS_DEV));
#else
    const uint32_t hi = (uint32_t)(m & ~(KEM_EXP_TABLE_SIZE - 1)) * 4096u + 0x3ff00000u;
#endif
    const double scale = bits_to_double((uint64_t)hi << 32);
    double q = fma(KEM_EXP_C[4], r, KEM_EXP_C[5]);
    q = fma(q, r, 0.5);
    const double e = fma(r * r, q, r);                                  // exp(r) - 1
    return fma(T, e, T) * scale;
}
#endif

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
    const uint64_t bx = double_to_bits(x);
    return ((((uint32_t)(bx >> 32) & 0x7fffffffu) | (uint32_t)bx) == 0u) ? x : res;
}

// x^1.5 = x sqrt(x): two roundings, <= 1 ulp (CUDA's pow is specified to 2 ulp).
KEM_HD double pow15(double x) { return x * kem::sqrt(x); }

// ---- log -------------------------------------------------------------------------------
// Table-assisted (round 2): x = 2^k m with m in [0.70703125, 1.4140625) -- the halving threshold
// sits on a table boundary (mantissa 0x6a000) instead of exactly sqrt 2.  With
// t = (high word of x) - 0x3fe6a000 the exponent is k = t >> 20, the table index
// j = (t >> 12) & 255 (entries ordered by m) and m's high word is that of x minus t's exponent
// bits: five integer instructions.  Entry j of a 4 KB table in shared memory holds
// (invc_j, logc_j): invc_j ~ 1 / centre of the interval, logc_j = -log(invc_j) of the ROUNDED
// invc_j, so that
//     log(x) = k ln2 + logc_j + log1p(r),   r = m invc_j - 1   (one FMA, |r| <= 2^-8)
// holds exactly; the two intervals next to 1 use invc = 1, logc = 0, i.e. r = m - 1, which keeps
// the result accurate relative to itself as x -> 1.  log1p(r) = r - r^2/2 + r^3 q(r), q of degree
// 4 (tools/fit_log_table.py: 1.1e-19).  12 FP64-pipe instructions + one LDS.128; the fdlibm
// arrangement of round 1 (s = f/(2+f), degree-14 polynomial, a reciprocal) took 27.  <= 1.6 ulp
// (one rounding each for the table entry, k ln2_hi + logc, and the final sum).
// Straight-line; special operands are steered through dk, the exponent as a double: every
// term that carries dk becomes -inf / +inf / NaN with it, so log(0) = -inf (zero, denormal, -0),
// log(+inf) = +inf, log(x < 0) = log(NaN) = NaN, for one 32-bit select on the normal path.
#define KEM_LOG_TABLE_SIZE 256
#define KEM_LOG_TABLE_VALUES \
    {0x1.6993f349cc726p+0, -0x1.61965cdb02c1ep-2}, {0x1.68954dd2390bap+0, -0x1.5ec433d5c35aep-2}, \
    {0x1.67980e0bf08c7p+0, -0x1.5bf406b543db1p-2}, {0x1.669c31075ab40p+0, -0x1.5925d2b112a59p-2}, \
    {0x1.65a1b3dd13357p+0, -0x1.565995069514cp-2}, {0x1.64a893adcd25fp+0, -0x1.538f4af8f72fcp-2}, \
    {0x1.63b0cda236e1cp+0, -0x1.50c6f1d11b97bp-2}, {0x1.62ba5eeade65ep+0, -0x1.4e0086dd8baccp-2}, \
    {0x1.61c544c0161c5p+0, -0x1.4b3c077267e9ap-2}, {0x1.60d17c61da198p+0, -0x1.487970e958771p-2}, \
    {0x1.5fdf0317b5c6fp+0, -0x1.45b8c0a17df12p-2}, {0x1.5eedd630a9fb3p+0, -0x1.42f9f3ff62641p-2}, \
    {0x1.5dfdf303137b6p+0, -0x1.403d086cea79bp-2}, {0x1.5d0f56ec91e57p+0, -0x1.3d81fb5946dbcp-2}, \
    {0x1.5c21ff51ef005p+0, -0x1.3ac8ca38e5c5dp-2}, {0x1.5b35e99f06714p+0, -0x1.3811728564cb2p-2}, \
    {0x1.5a4b1346add2bp+0, -0x1.355bf1bd82c8bp-2}, {0x1.596179c29d2cep+0, -0x1.32a84565120a9p-2}, \
    {0x1.58791a9357ccep+0, -0x1.2ff66b04ea9d5p-2}, {0x1.5791f34015792p+0, -0x1.2d46602adccefp-2}, \
    {0x1.56ac0156ac015p+0, -0x1.2a982269a3dbep-2}, {0x1.55c7426b79286p+0, -0x1.27ebaf58d8c9cp-2}, \
    {0x1.54e3b4194ce66p+0, -0x1.25410494e56c8p-2}, {0x1.5401540154015p+0, -0x1.22981fbef797ap-2}, \
    {0x1.53201fcb02fb1p+0, -0x1.1ff0fe7cf47a9p-2}, {0x1.5240152401524p+0, -0x1.1d4b9e796c245p-2}, \
    {0x1.516131c015161p+0, -0x1.1aa7fd638d33ep-2}, {0x1.508373590ec9cp+0, -0x1.180618ef18adep-2}, \
    {0x1.4fa6d7aeb597cp+0, -0x1.1565eed455fc2p-2}, {0x1.4ecb5c86b3d24p+0, -0x1.12c77cd00713cp-2}, \
    {0x1.4df0ffac83c01p+0, -0x1.102ac0a35cc1bp-2}, {0x1.4d17bef15cb4ep+0, -0x1.0d8fb813eb1efp-2}, \
    {0x1.4c3f982c20723p+0, -0x1.0af660eb9e278p-2}, {0x1.4b68893948d1cp+0, -0x1.085eb8f8ae799p-2}, \
    {0x1.4a928ffad5b5cp+0, -0x1.05c8be0d9635ap-2}, {0x1.49bdaa583b401p+0, -0x1.03346e0106062p-2}, \
    {0x1.48e9d63e504d1p+0, -0x1.00a1c6adda472p-2}, {0x1.4817119f3d325p+0, -0x1.fc218be620a5fp-3}, \
    {0x1.47455a726abf2p+0, -0x1.f702d36777df0p-3}, {0x1.4674aeb4717e9p+0, -0x1.f1e75fadf9bdep-3}, \
    {0x1.45a50c670938fp+0, -0x1.eccf2c8fe920bp-3}, {0x1.44d67190f8b43p+0, -0x1.e7ba35eb77e2ap-3}, \
    {0x1.4408dc3e05b22p+0, -0x1.e2a877a6b2c0fp-3}, {0x1.433c4a7ee52b4p+0, -0x1.dd99edaf6d7e9p-3}, \
    {0x1.4270ba692bc4dp+0, -0x1.d88e93fb2f451p-3}, {0x1.41a62a173e821p+0, -0x1.d38666871f467p-3}, \
    {0x1.40dc97a843ae8p+0, -0x1.ce816157f1985p-3}, {0x1.4014014014014p+0, -0x1.c97f8079d44ecp-3}, \
    {0x1.3f4c65072bf74p+0, -0x1.c480c0005cccfp-3}, {0x1.3e85c12a9d651p+0, -0x1.bf851c067555cp-3}, \
    {0x1.3dc013dc013dcp+0, -0x1.ba8c90ae4ad19p-3}, {0x1.3cfb5b51698ebp+0, -0x1.b5971a213acd9p-3}, \
    {0x1.3c3795c553afbp+0, -0x1.b0a4b48fc1b44p-3}, {0x1.3b74c1769aa5cp+0, -0x1.abb55c31693aep-3}, \
    {0x1.3ab2dca869b81p+0, -0x1.a6c90d44b704cp-3}, {0x1.39f1e5a22f36ep+0, -0x1.a1dfc40f1b7f1p-3}, \
    {0x1.3931daaf8f721p+0, -0x1.9cf97cdce0ec1p-3}, {0x1.3872ba2057e04p+0, -0x1.981634011aa74p-3}, \
    {0x1.37b4824872744p+0, -0x1.9335e5d594985p-3}, {0x1.36f7317fd9212p+0, -0x1.8e588ebac2dc1p-3}, \
    {0x1.363ac622898b1p+0, -0x1.897e2b17b19a6p-3}, {0x1.357f3e9078e5bp+0, -0x1.84a6b759f512dp-3}, \
    {0x1.34c4992d87fd9p+0, -0x1.7fd22ff599d4cp-3}, {0x1.340ad461776d3p+0, -0x1.7b0091651528bp-3}, \
    {0x1.3351ee97dbfc6p+0, -0x1.7631d82935a84p-3}, {0x1.3299e6401329ap+0, -0x1.716600c914055p-3}, \
    {0x1.31e2b9cd37dc2p+0, -0x1.6c9d07d203fc4p-3}, {0x1.312c67b6173eep+0, -0x1.67d6e9d785770p-3}, \
    {0x1.3076ee7525c2cp+0, -0x1.6313a37335d76p-3}, {0x1.2fc24c8874486p+0, -0x1.5e533144c1718p-3}, \
    {0x1.2f0e8071a5703p+0, -0x1.59958ff1d52f4p-3}, {0x1.2e5b88b5e3104p+0, -0x1.54dabc26105d3p-3}, \
    {0x1.2da963ddd3cfbp+0, -0x1.5022b292f6a45p-3}, {0x1.2cf8107590e67p+0, -0x1.4b6d6fefe22a5p-3}, \
    {0x1.2c478d0c9c013p+0, -0x1.46baf0f9f5db8p-3}, {0x1.2b97d835d548ep+0, -0x1.420b32740fdd6p-3}, \
    {0x1.2ae8f087718d0p+0, -0x1.3d5e3126bc281p-3}, {0x1.2a3ad49af0907p+0, -0x1.38b3e9e027477p-3}, \
    {0x1.298d830d13780p+0, -0x1.340c59741142dp-3}, {0x1.28e0fa7dd35a3p+0, -0x1.2f677cbbc0a98p-3}, \
    {0x1.2835399057efdp+0, -0x1.2ac55095f5c5bp-3}, {0x1.278a3eeaee650p+0, -0x1.2625d1e6ddf55p-3}, \
    {0x1.26e009370049cp+0, -0x1.2188fd9807266p-3}, {0x1.263697210aa18p+0, -0x1.1ceed09853755p-3}, \
    {0x1.258de75895121p+0, -0x1.185747dbecf34p-3}, {0x1.24e5f89029305p+0, -0x1.13c2605c398bfp-3}, \
    {0x1.243ec97d49eaep+0, -0x1.0f301717cf0fbp-3}, {0x1.239858d86b11fp+0, -0x1.0aa06912675d5p-3}, \
    {0x1.22f2a55ce8fc5p+0, -0x1.06135354d4b19p-3}, {0x1.224dadc900489p+0, -0x1.0188d2ecf613ep-3}, \
    {0x1.21a970ddc5ba7p+0, -0x1.fa01c9db57ce7p-4}, {0x1.2105ed5f1e336p+0, -0x1.f0f70cdd992e4p-4}, \
    {0x1.20632213b6c6dp+0, -0x1.e7f1691a32d3ap-4}, {0x1.1fc10dc4fce8bp+0, -0x1.def0d8d466dbbp-4}, \
    {0x1.1f1faf3f16b64p+0, -0x1.d5f55659210e1p-4}, {0x1.1e7f0550db594p+0, -0x1.ccfedbfee13a8p-4}, \
    {0x1.1ddf0ecbcb841p+0, -0x1.c40d6425a5cb4p-4}, {0x1.1d3fca840a074p+0, -0x1.bb20e936d6976p-4}, \
    {0x1.1ca13750547fep+0, -0x1.b23965a52ff04p-4}, {0x1.1c035409fc1dfp+0, -0x1.a956d3ecade60p-4}, \
    {0x1.1b661f8cde833p+0, -0x1.a0792e9277cadp-4}, {0x1.1ac998b75eb90p+0, -0x1.97a07024cbe6ep-4}, \
    {0x1.1a2dbe6a5e3e4p+0, -0x1.8ecc933aeb6e2p-4}, {0x1.19928f89362b7p+0, -0x1.85fd927506a46p-4}, \
    {0x1.18f80af9b06dcp+0, -0x1.7d33687c293c8p-4}, {0x1.185e2fa401186p+0, -0x1.746e100226edbp-4}, \
    {0x1.17c4fc72bfcb9p+0, -0x1.6bad83c1883bap-4}, {0x1.172c7052e1316p+0, -0x1.62f1be7d7774ap-4}, \
    {0x1.16948a33b08fap+0, -0x1.5a3abb01ade21p-4}, {0x1.15fd4906c96f1p+0, -0x1.5188742261311p-4}, \
    {0x1.1566abc011567p+0, -0x1.48dae4bc3101dp-4}, {0x1.14d0b155b19aep+0, -0x1.403207b414b79p-4}, \
    {0x1.143b58c01143bp+0, -0x1.378dd7f74970fp-4}, {0x1.13a6a0f9cf01ep+0, -0x1.2eee507b402ffp-4}, \
    {0x1.131288ffbb3b6p+0, -0x1.26536c3d8c36cp-4}, {0x1.127f0fd0d2295p+0, -0x1.1dbd2643d1913p-4}, \
    {0x1.11ec346e36092p+0, -0x1.152b799bb3cd0p-4}, {0x1.1159f5db29606p+0, -0x1.0c9e615ac4e19p-4}, \
    {0x1.10c8531d0952ep+0, -0x1.0415d89e7444bp-4}, {0x1.10374b3b480aap+0, -0x1.f723b517fc51fp-5}, \
    {0x1.0fa6dd3f67322p+0, -0x1.e624c4a0b5e15p-5}, {0x1.0f170834f27fap+0, -0x1.d52ed6405d87ap-5}, \
    {0x1.0e87cb297a51ep+0, -0x1.c441e06f72a93p-5}, {0x1.0df9252c8e5e6p+0, -0x1.b35dd9b58baa8p-5}, \
    {0x1.0d6b154fb86f9p+0, -0x1.a282b8a936174p-5}, {0x1.0cdd9aa677344p+0, -0x1.91b073efd7314p-5}, \
    {0x1.0c50b446391f3p+0, -0x1.80e7023d8ccc8p-5}, {0x1.0bc4614657569p+0, -0x1.70265a550e77bp-5}, \
    {0x1.0b38a0c010b39p+0, -0x1.5f6e73078efc3p-5}, {0x1.0aad71ce84d16p+0, -0x1.4ebf43349e26ap-5}, \
    {0x1.0a22d38eaf2bfp+0, -0x1.3e18c1ca0ae99p-5}, {0x1.0998c51f624d5p+0, -0x1.2d7ae5c3c5bb7p-5}, \
    {0x1.090f45a1430aap+0, -0x1.1ce5a62bc3540p-5}, {0x1.08865436c3cf7p+0, -0x1.0c58fa19dfaabp-5}, \
    {0x1.07fdf0041ff7cp+0, -0x1.f7a9b16782855p-6}, {0x1.0776182f57386p+0, -0x1.d6b272597981fp-6}, \
    {0x1.06eecbe029155p+0, -0x1.b5cc258b718e7p-6}, {0x1.06680a4010668p+0, -0x1.94f6b99a24473p-6}, \
    {0x1.05e1d27a3ee9cp+0, -0x1.74321d3d006d2p-6}, {0x1.055c23bb98e2ap+0, -0x1.537e3f45f354ep-6}, \
    {0x1.04d6fd32b0c7bp+0, -0x1.32db0ea132e10p-6}, {0x1.04525e0fc2fcbp+0, -0x1.12487a5507f68p-6}, \
    {0x1.03ce4584b19a0p+0, -0x1.e38ce30333100p-7}, {0x1.034ab2c50040dp+0, -0x1.a2a9c6c17044dp-7}, \
    {0x1.02c7a505cffbfp+0, -0x1.61e77e8b53f9fp-7}, {0x1.02451b7ddb2d2p+0, -0x1.2145e939ef1bcp-7}, \
    {0x1.01c315657186bp+0, -0x1.c189cbb0e283fp-8}, {0x1.014191f674111p+0, -0x1.40c8a7478788dp-8}, \
    {0x1.00c0906c513cfp+0, -0x1.809048289860ap-9}, {0x1.0000000000000p+0, 0x0.0p+0},        \
    {0x1.0000000000000p+0, 0x0.0p+0}, {0x1.fd04794a10e6ap-1, 0x1.7ee11ebd82ec4p-8},         \
    {0x1.fb0c610d5e939p-1, 0x1.3e7295d25a7d5p-7}, {0x1.f9182b6813bafp-1, 0x1.bcf712c743853p-7}, \
    {0x1.f727cce5f530ap-1, 0x1.1d7f7eb9eebf1p-6}, {0x1.f53b3a3fa204ep-1, 0x1.5c45a51b8d393p-6}, \
    {0x1.f3526859b8cecp-1, 0x1.9ace7551cc515p-6}, {0x1.f16d4c4401f17p-1, 0x1.d91a66c543cbep-6}, \
    {0x1.ef8bdb389ebadp-1, 0x1.0b94f7c196173p-5}, {0x1.edae0a9b3d3a5p-1, 0x1.2a7ec2214e879p-5}, \
    {0x1.ebd3cff850b0cp-1, 0x1.494acc34d911dp-5}, {0x1.e9fd21044e799p-1, 0x1.67f94f094bd92p-5}, \
    {0x1.e829f39aef509p-1, 0x1.868a83083f6d0p-5}, {0x1.e65a3dbe74d6bp-1, 0x1.a4fe9ffa3d233p-5}, \
    {0x1.e48df596f3394p-1, 0x1.c355dd0921f2fp-5}, {0x1.e2c511719ee16p-1, 0x1.e19070c276010p-5}, \
    {0x1.e0ff87c01e100p-1, 0x1.ffae9119b92fbp-5}, {0x1.df3d4f17de4dbp-1, 0x1.0ed839b5526fep-4}, \
    {0x1.dd7e5e316d94cp-1, 0x1.1dcb263db1944p-4}, {0x1.dbc2abe7d71d4p-1, 0x1.2cb0283f5de22p-4}, \
    {0x1.da0a2f3803b41p-1, 0x1.3b87598b1b6f0p-4}, {0x1.d854df401d855p-1, 0x1.4a50d3aa1b03fp-4}, \
    {0x1.d6a2b33ef7448p-1, 0x1.590cafdf01c26p-4}, {0x1.d4f3a293769cap-1, 0x1.67bb0726ec0fbp-4}, \
    {0x1.d347a4bc01d34p-1, 0x1.765bf23a6be17p-4}, {0x1.d19eb155f08a4p-1, 0x1.84ef898e82828p-4}, \
    {0x1.cff8c01cff8c0p-1, 0x1.9375e55595edfp-4}, {0x1.ce55c8eac7900p-1, 0x1.a1ef1d8061cd8p-4}, \
    {0x1.ccb5c3b636e3ap-1, 0x1.b05b49bee4403p-4}, {0x1.cb18a8930de60p-1, 0x1.beba818146764p-4}, \
    {0x1.c97e6fb15e44dp-1, 0x1.cd0cdbf8c13e0p-4}, {0x1.c7e7115d0ce95p-1, 0x1.db5270187d925p-4}, \
    {0x1.c65285fd56843p-1, 0x1.e98b54967146bp-4}, {0x1.c4c0c61456a8ep-1, 0x1.f7b79fec37de2p-4}, \
    {0x1.c331ca3e91679p-1, 0x1.02ebb42bf3d4ap-3}, {0x1.c1a58b327f576p-1, 0x1.09f561ee719c4p-3}, \
    {0x1.c01c01c01c01cp-1, 0x1.10f8e422539b1p-3}, {0x1.be9526d0769fap-1, 0x1.17f6458fca611p-3}, \
    {0x1.bd10f365451b6p-1, 0x1.1eed90e2dc2c3p-3}, {0x1.bb8f609879493p-1, 0x1.25ded0abc6ad3p-3}, \
    {0x1.ba10679bd8488p-1, 0x1.2cca0f5f5f252p-3}, {0x1.b89401b89401cp-1, 0x1.33af575770e4dp-3}, \
    {0x1.b71a284ee6b34p-1, 0x1.3a8eb2d31a375p-3}, {0x1.b5a2d4d5b081fp-1, 0x1.41682bf727bbfp-3}, \
    {0x1.b42e00da17007p-1, 0x1.483bccce6e3dcp-3}, {0x1.b2bba5ff26a23p-1, 0x1.4f099f4a230b1p-3}, \
    {0x1.b14bbdfd760e6p-1, 0x1.55d1ad4232d70p-3}, {0x1.afde42a2cb482p-1, 0x1.5c940075972b9p-3}, \
    {0x1.ae732dd1c2a09p-1, 0x1.6350a28aaa759p-3}, {0x1.ad0a798177693p-1, 0x1.6a079d0f7aad0p-3}, \
    {0x1.aba41fbd2e5b1p-1, 0x1.70b8f97a1aa74p-3}, {0x1.aa401aa401aa4p-1, 0x1.7764c128f2127p-3}, \
    {0x1.a8de64688ebabp-1, 0x1.7e0afd630c276p-3}, {0x1.a77ef750a56dap-1, 0x1.84abb75865137p-3}, \
    {0x1.a621cdb4f8fdfp-1, 0x1.8b46f8223625bp-3}, {0x1.a4c6e200d2637p-1, 0x1.91dcc8c340bdfp-3}, \
    {0x1.a36e2eb1c432dp-1, 0x1.986d3228180c8p-3}, {0x1.a217ae575ff2fp-1, 0x1.9ef83d2769a34p-3}, \
    {0x1.a0c35b92ecdf1p-1, 0x1.a57df28244dcbp-3}, {0x1.9f713117200d0p-1, 0x1.abfe5ae46124ap-3}, \
    {0x1.9e2129a7d5f0ap-1, 0x1.b2797ee46320cp-3}, {0x1.9cd34019cd340p-1, 0x1.b8ef670420c3bp-3}, \
    {0x1.9b876f5262dd1p-1, 0x1.bf601bb0e44e0p-3}, {0x1.9a3db2474fb98p-1, 0x1.c5cba543ae424p-3}, \
    {0x1.98f603fe670a0p-1, 0x1.cc320c0176501p-3}, {0x1.97b05f8d56652p-1, 0x1.d293581b6b3e7p-3}, \
    {0x1.966cc01966cc0p-1, 0x1.d8ef91af31d5ep-3}, {0x1.952b20d73ee97p-1, 0x1.df46c0c722d30p-3}, \
    {0x1.93eb7d0aa6759p-1, 0x1.e598ed5a87e2ep-3}, {0x1.92add0064ab74p-1, 0x1.ebe61f4dd7b0bp-3}, \
    {0x1.9172152b841ddp-1, 0x1.f22e5e72f105cp-3}, {0x1.903847ea1cec1p-1, 0x1.f871b28955045p-3}, \
    {0x1.8f0063c018f00p-1, 0x1.feb0233e607cep-3}, {0x1.8dca64397e408p-1, 0x1.0274dc16c232fp-2}, \
    {0x1.8c9644f01efbcp-1, 0x1.058f3c703ebc5p-2}, {0x1.8b64018b64019p-1, 0x1.08a73667c57aep-2}, \
    {0x1.8a3395c018a34p-1, 0x1.0bbccdb0d24bcp-2}, {0x1.8904fd503744bp-1, 0x1.0ed005f657da5p-2}, \
    {0x1.87d8340ab6e97p-1, 0x1.11e0e2dad9cb6p-2}, {0x1.86ad35cb59a84p-1, 0x1.14ef67f88685ap-2}, \
    {0x1.8583fe7a7c018p-1, 0x1.17fb98e15095ep-2}, {0x1.845c8a0ce5129p-1, 0x1.1b05791f07b4ap-2}, \
    {0x1.8336d48397a24p-1, 0x1.1e0d0c33716bdp-2}, {0x1.8212d9eba4018p-1, 0x1.211255986160cp-2}, \
    {0x1.80f0965dfabcbp-1, 0x1.241558bfd1405p-2}, {0x1.7fd005ff40180p-1, 0x1.27161913f853dp-2}, \
    {0x1.7eb124ffa053bp-1, 0x1.2a1499f762bcap-2}, {0x1.7d93ef9aa4b46p-1, 0x1.2d10dec508582p-2}, \
    {0x1.7c7862170949fp-1, 0x1.300aead06350cp-2}, {0x1.7b5e78c693733p-1, 0x1.3302c1658658ap-2}, \
    {0x1.7a463005e918cp-1, 0x1.35f865c93293ep-2}, {0x1.792f843c689c3p-1, 0x1.38ebdb38ed320p-2}, \
    {0x1.781a71dc01782p-1, 0x1.3bdd24eb14b69p-2}, {0x1.7706f5610d8d0p-1, 0x1.3ecc460ef5f50p-2}, \
    {0x1.75f50b522b17cp-1, 0x1.41b941cce0beep-2}, {0x1.74e4b040174e5p-1, 0x1.44a41b463c47bp-2}, \
    {0x1.73d5e0c5899f7p-1, 0x1.478cd5959b3d8p-2}, {0x1.72c899870f91fp-1, 0x1.4a7373cecf997p-2}, \
    {0x1.71bcd732e940ap-1, 0x1.4d57f8fefe27fp-2}, {0x1.70b29680e66fap-1, 0x1.503a682cb1cb3p-2}, \
    {0x1.6fa9d43244380p-1, 0x1.531ac457ee77fp-2}, {0x1.6ea28d118b474p-1, 0x1.55f9107a43ee2p-2}, \
    {0x1.6d9cbdf26eaefp-1, 0x1.58d54f86e02f3p-2}, {0x1.6c9863b1ab429p-1, 0x1.5baf846aa1b1ap-2}, \
    {0x1.6b957b34e7803p-1, 0x1.5e87b20c2954ap-2}, {0x1.6a94016a94017p-1, 0x1.615ddb4bec13cp-2}

#define KEM_LOG_CONSTS                                                                   \
    {0x1.62e42fee00000p-1,  /* [0] ln2 high: low 21 mantissa bits zero, k*ln2_hi exact */ \
     0x1.a39ef35793c76p-33, /* [1] ln2 low */                                             \
     0x1.249366bad3f6dp-3,  /* [2] q4 */                                                  \
     -0x1.5556969e595b5p-3, /* [3] q3 */                                                  \
     0x1.9999999951eadp-3,  /* [4] q2 */                                                  \
     -0x1.ffffffffaf5b6p-3, /* [5] q1 */                                                  \
     0x1.5555555555555p-2}  /* [6] q0 */

struct LogEntry {
    double invc, logc;
};

#if defined(__CUDACC__)
__constant__ double KEM_LOG_C_DEV[7] = KEM_LOG_CONSTS;
__device__ const LogEntry KEM_LOG_T_DEV[KEM_LOG_TABLE_SIZE] = {KEM_LOG_TABLE_VALUES};
__shared__ __align__(16) LogEntry kem_log_tab_s[KEM_LOG_TABLE_SIZE];
#endif
static const double KEM_LOG_C_HOST[7] = KEM_LOG_CONSTS;
static const LogEntry KEM_LOG_T_HOST[KEM_LOG_TABLE_SIZE] = {KEM_LOG_TABLE_VALUES};
#if defined(__CUDA_ARCH__)
#define KEM_LOG_C KEM_LOG_C_DEV
#define KEM_LOG_T kem_log_tab_s
#else
#define KEM_LOG_C KEM_LOG_C_HOST
#define KEM_LOG_T KEM_LOG_T_HOST
#endif

#if defined(__CUDACC__)
// Copy the look-up tables into this block's shared memory (coalesced, L2-resident: 2 KB for exp,
// 4 KB more for log when the model's code takes logarithms).  Call from every thread of the
// block, then __syncthreads().
template <bool WITH_LOG>
__device__ __forceinline__ void load_tables()
{
    for (int k = threadIdx.x; k < KEM_EXP_TABLE_SIZE; k += blockDim.x) kem_exp_tab_s[k] = KEM_EXP_T_DEV[k];
    if (WITH_LOG)
        for (int k = threadIdx.x; k < KEM_LOG_TABLE_SIZE; k += blockDim.x) kem_log_tab_s[k] = KEM_LOG_T_DEV[k];
}
#endif

KEM_HD double log(double x)
{
    const uint64_t bx = double_to_bits(x);
    const int hx0 = (int)(uint32_t)(bx >> 32);
    const uint32_t t = (uint32_t)hx0 - 0x3fe6a000u;
    const int k = (int)t >> 20;
    const LogEntry T = KEM_LOG_T[(t >> 12) & (KEM_LOG_TABLE_SIZE - 1)];
    const uint32_t mh = (uint32_t)hx0 - (t & 0xfff00000u);
    const double m = bits_to_double(((uint64_t)mh << 32) | (bx & 0xFFFFFFFFull));
    const uint32_t dk_hi = (uint32_t)(double_to_bits((double)k) >> 32);        // ((double)k has a zero low word)
    const bool ordinary = (uint32_t)(hx0 - 0x00100000) < 0x7fe00000u;           // positive, normal, finite
    // (a NaN whose payload sits in the low word only is reported as +inf: still not finite)
    uint32_t odd = hx0 >= 0x7ff00000 ? (uint32_t)hx0 : 0xfff00000u;              // inf, NaN : zero
    odd = (uint32_t)hx0 > 0x80000000u ? 0x7ff80000u : odd;                       // negative: NaN
    const double dk = bits_to_double((uint64_t)(ordinary ? dk_hi : odd) << 32);
    const double r = fma(m, T.invc, -1.0);
    const double w = fma(dk, KEM_LOG_C[0], T.logc);
    double q = fma(KEM_LOG_C[2], r, KEM_LOG_C[3]);
    q = fma(q, r, KEM_LOG_C[4]);
    q = fma(q, r, KEM_LOG_C[5]);
    q = fma(q, r, KEM_LOG_C[6]);
    const double u = fma(r, q, -0.5);
    // with dk = +-inf both w and the tail are infinities of dk's sign (ln2_lo > 0): no inf - inf
    const double p = fma(r * r, u, dk * KEM_LOG_C[1]);
    return w + (r + p);
}

}  // namespace kem
