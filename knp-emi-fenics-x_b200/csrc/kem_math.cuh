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

// ---- exp ------------------------------------------------------------------------------
// Table-assisted: x = (256 k + j) ln2/256 + r with |r| <= ln2/512 = 1.35e-3, so that
//     exp(x) = 2^k * T[j] * (1 + r + r^2 q(r)),   T[j] = 2^(j/256),   q of degree 2
// (tools/fit_exp_table.py: polynomial error 9.5e-18 = 0.09 ulp; T rounded to double).
// 10 FP64-pipe instructions (the degree-11 polynomial without a table took 16; this kernel is
// bound by the FP64 pipe, the table look-up is one LDS of the otherwise idle shared-memory
// pipe).  < 1 ulp: one rounding for T, one for T + T e, the scaling by 2^k is exact.
//
// The table lives in shared memory (per-thread index: constant memory would serialise):
// every kernel that evaluates kem::exp calls kem::load_tables() and __syncthreads() first.
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

#if defined(__CUDACC__)
// Copy the exp table into this block's shared memory (coalesced, L2-resident: 2 KB).
// Call from every thread of the block, then __syncthreads().
__device__ __forceinline__ void load_tables()
{
    for (int k = threadIdx.x; k < KEM_EXP_TABLE_SIZE; k += blockDim.x) kem_exp_tab_s[k] = KEM_EXP_T_DEV[k];
}
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
    const int hx0 = (int)(uint32_t)(bx >> 32);
    int hx = hx0;
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;          // mantissa above sqrt(2): halve it
    k += i >> 20;
    const uint64_t bm = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (bx & 0xFFFFFFFFull);
    const double f = bits_to_double(bm) - 1.0;
    // Special operands are steered through dk, the exponent as a double: the last operation is
    // fma(dk, ln2_hi, tail), so dk = -inf / +inf / NaN makes the result
    // log(0) = -inf (zero, denormal, -0), log(+inf) = +inf, log(x < 0) = log(NaN) = NaN.  Only
    // the high word of dk is selected ((double)k has a zero low word): three integer
    // instructions on the normal path instead of three 64-bit selects on the result.
    const uint32_t dk_hi = (uint32_t)(double_to_bits((double)k) >> 32);
    const bool ordinary = (uint32_t)(hx0 - 0x00100000) < 0x7fe00000u;   // positive, normal, finite
    // (a NaN whose payload sits in the low word only is reported as +inf: still not finite)
    uint32_t odd = hx0 >= 0x7ff00000 ? (uint32_t)hx0 : 0xfff00000u;              // inf, NaN : zero
    odd = (uint32_t)hx0 > 0x80000000u ? 0x7ff80000u : odd;                       // negative: NaN
    const double dk = bits_to_double((uint64_t)(ordinary ? dk_hi : odd) << 32);
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
    // dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f).  With dk = +-inf the inner term is an
    // infinity of the same sign (ln2_lo > 0, everything else is finite), so the outer FMA adds
    // two like-signed infinities: no inf - inf.
    const double a = fma(s, hfsq + R, dk * KEM_LOG_C[8]);
    return fma(dk, KEM_LOG_C[7], -((hfsq - a) - f));
}

}  // namespace kem
