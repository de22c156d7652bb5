// kem_kernel.cuh -- the fused membrane-ODE step kernel (sm_100a), generic over
// a generated model struct M (see knpemi_b200/codegen/emit.py).
//
// One thread advances one membrane DOF by one PDE step.  It replaces the row
// loop of MembraneModel.step_lsoda (reference src/knpemi/odeSolver.py:107-122):
//   * sticky stimulus write into the parameter table        (odeSolver.py:110-112)
//   * integration of the row from t0 to t0+dt               (odeSolver.py:116-120,
//     here scheme O1: classical RK4, n_sub sub-steps)
//   * the RHS side effect that leaves I_ch_* in parameter slots (mm_hh.py:220-225),
//     evaluated once at (t0+dt, y_end)
//   * write-back of the row                                  (odeSolver.py:122)
//
// Data layout: structure-of-arrays, one contiguous double column per state /
// per-DOF parameter / output slot, so a warp's 32 loads of one column are one
// 256-byte coalesced request.  Each column is read once and written once per
// PDE step; all states stay in registers across the 4*n_sub+1 RHS evaluations.
// Parameters that are uniform over the DOFs are read through a zero index mask
// (one broadcast transaction) from a small per-device table.
//
// Binding roofline: the FP64 pipe (about 25k DFMA-class instructions against
// 144 bytes per DOF-step for the HH models) -- see DESIGN.md.
#pragma once
#include "kem_math.cuh"
#include "kem_model_api.h"

template <class M>
struct KemArgs {
    long long n;
    double *y[M::NS];
    const double *p[M::NP > 0 ? M::NP : 1];      // (a model may have no parameters at all)
    long long pmask[M::NP > 0 ? M::NP : 1];
    double *out[M::NOUT > 0 ? M::NOUT : 1];
    const unsigned char *stim_mask;
    int n_stim;
    int stim_col[KEM_MAX_STIM];
    double stim_val[KEM_MAX_STIM];
    double *stim_ptr[KEM_MAX_STIM];
    const double *ttab;
    int n_sub;
    double h;
    int *flags;
    // scheme O3 (Dormand-Prince 5(4))
    double t0, dt, t_end, rtol, atol;
    double *hsug;
    unsigned long long *stats;
    const int *perm;
    const unsigned *perm_on;
};

// Prologue shared by the step kernels: parameters of DOF i (only the slots the RHS
// reads), sticky stimulus (odeSolver.py:110-112), parameter-only sub-expressions.
template <class M>
__device__ __forceinline__ void kem_prologue(const KemArgs<M> &a, long long i, typename M::H &q)
{
    constexpr int NP = M::NP;
    double p[NP > 0 ? NP : 1];
#pragma unroll
    for (int c = 0; c < NP; ++c)
        p[c] = M::used(c) ? __ldg(a.p[c] + (i & a.pmask[c])) : 0.0;

    if (a.n_stim > 0 && a.stim_mask[i]) {
#pragma unroll
        for (int s = 0; s < KEM_MAX_STIM; ++s) {
            if (s < a.n_stim) {
                const int col = a.stim_col[s];
                const double v = a.stim_val[s];
#pragma unroll
                for (int c = 0; c < NP; ++c)
                    if (M::used(c) && c == col) p[c] = v;
                a.stim_ptr[s][i] = v;
            }
        }
    }
    M::hoist(p, q);
}

// KEM_MIN_BLOCKS (compile-time, default 1) is the minimum-resident-blocks hint of
// __launch_bounds__ for the RK4 kernel: a register cap for occupancy experiments
// (`nvcc_flags=("-DKEM_MIN_BLOCKS=8",)`); the measured optimum is the uncapped default.
#ifndef KEM_MIN_BLOCKS
#define KEM_MIN_BLOCKS 1
#endif

template <class M, int BLOCK>
__global__ void __launch_bounds__(BLOCK, KEM_MIN_BLOCKS)
kem_step_kernel(const __grid_constant__ KemArgs<M> a)
{
    constexpr int NS = M::NS, NOUT = M::NOUT, NT = M::NT;
    extern __shared__ double s_tt[];

    // time-only factors of every stage time of this PDE step (host-evaluated), exp table
    if (NT > 0) {
        const int n_tt = (2 * a.n_sub + 2) * NT;
        for (int k = threadIdx.x; k < n_tt; k += BLOCK) s_tt[k] = a.ttab[k];
    }
    kem::load_tables<M::USES_LOG>();
    __syncthreads();

    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n) return;

    // ---- parameters, sticky stimulus, parameter-only sub-expressions (once per PDE step)
    typename M::H q;
    kem_prologue<M>(a, i, q);

    double y[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) y[c] = a.y[c][i];

    const double h = a.h;
    const double hh = 0.5 * h;
    const double h6 = h / 6.0;

    // ---- classical RK4, association identical to oracle/knpemi_oracle.c:step_row:
    //        acc = ((k1 + 2 k2) + 2 k3) + k4 ;  y += (h/6) acc
    // (With CUDA's libm the four inlined stages were a 59 KB loop body that missed the
    // instruction cache; with the branch-free math of kem_math.cuh one sub-step is ~13 KB.)
    double w[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) w[c] = y[c];
#pragma unroll 1
    for (int j = 0; j < a.n_sub; ++j) {
        const double *tj = s_tt + (2 * j) * NT;
        double acc[NS];
#pragma unroll
        for (int c = 0; c < NS; ++c) acc[c] = 0.0;
        // M::STAGE_UNROLL = 4 (the default for all but very large right-hand sides) takes the
        // four stages inline: stage weights become immediates and the loop control disappears
        // from the instruction stream (hh_ideal: 42 -> 24.5 non-FP64 instructions per stage)
#pragma unroll(M::STAGE_UNROLL)
        for (int s = 0; s < 4; ++s) {
            double k[NS];
            M::deriv(w, k, q, tj + ((s + 1) >> 1) * NT);      // stage times ta, tb, tb, tc
            const double bw = (s == 0 || s == 3) ? 1.0 : 2.0;  // weight in acc (products exact)
            const double aw = (s == 2) ? h : hh;               // offset of the next stage
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                acc[c] = acc[c] + bw * k[c];
                w[c] = y[c] + aw * k[c];
            }
        }
#pragma unroll
        for (int c = 0; c < NS; ++c) {
            y[c] = y[c] + h6 * acc[c];
            w[c] = y[c];
        }
    }

    // ---- current epilogue: I_ch(y(t0+dt)) into the output parameter slots
    if (NOUT > 0) {
        double o[NOUT > 0 ? NOUT : 1];
        M::outputs(y, o, q, s_tt + (2 * a.n_sub + 1) * NT);
#pragma unroll
        for (int c = 0; c < NOUT; ++c) a.out[c][i] = o[c];
    }

    bool finite = true;
#pragma unroll
    for (int c = 0; c < NS; ++c) {
        a.y[c][i] = y[c];
        finite = finite && isfinite(y[c]);
    }
    if (!finite) atomicOr(a.flags, 1);
}


// ------------------------------------------------------------------------------------------
// Scheme O3: error-controlled Dormand-Prince 5(4), one thread = one DOF, step sizes per DOF.
// Device twin of oracle/knpemi_oracle.c:dp45_row (same formulas, same association; the
// time logic uses non-contracted adds/multiplies so both sides see bit-identical stage
// times).  Time-only factors are evaluated on the device per stage (M::tonly_dev) because the
// stage times are not known on the host.  A warp runs until its slowest lane has reached
// t_end; lanes that are done idle (DOF numbering is spatially coherent on real meshes, so
// quiescent and active DOFs mostly sit in different warps).
template <class M, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
kem_step_dp45_kernel(const __grid_constant__ KemArgs<M> a)
{
    constexpr int NS = M::NS, NOUT = M::NOUT, NT = M::NT;
    kem::load_tables<true>();          // the step-size controller takes a logarithm
    __syncthreads();
    const long long tid = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (tid >= a.n) return;
    // activity-sorted execution: neighbouring threads take DOFs that needed similar step
    // counts last time, so a warp is not held up by one active lane (kem_runtime.cu:
    // build_activity_perm).  The loads and stores below become sector-granular gathers;
    // at 2 % of HBM bandwidth that is free.
    const long long i = (a.perm && a.perm_on[0]) ? (long long)a.perm[tid] : tid;

    typename M::H q;
    kem_prologue<M>(a, i, q);

    double y[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) y[c] = a.y[c][i];

    constexpr double c2 = 1.0 / 5.0, c3 = 3.0 / 10.0, c4 = 4.0 / 5.0, c5 = 8.0 / 9.0;
    constexpr double a21 = 1.0 / 5.0;
    constexpr double a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
    constexpr double a41 = 44.0 / 45.0, a42 = -56.0 / 15.0, a43 = 32.0 / 9.0;
    constexpr double a51 = 19372.0 / 6561.0, a52 = -25360.0 / 2187.0, a53 = 64448.0 / 6561.0,
                     a54 = -212.0 / 729.0;
    constexpr double a61 = 9017.0 / 3168.0, a62 = -355.0 / 33.0, a63 = 46732.0 / 5247.0,
                     a64 = 49.0 / 176.0, a65 = -5103.0 / 18656.0;
    constexpr double b1 = 35.0 / 384.0, b3 = 500.0 / 1113.0, b4 = 125.0 / 192.0,
                     b5 = -2187.0 / 6784.0, b6 = 11.0 / 84.0;
    constexpr double e1 = 71.0 / 57600.0, e3 = -71.0 / 16695.0, e4 = 71.0 / 1920.0,
                     e5 = -17253.0 / 339200.0, e6 = 22.0 / 525.0, e7 = -1.0 / 40.0;

    const double dt = a.dt, t_end = a.t_end, rtol = a.rtol, atol = a.atol;
    double t = a.t0;
    double h_try = a.hsug[i];
    if (!(h_try > 0.0) || !isfinite(h_try)) h_try = dt / 8.0;
    if (h_try > dt) h_try = dt;

    double ts[NT > 0 ? NT : 1];
    double k1[NS], k2[NS], k3[NS], k4[NS], k5[NS], k6[NS], k7[NS], w[NS], yn[NS];
    M::tonly_dev(t, ts);
    M::deriv(y, k1, q, ts);

    bool rejected = false, failed = true;
    unsigned n_acc = 0, n_rej = 0;
#pragma unroll 1
    for (int attempt = 0; attempt < 100000; ++attempt) {
        double h = h_try;
        bool last = false;
        if (__dadd_rn(t, __dmul_rn(h, 1.0 + 1e-9)) >= t_end) {
            h = __dsub_rn(t_end, t);
            last = true;
        }
#pragma unroll
        for (int c = 0; c < NS; ++c) w[c] = y[c] + h * (a21 * k1[c]);
        M::tonly_dev(__dadd_rn(t, __dmul_rn(c2, h)), ts);
        M::deriv(w, k2, q, ts);
#pragma unroll
        for (int c = 0; c < NS; ++c) w[c] = y[c] + h * (a31 * k1[c] + a32 * k2[c]);
        M::tonly_dev(__dadd_rn(t, __dmul_rn(c3, h)), ts);
        M::deriv(w, k3, q, ts);
#pragma unroll
        for (int c = 0; c < NS; ++c) w[c] = y[c] + h * ((a41 * k1[c] + a42 * k2[c]) + a43 * k3[c]);
        M::tonly_dev(__dadd_rn(t, __dmul_rn(c4, h)), ts);
        M::deriv(w, k4, q, ts);
#pragma unroll
        for (int c = 0; c < NS; ++c)
            w[c] = y[c] + h * (((a51 * k1[c] + a52 * k2[c]) + a53 * k3[c]) + a54 * k4[c]);
        M::tonly_dev(__dadd_rn(t, __dmul_rn(c5, h)), ts);
        M::deriv(w, k5, q, ts);
#pragma unroll
        for (int c = 0; c < NS; ++c)
            w[c] = y[c] + h * ((((a61 * k1[c] + a62 * k2[c]) + a63 * k3[c]) + a64 * k4[c]) + a65 * k5[c]);
        const double t_new = last ? t_end : __dadd_rn(t, h);
        M::tonly_dev(t_new, ts);
        M::deriv(w, k6, q, ts);
#pragma unroll
        for (int c = 0; c < NS; ++c)
            yn[c] = y[c] + h * ((((b1 * k1[c] + b3 * k3[c]) + b4 * k4[c]) + b5 * k5[c]) + b6 * k6[c]);
        M::deriv(yn, k7, q, ts);
        double sum = 0.0;
#pragma unroll
        for (int c = 0; c < NS; ++c) {
            const double ei = h * (((((e1 * k1[c] + e3 * k3[c]) + e4 * k4[c]) + e5 * k5[c]) + e6 * k6[c])
                                   + e7 * k7[c]);
            const double sc = atol + rtol * fmax(fabs(y[c]), fabs(yn[c]));
            const double r = ei / sc;
            sum += r * r;
        }
        // err = sqrt(err2); the twin tests err <= 1, which is err2 <= 1 exactly (sqrt is monotone
        // and sqrt(1) = 1), and 0.9 err^(-1/5) = 0.9 exp(-0.1 log err2): no sqrt, no libm pow
        const double err2 = sum / (double)NS;
        const double shrink = 0.9 * kem::exp(-0.1 * kem::log(err2));
        if (err2 <= 1.0) {
            double factor = (err2 == 0.0) ? 10.0 : fmin(10.0, shrink);
            if (rejected) factor = fmin(1.0, factor);
            t = t_new;
#pragma unroll
            for (int c = 0; c < NS; ++c) { y[c] = yn[c]; k1[c] = k7[c]; }
            ++n_acc;
            rejected = false;
            const double h_next = h * factor;
            h_try = last ? fmax(h_next, h_try) : h_next;
            if (last) {
                failed = false;
                break;
            }
        } else {
            const double factor = (err2 == err2) ? fmax(0.2, shrink) : 0.2;
            h_try = h * factor;
            rejected = true;
            ++n_rej;
            if (!(h_try > 1e-14 * fabs(dt))) break;
        }
    }
    a.hsug[i] = failed ? 0.0 : (h_try > dt ? dt : h_try);

    // ---- current epilogue at (t_end, y): ts still holds the factors of t_end
    if (NOUT > 0) {
        double o[NOUT > 0 ? NOUT : 1];
        M::outputs(y, o, q, ts);
#pragma unroll
        for (int c = 0; c < NOUT; ++c) a.out[c][i] = o[c];
    }
    bool finite = true;
#pragma unroll
    for (int c = 0; c < NS; ++c) {
        a.y[c][i] = y[c];
        finite = finite && isfinite(y[c]);
    }
    if (!finite) atomicOr(a.flags, 1);
    if (failed) atomicOr(a.flags, 2);      // step-size control gave up (step limit / h underflow)

    // ---- step statistics: one atomic pair per warp
    if (a.stats) {
        const unsigned mask = __activemask();
        const unsigned acc = __reduce_add_sync(mask, n_acc);
        const unsigned rej = __reduce_add_sync(mask, n_rej);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(mask) - 1)) {
            atomicAdd(a.stats, (unsigned long long)acc);
            atomicAdd(a.stats + 1, (unsigned long long)rej);
        }
    }
}

template <class M, int BLOCK>
static cudaError_t kem_launch_block(const KemArgs<M> &a, size_t smem, cudaStream_t stream)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kem_step_kernel<M, BLOCK>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long grid = (a.n + BLOCK - 1) / BLOCK;
    kem_step_kernel<M, BLOCK><<<(unsigned)grid, BLOCK, smem, stream>>>(a);
    return cudaGetLastError();
}

template <class M, int BLOCK>
static cudaError_t kem_launch_dp45_block(const KemArgs<M> &a, cudaStream_t stream)
{
    const long long grid = (a.n + BLOCK - 1) / BLOCK;
    kem_step_dp45_kernel<M, BLOCK><<<(unsigned)grid, BLOCK, 0, stream>>>(a);
    return cudaGetLastError();
}

template <class M>
static cudaError_t kem_launch(const KemLaunch *L, cudaStream_t stream)
{
    if (L->n <= 0) return cudaSuccess;
    if (L->n_stim > KEM_MAX_STIM) return cudaErrorInvalidValue;
    if (L->scheme == 0 && L->n_sub < 1) return cudaErrorInvalidValue;
    if (L->scheme == 1 && (!L->hsug || !(L->rtol > 0.0) || !(L->atol >= 0.0))) return cudaErrorInvalidValue;
    if (L->scheme != 0 && L->scheme != 1) return cudaErrorInvalidValue;
    if ((L->n + 63) / 64 > 0x7fffffffLL) return cudaErrorInvalidValue;
    KemArgs<M> a;
    a.n = L->n;
    for (int c = 0; c < M::NS; ++c) a.y[c] = L->y[c];
    for (int c = 0; c < M::NP; ++c) { a.p[c] = L->p[c]; a.pmask[c] = L->pmask[c]; }
    for (int c = 0; c < M::NOUT; ++c) a.out[c] = L->out[c];
    a.stim_mask = L->stim_mask;
    a.n_stim = L->stim_mask ? L->n_stim : 0;
    for (int s = 0; s < KEM_MAX_STIM; ++s) {
        a.stim_col[s] = L->stim_col[s];
        a.stim_val[s] = L->stim_val[s];
        a.stim_ptr[s] = L->stim_ptr[s];
    }
    a.ttab = L->ttab;
    a.n_sub = L->n_sub;
    a.h = L->h;
    a.flags = L->flags;
    a.t0 = L->t0; a.dt = L->dt; a.t_end = L->t_end; a.rtol = L->rtol; a.atol = L->atol;
    a.hsug = L->hsug; a.stats = L->stats; a.perm = L->perm; a.perm_on = L->perm_on;
    if (L->scheme == 1) {
        const int blk = L->block ? L->block : M::DEFAULT_BLOCK;
        switch (blk) {
            case 64:  return kem_launch_dp45_block<M, 64>(a, stream);
            case 128: return kem_launch_dp45_block<M, 128>(a, stream);
            case 256: return kem_launch_dp45_block<M, 256>(a, stream);
            default:  return cudaErrorInvalidValue;
        }
    }
    const size_t smem = (size_t)(2 * L->n_sub + 2) * M::NT * sizeof(double);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const int block = L->block ? L->block : M::DEFAULT_BLOCK;
    switch (block) {
        case 64:  return kem_launch_block<M, 64>(a, smem, stream);
        case 128: return kem_launch_block<M, 128>(a, smem, stream);
        case 256: return kem_launch_block<M, 256>(a, smem, stream);
        default:  return cudaErrorInvalidValue;
    }
}

template <class M>
static cudaError_t kem_launch_info(int *regs, int *max_blocks_per_sm, int block)
{
    cudaFuncAttributes fa;
    cudaError_t e;
    if (block == 0) block = M::DEFAULT_BLOCK;
    int nb = 0;
    switch (block) {
        case 64:
            e = cudaFuncGetAttributes(&fa, kem_step_kernel<M, 64>);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kem_step_kernel<M, 64>, 64, 0);
            break;
        case 128:
            e = cudaFuncGetAttributes(&fa, kem_step_kernel<M, 128>);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kem_step_kernel<M, 128>, 128, 0);
            break;
        case 256:
            e = cudaFuncGetAttributes(&fa, kem_step_kernel<M, 256>);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kem_step_kernel<M, 256>, 256, 0);
            break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    *regs = fa.numRegs;
    *max_blocks_per_sm = nb;
    return cudaSuccess;
}

// device-side numpy.mod (only reached if a model takes mod of a state-dependent value)
__device__ __forceinline__ double kem_npmod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((b < 0.0) != (r < 0.0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

static inline double kem_npmod_host(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((b < 0.0) != (r < 0.0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

#define KEM_DEFINE_MODEL(M, NAME_STR, HASH_STR, OUT_COLS, USED_COLS, N_USED, N_CONST, CONST_COLS,  \
                         CONST_VALS)                                                          \
    extern "C" __attribute__((visibility("default"))) const KemModelDesc *kem_model_descriptor(void) \
    {                                                                                         \
        static KemModelDesc d = {KEM_MODEL_ABI_VERSION, NAME_STR, HASH_STR, M::NS, M::NP,     \
                                 M::NOUT, OUT_COLS, N_USED, USED_COLS, M::NT,                 \
                                 &M::tonly, &kem_launch<M>, 0, &kem_launch_info<M>,           \
                                 N_CONST, CONST_COLS, CONST_VALS};                            \
        return &d;                                                                            \
    }
