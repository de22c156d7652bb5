// kem_model_api.h -- contract between the runtime (libknpemi_b200.so) and a
// generated model library (libkem_<model>_<hash>.so).
//
// A model library is produced by knpemi_b200.codegen from the Python source of
// a membrane model's `rhs_numba` (plugin protocol: reference
// examples/idealized_geometries/mm_hh.py:133-139) and exports ONE C symbol,
// `kem_model_descriptor`, returning the table below.  The runtime owns all
// device memory and streams; the model library only knows how to evaluate the
// host-side time-only factors and how to launch its fused step kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KEM_MODEL_ABI_VERSION 7
#define KEM_MAX_STIM 4

extern "C" {

// Everything one launch of the fused step kernel needs.  Pointer arrays are
// HOST arrays of DEVICE pointers (the launcher copies them into the kernel's
// parameter block).
typedef struct KemLaunch {
    int64_t n;                      // DOFs in this device's contiguous range
    double *const *y;               // [ns]  state columns (SoA), read + written
    const double *const *p;         // [np]  parameter column, or address of the uniform slot
    const int64_t *pmask;           // [np]  ~0 = per-DOF column, 0 = uniform (index i & mask)
    double *const *out;             // [n_out] output parameter columns (I_ch_*), written
    const uint8_t *stim_mask;       // per-DOF 0/1, or NULL (stimulus already folded into p)
    int n_stim;                     // sticky stimulus entries applied under stim_mask
    int stim_col[KEM_MAX_STIM];     //   parameter column
    double stim_val[KEM_MAX_STIM];  //   value
    double *stim_ptr[KEM_MAX_STIM]; //   that parameter's per-DOF column (written back: sticky)
    const double *ttab;             // [(2*n_sub+2) * n_tslots] host-evaluated time-only factors
    int n_sub;                      // RK4 sub-steps per PDE step
    double h;                       // dt / n_sub
    int *flags;                     // device int, OR-ed with 1 if any end state is non-finite
    int block;                      // threads per block (0 = model default)
    int scheme;                     // 0 = RK4 x n_sub (O1), 1 = Dormand-Prince 5(4) (O3)
    double t0, dt, t_end;           // O3: the PDE step [t0, t_end], t_end = t0 + dt formed on the host
    double rtol, atol;              // O3: error tolerances
    double *hsug;                   // O3: per-DOF warm-start step size (read + written)
    unsigned long long *stats;      // O3: device counters [accepted steps, rejected steps]
    const int *perm;                // O3: thread -> DOF permutation grouping DOFs of similar
                                    //     activity into the same warps (NULL = identity)
    const unsigned *perm_on;        // O3: device flag, 0 = ignore perm (all DOFs equally active)
} KemLaunch;

typedef struct KemModelDesc {
    int abi_version;
    const char *name;               // model module name
    const char *source_hash;        // hash of the generated source
    int ns, np;
    int n_out;                      // parameter slots the RHS writes
    const int *out_cols;            // [n_out] ascending
    int n_used;                     // parameter slots the RHS reads
    const int *used_cols;           // [n_used]
    int n_tslots;                   // time-only factors per stage time
    // host: evaluate the time-only factors at time t (glibc libm)
    void (*tonly)(double t, double *slots);
    // enqueue the fused step kernel on `stream`
    cudaError_t (*launch)(const KemLaunch *args, cudaStream_t stream);
    // static launch facts, for reporting
    int regs_per_thread;            // filled lazily by launch_info
    cudaError_t (*launch_info)(int *regs, int *max_blocks_per_sm, int block);
    // output slots whose right-hand side is a literal (I_ch_Cl = 0.0 in the HH models,
    // mm_hh.py:225): the kernel still stores them, but nothing has to ask the device for them
    int n_const_out;
    const int *const_out_cols;      // [n_const_out] parameter columns, subset of out_cols
    const double *const_out_vals;   // [n_const_out]
} KemModelDesc;

typedef const KemModelDesc *(*kem_model_descriptor_fn)(void);

}  // extern "C"
