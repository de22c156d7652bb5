/*
 * knpemi_b200.h -- C ABI of libknpemi_b200.so, the B200 (sm_100a) backend of the
 * membrane-ODE stage of knpemi.
 *
 * This is the drop-in boundary.  Each entry point names the
 * piece of the reference's MembraneModel (src/knpemi/odeSolver.py, cited as
 * odeSolver.py:LINE) it replaces.  Plain pointers and sizes only: host buffers
 * are caller-owned and may come from NumPy (`arr.ctypes.data`), DLPack or
 * kem_host_alloc().  Every function returns 0 on success, a negative KEM_E_*
 * code on argument / CUDA errors (text via kem_last_error()), and kem_step*
 * return a positive code when the integration failed -- KEM_NONFINITE (a state is
 * not finite) or KEM_STEP_FAILED (the error-controlled scheme could not reach
 * t0+dt) -- the counterpart of the reference's `assert success` (odeSolver.py:121).
 *
 * There is no CPU fallback: without a CUDA device kem_create() fails.
 * Thread-safety: one host thread per handle.
 */
#ifndef KNPEMI_B200_H
#define KNPEMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KEM_OK 0
#define KEM_NONFINITE 1      /* a state became non-finite */
#define KEM_STEP_FAILED 2    /* KEM_SCHEME_DP45: step-size control gave up before t0+dt */
#define KEM_E_ARG (-1)
#define KEM_E_CUDA (-2)
#define KEM_E_MODEL (-3)
#define KEM_E_NOMEM (-4)

#define KEM_STATE 0      /* `what == 'state'`     odeSolver.py:133 */
#define KEM_PARAM 1      /* `what == 'parameter'` odeSolver.py:134 */

#define KEM_SCHEME_RK4 0  /* scheme O1: classical RK4, n_sub sub-steps + current epilogue */
#define KEM_SCHEME_DP45 1 /* scheme O3: Dormand-Prince 5(4), per-DOF error-controlled steps
                             (the reference's LSODA is error-controlled too: odeSolver.py:116-120) */

typedef struct kem_handle_s *kem_handle;

typedef struct kem_model_info {
    int ns;          /* states per DOF        (len(init_state_values()),     odeSolver.py:41) */
    int np;          /* parameters per DOF    (len(init_parameter_values()), odeSolver.py:42) */
    int n_out;       /* parameter slots the RHS writes (I_ch_*; mm_hh.py:220-225) */
    int n_used;      /* parameter slots the RHS reads */
    int n_tslots;    /* host-evaluated time-only factors per stage time */
    int out_cols[64];
    char name[64];
    char source_hash[32];
} kem_model_info;

typedef struct kem_step_times {
    double ms_h2d;      /* host->device copies of the step's input columns  */
    double ms_kernel;   /* fused step kernel, max over the handle's devices */
    double ms_d2h;      /* device->host copies of the step's output columns */
    double ms_total;    /* first copy enqueued -> last copy landed, max over devices */
} kem_step_times;

/* one column moved by kem_step_io(): table[:, col] <-> host[0:n_dof] */
typedef struct kem_io_column {
    int kind;        /* KEM_STATE | KEM_PARAM */
    int col;
    double *host;    /* n_dof doubles */
} kem_io_column;

/* ---- library ---------------------------------------------------------------- */
int kem_version(void);
const char *kem_last_error(void);
int kem_device_count(int *n_out);

/* ---- models: replaces `ode.rhs_numba.address` (odeSolver.py:96) -------------- */
/* Load a generated model library (knpemi_b200.codegen) and register it. */
int kem_model_load(const char *so_path, int *model_id_out);
int kem_model_find(const char *name_or_hash, int *model_id_out);
int kem_model_get_info(int model_id, kem_model_info *out);
/* registers/thread and resident blocks/SM of the model's step kernel on device `dev` */
int kem_model_launch_info(int model_id, int dev, int block, int *regs_out, int *blocks_per_sm_out);

/* ---- construction: MembraneModel.__init__ (odeSolver.py:8-49) ---------------- */
/* Allocates the SoA tables for n_dof DOFs, split in contiguous ranges over the
 * n_dev devices dev_ids[] (NULL = device 0), every row initialised to the
 * model defaults (odeSolver.py:41-42). */
int kem_create(int model_id, int64_t n_dof, int n_dev, const int *dev_ids,
               const double *state_defaults, const double *param_defaults, kem_handle *out);
int kem_destroy(kem_handle h);
int kem_n_dof(kem_handle h, int64_t *n_out);
/* DOF range [begin, end) owned by the k-th device of the handle */
int kem_shard_range(kem_handle h, int k, int *dev_out, int64_t *begin_out, int64_t *end_out);

/* ---- table access: __set_ODE / __get_PDE / __set_ODE_values ------------------- */
/* table[:, col] = v for every DOF (odeSolver.py:183-187 with a constant value and no locator) */
int kem_set_uniform(kem_handle h, int kind, int col, double v);
/* table[:, col] = host_src[0:n]                      (odeSolver.py:142-144, locator None) */
int kem_set_column(kem_handle h, int kind, int col, const double *host_src, int64_t n);
/* table[mask, col] = host_src[mask]                  (odeSolver.py:138-144 with a locator) */
int kem_set_column_masked(kem_handle h, int kind, int col, const double *host_src,
                          const uint8_t *host_mask, int64_t n);
/* table[mask, col] = v                               (odeSolver.py:183-187, constant value) */
int kem_set_value_masked(kem_handle h, int kind, int col, double v, const uint8_t *host_mask,
                         int64_t n);
/* host_dst[0:n] = table[:, col]                      (odeSolver.py:159-164) */
int kem_get_column(kem_handle h, int kind, int col, double *host_dst, int64_t n);
/* 1 if the column is stored as one value for all DOFs */
int kem_column_is_uniform(kem_handle h, int kind, int col, int *is_uniform_out, double *value_out);

/* where a column currently lives: 0 = one value for all DOFs, 1 = per-DOF column in HBM,
 * 2 = per-DOF host shadow, 3 = discarded.  A parameter slot the generated right-hand side
 * neither reads nor writes (for the HH models: Cl_e, Cl_i) need not cross the host link; what
 * kem_set_column / kem_step_io do with a full-column write to such a slot is the policy below. */
int kem_column_location(kem_handle h, int kind, int col, int *location_out);

/* Policy for full-column writes to parameter slots the right-hand side never touches
 * (the reference stores them like any other column, odeSolver.py:142, and never reads them):
 *   KEM_UNREAD_SHADOW   keep the values in a host-side shadow copy (getters return them; they
 *                       are uploaded if a masked setter / stimulus / device gather needs them)
 *   KEM_UNREAD_UPLOAD   treat the slot like any other column (one more DMA per write)
 *   KEM_UNREAD_DISCARD  neither copy nor upload: the value is dropped; reading the column
 *                       afterwards fails with KEM_E_ARG until it is written again under another
 *                       policy.  For callers that push the PDE state every step and never read
 *                       these slots back (update_ode_variables, utils.py:227-228).
 *   KEM_UNREAD_AUTO     (default) pageable sources are shadowed; pinned sources are uploaded by
 *                       kem_set_column (the idle link beats a host copy) and, by kem_step_io,
 *                       shadowed with up to two GPUs on the host and uploaded with more (there
 *                       the host memory system is the bottleneck). */
#define KEM_UNREAD_AUTO 0
#define KEM_UNREAD_SHADOW 1
#define KEM_UNREAD_UPLOAD 2
#define KEM_UNREAD_DISCARD 3
int kem_set_unread_policy(kem_handle h, int policy);

/* ---- stimulus mask: `stimulus_mask` of step_lsoda (odeSolver.py:98-100) ------- */
/* Upload the 0/1 mask used by the next kem_step calls; NULL = every DOF (the
 * reference's default locator `lambda x: True`). */
int kem_set_stimulus_mask(kem_handle h, const uint8_t *host_mask_or_null, int64_t n);

/* ---- the step: MembraneModel.step_lsoda row loop (odeSolver.py:106-123) ------- */
/* Advance every DOF from t0 to t0+dt.  The n_stim (column, value) pairs are the
 * `stimulus` dict, written stickily into the parameter table under the current
 * stimulus mask (odeSolver.py:110-112).  status_flags != NULL: wait for the
 * kernel and report (bit 0: non-finite state, bit 1: DP45 step control failed); NULL: enqueue only, errors
 * surface at the next synchronising call. */
int kem_step(kem_handle h, double t0, double dt, int n_sub, int scheme,
             int n_stim, const int *stim_cols, const double *stim_vals, int *status_flags);
/* same, timed with CUDA events on the launching streams */
int kem_step_timed(kem_handle h, double t0, double dt, int n_sub, int scheme,
                   int n_stim, const int *stim_cols, const double *stim_vals,
                   int *status_flags, kem_step_times *times_out);
/* One coupled PDE/ODE exchange (utils.py:217-233 + step + run_2D.py:105-109):
 * copy n_in host columns in, step, copy n_out host columns out, pipelined in
 * DOF chunks over each device's streams.  Synchronous, except when every host buffer is
 * page-locked and both status_flags and times_out are NULL: then the call only enqueues
 * (the buffers must stay untouched until kem_sync or a kem_get_column, which follows the
 * kernel chunk by chunk).  Output slots the generated code assigns a literal (I_ch_Cl = 0.0,
 * mm_hh.py:225) and uniform columns are filled on the host instead of copied back. */
int kem_step_io(kem_handle h, double t0, double dt, int n_sub, int scheme,
                int n_stim, const int *stim_cols, const double *stim_vals,
                int n_in, const kem_io_column *in, int n_out, const kem_io_column *out,
                int *status_flags, kem_step_times *times_out);
int kem_sync(kem_handle h);
/* Launch the KEM_SCHEME_RK4 kernel of kem_step as `n_chunks` DOF chunks (two
 * compute streams) instead of one grid: a kem_get_column into page-locked memory that
 * follows then copies chunk c while chunk c+1 still computes.  1 (default) = one launch. */
int kem_set_step_chunks(kem_handle h, int n_chunks);
/* Tuning of kem_step_io's pipeline for this handle: DOF chunks per device (0 = default 16 /
 * KNPEMI_IO_CHUNKS) and host->device copy streams the input columns alternate over (0 = default
 * 2 / KNPEMI_IO_H2D_STREAMS, 1, 2).  Results do not depend on either. */
int kem_set_io_tuning(kem_handle h, int n_chunks, int h2d_streams);
/* The DOF chunks kem_step_io / a chunked kem_step cut a range of n DOFs into (offsets and
 * lengths, at most `cap` written, the count always returned): `target` equal chunks; with
 * taper = 1 the tail is halved repeatedly so that the pipeline drains through a small last
 * chunk (taper = -1: the runtime's default, off unless KNPEMI_IO_TAPER=1).  No device needed. */
int kem_plan_chunks(int64_t n, int target, int taper, int64_t *off_out, int64_t *len_out, int cap,
                    int *count_out);
/* tolerances of KEM_SCHEME_DP45; defaults are the reference's rtol 1e-8, atol 1e-10
 * (odeSolver.py:120) */
int kem_set_tolerances(kem_handle h, double rtol, double atol);
/* KEM_SCHEME_DP45 runs the DOFs in an order sorted by the step size each used last time, so
 * that a warp's lanes finish together (default on; results do not depend on it) */
int kem_set_activity_sort(kem_handle h, int enabled);
/* accepted / rejected DP45 steps summed over all DOFs since the last call (waits for the
 * device); RHS evaluations = 6 * (accepted + rejected) + 1 per DOF-step */
int kem_get_step_stats(kem_handle h, uint64_t *accepted_out, uint64_t *rejected_out);
/* launch configuration knobs: threads per block (64/128/256; 0 = model default) */
int kem_set_block(kem_handle h, int block);
/* number of kernels this handle has launched since creation */
int kem_launch_count(kem_handle h, int64_t *n_out);

/* CUDA-event stopwatch on the handle's launching streams: begin records an event on
 * every device's stream, end records another, waits, and returns the largest elapsed
 * time over the devices -- brackets K enqueue-only kem_step calls in the benchmark. */
int kem_timer_begin(kem_handle h);
int kem_timer_end(kem_handle h, double *ms_out);

/* ---- device-resident PDE vectors -------------------------
 * For a PDE side whose coefficient vectors already live on the GPU: the seven setter
 * copies and four getter copies of one PDE step (utils.py:217-233, run_2D.py:105-109)
 * become index gathers / scatters over membrane-DOF -> bulk-DOF maps (a CG-1 trace is a
 * vertex copy: utils.py:150-207), with no host traffic.  Maps are registered once;
 * `shard` selects the handle's k-th device, and the device pointers must live there.
 * Ordering: these calls enqueue on the handle's own stream; the caller makes sure its
 * producer of `dev_src` has finished (synchronise that stream or event) before calling, and
 * kem_device_scatter / kem_device_copy_out return after the data has landed. */
/* register map `map_id` (0..15): bulk index of every membrane DOF, n = n_dof entries */
int kem_device_map_set(kem_handle h, int map_id, const int64_t *host_map, int64_t n);
/* table[i, col] = dev_src[map[i]]             (update_ode_variables, utils.py:224-228) */
int kem_device_gather(kem_handle h, int shard, int kind, int col, const double *dev_src, int map_id);
/* dev_dst[map[i]] = table[i, col]             (get_membrane_potential / get_parameter) */
int kem_device_scatter(kem_handle h, int shard, int kind, int col, double *dev_dst, int map_id);
/* table[i, col] = dev_a[map_a[i]] - dev_b[map_b[i]]   (phi_M = tr(phi_i) - tr(phi_e), utils.py:247-293) */
int kem_device_gather_diff(kem_handle h, int shard, int kind, int col, const double *dev_a,
                           int map_a, const double *dev_b, int map_b);
/* contiguous device-to-device column copies for the DOFs of shard k (no map):
 * table[begin_k:end_k, col] = dev_src[0:n_k]   and   dev_dst[0:n_k] = table[begin_k:end_k, col] */
/* out[i] = a0 + sum_k coef[k] * in_k[i] over a bulk vector of n DOFs on device `dev`: the
 * eliminated-ion concentration of update_pde_variables (utils.py:247-267),
 * c_elim = -(1/z_e) (rho_z rho_tag + sum_k z_k c_k), summed in the order given.  At most 8
 * terms; synchronous on the default stream. */
int kem_device_affine_combine(int dev, int64_t n, double *dev_out, double a0, int n_terms,
                              const double *coef, const double *const *dev_in);
/* table[i, col] = a0 + sum_k coef[k] * in_k[map[i]]: the membrane trace of that combination
 * (what update_ode_variables pushes for the eliminated ion, utils.py:219-228) without forming
 * the bulk vector first.  Enqueued on the handle's stream like kem_device_gather. */
int kem_device_gather_affine(kem_handle h, int shard, int kind, int col, double a0, int n_terms,
                             const double *coef, const double *const *dev_in, int map_id);
int kem_device_copy_in(kem_handle h, int shard, int kind, int col, const double *dev_src);
int kem_device_copy_out(kem_handle h, int shard, int kind, int col, double *dev_dst);
/* plain device buffers for callers without their own CUDA allocations (tests, Python hosts) */
int kem_device_alloc(int dev, size_t bytes, void **ptr_out);
int kem_device_free(int dev, void *ptr);
int kem_device_upload(int dev, void *dev_dst, const void *host_src, size_t bytes);
int kem_device_download(int dev, void *host_dst, const void *dev_src, size_t bytes);

/* ---- pinned host memory for callers that want zero-staging transfers ---------- */
int kem_host_alloc(void **ptr_out, size_t bytes);
int kem_host_free(void *ptr);
/* Page-lock memory the caller owns -- the `u.x.array` the reference's setters read
 * (odeSolver.py:142) and its getters write (odeSolver.py:159-164) -- so that
 * kem_set_column / kem_get_column / kem_step_io copy it directly instead of through
 * the staging buffers.  Registrations are reference-counted per base address over all handles
 * of the process (registering twice needs two unregisters); a range that overlaps a registered
 * one, or memory another owner page-locked, is refused with KEM_E_ARG; unregistering unknown
 * memory is a no-op.  The memory must be unregistered before it is freed. */
int kem_host_register(void *ptr, size_t bytes);
int kem_host_unregister(void *ptr);
/* 1 if transfers from/to the WHOLE range [ptr, ptr+bytes) take the direct (page-locked) path,
 * 0 if they are staged (cudaHostRegister pins pages: the first byte alone proves nothing). */
int kem_host_is_pinned(const void *ptr, size_t bytes, int *pinned_out);

/* ---- measurement helpers ------------------------------------------------------ */
/* Dependent-chain-free DFMA micro-benchmark: the measured FP64 pipe peak of
 * device `dev` in TFLOP/s (2 flops per DFMA), used as roofline denominator. */
int kem_fp64_peak(int dev, double *tflops_out, double *ms_out);
/* Device-to-device copy bandwidth of device `dev` in GB/s (read + write bytes). */
int kem_hbm_copy_peak(int dev, double *gbs_out);
/* Host-link ceiling of device `dev`, the denominator of the end-to-end exchange
 * (the 7-in / 4-out column traffic of utils.py:217-233 + run_2D.py:105-109):
 * `reps_h2d` copies of `bytes` host->device and `reps_d2h` copies device->host run
 * concurrently on two streams from pinned memory of the calling thread, walking through
 * `span_bytes` of host memory per direction (0 = one buffer reused: cache-resident, the link
 * alone; far above the last-level cache: through DRAM, like a real exchange).  Each direction's
 * elapsed milliseconds are returned (0 reps = that direction idle).  Unequal rep counts measure
 * the shorter direction entirely under the other's load. */
int kem_link_probe(int dev, size_t bytes, size_t span_bytes, int reps_h2d, int reps_d2h,
                   double *ms_h2d_out, double *ms_d2h_out);

#ifdef __cplusplus
}
#endif
#endif /* KNPEMI_B200_H */
