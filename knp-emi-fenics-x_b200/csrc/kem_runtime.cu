// kem_runtime.cu -- host runtime of libknpemi_b200.so: model registry, SoA
// table storage in HBM, contiguous DOF ranges over 1..8 B200s, streams, pinned
// staging, the step driver, and the small utility / measurement kernels.
//
// C ABI: include/knpemi_b200.h (each entry point cites the part of the
// reference's src/knpemi/odeSolver.py it replaces).  No PyTorch, no CPU
// fallback: every path below ends in a CUDA call on the handle's devices.
#include "../../include/knpemi_b200.h"
#include "kem_model_api.h"
#include "kem_copy_pool.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a tool is attached
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            char b__[512];                                                               \
            snprintf(b__, sizeof b__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,      \
                     cudaGetErrorString(e__));                                           \
            return fail(KEM_E_CUDA, b__);                                                \
        }                                                                                \
    } while (0)

#define ARG(cond, msg)                                                                   \
    do {                                                                                 \
        if (!(cond)) return fail(KEM_E_ARG, std::string(__func__) + ": " + (msg));       \
    } while (0)

// NVTX range covering one C-ABI call (the counterpart of the reference's
// dolfinx.common.Timer('ODE step LSODA'), odeSolver.py:104)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ------------------------------------------------------------------ utility kernels
__global__ void k_fill(double *__restrict__ dst, long long n, double v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = v;
}

__global__ void k_set_value_masked(double *__restrict__ dst, const unsigned char *__restrict__ m,
                                   long long n, double v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        if (m[i]) dst[i] = v;
}

__global__ void k_copy_masked(double *__restrict__ dst, const double *__restrict__ src,
                              const unsigned char *__restrict__ m, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        if (m[i]) dst[i] = src[i];
}

__global__ void k_copy(double2 *__restrict__ dst, const double2 *__restrict__ src, long long n2)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) dst[i] = src[i];
}

// f1 / f3: membrane <- bulk gathers and bulk <- membrane scatters (HBM-bound, 8 B + 8 B index
// read and 8 B written per DOF; the membrane side is coalesced, the bulk side follows the map)
__global__ void k_gather(double *__restrict__ dst, const double *__restrict__ src,
                         const long long *__restrict__ map, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[map[i]];
}

__global__ void k_scatter(double *__restrict__ dst, const double *__restrict__ src,
                          const long long *__restrict__ map, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[map[i]] = src[i];
}

__global__ void k_gather_diff(double *__restrict__ dst, const double *__restrict__ a,
                              const long long *__restrict__ map_a, const double *__restrict__ b,
                              const long long *__restrict__ map_b, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = a[map_a[i]] - b[map_b[i]];
}

// ---- activity sort for scheme O3 (error-controlled stepping) ---------------------------
// bucket = quarter-octaves of dt/hsug (about the number of steps the DOF took last time),
// 0 for a DOF that has no history; 64 buckets cover up to 2^16 steps per PDE step.
constexpr int ACT_BUCKETS = 64;

__device__ __forceinline__ int activity_bucket(double hsug, double dt)
{
    if (!(hsug > 0.0) || !(hsug < dt)) return 0;
    const int b = (int)(4.0 * log2(dt / hsug) + 0.5);
    return b < 0 ? 0 : (b >= ACT_BUCKETS ? ACT_BUCKETS - 1 : b);
}

__global__ void k_activity_hist(const double *__restrict__ hsug, double dt, long long n,
                                unsigned *__restrict__ counts)
{
    __shared__ unsigned s_cnt[ACT_BUCKETS];
    for (int k = threadIdx.x; k < ACT_BUCKETS; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) atomicAdd(&s_cnt[activity_bucket(hsug[i], dt)], 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < ACT_BUCKETS; k += blockDim.x)
        if (s_cnt[k]) atomicAdd(&counts[k], s_cnt[k]);
}

// exclusive scan of the bucket counts -> running cursors; most active bucket first, so the
// long-running warps start early and the short ones fill the tail of the launch
__global__ void k_activity_scan(const unsigned *__restrict__ counts, unsigned *__restrict__ cursor)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned run = 0, used = 0;
        for (int k = ACT_BUCKETS - 1; k >= 0; --k) {
            cursor[k] = run;
            run += counts[k];
            used += counts[k] != 0;
        }
        cursor[ACT_BUCKETS] = used > 1;     // all DOFs equally active: keep the identity order
    }
}

__global__ void k_activity_scatter(const double *__restrict__ hsug, double dt, long long n,
                                   unsigned *__restrict__ cursor, int *__restrict__ perm)
{
    if (!cursor[ACT_BUCKETS]) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int b = activity_bucket(hsug[i], dt);
        // one atomic per (warp, bucket): lanes with the same bucket share a reservation
        const unsigned active = __activemask();
        const unsigned peers = __match_any_sync(active, b);
        const int leader = __ffs(peers) - 1;
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(&cursor[b], (unsigned)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        perm[base + __popc(peers & ((1u << lane) - 1u))] = (int)i;
    }
}

// FP64 pipe peak: 8 independent DFMA chains per thread, nothing else in the loop.
constexpr int PEAK_CHAINS = 8;
constexpr int PEAK_ITERS = 8192;
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, double a, double b)
{
    double x[PEAK_CHAINS];
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; ++c) x[c] = 1.0 + 1e-3 * (threadIdx.x + c);
#pragma unroll 1
    for (int it = 0; it < PEAK_ITERS; it += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < PEAK_CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; ++c) s += x[c];
    if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

int grid_for(long long n, int block = 256)
{
    long long g = (n + block - 1) / block;
    return (int)std::max(1LL, std::min(g, 148LL * 16));
}

// ------------------------------------------------------------------ model registry
struct LoadedModel {
    const KemModelDesc *desc;
    void *dl;
    std::string path;
};
std::vector<LoadedModel> g_models;
std::mutex g_models_mu;

const KemModelDesc *model_desc(int id)
{
    std::lock_guard<std::mutex> lk(g_models_mu);
    if (id < 0 || id >= (int)g_models.size()) return nullptr;
    return g_models[id].desc;
}

// ------------------------------------------------------------------ handle
constexpr int SMALL_RING = 8;
constexpr size_t SMALL_BYTES = 64 * 1024;
constexpr int N_STAGE = 3;
constexpr size_t STAGE_BYTES = 8u << 20;
constexpr int IO_MAX_CHUNKS = 64;
constexpr int KEM_MAX_MAPS = 16;
constexpr int IO_TARGET_CHUNKS = 16;   // measured best of 8/16/32/64 at 1e7 DOFs (profiles/r1_bench.md)

struct Shard {
    int dev = 0;
    int64_t begin = 0, n = 0;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    cudaStream_t stream2 = nullptr;   // second compute stream: chunk kernels of kem_step_io alternate
                                      // between the two so one chunk's tail overlaps the next one's head
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;   // kem_timer_begin / kem_timer_end
    std::vector<double *> ycol;   // ns per-DOF state columns
    std::vector<double *> pcol;   // np per-DOF parameter columns (allocation cached)
    double *d_uni = nullptr;      // np uniform parameter values
    unsigned char *d_mask = nullptr;
    bool has_mask = false;
    double *d_ttab = nullptr;
    size_t ttab_cap = 0;
    int *d_flags = nullptr;
    int *h_flags = nullptr;       // pinned
    // small pinned ring for time tables / uniform tables
    void *h_small[SMALL_RING] = {};
    cudaEvent_t small_ev[SMALL_RING] = {};
    bool small_busy[SMALL_RING] = {};
    int small_next = 0;
    // pinned staging for pageable host columns
    void *h_stage[N_STAGE] = {};
    cudaEvent_t stage_ev[N_STAGE] = {};
    // per-chunk events of kem_step_io
    std::vector<cudaEvent_t> io_in, io_k0, io_k1, io_out;
    // scheme O3: per-DOF warm-start step size, device counters [accepted, rejected]
    double *d_hsug = nullptr;
    int *d_perm = nullptr;                   // activity-sorted thread -> DOF map
    unsigned *d_act = nullptr;               // [2 * ACT_BUCKETS] counts, cursors
    bool perm_valid = false;
    unsigned long long *d_stats = nullptr;
    unsigned long long *h_stats = nullptr;   // pinned
    // membrane-DOF -> bulk-DOF maps of the device-resident exchange (f1/f3)
    long long *d_map[KEM_MAX_MAPS] = {};
};

}  // namespace

struct kem_handle_s {
    const KemModelDesc *m = nullptr;
    int model_id = -1;
    int64_t n = 0;
    std::vector<Shard> shards;
    std::vector<double> uni;          // np: value of uniform parameter columns
    std::vector<char> p_uniform;      // np: 1 = stored as one value
    // Parameter slots the generated right-hand side never reads (HH: Cl_e, Cl_i; the I_ch_*
    // inputs) need not travel to the GPU at all: a full-column write to such a slot is kept
    // in a host shadow (p_host = 1) and only uploaded if something on the device asks for it.
    std::vector<char> p_dead;         // np: 1 = neither read nor written by the RHS
    std::vector<char> p_host;         // np: 1 = current value lives in p_shadow, not on the device
    std::vector<std::vector<double>> p_shadow;
    bool shadow_pinned_io = true;     // kem_step_io: pinned inputs to dead slots go to the shadow too
    bool uni_dirty = true;
    int block = 0;
    int64_t launches = 0;
    double rtol = 1.0e-8, atol = 1.0e-10;   // odeSolver.py:120
    bool activity_sort = true;              // scheme O3: group DOFs of similar activity into warps
};

namespace {

int ensure_stage(Shard &s)
{
    if (s.h_stage[0]) return KEM_OK;
    CK(cudaSetDevice(s.dev));
    for (int k = 0; k < N_STAGE; ++k) {
        CK(cudaHostAlloc(&s.h_stage[k], STAGE_BYTES, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&s.stage_ev[k], cudaEventDisableTiming));
    }
    return KEM_OK;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// host -> device, enqueued on `st`; pageable sources go through the pinned
// staging ring (the source is fully consumed when this returns).
int copy_in(Shard &s, double *dst, const double *src, size_t bytes, cudaStream_t st, bool pinned)
{
    if (bytes == 0) return KEM_OK;
    CK(cudaSetDevice(s.dev));
    if (pinned) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return KEM_OK;
    }
    int rc = ensure_stage(s);
    if (rc) return rc;
    size_t off = 0;
    int k = 0;
    while (off < bytes) {
        const size_t len = std::min(STAGE_BYTES, bytes - off);
        const int slot = k % N_STAGE;
        if (k >= N_STAGE) CK(cudaEventSynchronize(s.stage_ev[slot]));
        CopyPool::get().copy(s.h_stage[slot], (const char *)src + off, len);
        CK(cudaMemcpyAsync((char *)dst + off, s.h_stage[slot], len, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(s.stage_ev[slot], st));
        off += len;
        ++k;
    }
    // the staging slots may be reused by the next call: wait for the DMAs
    for (int j = 0; j < std::min(k, N_STAGE); ++j) CK(cudaEventSynchronize(s.stage_ev[j]));
    return KEM_OK;
}

// device -> host; returns after the data is in `dst` when pageable, enqueued only when pinned
int copy_out(Shard &s, double *dst, const double *src, size_t bytes, cudaStream_t st, bool pinned)
{
    if (bytes == 0) return KEM_OK;
    CK(cudaSetDevice(s.dev));
    if (pinned) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        return KEM_OK;
    }
    int rc = ensure_stage(s);
    if (rc) return rc;
    const size_t n_chunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    for (size_t k = 0; k < n_chunks + N_STAGE; ++k) {
        const int slot = (int)(k % N_STAGE);
        if (k >= (size_t)N_STAGE) {
            const size_t kk = k - N_STAGE;   // chunk that used this slot before
            if (kk < n_chunks) {
                CK(cudaEventSynchronize(s.stage_ev[slot]));
                const size_t off = kk * STAGE_BYTES;
                CopyPool::get().copy((char *)dst + off, s.h_stage[slot], std::min(STAGE_BYTES, bytes - off));
            }
        }
        if (k < n_chunks) {
            const size_t off = k * STAGE_BYTES;
            CK(cudaMemcpyAsync(s.h_stage[slot], (const char *)src + off,
                               std::min(STAGE_BYTES, bytes - off), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(s.stage_ev[slot], st));
        }
    }
    return KEM_OK;
}

// small host->device upload through the pinned ring (time tables, uniform tables)
int small_upload(Shard &s, void *dst, const void *src, size_t bytes)
{
    CK(cudaSetDevice(s.dev));
    if (bytes > SMALL_BYTES) {   // rare: huge n_sub; synchronous pageable copy
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
        return KEM_OK;
    }
    const int slot = s.small_next;
    s.small_next = (slot + 1) % SMALL_RING;
    if (s.small_busy[slot]) CK(cudaEventSynchronize(s.small_ev[slot]));
    memcpy(s.h_small[slot], src, bytes);
    CK(cudaMemcpyAsync(dst, s.h_small[slot], bytes, cudaMemcpyHostToDevice, s.stream));
    CK(cudaEventRecord(s.small_ev[slot], s.stream));
    s.small_busy[slot] = true;
    return KEM_OK;
}

// true if the parameter column's current value is not in a per-DOF device column
bool not_on_device(kem_handle h, int col)
{
    return h->p_uniform[col] || h->p_host[col] || (!h->shards.empty() && !h->shards[0].pcol[col]);
}

// make the parameter column a per-DOF device column holding its current value
int ensure_pcol(kem_handle h, int col)
{
    for (Shard &s : h->shards) {
        CK(cudaSetDevice(s.dev));
        if (!s.pcol[col] && s.n > 0)
            CK(cudaMalloc(&s.pcol[col], (size_t)s.n * sizeof(double)));
        if (h->p_uniform[col] && s.n > 0) {
            k_fill<<<grid_for(s.n), 256, 0, s.stream>>>(s.pcol[col], s.n, h->uni[col]);
            CK(cudaGetLastError());
            h->launches++;
        } else if (h->p_host[col] && s.n > 0) {
            int rc = copy_in(s, s.pcol[col], h->p_shadow[col].data() + s.begin, (size_t)s.n * sizeof(double),
                             s.stream, false);
            if (rc) return rc;
        }
    }
    h->p_uniform[col] = 0;
    if (h->p_host[col]) {
        h->p_host[col] = 0;
        std::vector<double>().swap(h->p_shadow[col]);
    }
    return KEM_OK;
}

// Where kem_step_io puts a PINNED input to a slot the RHS never touches.  The shadow copy
// replaces one DMA read of the column by a host read + write of it.  With one or two GPUs
// on the host the link is the bottleneck and the shadow wins (1 GPU: 10.9 vs 12.9 ms per
// exchange of 1e7 DOFs); with four or more the host memory system is, and the DMA wins
// (4 GPUs, same box: 2.02e9 vs 1.70e9 DOF-steps/s; profiles/r1_bench.md).  Pageable inputs
// always go to the shadow: their alternative is a staging copy plus the DMA.
bool shadow_pinned_inputs(int n_dev)
{
    if (const char *e = getenv("KNPEMI_HOST_SHADOW_IO")) return atoi(e) != 0;
    int sharing = n_dev;                                   // GPUs fed from this host's memory
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) sharing = std::max(sharing, n_dev * atoi(e));
    return sharing <= 2;
}

// full-column write to a parameter slot the RHS never touches: host shadow only
void shadow_store(kem_handle h, int col, const double *src)
{
    h->p_shadow[col].resize((size_t)h->n);
    CopyPool::get().copy(h->p_shadow[col].data(), src, (size_t)h->n * sizeof(double));
    h->p_uniform[col] = 0;
    h->p_host[col] = 1;
}

int check_col(kem_handle h, int kind, int col, const char *fn)
{
    if (!h) return fail(KEM_E_ARG, std::string(fn) + ": null handle");
    if (kind != KEM_STATE && kind != KEM_PARAM) return fail(KEM_E_ARG, std::string(fn) + ": bad kind");
    const int lim = kind == KEM_STATE ? h->m->ns : h->m->np;
    if (col < 0 || col >= lim) return fail(KEM_E_ARG, std::string(fn) + ": column out of range");
    return KEM_OK;
}

double *col_ptr(Shard &s, int kind, int col) { return kind == KEM_STATE ? s.ycol[col] : s.pcol[col]; }

int sync_all(kem_handle h)
{
    for (Shard &s : h->shards) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_in));
        CK(cudaStreamSynchronize(s.stream2));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaStreamSynchronize(s.s_out));
    }
    return KEM_OK;
}

// (2*n_sub+2) stage times, formed exactly as oracle/knpemi_oracle.c:step_row does
void build_ttab(const KemModelDesc *m, double t0, double dt, int n_sub, std::vector<double> &tab)
{
    const int nt = m->n_tslots;
    tab.assign((size_t)(2 * n_sub + 2) * nt, 0.0);
    if (nt == 0) return;
    const double hstep = dt / (double)n_sub;
    for (int j = 0; j < n_sub; ++j) {
        const double ta = t0 + (double)j * hstep;
        const double tb = t0 + ((double)j + 0.5) * hstep;
        m->tonly(ta, &tab[(size_t)(2 * j) * nt]);
        m->tonly(tb, &tab[(size_t)(2 * j + 1) * nt]);
    }
    const double tc = t0 + ((double)(n_sub - 1) + 1.0) * hstep;
    m->tonly(tc, &tab[(size_t)(2 * n_sub) * nt]);
    m->tonly(t0 + dt, &tab[(size_t)(2 * n_sub + 1) * nt]);
}

// DOF chunks of one pipelined exchange (kem_step_io): enough chunks that the pipeline
// fill/drain (one chunk of kernel + D2H) is a few percent of the exchange, chunks large
// enough (>= 128k DOFs) that each launch still fills the GPU for several waves.
int io_chunks(int64_t n, int64_t *chunk_out)
{
    int target = IO_TARGET_CHUNKS;
    if (const char *e = getenv("KNPEMI_IO_CHUNKS")) target = std::max(1, std::min(atoi(e), IO_MAX_CHUNKS));
    int64_t chunk = std::max<int64_t>((n + target - 1) / target, 1 << 17);
    chunk = (chunk + 1023) / 1024 * 1024;
    *chunk_out = chunk;
    return (int)((n + chunk - 1) / chunk);
}

struct StepPlan {
    int scheme = KEM_SCHEME_RK4;
    double t0 = 0.0, dt = 0.0, t_end = 0.0;
    int n_stim = 0;
    int stim_col[KEM_MAX_STIM];
    double stim_val[KEM_MAX_STIM];
    bool masked = false;
    std::vector<double> ttab;
    double hstep = 0.0;
    int n_sub = 0;
};

// validates the step arguments, folds an unmasked stimulus into the uniform
// table, makes masked stimulus columns per-DOF, uploads time/uniform tables
int prepare_step(kem_handle h, double t0, double dt, int n_sub, int scheme, int n_stim,
                 const int *stim_cols, const double *stim_vals, StepPlan &pl)
{
    ARG(h, "null handle");
    ARG(scheme == KEM_SCHEME_RK4 || scheme == KEM_SCHEME_DP45, "unknown scheme");
    ARG(scheme != KEM_SCHEME_RK4 || (n_sub >= 1 && n_sub <= 100000), "n_sub out of range");
    ARG(scheme != KEM_SCHEME_DP45 || dt > 0.0, "KEM_SCHEME_DP45 needs dt > 0");
    ARG(n_stim >= 0 && n_stim <= KEM_MAX_STIM, "too many stimulus entries");
    ARG(n_stim == 0 || (stim_cols && stim_vals), "null stimulus arrays");
    ARG(isfinite(t0) && isfinite(dt), "non-finite time");
    const KemModelDesc *m = h->m;
    pl.scheme = scheme;
    pl.t0 = t0;
    pl.dt = dt;
    pl.t_end = t0 + dt;
    pl.n_sub = scheme == KEM_SCHEME_RK4 ? n_sub : 0;
    pl.hstep = scheme == KEM_SCHEME_RK4 ? dt / (double)n_sub : 0.0;
    if (scheme == KEM_SCHEME_DP45)
        for (Shard &s : h->shards) {      // per-DOF step sizes, counters, activity sort buffers
            CK(cudaSetDevice(s.dev));
            const size_t nn = std::max<size_t>((size_t)s.n, 1);
            if (!s.d_hsug) {
                CK(cudaMalloc(&s.d_hsug, nn * sizeof(double)));
                CK(cudaMemsetAsync(s.d_hsug, 0, nn * sizeof(double), s.stream));
            }
            if (!s.d_stats) {
                CK(cudaMalloc(&s.d_stats, 2 * sizeof(unsigned long long)));
                CK(cudaMemsetAsync(s.d_stats, 0, 2 * sizeof(unsigned long long), s.stream));
            }
            if (!s.h_stats)
                CK(cudaHostAlloc((void **)&s.h_stats, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
            if (!s.d_perm) CK(cudaMalloc(&s.d_perm, nn * sizeof(int)));
            if (!s.d_act) CK(cudaMalloc(&s.d_act, (2 * ACT_BUCKETS + 1) * sizeof(unsigned)));
        }
    pl.masked = !h->shards.empty() && h->shards[0].has_mask;
    for (int s = 0; s < n_stim; ++s) {
        ARG(stim_cols[s] >= 0 && stim_cols[s] < m->np, "stimulus column out of range");
        if (pl.masked) {
            pl.stim_col[pl.n_stim] = stim_cols[s];
            pl.stim_val[pl.n_stim] = stim_vals[s];
            pl.n_stim++;
            if (not_on_device(h, stim_cols[s])) {
                int rc = ensure_pcol(h, stim_cols[s]);
                if (rc) return rc;
            }
        } else {
            // every DOF is stimulated: parameters[:, col] = value (odeSolver.py:110-112)
            int rc = kem_set_uniform(h, KEM_PARAM, stim_cols[s], stim_vals[s]);
            if (rc) return rc;
        }
    }
    if (scheme == KEM_SCHEME_RK4) build_ttab(m, t0, dt, n_sub, pl.ttab);
    for (Shard &s : h->shards) {
        CK(cudaSetDevice(s.dev));
        const size_t tb = pl.ttab.size() * sizeof(double);
        if (tb > s.ttab_cap) {
            CK(cudaStreamSynchronize(s.stream));
            if (s.d_ttab) CK(cudaFree(s.d_ttab));
            CK(cudaMalloc(&s.d_ttab, tb));
            s.ttab_cap = tb;
        }
        if (tb) {
            int rc = small_upload(s, s.d_ttab, pl.ttab.data(), tb);
            if (rc) return rc;
        }
        if (h->uni_dirty) {
            int rc = small_upload(s, s.d_uni, h->uni.data(), h->uni.size() * sizeof(double));
            if (rc) return rc;
        }
    }
    h->uni_dirty = false;
    return KEM_OK;
}

// scheme O3: counting sort of the shard's DOFs by the step size they used last time
int build_activity_perm(kem_handle h, Shard &s, double dt)
{
    if (s.n == 0 || s.n > 0x7fffffffLL) {
        s.perm_valid = false;
        return KEM_OK;
    }
    CK(cudaSetDevice(s.dev));
    CK(cudaMemsetAsync(s.d_act, 0, (2 * ACT_BUCKETS + 1) * sizeof(unsigned), s.stream));
    k_activity_hist<<<grid_for(s.n), 256, 0, s.stream>>>(s.d_hsug, dt, s.n, s.d_act);
    k_activity_scan<<<1, 32, 0, s.stream>>>(s.d_act, s.d_act + ACT_BUCKETS);
    k_activity_scatter<<<grid_for(s.n), 256, 0, s.stream>>>(s.d_hsug, dt, s.n, s.d_act + ACT_BUCKETS,
                                                            s.d_perm);
    CK(cudaGetLastError());
    h->launches += 3;
    s.perm_valid = true;
    return KEM_OK;
}

// enqueue the fused kernel for DOFs [off, off+len) of shard s on its compute stream
int launch_range(kem_handle h, Shard &s, const StepPlan &pl, int64_t off, int64_t len,
                 cudaStream_t on = nullptr)
{
    if (!on) on = s.stream;
    const KemModelDesc *m = h->m;
    std::vector<double *> y(m->ns);
    std::vector<const double *> p(m->np);
    std::vector<int64_t> pm(m->np);
    std::vector<double *> o(std::max(m->n_out, 1));
    for (int c = 0; c < m->ns; ++c) y[c] = s.ycol[c] + off;
    for (int c = 0; c < m->np; ++c) {
        if (h->p_uniform[c] || h->p_host[c]) {      // (a host-shadowed slot is never read by the kernel)
            p[c] = s.d_uni + c;
            pm[c] = 0;
        } else {
            p[c] = s.pcol[c] + off;
            pm[c] = ~(int64_t)0;
        }
    }
    for (int k = 0; k < m->n_out; ++k) o[k] = s.pcol[m->out_cols[k]] + off;
    KemLaunch L;
    memset(&L, 0, sizeof L);
    L.n = len;
    L.y = y.data();
    L.p = p.data();
    L.pmask = pm.data();
    L.out = o.data();
    L.stim_mask = (pl.masked && pl.n_stim > 0) ? s.d_mask + off : nullptr;
    L.n_stim = pl.n_stim;
    for (int k = 0; k < pl.n_stim; ++k) {
        L.stim_col[k] = pl.stim_col[k];
        L.stim_val[k] = pl.stim_val[k];
        L.stim_ptr[k] = s.pcol[pl.stim_col[k]] + off;
    }
    L.ttab = s.d_ttab;
    L.n_sub = pl.n_sub;
    L.h = pl.hstep;
    L.flags = s.d_flags;
    L.block = h->block;
    L.scheme = pl.scheme;
    L.t0 = pl.t0;
    L.dt = pl.dt;
    L.t_end = pl.t_end;
    L.rtol = h->rtol;
    L.atol = h->atol;
    L.hsug = s.d_hsug ? s.d_hsug + off : nullptr;
    L.stats = s.d_stats;
    L.perm = nullptr;
    L.perm_on = nullptr;
    if (pl.scheme == KEM_SCHEME_DP45 && h->activity_sort && off == 0 && len == s.n) {
        int rc = build_activity_perm(h, s, pl.dt);     // whole-range launches only (not the chunks
        if (rc) return rc;                             // of kem_step_io)
        if (s.perm_valid) {
            L.perm = s.d_perm;
            L.perm_on = s.d_act + 2 * ACT_BUCKETS;
        }
    }
    CK(cudaSetDevice(s.dev));
    cudaError_t e = m->launch(&L, on);
    if (e != cudaSuccess)
        return fail(KEM_E_CUDA, std::string("step kernel launch failed: ") + cudaGetErrorString(e));
    h->launches++;
    return KEM_OK;
}

int read_flags(kem_handle h, int *status_flags)
{
    int flags = 0;
    for (Shard &s : h->shards) {
        CK(cudaSetDevice(s.dev));
        CK(cudaMemcpyAsync(s.h_flags, s.d_flags, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    }
    for (Shard &s : h->shards) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.stream));
        flags |= *s.h_flags;
    }
    *status_flags = flags;
    if (flags & 1) {
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            CK(cudaMemsetAsync(s.d_flags, 0, sizeof(int), s.stream));
        }
        g_err = "kem_step: a membrane state became non-finite";
        return KEM_NONFINITE;
    }
    return KEM_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int kem_version(void) { return 100; }

const char *kem_last_error(void) { return g_err.c_str(); }

int kem_device_count(int *n_out)
{
    ARG(n_out, "null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *n_out = 0;
        return fail(KEM_E_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *n_out = n;
    return KEM_OK;
}

// ------------------------------------------------------------------------- models
int kem_model_load(const char *so_path, int *model_id_out)
{
    ARG(so_path && model_id_out, "null argument");
    std::lock_guard<std::mutex> lk(g_models_mu);
    for (size_t i = 0; i < g_models.size(); ++i)
        if (g_models[i].path == so_path) {
            *model_id_out = (int)i;
            return KEM_OK;
        }
    void *dl = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) return fail(KEM_E_MODEL, std::string("dlopen failed: ") + dlerror());
    auto fn = (kem_model_descriptor_fn)dlsym(dl, "kem_model_descriptor");
    if (!fn) {
        dlclose(dl);
        return fail(KEM_E_MODEL, std::string(so_path) + " exports no kem_model_descriptor");
    }
    const KemModelDesc *d = fn();
    if (!d || d->abi_version != KEM_MODEL_ABI_VERSION) {
        dlclose(dl);
        return fail(KEM_E_MODEL, std::string(so_path) + ": model ABI version mismatch (regenerate)");
    }
    if (d->ns < 1 || d->np < 0 || d->n_out < 0 || d->n_out > 64 || !d->launch || !d->tonly) {
        dlclose(dl);
        return fail(KEM_E_MODEL, std::string(so_path) + ": malformed model descriptor");
    }
    g_models.push_back({d, dl, so_path});
    *model_id_out = (int)g_models.size() - 1;
    return KEM_OK;
}

int kem_model_find(const char *key, int *model_id_out)
{
    ARG(key && model_id_out, "null argument");
    std::lock_guard<std::mutex> lk(g_models_mu);
    for (int i = (int)g_models.size() - 1; i >= 0; --i)
        if (!strcmp(g_models[i].desc->name, key) || !strcmp(g_models[i].desc->source_hash, key)) {
            *model_id_out = i;
            return KEM_OK;
        }
    return fail(KEM_E_MODEL, std::string("no loaded model named ") + key);
}

int kem_model_get_info(int model_id, kem_model_info *out)
{
    ARG(out, "null output");
    const KemModelDesc *d = model_desc(model_id);
    if (!d) return fail(KEM_E_MODEL, "unknown model id");
    memset(out, 0, sizeof *out);
    out->ns = d->ns;
    out->np = d->np;
    out->n_out = d->n_out;
    out->n_used = d->n_used;
    out->n_tslots = d->n_tslots;
    for (int k = 0; k < d->n_out; ++k) out->out_cols[k] = d->out_cols[k];
    snprintf(out->name, sizeof out->name, "%s", d->name);
    snprintf(out->source_hash, sizeof out->source_hash, "%s", d->source_hash);
    return KEM_OK;
}

int kem_model_launch_info(int model_id, int dev, int block, int *regs_out, int *blocks_per_sm_out)
{
    ARG(regs_out && blocks_per_sm_out, "null output");
    const KemModelDesc *d = model_desc(model_id);
    if (!d) return fail(KEM_E_MODEL, "unknown model id");
    CK(cudaSetDevice(dev));
    cudaError_t e = d->launch_info(regs_out, blocks_per_sm_out, block);
    if (e != cudaSuccess) return fail(KEM_E_CUDA, std::string("launch_info: ") + cudaGetErrorString(e));
    return KEM_OK;
}

// ------------------------------------------------------------------- construction
int kem_create(int model_id, int64_t n_dof, int n_dev, const int *dev_ids,
               const double *state_defaults, const double *param_defaults, kem_handle *out)
{
    NvtxRange nvtx_range("kem_create");
    ARG(out, "null output");
    *out = nullptr;
    const KemModelDesc *m = model_desc(model_id);
    if (!m) return fail(KEM_E_MODEL, "unknown model id");
    ARG(n_dof >= 0, "negative n_dof");
    ARG(n_dev >= 1 && n_dev <= 64, "n_dev out of range");
    ARG(state_defaults && (param_defaults || m->np == 0), "null defaults");
    int avail = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&avail);
        if (e != cudaSuccess || avail < 1) {
            cudaGetLastError();
            return fail(KEM_E_CUDA,
                        "no CUDA device: libknpemi_b200 has no CPU fallback (cudaGetDeviceCount: " +
                            std::string(cudaGetErrorString(e)) + ")");
        }
    }
    kem_handle h = new kem_handle_s;
    h->m = m;
    h->model_id = model_id;
    h->n = n_dof;
    h->uni.assign(param_defaults, param_defaults + m->np);
    h->p_uniform.assign(m->np, 1);
    h->p_host.assign(m->np, 0);
    h->p_shadow.resize(m->np);
    h->p_dead.assign(m->np, 1);
    for (int k = 0; k < m->n_used; ++k) h->p_dead[m->used_cols[k]] = 0;
    for (int k = 0; k < m->n_out; ++k) h->p_dead[m->out_cols[k]] = 0;
    if (getenv("KNPEMI_NO_HOST_SHADOW")) h->p_dead.assign(m->np, 0);
    h->shadow_pinned_io = shadow_pinned_inputs(n_dev);
    h->shards.resize(n_dev);
    const int64_t per = (n_dof + n_dev - 1) / n_dev;   // contiguous ranges, remainder on the last
    auto bail = [&](int rc) {
        std::string keep = g_err;
        kem_destroy(h);
        g_err = keep;
        return rc;
    };
#define CKB(call)                                                                          \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            fail(KEM_E_CUDA, std::string(#call) + " -> " + cudaGetErrorString(e__));       \
            return bail(e__ == cudaErrorMemoryAllocation ? KEM_E_NOMEM : KEM_E_CUDA);      \
        }                                                                                  \
    } while (0)
    for (int k = 0; k < n_dev; ++k) {
        Shard &s = h->shards[k];
        s.dev = dev_ids ? dev_ids[k] : k;
        if (s.dev < 0 || s.dev >= avail) {
            fail(KEM_E_ARG, "kem_create: device id out of range");
            return bail(KEM_E_ARG);
        }
        s.begin = std::min<int64_t>((int64_t)k * per, n_dof);
        s.n = std::min<int64_t>(s.begin + per, n_dof) - s.begin;
        s.ycol.assign(m->ns, nullptr);
        s.pcol.assign(m->np, nullptr);
        CKB(cudaSetDevice(s.dev));
        CKB(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CKB(cudaStreamCreateWithFlags(&s.s_in, cudaStreamNonBlocking));
        CKB(cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking));
        CKB(cudaStreamCreateWithFlags(&s.s_out, cudaStreamNonBlocking));
        CKB(cudaEventCreate(&s.ev_a));
        CKB(cudaEventCreate(&s.ev_b));
        CKB(cudaEventCreate(&s.ev_c));
        CKB(cudaEventCreate(&s.ev_d));
        CKB(cudaEventCreate(&s.ev_t0));
        CKB(cudaEventCreate(&s.ev_t1));
        for (int r = 0; r < SMALL_RING; ++r) {
            CKB(cudaHostAlloc(&s.h_small[r], SMALL_BYTES, cudaHostAllocDefault));
            CKB(cudaEventCreateWithFlags(&s.small_ev[r], cudaEventDisableTiming));
        }
        CKB(cudaHostAlloc((void **)&s.h_flags, sizeof(int), cudaHostAllocDefault));
        *s.h_flags = 0;
        CKB(cudaMalloc(&s.d_flags, sizeof(int)));
        CKB(cudaMemsetAsync(s.d_flags, 0, sizeof(int), s.stream));
        CKB(cudaMalloc(&s.d_uni, std::max(m->np, 1) * sizeof(double)));
        if (s.n > 0) {
            CKB(cudaMalloc(&s.d_mask, (size_t)s.n));
            for (int c = 0; c < m->ns; ++c) {
                CKB(cudaMalloc(&s.ycol[c], (size_t)s.n * sizeof(double)));
                k_fill<<<grid_for(s.n), 256, 0, s.stream>>>(s.ycol[c], s.n, state_defaults[c]);
                CKB(cudaGetLastError());
                h->launches++;
            }
        }
    }
#undef CKB
    // output slots are always per-DOF columns
    for (int k = 0; k < m->n_out; ++k) {
        int rc = ensure_pcol(h, m->out_cols[k]);
        if (rc) return bail(rc);
    }
    int rc = sync_all(h);
    if (rc) return bail(rc);
    *out = h;
    return KEM_OK;
}

int kem_destroy(kem_handle h)
{
    if (!h) return KEM_OK;
    for (Shard &s : h->shards) {
        if (cudaSetDevice(s.dev) != cudaSuccess) continue;
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.s_in) cudaStreamSynchronize(s.s_in);
        if (s.stream2) cudaStreamSynchronize(s.stream2);
        if (s.s_out) cudaStreamSynchronize(s.s_out);
        for (double *p : s.ycol) if (p) cudaFree(p);
        for (double *p : s.pcol) if (p) cudaFree(p);
        if (s.d_uni) cudaFree(s.d_uni);
        if (s.d_mask) cudaFree(s.d_mask);
        if (s.d_ttab) cudaFree(s.d_ttab);
        if (s.d_flags) cudaFree(s.d_flags);
        if (s.d_hsug) cudaFree(s.d_hsug);
        if (s.d_perm) cudaFree(s.d_perm);
        if (s.d_act) cudaFree(s.d_act);
        if (s.d_stats) cudaFree(s.d_stats);
        if (s.h_stats) cudaFreeHost(s.h_stats);
        for (long long *m : s.d_map) if (m) cudaFree(m);
        if (s.h_flags) cudaFreeHost(s.h_flags);
        for (int r = 0; r < SMALL_RING; ++r) {
            if (s.h_small[r]) cudaFreeHost(s.h_small[r]);
            if (s.small_ev[r]) cudaEventDestroy(s.small_ev[r]);
        }
        for (int r = 0; r < N_STAGE; ++r) {
            if (s.h_stage[r]) cudaFreeHost(s.h_stage[r]);
            if (s.stage_ev[r]) cudaEventDestroy(s.stage_ev[r]);
        }
        for (auto *v : {&s.io_in, &s.io_k0, &s.io_k1, &s.io_out})
            for (cudaEvent_t e : *v) cudaEventDestroy(e);
        for (cudaEvent_t e : {s.ev_a, s.ev_b, s.ev_c, s.ev_d, s.ev_t0, s.ev_t1}) if (e) cudaEventDestroy(e);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.s_in) cudaStreamDestroy(s.s_in);
        if (s.stream2) cudaStreamDestroy(s.stream2);
        if (s.s_out) cudaStreamDestroy(s.s_out);
    }
    cudaGetLastError();
    delete h;
    return KEM_OK;
}

int kem_n_dof(kem_handle h, int64_t *n_out)
{
    ARG(h && n_out, "null argument");
    *n_out = h->n;
    return KEM_OK;
}

int kem_shard_range(kem_handle h, int k, int *dev_out, int64_t *begin_out, int64_t *end_out)
{
    ARG(h && dev_out && begin_out && end_out, "null argument");
    ARG(k >= 0 && k < (int)h->shards.size(), "shard index out of range");
    *dev_out = h->shards[k].dev;
    *begin_out = h->shards[k].begin;
    *end_out = h->shards[k].begin + h->shards[k].n;
    return KEM_OK;
}

// ------------------------------------------------------------------- table access
int kem_set_uniform(kem_handle h, int kind, int col, double v)
{
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    bool is_out = false;
    if (kind == KEM_PARAM)
        for (int k = 0; k < h->m->n_out; ++k) is_out |= (h->m->out_cols[k] == col);
    if (kind == KEM_PARAM && !is_out) {
        // the per-DOF allocation (if any) stays cached for a later kem_set_column
        if (h->p_uniform[col] && memcmp(&h->uni[col], &v, sizeof v) == 0) return KEM_OK;
        h->uni[col] = v;
        h->p_uniform[col] = 1;
        h->p_host[col] = 0;
        h->uni_dirty = true;
        return KEM_OK;
    }
    if (kind == KEM_PARAM) h->uni[col] = v;
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        CK(cudaSetDevice(s.dev));
        k_fill<<<grid_for(s.n), 256, 0, s.stream>>>(col_ptr(s, kind, col), s.n, v);
        CK(cudaGetLastError());
        h->launches++;
    }
    return KEM_OK;
}

int kem_set_column(kem_handle h, int kind, int col, const double *src, int64_t n)
{
    NvtxRange nvtx_range("kem_set_column");
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    ARG(n == h->n, "length must equal the handle's n_dof");
    ARG(src || n == 0, "null source");
    if (kind == KEM_PARAM && h->p_dead[col] && n > 0) {
        shadow_store(h, col, src);
        return KEM_OK;
    }
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        // becomes a per-DOF column; no need to pre-fill, every row is overwritten
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            if (!s.pcol[col] && s.n > 0) CK(cudaMalloc(&s.pcol[col], (size_t)s.n * sizeof(double)));
        }
        h->p_uniform[col] = 0;
        h->p_host[col] = 0;
    }
    const bool pinned = n > 0 && is_pinned(src);
    for (Shard &s : h->shards) {
        rc = copy_in(s, col_ptr(s, kind, col), src + s.begin, (size_t)s.n * sizeof(double), s.stream,
                     pinned);
        if (rc) return rc;
    }
    if (pinned)   // the caller may overwrite `src` as soon as we return
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            CK(cudaStreamSynchronize(s.stream));
        }
    return KEM_OK;
}

// scratch device buffer of a setter call, released on every exit path
struct ScratchBuf {
    void *p = nullptr;
    ~ScratchBuf()
    {
        if (p) cudaFree(p);
    }
};

static int upload_mask_tmp(Shard &s, const uint8_t *host_mask, ScratchBuf &buf)
{
    CK(cudaSetDevice(s.dev));
    CK(cudaMalloc(&buf.p, (size_t)s.n));
    CK(cudaMemcpyAsync(buf.p, host_mask + s.begin, (size_t)s.n, cudaMemcpyHostToDevice, s.stream));
    return KEM_OK;
}

int kem_set_column_masked(kem_handle h, int kind, int col, const double *src,
                          const uint8_t *host_mask, int64_t n)
{
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    ARG(n == h->n, "length must equal the handle's n_dof");
    ARG((src && host_mask) || n == 0, "null source or mask");
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        rc = ensure_pcol(h, col);
        if (rc) return rc;
    }
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        ScratchBuf d_m, d_src;
        rc = upload_mask_tmp(s, host_mask, d_m);
        if (rc) return rc;
        CK(cudaMalloc(&d_src.p, (size_t)s.n * sizeof(double)));
        rc = copy_in(s, (double *)d_src.p, src + s.begin, (size_t)s.n * sizeof(double), s.stream, false);
        if (rc) return rc;
        k_copy_masked<<<grid_for(s.n), 256, 0, s.stream>>>(col_ptr(s, kind, col), (const double *)d_src.p,
                                                           (const unsigned char *)d_m.p, s.n);
        CK(cudaGetLastError());
        h->launches++;
        CK(cudaStreamSynchronize(s.stream));
    }
    return KEM_OK;
}

int kem_set_value_masked(kem_handle h, int kind, int col, double v, const uint8_t *host_mask,
                         int64_t n)
{
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    ARG(n == h->n, "length must equal the handle's n_dof");
    ARG(host_mask || n == 0, "null mask");
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        rc = ensure_pcol(h, col);
        if (rc) return rc;
    }
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        ScratchBuf d_m;
        rc = upload_mask_tmp(s, host_mask, d_m);
        if (rc) return rc;
        k_set_value_masked<<<grid_for(s.n), 256, 0, s.stream>>>(col_ptr(s, kind, col),
                                                                (const unsigned char *)d_m.p, s.n, v);
        CK(cudaGetLastError());
        h->launches++;
        CK(cudaStreamSynchronize(s.stream));
    }
    return KEM_OK;
}

int kem_get_column(kem_handle h, int kind, int col, double *dst, int64_t n)
{
    NvtxRange nvtx_range("kem_get_column");
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    ARG(n == h->n, "length must equal the handle's n_dof");
    ARG(dst || n == 0, "null destination");
    if (kind == KEM_PARAM && h->p_uniform[col]) {
        std::fill(dst, dst + n, h->uni[col]);
        return KEM_OK;
    }
    if (kind == KEM_PARAM && h->p_host[col]) {
        CopyPool::get().copy(dst, h->p_shadow[col].data(), (size_t)n * sizeof(double));
        return KEM_OK;
    }
    const bool pinned = n > 0 && is_pinned(dst);
    for (Shard &s : h->shards) {
        rc = copy_out(s, dst + s.begin, col_ptr(s, kind, col), (size_t)s.n * sizeof(double), s.stream,
                      pinned);
        if (rc) return rc;
    }
    if (pinned)
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            CK(cudaStreamSynchronize(s.stream));
        }
    return KEM_OK;
}

int kem_column_is_uniform(kem_handle h, int kind, int col, int *is_uniform_out, double *value_out)
{
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    ARG(is_uniform_out, "null output");
    const bool u = kind == KEM_PARAM && h->p_uniform[col];
    *is_uniform_out = u ? 1 : 0;
    if (value_out) *value_out = u ? h->uni[col] : 0.0;
    return KEM_OK;
}

int kem_column_location(kem_handle h, int kind, int col, int *location_out)
{
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    ARG(location_out, "null output");
    if (kind == KEM_STATE) *location_out = 1;
    else *location_out = h->p_uniform[col] ? 0 : (h->p_host[col] ? 2 : 1);
    return KEM_OK;
}

int kem_set_stimulus_mask(kem_handle h, const uint8_t *host_mask, int64_t n)
{
    ARG(h, "null handle");
    if (!host_mask) {
        for (Shard &s : h->shards) s.has_mask = false;
        return KEM_OK;
    }
    ARG(n == h->n, "length must equal the handle's n_dof");
    for (Shard &s : h->shards) {
        s.has_mask = true;
        if (s.n == 0) continue;
        CK(cudaSetDevice(s.dev));
        CK(cudaMemcpyAsync(s.d_mask, host_mask + s.begin, (size_t)s.n, cudaMemcpyHostToDevice,
                           s.stream));
        CK(cudaStreamSynchronize(s.stream));   // pageable source: consumed on return
    }
    return KEM_OK;
}

// -------------------------------------------------------------------------- step
int kem_step_timed(kem_handle h, double t0, double dt, int n_sub, int scheme, int n_stim,
                   const int *stim_cols, const double *stim_vals, int *status_flags,
                   kem_step_times *times)
{
    NvtxRange nvtx_range("kem_step_timed");
    StepPlan pl;
    int rc = prepare_step(h, t0, dt, n_sub, scheme, n_stim, stim_cols, stim_vals, pl);
    if (rc) return rc;
    for (Shard &s : h->shards) {
        if (times) {
            CK(cudaSetDevice(s.dev));
            CK(cudaEventRecord(s.ev_a, s.stream));
        }
        rc = launch_range(h, s, pl, 0, s.n);
        if (rc) return rc;
        if (times) CK(cudaEventRecord(s.ev_b, s.stream));
    }
    if (times) {
        memset(times, 0, sizeof *times);
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            CK(cudaEventSynchronize(s.ev_b));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, s.ev_a, s.ev_b));
            times->ms_kernel = std::max(times->ms_kernel, (double)ms);
        }
        times->ms_total = times->ms_kernel;
    }
    if (status_flags) return read_flags(h, status_flags);
    return KEM_OK;
}

int kem_step(kem_handle h, double t0, double dt, int n_sub, int scheme, int n_stim,
             const int *stim_cols, const double *stim_vals, int *status_flags)
{
    return kem_step_timed(h, t0, dt, n_sub, scheme, n_stim, stim_cols, stim_vals, status_flags,
                          nullptr);
}

int kem_step_io(kem_handle h, double t0, double dt, int n_sub, int scheme, int n_stim,
                const int *stim_cols, const double *stim_vals, int n_in, const kem_io_column *in,
                int n_out, const kem_io_column *out, int *status_flags, kem_step_times *times)
{
    NvtxRange nvtx_range("kem_step_io");
    ARG(h, "null handle");
    ARG(n_in >= 0 && n_out >= 0 && (in || !n_in) && (out || !n_out), "bad io arrays");
    int rc;
    bool all_pinned = true;
    // columns that really cross the host link; inputs to slots the RHS never reads stay in
    // their host shadow, outputs that live in a host shadow are copied from it
    std::vector<kem_io_column> dev_in, dev_out, shadow_in, shadow_out;
    for (int k = 0; k < n_in; ++k) {
        rc = check_col(h, in[k].kind, in[k].col, __func__);
        if (rc) return rc;
        ARG(in[k].host || h->n == 0, "null input column");
        if (in[k].kind == KEM_PARAM && h->p_dead[in[k].col] && h->n > 0 &&
            (h->shadow_pinned_io || !is_pinned(in[k].host))) {
            shadow_in.push_back(in[k]);
            continue;
        }
        dev_in.push_back(in[k]);
        all_pinned = all_pinned && (h->n == 0 || is_pinned(in[k].host));
        if (in[k].kind == KEM_PARAM && not_on_device(h, in[k].col)) {
            for (Shard &s : h->shards) {
                CK(cudaSetDevice(s.dev));
                if (!s.pcol[in[k].col] && s.n > 0)
                    CK(cudaMalloc(&s.pcol[in[k].col], (size_t)s.n * sizeof(double)));
            }
            h->p_uniform[in[k].col] = 0;
            h->p_host[in[k].col] = 0;
            std::vector<double>().swap(h->p_shadow[in[k].col]);
        }
    }
    for (int k = 0; k < n_out; ++k) {
        rc = check_col(h, out[k].kind, out[k].col, __func__);
        if (rc) return rc;
        ARG(out[k].host || h->n == 0, "null output column");
        ARG(!(out[k].kind == KEM_PARAM && h->p_uniform[out[k].col]),
            "output column is uniform; read it with kem_get_column");
        bool from_shadow = out[k].kind == KEM_PARAM && h->p_host[out[k].col];
        for (const kem_io_column &c : shadow_in)
            from_shadow = from_shadow || (out[k].kind == KEM_PARAM && c.col == out[k].col);
        if (from_shadow) {
            shadow_out.push_back(out[k]);
            continue;
        }
        dev_out.push_back(out[k]);
        all_pinned = all_pinned && (h->n == 0 || is_pinned(out[k].host));
    }
    in = dev_in.data();
    n_in = (int)dev_in.size();
    out = dev_out.data();
    n_out = (int)dev_out.size();
    // host-side part of the exchange; runs while the devices work (the calls below only enqueue)
    auto host_side = [&]() {
        for (const kem_io_column &c : shadow_in) shadow_store(h, c.col, c.host);
        for (const kem_io_column &c : shadow_out)
            CopyPool::get().copy(c.host, h->p_shadow[c.col].data(), (size_t)h->n * sizeof(double));
    };
    StepPlan pl;
    rc = prepare_step(h, t0, dt, n_sub, scheme, n_stim, stim_cols, stim_vals, pl);
    if (rc) return rc;

    if (!all_pinned) {
        // pageable host buffers: staged column copies around one kernel launch
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            CK(cudaEventRecord(s.ev_a, s.stream));
            for (int k = 0; k < n_in; ++k) {
                rc = copy_in(s, col_ptr(s, in[k].kind, in[k].col), in[k].host + s.begin,
                             (size_t)s.n * sizeof(double), s.stream, false);
                if (rc) return rc;
            }
            CK(cudaEventRecord(s.ev_b, s.stream));
            rc = launch_range(h, s, pl, 0, s.n);
            if (rc) return rc;
            CK(cudaEventRecord(s.ev_c, s.stream));
        }
        host_side();
        for (Shard &s : h->shards) {
            for (int k = 0; k < n_out; ++k) {
                rc = copy_out(s, out[k].host + s.begin, col_ptr(s, out[k].kind, out[k].col),
                              (size_t)s.n * sizeof(double), s.stream, false);
                if (rc) return rc;
            }
            CK(cudaSetDevice(s.dev));
            CK(cudaEventRecord(s.ev_d, s.stream));
        }
        if (times) memset(times, 0, sizeof *times);
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            CK(cudaEventSynchronize(s.ev_d));
            if (times) {
                float a = 0, b = 0, c = 0, d = 0;
                CK(cudaEventElapsedTime(&a, s.ev_a, s.ev_b));
                CK(cudaEventElapsedTime(&b, s.ev_b, s.ev_c));
                CK(cudaEventElapsedTime(&c, s.ev_c, s.ev_d));
                CK(cudaEventElapsedTime(&d, s.ev_a, s.ev_d));
                times->ms_h2d = std::max(times->ms_h2d, (double)a);
                times->ms_kernel = std::max(times->ms_kernel, (double)b);
                times->ms_d2h = std::max(times->ms_d2h, (double)c);
                times->ms_total = std::max(times->ms_total, (double)d);
            }
        }
        if (status_flags) return read_flags(h, status_flags);
        return KEM_OK;
    }

    // pinned host buffers: DOF-chunked pipeline, H2D (s_in) | kernel (stream) | D2H (s_out)
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        CK(cudaSetDevice(s.dev));
        int64_t chunk = 0;
        const int n_chunks = io_chunks(s.n, &chunk);
        while ((int)s.io_in.size() < n_chunks) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e)); s.io_in.push_back(e);
            CK(cudaEventCreate(&e)); s.io_k0.push_back(e);
            CK(cudaEventCreate(&e)); s.io_k1.push_back(e);
            CK(cudaEventCreate(&e)); s.io_out.push_back(e);
        }
        // the copy streams must see the table/uniform uploads and earlier work on `stream`
        CK(cudaEventRecord(s.ev_a, s.stream));
        CK(cudaStreamWaitEvent(s.s_in, s.ev_a, 0));
        CK(cudaStreamWaitEvent(s.stream2, s.ev_a, 0));
        CK(cudaEventRecord(s.ev_b, s.s_in));   // t = 0 of this shard's exchange
        const bool two = getenv("KNPEMI_IO_ONE_COMPUTE_STREAM") == nullptr;
        for (int c = 0; c < n_chunks; ++c) {
            const int64_t off = (int64_t)c * chunk;
            const int64_t len = std::min(chunk, s.n - off);
            for (int k = 0; k < n_in; ++k)
                CK(cudaMemcpyAsync(col_ptr(s, in[k].kind, in[k].col) + off, in[k].host + s.begin + off,
                                   (size_t)len * sizeof(double), cudaMemcpyHostToDevice, s.s_in));
            CK(cudaEventRecord(s.io_in[c], s.s_in));
            cudaStream_t sk = (two && (c & 1)) ? s.stream2 : s.stream;
            CK(cudaStreamWaitEvent(sk, s.io_in[c], 0));
            CK(cudaEventRecord(s.io_k0[c], sk));
            rc = launch_range(h, s, pl, off, len, sk);
            if (rc) return rc;
            CK(cudaEventRecord(s.io_k1[c], sk));
            CK(cudaStreamWaitEvent(s.s_out, s.io_k1[c], 0));
            for (int k = 0; k < n_out; ++k)
                CK(cudaMemcpyAsync(out[k].host + s.begin + off, col_ptr(s, out[k].kind, out[k].col) + off,
                                   (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, s.s_out));
            CK(cudaEventRecord(s.io_out[c], s.s_out));
        }
        // later work on `stream` (the next step) must not overtake the D2H copies
        CK(cudaStreamWaitEvent(s.stream, s.io_out[n_chunks - 1], 0));
    }
    host_side();
    if (times) memset(times, 0, sizeof *times);
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_out));
        if (times) {
            int64_t chunk = 0;
            const int n_chunks = io_chunks(s.n, &chunk);
            float tot = 0, kern = 0, h2d = 0, d2h = 0, f = 0;
            CK(cudaEventElapsedTime(&tot, s.ev_b, s.io_out[n_chunks - 1]));
            CK(cudaEventElapsedTime(&h2d, s.ev_b, s.io_in[n_chunks - 1]));
            // chunk kernels overlap across the two compute streams: report the span from the
            // first kernel's start to the last kernel's end, not the sum
            for (int c = std::max(0, n_chunks - 2); c < n_chunks; ++c) {
                CK(cudaEventElapsedTime(&f, s.io_k0[0], s.io_k1[c]));
                kern = std::max(kern, f);
            }
            CK(cudaEventElapsedTime(&d2h, s.io_k1[0], s.io_out[n_chunks - 1]));
            times->ms_total = std::max(times->ms_total, (double)tot);
            times->ms_kernel = std::max(times->ms_kernel, (double)kern);
            times->ms_h2d = std::max(times->ms_h2d, (double)h2d);
            times->ms_d2h = std::max(times->ms_d2h, (double)d2h);
        }
    }
    if (status_flags) return read_flags(h, status_flags);
    return KEM_OK;
}

int kem_sync(kem_handle h)
{
    ARG(h, "null handle");
    int rc = sync_all(h);
    if (rc) return rc;
    int flags = 0;
    return read_flags(h, &flags);
}

int kem_set_activity_sort(kem_handle h, int enabled)
{
    ARG(h, "null handle");
    h->activity_sort = enabled != 0;
    return KEM_OK;
}

int kem_set_tolerances(kem_handle h, double rtol, double atol)
{
    ARG(h, "null handle");
    ARG(rtol > 0.0 && atol >= 0.0 && isfinite(rtol) && isfinite(atol), "tolerances must be positive");
    h->rtol = rtol;
    h->atol = atol;
    return KEM_OK;
}

int kem_get_step_stats(kem_handle h, uint64_t *accepted_out, uint64_t *rejected_out)
{
    ARG(h && accepted_out && rejected_out, "null argument");
    *accepted_out = *rejected_out = 0;
    for (Shard &s : h->shards) {
        if (!s.d_stats) continue;
        CK(cudaSetDevice(s.dev));
        CK(cudaMemcpyAsync(s.h_stats, s.d_stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           s.stream));
        CK(cudaMemsetAsync(s.d_stats, 0, 2 * sizeof(unsigned long long), s.stream));
        CK(cudaStreamSynchronize(s.stream));
        *accepted_out += s.h_stats[0];
        *rejected_out += s.h_stats[1];
    }
    return KEM_OK;
}

int kem_timer_begin(kem_handle h)
{
    ARG(h, "null handle");
    for (Shard &s : h->shards) {
        CK(cudaSetDevice(s.dev));
        CK(cudaEventRecord(s.ev_t0, s.stream));
    }
    return KEM_OK;
}

int kem_timer_end(kem_handle h, double *ms_out)
{
    ARG(h && ms_out, "null argument");
    for (Shard &s : h->shards) {
        CK(cudaSetDevice(s.dev));
        CK(cudaEventRecord(s.ev_t1, s.stream));
    }
    double worst = 0.0;
    for (Shard &s : h->shards) {
        CK(cudaSetDevice(s.dev));
        CK(cudaEventSynchronize(s.ev_t1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s.ev_t0, s.ev_t1));
        worst = std::max(worst, (double)ms);
    }
    *ms_out = worst;
    return KEM_OK;
}

int kem_set_block(kem_handle h, int block)
{
    ARG(h, "null handle");
    ARG(block == 0 || block == 64 || block == 128 || block == 256, "block must be 0, 64, 128 or 256");
    h->block = block;
    return KEM_OK;
}

int kem_launch_count(kem_handle h, int64_t *n_out)
{
    ARG(h && n_out, "null argument");
    *n_out = h->launches;
    return KEM_OK;
}

// ------------------------------------------------- device-resident exchange (f1, f3)
int kem_device_map_set(kem_handle h, int map_id, const int64_t *host_map, int64_t n)
{
    ARG(h, "null handle");
    ARG(map_id >= 0 && map_id < KEM_MAX_MAPS, "map_id out of range");
    ARG(n == h->n, "length must equal the handle's n_dof");
    ARG(host_map || n == 0, "null map");
    for (int64_t i = 0; i < n; ++i) ARG(host_map[i] >= 0, "negative bulk index in map");
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        CK(cudaSetDevice(s.dev));
        if (!s.d_map[map_id]) CK(cudaMalloc(&s.d_map[map_id], (size_t)s.n * sizeof(long long)));
        CK(cudaMemcpyAsync(s.d_map[map_id], host_map + s.begin, (size_t)s.n * sizeof(long long),
                           cudaMemcpyHostToDevice, s.stream));
        CK(cudaStreamSynchronize(s.stream));
    }
    return KEM_OK;
}

static int device_xfer_check(kem_handle h, int shard, int kind, int col, const void *p, int map_id,
                             const char *fn)
{
    int rc = check_col(h, kind, col, fn);
    if (rc) return rc;
    if (shard < 0 || shard >= (int)h->shards.size()) return fail(KEM_E_ARG, std::string(fn) + ": shard out of range");
    if (map_id < 0 || map_id >= KEM_MAX_MAPS || (!h->shards[shard].d_map[map_id] && h->shards[shard].n > 0))
        return fail(KEM_E_ARG, std::string(fn) + ": map not registered (kem_device_map_set)");
    if (!p && h->shards[shard].n > 0) return fail(KEM_E_ARG, std::string(fn) + ": null device pointer");
    return KEM_OK;
}

int kem_device_gather(kem_handle h, int shard, int kind, int col, const double *dev_src, int map_id)
{
    int rc = device_xfer_check(h, shard, kind, col, dev_src, map_id, __func__);
    if (rc) return rc;
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        rc = ensure_pcol(h, col);
        if (rc) return rc;
    }
    Shard &s = h->shards[shard];
    if (s.n == 0) return KEM_OK;
    CK(cudaSetDevice(s.dev));
    k_gather<<<grid_for(s.n), 256, 0, s.stream>>>(col_ptr(s, kind, col), dev_src, s.d_map[map_id], s.n);
    CK(cudaGetLastError());
    h->launches++;
    return KEM_OK;
}

int kem_device_scatter(kem_handle h, int shard, int kind, int col, double *dev_dst, int map_id)
{
    int rc = device_xfer_check(h, shard, kind, col, dev_dst, map_id, __func__);
    if (rc) return rc;
    Shard &s = h->shards[shard];
    if (s.n == 0) return KEM_OK;
    CK(cudaSetDevice(s.dev));
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        rc = ensure_pcol(h, col);       // materialise a uniform / host-shadowed value as a column first
        if (rc) return rc;
    }
    k_scatter<<<grid_for(s.n), 256, 0, s.stream>>>(dev_dst, col_ptr(s, kind, col), s.d_map[map_id], s.n);
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaStreamSynchronize(s.stream));     // the caller's own stream may read dev_dst next
    return KEM_OK;
}

int kem_device_gather_diff(kem_handle h, int shard, int kind, int col, const double *dev_a, int map_a,
                           const double *dev_b, int map_b)
{
    int rc = device_xfer_check(h, shard, kind, col, dev_a, map_a, __func__);
    if (rc) return rc;
    rc = device_xfer_check(h, shard, kind, col, dev_b, map_b, __func__);
    if (rc) return rc;
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        rc = ensure_pcol(h, col);
        if (rc) return rc;
    }
    Shard &s = h->shards[shard];
    if (s.n == 0) return KEM_OK;
    CK(cudaSetDevice(s.dev));
    k_gather_diff<<<grid_for(s.n), 256, 0, s.stream>>>(col_ptr(s, kind, col), dev_a, s.d_map[map_a],
                                                       dev_b, s.d_map[map_b], s.n);
    CK(cudaGetLastError());
    h->launches++;
    return KEM_OK;
}

int kem_device_copy_in(kem_handle h, int shard, int kind, int col, const double *dev_src)
{
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    ARG(shard >= 0 && shard < (int)h->shards.size(), "shard out of range");
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        rc = ensure_pcol(h, col);
        if (rc) return rc;
    }
    Shard &s = h->shards[shard];
    if (s.n == 0) return KEM_OK;
    ARG(dev_src, "null device pointer");
    CK(cudaSetDevice(s.dev));
    CK(cudaMemcpyAsync(col_ptr(s, kind, col), dev_src, (size_t)s.n * sizeof(double), cudaMemcpyDeviceToDevice,
                       s.stream));
    return KEM_OK;
}

int kem_device_copy_out(kem_handle h, int shard, int kind, int col, double *dev_dst)
{
    int rc = check_col(h, kind, col, __func__);
    if (rc) return rc;
    ARG(shard >= 0 && shard < (int)h->shards.size(), "shard out of range");
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        rc = ensure_pcol(h, col);
        if (rc) return rc;
    }
    Shard &s = h->shards[shard];
    if (s.n == 0) return KEM_OK;
    ARG(dev_dst, "null device pointer");
    CK(cudaSetDevice(s.dev));
    CK(cudaMemcpyAsync(dev_dst, col_ptr(s, kind, col), (size_t)s.n * sizeof(double), cudaMemcpyDeviceToDevice,
                       s.stream));
    CK(cudaStreamSynchronize(s.stream));     // the caller's own stream may read dev_dst next
    return KEM_OK;
}

int kem_device_alloc(int dev, size_t bytes, void **ptr_out)
{
    ARG(ptr_out, "null output");
    *ptr_out = nullptr;
    CK(cudaSetDevice(dev));
    CK(cudaMalloc(ptr_out, std::max<size_t>(bytes, 8)));
    return KEM_OK;
}

int kem_device_free(int dev, void *ptr)
{
    CK(cudaSetDevice(dev));
    if (ptr) CK(cudaFree(ptr));
    return KEM_OK;
}

int kem_device_upload(int dev, void *dev_dst, const void *host_src, size_t bytes)
{
    ARG((dev_dst && host_src) || bytes == 0, "null pointer");
    CK(cudaSetDevice(dev));
    CK(cudaMemcpy(dev_dst, host_src, bytes, cudaMemcpyHostToDevice));
    return KEM_OK;
}

int kem_device_download(int dev, void *host_dst, const void *dev_src, size_t bytes)
{
    ARG((host_dst && dev_src) || bytes == 0, "null pointer");
    CK(cudaSetDevice(dev));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
    return KEM_OK;
}

// ------------------------------------------------------------------ pinned memory
int kem_host_alloc(void **ptr_out, size_t bytes)
{
    ARG(ptr_out, "null output");
    *ptr_out = nullptr;
    CK(cudaHostAlloc(ptr_out, std::max<size_t>(bytes, 8), cudaHostAllocPortable));
    return KEM_OK;
}

int kem_host_free(void *ptr)
{
    if (ptr) CK(cudaFreeHost(ptr));
    return KEM_OK;
}

// Page-lock memory the caller owns (the array behind a dolfinx Function): afterwards the
// setters, getters and kem_step_io see it as pinned and DMA directly instead of staging.
int kem_host_register(void *ptr, size_t bytes)
{
    ARG(ptr && bytes > 0, "null pointer or zero size");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return KEM_OK;
    }
    CK(e);
    return KEM_OK;
}

int kem_host_unregister(void *ptr)
{
    ARG(ptr, "null pointer");
    cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return KEM_OK;
    }
    CK(e);
    return KEM_OK;
}

int kem_host_is_pinned(const void *ptr, int *pinned_out)
{
    ARG(ptr && pinned_out, "null argument");
    *pinned_out = is_pinned(ptr) ? 1 : 0;
    return KEM_OK;
}

// ------------------------------------------------------------------- measurement
int kem_fp64_peak(int dev, double *tflops_out, double *ms_out)
{
    ARG(tflops_out, "null output");
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    double *d_out = nullptr;
    CK(cudaMalloc(&d_out, sizeof(double)));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int block = 256, grid = prop.multiProcessorCount * 8;
    const int reps = 8;
    for (int w = 0; w < 3; ++w) k_dfma_peak<<<grid, block, 0, st>>>(d_out, 0.999999, 1e-6);
    CK(cudaGetLastError());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, st));
        for (int k = 0; k < reps; ++k) k_dfma_peak<<<grid, block, 0, st>>>(d_out, 0.999999, 1e-6);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, (double)ms / reps);
    }
    const double flops = 2.0 * (double)grid * block * PEAK_CHAINS * PEAK_ITERS;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
    CK(cudaStreamDestroy(st));
    CK(cudaFree(d_out));
    return KEM_OK;
}

int kem_hbm_copy_peak(int dev, double *gbs_out)
{
    ARG(gbs_out, "null output");
    CK(cudaSetDevice(dev));
    const size_t bytes = (size_t)1 << 30;
    double2 *a = nullptr, *b = nullptr;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMalloc(&b, bytes));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CK(cudaMemsetAsync(a, 0, bytes, st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const long long n2 = (long long)(bytes / sizeof(double2));
    double best = 1e30;
    for (int r = 0; r < 8; ++r) {
        CK(cudaEventRecord(e0, st));
        k_copy<<<148 * 16, 256, 0, st>>>(b, a, n2);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) best = std::min(best, (double)ms);
    }
    CK(cudaGetLastError());
    *gbs_out = 2.0 * (double)bytes / (best * 1e-3) / 1e9;
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
    CK(cudaStreamDestroy(st));
    CK(cudaFree(a));
    CK(cudaFree(b));
    return KEM_OK;
}

}  // extern "C"
