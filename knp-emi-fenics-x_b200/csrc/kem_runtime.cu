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
#include <map>
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

// f3: out[i] = a0 + sum_k coef[k] * in_k[i], the eliminated-ion concentration of
// update_pde_variables (utils.py:247-267): c_elim = -(1/z_e) (rho_z rho_tag + sum_k z_k c_k),
// evaluated per bulk DOF in the order the sum is written.  Streaming and HBM-bound
// (8 (n_terms + 1) bytes per DOF): double2 accesses when every pointer is 16-byte aligned.
constexpr int KEM_MAX_TERMS = 8;
struct AffineArgs {
    double a0;
    int n_terms;
    double coef[KEM_MAX_TERMS];
    const double *in[KEM_MAX_TERMS];
};

__global__ void k_affine_combine(double *__restrict__ out, const __grid_constant__ AffineArgs a, long long n,
                                 int vec2)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec2) {
        const long long n2 = n >> 1;
        for (; i < n2; i += stride) {
            double2 r = make_double2(a.a0, a.a0);
            for (int k = 0; k < a.n_terms; ++k) {
                const double2 v = reinterpret_cast<const double2 *>(a.in[k])[i];
                r.x = r.x + a.coef[k] * v.x;
                r.y = r.y + a.coef[k] * v.y;
            }
            reinterpret_cast<double2 *>(out)[i] = r;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) {
            double r = a.a0;
            for (int k = 0; k < a.n_terms; ++k) r = r + a.coef[k] * a.in[k][n - 1];
            out[n - 1] = r;
        }
        return;
    }
    for (; i < n; i += stride) {
        double r = a.a0;
        for (int k = 0; k < a.n_terms; ++k) r = r + a.coef[k] * a.in[k][i];
        out[i] = r;
    }
}

// the same combination taken at the bulk DOF of every membrane DOF: the trace of the
// eliminated ion lands in a table column without the bulk vector being formed first
__global__ void k_gather_affine(double *__restrict__ dst, const __grid_constant__ AffineArgs a,
                                const long long *__restrict__ map, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const long long j = map[i];
        double r = a.a0;
        for (int k = 0; k < a.n_terms; ++k) r = r + a.coef[k] * a.in[k][j];
        dst[i] = r;
    }
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
constexpr size_t STAGE_BYTES = 16u << 20;
constexpr int IO_MAX_CHUNKS = 96;
constexpr int KEM_MAX_MAPS = 16;
const int IO_TARGET_CHUNKS = [] {       // 16: measured best of 8/16/32/64 at 1e7 DOFs (profiles/r1_bench.md)
    const char *e = getenv("KNPEMI_IO_CHUNKS");
    return e && atoi(e) > 0 ? atoi(e) : 16;
}();
constexpr int64_t IO_MIN_CHUNK = 1 << 16;   // smallest tail chunk of the pipeline (0.5 MB per column)
// Columns below this size always go through the pinned staging buffer: one memcpy of a few KB
// costs less than asking the driver (twice) whether the caller's memory is page-locked.  The
// reference's real meshes have 124 - 2 952 membrane DOFs (1 - 24 KB per column), where the
// per-call latency of the setters and getters is what a PDE step pays.
constexpr size_t DIRECT_COPY_MIN_BYTES = 64 * 1024;

struct Shard {
    int dev = 0;
    int64_t begin = 0, n = 0;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    cudaStream_t s_in2 = nullptr;     // second host->device stream of kem_step_io
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
    bool stage_busy[N_STAGE] = {};    // a host->device DMA out of this slot may still be in flight
    int stage_next = 0;
    // per-chunk events of kem_step_io / a chunked kem_step
    std::vector<cudaEvent_t> io_in, io_k0, io_k1, io_out, io_in2;
    // DOF chunks of the last chunked step; `chunks_live` while nothing else has been enqueued
    // since, so that a getter may follow the kernel chunk by chunk instead of waiting for all
    std::vector<int64_t> ch_off, ch_len;
    bool chunks_live = false;
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
    std::vector<char> p_discarded;    // np: 1 = value dropped on request (KEM_UNREAD_DISCARD)
    int unread_policy = KEM_UNREAD_AUTO;
    bool shadow_pinned_io = true;     // KEM_UNREAD_AUTO: pinned inputs of kem_step_io to dead slots are shadowed
    std::vector<char> out_const_valid;   // per constant output slot: 1 = the column holds the literal
    int step_chunks = 1;              // kem_step: launch the range as this many chunks (getter overlap)
    int io_chunks = 0;                // kem_step_io: chunks per shard (0 = IO_TARGET_CHUNKS / KNPEMI_IO_CHUNKS)
    int io_h2d_streams = 0;           // kem_step_io: 1 or 2 host->device streams (0 = default / KNPEMI_IO_H2D_STREAMS)
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

// ---- page-locked host ranges -------------------------------------------------------
// A host buffer takes the direct DMA path only if EVERY byte of it is page-locked.
// cudaHostRegister pins whole pages, so a small array can start inside a page that a
// registered neighbour pinned while its tail is pageable: the first byte alone proves
// nothing.  Ranges this library pinned itself (kem_host_alloc, kem_host_register) are
// kept in a process-wide, reference-counted table and answer by containment; memory
// somebody else pinned (torch, the caller) is accepted when the driver reports one
// allocation range that covers the whole buffer.
struct HostRange {
    size_t bytes = 0;
    int refs = 0;
    bool registered = false;   // cudaHostRegister by this library (unregister on last release)
};
std::map<uintptr_t, HostRange> g_host_ranges;     // keyed by base address
std::map<uintptr_t, uintptr_t> g_host_aliases;    // registered sub-range -> base of the owning range
std::mutex g_host_mu;

bool in_own_range(const void *p, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_host_mu);
    const uintptr_t a = (uintptr_t)p;
    auto it = g_host_ranges.upper_bound(a);
    if (it == g_host_ranges.begin()) return false;
    --it;
    return a >= it->first && a + bytes <= it->first + it->second.bytes;
}

// driver API, resolved at run time so that the library loads on a box without libcuda
typedef int (*cuPointerGetAttribute_fn)(void *data, int attribute, unsigned long long ptr);
cuPointerGetAttribute_fn driver_pointer_attribute()
{
    static cuPointerGetAttribute_fn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return (cuPointerGetAttribute_fn)f;
    }();
    return fn;
}

bool is_pinned(const void *p, size_t bytes)
{
    if (!p || bytes == 0) return false;
    if (in_own_range(p, bytes)) return true;
    const char *first = (const char *)p, *last = first + bytes - 1;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, first) != cudaSuccess ||
        cudaPointerGetAttributes(&a1, last) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost) return false;
    // both ends are page-locked: they must belong to ONE allocation that spans the buffer
    if (cuPointerGetAttribute_fn get = driver_pointer_attribute()) {
        unsigned long long start = 0;
        size_t size = 0;
        const int RANGE_START_ADDR = 11, RANGE_SIZE = 12;   // CU_POINTER_ATTRIBUTE_RANGE_*
        if (get(&start, RANGE_START_ADDR, (unsigned long long)(uintptr_t)first) == 0 &&
            get(&size, RANGE_SIZE, (unsigned long long)(uintptr_t)first) == 0)
            return (uintptr_t)first >= start && (uintptr_t)first + bytes <= start + size;
    }
    return false;   // cannot prove it: take the staged path (always correct)
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
    // The slots rotate across calls and a slot is waited for only when it comes round again:
    // the source has been consumed once it is in the staging buffer, so a run of small setter
    // calls (seven per PDE step, a few KB each on the reference's real meshes) never blocks on
    // its own DMAs.
    size_t off = 0;
    while (off < bytes) {
        const size_t len = std::min(STAGE_BYTES, bytes - off);
        const int slot = s.stage_next;
        s.stage_next = (slot + 1) % N_STAGE;
        if (s.stage_busy[slot]) CK(cudaEventSynchronize(s.stage_ev[slot]));
        CopyPool::get().copy(s.h_stage[slot], (const char *)src + off, len);
        CK(cudaMemcpyAsync((char *)dst + off, s.h_stage[slot], len, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(s.stage_ev[slot], st));
        s.stage_busy[slot] = true;
        off += len;
    }
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
    for (int j = 0; j < N_STAGE; ++j)          // uploads that still read from the staging slots
        if (s.stage_busy[j]) {
            CK(cudaEventSynchronize(s.stage_ev[j]));
            s.stage_busy[j] = false;
        }
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
    return h->p_uniform[col] || h->p_host[col] || h->p_discarded[col] ||
           (!h->shards.empty() && !h->shards[0].pcol[col]);
}

// make the parameter column a per-DOF device column holding its current value
int ensure_pcol(kem_handle h, int col)
{
    if (h->p_discarded[col])
        return fail(KEM_E_ARG, "parameter column " + std::to_string(col) +
                                   " was discarded (unread-input policy KEM_UNREAD_DISCARD): its value "
                                   "is not available; write it again under another policy");
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
        CK(cudaStreamSynchronize(s.s_in2));
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

// DOF chunks of one pipelined exchange (kem_step_io) or of a chunked kem_step: `target` equal
// chunks (each launch still fills the GPU for several waves, each column copy is several MB).
// With `taper` the tail is cut finer -- whenever at most four chunks of the current size
// remain the size is halved, down to IO_MIN_CHUNK -- so that the pipeline drains through a
// small last chunk.  Measured on B200 (profiles/r2_exchange.md): the drain it saves (0.5 ms of
// 10) is less than what the small copies cost, because a copy of 1 MB moves at 32 GB/s when
// both directions are busy and one of 5 MB at 48; equal chunks are the default,
// KNPEMI_IO_TAPER=1 turns the taper on.
bool taper_default()
{
    static const bool on = getenv("KNPEMI_IO_TAPER") && atoi(getenv("KNPEMI_IO_TAPER")) != 0;
    return on;
}

void plan_chunks(int64_t n, int target, std::vector<int64_t> &off, std::vector<int64_t> &len,
                 bool taper = taper_default())
{
    off.clear();
    len.clear();
    if (n <= 0) return;
    target = std::max(1, std::min(target, IO_MAX_CHUNKS / 2));
    int64_t chunk = std::max<int64_t>((n + target - 1) / target, 2 * IO_MIN_CHUNK);
    chunk = (chunk + 1023) / 1024 * 1024;
    int64_t at = 0;
    while (at < n) {
        const int64_t rem = n - at;
        if (taper && target > 1 && rem <= 4 * chunk && chunk > IO_MIN_CHUNK &&
            (int)off.size() + 8 < IO_MAX_CHUNKS) {
            chunk = std::max<int64_t>(((chunk / 2) + 1023) / 1024 * 1024, IO_MIN_CHUNK);
            continue;
        }
        const int64_t take = ((int)off.size() + 1 >= IO_MAX_CHUNKS) ? rem : std::min(chunk, rem);
        off.push_back(at);
        len.push_back(take);
        at += take;
    }
}

int ensure_chunk_events(Shard &s, size_t n_chunks)
{
    CK(cudaSetDevice(s.dev));
    while (s.io_in.size() < n_chunks) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e)); s.io_in.push_back(e);
        CK(cudaEventCreate(&e)); s.io_k0.push_back(e);
        CK(cudaEventCreate(&e)); s.io_k1.push_back(e);
        CK(cudaEventCreate(&e)); s.io_out.push_back(e);
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); s.io_in2.push_back(e);
    }
    return KEM_OK;
}

// something other than a chunked step is about to be enqueued: getters go back to stream order
void drop_live_chunks(kem_handle h)
{
    for (Shard &s : h->shards) s.chunks_live = false;
}

int const_out_index(kem_handle h, int kind, int col)
{
    if (kind != KEM_PARAM) return -1;
    for (int k = 0; k < h->m->n_const_out; ++k)
        if (h->m->const_out_cols[k] == col) return k;
    return -1;
}

// a caller wrote to parameter column `col`: it no longer provably holds the literal
void touch_param(kem_handle h, int kind, int col)
{
    const int k = const_out_index(h, kind, col);
    if (k >= 0) h->out_const_valid[k] = 0;
}

// every step stores the literals again
void mark_outputs_stored(kem_handle h)
{
    std::fill(h->out_const_valid.begin(), h->out_const_valid.end(), 1);
}

bool holds_literal(kem_handle h, int kind, int col, double *v)
{
    const int k = const_out_index(h, kind, col);
    if (k < 0 || !h->out_const_valid[k]) return false;
    *v = h->m->const_out_vals[k];
    return true;
}

// Where a full-column write to a slot the right-hand side never touches goes.
enum UnreadDest { DEST_SHADOW, DEST_UPLOAD, DEST_DISCARD };
UnreadDest unread_destination(kem_handle h, bool io_call, bool pinned_src)
{
    switch (h->unread_policy) {
        case KEM_UNREAD_SHADOW: return DEST_SHADOW;
        case KEM_UNREAD_UPLOAD: return DEST_UPLOAD;
        case KEM_UNREAD_DISCARD: return DEST_DISCARD;
        default: break;
    }
    // KEM_UNREAD_AUTO.  A plain setter runs alone: a pinned source goes over the idle link
    // faster (one DMA) than the host can copy it (80 MB: 1.5 ms against 3 ms), a pageable one
    // would need the same host copy into staging plus the DMA, so it is shadowed.  kem_step_io
    // overlaps everything: see shadow_pinned_inputs().
    if (!pinned_src) return DEST_SHADOW;
    if (!io_call) return DEST_UPLOAD;
    return h->shadow_pinned_io ? DEST_SHADOW : DEST_UPLOAD;
}

void discard_store(kem_handle h, int col)
{
    h->p_uniform[col] = 0;
    h->p_host[col] = 0;
    h->p_discarded[col] = 1;
    std::vector<double>().swap(h->p_shadow[col]);
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
        if (h->p_uniform[c] || h->p_host[c] || h->p_discarded[c]) {   // (shadowed / discarded slots are never read)
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
    if (flags & 3) {
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            CK(cudaMemsetAsync(s.d_flags, 0, sizeof(int), s.stream));
        }
        if (flags & 1) {
            g_err = "kem_step: a membrane state became non-finite";
            return KEM_NONFINITE;
        }
        g_err = "kem_step: the error-controlled integrator (KEM_SCHEME_DP45) could not reach t0+dt within "
                "its step limit / minimum step size at the requested tolerances";
        return KEM_STEP_FAILED;
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
    if (d->ns < 1 || d->np < 0 || d->n_out < 0 || d->n_out > 64 || !d->launch || !d->tonly ||
        d->n_const_out < 0 || d->n_const_out > d->n_out) {
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
    h->p_discarded.assign(m->np, 0);
    h->out_const_valid.assign(std::max(m->n_const_out, 0), 0);
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
        CKB(cudaStreamCreateWithFlags(&s.s_in2, cudaStreamNonBlocking));
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
        if (s.s_in2) cudaStreamSynchronize(s.s_in2);
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
        for (auto *v : {&s.io_in, &s.io_k0, &s.io_k1, &s.io_out, &s.io_in2})
            for (cudaEvent_t e : *v) cudaEventDestroy(e);
        for (cudaEvent_t e : {s.ev_a, s.ev_b, s.ev_c, s.ev_d, s.ev_t0, s.ev_t1}) if (e) cudaEventDestroy(e);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.s_in) cudaStreamDestroy(s.s_in);
        if (s.s_in2) cudaStreamDestroy(s.s_in2);
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
        h->p_discarded[col] = 0;
        std::vector<double>().swap(h->p_shadow[col]);
        h->uni_dirty = true;
        return KEM_OK;
    }
    if (kind == KEM_PARAM) h->uni[col] = v;
    touch_param(h, kind, col);
    drop_live_chunks(h);
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
    if (n == 0) return KEM_OK;
    const size_t col_bytes = (size_t)n * sizeof(double);
    const bool pinned = col_bytes >= DIRECT_COPY_MIN_BYTES && is_pinned(src, col_bytes);
    touch_param(h, kind, col);
    if (kind == KEM_PARAM && h->p_dead[col]) {
        switch (unread_destination(h, false, pinned)) {
            case DEST_SHADOW: shadow_store(h, col, src); return KEM_OK;
            case DEST_DISCARD: discard_store(h, col); return KEM_OK;
            case DEST_UPLOAD: break;
        }
    }
    drop_live_chunks(h);
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        // becomes a per-DOF column; no need to pre-fill, every row is overwritten
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            if (!s.pcol[col] && s.n > 0) CK(cudaMalloc(&s.pcol[col], (size_t)s.n * sizeof(double)));
        }
        h->p_uniform[col] = 0;
        h->p_host[col] = 0;
        h->p_discarded[col] = 0;
        std::vector<double>().swap(h->p_shadow[col]);
    }
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
    touch_param(h, kind, col);
    drop_live_chunks(h);
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
    touch_param(h, kind, col);
    drop_live_chunks(h);
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
    if (n == 0) return KEM_OK;
    double lit = 0.0;
    if (kind == KEM_PARAM && h->p_uniform[col]) {
        CopyPool::get().fill(dst, h->uni[col], (size_t)n);
        return KEM_OK;
    }
    if (holds_literal(h, kind, col, &lit)) {      // e.g. I_ch_Cl = 0.0: known without asking the device
        CopyPool::get().fill(dst, lit, (size_t)n);
        return KEM_OK;
    }
    if (kind == KEM_PARAM && h->p_host[col]) {
        CopyPool::get().copy(dst, h->p_shadow[col].data(), (size_t)n * sizeof(double));
        return KEM_OK;
    }
    if (kind == KEM_PARAM && h->p_discarded[col])
        return fail(KEM_E_ARG, "kem_get_column: parameter column " + std::to_string(col) +
                                   " was discarded (KEM_UNREAD_DISCARD); its value is not kept");
    const size_t col_bytes = (size_t)n * sizeof(double);
    const bool pinned = col_bytes >= DIRECT_COPY_MIN_BYTES && is_pinned(dst, col_bytes);
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        if (pinned && s.chunks_live) {
            // the last thing enqueued is a chunked step: follow it chunk by chunk on the
            // device->host stream instead of waiting for the whole range
            CK(cudaSetDevice(s.dev));
            for (size_t c = 0; c < s.ch_off.size(); ++c) {
                CK(cudaStreamWaitEvent(s.s_out, s.io_k1[c], 0));
                CK(cudaMemcpyAsync(dst + s.begin + s.ch_off[c], col_ptr(s, kind, col) + s.ch_off[c],
                                   (size_t)s.ch_len[c] * sizeof(double), cudaMemcpyDeviceToHost, s.s_out));
            }
            continue;
        }
        rc = copy_out(s, dst + s.begin, col_ptr(s, kind, col), (size_t)s.n * sizeof(double), s.stream,
                      pinned);
        if (rc) return rc;
    }
    if (pinned)
        for (Shard &s : h->shards) {
            if (s.n == 0) continue;
            CK(cudaSetDevice(s.dev));
            CK(cudaStreamSynchronize(s.chunks_live ? s.s_out : s.stream));
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
    else *location_out = h->p_uniform[col] ? 0 : (h->p_host[col] ? 2 : (h->p_discarded[col] ? 3 : 1));
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
// Enqueue the step of shard `s` as the chunks of its plan, alternating over the two compute
// streams (one chunk's tail wave overlaps the next one's head), each chunk after `after[c]`
// if given.  Records io_k0/io_k1 per chunk; `stream` is made to wait for all of them.
static int launch_chunked(kem_handle h, Shard &s, const StepPlan &pl, const cudaEvent_t *after)
{
    const size_t n_chunks = s.ch_off.size();
    static const bool two = getenv("KNPEMI_IO_ONE_COMPUTE_STREAM") == nullptr;
    CK(cudaSetDevice(s.dev));
    for (size_t c = 0; c < n_chunks; ++c) {
        cudaStream_t sk = (two && (c & 1)) ? s.stream2 : s.stream;
        if (after) CK(cudaStreamWaitEvent(sk, after[c], 0));
        CK(cudaEventRecord(s.io_k0[c], sk));
        int rc = launch_range(h, s, pl, s.ch_off[c], s.ch_len[c], sk);
        if (rc) return rc;
        CK(cudaEventRecord(s.io_k1[c], sk));
    }
    return KEM_OK;
}

int kem_step_timed(kem_handle h, double t0, double dt, int n_sub, int scheme, int n_stim,
                   const int *stim_cols, const double *stim_vals, int *status_flags,
                   kem_step_times *times)
{
    NvtxRange nvtx_range("kem_step_timed");
    StepPlan pl;
    int rc = prepare_step(h, t0, dt, n_sub, scheme, n_stim, stim_cols, stim_vals, pl);
    if (rc) return rc;
    drop_live_chunks(h);
    // a chunked launch lets the getters that follow copy chunk c while chunk c+1 still runs
    // (kem_set_step_chunks; scheme O3 sorts the whole range and stays one launch)
    const bool chunked = h->step_chunks > 1 && scheme == KEM_SCHEME_RK4;
    for (Shard &s : h->shards) {
        if (times) {
            CK(cudaSetDevice(s.dev));
            CK(cudaEventRecord(s.ev_a, s.stream));
        }
        if (chunked && s.n >= 4 * IO_MIN_CHUNK) {
            plan_chunks(s.n, h->step_chunks, s.ch_off, s.ch_len);
            rc = ensure_chunk_events(s, s.ch_off.size());
            if (rc) return rc;
            CK(cudaEventRecord(s.ev_c, s.stream));          // stream2 must see the table uploads
            CK(cudaStreamWaitEvent(s.stream2, s.ev_c, 0));
            rc = launch_chunked(h, s, pl, nullptr);
            if (rc) return rc;
            const size_t nc = s.ch_off.size();
            if (nc >= 2) CK(cudaStreamWaitEvent(s.stream, s.io_k1[(nc - 1) & 1 ? nc - 1 : nc - 2], 0));
            s.chunks_live = true;
        } else {
            rc = launch_range(h, s, pl, 0, s.n);
            if (rc) return rc;
        }
        if (times) CK(cudaEventRecord(s.ev_b, s.stream));
    }
    mark_outputs_stored(h);
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

// One whole PDE -> ODE -> PDE exchange.  Order of effects, as in the reference: the input
// columns are written first (the setters of utils.py:227-233), then the step applies the
// sticky stimulus and integrates (odeSolver.py:108-122), then the outputs are read
// (run_2D.py:105-109).  Everything is validated and classified before anything changes.
int kem_step_io(kem_handle h, double t0, double dt, int n_sub, int scheme, int n_stim,
                const int *stim_cols, const double *stim_vals, int n_in, const kem_io_column *in,
                int n_out, const kem_io_column *out, int *status_flags, kem_step_times *times)
{
    NvtxRange nvtx_range("kem_step_io");
    ARG(h, "null handle");
    ARG(n_in >= 0 && n_out >= 0 && (in || !n_in) && (out || !n_out), "bad io arrays");
    ARG(n_stim >= 0 && n_stim <= KEM_MAX_STIM && (n_stim == 0 || (stim_cols && stim_vals)),
        "bad stimulus arrays");
    int rc;
    const size_t col_bytes = (size_t)h->n * sizeof(double);
    const bool masked = !h->shards.empty() && h->shards[0].has_mask;
    // what the stimulus of this step does to a parameter column: 0 nothing, 1 rows under the
    // mask, 2 every row (the column ends up uniform, whatever the inputs say)
    auto stim_effect = [&](int kind, int col) {
        if (kind != KEM_PARAM) return 0;
        for (int k = 0; k < n_stim; ++k)
            if (stim_cols[k] == col) return masked ? 1 : 2;
        return 0;
    };
    auto stim_value = [&](int col) {
        double v = 0.0;
        for (int k = 0; k < n_stim; ++k)
            if (stim_cols[k] == col) v = stim_vals[k];
        return v;
    };

    // ---- 1. validate and classify; the handle is not modified in this phase
    struct FillOut { double *host; double value; };
    std::vector<kem_io_column> dev_in, dev_out, shadow_in, shadow_out, discard_in;
    std::vector<FillOut> fill_out;
    bool all_pinned = true;
    for (int k = 0; k < n_in; ++k) {
        rc = check_col(h, in[k].kind, in[k].col, __func__);
        if (rc) return rc;
        ARG(in[k].host || h->n == 0, "null input column");
        if (h->n == 0) continue;
        const int eff = stim_effect(in[k].kind, in[k].col);
        if (eff == 2) continue;                     // overwritten on every row by the stimulus
        const bool pinned = is_pinned(in[k].host, col_bytes);
        if (in[k].kind == KEM_PARAM && h->p_dead[in[k].col] && eff == 0) {
            const UnreadDest d = unread_destination(h, true, pinned);
            if (d == DEST_SHADOW) { shadow_in.push_back(in[k]); continue; }
            if (d == DEST_DISCARD) { discard_in.push_back(in[k]); continue; }
        }
        dev_in.push_back(in[k]);
        all_pinned = all_pinned && pinned;
    }
    auto listed = [](const std::vector<kem_io_column> &v, int kind, int col) {
        for (const kem_io_column &c : v)
            if (c.kind == kind && c.col == col) return true;
        return false;
    };
    for (int k = 0; k < n_out; ++k) {
        rc = check_col(h, out[k].kind, out[k].col, __func__);
        if (rc) return rc;
        ARG(out[k].host || h->n == 0, "null output column");
        if (h->n == 0) continue;
        const int kind = out[k].kind, col = out[k].col;
        const int ci = const_out_index(h, kind, col);
        if (ci >= 0) {                              // stored as a literal by every step
            fill_out.push_back({out[k].host, h->m->const_out_vals[ci]});
            continue;
        }
        const int eff = stim_effect(kind, col);
        if (eff == 2) {                             // uniform after this step's stimulus
            fill_out.push_back({out[k].host, stim_value(col)});
            continue;
        }
        const bool as_input = listed(dev_in, kind, col);
        if (kind == KEM_PARAM && listed(discard_in, kind, col))
            return fail(KEM_E_ARG, "kem_step_io: an output column is discarded by this call's inputs "
                                   "(KEM_UNREAD_DISCARD)");
        if (kind == KEM_PARAM && listed(shadow_in, kind, col)) { shadow_out.push_back(out[k]); continue; }
        if (kind == KEM_PARAM && !as_input && eff == 0) {
            if (h->p_uniform[col]) { fill_out.push_back({out[k].host, h->uni[col]}); continue; }
            if (h->p_host[col]) { shadow_out.push_back(out[k]); continue; }
            if (h->p_discarded[col])
                return fail(KEM_E_ARG, "kem_step_io: output column was discarded (KEM_UNREAD_DISCARD)");
        }
        dev_out.push_back(out[k]);
        all_pinned = all_pinned && is_pinned(out[k].host, col_bytes);
    }

    // ---- 2. stimulus, time tables (may make a stimulus column uniform or per-DOF)
    // A discarded slot that is a masked stimulus target AND an input of this call gets its
    // default back first, so that prepare_step can materialise it; the input then overwrites it.
    for (const kem_io_column &c : dev_in)
        if (c.kind == KEM_PARAM && h->p_discarded[c.col] && stim_effect(c.kind, c.col) == 1) {
            h->p_discarded[c.col] = 0;
            h->p_uniform[c.col] = 1;
            h->uni_dirty = true;
        }
    StepPlan pl;
    rc = prepare_step(h, t0, dt, n_sub, scheme, n_stim, stim_cols, stim_vals, pl);
    if (rc) return rc;
    drop_live_chunks(h);

    // ---- 3. residency of the input columns, just before their copies are enqueued
    for (const kem_io_column &c : dev_in) {
        touch_param(h, c.kind, c.col);
        if (c.kind != KEM_PARAM || !not_on_device(h, c.col)) continue;
        for (Shard &s : h->shards) {
            CK(cudaSetDevice(s.dev));
            if (!s.pcol[c.col] && s.n > 0) CK(cudaMalloc(&s.pcol[c.col], (size_t)s.n * sizeof(double)));
        }
        h->p_uniform[c.col] = 0;                    // every row is overwritten by the copy below
        h->p_host[c.col] = 0;
        h->p_discarded[c.col] = 0;
        std::vector<double>().swap(h->p_shadow[c.col]);
    }
    for (const kem_io_column &c : discard_in) discard_store(h, c.col);
    in = dev_in.data();
    n_in = (int)dev_in.size();
    out = dev_out.data();
    n_out = (int)dev_out.size();
    // host-side part of the exchange; runs while the devices work (the calls below only enqueue)
    auto host_side = [&]() {
        for (const kem_io_column &c : shadow_in) shadow_store(h, c.col, c.host);
        for (const kem_io_column &c : shadow_out)
            CopyPool::get().copy(c.host, h->p_shadow[c.col].data(), col_bytes);
        for (const FillOut &f : fill_out) CopyPool::get().fill(f.host, f.value, (size_t)h->n);
    };

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
        mark_outputs_stored(h);
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

    // pinned host buffers: DOF-chunked pipeline, H2D (s_in) | kernel (stream, stream2) | D2H (s_out)
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        CK(cudaSetDevice(s.dev));
        plan_chunks(s.n, h->io_chunks > 0 ? h->io_chunks : IO_TARGET_CHUNKS, s.ch_off, s.ch_len);
        const size_t n_chunks = s.ch_off.size();
        rc = ensure_chunk_events(s, n_chunks);
        if (rc) return rc;
        // the copy streams must see the table/uniform uploads and earlier work on `stream`
        CK(cudaEventRecord(s.ev_a, s.stream));
        CK(cudaStreamWaitEvent(s.s_in, s.ev_a, 0));
        CK(cudaStreamWaitEvent(s.stream2, s.ev_a, 0));
        CK(cudaStreamWaitEvent(s.s_out, s.ev_a, 0));
        CK(cudaEventRecord(s.ev_b, s.s_in));   // t = 0 of this shard's exchange
        // the input columns alternate over two copy streams: more read requests in flight while
        // the device->host writes share the link (measured -4 % per exchange, profiles/r2_exchange.md;
        // KNPEMI_IO_H2D_STREAMS=1 puts them back on one)
        static const bool two_in_env = !(getenv("KNPEMI_IO_H2D_STREAMS") && atoi(getenv("KNPEMI_IO_H2D_STREAMS")) == 1);
        const bool two_in = h->io_h2d_streams ? h->io_h2d_streams == 2 : two_in_env;
        if (two_in) CK(cudaStreamWaitEvent(s.s_in2, s.ev_a, 0));
        for (size_t c = 0; c < n_chunks; ++c) {
            for (int k = 0; k < n_in; ++k)
                CK(cudaMemcpyAsync(col_ptr(s, in[k].kind, in[k].col) + s.ch_off[c],
                                   in[k].host + s.begin + s.ch_off[c],
                                   (size_t)s.ch_len[c] * sizeof(double), cudaMemcpyHostToDevice,
                                   (two_in && (k & 1)) ? s.s_in2 : s.s_in));
            if (two_in) {
                CK(cudaEventRecord(s.io_in2[c], s.s_in2));
                CK(cudaStreamWaitEvent(s.s_in, s.io_in2[c], 0));
            }
            CK(cudaEventRecord(s.io_in[c], s.s_in));
        }
        rc = launch_chunked(h, s, pl, s.io_in.data());
        if (rc) return rc;
        for (size_t c = 0; c < n_chunks; ++c) {
            CK(cudaStreamWaitEvent(s.s_out, s.io_k1[c], 0));
            for (int k = 0; k < n_out; ++k)
                CK(cudaMemcpyAsync(out[k].host + s.begin + s.ch_off[c],
                                   col_ptr(s, out[k].kind, out[k].col) + s.ch_off[c],
                                   (size_t)s.ch_len[c] * sizeof(double), cudaMemcpyDeviceToHost, s.s_out));
            CK(cudaEventRecord(s.io_out[c], s.s_out));
        }
        // later work on `stream` (the next step) must not overtake the kernels of stream2 or the
        // D2H copies (io_out of the last chunk follows every kernel chunk)
        CK(cudaStreamWaitEvent(s.stream, s.io_out[n_chunks - 1], 0));
    }
    mark_outputs_stored(h);
    host_side();
    if (!status_flags && !times) {
        // enqueue-only (like kem_step with NULL flags): the caller keeps the pinned buffers
        // untouched until a synchronising call; a getter that follows copies chunk by chunk
        for (Shard &s : h->shards) s.chunks_live = s.n > 0;
        return KEM_OK;
    }
    if (times) memset(times, 0, sizeof *times);
    for (Shard &s : h->shards) {
        if (s.n == 0) continue;
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_out));
        if (times) {
            const size_t n_chunks = s.ch_off.size();
            float tot = 0, kern = 0, h2d = 0, d2h = 0, f = 0;
            CK(cudaEventElapsedTime(&tot, s.ev_b, s.io_out[n_chunks - 1]));
            CK(cudaEventElapsedTime(&h2d, s.ev_b, s.io_in[n_chunks - 1]));
            // chunk kernels overlap across the two compute streams: report the span from the
            // first kernel's start to the last kernel's end, not the sum
            for (size_t c = n_chunks >= 2 ? n_chunks - 2 : 0; c < n_chunks; ++c) {
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

int kem_set_unread_policy(kem_handle h, int policy)
{
    ARG(h, "null handle");
    ARG(policy >= KEM_UNREAD_AUTO && policy <= KEM_UNREAD_DISCARD, "unknown policy");
    h->unread_policy = policy;
    return KEM_OK;
}

int kem_plan_chunks(int64_t n, int target, int taper, int64_t *off_out, int64_t *len_out, int cap,
                    int *count_out)
{
    ARG(count_out && n >= 0, "bad arguments");
    std::vector<int64_t> off, len;
    plan_chunks(n, target, off, len, taper < 0 ? taper_default() : taper != 0);
    *count_out = (int)off.size();
    for (int k = 0; k < std::min<int>(cap, (int)off.size()); ++k) {
        if (off_out) off_out[k] = off[k];
        if (len_out) len_out[k] = len[k];
    }
    return KEM_OK;
}

int kem_set_io_tuning(kem_handle h, int n_chunks, int h2d_streams)
{
    ARG(h, "null handle");
    ARG(n_chunks >= 0 && n_chunks <= IO_MAX_CHUNKS / 2, "n_chunks out of range");
    ARG(h2d_streams >= 0 && h2d_streams <= 2, "h2d_streams must be 0 (default), 1 or 2");
    h->io_chunks = n_chunks;
    h->io_h2d_streams = h2d_streams;
    return KEM_OK;
}

int kem_set_step_chunks(kem_handle h, int n_chunks)
{
    ARG(h, "null handle");
    ARG(n_chunks >= 1 && n_chunks <= IO_MAX_CHUNKS / 2, "n_chunks out of range");
    h->step_chunks = n_chunks;
    return KEM_OK;
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
    if (h) { touch_param(h, kind, col); drop_live_chunks(h); }
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
    if (h) { touch_param(h, kind, col); drop_live_chunks(h); }
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

static int fill_affine(AffineArgs &a, double a0, int n_terms, const double *coef,
                       const double *const *dev_in, const char *fn)
{
    if (n_terms < 0 || n_terms > KEM_MAX_TERMS || (n_terms && (!coef || !dev_in)))
        return fail(KEM_E_ARG, std::string(fn) + ": 0.." + std::to_string(KEM_MAX_TERMS) + " terms");
    memset(&a, 0, sizeof a);
    a.a0 = a0;
    a.n_terms = n_terms;
    for (int k = 0; k < n_terms; ++k) {
        if (!dev_in[k]) return fail(KEM_E_ARG, std::string(fn) + ": null device pointer");
        a.coef[k] = coef[k];
        a.in[k] = dev_in[k];
    }
    return KEM_OK;
}

int kem_device_affine_combine(int dev, int64_t n, double *dev_out, double a0, int n_terms,
                              const double *coef, const double *const *dev_in)
{
    ARG(n >= 0 && (dev_out || n == 0), "bad output");
    AffineArgs a;
    int rc = fill_affine(a, a0, n_terms, coef, dev_in, __func__);
    if (rc) return rc;
    if (n == 0) return KEM_OK;
    CK(cudaSetDevice(dev));
    bool aligned = ((uintptr_t)dev_out & 15) == 0;
    for (int k = 0; k < n_terms; ++k) aligned = aligned && ((uintptr_t)dev_in[k] & 15) == 0;
    // legacy default stream: ordered after the caller's earlier default-stream work
    k_affine_combine<<<grid_for(aligned ? n / 2 + 1 : n), 256>>>(dev_out, a, n, aligned ? 1 : 0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(0));
    return KEM_OK;
}

int kem_device_gather_affine(kem_handle h, int shard, int kind, int col, double a0, int n_terms,
                             const double *coef, const double *const *dev_in, int map_id)
{
    int rc = device_xfer_check(h, shard, kind, col, n_terms > 0 && dev_in ? dev_in[0] : (const void *)h, map_id,
                               __func__);
    if (rc) return rc;
    AffineArgs a;
    rc = fill_affine(a, a0, n_terms, coef, dev_in, __func__);
    if (rc) return rc;
    touch_param(h, kind, col);
    drop_live_chunks(h);
    if (kind == KEM_PARAM && not_on_device(h, col)) {
        if (h->p_discarded[col]) {                 // every row is overwritten below
            h->p_discarded[col] = 0;
            h->p_uniform[col] = 1;
        }
        rc = ensure_pcol(h, col);
        if (rc) return rc;
    }
    Shard &s = h->shards[shard];
    if (s.n == 0) return KEM_OK;
    CK(cudaSetDevice(s.dev));
    k_gather_affine<<<grid_for(s.n), 256, 0, s.stream>>>(col_ptr(s, kind, col), a, s.d_map[map_id], s.n);
    CK(cudaGetLastError());
    h->launches++;
    return KEM_OK;
}

int kem_device_copy_in(kem_handle h, int shard, int kind, int col, const double *dev_src)
{
    if (h) { touch_param(h, kind, col); drop_live_chunks(h); }
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
    bytes = std::max<size_t>(bytes, 8);
    CK(cudaHostAlloc(ptr_out, bytes, cudaHostAllocPortable));
    std::lock_guard<std::mutex> lk(g_host_mu);
    HostRange &r = g_host_ranges[(uintptr_t)*ptr_out];
    r.bytes = bytes;
    r.refs = 1;
    r.registered = false;
    return KEM_OK;
}

int kem_host_free(void *ptr)
{
    if (!ptr) return KEM_OK;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        g_host_ranges.erase((uintptr_t)ptr);
    }
    CK(cudaFreeHost(ptr));
    return KEM_OK;
}

// Page-lock memory the caller owns (the array behind a dolfinx Function): afterwards the
// setters, getters and kem_step_io see it as pinned and DMA directly instead of staging.
// Registrations are reference-counted per base address across all handles of the process:
// two models that register the same array share one cudaHostRegister, and the memory is
// unpinned when the last of them releases it.
int kem_host_register(void *ptr, size_t bytes)
{
    ARG(ptr && bytes > 0, "null pointer or zero size");
    const uintptr_t a = (uintptr_t)ptr;
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto it = g_host_ranges.find(a);
    if (it != g_host_ranges.end()) {
        ARG(bytes <= it->second.bytes, "range is registered with a smaller size; unregister it first");
        it->second.refs++;
        return KEM_OK;
    }
    // a different base inside / across a range of ours: the caller would end up with a
    // partially pinned buffer; say so instead of pretending
    auto up = g_host_ranges.upper_bound(a);
    if (up != g_host_ranges.end() && up->first < a + bytes)
        return fail(KEM_E_ARG, "kem_host_register: range overlaps a registered range");
    if (up != g_host_ranges.begin()) {
        auto prev = std::prev(up);
        if (prev->first + prev->second.bytes > a) {
            if (a + bytes <= prev->first + prev->second.bytes) {   // a view into a pinned range
                if (!prev->second.registered) return KEM_OK;       // kem_host_alloc memory: nothing to count
                if (g_host_aliases.count(a))
                    return fail(KEM_E_ARG, "kem_host_register: sub-range is already registered");
                prev->second.refs++;
                g_host_aliases[a] = prev->first;                   // unregister(ptr) finds the owner
                return KEM_OK;
            }
            return fail(KEM_E_ARG, "kem_host_register: range overlaps a registered range");
        }
    }
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        // pinned by somebody else (not through this library): usable only if that pinning
        // covers the whole buffer, which is_pinned() checks on every transfer
        cudaGetLastError();
        return fail(KEM_E_ARG, "kem_host_register: memory is (partly) page-locked by another owner");
    }
    CK(e);
    HostRange &r = g_host_ranges[a];
    r.bytes = bytes;
    r.refs = 1;
    r.registered = true;
    return KEM_OK;
}

int kem_host_unregister(void *ptr)
{
    ARG(ptr, "null pointer");
    uintptr_t a = (uintptr_t)ptr;
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto al = g_host_aliases.find(a);
    if (al != g_host_aliases.end()) {
        a = al->second;
        g_host_aliases.erase(al);
    }
    auto it = g_host_ranges.find(a);
    if (it == g_host_ranges.end() || !it->second.registered) return KEM_OK;   // unknown memory: no-op
    if (--it->second.refs > 0) return KEM_OK;
    g_host_ranges.erase(it);
    cudaError_t e = cudaHostUnregister((void *)a);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return KEM_OK;
    }
    CK(e);
    return KEM_OK;
}

int kem_host_is_pinned(const void *ptr, size_t bytes, int *pinned_out)
{
    ARG(ptr && pinned_out, "null argument");
    *pinned_out = is_pinned(ptr, std::max<size_t>(bytes, 1)) ? 1 : 0;
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

// Host-link ceiling of device `dev`: `reps_h2d` copies of `bytes` host->device on one stream
// and `reps_d2h` copies device->host on another, enqueued interleaved (start skew ~10 us), pinned
// host memory allocated by the calling thread (so its NUMA placement is the bench's).  Successive
// copies walk through `span_bytes` of host memory per direction (0 = reuse one buffer): with a
// span far above the CPU's last-level cache the copies stream through DRAM like the exchange
// of a real PDE step does; with one small buffer they are served from the cache (DDIO) and show
// the PCIe link alone.  A direction's elapsed time is first copy start -> its last copy end.
// With unequal rep counts the shorter direction is measured entirely under the other's traffic.
int kem_link_probe(int dev, size_t bytes, size_t span_bytes, int reps_h2d, int reps_d2h,
                   double *ms_h2d_out, double *ms_d2h_out)
{
    ARG(bytes >= 8 && reps_h2d >= 0 && reps_d2h >= 0 && reps_h2d + reps_d2h > 0, "bad probe size");
    ARG(ms_h2d_out && ms_d2h_out, "null output");
    if (span_bytes < bytes) span_bytes = bytes;
    const size_t slots = span_bytes / bytes;
    span_bytes = slots * bytes;
    CK(cudaSetDevice(dev));
    char *h_a = nullptr, *h_b = nullptr;
    void *d_a = nullptr, *d_b = nullptr;
    CK(cudaHostAlloc((void **)&h_a, span_bytes, cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&h_b, span_bytes, cudaHostAllocDefault));
    CopyPool::get().fill((double *)h_a, 1.0, span_bytes / sizeof(double));     // touch every page
    CopyPool::get().fill((double *)h_b, 2.0, span_bytes / sizeof(double));
    CK(cudaMalloc(&d_a, bytes));
    CK(cudaMalloc(&d_b, bytes));
    CK(cudaMemset(d_b, 0, bytes));
    cudaStream_t s_in, s_out;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    cudaEvent_t i0, i1, o0, o1;
    CK(cudaEventCreate(&i0));
    CK(cudaEventCreate(&i1));
    CK(cudaEventCreate(&o0));
    CK(cudaEventCreate(&o1));
    // warm-up in both directions (first touch of the mappings)
    CK(cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s_in));
    CK(cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s_out));
    CK(cudaStreamSynchronize(s_in));
    CK(cudaStreamSynchronize(s_out));
    CK(cudaEventRecord(i0, s_in));
    CK(cudaEventRecord(o0, s_out));
    for (int k = 0; k < std::max(reps_h2d, reps_d2h); ++k) {      // interleaved enqueue
        const size_t off = ((size_t)k % slots) * bytes;
        if (k < reps_h2d) CK(cudaMemcpyAsync(d_a, h_a + off, bytes, cudaMemcpyHostToDevice, s_in));
        if (k < reps_d2h) CK(cudaMemcpyAsync(h_b + off, d_b, bytes, cudaMemcpyDeviceToHost, s_out));
    }
    CK(cudaEventRecord(i1, s_in));
    CK(cudaEventRecord(o1, s_out));
    CK(cudaEventSynchronize(i1));
    CK(cudaEventSynchronize(o1));
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, i0, i1));
    CK(cudaEventElapsedTime(&b, o0, o1));
    *ms_h2d_out = reps_h2d ? (double)a : 0.0;
    *ms_d2h_out = reps_d2h ? (double)b : 0.0;
    for (cudaEvent_t e : {i0, i1, o0, o1}) CK(cudaEventDestroy(e));
    for (cudaStream_t s : {s_in, s_out}) CK(cudaStreamDestroy(s));
    CK(cudaFree(d_a));
    CK(cudaFree(d_b));
    CK(cudaFreeHost(h_a));
    CK(cudaFreeHost(h_b));
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
