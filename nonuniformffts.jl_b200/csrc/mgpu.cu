// mgpu.cu — the multi-GPU transforms behind the C ABI (include/nufft_b200.h, nufft_mgpu_*): one B200 per rank, NCCL over
// NVLink 5 / NVSwitch for the exchanges.  The reference has no distributed code; this is the sharding the hot
// path offers, in three strategies:
//
//   NUFFT_MGPU_SLAB        z-slab spatial decomposition (large single transforms, C5).  Rank r owns the planes
//                          [r Nz/G, (r + 1) Nz/G) of the oversampled grid plus 2M - 1 halo planes.
//                            set_points  every rank sorts ITS points by owner slab (one counting pass) and the coordinates
//                                        are exchanged all-to-all (12 bytes per point); each rank then bins the points of its
//                                        slab with the single-GPU set_points (K-bin).
//                            type 1      values follow the points (8 bytes per point) -> ring-window spreading into the slab ->
//                                        halo planes are sent to the two neighbours and added (2M - 1 planes per rank) ->
//                                        pruned FFT passes along x and y on the slab (K-pfft) -> all-to-all transpose of the
//                                        1/4-size intermediate (z slabs -> y slabs of kept modes) -> pass along z with the
//                                        deconvolution and normalisation fused.  Output: rank r holds the kept modes
//                                        [:, r Ky/G : (r + 1) Ky/G, :] (nufft_mgpu_local_block).
//                            type 2      the same in reverse: z pass -> transpose -> y, x passes -> halo planes copied from
//                                        the neighbours -> interpolation at the slab's points -> values return to their owners.
//                          Per-GPU work and memory are 1/G of the single-GPU transform; the wire carries 20 bytes per point
//                          and 1/4 of the grid instead of the whole grid of a partial-grid reduction.
//   NUFFT_MGPU_POINTS      points partitioned, every rank transforms its points on a FULL grid; type 1 sums the partial
//                          outputs (ncclAllReduce of prod(size(p)) values: FFT and deconvolution are linear), type 2
//                          broadcasts the spectrum.  Any plan the single-GPU library supports.
//   NUFFT_MGPU_TRANSFORMS  independent ntransforms dealt round-robin to ranks (c mod G == r); no data-path collective.
//
// One handle drives `nlocal` ranks of the job: nlocal == nranks in a single process (one host thread, all GPUs of the box:
// what a Julia host does through ccall), nlocal == 1 under one process per GPU (torchrun; the NCCL id travels through the
// host's own channel).  Every stage is issued for all local ranks before the next one starts, NCCL calls inside
// ncclGroupStart / ncclGroupEnd, so a single thread never blocks on a peer it has not served yet.  NCCL is loaded at run time
// (dlopen, reusing the copy the process already has): single-GPU users do not need it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>
#include <new>
#include <vector>

#include "common.cuh"
#include "kernel_eval.cuh"

namespace nufft {

int pfft_pass(Plan &p, int d, bool fwd, const void *in, void *out, int64_t n_lo, int64_t n_hi, double scale, int blk,
              int64_t blk_stride);      // pfft.cu

// ---- NCCL through dlopen -----------------------------------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    bool tried = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl_api()
{
    static NcclApi a;
    if (a.tried) return a;
    a.tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {                   // the copy already in the process (e.g. PyTorch's) first
        a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
        if (a.lib) break;
    }
    if (!a.lib && getenv("NUFFT_B200_NCCL")) a.lib = dlopen(getenv("NUFFT_B200_NCCL"), RTLD_NOW | RTLD_GLOBAL);
    for (const char *n : names) {
        if (a.lib) break;
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!a.lib) return a;
    bool ok = true;
    auto sym = [&](const char *name) -> void * {
        void *s = dlsym(a.lib, name);
        if (!s) ok = false;
        return s;
    };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.Broadcast = (decltype(a.Broadcast))sym("ncclBroadcast");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { a.lib = nullptr; }
    return a;
}

#define NCCL_TRY(expr)                                                                                      \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess) {                                                                            \
            ::nufft::set_error("NCCL error %d at %s:%d: %s", (int)_r, __FILE__, __LINE__,                   \
                               nccl_api().GetErrorString ? nccl_api().GetErrorString(_r) : "?");            \
            return NUFFT_ERR_CUDA;                                                                          \
        }                                                                                                   \
    } while (0)

// ---- per-rank state --------------------------------------------------------------------------------------------------------
constexpr int MG_MAX_RANKS = 16;
struct DestOffsets {
    unsigned long long off[MG_MAX_RANKS];
};

// peer windows: the receive buffers of the z-slab exchanges, written directly by the sending rank (cudaMemcpyAsync over NVLink
// into memory mapped with CUDA IPC, or plain peer access inside one process) instead of NCCL send / recv pairs
enum { W_RX0 = 0, W_RX1, W_RX2, W_RV, W_SV, W_HALO, W_GRID, W_A, W_N };
constexpr unsigned W_DYNAMIC = 0x1fu, W_STATIC = 0xe0u;      // sized by the points (reallocated as they grow) / by the plan
struct WinMsg {                                            // what a rank publishes about its windows
    cudaIpcMemHandle_t h[W_N];
    unsigned long long ptr[W_N];
    int32_t ok, pad;
};

struct MgRank {
    int rank = 0, dev = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    ncclComm_t comm = nullptr;
    Plan *plan = nullptr;
    // slab
    int z0 = 0, nz = 0;
    int64_t np_user = -1, np_slab = 0, cap_user = 0, cap_slab = 0;
    int32_t *d_sendperm = nullptr;
    float *d_sx[3] = {nullptr, nullptr, nullptr}, *d_rx[3] = {nullptr, nullptr, nullptr};
    void *d_sv = nullptr, *d_rv = nullptr;            // values grouped by destination / of the slab's points
    unsigned long long *d_cnt = nullptr;               // [G] send counts | [G] cursors | [G * G] count matrix
    unsigned long long *h_cnt = nullptr;               // pinned [G * G]
    std::vector<int64_t> sendcnt, sendoff, recvcnt, recvoff;
    void *d_halo = nullptr;                            // 2M - 1 planes received from the neighbours
    void *d_gather = nullptr;                          // nufft_mgpu_gather_output scratch
    void *win[W_N][MG_MAX_RANKS] = {};                 // rank r's window buffer as this rank's device addresses it
    bool win_ipc[W_N][MG_MAX_RANKS] = {};              // mapped with cudaIpcOpenMemHandle (closed before the owner frees it)
    WinMsg *d_msg = nullptr, *h_msg = nullptr;         // [1 + G]: own message, all-gathered messages
    int32_t *d_bar = nullptr, *h_flag = nullptr;       // barrier scratch / agreement flag
    std::vector<int64_t> put_fwd, put_bwd;             // where my block starts inside rank r's receive buffer (elements)
    std::vector<int> comps;                            // TRANSFORMS: components of this rank
    cudaEvent_t ev[24] = {};
    bool ev_ok = false;
    float ms[16] = {};
};

struct Mgpu {
    nufft_opts opts{};
    int nranks = 1, strategy = 0;
    std::vector<MgRank> L;
    int D = 0, C = 1;
    bool f64 = false, cplx = true;
    int64_t nk[3] = {1, 1, 1}, Nos[3] = {1, 1, 1};
    int kyl = 0;                                       // kept y modes per rank (SLAB)
    size_t zbytes = 8;                                 // bytes of a non-uniform value
    size_t cbytes = 8;                                 // bytes of a complex coefficient
    bool p2p = false;                                  // SLAB exchanges through peer windows (else NCCL send / recv)
    unsigned busy = 0;                                 // windows that local work has touched since the last barrier (same on every rank)
    std::vector<int64_t> cap_slab_all, cap_user_all;   // every rank's buffer capacities (same rule everywhere: who reallocates is known to all)
};

// ---- kernels -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mg_dest(float z, int convention, int Nz, int nzs, int G)
{
    float r;
    const float f = fold_point<float>(z, convention);
    int d = point_to_cell0<float>(f, Nz, r) / nzs;
    return d < G ? d : G - 1;
}

constexpr int MG_ITEMS = 4;                     // points per thread: the loads are issued before the cell arithmetic
constexpr int MG_TILE = 256 * MG_ITEMS;

__global__ void __launch_bounds__(256) mg_dest_count_kernel(const float *__restrict__ z, int64_t np, int convention, int Nz, int nzs, int G,
                                                            unsigned long long *__restrict__ counts)
{
    __shared__ unsigned int sc[MG_MAX_RANKS];
    if (threadIdx.x < MG_MAX_RANKS) sc[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * MG_TILE;
    float zi[MG_ITEMS];
#pragma unroll
    for (int k = 0; k < MG_ITEMS; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        zi[k] = i < np ? z[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < MG_ITEMS; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        const int d = i < np ? mg_dest(zi[k], convention, Nz, nzs, G) : MG_MAX_RANKS + (threadIdx.x & 31);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (i < np && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sc[d], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < G && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
}

// points grouped by destination rank (any order inside a group): sendperm[pos] = i, coordinates to pos.  A CTA ranks its tile of
// points per destination, reserves one range per destination with a single global atomic, sorts the tile by destination in shared
// memory and writes runs of consecutive positions.  (Writing the runs straight into the destination ranks' windows, and the same
// fusion for the value exchanges, was measured SLOWER than copy-engine copies: 4- and 8-byte remote accesses reach ~120 GB/s
// against ~330 GB/s — profiles/r2_mgpu_fused_put_get_experiment.txt.)
__global__ void __launch_bounds__(256) mg_dest_scatter_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
                                                              int64_t np, int convention, int Nz, int nzs, int G, DestOffsets base,
                                                              unsigned long long *__restrict__ cursor, int32_t *__restrict__ sendperm,
                                                              float *__restrict__ sx, float *__restrict__ sy, float *__restrict__ sz)
{
    __shared__ unsigned int sc[MG_MAX_RANKS], stoff[MG_MAX_RANKS + 1];
    __shared__ unsigned long long sbase[MG_MAX_RANKS];
    __shared__ float bx[MG_TILE], by[MG_TILE], bz[MG_TILE];
    __shared__ int32_t bp[MG_TILE];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < MG_MAX_RANKS) sc[threadIdx.x] = 0;
    __syncthreads();
    const int64_t tile = (int64_t)blockIdx.x * MG_TILE;
    float zi[MG_ITEMS], xi[MG_ITEMS], yi[MG_ITEMS];
    int dest[MG_ITEMS];
    unsigned rank[MG_ITEMS];
#pragma unroll
    for (int k = 0; k < MG_ITEMS; ++k) {
        const int64_t i = tile + k * 256 + threadIdx.x;
        zi[k] = xi[k] = yi[k] = 0.f;
        if (i < np) { zi[k] = z[i]; xi[k] = x[i]; yi[k] = y[i]; }
    }
#pragma unroll
    for (int k = 0; k < MG_ITEMS; ++k) {
        const int64_t i = tile + k * 256 + threadIdx.x;
        const bool valid = i < np;
        dest[k] = valid ? mg_dest(zi[k], convention, Nz, nzs, G) : MG_MAX_RANKS + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, dest[k]);
        const int leader = __ffs(peers) - 1;
        unsigned r = 0;
        if (valid && lane == leader) r = atomicAdd(&sc[dest[k]], (unsigned)__popc(peers));
        r = __shfl_sync(0xffffffffu, r, leader);
        rank[k] = r + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < G) sbase[threadIdx.x] = base.off[threadIdx.x] + (sc[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (unsigned long long)sc[threadIdx.x]) : 0ull);
    if (threadIdx.x == 32) {
        unsigned acc = 0;
        for (int r = 0; r < G; ++r) { stoff[r] = acc; acc += sc[r]; }
        stoff[G] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MG_ITEMS; ++k) {
        const int64_t i = tile + k * 256 + threadIdx.x;
        if (i < np) {
            const unsigned slot = stoff[dest[k]] + rank[k];
            bx[slot] = xi[k]; by[slot] = yi[k]; bz[slot] = zi[k];
            bp[slot] = (int32_t)i;
        }
    }
    __syncthreads();
    const unsigned ntile = stoff[G];
    for (unsigned j = threadIdx.x; j < ntile; j += 256) {
        int r = 0;
        for (int t = 1; t < G; ++t) r += (j >= stoff[t]);
        const unsigned long long pos = sbase[r] + (j - stoff[r]);
        sendperm[pos] = bp[j];
        sx[pos] = bx[j]; sy[pos] = by[j]; sz[pos] = bz[j];
    }
}

// every block copy of an exchange in ONE launch (blockIdx.y = job), written into the peers' windows over NVLink: all peers are
// served at once (measured 670 GB/s of egress per GPU at 8 GPUs, against ~350 GB/s for one copy-engine copy per peer in turn).
// Units are 4-byte words at any 4-byte alignment: the REMOTE side is always written with aligned 16-byte stores (4- and 8-byte
// remote accesses reach ~120 GB/s), the local side is read with 16-byte loads when it happens to share the alignment, else with
// four scalar loads; the <= 3 words before / after the aligned body are stored one by one.
constexpr int MG_MAX_JOBS = 3 * MG_MAX_RANKS;
struct CopyJobs {
    const uint32_t *src[MG_MAX_JOBS];
    uint32_t *dst[MG_MAX_JOBS];
    unsigned long long n[MG_MAX_JOBS];
};
__global__ void __launch_bounds__(256) mg_multi_copy_kernel(CopyJobs j)
{
    const int job = blockIdx.y;
    const unsigned long long n = j.n[job];
    const uint32_t *__restrict__ s = j.src[job];
    uint32_t *__restrict__ d = j.dst[job];
    unsigned long long head = ((16u - (unsigned)((uintptr_t)d & 15u)) & 15u) >> 2;
    if (head > n) head = n;
    const unsigned long long nv = (n - head) >> 2;
    const uint32_t *__restrict__ sb = s + head;
    uint4 *__restrict__ dv = reinterpret_cast<uint4 *>(d + head);
    const unsigned long long stride = (unsigned long long)gridDim.x * 256;
    unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if ((((uintptr_t)sb) & 15u) == 0) {
        const uint4 *__restrict__ sv = reinterpret_cast<const uint4 *>(sb);
        for (; i + 3 * stride < nv; i += 4 * stride) {
            const uint4 a = sv[i], b = sv[i + stride], c = sv[i + 2 * stride], e = sv[i + 3 * stride];
            dv[i] = a; dv[i + stride] = b; dv[i + 2 * stride] = c; dv[i + 3 * stride] = e;
        }
        for (; i < nv; i += stride) dv[i] = sv[i];
    } else {
        for (; i + stride < nv; i += 2 * stride) {
            const uint32_t *p0 = sb + 4 * i, *p1 = sb + 4 * (i + stride);
            const uint4 a = make_uint4(p0[0], p0[1], p0[2], p0[3]), b = make_uint4(p1[0], p1[1], p1[2], p1[3]);
            dv[i] = a; dv[i + stride] = b;
        }
        for (; i < nv; i += stride) {
            const uint32_t *p0 = sb + 4 * i;
            dv[i] = make_uint4(p0[0], p0[1], p0[2], p0[3]);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        if (threadIdx.x < 4) { if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x]; }
        else {
            const unsigned long long t = head + 4 * nv + (threadIdx.x - 4);
            if (t < n) d[t] = s[t];
        }
    }
}

template <typename V> __global__ void __launch_bounds__(256) mg_gather_kernel(V *__restrict__ dst, const V *__restrict__ src, const int32_t *__restrict__ perm, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[k] = src[perm[k]];
}
template <typename V> __global__ void __launch_bounds__(256) mg_scatter_kernel(V *__restrict__ dst, const V *__restrict__ src, const int32_t *__restrict__ perm, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[perm[k]] = src[k];
}

__global__ void __launch_bounds__(256) mg_add_kernel(float4 *__restrict__ dst, const float4 *__restrict__ src, int64_t n4)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 a = dst[i];
        const float4 b = src[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        dst[i] = a;
    }
}

// tmp [G][Kx][Kyl][Kz] (all-gathered local blocks) -> full [Kx][Ky][Kz]
__global__ void __launch_bounds__(256) mg_unblock_kernel(float2 *__restrict__ full, const float2 *__restrict__ tmp, int Kx, int Kyl, int Ky, int Kz)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)Kx * Ky * Kz;
    if (i >= n) return;
    const int kx = (int)(i % Kx);
    const int ky = (int)((i / Kx) % Ky);
    const int kz = (int)(i / ((int64_t)Kx * Ky));
    const int r = ky / Kyl, kl = ky - r * Kyl;
    full[i] = tmp[(((int64_t)r * Kz + kz) * Kyl + kl) * Kx + kx];
}

// ---- helpers ---------------------------------------------------------------------------------------------------------------
static void mg_free_rank(MgRank &R)
{
    cudaSetDevice(R.dev);
    if (R.stream) cudaStreamSynchronize(R.stream);
    if (R.plan) { host_plan_free(*R.plan); delete R.plan; R.plan = nullptr; }
    auto f = [](auto *&p) { if (p) { cudaFree((void *)p); p = nullptr; } };
    f(R.d_sendperm);
    for (int d = 0; d < 3; ++d) { f(R.d_sx[d]); f(R.d_rx[d]); }
    f(R.d_sv); f(R.d_rv); f(R.d_cnt); f(R.d_halo); f(R.d_gather); f(R.d_msg); f(R.d_bar);
    if (R.h_cnt) { cudaFreeHost(R.h_cnt); R.h_cnt = nullptr; }
    if (R.h_msg) { cudaFreeHost(R.h_msg); R.h_msg = nullptr; }
    if (R.h_flag) { cudaFreeHost(R.h_flag); R.h_flag = nullptr; }
    if (R.ev_ok) { for (auto &e : R.ev) cudaEventDestroy(e); R.ev_ok = false; }
    if (R.comm && nccl_api().CommDestroy) { nccl_api().CommDestroy(R.comm); R.comm = nullptr; }
    if (R.own_stream && R.stream) { cudaStreamDestroy(R.stream); R.stream = nullptr; }
}

static inline void mg_rec(MgRank &R, int i)
{
    if (R.ev_ok) cudaEventRecord(R.ev[i], R.stream);
}

static inline int64_t mg_grown(int64_t need) { return need + need / 8 + 1024; }

static int mg_ensure_user(Mgpu &m, MgRank &R, int64_t np)
{
    if (np <= R.cap_user && R.cap_user > 0) return NUFFT_SUCCESS;          // (a rank without points still owns minimal buffers: they are windows)
    auto f = [](auto *&p) { if (p) { cudaFree((void *)p); p = nullptr; } };
    f(R.d_sendperm); f(R.d_sv);
    for (int d = 0; d < 3; ++d) f(R.d_sx[d]);
    R.cap_user = 0;
    const int64_t cap = mg_grown(np);
    CUDA_TRY(cudaMalloc(&R.d_sendperm, (size_t)cap * sizeof(int32_t)));
    for (int d = 0; d < 3; ++d) CUDA_TRY(cudaMalloc(&R.d_sx[d], (size_t)cap * sizeof(float)));
    CUDA_TRY(cudaMalloc(&R.d_sv, (size_t)cap * m.zbytes));
    R.cap_user = cap;
    return NUFFT_SUCCESS;
}

static int mg_ensure_slab(Mgpu &m, MgRank &R, int64_t np)
{
    if (np <= R.cap_slab && R.cap_slab > 0) return NUFFT_SUCCESS;
    auto f = [](auto *&p) { if (p) { cudaFree((void *)p); p = nullptr; } };
    f(R.d_rv);
    for (int d = 0; d < 3; ++d) f(R.d_rx[d]);
    R.cap_slab = 0;
    const int64_t cap = mg_grown(np);
    for (int d = 0; d < 3; ++d) CUDA_TRY(cudaMalloc(&R.d_rx[d], (size_t)cap * sizeof(float)));
    CUDA_TRY(cudaMalloc(&R.d_rv, (size_t)cap * m.zbytes));
    R.cap_slab = cap;
    return NUFFT_SUCCESS;
}

// ---- peer windows ----------------------------------------------------------------------------------------------------------
static void *mg_win_local(MgRank &R, int w)
{
    switch (w) {
        case W_RX0: case W_RX1: case W_RX2: return R.d_rx[w - W_RX0];
        case W_RV: return R.d_rv;
        case W_SV: return R.d_sv;
        case W_HALO: return R.d_halo;
        case W_GRID: return R.plan ? R.plan->d_us : nullptr;
        case W_A: return R.plan ? R.plan->d_pf_a : nullptr;
    }
    return nullptr;
}

// stream-ordered barrier over all ranks: a rank's later work starts after every rank's earlier work has finished
static int mg_barrier(Mgpu &m)
{
    NcclApi &n = nccl_api();
    NCCL_TRY(n.GroupStart());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        NCCL_TRY(n.AllReduce(R.d_bar, R.d_bar, 1, ncclInt32, ncclSum, R.comm, R.stream));
    }
    NCCL_TRY(n.GroupEnd());
    return NUFFT_SUCCESS;
}

// Around an exchange through peer windows.  BEFORE the copies a barrier is needed only if some rank may still be working on a
// window the copies write (local work since the last barrier: the calls are collective, so `busy` is the same everywhere);
// AFTER them always: the data has arrived once every rank's copies have completed.
static int mg_exchange_begin(Mgpu &m, unsigned windows)
{
    if (m.busy & windows) { NUFFT_TRY(mg_barrier(m)); m.busy = 0; }
    return NUFFT_SUCCESS;
}
static int mg_exchange_end(Mgpu &m)
{
    NUFFT_TRY(mg_barrier(m));
    m.busy = 0;
    return NUFFT_SUCCESS;
}

// the same, and the host waits for it (bounded: a peer that died must not hang this process in a destructor)
static int mg_host_barrier(Mgpu &m, double timeout_s)
{
    NUFFT_TRY(mg_barrier(m));
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        if (timeout_s <= 0) { CUDA_TRY(cudaStreamSynchronize(R.stream)); continue; }
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            const cudaError_t q = cudaStreamQuery(R.stream);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) { cudaGetLastError(); set_error("stream error while waiting for the other ranks"); return NUFFT_ERR_CUDA; }
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
                set_error("timed out waiting for the other ranks");
                return NUFFT_ERR_STATE;
            }
            std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
    }
    return NUFFT_SUCCESS;
}

static void mg_win_close(Mgpu &m, unsigned mask)
{
    for (auto &R : m.L) {
        cudaSetDevice(R.dev);
        for (int w = 0; w < W_N; ++w) {
            if (!((mask >> w) & 1u)) continue;
            for (int r = 0; r < MG_MAX_RANKS; ++r) {
                if (R.win_ipc[w][r] && R.win[w][r]) cudaIpcCloseMemHandle(R.win[w][r]);
                R.win[w][r] = nullptr;
                R.win_ipc[w][r] = false;
            }
        }
    }
    cudaGetLastError();
}

// collective: every rank publishes the buffers of the windows in `mask` and maps its peers' ones; *all_ok = every rank succeeded
// (veto: this rank takes part in the agreement but declines)
static int mg_win_open(Mgpu &m, unsigned mask, bool *all_ok, bool veto = false)
{
    NcclApi &n = nccl_api();
    const int G = m.nranks;
    const bool other_processes = (int)m.L.size() < G;
    *all_ok = false;
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        WinMsg &s = R.h_msg[0];
        memset(&s, 0, sizeof(s));
        s.ok = veto ? 0 : 1;
        for (int w = 0; w < W_N && !veto; ++w) {
            if (!((mask >> w) & 1u)) continue;
            void *q = mg_win_local(R, w);
            s.ptr[w] = (unsigned long long)q;
            if (!q) { s.ok = 0; continue; }
            if (other_processes && cudaIpcGetMemHandle(&s.h[w], q) != cudaSuccess) { cudaGetLastError(); s.ok = 0; }
        }
        CUDA_TRY(cudaMemcpyAsync(R.d_msg, R.h_msg, sizeof(WinMsg), cudaMemcpyHostToDevice, R.stream));
    }
    NCCL_TRY(n.GroupStart());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        NCCL_TRY(n.AllGather(R.d_msg, R.d_msg + 1, sizeof(WinMsg), ncclUint8, R.comm, R.stream));
    }
    NCCL_TRY(n.GroupEnd());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        CUDA_TRY(cudaMemcpyAsync(R.h_msg + 1, R.d_msg + 1, (size_t)G * sizeof(WinMsg), cudaMemcpyDeviceToHost, R.stream));
    }
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        CUDA_TRY(cudaStreamSynchronize(R.stream));
    }
    bool ok = true;
    for (int r = 0; r < G; ++r) if (!m.L[0].h_msg[1 + r].ok) ok = false;      // the same verdict on every rank
    if (ok) {
        for (auto &R : m.L) {
            CUDA_TRY(cudaSetDevice(R.dev));
            for (int w = 0; w < W_N; ++w) {
                if (!((mask >> w) & 1u)) continue;
                for (int r = 0; r < G; ++r) {
                    MgRank *local = nullptr;
                    for (auto &Q : m.L) if (Q.rank == r) local = &Q;
                    if (local) { R.win[w][r] = mg_win_local(*local, w); R.win_ipc[w][r] = false; continue; }
                    void *q = nullptr;
                    if (cudaIpcOpenMemHandle(&q, R.h_msg[1 + r].h[w], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; continue; }
                    R.win[w][r] = q;
                    R.win_ipc[w][r] = true;
                }
            }
        }
    }
    // mapping can fail on one rank only: agree
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        *R.h_flag = ok ? 1 : 0;
        CUDA_TRY(cudaMemcpyAsync(R.d_bar + 1, R.h_flag, sizeof(int32_t), cudaMemcpyHostToDevice, R.stream));
    }
    NCCL_TRY(n.GroupStart());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        NCCL_TRY(n.AllReduce(R.d_bar + 1, R.d_bar + 1, 1, ncclInt32, ncclMin, R.comm, R.stream));
    }
    NCCL_TRY(n.GroupEnd());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        CUDA_TRY(cudaMemcpyAsync(R.h_flag, R.d_bar + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, R.stream));
    }
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        CUDA_TRY(cudaStreamSynchronize(R.stream));
        if (*R.h_flag != 1) ok = false;
    }
    if (!ok) mg_win_close(m, mask);
    *all_ok = ok;
    return NUFFT_SUCCESS;
}

// the block copies of one exchange into the peers' windows, one launch
struct CopyList {
    CopyJobs j{};
    int n = 0;
    void add(void *dst, const void *src, size_t bytes)          // bytes and both addresses are multiples of 4
    {
        if (!bytes) return;
        j.dst[n] = (uint32_t *)dst; j.src[n] = (const uint32_t *)src; j.n[n] = bytes / 4;
        ++n;
    }
};
static int mg_copy_run(MgRank &R, const CopyList &c)
{
    if (!c.n) return NUFFT_SUCCESS;
    unsigned long long nmax = 0;
    for (int i = 0; i < c.n; ++i) nmax = std::max(nmax, c.j.n[i]);
    const int sms = R.plan ? R.plan->num_sms : 148;
    const unsigned gx = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>(cdiv((int64_t)nmax, 4096), (unsigned long long)cdiv(4 * sms, c.n)));
    mg_multi_copy_kernel<<<dim3(gx, (unsigned)c.n), 256, 0, R.stream>>>(c.j);
    NUFFT_COUNT_LAUNCH();
    return NUFFT_SUCCESS;
}

// my groups -> the peers' windows (element offsets doff inside them), staggered so that the jobs of a launch start on different peers
static void mg_put_jobs(Mgpu &m, MgRank &R, CopyList &c, int w, const void *send, const std::vector<int64_t> &scnt, const std::vector<int64_t> &soff,
                        const std::vector<int64_t> &doff, size_t esize)
{
    const int G = m.nranks;
    for (int k = 0; k < G; ++k) {
        const int r = (R.rank + k) % G;
        if (scnt[r] <= 0) continue;
        c.add((char *)R.win[w][r] + (size_t)doff[r] * esize, (const char *)send + (size_t)soff[r] * esize, (size_t)scnt[r] * esize);
    }
}

// all-to-all with per-peer element counts (bytes = count * esize) over NCCL, self part copied on the stream
static int mg_alltoallv(Mgpu &m, MgRank &R, const void *send, const std::vector<int64_t> &scnt, const std::vector<int64_t> &soff, void *recv,
                        const std::vector<int64_t> &rcnt, const std::vector<int64_t> &roff, size_t esize)
{
    NcclApi &n = nccl_api();
    for (int r = 0; r < m.nranks; ++r) {
        const char *sp = (const char *)send + (size_t)soff[r] * esize;
        char *rp = (char *)recv + (size_t)roff[r] * esize;
        if (r == R.rank) {
            if (scnt[r] > 0) CUDA_TRY(cudaMemcpyAsync(rp, sp, (size_t)scnt[r] * esize, cudaMemcpyDeviceToDevice, R.stream));
            continue;
        }
        if (scnt[r] > 0) NCCL_TRY(n.Send(sp, (size_t)scnt[r] * esize, ncclUint8, r, R.comm, R.stream));
        if (rcnt[r] > 0) NCCL_TRY(n.Recv(rp, (size_t)rcnt[r] * esize, ncclUint8, r, R.comm, R.stream));
    }
    return NUFFT_SUCCESS;
}

// the entry points switch between the devices of their local ranks: the caller's current device is restored on return
struct MgDeviceRestore {
    int prev = -1;
    MgDeviceRestore() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~MgDeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
};

static int mg_check(nufft_mgpu h)
{
    if (!h) { set_error("null multi-GPU handle"); return NUFFT_ERR_STATE; }
    return NUFFT_SUCCESS;
}

// ---- SLAB strategy ---------------------------------------------------------------------------------------------------------
static int slab_set_points(Mgpu &m, const int64_t np[], const void *const x[])
{
    NcclApi &n = nccl_api();
    const int G = m.nranks, Nz = (int)m.Nos[2], nzs = Nz / G;
    // (1) destination counts of the local points, all-gathered into the G x G matrix (row = sender)
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        CUDA_TRY(cudaSetDevice(R.dev));
        if (np[l] < 0 || np[l] >= ((int64_t)1 << 31) - 4096) { set_error("invalid number of points"); return NUFFT_ERR_ARG; }
        for (int d = 0; d < 3; ++d) if (np[l] > 0 && !x[3 * l + d]) { set_error("null point array"); return NUFFT_ERR_ARG; }
        mg_rec(R, 0);
        CUDA_TRY(cudaMemsetAsync(R.d_cnt, 0, 2 * MG_MAX_RANKS * sizeof(unsigned long long), R.stream));
        if (np[l] > 0) {
            const unsigned grid = (unsigned)cdiv(np[l], MG_TILE);
            mg_dest_count_kernel<<<grid, 256, 0, R.stream>>>((const float *)x[3 * l + 2], np[l], m.opts.point_convention, Nz, nzs, G, R.d_cnt);
            NUFFT_COUNT_LAUNCH();
        }
    }
    NCCL_TRY(n.GroupStart());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        NCCL_TRY(n.AllGather(R.d_cnt, R.d_cnt + 2 * MG_MAX_RANKS, MG_MAX_RANKS, ncclUint64, R.comm, R.stream));
    }
    NCCL_TRY(n.GroupEnd());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        CUDA_TRY(cudaMemcpyAsync(R.h_cnt, R.d_cnt + 2 * MG_MAX_RANKS, (size_t)G * MG_MAX_RANKS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, R.stream));
    }
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        CUDA_TRY(cudaStreamSynchronize(R.stream));
    }
    // (2) counts and offsets of every exchange of this point set
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        R.sendcnt.assign(G, 0); R.sendoff.assign(G, 0); R.recvcnt.assign(G, 0); R.recvoff.assign(G, 0);
        R.put_fwd.assign(G, 0); R.put_bwd.assign(G, 0);
        int64_t so = 0, ro = 0;
        for (int r = 0; r < G; ++r) {
            R.sendcnt[r] = (int64_t)R.h_cnt[(size_t)R.rank * MG_MAX_RANKS + r];
            R.recvcnt[r] = (int64_t)R.h_cnt[(size_t)r * MG_MAX_RANKS + R.rank];
            R.sendoff[r] = so; so += R.sendcnt[r];
            R.recvoff[r] = ro; ro += R.recvcnt[r];
            for (int t = 0; t < R.rank; ++t) {
                R.put_fwd[r] += (int64_t)R.h_cnt[(size_t)t * MG_MAX_RANKS + r];        // senders before me in rank r's receive order
                R.put_bwd[r] += (int64_t)R.h_cnt[(size_t)r * MG_MAX_RANKS + t];        // slabs before mine in rank r's send order
            }
        }
        if (so != np[l]) { set_error("internal error: destination counts do not add up"); return NUFFT_ERR_STATE; }
        if (ro >= ((int64_t)1 << 31) - 4096) { set_error("too many points in one slab"); return NUFFT_ERR_UNSUPPORTED; }
        R.np_user = np[l];
        R.np_slab = ro;
    }
    // buffers: with peer windows a reallocation anywhere is a collective (the old mappings are closed before the owner frees)
    bool remap = false;
    if (m.p2p) {
        const unsigned long long *cm = m.L[0].h_cnt;
        for (int r = 0; r < G; ++r) {
            int64_t row = 0, col = 0;
            for (int t = 0; t < G; ++t) { row += (int64_t)cm[(size_t)r * MG_MAX_RANKS + t]; col += (int64_t)cm[(size_t)t * MG_MAX_RANKS + r]; }
            if (row > m.cap_user_all[r] || m.cap_user_all[r] == 0) { m.cap_user_all[r] = mg_grown(row); remap = true; }
            if (col > m.cap_slab_all[r] || m.cap_slab_all[r] == 0) { m.cap_slab_all[r] = mg_grown(col); remap = true; }
        }
        if (remap) {
            mg_win_close(m, W_DYNAMIC);
            NUFFT_TRY(mg_host_barrier(m, 0));
        }
    }
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        NUFFT_TRY(mg_ensure_user(m, R, R.np_user));
        NUFFT_TRY(mg_ensure_slab(m, R, R.np_slab));
    }
    if (remap) {
        bool ok = false;
        NUFFT_TRY(mg_win_open(m, W_DYNAMIC, &ok));
        if (!ok) { set_error("cannot map the peers' exchange buffers (CUDA IPC)"); return NUFFT_ERR_CUDA; }
    }
    // (3) group the local points by destination, exchange the coordinates
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        CUDA_TRY(cudaSetDevice(R.dev));
        DestOffsets base{};
        for (int r = 0; r < G; ++r) base.off[r] = (unsigned long long)R.sendoff[r];
        if (np[l] > 0) {
            const unsigned grid = (unsigned)cdiv(np[l], MG_TILE);
            mg_dest_scatter_kernel<<<grid, 256, 0, R.stream>>>((const float *)x[3 * l], (const float *)x[3 * l + 1], (const float *)x[3 * l + 2], np[l],
                                                              m.opts.point_convention, Nz, nzs, G, base, R.d_cnt + MG_MAX_RANKS, R.d_sendperm,
                                                              R.d_sx[0], R.d_sx[1], R.d_sx[2]);
            NUFFT_COUNT_LAUNCH();
        }
    }
    if (m.p2p) {
        NUFFT_TRY(mg_exchange_begin(m, 7u << W_RX0));
        for (auto &R : m.L) {
            CUDA_TRY(cudaSetDevice(R.dev));
            CopyList c;
            for (int d = 0; d < 3; ++d) mg_put_jobs(m, R, c, W_RX0 + d, R.d_sx[d], R.sendcnt, R.sendoff, R.put_fwd, sizeof(float));
            NUFFT_TRY(mg_copy_run(R, c));
        }
        NUFFT_TRY(mg_exchange_end(m));
    } else {
        NCCL_TRY(n.GroupStart());
        for (auto &R : m.L) {
            CUDA_TRY(cudaSetDevice(R.dev));
            for (int d = 0; d < 3; ++d)
                NUFFT_TRY(mg_alltoallv(m, R, R.d_sx[d], R.sendcnt, R.sendoff, R.d_rx[d], R.recvcnt, R.recvoff, sizeof(float)));
        }
        NCCL_TRY(n.GroupEnd());
    }
    // (4) bin the slab's points
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 1);
        const void *rx[3] = {R.d_rx[0], R.d_rx[1], R.d_rx[2]};
        NUFFT_TRY(binning_set_points(*R.plan, R.np_slab, rx));
        mg_rec(R, 2);
    }
    m.busy |= 7u << W_RX0;
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

// values of the local points -> values of the slab's points (forward) or back
static int slab_exchange_values(Mgpu &m, bool forward, const void *const vp_in[], void *const vp_out[])
{
    NcclApi &n = nccl_api();
    if (forward) {
        for (size_t l = 0; l < m.L.size(); ++l) {
            MgRank &R = m.L[l];
            CUDA_TRY(cudaSetDevice(R.dev));
            if (R.np_user > 0) {
                mg_gather_kernel<unsigned long long><<<(unsigned)cdiv(R.np_user, 256), 256, 0, R.stream>>>(
                    (unsigned long long *)R.d_sv, (const unsigned long long *)vp_in[l], R.d_sendperm, R.np_user);
                NUFFT_COUNT_LAUNCH();
            }
        }
    }
    if (forward) m.busy |= 1u << W_SV;                 // the gather kernel writes d_sv
    if (m.p2p) {
        NUFFT_TRY(mg_exchange_begin(m, forward ? 1u << W_RV : 1u << W_SV));
        for (auto &R : m.L) {
            CUDA_TRY(cudaSetDevice(R.dev));
            CopyList c;
            if (forward) mg_put_jobs(m, R, c, W_RV, R.d_sv, R.sendcnt, R.sendoff, R.put_fwd, m.zbytes);
            else mg_put_jobs(m, R, c, W_SV, R.d_rv, R.recvcnt, R.recvoff, R.put_bwd, m.zbytes);
            NUFFT_TRY(mg_copy_run(R, c));
        }
        NUFFT_TRY(mg_exchange_end(m));
    } else {
        NCCL_TRY(n.GroupStart());
        for (auto &R : m.L) {
            CUDA_TRY(cudaSetDevice(R.dev));
            if (forward) NUFFT_TRY(mg_alltoallv(m, R, R.d_sv, R.sendcnt, R.sendoff, R.d_rv, R.recvcnt, R.recvoff, m.zbytes));
            else NUFFT_TRY(mg_alltoallv(m, R, R.d_rv, R.recvcnt, R.recvoff, R.d_sv, R.sendcnt, R.sendoff, m.zbytes));
        }
        NCCL_TRY(n.GroupEnd());
    }
    if (!forward) {
        for (size_t l = 0; l < m.L.size(); ++l) {
            MgRank &R = m.L[l];
            CUDA_TRY(cudaSetDevice(R.dev));
            if (R.np_user > 0) {
                mg_scatter_kernel<unsigned long long><<<(unsigned)cdiv(R.np_user, 256), 256, 0, R.stream>>>(
                    (unsigned long long *)vp_out[l], (const unsigned long long *)R.d_sv, R.d_sendperm, R.np_user);
                NUFFT_COUNT_LAUNCH();
            }
        }
        m.busy |= 1u << W_SV;                           // the scatter kernel reads d_sv
    }
    return NUFFT_SUCCESS;
}

// z slabs of [Kx][Ky][nz] blocked by destination  <->  y slabs [Kx][Kyl][Nz]: contiguous blocks of Kx Kyl nz values both sides
static int slab_transpose(Mgpu &m, bool forward)
{
    NcclApi &n = nccl_api();
    const int G = m.nranks;
    if (m.p2p) {
        NUFFT_TRY(mg_exchange_begin(m, forward ? 1u << W_A : 1u << W_GRID));
        for (auto &R : m.L) {
            CUDA_TRY(cudaSetDevice(R.dev));
            const size_t blk = (size_t)(m.nk[0] * m.kyl * R.nz) * m.cbytes;
            const char *src = (const char *)(forward ? R.plan->d_us : R.plan->d_pf_a);
            CopyList c;
            for (int k = 0; k < G; ++k) {
                const int r = (R.rank + k) % G;
                c.add((char *)R.win[forward ? W_A : W_GRID][r] + (size_t)R.rank * blk, src + (size_t)r * blk, blk);
            }
            NUFFT_TRY(mg_copy_run(R, c));
        }
        NUFFT_TRY(mg_exchange_end(m));
        return NUFFT_SUCCESS;
    }
    NCCL_TRY(n.GroupStart());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        const int64_t blk = m.nk[0] * m.kyl * R.nz;
        std::vector<int64_t> cnt(G, blk), off(G);
        for (int r = 0; r < G; ++r) off[r] = (int64_t)r * blk;
        void *grid = R.plan->d_us, *A = R.plan->d_pf_a;
        if (forward) NUFFT_TRY(mg_alltoallv(m, R, grid, cnt, off, A, cnt, off, m.cbytes));
        else NUFFT_TRY(mg_alltoallv(m, R, A, cnt, off, grid, cnt, off, m.cbytes));
    }
    NCCL_TRY(n.GroupEnd());
    return NUFFT_SUCCESS;
}

static int slab_exec_type1(Mgpu &m, void *const uhat[], const void *const vp[])
{
    NcclApi &n = nccl_api();
    const int G = m.nranks, M = m.opts.half_support;
    const int64_t pl = m.Nos[0] * m.Nos[1];              // cells per plane
    for (auto &R : m.L) { CUDA_TRY(cudaSetDevice(R.dev)); mg_rec(R, 3); }
    NUFFT_TRY(slab_exchange_values(m, true, vp, nullptr));
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 4);
        Plan &p = *R.plan;
        CUDA_TRY(cudaMemsetAsync(p.d_us, 0, (size_t)p.ncells * m.zbytes, R.stream));
        NUFFT_COUNT_LAUNCH();
        const void *rv[1] = {R.d_rv};
        NUFFT_TRY(spread_run(p, rv, nullptr));
        mg_rec(R, 5);
    }
    // halo planes -> neighbours: the M - 1 planes below the slab belong to rank - 1, the M planes above to rank + 1
    m.busy |= (1u << W_RV) | (1u << W_GRID);            // spreading
    if (m.p2p) {
        NUFFT_TRY(mg_exchange_begin(m, 1u << W_HALO));
        for (auto &R : m.L) {
            CUDA_TRY(cudaSetDevice(R.dev));
            const int dn = (R.rank + G - 1) % G, up = (R.rank + 1) % G;
            const char *grid = (const char *)R.plan->d_us;
            const size_t lo_b = (size_t)(M - 1) * pl * m.zbytes, hi_b = (size_t)M * pl * m.zbytes;
            CopyList c;
            c.add((char *)R.win[W_HALO][dn], grid, lo_b);
            c.add((char *)R.win[W_HALO][up] + lo_b, grid + (size_t)(M - 1 + R.nz) * pl * m.zbytes, hi_b);
            NUFFT_TRY(mg_copy_run(R, c));
        }
        NUFFT_TRY(mg_exchange_end(m));
    }
    if (!m.p2p) NCCL_TRY(n.GroupStart());
    for (auto &R : m.L) {
        if (m.p2p) break;
        CUDA_TRY(cudaSetDevice(R.dev));
        const int dn = (R.rank + G - 1) % G, up = (R.rank + 1) % G;
        char *grid = (char *)R.plan->d_us, *halo = (char *)R.d_halo;
        const size_t lo_b = (size_t)(M - 1) * pl * m.zbytes, hi_b = (size_t)M * pl * m.zbytes;
        NCCL_TRY(n.Send(grid, lo_b, ncclUint8, dn, R.comm, R.stream));
        NCCL_TRY(n.Send(grid + (size_t)(M - 1 + R.nz) * pl * m.zbytes, hi_b, ncclUint8, up, R.comm, R.stream));
        NCCL_TRY(n.Recv(halo, lo_b, ncclUint8, up, R.comm, R.stream));               // the upper neighbour's lower halo: my top planes
        NCCL_TRY(n.Recv(halo + lo_b, hi_b, ncclUint8, dn, R.comm, R.stream));        // the lower neighbour's upper halo: my first planes
    }
    if (!m.p2p) NCCL_TRY(n.GroupEnd());
    double nf = 1.0;
    for (int d = 0; d < 3; ++d) nf *= 2.0 * M_PI / (double)m.Nos[d];
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        Plan &p = *R.plan;
        char *grid = (char *)p.d_us, *halo = (char *)R.d_halo;
        const int64_t lo4 = (int64_t)(M - 1) * pl * (int64_t)m.zbytes / 16, hi4 = (int64_t)M * pl * (int64_t)m.zbytes / 16;
        // owned plane i is stored at M - 1 + i
        mg_add_kernel<<<(unsigned)cdiv(lo4, 256), 256, 0, R.stream>>>((float4 *)(grid + (size_t)R.nz * pl * m.zbytes), (const float4 *)halo, lo4);
        mg_add_kernel<<<(unsigned)cdiv(hi4, 256), 256, 0, R.stream>>>((float4 *)(grid + (size_t)(M - 1) * pl * m.zbytes),
                                                                     (const float4 *)(halo + (size_t)(M - 1) * pl * m.zbytes), hi4);
        NUFFT_COUNT_LAUNCH(); NUFFT_COUNT_LAUNCH();
        mg_rec(R, 6);
        // pruned passes along x and y on the owned planes; the y pass leaves blocks of kyl kept modes per destination rank
        void *own = grid + (size_t)(M - 1) * pl * m.zbytes;
        NUFFT_TRY(pfft_pass(p, 0, true, own, p.d_pf_a, 1, m.Nos[1] * R.nz, 1.0, 0, 0));
        NUFFT_TRY(pfft_pass(p, 1, true, p.d_pf_a, p.d_us, m.nk[0], R.nz, 1.0, m.kyl, m.nk[0] * m.kyl * R.nz));
        mg_rec(R, 7);
    }
    m.busy |= (1u << W_HALO) | (1u << W_A) | (1u << W_GRID);
    NUFFT_TRY(slab_transpose(m, true));
    m.busy |= 1u << W_A;                                // the z pass reads the transposed intermediate
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 8);
        NUFFT_TRY(pfft_pass(*R.plan, 2, true, R.plan->d_pf_a, uhat[l], m.nk[0] * m.kyl, 1, nf, 0, 0));
        mg_rec(R, 9);
    }
    return NUFFT_SUCCESS;
}

static int slab_exec_type2(Mgpu &m, void *const vp[], const void *const uhat[])
{
    NcclApi &n = nccl_api();
    const int G = m.nranks, M = m.opts.half_support;
    const int64_t pl = m.Nos[0] * m.Nos[1];
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 10);
        NUFFT_TRY(pfft_pass(*R.plan, 2, false, uhat[l], R.plan->d_pf_a, m.nk[0] * m.kyl, 1, 1.0, 0, 0));
        mg_rec(R, 11);
    }
    m.busy |= 1u << W_A;
    NUFFT_TRY(slab_transpose(m, false));
    m.busy |= (1u << W_GRID) | (1u << W_A);             // the y and x passes
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 12);
        Plan &p = *R.plan;
        void *own = (char *)p.d_us + (size_t)(M - 1) * pl * m.zbytes;
        // the blocked intermediate sits at the start of the grid memory; the x pass writes the owned planes (M - 1 planes
        // further up), so the y pass goes through the scratch array
        NUFFT_TRY(pfft_pass(p, 1, false, p.d_us, p.d_pf_a, m.nk[0], R.nz, 1.0, m.kyl, m.nk[0] * m.kyl * R.nz));
        NUFFT_TRY(pfft_pass(p, 0, false, p.d_pf_a, own, 1, m.Nos[1] * R.nz, 1.0, 0, 0));
        mg_rec(R, 13);
    }
    // halo planes <- neighbours: my lower halo = the last M - 1 owned planes of rank - 1, my upper halo = the first M of rank + 1
    if (m.p2p) {
        NUFFT_TRY(mg_exchange_begin(m, 1u << W_GRID));
        for (auto &R : m.L) {
            CUDA_TRY(cudaSetDevice(R.dev));
            const int dn = (R.rank + G - 1) % G, up = (R.rank + 1) % G;
            const char *grid = (const char *)R.plan->d_us;
            const size_t lo_b = (size_t)(M - 1) * pl * m.zbytes, hi_b = (size_t)M * pl * m.zbytes;
            CopyList c;
            c.add((char *)R.win[W_GRID][up], grid + (size_t)R.nz * pl * m.zbytes, lo_b);
            c.add((char *)R.win[W_GRID][dn] + (size_t)(M - 1 + R.nz) * pl * m.zbytes, grid + (size_t)(M - 1) * pl * m.zbytes, hi_b);
            NUFFT_TRY(mg_copy_run(R, c));
        }
        NUFFT_TRY(mg_exchange_end(m));
    }
    if (!m.p2p) NCCL_TRY(n.GroupStart());
    for (auto &R : m.L) {
        if (m.p2p) break;
        CUDA_TRY(cudaSetDevice(R.dev));
        const int dn = (R.rank + G - 1) % G, up = (R.rank + 1) % G;
        char *grid = (char *)R.plan->d_us;
        const size_t lo_b = (size_t)(M - 1) * pl * m.zbytes, hi_b = (size_t)M * pl * m.zbytes;
        NCCL_TRY(n.Send(grid + (size_t)R.nz * pl * m.zbytes, lo_b, ncclUint8, up, R.comm, R.stream));          // my top M - 1 owned planes
        NCCL_TRY(n.Send(grid + (size_t)(M - 1) * pl * m.zbytes, hi_b, ncclUint8, dn, R.comm, R.stream));       // my first M owned planes
        NCCL_TRY(n.Recv(grid, lo_b, ncclUint8, dn, R.comm, R.stream));
        NCCL_TRY(n.Recv(grid + (size_t)(M - 1 + R.nz) * pl * m.zbytes, hi_b, ncclUint8, up, R.comm, R.stream));
    }
    if (!m.p2p) NCCL_TRY(n.GroupEnd());
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 14);
        void *rv[1] = {R.d_rv};
        NUFFT_TRY(interp_run(*R.plan, rv, nullptr));
        mg_rec(R, 15);
    }
    m.busy |= (1u << W_GRID) | (1u << W_RV);
    NUFFT_TRY(slab_exchange_values(m, false, nullptr, vp));
    for (auto &R : m.L) { CUDA_TRY(cudaSetDevice(R.dev)); mg_rec(R, 16); }
    return NUFFT_SUCCESS;
}

}  // namespace nufft

using namespace nufft;

extern "C" {

int nufft_mgpu_unique_id(void *id128)
{
    NcclApi &n = nccl_api();
    if (!n.lib) { set_error("NCCL (libnccl.so.2) could not be loaded: multi-GPU transforms are unavailable"); return NUFFT_ERR_UNSUPPORTED; }
    if (!id128) { set_error("null id buffer"); return NUFFT_ERR_ARG; }
    static_assert(sizeof(ncclUniqueId) == NUFFT_MGPU_ID_BYTES, "NCCL unique id size");
    ncclUniqueId id;
    NCCL_TRY(n.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return NUFFT_SUCCESS;
}

int nufft_mgpu_create(nufft_mgpu *out, const nufft_opts *opts, int32_t nranks, int32_t nlocal, const int32_t *local_ranks,
                      const int32_t *devices, const void *id128, int32_t strategy)
{
    if (!out || !opts || !local_ranks || !devices) { set_error("null argument"); return NUFFT_ERR_ARG; }
    *out = nullptr;
    if (opts->struct_size != sizeof(nufft_opts)) { set_error("nufft_opts.struct_size does not match this library: ABI mismatch"); return NUFFT_ERR_ARG; }
    if (nranks < 1 || nranks > MG_MAX_RANKS || nlocal < 1 || nlocal > nranks) { set_error("invalid rank counts (1 <= nlocal <= nranks <= %d)", MG_MAX_RANKS); return NUFFT_ERR_ARG; }
    if (strategy < NUFFT_MGPU_AUTO || strategy > NUFFT_MGPU_TRANSFORMS) { set_error("unknown multi-GPU strategy %d", strategy); return NUFFT_ERR_ARG; }
    MgDeviceRestore restore_device;
    NcclApi &n = nccl_api();
    if (nranks > 1 && !n.lib) { set_error("NCCL (libnccl.so.2) could not be loaded: multi-GPU transforms are unavailable"); return NUFFT_ERR_UNSUPPORTED; }
    if (nranks > 1 && !id128) { set_error("null NCCL id"); return NUFFT_ERR_ARG; }
    Mgpu *m = new (std::nothrow) Mgpu();
    if (!m) { set_error("out of host memory"); return NUFFT_ERR_ALLOC; }
    m->opts = *opts;
    m->nranks = nranks;
    m->D = opts->dim;
    m->C = opts->ntransforms;
    m->f64 = opts->dtype == NUFFT_F64;
    m->cplx = opts->is_complex != 0;
    m->zbytes = (m->f64 ? 8 : 4) * (m->cplx ? 2 : 1);
    m->cbytes = (m->f64 ? 8 : 4) * 2;
    auto fail = [&](int rc) {
        for (auto &R : m->L) mg_free_rank(R);
        delete m;
        return rc;
    };

    // sizes of the problem from a throw-away description of the plan (no device memory): reuse host_plan_init on rank 0's
    // device would allocate the full grid, so the few numbers needed are recomputed by the slab rule below after the first
    // plan exists; AUTO picks SLAB when the plan is eligible
    int strat = strategy;
    const bool slab_ok = nranks > 1 && opts->dim == 3 && opts->is_complex && opts->dtype == NUFFT_F32 && opts->half_support == 4 &&
                         opts->ntransforms == 1 && opts->gpu_method != NUFFT_METHOD_GLOBAL_MEMORY && !opts->fftshift;
    if (strat == NUFFT_MGPU_AUTO) strat = slab_ok ? NUFFT_MGPU_SLAB : (opts->ntransforms >= nranks ? NUFFT_MGPU_TRANSFORMS : NUFFT_MGPU_POINTS);
    if (strat == NUFFT_MGPU_SLAB && !slab_ok) {
        set_error("the z-slab strategy needs a 3-D ComplexF32 HalfSupport(4) plan with ntransforms = 1 and more than one rank");
        return fail(NUFFT_ERR_UNSUPPORTED);
    }
    m->strategy = strat;

    m->L.resize(nlocal);
    for (int l = 0; l < nlocal; ++l) {
        MgRank &R = m->L[l];
        R.rank = local_ranks[l];
        R.dev = devices[l];
        if (R.rank < 0 || R.rank >= nranks) { set_error("local rank out of range"); return fail(NUFFT_ERR_ARG); }
        if (cudaSetDevice(R.dev) != cudaSuccess) { cudaGetLastError(); set_error("cannot select device %d", R.dev); return fail(NUFFT_ERR_CUDA); }
        // one rank per process: the caller's stream, INCLUDING the default stream (opts.stream == NULL) — the work must be ordered
        // with the caller's copies and events; several ranks in one process: a private stream per rank
        if (nlocal == 1) R.stream = (cudaStream_t)opts->stream;
        else {
            if (cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); set_error("cannot create a stream"); return fail(NUFFT_ERR_CUDA); }
            R.own_stream = true;
        }
    }
    if (nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        if (n.GroupStart() != ncclSuccess) { set_error("ncclGroupStart failed"); return fail(NUFFT_ERR_CUDA); }
        for (auto &R : m->L) {
            cudaSetDevice(R.dev);
            const ncclResult_t r = n.CommInitRank(&R.comm, nranks, id, R.rank);
            if (r != ncclSuccess) { n.GroupEnd(); set_error("ncclCommInitRank failed: %s", n.GetErrorString(r)); return fail(NUFFT_ERR_CUDA); }
        }
        const ncclResult_t r = n.GroupEnd();
        if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", n.GetErrorString(r)); return fail(NUFFT_ERR_CUDA); }
    }

    for (auto &R : m->L) {
        cudaSetDevice(R.dev);
        Plan *p = new (std::nothrow) Plan();
        if (!p) { set_error("out of host memory"); return fail(NUFFT_ERR_ALLOC); }
        R.plan = p;
        p->opts = *opts;
        p->opts.device = R.dev;
        p->opts.stream = R.stream;
        if (strat == NUFFT_MGPU_TRANSFORMS) {
            for (int c = 0; c < opts->ntransforms; ++c) if (c % nranks == R.rank) R.comps.push_back(c);
            if (R.comps.empty()) { delete p; R.plan = nullptr; continue; }
            p->opts.ntransforms = (int)R.comps.size();
        }
        if (strat == NUFFT_MGPU_SLAB) {
            // oversampled z size by the plan's own rule (sigma in T first, src/plan.jl:575-576; power of two required below)
            const int64_t Nz = (int64_t)std::floor((double)((float)opts->sigma * (float)opts->n_modes[2]));
            int64_t nzo = 1;
            while (nzo < Nz) nzo *= 2;
            if (nzo % nranks != 0 || nzo / nranks < opts->half_support || opts->n_modes[1] % nranks != 0) {
                set_error("z-slab strategy: the oversampled z size (%lld) and the kept y size (%lld) must be divisible by the number of ranks, "
                          "with at least M planes per rank", (long long)nzo, (long long)opts->n_modes[1]);
                return fail(NUFFT_ERR_UNSUPPORTED);
            }
            R.nz = (int)(nzo / nranks);
            R.z0 = R.rank * R.nz;
            p->slab_z0 = R.z0;
            p->slab_nz = R.nz;
        }
        const int rc = host_plan_init(*p);
        if (rc != NUFFT_SUCCESS) return fail(rc);
        if (strat == NUFFT_MGPU_SLAB && (p->Nos[2] != (int64_t)R.nz * nranks)) {
            set_error("z-slab strategy: oversampled z size is not a power of two");
            return fail(NUFFT_ERR_UNSUPPORTED);
        }
        for (int d = 0; d < 3; ++d) { m->nk[d] = p->nk[d]; m->Nos[d] = p->Nos[d]; }
        if (strat == NUFFT_MGPU_SLAB) {
            m->kyl = (int)(p->nk[1] / nranks);
            if (cudaMalloc(&R.d_cnt, (size_t)(2 + MG_MAX_RANKS) * MG_MAX_RANKS * sizeof(unsigned long long)) != cudaSuccess ||
                cudaMallocHost(&R.h_cnt, (size_t)MG_MAX_RANKS * MG_MAX_RANKS * sizeof(unsigned long long)) != cudaSuccess ||
                cudaMalloc(&R.d_halo, (size_t)(2 * p->M - 1) * p->Nos[0] * p->Nos[1] * m->zbytes) != cudaSuccess ||
                cudaMalloc(&R.d_msg, (size_t)(1 + MG_MAX_RANKS) * sizeof(WinMsg)) != cudaSuccess ||
                cudaMallocHost(&R.h_msg, (size_t)(1 + MG_MAX_RANKS) * sizeof(WinMsg)) != cudaSuccess ||
                cudaMalloc(&R.d_bar, 4 * sizeof(int32_t)) != cudaSuccess || cudaMemset(R.d_bar, 0, 4 * sizeof(int32_t)) != cudaSuccess ||
                cudaMallocHost(&R.h_flag, 64) != cudaSuccess) {
                cudaGetLastError();
                set_error("cannot allocate the exchange buffers");
                return fail(NUFFT_ERR_ALLOC);
            }
        }
        if (opts->record_timings) {
            for (auto &e : R.ev) if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); set_error("cannot create events"); return fail(NUFFT_ERR_CUDA); }
            R.ev_ok = true;
        }
    }
    // z-slab exchanges through peer windows (NUFFT_B200_MGPU_P2P=0: NCCL send / recv): every GPU must reach every other one
    if (strat == NUFFT_MGPU_SLAB) {
        const char *e = getenv("NUFFT_B200_MGPU_P2P");
        bool want = !(e && atoi(e) == 0);
        if (want && nlocal > 1) {
            for (auto &R : m->L) {
                cudaSetDevice(R.dev);
                for (auto &Q : m->L) {
                    if (Q.dev == R.dev) continue;
                    int can = 0;
                    if (cudaDeviceCanAccessPeer(&can, R.dev, Q.dev) != cudaSuccess || !can) { cudaGetLastError(); want = false; continue; }
                    const cudaError_t pe = cudaDeviceEnablePeerAccess(Q.dev, 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) want = false;
                    cudaGetLastError();
                }
            }
        }
        // a collective decision: mg_win_open agrees on it
        m->cap_slab_all.assign(nranks, 0);
        m->cap_user_all.assign(nranks, 0);
        bool ok = false;
        const int rc = mg_win_open(*m, W_STATIC, &ok, !want);
        if (rc != NUFFT_SUCCESS) return fail(rc);
        m->p2p = ok;
    }
    *out = reinterpret_cast<nufft_mgpu>(m);
    return NUFFT_SUCCESS;
}

int nufft_mgpu_destroy(nufft_mgpu h)
{
    MgDeviceRestore restore_device;
    if (!h) return NUFFT_SUCCESS;
    Mgpu *m = reinterpret_cast<Mgpu *>(h);
    if (m->p2p) {
        // the peers' mappings of my buffers must be closed before I free them: close mine, then wait (bounded) for everybody
        for (auto &R : m->L) { cudaSetDevice(R.dev); cudaStreamSynchronize(R.stream); }
        mg_win_close(*m, W_DYNAMIC | W_STATIC);
        mg_host_barrier(*m, 5.0);
        cudaGetLastError();
    }
    for (auto &R : m->L) mg_free_rank(R);
    delete m;
    return NUFFT_SUCCESS;
}

int nufft_mgpu_info(nufft_mgpu h, int32_t *strategy, int32_t *nranks, int32_t *nlocal, int64_t size_out[3], int64_t os_dims[3])
{
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    if (strategy) *strategy = m.strategy;
    if (nranks) *nranks = m.nranks;
    if (nlocal) *nlocal = (int32_t)m.L.size();
    for (int d = 0; d < 3; ++d) {
        if (size_out) size_out[d] = m.nk[d];
        if (os_dims) os_dims[d] = m.Nos[d];
    }
    return NUFFT_SUCCESS;
}

int nufft_mgpu_local_block(nufft_mgpu h, int32_t l, int64_t offset[3], int64_t size[3])
{
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    if (l < 0 || l >= (int)m.L.size()) { set_error("local rank index out of range"); return NUFFT_ERR_ARG; }
    for (int d = 0; d < 3; ++d) { if (offset) offset[d] = 0; if (size) size[d] = m.nk[d]; }
    if (m.strategy == NUFFT_MGPU_SLAB) {
        if (offset) offset[1] = (int64_t)m.L[l].rank * m.kyl;
        if (size) size[1] = m.kyl;
    }
    return NUFFT_SUCCESS;
}

int nufft_mgpu_set_points(nufft_mgpu h, const int64_t np[], const void *const x[])
{
    MgDeviceRestore restore_device;
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    if (!np || !x) { set_error("null argument"); return NUFFT_ERR_ARG; }
    if (m.strategy == NUFFT_MGPU_SLAB) return slab_set_points(m, np, x);
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        if (!R.plan) continue;
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 0);
        mg_rec(R, 1);
        NUFFT_TRY(binning_set_points(*R.plan, np[l], x + 3 * l));
        mg_rec(R, 2);
        R.np_user = np[l];
    }
    return NUFFT_SUCCESS;
}

int nufft_mgpu_exec_type1(nufft_mgpu h, void *const uhat[], const void *const vp[], const nufft_callbacks *cb)
{
    MgDeviceRestore restore_device;
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    if (!uhat || !vp) { set_error("null argument"); return NUFFT_ERR_ARG; }
    for (auto &R : m.L) if (R.plan && R.np_user < 0) { set_error("set_points must be called before exec_type1"); return NUFFT_ERR_STATE; }
    if (m.strategy == NUFFT_MGPU_SLAB) {
        if (cb) { set_error("callbacks are not supported by the z-slab strategy"); return NUFFT_ERR_UNSUPPORTED; }
        return slab_exec_type1(m, uhat, vp);
    }
    NcclApi &n = nccl_api();
    const int C = m.C;
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        if (!R.plan) continue;
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 3);
        std::vector<void *> u;
        std::vector<const void *> v;
        if (m.strategy == NUFFT_MGPU_TRANSFORMS) for (int c : R.comps) { u.push_back(uhat[l * C + c]); v.push_back(vp[l * C + c]); }
        else for (int c = 0; c < C; ++c) { u.push_back(uhat[l * C + c]); v.push_back(vp[l * C + c]); }
        NUFFT_TRY(nufft_exec_type1(reinterpret_cast<nufft_plan>(R.plan), u.data(), v.data(), cb));
        mg_rec(R, 8);
    }
    if (m.strategy == NUFFT_MGPU_POINTS && m.nranks > 1) {
        const size_t cnt = (size_t)(m.nk[0] * m.nk[1] * m.nk[2]) * 2;
        NCCL_TRY(n.GroupStart());
        for (size_t l = 0; l < m.L.size(); ++l) {
            MgRank &R = m.L[l];
            CUDA_TRY(cudaSetDevice(R.dev));
            for (int c = 0; c < C; ++c)
                NCCL_TRY(n.AllReduce(uhat[l * C + c], uhat[l * C + c], cnt, m.f64 ? ncclDouble : ncclFloat, ncclSum, R.comm, R.stream));
        }
        NCCL_TRY(n.GroupEnd());
    }
    for (auto &R : m.L) if (R.plan) { CUDA_TRY(cudaSetDevice(R.dev)); mg_rec(R, 9); }
    return NUFFT_SUCCESS;
}

int nufft_mgpu_exec_type2(nufft_mgpu h, void *const vp[], const void *const uhat[], const nufft_callbacks *cb)
{
    MgDeviceRestore restore_device;
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    if (!uhat || !vp) { set_error("null argument"); return NUFFT_ERR_ARG; }
    for (auto &R : m.L) if (R.plan && R.np_user < 0) { set_error("set_points must be called before exec_type2"); return NUFFT_ERR_STATE; }
    if (m.strategy == NUFFT_MGPU_SLAB) {
        if (cb) { set_error("callbacks are not supported by the z-slab strategy"); return NUFFT_ERR_UNSUPPORTED; }
        return slab_exec_type2(m, vp, uhat);
    }
    NcclApi &n = nccl_api();
    const int C = m.C;
    for (auto &R : m.L) if (R.plan) { CUDA_TRY(cudaSetDevice(R.dev)); mg_rec(R, 10); }
    if (m.strategy == NUFFT_MGPU_POINTS && m.nranks > 1) {       // the spectrum of rank 0 goes to everybody
        const size_t cnt = (size_t)(m.nk[0] * m.nk[1] * m.nk[2]) * 2;
        NCCL_TRY(n.GroupStart());
        for (size_t l = 0; l < m.L.size(); ++l) {
            MgRank &R = m.L[l];
            CUDA_TRY(cudaSetDevice(R.dev));
            for (int c = 0; c < C; ++c)
                NCCL_TRY(n.Broadcast(uhat[l * C + c], (void *)uhat[l * C + c], cnt, m.f64 ? ncclDouble : ncclFloat, 0, R.comm, R.stream));
        }
        NCCL_TRY(n.GroupEnd());
    }
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        if (!R.plan) continue;
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_rec(R, 11);
        std::vector<void *> v;
        std::vector<const void *> u;
        if (m.strategy == NUFFT_MGPU_TRANSFORMS) for (int c : R.comps) { u.push_back(uhat[l * C + c]); v.push_back(vp[l * C + c]); }
        else for (int c = 0; c < C; ++c) { u.push_back(uhat[l * C + c]); v.push_back(vp[l * C + c]); }
        NUFFT_TRY(nufft_exec_type2(reinterpret_cast<nufft_plan>(R.plan), v.data(), u.data(), cb));
        mg_rec(R, 16);
    }
    return NUFFT_SUCCESS;
}

int nufft_mgpu_gather_output(nufft_mgpu h, void *const full[], const void *const local[])
{
    MgDeviceRestore restore_device;
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    if (!full || !local) { set_error("null argument"); return NUFFT_ERR_ARG; }
    const int64_t K = m.nk[0] * m.nk[1] * m.nk[2];
    if (m.strategy != NUFFT_MGPU_SLAB) {
        for (size_t l = 0; l < m.L.size(); ++l) {
            MgRank &R = m.L[l];
            CUDA_TRY(cudaSetDevice(R.dev));
            for (int c = 0; c < m.C; ++c)
                if (full[l * m.C + c] != local[l * m.C + c] && local[l * m.C + c])
                    CUDA_TRY(cudaMemcpyAsync(full[l * m.C + c], local[l * m.C + c], (size_t)K * m.cbytes, cudaMemcpyDeviceToDevice, R.stream));
        }
        return NUFFT_SUCCESS;
    }
    NcclApi &n = nccl_api();
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        if (!R.d_gather) CUDA_TRY(cudaMalloc(&R.d_gather, (size_t)K * m.cbytes));
    }
    NCCL_TRY(n.GroupStart());
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        CUDA_TRY(cudaSetDevice(R.dev));
        NCCL_TRY(n.AllGather(local[l], R.d_gather, (size_t)(K / m.nranks) * m.cbytes, ncclUint8, R.comm, R.stream));
    }
    NCCL_TRY(n.GroupEnd());
    for (size_t l = 0; l < m.L.size(); ++l) {
        MgRank &R = m.L[l];
        CUDA_TRY(cudaSetDevice(R.dev));
        mg_unblock_kernel<<<(unsigned)cdiv(K, 256), 256, 0, R.stream>>>((float2 *)full[l], (const float2 *)R.d_gather, (int)m.nk[0], m.kyl, (int)m.nk[1], (int)m.nk[2]);
        NUFFT_COUNT_LAUNCH();
    }
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

int nufft_mgpu_synchronize(nufft_mgpu h)
{
    MgDeviceRestore restore_device;
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    for (auto &R : m.L) {
        CUDA_TRY(cudaSetDevice(R.dev));
        CUDA_TRY(cudaStreamSynchronize(R.stream));
    }
    return NUFFT_SUCCESS;
}

int nufft_mgpu_get_stream(nufft_mgpu h, int32_t l, void **stream)
{
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    if (l < 0 || l >= (int)m.L.size() || !stream) { set_error("invalid argument"); return NUFFT_ERR_ARG; }
    *stream = (void *)m.L[l].stream;
    return NUFFT_SUCCESS;
}

int nufft_mgpu_exchange_mode(nufft_mgpu h, int32_t *mode)
{
    NUFFT_TRY(mg_check(h));
    if (!mode) { set_error("null argument"); return NUFFT_ERR_ARG; }
    *mode = reinterpret_cast<Mgpu *>(h)->p2p ? 1 : 0;
    return NUFFT_SUCCESS;
}

int nufft_mgpu_get_timings(nufft_mgpu h, int32_t l, float ms[16])
{
    MgDeviceRestore restore_device;
    NUFFT_TRY(mg_check(h));
    Mgpu &m = *reinterpret_cast<Mgpu *>(h);
    if (l < 0 || l >= (int)m.L.size() || !ms) { set_error("invalid argument"); return NUFFT_ERR_ARG; }
    MgRank &R = m.L[l];
    if (!R.ev_ok) { set_error("handle was created with record_timings = 0"); return NUFFT_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(R.dev));
    CUDA_TRY(cudaStreamSynchronize(R.stream));
    // [0] point exchange  [1] local set_points  [2] type-1 value exchange  [3] zero fill + spreading  [4] halo exchange + add
    // [5] FFT passes x, y  [6] transpose  [7] FFT pass z   [8] type-2 FFT pass z  [9] transpose  [10] FFT passes y, x
    // [11] halo exchange  [12] interpolation  [13] value return
    const int pairs[14][2] = {{0, 1}, {1, 2}, {3, 4}, {4, 5}, {5, 6}, {6, 7}, {7, 8}, {8, 9}, {10, 11}, {11, 12}, {12, 13}, {13, 14}, {14, 15}, {15, 16}};
    for (int i = 0; i < 14; ++i) {
        float t = 0;
        if (cudaEventElapsedTime(&t, R.ev[pairs[i][0]], R.ev[pairs[i][1]]) == cudaSuccess) R.ms[i] = t;
        else cudaGetLastError();
    }
    for (int i = 0; i < 16; ++i) ms[i] = R.ms[i];
    return NUFFT_SUCCESS;
}

}  // extern "C"
