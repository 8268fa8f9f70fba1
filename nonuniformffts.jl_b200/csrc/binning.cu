// binning.cu — K-bin: set_points! on the GPU.
//
// Replaces src/blocking/gpu.jl:73-212 (assign_blocks_kernel!, AK.accumulate!, sortperm_kernel!,
// permute_kernel!) and src/set_points.jl:33-52.
//
// The reference ranks points inside a bin with an atomic counter, so its intra-bin order is
// non-deterministic.  Here the (bin id -> point index) pairs go through a stable LSD radix sort
// restricted to ceil(log2(nbins)) key bits, which yields exactly the order of the reference's
// single-thread counting sort (src/blocking/cpu.jl:73-111): deterministic and bit-comparable.
//
// Kernels
//   bin_keys_kernel      fold + point_to_cell + block_index -> 32-bit key, bin histogram (warp-aggregated atomics)
//   radix_hist_kernel    per-CTA digit histogram (8-bit digits)
//   scan kernels         exclusive prefix sum (digit-major histogram; bin offsets; work items)
//   radix_scatter_kernel stable scatter (warp match_any ranks + per-warp digit counters)
//   gather_points_kernel sorted, folded copy of the coordinates (coalesced reads in spread/interp)
//   work_items_kernel    items per bin = ceil(count / chunk) (bins over `chunk` points are split)
#include <algorithm>
#include <utility>
#include "common.cuh"
#include "kernel_eval.cuh"

namespace nufft {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
#ifndef SCATTER_MIN_CTAS
#define SCATTER_MIN_CTAS 3      // 80 registers: 3 CTAs (24 warps) per SM; ptxas picks 113 (2 CTAs) on its own, 4 CTAs (64 registers, spills) measured slower
#endif
static_assert(SORT_THREADS == RADIX, "radix_scatter_kernel: one thread per digit");

struct BinGeom {
    int D;
    int N[3];
    int B[3];
    int nb[3];
    int convention;
    int xstride;     // element stride of the point arrays: 1 (D separate vectors) or D (one (D, Np) matrix)
    int rt;          // sub-bin refinement active (TileGeom::rt)
    int sub[3];
    int nsub;
    int zoff;        // z slab plans: global cell of the slab's first owned plane (0 otherwise)
    int nzown;       // z cells this plan bins (N[2] unless a slab)
};

// interleaved record of the folded coordinates of one point (D > 1): float4 / double4
template <typename T> struct PointRec;
template <> struct PointRec<float> {
    using type = float4;
    __device__ static __forceinline__ float4 make(float a, float b, float c) { return make_float4(a, b, c, 0.f); }
};
template <> struct PointRec<double> {
    using type = double4;
    __device__ static __forceinline__ double4 make(double a, double b, double c) { return make_double4(a, b, c, 0.); }
};

// ---- keys --------------------------------------------------------------------------------------
constexpr int KEYS_ITEMS = 4;
constexpr int MAX_PASSES = 4;
// The CTA walks whole tiles of SORT_TILE consecutive points (grid-stride) and, when `ghist` is given, accumulates the digit
// histograms of ALL radix passes of its keys in shared memory and adds them to ghist[pass][digit] once at the end: the
// single-sweep sort passes (onesweep_kernel) then need no histogram pass of their own.
template <typename T, bool HIST>
__global__ void __launch_bounds__(256)
bin_keys_kernel(BinGeom g, int64_t np, const T *__restrict__ x0, const T *__restrict__ x1, const T *__restrict__ x2,
                uint32_t *__restrict__ keys, uint32_t *__restrict__ bin_count, typename PointRec<T>::type *__restrict__ rec,
                uint32_t *__restrict__ ghist, int passes, int nblk)
{
    static_assert(SORT_TILE % (256 * KEYS_ITEMS) == 0 && SORT_THREADS == 256, "tiles of the sort are walked by whole iterations");
    __shared__ uint32_t dh[MAX_PASSES][RADIX];
    // KEYS_ITEMS points per thread and iteration: all coordinate loads are issued before the (division-heavy) cell arithmetic,
    // which keeps enough bytes in flight to cover the HBM latency (one point per thread ran at half the copy bandwidth)
    const int lane = threadIdx.x & 31;
    if (ghist) {
#pragma unroll
        for (int q = 0; q < MAX_PASSES; ++q) dh[q][threadIdx.x] = 0;
        __syncthreads();
    }
    for (int64_t tile = blockIdx.x; tile < nblk; tile += gridDim.x) {
    for (int64_t base = tile * SORT_TILE; base < (tile + 1) * SORT_TILE && base < np; base += 256 * KEYS_ITEMS) {
        T xin[KEYS_ITEMS][3];
#pragma unroll
        for (int it = 0; it < KEYS_ITEMS; ++it) {
            const int64_t i = base + (int64_t)it * blockDim.x + threadIdx.x;
            xin[it][0] = xin[it][1] = xin[it][2] = (T)0;
            if (i < np) {
                xin[it][0] = x0[i * g.xstride];
                if (g.D > 1) xin[it][1] = x1[i * g.xstride];
                if (g.D > 2) xin[it][2] = x2[i * g.xstride];
            }
        }
#pragma unroll
        for (int it = 0; it < KEYS_ITEMS; ++it) {
            const int64_t i = base + (int64_t)it * blockDim.x + threadIdx.x;
            const bool valid = i < np;
            uint32_t key = 0xffffffffu;
            if (valid) {
                T r;
                T f0 = fold_point<T>(xin[it][0], g.convention), f1 = (T)0, f2 = (T)0;
                int c = point_to_cell0<T>(f0, g.N[0], r);
                int b = c / g.B[0];
                uint32_t k = (uint32_t)b;
                int sx = (c - b * g.B[0]) >> 2, sy = 0, sz = 0;   // refined plans: sub-bins of 4 cells in x and y
                if (g.D > 1) {
                    f1 = fold_point<T>(xin[it][1], g.convention);
                    c = point_to_cell0<T>(f1, g.N[1], r);
                    b = c / g.B[1];
                    sy = (c - b * g.B[1]) >> 2;
                    k += (uint32_t)b * (uint32_t)g.nb[0];
                }
                if (g.D > 2) {
                    f2 = fold_point<T>(xin[it][2], g.convention);
                    c = point_to_cell0<T>(f2, g.N[2], r) - g.zoff;  // z slab plans: cells relative to the slab's first plane
                    c = c < 0 ? 0 : (c >= g.nzown ? g.nzown - 1 : c);
                    b = c / g.B[2];
                    sz = c - b * g.B[2];                           // single cells along z
                    k += (uint32_t)b * (uint32_t)(g.nb[0] * g.nb[1]);
                }
                // column-streaming plans: refine by the z cell inside the bin (a column of 4 x 4 cells) so that the points of a
                // column are contiguous and ordered along z; the histogram stays per bin
                key = k;
                const uint32_t kfull = g.rt ? k * (uint32_t)g.nsub + (uint32_t)((sy * g.sub[0] + sx) * g.sub[2] + sz) : k;
                keys[i] = kfull;
                if (ghist) {
                    for (int q = 0; q < passes; ++q) atomicAdd(&dh[q][(kfull >> (q * RADIX_BITS)) & (RADIX - 1)], 1u);
                }
                // folded coordinates as one 16- / 32-byte record: the gather after the sort then touches one sector per point
                if (rec) rec[i] = PointRec<T>::make(f0, f1, f2);
            }
            // warp-aggregated histogram: one atomic per distinct bin in the warp (clustered inputs).  Column-streaming plans
            // (HIST = false) do not need bin offsets on the path: they are built on demand (binning_ensure_offsets)
            if (HIST) {
                const unsigned active = __ballot_sync(0xffffffffu, valid);
                if (valid) {
                    const unsigned peers = __match_any_sync(active, key);
                    if (lane == __ffs(peers) - 1) atomicAdd(&bin_count[key + 1], (uint32_t)__popc(peers));
                }
            }
        }
    }
    }
    if (ghist) {
        __syncthreads();
        for (int q = 0; q < passes; ++q) {
            const uint32_t v = dh[q][threadIdx.x];
            if (v) atomicAdd(&ghist[q * RADIX + threadIdx.x], v);
        }
    }
}

// bin histogram from the folded records (same cell arithmetic as bin_keys_kernel): introspection path of plans whose
// set_points skips it
__global__ void __launch_bounds__(256)
bin_hist_from_rec_kernel(BinGeom g, int64_t np, const float4 *__restrict__ rec, uint32_t *__restrict__ bin_count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const float4 r = rec[i];
    float t;
    const int b0 = point_to_cell0<float>(r.x, g.N[0], t) / g.B[0];
    const int b1 = point_to_cell0<float>(r.y, g.N[1], t) / g.B[1];
    int c2 = point_to_cell0<float>(r.z, g.N[2], t) - g.zoff;
    c2 = c2 < 0 ? 0 : (c2 >= g.nzown ? g.nzown - 1 : c2);
    const int b2 = c2 / g.B[2];
    atomicAdd(&bin_count[(b2 * g.nb[1] + b1) * g.nb[0] + b0 + 1], 1u);
}

// ---- exclusive scan (uint32, in place), three kernels -------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem /*>= 32*/, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;          // exclusive warp offsets
        if (lane == 31) smem[32] = winc; // block total
    }
    __syncthreads();
    const uint32_t res = smem[warp] + inc - v;
    total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ data, int64_t n, uint32_t *__restrict__ sums)
{
    __shared__ uint32_t sm[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += data[i];
    }
    uint32_t total;
    block_exclusive_scan(s, sm, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the block sums (loops over chunks, carries a running prefix)
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(uint32_t *__restrict__ sums, int64_t m)
{
    __shared__ uint32_t sm[33];
    uint32_t carry = 0;
    for (int64_t base = 0; base < m; base += SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = (i < m) ? sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, sm, total);
        if (i < m) sums[i] = ex + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(uint32_t *__restrict__ data, int64_t n, const uint32_t *__restrict__ sums, int inclusive)
{
    __shared__ uint32_t sm[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;   // blocked arrangement
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, sm, total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) data[base + k] = inclusive ? run + v[k] : run;
        run += v[k];
    }
}

int scan_u32(Plan &p, uint32_t *data, int64_t n, bool inclusive)
{
    if (n <= 0) return NUFFT_SUCCESS;
    const int64_t nblk = cdiv(n, SCAN_TILE);
    if ((size_t)nblk > p.scan_tmp_cap) {
        if (p.d_scan_tmp) cudaFree(p.d_scan_tmp);
        p.scan_tmp_cap = (size_t)nblk * 2 + 64;
        CUDA_TRY(cudaMalloc(&p.d_scan_tmp, p.scan_tmp_cap * sizeof(uint32_t)));
    }
    scan_reduce_kernel<<<(unsigned)nblk, SCAN_THREADS, 0, p.stream>>>(data, n, p.d_scan_tmp);
    NUFFT_COUNT_LAUNCH();
    scan_sums_kernel<<<1, SCAN_THREADS, 0, p.stream>>>(p.d_scan_tmp, nblk);
    NUFFT_COUNT_LAUNCH();
    scan_apply_kernel<<<(unsigned)nblk, SCAN_THREADS, 0, p.stream>>>(data, n, p.d_scan_tmp, inclusive ? 1 : 0);
    NUFFT_COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

// ---- radix sort passes -----------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ hist, int nblk)
{
    __shared__ uint32_t h[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) h[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & (RADIX - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) hist[(size_t)i * nblk + blockIdx.x] = h[i];
}

// Stable scatter.  Order inside a CTA tile: warp-major, then iteration k, then lane (== index order).  The tile is first
// reordered by digit in shared memory so that the global writes are runs of consecutive addresses per digit.
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(SORT_THREADS, SCATTER_MIN_CTAS)
radix_scatter_kernel(const uint32_t *__restrict__ keys_in, const int32_t *__restrict__ vals_in, int64_t n, int shift,
                     const uint32_t *__restrict__ hist_scanned, int nblk,
                     uint32_t *__restrict__ keys_out, int32_t *__restrict__ vals_out)
{
    __shared__ uint32_t warp_cnt[SORT_WARPS][RADIX];
    __shared__ uint32_t digit_base[RADIX];       // global position of the first element of (digit, this CTA)
    __shared__ uint32_t local_off[RADIX];        // position of the digit's first element inside the reordered tile
    __shared__ uint32_t scan_tmp[40];
    __shared__ uint32_t s_keys[SORT_TILE];
    __shared__ int32_t s_vals[SORT_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();

    const int64_t tbase = (int64_t)blockIdx.x * SORT_TILE;
    const int64_t wbase = tbase + (int64_t)warp * (32 * SORT_ITEMS);
    const int ntile = (int)(n - tbase < SORT_TILE ? n - tbase : SORT_TILE);
    uint32_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    int32_t val[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        key[k] = i < n ? keys_in[i] : 0u;
        val[k] = FIRST ? (int32_t)i : (i < n ? vals_in[i] : 0);
    }
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        const bool valid = i < n;
        const uint32_t digit = (key[k] >> shift) & (RADIX - 1);
        const uint32_t tag = valid ? digit : (uint32_t)(RADIX + lane);   // invalid lanes never match anyone
        const unsigned peers = __match_any_sync(0xffffffffu, tag);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (lane == leader && valid) {
            pre = warp_cnt[warp][digit];
            warp_cnt[warp][digit] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[k] = pre + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    // per digit (one thread each): exclusive scan over the warps of this CTA, CTA total, global base of (digit, CTA)
    uint32_t total = 0;
    {
        const int d = threadIdx.x;               // SORT_THREADS == RADIX
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t t = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += t;
        }
        total = run;
        digit_base[d] = hist_scanned[(size_t)d * nblk + blockIdx.x];
    }
    {
        uint32_t sum;
        const uint32_t ex = block_exclusive_scan(total, scan_tmp, sum);
        local_off[threadIdx.x] = ex;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t digit = (key[k] >> shift) & (RADIX - 1);
            const uint32_t lpos = local_off[digit] + warp_cnt[warp][digit] + rank[k];
            s_keys[lpos] = key[k];
            s_vals[lpos] = val[k];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ntile; t += SORT_THREADS) {
        const uint32_t kk = s_keys[t];
        const uint32_t digit = (kk >> shift) & (RADIX - 1);
        const uint32_t pos = digit_base[digit] + ((uint32_t)t - local_off[digit]);
        if (!LAST) keys_out[pos] = kk;
        vals_out[pos] = s_vals[t];
    }
}

// ---- single-sweep radix passes (decoupled look-back) ----------------------------------------------------------
// One kernel per digit, no separate histogram / scan launches: bin_keys_kernel has left the global digit histograms of all
// passes (os_scan_kernel turns them into exclusive digit offsets), and a tile learns the number of equal digits in all
// EARLIER tiles by looking back through a status array — per (tile, digit) one 32-bit word holding a count and a flag:
//   00 not ready   01 count of this tile only (aggregate)   10 count of all tiles up to and including this one (inclusive).
// Tiles are numbered by an atomic ticket taken at CTA start, so every tile a CTA waits for is running or finished.
// Stability: tile order == index order, and inside a tile the order of radix_scatter_kernel (warp, iteration, lane).
constexpr uint32_t OS_AGG = 0x40000000u, OS_INC = 0x80000000u, OS_MASK = 0x3fffffffu;

__global__ void __launch_bounds__(RADIX) os_scan_kernel(uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t sm[33];
    uint32_t *h = ghist + (size_t)blockIdx.x * RADIX;
    const uint32_t v = h[threadIdx.x];
    uint32_t total;
    h[threadIdx.x] = block_exclusive_scan(v, sm, total);
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(SORT_THREADS, SCATTER_MIN_CTAS)
onesweep_kernel(const uint32_t *__restrict__ keys_in, const int32_t *__restrict__ vals_in, int64_t n, int shift,
                const uint32_t *__restrict__ gbase, uint32_t *__restrict__ status, uint32_t *__restrict__ ticket,
                uint32_t *__restrict__ keys_out, int32_t *__restrict__ vals_out)
{
    __shared__ uint32_t warp_cnt[SORT_WARPS][RADIX];
    __shared__ uint32_t digit_base[RADIX];       // global position of the first element of (digit, this tile)
    __shared__ uint32_t local_off[RADIX];        // position of the digit's first element inside the reordered tile
    __shared__ uint32_t scan_tmp[40];
    __shared__ uint32_t s_keys[SORT_TILE];
    __shared__ int32_t s_vals[SORT_TILE];
    __shared__ uint32_t s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;

    const int64_t tbase = (int64_t)tile * SORT_TILE;
    const int64_t wbase = tbase + (int64_t)warp * (32 * SORT_ITEMS);
    const int ntile = (int)(n - tbase < SORT_TILE ? n - tbase : SORT_TILE);
    uint32_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    int32_t val[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        key[k] = i < n ? keys_in[i] : 0u;
        val[k] = FIRST ? (int32_t)i : (i < n ? vals_in[i] : 0);
    }
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        const bool valid = i < n;
        const uint32_t digit = (key[k] >> shift) & (RADIX - 1);
        const uint32_t tag = valid ? digit : (uint32_t)(RADIX + lane);   // invalid lanes never match anyone
        const unsigned peers = __match_any_sync(0xffffffffu, tag);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (lane == leader && valid) {
            pre = warp_cnt[warp][digit];
            warp_cnt[warp][digit] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[k] = pre + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    // per digit (one thread each): exclusive scan over the warps of this tile, tile total, look-back for the global base
    uint32_t total = 0;
    {
        const int d = threadIdx.x;               // SORT_THREADS == RADIX
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t t = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += t;
        }
        total = run;
        uint32_t *my = status + (size_t)tile * RADIX + d;
        if (tile == 0) {
            st_volatile_u32(my, total | OS_INC);
            digit_base[d] = gbase[d];
        } else {
            st_volatile_u32(my, total | OS_AGG);
            uint32_t excl = 0;
            for (int64_t t = (int64_t)tile - 1; t >= 0; --t) {
                const uint32_t *ps = status + (size_t)t * RADIX + d;
                uint32_t v;
                do { v = ld_volatile_u32(ps); } while ((v & (OS_AGG | OS_INC)) == 0);
                excl += v & OS_MASK;
                if (v & OS_INC) break;
            }
            st_volatile_u32(my, (excl + total) | OS_INC);
            digit_base[d] = gbase[d] + excl;
        }
    }
    {
        uint32_t sum;
        const uint32_t ex = block_exclusive_scan(total, scan_tmp, sum);
        local_off[threadIdx.x] = ex;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t digit = (key[k] >> shift) & (RADIX - 1);
            const uint32_t lpos = local_off[digit] + warp_cnt[warp][digit] + rank[k];
            s_keys[lpos] = key[k];
            s_vals[lpos] = val[k];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ntile; t += SORT_THREADS) {
        const uint32_t kk = s_keys[t];
        const uint32_t digit = (kk >> shift) & (RADIX - 1);
        const uint32_t pos = digit_base[digit] + ((uint32_t)t - local_off[digit]);
        if (!LAST) keys_out[pos] = kk;
        vals_out[pos] = s_vals[t];
    }
}

// ---- sorted, folded copy of the points ---------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_points_kernel(int D, int convention, int xstride, int64_t np, const int32_t *__restrict__ perm,
                     const T *__restrict__ x0, const typename PointRec<T>::type *__restrict__ rec,
                     T *__restrict__ y0, T *__restrict__ y1, T *__restrict__ y2)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np) return;
    const int32_t i = perm[k];
    if (D == 1) { y0[k] = fold_point<T>(x0[(int64_t)i * xstride], convention); return; }
    const typename PointRec<T>::type r = rec[i];          // folded in bin_keys_kernel: one sector per point
    y0[k] = r.x;
    y1[k] = r.y;
    if (D > 2) y2[k] = r.z;
}

// bin_count[b+1] holds the count of bin b -> items[b+1] = ceil(count / chunk); [0] = 0
__global__ void work_items_kernel(const int32_t *__restrict__ bin_offsets, int64_t nbins, int chunk, int32_t *__restrict__ items)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) items[0] = 0;
    if (b < nbins) {
        const int32_t c = bin_offsets[b + 1] - bin_offsets[b];
        items[b + 1] = (c + chunk - 1) / chunk;
    }
}

// item_table[item_start[b] + c] = (b, c) for every chunk c of bin b
__global__ void item_table_kernel(const int32_t *__restrict__ item_start, int64_t nbins, int2 *__restrict__ table)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    const int s = item_start[b], e = item_start[b + 1];
    for (int i = s; i < e; ++i) table[i] = make_int2((int)b, i - s);
}

// identity permutation when there is a single bin and nothing to sort
__global__ void iota_kernel(int32_t *v, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}

// keys[perm[k]] = bin of sorted position k (binary search in bin_offsets); introspection path only
__global__ void coarse_keys_kernel(const int32_t *__restrict__ perm, const int32_t *__restrict__ bin_offsets, int64_t nbins,
                                   int64_t np, uint32_t *__restrict__ keys)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np) return;
    int64_t lo = 0, hi = nbins;            // largest b with bin_offsets[b] <= k
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (bin_offsets[mid] <= (int32_t)k) lo = mid; else hi = mid;
    }
    keys[perm[k]] = (uint32_t)lo;
}

static int ensure_capacity(Plan &p, int64_t np)
{
    if (np <= p.cap) return NUFFT_SUCCESS;
    auto f = [](auto *&ptr) { if (ptr) { cudaFree((void *)ptr); ptr = nullptr; } };
    f(p.d_keys[0]); f(p.d_keys[1]); f(p.d_vals[0]); f(p.d_vals[1]);
    for (int d = 0; d < 3; ++d) f(p.d_xs[d]);
    f(p.d_rec);
    const int64_t cap = np + np / 8 + 1024;   // head-room: no reallocation when Np fluctuates
    p.cap = 0;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaMalloc(&p.d_keys[i], (size_t)cap * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&p.d_vals[i], (size_t)cap * sizeof(int32_t));
    }
    for (int d = 0; d < p.D && e == cudaSuccess; ++d) e = cudaMalloc(&p.d_xs[d], (size_t)cap * p.real_bytes);
    if (p.D > 1 && e == cudaSuccess) e = cudaMalloc(&p.d_rec, (size_t)cap * 4 * p.real_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cannot allocate point buffers for %lld points", (long long)np);
        return NUFFT_ERR_ALLOC;
    }
    p.cap = cap;
    return NUFFT_SUCCESS;
}

// capacity of the per-tile digit histograms of the radix sort (RADIX counters per tile of SORT_TILE points)
static int ensure_hist(Plan &p, int64_t np)
{
    const size_t hneed = (size_t)cdiv(np, SORT_TILE) * RADIX;
    if (hneed > p.hist_cap) {
        if (p.d_hist) cudaFree(p.d_hist);
        p.hist_cap = hneed + hneed / 8;
        CUDA_TRY(cudaMalloc(&p.d_hist, p.hist_cap * sizeof(uint32_t)));
    }
    return NUFFT_SUCCESS;
}

// scratch of the single-sweep sort: [MAX_PASSES][RADIX] global digit histograms | [MAX_PASSES] tile tickets (padded to RADIX
// words) | [passes][nblk][RADIX] look-back status words; zeroed by one memset per set_points
static int ensure_onesweep(Plan &p, int64_t np, int passes, size_t *bytes)
{
    const size_t need = ((size_t)(MAX_PASSES + 1) * RADIX + (size_t)passes * cdiv(np, SORT_TILE) * RADIX) * sizeof(uint32_t);
    if (need > p.os_cap) {
        if (p.d_os) cudaFree(p.d_os);
        p.d_os = nullptr;
        p.os_cap = 0;
        CUDA_TRY(cudaMalloc(&p.d_os, need + need / 8));
        p.os_cap = need + need / 8;
    }
    *bytes = need;
    return NUFFT_SUCCESS;
}

template <typename T> static int set_points_impl(Plan &p, int64_t np, const void *const x[], int os_passes)
{
    const TileGeom &g = p.geom;
    BinGeom bg;
    bg.D = p.D;
    bg.convention = p.opts.point_convention;
    bg.xstride = p.x_stride;
    for (int d = 0; d < 3; ++d) { bg.N[d] = g.N[d]; bg.B[d] = g.B[d]; bg.nb[d] = g.nb[d]; bg.sub[d] = g.sub[d]; }
    bg.rt = g.rt;
    bg.nsub = g.nsub;
    bg.zoff = p.slab_z0;
    bg.nzown = p.slab_nz > 0 ? p.slab_nz : g.N[2];
    const T *x0 = (const T *)x[0];
    const T *x1 = p.D > 1 ? (const T *)x[1] : nullptr;
    const T *x2 = p.D > 2 ? (const T *)x[2] : nullptr;
    cudaStream_t st = p.stream;

    uint32_t *bin_count = (uint32_t *)p.d_bin_offsets;
    const bool hist = g.rt != 3;            // column-streaming plans: offsets on demand
    p.offsets_valid = hist;
    if (hist) CUDA_TRY(cudaMemsetAsync(bin_count, 0, (size_t)(p.nbins + 1) * sizeof(uint32_t), st));
    if (np > 0) {
        const int nblk = (int)cdiv(np, SORT_TILE);
        const int grid = std::min(nblk, p.num_sms * 8);       // persistent CTAs: few global histogram atomics
        auto *rec = p.D > 1 ? (typename PointRec<T>::type *)p.d_rec : nullptr;
        uint32_t *ghist = nullptr;
        if (os_passes > 0) {
            size_t bytes = 0;
            NUFFT_TRY(ensure_onesweep(p, np, os_passes, &bytes));
            CUDA_TRY(cudaMemsetAsync(p.d_os, 0, bytes, st));
            ghist = p.d_os;
        }
        if (hist) bin_keys_kernel<T, true><<<grid, 256, 0, st>>>(bg, np, x0, x1, x2, p.d_keys[0], bin_count, rec, ghist, os_passes, nblk);
        else bin_keys_kernel<T, false><<<grid, 256, 0, st>>>(bg, np, x0, x1, x2, p.d_keys[0], bin_count, rec, ghist, os_passes, nblk);
        NUFFT_COUNT_LAUNCH();
    }
    // slot b+1 holds the count of bin b and slot 0 stays 0: an inclusive scan of this array is exactly
    // cumulative_npoints_per_block (reference layout, src/blocking/gpu.jl:12); done by the caller.
    return NUFFT_SUCCESS;
}

// stable LSD radix sort of (key, index) pairs, one kernel per digit (see onesweep_kernel); the digit histograms of all
// passes are in p.d_os (bin_keys_kernel).  Same result as radix_sort_pairs.
static int onesweep_sort_pairs(Plan &p, uint32_t *k0, uint32_t *k1, int32_t *va, int32_t *vb, int64_t np, int passes, int32_t **result)
{
    cudaStream_t st = p.stream;
    const int nblk = (int)cdiv(np, SORT_TILE);
    uint32_t *ghist = p.d_os, *tickets = p.d_os + MAX_PASSES * RADIX, *status = p.d_os + (MAX_PASSES + 1) * RADIX;
    os_scan_kernel<<<passes, RADIX, 0, st>>>(ghist);
    NUFFT_COUNT_LAUNCH();
    uint32_t *kin = k0, *kout = k1;
    int32_t *vin = vb, *vout = va;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = pass * RADIX_BITS;
        const bool first = pass == 0, last = pass == passes - 1;
        const uint32_t *gb = ghist + pass * RADIX;
        uint32_t *stt = status + (size_t)pass * nblk * RADIX, *tk = tickets + pass;
        if (first && last) onesweep_kernel<true, true><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, np, shift, gb, stt, tk, kout, vout);
        else if (first) onesweep_kernel<true, false><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, np, shift, gb, stt, tk, kout, vout);
        else if (last) onesweep_kernel<false, true><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, np, shift, gb, stt, tk, kout, vout);
        else onesweep_kernel<false, false><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, np, shift, gb, stt, tk, kout, vout);
        NUFFT_COUNT_LAUNCH();
        std::swap(kin, kout);
        *result = vout;
        std::swap(vin, vout);
    }
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

// stable LSD radix sort of (key, index) pairs restricted to `bits` key bits; the first pass generates the
// indices.  keys ping-pong between k0/k1 (both clobbered), values between va/vb; *result = final values.
static int radix_sort_pairs(Plan &p, uint32_t *k0, uint32_t *k1, int32_t *va, int32_t *vb, int64_t np, int bits,
                            int32_t **result)
{
    cudaStream_t st = p.stream;
    const int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    const int nblk = (int)cdiv(np, SORT_TILE);
    const size_t hneed = (size_t)nblk * RADIX;
    NUFFT_TRY(ensure_hist(p, np));
    uint32_t *kin = k0, *kout = k1;
    int32_t *vin = vb, *vout = va;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = pass * RADIX_BITS;
        const bool first = pass == 0, last = pass == passes - 1;
        radix_hist_kernel<<<nblk, SORT_THREADS, 0, st>>>(kin, np, shift, p.d_hist, nblk);
        NUFFT_COUNT_LAUNCH();
        NUFFT_TRY(scan_u32(p, p.d_hist, (int64_t)hneed, false));
        if (first && last) radix_scatter_kernel<true, true><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, np, shift, p.d_hist, nblk, kout, vout);
        else if (first) radix_scatter_kernel<true, false><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, np, shift, p.d_hist, nblk, kout, vout);
        else if (last) radix_scatter_kernel<false, true><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, np, shift, p.d_hist, nblk, kout, vout);
        else radix_scatter_kernel<false, false><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, np, shift, p.d_hist, nblk, kout, vout);
        NUFFT_COUNT_LAUNCH();
        std::swap(kin, kout);
        *result = vout;
        std::swap(vin, vout);
    }
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

template <typename T> static int run_set_points(Plan &p, int64_t np, const void *const x[])
{
    p.Np = -1;                              // a failure below leaves the plan without points (exec then reports NUFFT_ERR_STATE)
    NUFFT_TRY(ensure_capacity(p, np));
    const bool sorts = p.nbins * p.geom.nsub > 1 && np > 0;
    const int passes = (p.key_bits + RADIX_BITS - 1) / RADIX_BITS;
    // single-sweep passes: counts and prefixes share a 30-bit field of the status words
    bool onesweep = sorts && passes <= MAX_PASSES && np < ((int64_t)1 << 30);
    if (const char *e = getenv("NUFFT_B200_ONESWEEP")) onesweep = onesweep && atoi(e) != 0;
    NUFFT_TRY(set_points_impl<T>(p, np, x, onesweep ? passes : 0));
    cudaStream_t st = p.stream;
    const int64_t nb1 = p.nbins + 1;
    if (p.offsets_valid) NUFFT_TRY(scan_u32(p, (uint32_t *)p.d_bin_offsets, nb1, true));
    p.perm_coarse_ptr = nullptr;

    // stable LSD radix sort of (key, index)
    if (!sorts) {
        if (np > 0) {
            iota_kernel<<<(unsigned)cdiv(np, 256), 256, 0, st>>>(p.d_vals[0], np);
            NUFFT_COUNT_LAUNCH();
        }
        p.d_perm = p.d_vals[0];
        p.sort_cur = 0;
    } else {
        int32_t *res = nullptr;
        if (onesweep) NUFFT_TRY(onesweep_sort_pairs(p, p.d_keys[0], p.d_keys[1], p.d_vals[1], p.d_vals[0], np, passes, &res));
        else NUFFT_TRY(radix_sort_pairs(p, p.d_keys[0], p.d_keys[1], p.d_vals[1], p.d_vals[0], np, p.key_bits, &res));
        p.d_perm = res;
        p.sort_cur = (res == p.d_vals[0]) ? 0 : 1;
    }
    // column-streaming plans (cs_spread.cuh / cs_interp.cuh) read the folded records through the permutation and pull
    // fixed-size chunks of the sorted order: no sorted copy of the coordinates, no work-item table
    if (p.geom.rt == 3) {
        CUDA_TRY(cudaGetLastError());
        return NUFFT_SUCCESS;
    }
    if (np > 0) {
        gather_points_kernel<T><<<(unsigned)cdiv(np, 256), 256, 0, st>>>(
            p.D, p.opts.point_convention, p.x_stride, np, p.d_perm, (const T *)x[0], (const typename PointRec<T>::type *)p.d_rec,
            (T *)p.d_xs[0], (T *)p.d_xs[1], (T *)p.d_xs[2]);
        NUFFT_COUNT_LAUNCH();
    }
    // work items: bins with more than `chunk` points are split
    work_items_kernel<<<(unsigned)cdiv(nb1, 256), 256, 0, st>>>(p.d_bin_offsets, p.nbins, p.geom.chunk, p.d_item_start);
    NUFFT_COUNT_LAUNCH();
    // inclusive scan of [0, n0, n1, ...]: item_start[b] = first item of bin b, item_start[nbins] = total
    NUFFT_TRY(scan_u32(p, (uint32_t *)p.d_item_start, nb1, true));
    const size_t max_items = (size_t)p.nbins + (size_t)cdiv(np, p.geom.chunk) + 1;
    if (max_items > p.item_cap) {
        if (p.d_item_table) cudaFree(p.d_item_table);
        p.item_cap = max_items + max_items / 8;
        CUDA_TRY(cudaMalloc(&p.d_item_table, p.item_cap * sizeof(int2)));
    }
    item_table_kernel<<<(unsigned)cdiv(p.nbins, 256), 256, 0, st>>>(p.d_item_start, p.nbins, p.d_item_table);
    NUFFT_COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

// rt plans order the points by (bin, sub-bin); the reference-order permutation (stable by bin only, what
// BlockDataGPU.pointperm holds after the 1-thread counting sort, src/blocking/cpu.jl:73-111) is rebuilt on demand:
// bin of every ORIGINAL index (scatter through the fine permutation), then the same stable radix sort on the bin bits.
// bin offsets (cumulative_npoints_per_block) of plans whose set_points does not build them (column-streaming)
int binning_ensure_offsets(Plan &p)
{
    if (p.offsets_valid) return NUFFT_SUCCESS;
    const TileGeom &g = p.geom;
    BinGeom bg{};
    bg.D = p.D;
    for (int d = 0; d < 3; ++d) { bg.N[d] = g.N[d]; bg.B[d] = g.B[d]; bg.nb[d] = g.nb[d]; }
    bg.zoff = p.slab_z0;
    bg.nzown = p.slab_nz > 0 ? p.slab_nz : g.N[2];
    uint32_t *bin_count = (uint32_t *)p.d_bin_offsets;
    CUDA_TRY(cudaMemsetAsync(bin_count, 0, (size_t)(p.nbins + 1) * sizeof(uint32_t), p.stream));
    if (p.Np > 0) {
        bin_hist_from_rec_kernel<<<(unsigned)cdiv(p.Np, 256), 256, 0, p.stream>>>(bg, p.Np, (const float4 *)p.d_rec, bin_count);
        NUFFT_COUNT_LAUNCH();
    }
    NUFFT_TRY(scan_u32(p, bin_count, p.nbins + 1, true));
    CUDA_TRY(cudaGetLastError());
    p.offsets_valid = true;
    return NUFFT_SUCCESS;
}

int binning_coarse_perm(Plan &p, const int32_t **perm)
{
    NUFFT_TRY(binning_ensure_offsets(p));
    if (p.geom.nsub <= 1 || p.Np <= 0 || p.nbins <= 1) { *perm = p.d_perm; return NUFFT_SUCCESS; }
    if (p.perm_coarse_ptr) { *perm = p.perm_coarse_ptr; return NUFFT_SUCCESS; }
    const int64_t np = p.Np;
    if (np > p.perm_coarse_cap) {
        if (p.d_perm_coarse) cudaFree(p.d_perm_coarse);
        p.perm_coarse_cap = 0;
        CUDA_TRY(cudaMalloc(&p.d_perm_coarse, (size_t)p.cap * sizeof(int32_t)));
        p.perm_coarse_cap = p.cap;
    }
    coarse_keys_kernel<<<(unsigned)cdiv(np, 256), 256, 0, p.stream>>>(p.d_perm, p.d_bin_offsets, p.nbins, np, p.d_keys[0]);
    NUFFT_COUNT_LAUNCH();
    int bits = 1;
    while (((int64_t)1 << bits) < p.nbins) ++bits;
    int32_t *res = nullptr;
    NUFFT_TRY(radix_sort_pairs(p, p.d_keys[0], p.d_keys[1], p.d_perm_coarse, p.d_vals[p.sort_cur ^ 1], np, bits, &res));
    p.perm_coarse_ptr = res;
    *perm = res;
    return NUFFT_SUCCESS;
}

int binning_set_points(Plan &p, int64_t np, const void *const x[], int xstride)
{
    p.x_stride = xstride;
    if (np < 0) { set_error("negative number of points"); return NUFFT_ERR_ARG; }
    if (np >= ((int64_t)1 << 31) - 4096) { set_error("number of points exceeds maximum allowed: 2^31"); return NUFFT_ERR_ARG; }
    for (int d = 0; d < p.D; ++d) {
        if (np > 0 && x[d] == nullptr) { set_error("null point array for dimension %d", d); return NUFFT_ERR_ARG; }
        p.user_x[d] = x[d];
    }
    if (p.dual_geom) {
        // column-streaming kernels when the register windows see enough points, shared-memory tiles otherwise
        const int w = (p.cs_min_cells <= 0 || (double)np * p.cs_min_cells >= (double)p.ncells) ? 0 : 1;
        p.geom = p.geom_alt[w]; p.nbins = p.nbins_alt[w]; p.key_bits = p.key_bits_alt[w];
    }
    if (p.ev_ok) { cudaEventRecord(p.ev[0], p.stream); }
    int rc = p.f64 ? run_set_points<double>(p, np, x) : run_set_points<float>(p, np, x);
    if (rc != NUFFT_SUCCESS) return rc;
    if (p.ev_ok) { cudaEventRecord(p.ev[1], p.stream); p.ev_rec[0] = true; }
    p.Np = np;
    return NUFFT_SUCCESS;
}

}  // namespace nufft
