// tile_common.cuh — pieces shared by the shared-memory spreading (K-spread) and interpolation (K-interp) kernels:
// work-item decoding, the lane map of a warp over a point's support, the per-point record produced by the
// producer warps (kernel values + local indices + value), and small vector-load helpers.
#pragma once
#include "common.cuh"
#include "kernel_eval.cuh"

namespace nufft {

constexpr int MAX_PACK = 8;
struct PtrPack {
    const void *p[MAX_PACK];
};
struct MutPtrPack {
    void *p[MAX_PACK];
};

struct SmArgs {
    const int32_t *perm;
    const int32_t *bin_offsets;   // nbins + 1
    const int32_t *item_start;    // nbins + 1 (inclusive-scan form: item_start[b] = first item of bin b)
    const int2 *item_table;       // per work item: (bin, chunk index)
    int32_t *work_counter;
    int nbins;
};

// lane map of one warp over the (x, y) footprint of a point: lane -> (lx, lg); rows jy = lg + i * G
template <int M> struct LaneMap {
    static constexpr int W = 2 * M;
    static constexpr int G = 32 / W;                       // rows handled concurrently (W <= 24 -> G >= 1)
    static constexpr int NI = (W + G - 1) / G;             // row iterations
    static constexpr int WSLOT = ((W + 3) / 4) * 4;        // slot of a naturally ordered weight vector
    static constexpr int YSLOT = ((G * NI + 3) / 4) * 4;   // slot of the lane-transposed weight vector
};

// Per-point record layout in shared memory (units of T).
//   SPREAD = true : the last ("owned") dimension is stored by ABSOLUTE residue class of the tile coordinate:
//                   slot[((s + j) % M) * 2 + j / M] = w[j]  (s = local start index of the support), so consumer
//                   warp `w` (owner of tile coordinates = w mod M) reads its two values at slot[2w .. 2w+1]
//                   with one vector load that does not depend on the point.
//   D == 3        : the middle dimension is lane-transposed: slot[(j % G) * NI + j / G] = w[j].
template <int D, int M, bool SPREAD> struct WRecord {
    using LM = LaneMap<M>;
    static constexpr int W = 2 * M;
    static constexpr int OFF_X = 0;
    static constexpr int OFF_Y = LM::WSLOT;
    static constexpr int OFF_Z = LM::WSLOT + (D == 3 ? LM::YSLOT : LM::WSLOT);
    static constexpr int SIZE = (D == 1) ? LM::WSLOT : (D == 2 ? 2 * LM::WSLOT : 2 * LM::WSLOT + LM::YSLOT);
    __host__ __device__ static constexpr int offset(int d) { return d == 0 ? OFF_X : (d == 1 ? OFF_Y : OFF_Z); }

    template <typename T> __device__ static __forceinline__ void store(T *rec, int d, const T *w, int start)
    {
        T *dst = rec + offset(d);
        const bool owned = (d == D - 1);
        if (SPREAD && owned) {
            int rho = start % M;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                dst[rho * 2 + j / M] = w[j];
                rho = (rho + 1 == M) ? 0 : rho + 1;
            }
        } else if (D == 3 && d == 1) {
#pragma unroll
            for (int j = 0; j < W; ++j) dst[(j % LM::G) * LM::NI + j / LM::G] = w[j];
#pragma unroll
            for (int j = W; j < LM::G * LM::NI; ++j) dst[(j % LM::G) * LM::NI + j / LM::G] = (T)0;
        } else {
#pragma unroll
            for (int j = 0; j < W; ++j) dst[j] = w[j];
        }
    }
};

__device__ __forceinline__ int wrap1(int g, int N)   // g in [-N, 2N)
{
    g += (g < 0) ? N : 0;
    g -= (g >= N) ? N : 0;
    return g;
}
__device__ __forceinline__ int wrap_any(int g, int N)
{
    g %= N;
    return g < 0 ? g + N : g;
}

__device__ __forceinline__ int pmod(int a, int m)    // a mod m in [0, m), m > 0 compile-time in practice
{
    int r = a % m;
    return r < 0 ? r + m : r;
}

// aligned vector loads of N consecutive T (N * sizeof(T) a power of two <= 16 and the address aligned to it)
template <typename T, int N> struct VecLoad {
    __device__ static __forceinline__ void load(const T *p, T *out)
    {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = p[i];
    }
};
template <> struct VecLoad<float, 2> {
    __device__ static __forceinline__ void load(const float *p, float *out)
    {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        out[0] = v.x; out[1] = v.y;
    }
};
template <> struct VecLoad<float, 4> {
    __device__ static __forceinline__ void load(const float *p, float *out)
    {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    }
};
template <> struct VecLoad<float, 8> {
    __device__ static __forceinline__ void load(const float *p, float *out)
    {
        VecLoad<float, 4>::load(p, out);
        VecLoad<float, 4>::load(p + 4, out + 4);
    }
};
template <> struct VecLoad<double, 2> {
    __device__ static __forceinline__ void load(const double *p, double *out)
    {
        const double2 v = *reinterpret_cast<const double2 *>(p);
        out[0] = v.x; out[1] = v.y;
    }
};
template <> struct VecLoad<double, 4> {
    __device__ static __forceinline__ void load(const double *p, double *out)
    {
        VecLoad<double, 2>::load(p, out);
        VecLoad<double, 2>::load(p + 2, out + 2);
    }
};
template <> struct VecLoad<double, 8> {
    __device__ static __forceinline__ void load(const double *p, double *out)
    {
        VecLoad<double, 4>::load(p, out);
        VecLoad<double, 4>::load(p + 4, out + 4);
    }
};

// asynchronous global -> shared copy of one cell (LDGSTS; completion via cp_async_wait_all)
template <int BYTES> __device__ __forceinline__ void cp_async_cell(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <typename T, bool CPLX> __device__ __forceinline__ typename CellOf<T, CPLX>::type load_value(const void *vp, int64_t i)
{
    using Cell = typename CellOf<T, CPLX>::type;
    return ((const Cell *)vp)[i];
}

template <typename T, bool CPLX> __device__ __forceinline__ void store_value(void *vp, int64_t i, typename CellOf<T, CPLX>::type v)
{
    using Cell = typename CellOf<T, CPLX>::type;
    ((Cell *)vp)[i] = v;
}

// Work item -> (bin, [k0, k1)); one table lookup + two offsets (no search).
__device__ __forceinline__ void decode_item(const SmArgs &a, int item, int chunk, int &bin, int &k0, int &k1)
{
    const int2 e = a.item_table[item];
    bin = e.x;
    const int off = a.bin_offsets[bin], end = a.bin_offsets[bin + 1];
    k0 = off + e.y * chunk;
    k1 = min(k0 + chunk, end);
}

// Shared-memory carve-up shared by both kernels.
//   tile | values[2][batch] | cs tables | records[2][batch][REC] | starts[2][batch] (int4) | results (interp)
template <typename T, bool CPLX, int D, int M, bool SPREAD>
__host__ __device__ inline size_t sm_dynamic_bytes(const TileGeom &g, int cs_stride)
{
    using Cell = typename CellOf<T, CPLX>::type;
    size_t b = (size_t)g.tile_cells * sizeof(Cell);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)2 * g.batch * sizeof(Cell);                                  // values / results, double buffered
    b += (size_t)(((D * cs_stride) + 3) / 4 * 4) * sizeof(T);                  // kernel coefficient tables
    b += (size_t)2 * g.batch * WRecord<D, M, SPREAD>::SIZE * sizeof(T);        // weight records, double buffered
    b += (size_t)2 * g.batch * sizeof(int4);                                   // local start indices (+ original index)
    return b + 32;
}

}  // namespace nufft
