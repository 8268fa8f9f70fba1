// ring_inst.cu — launches of the ring-window kernels (ring_spread.cuh / ring_interp.cuh: 3-D, HalfSupport(4), ComplexF32).
// A translation unit of its own: the two kernels compile in a minute, the generic instantiations of spread_inst.cu /
// interp_inst.cu (every D, M, precision) take several.
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include "ring_spread.cuh"
#include "ring_interp.cuh"

namespace nufft {

// ---- tensor map of the oversampled grid (for the TMA plane reductions / loads of the ring kernels) -------------------------
// The grid of a plan is C contiguous arrays of Nx x Ny x nz complex cells: a 3-D Float32 tensor of 2 Nx x Ny x (C nz) elements,
// boxes of 24 x 11 x 1 (12 cells x 11 rows of one plane).  cuTensorMapEncodeTiled is a driver entry point: fetched through the
// runtime (no link against libcuda).  One map per (grid pointer, dims), cached.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool grid_tensor_map(Plan &p, CUtensorMap *out)
{
    static std::mutex mu;
    static std::map<std::tuple<void *, int64_t, int64_t, int64_t, int>, CUtensorMap> cache;
    static EncodeTiledFn encode = nullptr;
    static bool tried = false;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried) {
        tried = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            encode = (EncodeTiledFn)fn;
        else
            cudaGetLastError();
    }
    if (!encode || (p.Nos[0] & 1)) return false;
    const auto key = std::make_tuple(p.d_us, (int64_t)p.Nos[0], (int64_t)p.Nos[1], (int64_t)p.nz_local, p.C);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)(2 * p.Nos[0]), (cuuint64_t)p.Nos[1], (cuuint64_t)p.nz_local * (cuuint64_t)p.C};
    const cuuint64_t strides[2] = {(cuuint64_t)(2 * p.Nos[0]) * 4, (cuuint64_t)(2 * p.Nos[0]) * 4 * (cuuint64_t)p.Nos[1]};
    const cuuint32_t box[3] = {2 * ring::TMA_ROW, 11, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.d_us, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (cache.size() > 256) cache.clear();
    cache[key] = tm;
    *out = tm;
    return true;
}

// points per work item: a z segment of <= 128 planes holds ~ 256 points per column at the headline density, and an item that
// spans more than one column pays a second window set-up (profiles/r2_ring_segment_chunk_sweep.txt: 3.09 -> 2.83 ms)
static int ring_chunk(const Plan &p)
{
    if (getenv("NUFFT_B200_CS_CHUNK")) return cs::chunk_points();
    return p.geom.B[2] <= 128 ? 256 : cs::chunk_points();
}

template <bool TMA>
static int ring_spread_go(Plan &p, const CUtensorMap &tm, const KernelParams<float> &kp, const PtrPack &pack, int cn, float2 *us, const float *nuw,
                          int zlo, int nzwrap)
{
    auto kern = ring::ring_spread_kernel<ring::NWARP, TMA>;
    const size_t smem = ring::spread_smem_bytes(p.cs_stride, TMA);
    static bool attr_done[64] = {};                // per device: function attributes are per device
    if (!attr_done[p.device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[p.device & 63] = true;
    }
    kern<<<p.num_sms, 32 * ring::NWARP, smem, p.stream>>>(tm, kp, p.geom, (int)p.Np, ring_chunk(p), p.d_perm, p.d_counters,
                                                          (const float4 *)p.d_rec, pack, cn, us, p.ncells, nuw, zlo, nzwrap, p.nz_local);
    NUFFT_COUNT_LAUNCH();
    return NUFFT_SUCCESS;
}

int ring_spread_run(Plan &p, const KernelParams<float> &kp, const PtrPack &pack, int cn, float2 *us, const float *nuw, int zlo, int nzwrap)
{
    // TMA plane retirement is OPT-IN (NUFFT_B200_TMA_SPREAD=1): measured at C3 it is 2.6 x slower than per-lane red.global.add.v2.f32
    // (7.2 against 2.8 ms) — like red.global.add.v4.f32 (7.4 ms), the 16-byte float reductions of L2 are slow on this part.
    static const bool tma_on = getenv("NUFFT_B200_TMA_SPREAD") && atoi(getenv("NUFFT_B200_TMA_SPREAD")) != 0;
    CUtensorMap tm{};
    // (the map describes the plan's whole grid: the component offset of `us` inside it is applied through the plane coordinate)
    if (tma_on && us == (float2 *)p.d_us && grid_tensor_map(p, &tm)) return ring_spread_go<true>(p, tm, kp, pack, cn, us, nuw, zlo, nzwrap);
    return ring_spread_go<false>(p, tm, kp, pack, cn, us, nuw, zlo, nzwrap);
}

template <bool TMA>
static int ring_interp_go(Plan &p, const CUtensorMap &tm, const KernelParams<float> &kp, const MutPtrPack &pack, int cn, const float2 *us,
                          float prefactor, const float *nuw, int zlo, int nzwrap)
{
    auto kern = ring::ring_interp_kernel<ring::NWARP, TMA>;
    const size_t smem = ring::interp_smem_bytes(p.cs_stride, TMA);
    static bool attr_done[64] = {};
    if (!attr_done[p.device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[p.device & 63] = true;
    }
    kern<<<p.num_sms, 32 * ring::NWARP, smem, p.stream>>>(tm, kp, p.geom, (int)p.Np, ring_chunk(p), p.d_perm, p.d_counters,
                                                          (const float4 *)p.d_rec, pack, cn, us, p.ncells, prefactor, nuw, zlo, nzwrap,
                                                          p.nz_local);
    NUFFT_COUNT_LAUNCH();
    return NUFFT_SUCCESS;
}

int ring_interp_run(Plan &p, const KernelParams<float> &kp, const MutPtrPack &pack, int cn, const float2 *us, float prefactor, const float *nuw,
                    int zlo, int nzwrap)
{
    // TMA plane loads are OPT-IN (NUFFT_B200_TMA_INTERP=1, experimental): measured at C3 they are slower than the per-lane cp.async
    // staging (4.2 against 3.0 ms: a 3-D box of 11 rows arrives later than three planes of lookahead cover) and the path is
    // not covered by the parity suite.
    static const bool tma_on = getenv("NUFFT_B200_TMA_INTERP") && atoi(getenv("NUFFT_B200_TMA_INTERP")) != 0;
    CUtensorMap tm{};
    if (tma_on && us == (const float2 *)p.d_us && grid_tensor_map(p, &tm))
        return ring_interp_go<true>(p, tm, kp, pack, cn, us, prefactor, nuw, zlo, nzwrap);
    return ring_interp_go<false>(p, tm, kp, pack, cn, us, prefactor, nuw, zlo, nzwrap);
}

}  // namespace nufft
