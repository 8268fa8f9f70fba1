// spread_inst.cu — instantiates the K-spread kernels for one (T, CPLX) pair.
// Compiled four times: -DINST_T=float|double -DINST_CPLX=0|1 (keeps each nvcc job short).
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "spread.cuh"
#include "cs_spread.cuh"

#ifndef INST_T
#define INST_T float
#endif
#ifndef INST_CPLX
#define INST_CPLX 1
#endif

namespace nufft {

int ring_spread_run(Plan &p, const KernelParams<float> &kp, const PtrPack &pack, int cn, float2 *us, const float *nuw, int zlo, int nzwrap);

template <typename T, bool CPLX, int D, int M>
static int spread_launch(Plan &p, const void *const vp[], const nufft_callbacks *cb)
{
    using Cell = typename CellOf<T, CPLX>::type;
    cudaStream_t st = p.stream;
    const int64_t np = p.Np;
    if (np == 0) return NUFFT_SUCCESS;
    const KernelParams<T> kp = make_kernel_params<T>(p);
    const T *nuw = (cb && cb->nu_weights) ? (const T *)cb->nu_weights : nullptr;
    const T *xs0 = (const T *)p.d_xs[0], *xs1 = (const T *)p.d_xs[1], *xs2 = (const T *)p.d_xs[2];
    for (int c0 = 0; c0 < p.C; c0 += MAX_PACK) {
        const int cn = std::min(MAX_PACK, p.C - c0);
        PtrPack pack{};
        for (int c = 0; c < cn; ++c) pack.p[c] = vp[c0 + c];
        Cell *us = (Cell *)p.d_us + (int64_t)c0 * p.ncells;
        if constexpr (std::is_same<T, float>::value && D == 3 && M == 4) {
            if (p.geom.rt == 3 && p.method == NUFFT_METHOD_SHARED_MEMORY) {
                // z-slab plans address planes relative to the first stored one and never wrap along z
                const int zlo = p.slab_nz > 0 ? p.slab_z0 - (M - 1) : 0, nzwrap = p.slab_nz > 0 ? (1 << 30) : (int)p.Nos[2];
                CUDA_TRY(cudaMemsetAsync(p.d_counters, 0, sizeof(int32_t), st));
                static bool ring_off = getenv("NUFFT_B200_RING") && atoi(getenv("NUFFT_B200_RING")) == 0;
                if constexpr (CPLX) {
                    if (!ring_off) {                       // ring-window kernel: its own translation unit (ring_inst.cu)
                        NUFFT_TRY(ring_spread_run(p, kp, pack, cn, us, nuw, zlo, nzwrap));
                        continue;
                    }
                }
                if (p.slab_nz > 0) { set_error("z-slab plans need the ring kernels (complex data)"); return NUFFT_ERR_UNSUPPORTED; }
                auto kern = cs::cs_spread_kernel<CPLX>;
                const size_t smem = cs::spread_smem_bytes(p.cs_stride);
                static bool attr_done[64] = {};          // per instantiation AND device: function attributes are per device
                if (!attr_done[p.device & 63]) {
                    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    attr_done[p.device & 63] = true;
                }
                kern<<<p.num_sms, 32 * cs::NWARP, smem, st>>>(kp, p.geom, (int)np, cs::chunk_points(), p.d_perm, p.d_counters, (const float4 *)p.d_rec, pack, cn, us,
                                                        p.ncells, nuw);
                NUFFT_COUNT_LAUNCH();
                continue;
            }
        }
        if (p.method == NUFFT_METHOD_GLOBAL_MEMORY) {
            spread_gm_kernel<T, CPLX, D, M><<<(unsigned)cdiv(np, 128), 128, 0, st>>>(
                kp, np, xs0, xs1, xs2, p.d_perm, pack, cn, us, p.ncells, nuw);
            NUFFT_COUNT_LAUNCH();
        } else {
            auto kern = spread_sm_kernel<T, CPLX, D, M>;
            const size_t smem = sm_dynamic_bytes<T, CPLX, D, M, true>(p.geom, p.cs_stride);
            static size_t smem_set_dev[64] = {};               // per instantiation and device: attribute / occupancy queried once per size
            static int occ_dev[64] = {};
            size_t &smem_set = smem_set_dev[p.device & 63];
            int &occ = occ_dev[p.device & 63];
            if (smem_set != smem + 1) {
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * (M + SPREAD_NPROD), smem));
                smem_set = smem + 1;
            }
            const int nsm = p.num_sms;
            if (occ < 1) { set_error("spread_sm_kernel cannot be resident (smem %zu bytes)", smem); return NUFFT_ERR_UNSUPPORTED; }
            SmArgs a{p.d_perm, p.d_bin_offsets, p.d_item_start, p.d_item_table, p.d_counters, (int)p.nbins};
            CUDA_TRY(cudaMemsetAsync(p.d_counters, 0, sizeof(int32_t), st));
            kern<<<nsm * occ, 32 * (M + SPREAD_NPROD), smem, st>>>(kp, p.geom, a, xs0, xs1, xs2, pack, cn, us, p.ncells, nuw);
            NUFFT_COUNT_LAUNCH();
        }
    }
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

// NUFFT_DEV_M=<M> (development builds only) restricts the instantiated half supports to one value
#ifdef NUFFT_DEV_M
#define NUFFT_M_CASES(D_) case NUFFT_DEV_M: return spread_launch<T, CPLX, D_, NUFFT_DEV_M>(p, vp, cb);
#else
#define NUFFT_M_CASES(D_) \
    case 2: return spread_launch<T, CPLX, D_, 2>(p, vp, cb);   \
    case 3: return spread_launch<T, CPLX, D_, 3>(p, vp, cb);   \
    case 4: return spread_launch<T, CPLX, D_, 4>(p, vp, cb);   \
    case 5: return spread_launch<T, CPLX, D_, 5>(p, vp, cb);   \
    case 6: return spread_launch<T, CPLX, D_, 6>(p, vp, cb);   \
    case 7: return spread_launch<T, CPLX, D_, 7>(p, vp, cb);   \
    case 8: return spread_launch<T, CPLX, D_, 8>(p, vp, cb);   \
    case 9: return spread_launch<T, CPLX, D_, 9>(p, vp, cb);   \
    case 10: return spread_launch<T, CPLX, D_, 10>(p, vp, cb);   \
    case 11: return spread_launch<T, CPLX, D_, 11>(p, vp, cb);   \
    case 12: return spread_launch<T, CPLX, D_, 12>(p, vp, cb);
#endif

template <typename T, bool CPLX> int spread_dispatch(Plan &p, const void *const vp[], const nufft_callbacks *cb)
{
    switch (p.D) {
    case 1: switch (p.M) { NUFFT_M_CASES(1) } break;
    case 2: switch (p.M) { NUFFT_M_CASES(2) } break;
    case 3: switch (p.M) { NUFFT_M_CASES(3) } break;
    }
    set_error("spreading kernel not instantiated for D = %d, M = %d", p.D, p.M);
    return NUFFT_ERR_UNSUPPORTED;
}

template int spread_dispatch<INST_T, (INST_CPLX != 0)>(Plan &, const void *const[], const nufft_callbacks *);

}  // namespace nufft
