// wp_interp.cuh — K-interp, warp-private-tile variant (3-D, HalfSupport(4), ComplexF32).  Mirror image of
// wp_spread.cuh; replaces src/interpolation/gpu.jl:211-395 (same sums, different order).
//
//   global -> shared   the bin's padded tile (15 x 15 x 15 cells, private to the warp) is staged with cp.async: 8 lanes per
//                      tile row, one 8-byte copy per cell into the rotated row layout of wp_spread.cuh; all copies of a
//                      tile are in flight at once, the other warps of the SM compute meanwhile;
//   shared -> regs     the padded footprint (11 x 11 x 11 cells) of a 4 x 4 x 4-cell sub-bin is loaded into the warp's
//                      registers when the sub-bin changes (44 LDS.64 per lane, conflict-free, about once per 8 points);
//   per point          branch-free register dot product: 9 shared-memory loads (zero-padded weights), 48 FFMA2 / 2 FMUL2
//                      per lane, then a 5-step butterfly over the lanes; results of a batch stay in lanes and are
//                      scattered through the permutation once per batch.
#pragma once
#include "wp_spread.cuh"

namespace nufft {
namespace wp {

constexpr int IREC_F = 56;                // floats per point record (interpolation)
constexpr int IOFF_WXP = 0;               // [0..11]  wx_pad[0..10], 0
constexpr int IOFF_KEY = 12;              // [12]     sub-bin index
constexpr int IOFF_WY = rt::OFF_WY;       // [16..39] wyT rows (rt::store_y)
constexpr int IOFF_WZ = 40;               // [40..51] wz_pad[0..10], 0

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { return rt::ffma2(a, b, c); }

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m)
{
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
    const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((u64)hi << 32) | lo;
}

template <typename Inst>                   // instantiated only by the ComplexF32 translation unit
__global__ void __launch_bounds__(32 * NWARP, 1)
wp_interp_kernel(KernelParams<float> kp, TileGeom g, SmArgs a, const float *__restrict__ xs0, const float *__restrict__ xs1,
                 const float *__restrict__ xs2, MutPtrPack vp, int C, const float2 *__restrict__ us, int64_t ncells,
                 float prefactor, const float *__restrict__ nu_weights)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tiles = (float2 *)smem_raw;                                      // [NWARP][TILE_CELLS]
    float *rec_all = (float *)(smem_raw + (size_t)NWARP * TILE_CELLS * sizeof(float2));   // [NWARP][BATCH][IREC_F]
    float *cs_s = rec_all + NWARP * BATCH * IREC_F;                          // [3][cs_stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const int total_items = a.item_start[a.nbins];
    float2 *tile = tiles + warp * TILE_CELLS;
    const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
    float *rec_w = rec_all + warp * BATCH * IREC_F;

    for (int i = tid; i < 3 * kp.cs_stride; i += 32 * NWARP) cs_s[i] = kp.cs[i];
    __syncthreads();                                   // the only CTA barrier: coefficient tables

    const rt::LaneSlots ls = rt::lane_slots(lane);
    const int ep = lane / 3, ed = lane - 3 * ep;       // evaluation role: lane = 3 * point + dimension
    const float *xs_d = ed == 0 ? xs0 : (ed == 1 ? xs1 : xs2);
    const KernelParams<float> kl = lane_kernel_params(kp, ed);
    const float *cs_d = cs_s + ed * kp.cs_stride;
    const int fj = lane & 7, fr = lane >> 3;           // staging role: 8 lanes per tile row, 4 rows per pass
    const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];

    u64 G[4][P];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < P; ++i) G[k][i] = 0ull;

    int nxt_item, nxt_bin = 0, nxt_k0 = 0, nxt_k1 = 0;
    auto fetch_item = [&]() {
        int it = 0;
        if (lane == 0) it = atomicAdd(a.work_counter, 1);
        nxt_item = __shfl_sync(FULL, it, 0);
        if (nxt_item < total_items) decode_item(a, nxt_item, g.chunk, nxt_bin, nxt_k0, nxt_k1);
    };
    fetch_item();

    while (nxt_item < total_items) {
        const int bin = nxt_bin, k0 = nxt_k0, k1 = nxt_k1;
        fetch_item();
        int b = bin;
        const int bx = b % g.nb[0]; b /= g.nb[0];
        const int by = b % g.nb[1]; b /= g.nb[1];
        const int bz = b;
        const int org0 = bx * BIN, org1 = by * BIN, org2 = bz * BIN;          // first cell of the bin
        const int org_d = ed == 0 ? org0 : (ed == 1 ? org1 : org2);
        const int nbatches = (k1 - k0 + BATCH - 1) / BATCH;

        for (int c = 0; c < C; ++c) {
            // ---- stage the tile: global -> shared (periodic), every copy in flight at once ------------------------------
            {
                const float2 *u = us + (int64_t)c * ncells;
                int gxa = wrap1(org0 - (M - 1) - 1 + 2 * fj, Nx);      // cells 2j - 1, 2j of the tile row
                const int gxb = wrap1(gxa + 1, Nx);
                for (int rb = 0; rb < TE * TE; rb += 4) {
                    const int row = rb + fr;
                    if (row < TE * TE) {
                        const int tz = row / TE, ty = row - TE * tz;
                        const int rot = 11 * ty + 2 * fj;
                        const int gy = wrap1(org1 - (M - 1) + ty, Ny), gz = wrap1(org2 - (M - 1) + tz, Nz);
                        const float2 *grow = u + ((int64_t)gz * Ny + gy) * Nx;
                        float2 *trow = tile + row * ROW;
                        if (fj > 0) cp_async_cell<8>(trow + ((rot - 1) & (ROW - 1)), grow + gxa);
                        cp_async_cell<8>(trow + (rot & (ROW - 1)), grow + gxb);
                    }
                }
            }
            float2 *vc = (float2 *)vp.p[c];
            float xq = 0.f;
            auto prefetch = [&](int bi) {
                const int k = k0 + bi * BATCH + ep;
                if (bi < nbatches && lane < 3 * BATCH && k < k1) xq = xs_d[k];
            };
            prefetch(0);
            cp_async_wait_all();
            __syncwarp();
            int cur_sub = -1;

            auto load_window = [&](int sub) {
                const int sz = sub & 1, sx = (sub >> 1) & 1, sy = sub >> 2;
                int off[4];
#pragma unroll
                for (int k = 0; k < 3; ++k) off[k] = 8 * plane_cell(4 * sx + ls.x, 4 * sy + ls.g + 3 * k);
                off[3] = 8 * plane_cell(4 * sx + ls.x3, 4 * sy + ls.y3);
                const unsigned win_s = tile_s + 4 * sz * (PLANE * 8);
#pragma unroll
                for (int i = 0; i < P; ++i) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) lds64_to(G[k][i], win_s + off[k] + i * (PLANE * 8));
                }
            };

            for (int bi = 0; bi < nbatches; ++bi) {
                const int kb = k0 + bi * BATCH;
                const int nb = min(BATCH, k1 - kb);
                const float x = xq;
                prefetch(bi + 1);
                // original index of the point this lane will store (lanes 0 .. nb - 1)
                int32_t n_out = 0;
                if (lane < nb) n_out = a.perm[kb + lane];

                // ---- evaluate: 3 lanes per point -----------------------------------------------------------------
                const bool act = lane < 3 * BATCH && ep < nb;
                int t = 0;
                if (act) {
                    float *r = rec_w + ep * IREC_F;
                    float w[W];
                    t = eval_kernel_values<float, M>(kl, cs_d, 0, x, w) - org_d;
                    float pw[P];
                    rt::pad_shift(w, t & 3, pw);
                    if (ed == 1) {
                        rt::store_y(r, pw);
                    } else {
                        float4 *q = reinterpret_cast<float4 *>(r + (ed == 0 ? IOFF_WXP : IOFF_WZ));
                        q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
                        q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
                        q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
                    }
                }
                {
                    const int src = min(3 * ep, 27);
                    const int t0 = __shfl_sync(FULL, t, src), t1 = __shfl_sync(FULL, t, src + 1), t2 = __shfl_sync(FULL, t, src + 2);
                    if (act && ed == 0) {
                        const int key = ((((t1 >> 2) << 1) | (t0 >> 2)) << 1) | (t2 >> 2);
                        rec_w[ep * IREC_F + IOFF_KEY] = __int_as_float(key);
                    }
                }
                __syncwarp();

                // ---- per point: register dot product + butterfly over the lanes ------------------------------------
                u64 res = 0ull;                            // lane p keeps the result of point p of the batch
                for (int p = 0; p < nb; ++p) {
                    const float *r = rec_w + p * IREC_F;
                    const float4 wy = *reinterpret_cast<const float4 *>(r + IOFF_WY + 4 * ls.row);
                    const float wx = r[IOFF_WXP + ls.x], wx3 = r[IOFF_WXP + ls.x3];
                    const float4 *zq = reinterpret_cast<const float4 *>(r + IOFF_WZ);
                    const float4 z0 = zq[0], z1 = zq[1], z2 = zq[2];
                    const int sub = __float_as_int(r[IOFF_KEY]);
                    if (sub != cur_sub) {
                        load_window(sub);
                        cur_sub = sub;
                    }
                    const float wz[P] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w, z2.x, z2.y, z2.z};
                    u64 tk[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const u64 wzz = pk2(wz[i], wz[i]);
#pragma unroll
                        for (int k = 0; k < 4; ++k) tk[k] = ffma2(G[k][i], wzz, tk[k]);
                    }
                    const u64 w01 = fmul2(pk2(wx, wx), pk2(wy.x, wy.y));
                    const u64 w23 = fmul2(pk2(wx, wx3), pk2(wy.z, wy.w));
                    const float2 wa = unpk2(w01), wb = unpk2(w23);
                    u64 acc = fmul2(tk[0], pk2(wa.x, wa.x));
                    acc = ffma2(tk[1], pk2(wa.y, wa.y), acc);
                    acc = ffma2(tk[2], pk2(wb.x, wb.x), acc);
                    acc = ffma2(tk[3], pk2(wb.y, wb.y), acc);
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) acc = rt::fadd2(acc, shfl_xor_u64(acc, o));
                    if (lane == p) res = acc;
                }
                if (lane < nb) {
                    const float2 rv = unpk2(res);
                    const float scale = prefactor * (nu_weights ? nu_weights[n_out] : 1.f);
                    vc[n_out] = make_float2(rv.x * scale, rv.y * scale);
                }
                __syncwarp();
            }
            __syncwarp();                                  // the tile is re-staged by the next component / item
        }
    }
}

inline size_t interp_smem_bytes(int cs_stride)
{
    return (size_t)NWARP * (TILE_CELLS * sizeof(float2) + BATCH * IREC_F * sizeof(float)) + (size_t)(3 * cs_stride + 4) * sizeof(float) + 16;
}

}  // namespace wp
}  // namespace nufft
