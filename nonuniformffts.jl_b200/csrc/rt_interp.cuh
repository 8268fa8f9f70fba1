// rt_interp.cuh — K-interp, register-window variant (3-D, HalfSupport(4), Float32).  See rt_common.cuh for the idea.
//
// Replaces src/interpolation/gpu.jl:211-395 for this configuration class (same sums, different order).
//
// CTA = 4 independent warps sharing one read-only tile (the bin's padded subgrid, staged with cp.async), persistent,
// pulling (bin, chunk) work items from a device counter.  Inside an item the warps pull 32-point batches from a
// shared-memory counter; a batch never leaves its warp (no CTA barrier between tile loads):
//   evaluate   one thread per point: x / y / z kernel values zero-padded to the footprint of the point's 4 x 4 x 4-cell
//              sub-bin (11 cells per dimension) -> the warp's private record buffer; the (column, z block) key stays in
//              the lane's register;
//   accumulate lane L keeps its 4 footprint cells of 12 planes in registers: three GROUPS of 4 planes, group g = planes
//              4g .. 4g + 3 of the column, held in register group g % 3.  Points arrive ordered by (column, z): a run of
//              points of the same (column, z block) is a branch-free loop of 6 shared-memory loads and 50 FFMA2 per
//              point; moving to the next z block loads ONE new group (16 LDS.64).  The 32 lane partials of 32
//              consecutive points are summed by a streaming butterfly transposition (31 exchanges per 32 points)
//              that leaves point bitrev5(L) in lane L;
//   store      prefactor, non-uniform callback, scatter through the permutation.
#pragma once
#include <type_traits>
#include "rt_common.cuh"

namespace nufft {
namespace rt {

constexpr int INTERP_NW = 4;
constexpr int IOFF_WZ = 0;                        // record: [0..11] wz_pad, [16..39] wyT, [40..51] wx_pad
constexpr int PLANE_IT = (23 * 23 + 31) / 32;     // cp.async per lane and tile plane (bins of 16 x 16 cells in x, y)

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m)
{
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
    const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((u64)hi << 32) | lo;
}

template <bool CPLX>
__global__ void __launch_bounds__(32 * INTERP_NW, 2)
rt_interp_kernel(KernelParams<float> kp, TileGeom g, SmArgs a, const float *__restrict__ xs0, const float *__restrict__ xs1,
                 const float *__restrict__ xs2, MutPtrPack vp, int C, const typename CellOf<float, CPLX>::type *__restrict__ us,
                 int64_t ncells, float prefactor, const float *__restrict__ nu_weights)
{
    using Cell = typename CellOf<float, CPLX>::type;
    constexpr int NT = 32 * INTERP_NW;
    constexpr int NG = CPLX ? 4 : 2;               // u64 registers per plane: complex (re, im) per slot; real (slot k, k + 1)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_bytes = (g.tile_cells * (int)sizeof(Cell) + 15) & ~15;
    Cell *tile = (Cell *)smem_raw;
    float *rec_all = (float *)(smem_raw + tile_bytes);                // [NW][BATCH][REC_F]
    float *cs_s = rec_all + INTERP_NW * BATCH * REC_F;                // [3][cs_stride]
    __shared__ int s_item[2][4];
    __shared__ int s_batch;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Tx = g.T[0], Ty = g.T[1], Tz = g.T[2], Sx = g.S[0], S2 = g.S[2];
    const int total_items = a.item_start[a.nbins];
    float *rec_w = rec_all + warp * BATCH * REC_F;

    const LaneSlots ls = lane_slots(lane);
    const int off0 = ls.g * Sx + ls.x, off3 = ls.y3 * Sx + ls.x3;     // slot k < 3: off0 + 3 k Sx
    const int brev = (int)(__brev((unsigned)lane) >> 27);

    for (int i = tid; i < 3 * kp.cs_stride; i += NT) cs_s[i] = kp.cs[i];
    if (tid == 0) {
        const int item = atomicAdd(a.work_counter, 1);
        s_item[0][0] = item;
        if (item < total_items) decode_item(a, item, g.chunk, s_item[0][1], s_item[0][2], s_item[0][3]);
        s_batch = 0;
    }
    __syncthreads();

    for (int it = 0;; ++it) {
        const int *cur = s_item[it & 1];
        if (cur[0] >= total_items) break;
        const int bin = cur[1], k0 = cur[2], k1 = cur[3];
        if (tid == NT - 1) {
            int *nxt = s_item[(it + 1) & 1];
            const int item = atomicAdd(a.work_counter, 1);
            nxt[0] = item;
            if (item < total_items) decode_item(a, item, g.chunk, nxt[1], nxt[2], nxt[3]);
        }
        int b = bin;
        const int bx = b % g.nb[0]; b /= g.nb[0];
        const int by = b % g.nb[1]; b /= g.nb[1];
        const int bz = b;
        const int org0 = bx * g.B[0], org1 = by * g.B[1], org2 = bz * g.B[2];
        const int nbatches = (k1 - k0 + BATCH - 1) / BATCH;

        for (int c = 0; c < C; ++c) {
            // ---- tile staging: lanes sweep the flattened (x, y) plane, warp w takes planes w, w + 4, ... ----------
            {
                const Cell *u = us + (int64_t)c * ncells;
                const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];
                const int x0 = org0 - (M - 1), y0 = org1 - (M - 1), z0 = org2 - (M - 1);
                const int nxy = Tx * Ty;
                int goff[PLANE_IT], soff[PLANE_IT];
#pragma unroll
                for (int j = 0; j < PLANE_IT; ++j) {
                    const int i = lane + 32 * j;
                    const int y = i / Tx, x = i - y * Tx;
                    const bool in = i < nxy;
                    goff[j] = in ? wrap1(y0 + y, Ny) * Nx + wrap1(x0 + x, Nx) : -1;
                    soff[j] = y * Sx + x;
                }
                for (int z = warp; z < Tz; z += INTERP_NW) {
                    const Cell *gplane = u + (int64_t)wrap1(z0 + z, Nz) * Ny * Nx;
                    Cell *tplane = tile + z * S2;
#pragma unroll
                    for (int j = 0; j < PLANE_IT; ++j)
                        if (goff[j] >= 0) cp_async_cell<(int)sizeof(Cell)>(tplane + soff[j], gplane + goff[j]);
                }
                cp_async_wait_all();
            }
            __syncthreads();

            // ---- batches of this item, pulled by the warps ------------------------------------------------------
            int bi = 0;
            if (lane == 0) bi = atomicAdd(&s_batch, 1);
            bi = __shfl_sync(0xffffffffu, bi, 0);
            float xq = 0.f, yq = 0.f, zq = 0.f;
            {
                const int k = k0 + bi * BATCH + lane;
                if (bi < nbatches && k < k1) { xq = xs0[k]; yq = xs1[k]; zq = xs2[k]; }
            }
            while (bi < nbatches) {
                const int kb = k0 + bi * BATCH;
                const int nb = min(BATCH, k1 - kb);
                // next batch of this warp: index and coordinates are fetched while this one is processed
                int bn = 0;
                if (lane == 0) bn = atomicAdd(&s_batch, 1);
                bn = __shfl_sync(0xffffffffu, bn, 0);
                const float x = xq, y = yq, z = zq;
                {
                    const int k = k0 + bn * BATCH + lane;
                    if (bn < nbatches && k < k1) { xq = xs0[k]; yq = xs1[k]; zq = xs2[k]; }
                }
                // output slot of this lane after the butterfly: point bitrev5(lane) of the batch
                const bool has_out = brev < nb;
                const int32_t n_out = has_out ? a.perm[kb + brev] : 0;

                // ---- evaluate: one thread per point -----------------------------------------------------------
                int mykey = -1;                            // (column y << 12) | (column x << 8) | z block
                if (lane < nb) {
                    float *r = rec_w + lane * REC_F;
                    float w[W], pw[P];
                    const int tx = eval_kernel_values<float, M>(kp, cs_s, 0, x, w) - org0;
                    pad_shift(w, tx & 3, pw);
                    float4 *q = reinterpret_cast<float4 *>(r + OFF_WX);
                    q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
                    q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
                    q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
                    const int ty = eval_kernel_values<float, M>(kp, cs_s + kp.cs_stride, 1, y, w) - org1;
                    pad_shift(w, ty & 3, pw);
                    store_y(r, pw);
                    const int tz = eval_kernel_values<float, M>(kp, cs_s + 2 * kp.cs_stride, 2, z, w) - org2;
                    pad_shift(w, tz & 3, pw);
                    q = reinterpret_cast<float4 *>(r + IOFF_WZ);
                    q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
                    q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
                    q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
                    mykey = ((ty >> 2) << 12) | ((tx >> 2) << 8) | (tz >> 2);
                }
                __syncwarp();
                // runs of equal key: bit p of `starts` is set when point p opens a new (column, z block)
                unsigned starts;
                {
                    const int prev = __shfl_up_sync(0xffffffffu, mykey, 1);
                    starts = __ballot_sync(0xffffffffu, lane < nb && (lane == 0 || mykey != prev));
                }

                // ---- accumulate ---------------------------------------------------------------------------------
                u64 G[3][4][NG];
#pragma unroll
                for (int g3 = 0; g3 < 3; ++g3)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < NG; ++k) G[g3][i][k] = 0ull;
                u64 c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, res = 0;
                const float *pY = rec_w + OFF_WY + 4 * ls.row, *pX = rec_w + OFF_WX + ls.x, *pX3 = rec_w + OFF_WX + ls.x3;

                // streaming butterfly: a binary counter of partial transposes (levels xor 16, 8, 4, 2, 1); the earlier
                // point of a pair stays in the lanes whose bit is clear
                auto xchg = [&](u64 ea, u64 la, int m) -> u64 {
                    const bool bit = (lane & m) != 0;
                    const u64 send = bit ? ea : la, keep = bit ? la : ea;
                    return fadd2(keep, shfl_xor_u64(send, m));
                };
                auto bfly = [&](int p, u64 v) {
                    if (!(p & 1)) { c0 = v; return; }
                    v = xchg(c0, v, 16);
                    if (!(p & 2)) { c1 = v; return; }
                    v = xchg(c1, v, 8);
                    if (!(p & 4)) { c2 = v; return; }
                    v = xchg(c2, v, 4);
                    if (!(p & 8)) { c3 = v; return; }
                    v = xchg(c3, v, 2);
                    if (!(p & 16)) { c4 = v; return; }
                    res = xchg(c4, v, 1);
                };
                // group g of column base cb -> register group g % 3 (in-place loads keep the register state put)
                auto load_group = [&](const Cell *cb, int g) {
#define NUFFT_RT_LOADG(R3)                                                                                    \
    _Pragma("unroll") for (int i = 0; i < 4; ++i) {                                                           \
        const Cell *pl = cb + min(4 * g + i, Tz - 1) * S2;                                                    \
        if constexpr (CPLX) {                                                                                 \
            lds64_inplace(G[R3][i][0], pl + off0); lds64_inplace(G[R3][i][1], pl + off0 + 3 * Sx);            \
            lds64_inplace(G[R3][i][2], pl + off0 + 6 * Sx); lds64_inplace(G[R3][i][3], pl + off3);            \
        } else {                                                                                              \
            lds32x2_inplace(G[R3][i][0], pl + off0, pl + off0 + 3 * Sx);                                      \
            lds32x2_inplace(G[R3][i][1], pl + off0 + 6 * Sx, pl + off3);                                      \
        }                                                                                                     \
    }
                    const int r3 = g % 3;
                    if (r3 == 0) { NUFFT_RT_LOADG(0) } else if (r3 == 1) { NUFFT_RT_LOADG(1) } else { NUFFT_RT_LOADG(2) }
#undef NUFFT_RT_LOADG
                };
                // points [p0, p1) share the window: plane i of the window = register group (ROT + i / 4) % 3, plane i % 4
                auto run = [&](auto rot, int p0, int p1) {
                    constexpr int ROT = decltype(rot)::value;
                    for (int p = p0; p < p1; ++p) {
                        const float *r = rec_w + p * REC_F;
                        const float4 za = *reinterpret_cast<const float4 *>(r + IOFF_WZ);
                        const float4 zb = *reinterpret_cast<const float4 *>(r + IOFF_WZ + 4);
                        const float4 zc = *reinterpret_cast<const float4 *>(r + IOFF_WZ + 8);
                        const float4 wy = *reinterpret_cast<const float4 *>(pY + p * REC_F);
                        const float wx = pX[p * REC_F], wx3 = pX3[p * REC_F];
                        const float wz[11] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w, zc.x, zc.y, zc.z};
                        u64 h[NG];
#pragma unroll
                        for (int k = 0; k < NG; ++k) h[k] = fmul2(pk2(wz[0], wz[0]), G[ROT % 3][0][k]);
#pragma unroll
                        for (int i = 1; i < 11; ++i)
#pragma unroll
                            for (int k = 0; k < NG; ++k) h[k] = ffma2(pk2(wz[i], wz[i]), G[(ROT + i / 4) % 3][i % 4][k], h[k]);
                        const u64 w01 = fmul2(pk2(wx, wx), pk2(wy.x, wy.y));
                        const u64 w23 = fmul2(pk2(wx, wx3), pk2(wy.z, wy.w));
                        u64 v;
                        if constexpr (CPLX) {
                            const float2 wa = unpk2(w01), wb = unpk2(w23);
                            v = fmul2(pk2(wa.x, wa.x), h[0]);
                            v = ffma2(pk2(wa.y, wa.y), h[1], v);
                            v = ffma2(pk2(wb.x, wb.x), h[2], v);
                            v = ffma2(pk2(wb.y, wb.y), h[3], v);
                        } else {
                            v = ffma2(w23, h[1], fmul2(w01, h[0]));      // (slots 0 + 2, slots 1 + 3): summed at the end
                        }
                        bfly(p, v);
                    }
                };

                int cur_col = -1, hi_g = 0;               // groups [.., hi_g) of column cur_col are in registers
                for (int p0 = 0; p0 < nb;) {
                    const unsigned rest = starts & ~((2u << p0) - 1u);       // run starts after p0
                    const int p1 = rest ? __ffs(rest) - 1 : nb;
                    const int key = __shfl_sync(0xffffffffu, mykey, p0);
                    const int col = key >> 8, zb = key & 0xff;
                    if (col != cur_col) { cur_col = col; hi_g = 0; }
                    {
                        const Cell *cb = tile + (4 * (col >> 4)) * Sx + 4 * (col & 15);
                        for (int g = max(hi_g, zb); g < zb + 3; ++g) load_group(cb, g);
                        hi_g = zb + 3;
                    }
                    const int r3 = zb % 3;
                    if (r3 == 0) run(std::integral_constant<int, 0>{}, p0, p1);
                    else if (r3 == 1) run(std::integral_constant<int, 1>{}, p0, p1);
                    else run(std::integral_constant<int, 2>{}, p0, p1);
                    p0 = p1;
                }
                for (int p = nb; p < BATCH; ++p) bfly(p, 0ull);      // partial batch: complete the transposition
                // ---- store -----------------------------------------------------------------------------------------
                if (has_out) {
                    const float2 rv = unpk2(res);
                    float sc = prefactor;
                    if (nu_weights) sc *= nu_weights[n_out];
                    if constexpr (CPLX) store_value<float, true>(vp.p[c], n_out, make_float2(rv.x * sc, rv.y * sc));
                    else store_value<float, false>(vp.p[c], n_out, (rv.x + rv.y) * sc);
                }
                __syncwarp();
                bi = bn;
            }
            __syncthreads();                 // every warp is done with the tile (and with s_batch)
            if (tid == 0) s_batch = 0;
        }
    }
}

inline size_t interp_smem_bytes(const TileGeom &g, int cs_stride, size_t cell_bytes)
{
    size_t b = ((size_t)g.tile_cells * cell_bytes + 15) & ~(size_t)15;
    b += (size_t)INTERP_NW * BATCH * REC_F * sizeof(float);
    b += (size_t)(3 * cs_stride + 4) * sizeof(float);
    return b + 16;
}

}  // namespace rt
}  // namespace nufft
