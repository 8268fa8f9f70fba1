// rt_spread.cuh — K-spread, register-window variant (3-D, HalfSupport(4), Float32).  See rt_common.cuh for the idea.
//
// Replaces src/spreading/gpu.jl:237-434 for this configuration class (same sums, different order).
//
// CTA = 4 independent warps sharing one tile (the bin's padded subgrid in shared memory), persistent, pulling
// (bin, chunk) work items from a device counter.  Inside an item the warps pull 32-point batches from a shared-memory
// counter; a batch never leaves its warp (CTA barriers only around the tile flush):
//   evaluate   one thread per point: x / y kernel values zero-padded to the footprint of the point's 4 x 4 x 4-cell
//              sub-bin (11 cells per dimension), z values likewise and multiplied by the point's value -> the warp's
//              private record buffer; the (column, z block) key stays in the lane's register;
//   accumulate lane L keeps its 4 footprint cells of 12 planes in REGISTERS: three groups of 4 planes, group g = planes
//              4g .. 4g + 3 of the column in register group g % 3.  Points arrive ordered by (column, z): a run of points
//              of the same (column, z block) is a branch-free loop of 9 shared-memory loads and 46 FFMA2 / FMUL2
//              (packed re/im) per point and touches no tile memory;
//   retire     when the window moves on, the group that falls out is added to the tile.  Warps work on different
//              columns whose footprints overlap, so the add is a 64-bit compare-and-swap loop on the (re, im) cell
//              (shared memory has no native floating-point atomics); it runs once per (column, plane), not per point;
//   flush      tile -> oversampled grid with red.global.add.v4.f32 (2 complex cells); every lane owns a fixed set of
//              16-byte vectors of the flattened (x, y) plane, so a plane costs 9 loads + reductions per lane.
#pragma once
#include <type_traits>
#include "rt_common.cuh"
#include "spread.cuh"

namespace nufft {
namespace rt {

constexpr int SPREAD_NWARP = 4;
// record (floats): [0..23] value x wz_pad (complex: 11 x (re, im) + pad; real: 11 + pad), [24..47] wyT, [48..59] wx_pad
constexpr int SREC_F = RT_SREC_F;
constexpr int SOFF_S = 0, SOFF_WY = 24, SOFF_WX = 48;
constexpr int FLJ = 10;            // 16-byte vectors per lane and tile plane in the flush (23 rows x <= 13 vectors / 32)

// cell += v in shared memory, safe against concurrent adders
__device__ __forceinline__ void smem_add(float2 *cell, u64 v)
{
    unsigned long long *a = reinterpret_cast<unsigned long long *>(cell);
    u64 old = *a;
    while (true) {
        const u64 prev = atomicCAS(a, old, fadd2(old, v));
        if (prev == old) break;
        old = prev;
    }
}
__device__ __forceinline__ void smem_add(float *cell, float v)
{
    int *a = reinterpret_cast<int *>(cell);
    int old = *a;
    while (true) {
        const int prev = atomicCAS(a, old, __float_as_int(__int_as_float(old) + v));
        if (prev == old) break;
        old = prev;
    }
}

template <bool CPLX>
__global__ void __launch_bounds__(32 * SPREAD_NWARP, 2)
rt_spread_kernel(KernelParams<float> kp, TileGeom g, SmArgs a, const float *__restrict__ xs0, const float *__restrict__ xs1,
                 const float *__restrict__ xs2, PtrPack vp, int C, typename CellOf<float, CPLX>::type *__restrict__ us,
                 int64_t ncells, const float *__restrict__ nu_weights)
{
    using Cell = typename CellOf<float, CPLX>::type;
    constexpr int NWARP = SPREAD_NWARP;
    constexpr int NT = 32 * NWARP;
    constexpr int VEC = FlushVec<Cell>::VEC;
    constexpr int NG = CPLX ? 4 : 2;               // u64 registers per plane: complex (re, im) per slot; real (slot k, k + 1)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_bytes = (g.tile_cells * (int)sizeof(Cell) + 15) & ~15;
    Cell *tile = (Cell *)smem_raw;
    float *rec_all = (float *)(smem_raw + tile_bytes);                // [NWARP][BATCH][SREC_F]
    float *cs_s = rec_all + NWARP * BATCH * SREC_F;                   // [3][cs_stride]
    __shared__ int s_item[2][4];
    __shared__ int s_batch;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Tx = g.T[0], Ty = g.T[1], Tz = g.T[2], Sx = g.S[0], S2 = g.S[2];
    const int total_items = a.item_start[a.nbins];
    float *rec_w = rec_all + warp * BATCH * SREC_F;

    const LaneSlots ls = lane_slots(lane);
    const int off0 = ls.g * Sx + ls.x, off3 = ls.y3 * Sx + ls.x3;     // slot k < 3: off0 + 3 k Sx

    for (int i = tid; i < 3 * kp.cs_stride; i += NT) cs_s[i] = kp.cs[i];
    {
        const Cell zero = cell_zero((Cell *)nullptr);
        for (int i = tid; i < g.tile_cells; i += NT) tile[i] = zero;
    }
    if (tid == 0) {
        const int item = atomicAdd(a.work_counter, 1);
        s_item[0][0] = item;
        if (item < total_items) decode_item(a, item, g.chunk, s_item[0][1], s_item[0][2], s_item[0][3]);
        s_batch = 0;
    }
    __syncthreads();

    // register window: three groups of four planes
    u64 G[3][4][NG];
#pragma unroll
    for (int g3 = 0; g3 < 3; ++g3)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < NG; ++k) G[g3][i][k] = 0ull;

    for (int it = 0;; ++it) {
        const int *cur = s_item[it & 1];
        if (cur[0] >= total_items) break;
        const int bin = cur[1], k0 = cur[2], k1 = cur[3];
        if (tid == NT - 1) {                      // prefetch the next work item (read after the barriers below)
            int *nxt = s_item[(it + 1) & 1];
            const int item = atomicAdd(a.work_counter, 1);
            nxt[0] = item;
            if (item < total_items) decode_item(a, item, g.chunk, nxt[1], nxt[2], nxt[3]);
        }
        int b = bin;
        const int bx = b % g.nb[0]; b /= g.nb[0];
        const int by = b % g.nb[1]; b /= g.nb[1];
        const int bz = b;
        const int org0 = bx * g.B[0], org1 = by * g.B[1], org2 = bz * g.B[2];   // first cell of the bin
        const int nbatches = (k1 - k0 + BATCH - 1) / BATCH;

        for (int c = 0; c < C; ++c) {
            auto load_v = [&](int32_t n) -> Cell {
                Cell v = load_value<float, CPLX>(vp.p[c], n);
                if (nu_weights) v = cmul(v, nu_weights[n]);
                return v;
            };
            // ---- batches of this item, pulled by the warps; coordinates / values of the next batch are prefetched ----
            int bi = 0;
            if (lane == 0) bi = atomicAdd(&s_batch, 1);
            bi = __shfl_sync(0xffffffffu, bi, 0);
            float xq = 0.f, yq = 0.f, zq = 0.f;
            Cell vq = cell_zero((Cell *)nullptr);
            {
                const int k = k0 + bi * BATCH + lane;
                if (bi < nbatches && k < k1) { xq = xs0[k]; yq = xs1[k]; zq = xs2[k]; vq = load_v(a.perm[k]); }
            }
            while (bi < nbatches) {
                const int kb = k0 + bi * BATCH;
                const int nb = min(BATCH, k1 - kb);
                int bn = 0;
                if (lane == 0) bn = atomicAdd(&s_batch, 1);
                bn = __shfl_sync(0xffffffffu, bn, 0);
                const float x = xq, y = yq, z = zq;
                const Cell v = vq;
                {
                    const int k = k0 + bn * BATCH + lane;
                    if (bn < nbatches && k < k1) { xq = xs0[k]; yq = xs1[k]; zq = xs2[k]; vq = load_v(a.perm[k]); }
                }

                // ---- evaluate: one thread per point -----------------------------------------------------------
                int mykey = -1;                            // (column y << 12) | (column x << 8) | z block
                if (lane < nb) {
                    float *r = rec_w + lane * SREC_F;
                    float w[W], pw[P];
                    const int tx = eval_kernel_values<float, M>(kp, cs_s, 0, x, w) - org0;
                    pad_shift(w, tx & 3, pw);
                    float4 *q = reinterpret_cast<float4 *>(r + SOFF_WX);
                    q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
                    q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
                    q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
                    const int ty = eval_kernel_values<float, M>(kp, cs_s + kp.cs_stride, 1, y, w) - org1;
                    pad_shift(w, ty & 3, pw);
                    store_y(r + (SOFF_WY - OFF_WY), pw);
                    const int tz = eval_kernel_values<float, M>(kp, cs_s + 2 * kp.cs_stride, 2, z, w) - org2;
                    pad_shift(w, tz & 3, pw);
                    q = reinterpret_cast<float4 *>(r + SOFF_S);
                    if constexpr (CPLX) {
#pragma unroll
                        for (int i = 0; i < 5; ++i)
                            q[i] = make_float4(v.x * pw[2 * i], v.y * pw[2 * i], v.x * pw[2 * i + 1], v.y * pw[2 * i + 1]);
                        q[5] = make_float4(v.x * pw[10], v.y * pw[10], 0.f, 0.f);
                    } else {
                        q[0] = make_float4(v * pw[0], v * pw[1], v * pw[2], v * pw[3]);
                        q[1] = make_float4(v * pw[4], v * pw[5], v * pw[6], v * pw[7]);
                        q[2] = make_float4(v * pw[8], v * pw[9], v * pw[10], 0.f);
                    }
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
                const float *pY = rec_w + SOFF_WY + 4 * ls.row, *pX = rec_w + SOFF_WX + ls.x, *pX3 = rec_w + SOFF_WX + ls.x3;
                // add register group g % 3 (= planes 4g .. 4g + 3 of column base cb) to the tile and clear it
                auto retire_group = [&](Cell *cb, int gq) {
#define NUFFT_RT_RETIRE(R3)                                                                                   \
    _Pragma("unroll") for (int i = 0; i < 4; ++i) {                                                           \
        const int zp = 4 * gq + i;                                                                            \
        if (zp < Tz) {                                                                                        \
            Cell *pl = cb + zp * S2;                                                                          \
            if constexpr (CPLX) {                                                                             \
                smem_add(pl + off0, G[R3][i][0]); smem_add(pl + off0 + 3 * Sx, G[R3][i][1]);                  \
                smem_add(pl + off0 + 6 * Sx, G[R3][i][2]);                                                    \
                if (ls.has3) smem_add(pl + off3, G[R3][i][3]);                                                \
            } else {                                                                                          \
                const float2 ta = unpk2(G[R3][i][0]), tb = unpk2(G[R3][i][1]);                                \
                smem_add(pl + off0, ta.x); smem_add(pl + off0 + 3 * Sx, ta.y);                                \
                smem_add(pl + off0 + 6 * Sx, tb.x);                                                           \
                if (ls.has3) smem_add(pl + off3, tb.y);                                                       \
            }                                                                                                 \
        }                                                                                                     \
        _Pragma("unroll") for (int k = 0; k < NG; ++k) zero_inplace(G[R3][i][k]);                             \
    }
                    const int r3 = gq % 3;
                    if (r3 == 0) { NUFFT_RT_RETIRE(0) } else if (r3 == 1) { NUFFT_RT_RETIRE(1) } else { NUFFT_RT_RETIRE(2) }
#undef NUFFT_RT_RETIRE
                };
                // points [p0, p1) share the window: plane i of the window = register group (ROT + i / 4) % 3, plane i % 4
                auto run = [&](auto rot, int p0, int p1) {
                    constexpr int ROT = decltype(rot)::value;
                    for (int p = p0; p < p1; ++p) {
                        const float *r = rec_w + p * SREC_F;
                        const float4 wy = *reinterpret_cast<const float4 *>(pY + p * SREC_F);
                        const float wx = pX[p * SREC_F], wx3 = pX3[p * SREC_F];
                        const u64 w01 = fmul2(pk2(wx, wx), pk2(wy.x, wy.y));
                        const u64 w23 = fmul2(pk2(wx, wx3), pk2(wy.z, wy.w));
                        if constexpr (CPLX) {
                            const float2 wa = unpk2(w01), wb = unpk2(w23);
                            const float4 *sq = reinterpret_cast<const float4 *>(r + SOFF_S);
#pragma unroll
                            for (int h = 0; h < 6; ++h) {
                                const float4 s2 = sq[h];
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int i = 2 * h + e;
                                    if (i < 11) {
                                        const u64 sv = e == 0 ? pk2(s2.x, s2.y) : pk2(s2.z, s2.w);
                                        u64(&acc)[NG] = G[(ROT + i / 4) % 3][i % 4];
                                        acc[0] = ffma2(pk2(wa.x, wa.x), sv, acc[0]);
                                        acc[1] = ffma2(pk2(wa.y, wa.y), sv, acc[1]);
                                        acc[2] = ffma2(pk2(wb.x, wb.x), sv, acc[2]);
                                        acc[3] = ffma2(pk2(wb.y, wb.y), sv, acc[3]);
                                    }
                                }
                            }
                        } else {
                            const float4 *sq = reinterpret_cast<const float4 *>(r + SOFF_S);
                            const float4 sa = sq[0], sb = sq[1], sc = sq[2];
                            const float sv[11] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w, sc.x, sc.y, sc.z};
#pragma unroll
                            for (int i = 0; i < 11; ++i) {
                                u64(&acc)[NG] = G[(ROT + i / 4) % 3][i % 4];
                                acc[0] = ffma2(w01, pk2(sv[i], sv[i]), acc[0]);
                                acc[1] = ffma2(w23, pk2(sv[i], sv[i]), acc[1]);
                            }
                        }
                    }
                };

                int cur_col = -1, lo_g = 0;               // register groups hold groups [lo_g, lo_g + 3) of column cur_col
                Cell *cur_cb = tile;
                for (int p0 = 0; p0 < nb;) {
                    const unsigned rest = starts & ~((2u << p0) - 1u);       // run starts after p0
                    const int p1 = rest ? __ffs(rest) - 1 : nb;
                    const int key = __shfl_sync(0xffffffffu, mykey, p0);
                    const int col = key >> 8, zb = key & 0xff;
                    if (col != cur_col) {
                        if (cur_col >= 0)
                            for (int gq = lo_g; gq < lo_g + 3; ++gq) retire_group(cur_cb, gq);
                        cur_col = col;
                        cur_cb = tile + (4 * (col >> 4)) * Sx + 4 * (col & 15);
                        lo_g = zb;
                    } else if (zb > lo_g) {                // z only grows inside a column
                        for (int gq = lo_g; gq < min(zb, lo_g + 3); ++gq) retire_group(cur_cb, gq);
                        lo_g = zb;
                    }
                    const int r3 = zb % 3;
                    if (r3 == 0) run(std::integral_constant<int, 0>{}, p0, p1);
                    else if (r3 == 1) run(std::integral_constant<int, 1>{}, p0, p1);
                    else run(std::integral_constant<int, 2>{}, p0, p1);
                    p0 = p1;
                }
                if (cur_col >= 0)
                    for (int gq = lo_g; gq < lo_g + 3; ++gq) retire_group(cur_cb, gq);
                __syncwarp();
                bi = bn;
            }
            __syncthreads();
            if (tid == 0) s_batch = 0;             // every warp is done pulling batches; ordered by the barrier after the flush
            // ---- flush: tile -> global grid (periodic), vector reductions; re-zero the tile ------------
            {
                Cell *u = us + (int64_t)c * ncells;
                const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];
                const int x0 = org0 - (M - 1), y0 = org1 - (M - 1), z0 = org2 - (M - 1);
                const Cell zero = cell_zero((Cell *)nullptr);
                const bool vec_ok = (Nx % VEC) == 0 && VEC > 1;
                const int a0 = vec_ok ? pmod(x0, VEC) : 0;     // tile x of vector q starts at VEC * q - a0
                const int nvec = (Tx + a0 + VEC - 1) / VEC;
                if (vec_ok && Ty * nvec <= 32 * FLJ) {
                    // lane-owned vectors of the flattened (row, vector) plane: offsets are computed once per item
                    int soff[FLJ], goff[FLJ];
                    unsigned long long inmask = 0;             // VEC bits per vector: cell inside the tile row
#pragma unroll
                    for (int j = 0; j < FLJ; ++j) {
                        const int f = lane + 32 * j;
                        const int y = f / nvec, q = f - y * nvec;
                        const int xt = VEC * q - a0;
                        soff[j] = y * Sx + xt;
                        goff[j] = -1;
                        if (y < Ty) {
                            goff[j] = wrap1(y0 + y, Ny) * Nx + wrap1(x0 + xt, Nx);
#pragma unroll
                            for (int e = 0; e < VEC; ++e)
                                if (xt + e >= 0 && xt + e < Tx) inmask |= 1ull << (VEC * j + e);
                        }
                    }
                    for (int zp = warp; zp < Tz; zp += NWARP) {
                        Cell *gplane = u + (int64_t)wrap1(z0 + zp, Nz) * Ny * Nx;
                        Cell *tplane = tile + zp * S2;
#pragma unroll
                        for (int j = 0; j < FLJ; ++j) {
                            if (goff[j] >= 0) {
                                Cell val[VEC];
                                bool nz = false;
#pragma unroll
                                for (int e = 0; e < VEC; ++e) {
                                    val[e] = zero;
                                    if ((inmask >> (VEC * j + e)) & 1ull) { val[e] = tplane[soff[j] + e]; tplane[soff[j] + e] = zero; }
                                    nz = nz || cnonzero(val[e]);
                                }
                                if (nz) red_vec(gplane + goff[j], val);
                            }
                        }
                    }
                } else {
                    for (int zp = 0; zp < Tz; ++zp) {
                        const int gz = wrap1(z0 + zp, Nz);
                        for (int y = warp; y < Ty; y += NWARP) {
                            const int gy = wrap1(y0 + y, Ny);
                            Cell *grow = u + ((int64_t)gz * Ny + gy) * Nx;
                            Cell *trow = tile + zp * S2 + y * Sx;
                            for (int x = lane; x < Tx; x += 32) {
                                const Cell val = trow[x];
                                trow[x] = zero;
                                if (cnonzero(val)) catomic_add(grow + wrap1(x0 + x, Nx), val);
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

inline size_t spread_smem_bytes(const TileGeom &g, int cs_stride, size_t cell_bytes)
{
    size_t b = ((size_t)g.tile_cells * cell_bytes + 15) & ~(size_t)15;
    b += (size_t)SPREAD_NWARP * BATCH * SREC_F * sizeof(float);
    b += (size_t)(3 * cs_stride + 4) * sizeof(float);
    return b + 16;
}

}  // namespace rt
}  // namespace nufft
