// cs_interp.cuh — K-interp, column-streaming variant (3-D, HalfSupport(4), Float32 complex or real data).  Mirror image of
// cs_spread.cuh (real data: two consecutive z planes per packed register, as there);
// replaces src/interpolation/gpu.jl:211-395 for this configuration class (same sums, different order).
//
// A warp walks through a chunk of the points ordered by (z segment, column of 4 x 4 cells, layer of 4 cells) and keeps
// the padded footprint of the current layer (11 x 11 x 11 grid values) in registers, lane L holding the columns
// rt::lane_slots(L):
//   per point   8 shared-memory loads of the point's record (zero-padded 1-D weights), 44 FFMA2 (t_k = sum_i G[k][i] wz[i])
//               + 2 FMUL2 + 4 FFMA2 (sum_k t_k wx wy) per lane; the lane partial goes to a warp-private shared-memory
//               row (one STS.64) — no dependent shuffle chain in the loop;
//   per 16 pts  2 lanes per point add up the 32 partials of that point (16 LDS.64 each, rows padded to 34 -> no bank
//               conflicts) + 1 shuffle; prefactor, non-uniform callback, scatter through the permutation;
//   next layer  7 planes move down in the register file; the 4 new planes were requested from global memory (L2) when
//               the window arrived at the CURRENT layer (16 cp.async per lane into a warp-private staging buffer, so no
//               registers and no scoreboard shared with other loads), their latency is covered by the ~8 points of
//               the layer; 16 LDS.64 bring them in;
//   new column  44 loads per lane (rows of 11 consecutive cells).
#pragma once
#include "cs_spread.cuh"

namespace nufft {
namespace cs {

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m)
{
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
    const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_down_u64(u64 v, int d)
{
    const unsigned lo = __shfl_down_sync(0xffffffffu, (unsigned)v, d);
    const unsigned hi = __shfl_down_sync(0xffffffffu, (unsigned)(v >> 32), d);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_idx_u64(u64 v, int src)
{
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
    const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
constexpr int HALF = 16;                  // points per reduction round
constexpr int PART_LD = 34;               // row length of the lane-partial buffer (u64): 34 = 2 mod 16 -> lane 2p + h reads bank pair (lane + 2j) mod 16

__device__ __forceinline__ u64 ldg_cell(const float2 *p)
{
    u64 v;
    asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_cell(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <bool CPLX>                       // complex / real Float32 data (one instantiation per translation unit)
__global__ void __launch_bounds__(32 * NWARP)
cs_interp_kernel(KernelParams<float> kp, TileGeom g, int np, int chunk, const int32_t *__restrict__ perm, int32_t *work_counter,
                 const float4 *__restrict__ prec, MutPtrPack vp, int C,
                 const typename CellOf<float, CPLX>::type *__restrict__ us, int64_t ncells, float prefactor,
                 const float *__restrict__ nu_weights)
{
    using Cell = typename CellOf<float, CPLX>::type;
    constexpr int NR = Win<CPLX>::NR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *rec_all = (float *)smem_raw;                                      // [NWARP][BATCH][REC_F]
    float *stage_all = rec_all + NWARP * BATCH * REC_F;                      // [NWARP][STAGE_F] coordinates / index staging
    u64 *part_all = (u64 *)(stage_all + NWARP * STAGE_F);                    // [NWARP][HALF][PART_LD] lane partials
    u64 *hst_all = part_all + NWARP * HALF * PART_LD;                        // [NWARP][4 planes x 4 columns][32 lanes]
    float *cs_s = (float *)(hst_all + NWARP * 16 * 32);                      // [3][cs_stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    float *rec_w = rec_all + warp * BATCH * REC_F;
    u64 *part_w = part_all + warp * HALF * PART_LD;
    Cell *hst_w = reinterpret_cast<Cell *>(hst_all + warp * 16 * 32) + lane;
    float4 *st_x = reinterpret_cast<float4 *>(stage_all + warp * STAGE_F) + lane;      // folded (x, y, z, -) of the point
    int32_t *st_n = reinterpret_cast<int32_t *>(stage_all + warp * STAGE_F + 192) + lane;

    for (int i = tid; i < 3 * kp.cs_stride; i += 32 * NWARP) cs_s[i] = kp.cs[i];
    __syncthreads();                                   // the only CTA barrier: coefficient tables

    const rt::LaneSlots ls = rt::lane_slots(lane);
    const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];
    const int plane = Nx * Ny;
    const unsigned long long pol = l2_evict_first_policy();

    u64 G[4][NR];                                      // window: planes COL * wl - 3 .. COL * wl + 7
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < NR; ++i) G[k][i] = 0ull;

    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1);
        item = __shfl_sync(FULL, item, 0);
        const int64_t k0l = (int64_t)item * chunk;
        if (k0l >= np) break;
        const int k0 = (int)k0l, k1 = min(k0 + chunk, np);
        const int nbatches = (k1 - k0 + BATCH - 1) / BATCH;

        for (int c = 0; c < C; ++c) {
            Cell *vc = (Cell *)vp.p[c];
            const Cell *u = us + (int64_t)c * ncells;
            int wcol = -1, wl = 0;                         // column id (cy << 16 | cx), layer
            int goff[4] = {0, 0, 0, 0};                    // cell offsets of the lane's 4 columns inside a z plane

            auto request_ahead = [&]() {                   // planes P .. P + 3 relative to the window (next layer's new planes)
#pragma unroll
                for (int i = 0; i < COL; ++i) {
                    const Cell *pl = u + (int64_t)wrap1(COL * wl - (M - 1) + P + i, Nz) * plane;
#pragma unroll
                    for (int k = 0; k < 4; ++k) cp_async_cell<(int)sizeof(Cell)>(hst_w + (4 * i + k) * 32, pl + goff[k]);
                }
                cp_async_commit();
            };
            auto load_all = [&]() {
                cp_async_wait0();                          // a request of the previous window may be in flight
                if constexpr (CPLX) {
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const Cell *pl = u + (int64_t)wrap1(COL * wl - (M - 1) + i, Nz) * plane;
#pragma unroll
                        for (int k = 0; k < 4; ++k) G[k][i] = ldg_cell(pl + goff[k]);
                    }
                } else {                                   // planes 2j, 2j + 1 share register j; the 12th plane is padding
                    float f[4][P + 1];
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const Cell *pl = u + (int64_t)wrap1(COL * wl - (M - 1) + i, Nz) * plane;
#pragma unroll
                        for (int k = 0; k < 4; ++k) f[k][i] = ldg_cell(pl + goff[k]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        f[k][P] = 0.f;
#pragma unroll
                        for (int j = 0; j < NR; ++j) G[k][j] = pk2(f[k][2 * j], f[k][2 * j + 1]);
                    }
                }
                request_ahead();
            };
            auto shift_one = [&]() {                       // next layer: 7 planes slide down, the requested 4 come in
                cp_async_wait0();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (CPLX) {
#pragma unroll
                        for (int i = 0; i < P - COL; ++i) G[k][i] = G[k][i + COL];
#pragma unroll
                        for (int i = 0; i < COL; ++i) G[k][P - COL + i] = *reinterpret_cast<const u64 *>(hst_w + (4 * i + k) * 32);
                    } else {                               // old planes 4 .. 10 -> 0 .. 6; plane 6 shares register 3 with new plane 7
                        G[k][0] = G[k][2]; G[k][1] = G[k][3]; G[k][2] = G[k][4];
                        G[k][3] = pk2(unpk2(G[k][5]).x, hst_w[(4 * 0 + k) * 32]);
                        G[k][4] = pk2(hst_w[(4 * 1 + k) * 32], hst_w[(4 * 2 + k) * 32]);
                        G[k][5] = pk2(hst_w[(4 * 3 + k) * 32], 0.f);
                    }
                }
                ++wl;
                request_ahead();
            };

            // ---- global loads: one lane per point, staged through cp.async: the index two steps ahead, the folded
            //      coordinates (gathered through it from set_points' input-order records) one batch ahead ---------------------
            auto issue_n = [&](int bi) {
                const int k = k0 + bi * BATCH + lane;
                if (k < k1) cp_async_stream<4>(st_n, perm + k, pol);
            };
            auto issue_x = [&](int bi, int32_t n) {
                const int k = k0 + bi * BATCH + lane;
                if (k < k1) cp_async_stream<16>(st_x, prec + n, pol);
            };
            issue_n(0);
            cp_async_commit();
            cp_async_wait0();
            int32_t n_nxt = *st_n;                         // original index of this lane's point of the batch being staged
            issue_x(0, n_nxt);
            issue_n(1);
            cp_async_commit();

            for (int bi = 0; bi < nbatches; ++bi) {
                const int kb = k0 + bi * BATCH;
                const int nb = min(BATCH, k1 - kb);
                cp_async_wait0();
                const float4 xyz = *st_x;
                const float x = xyz.x, y = xyz.y, z = xyz.z;
                const int32_t n_mine = n_nxt;              // original index of point `lane` of the batch
                n_nxt = *st_n;
                issue_x(bi + 1, n_nxt);
                issue_n(bi + 2);
                cp_async_commit();

                // ---- evaluate: one lane per point ----------------------------------------------------------------------
                int mycol = -1, mylay = 0;                 // lane p < nb: column id and layer of point p of the batch
                if (lane < nb) {
                    int cx, cy, cz;
                    evaluate_point(kp, cs_s, x, y, z, rec_w + lane * REC_F, cx, cy, cz);
                    mycol = ((cy >> 2) << 16) | (cx >> 2);
                    mylay = cz >> 2;
                }
                // runs of points sharing the window: bit p of `starts` is set when point p opens a new (column, layer)
                unsigned starts;
                {
                    const int pc = __shfl_up_sync(FULL, mycol, 1), pl = __shfl_up_sync(FULL, mylay, 1);
                    starts = __ballot_sync(FULL, lane < nb && (lane == 0 || lane == HALF || mycol != pc || mylay != pl));
                }
                __syncwarp();

                for (int h0 = 0; h0 < nb; h0 += HALF) {    // two rounds of 16 points share the lane-partial buffer
                    const int h1 = min(h0 + HALF, nb);
                    for (int p0 = h0; p0 < h1;) {
                        const unsigned rest = starts & ~((2u << p0) - 1u);       // run starts after p0
                        const int p1 = rest ? min(__ffs(rest) - 1, h1) : h1;
                        const int col = __shfl_sync(FULL, mycol, p0), lay = __shfl_sync(FULL, mylay, p0);
                        if (col != wcol || lay != wl) {    // move the window (cold path)
                            const int d = lay - wl;
                            if (col == wcol && d == 1) {
                                shift_one();
                            } else {
                                wcol = col;
                                wl = lay;
                                const int cx = col & 0xffff, cy = col >> 16;
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                    goff[k] = wrap1(COL * cy - (M - 1) + ls.g + 3 * k, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                                goff[3] = wrap1(COL * cy - (M - 1) + ls.y3, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x3, Nx);
                                load_all();
                            }
                        }
                        // hot loop (rotated: the next record is requested at the end of the body)
                        PointRec A = load_rec(rec_w + p0 * REC_F, ls);
#pragma unroll 1
                        for (int p = p0; p < p1; ++p) {
                            u64 tk[4] = {0ull, 0ull, 0ull, 0ull};
                            if constexpr (CPLX) {
                                const float wz[P] = {A.z0.x, A.z0.y, A.z0.z, A.z0.w, A.z1.x, A.z1.y, A.z1.z, A.z1.w, A.z2.x, A.z2.y, A.z2.z};
#pragma unroll
                                for (int i = 0; i < P; ++i) {
                                    const u64 wzz = pk2(wz[i], wz[i]);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) tk[k] = ffma2(G[k][i], wzz, tk[k]);
                                }
                            } else {                       // (even planes, odd planes) partial sums; added up after the reduction
                                const u64 wz[6] = {pk2(A.z0.x, A.z0.y), pk2(A.z0.z, A.z0.w), pk2(A.z1.x, A.z1.y),
                                                   pk2(A.z1.z, A.z1.w), pk2(A.z2.x, A.z2.y), pk2(A.z2.z, A.z2.w)};      // z2.w = 0
#pragma unroll
                                for (int j = 0; j < 6; ++j)
#pragma unroll
                                    for (int k = 0; k < 4; ++k) tk[k] = ffma2(G[k][j], wz[j], tk[k]);
                            }
                            const u64 w01 = fmul2(pk2(A.wx, A.wx), pk2(A.wy.x, A.wy.y));
                            const u64 w23 = fmul2(pk2(A.wx, A.wx3), pk2(A.wy.z, A.wy.w));
                            const float2 wa = unpk2(w01), wb = unpk2(w23);
                            u64 acc = fmul2(tk[0], pk2(wa.x, wa.x));
                            acc = ffma2(tk[1], pk2(wa.y, wa.y), acc);
                            acc = ffma2(tk[2], pk2(wb.x, wb.x), acc);
                            acc = ffma2(tk[3], pk2(wb.y, wb.y), acc);
                            part_w[(p - h0) * PART_LD + lane] = acc;
                            A = load_rec(rec_w + min(p + 1, p1 - 1) * REC_F, ls);
                        }
                        p0 = p1;
                    }
                    __syncwarp();
                    // ---- lanes 2q, 2q + 1 add up the 32 lane partials of point h0 + q ---------------------------------------
                    {
                        const int q = lane >> 1, hh = lane & 1;
                        u64 sum = 0ull;
                        if (h0 + q < h1) {
                            const u64 *row = part_w + q * PART_LD + hh;
#pragma unroll
                            for (int j = 0; j < 16; ++j) sum = rt::fadd2(sum, row[2 * j]);
                        }
                        sum = rt::fadd2(sum, shfl_xor_u64(sum, 1));
                        // lane h0 + q' stores point h0 + q': fetch its sum from lane 2 q'
                        const u64 res = shfl_idx_u64(sum, (2 * (lane - h0)) & 31);
                        if (lane >= h0 && lane < h1) {
                            const float2 rv = unpk2(res);
                            const float scale = prefactor * (nu_weights ? nu_weights[n_mine] : 1.f);
                            if constexpr (CPLX) __stcs(vc + n_mine, make_float2(rv.x * scale, rv.y * scale));     // streaming store: written once
                            else __stcs(vc + n_mine, (rv.x + rv.y) * scale);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
}

inline size_t interp_smem_bytes(int cs_stride)
{
    return spread_smem_bytes(cs_stride) + (size_t)NWARP * (HALF * PART_LD + 16 * 32) * sizeof(u64);
}

}  // namespace cs
}  // namespace nufft
