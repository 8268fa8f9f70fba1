// rt_spread.cuh — K-spread, register-tile variant (3-D, HalfSupport(4), Float32).  See rt_common.cuh for the idea.
//
// Replaces src/spreading/gpu.jl:237-434 for this configuration class (same sums, different order).
//
// CTA = 4 consumer warps + NPROD producer warps, persistent, pulling (bin, chunk) work items from a device counter.
//   producers : three warps, one per DIMENSION, one thread per point of the 32-point batch — kernel values
//               (eval_kernel_values, same polynomials as every other path) zero-padded to the column footprint (x, y)
//               or pre-multiplied by the point's value (z) — written as a 208-byte record, double buffered (one CTA
//               barrier per batch).  Coordinates / permutation / values of the following batches are prefetched
//               into registers (the value gather vp[perm[k]] is two dependent loads), so the evaluation never
//               waits on global memory;
//   consumers : warp w owns the tile planes z = w (mod 4).  A point's 8 planes meet every class exactly twice, so
//               every consumer does the same work for every point (no imbalance, nothing shared between warps, no
//               shared-memory atomics).  Per point and warp: 4 shared-memory loads, 4 FMUL, 16 FFMA into the
//               register tile acc[6 planes][4 slots]; the tile in shared memory is only read-modified-written when
//               the COLUMN changes (~ once per 32 points at density 1/8 per cell), by its only owner.
//   flush     : tile -> oversampled grid with red.global.add.v4.f32 (2 complex cells), as spread_sm_kernel.
// Issue-bound (~27 warp instructions per point and consumer), not shared-memory-bandwidth-bound like the
// cell-per-lane read-modify-write formulation (512 cell updates = 8 KiB of shared-memory traffic per point).
#pragma once
#include "rt_common.cuh"
#include "spread.cuh"

namespace nufft {
namespace rt {

constexpr int SPREAD_NCONS = 4;
constexpr int SPREAD_NPROD = 3;        // producer warps: one per dimension

template <bool CPLX>
__global__ void __launch_bounds__(32 * (SPREAD_NCONS + SPREAD_NPROD), 2)
rt_spread_kernel(KernelParams<float> kp, TileGeom g, SmArgs a, const float *__restrict__ xs0, const float *__restrict__ xs1,
                 const float *__restrict__ xs2, PtrPack vp, int C, typename CellOf<float, CPLX>::type *__restrict__ us,
                 int64_t ncells, const float *__restrict__ nu_weights)
{
    using Cell = typename CellOf<float, CPLX>::type;
    constexpr int NWARP = SPREAD_NCONS + SPREAD_NPROD;
    constexpr int NT = 32 * NWARP;
    constexpr int VEC = FlushVec<Cell>::VEC;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_bytes = (g.tile_cells * (int)sizeof(Cell) + 15) & ~15;
    Cell *tile = (Cell *)smem_raw;
    float *rec_s = (float *)(smem_raw + tile_bytes);                  // [2][batch][REC_F]
    float *cs_s = rec_s + 2 * g.batch * REC_F;                        // [3][cs_stride]
    __shared__ int s_item[2][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool consumer = warp < SPREAD_NCONS;
    const int role = warp - SPREAD_NCONS;                              // producers: dimension evaluated by this warp
    const float *xs_r = role == 0 ? xs0 : (role == 1 ? xs1 : xs2);
    const int Tx = g.T[0], Ty = g.T[1], Tz = g.T[2], Sx = g.S[0], S2 = g.S[2];
    const int total_items = a.item_start[a.nbins];

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
    }
    __syncthreads();

    // register tile of this consumer warp: planes warp + 4 q, q = 0..5; re / im parts
    float ar[NPL][4], ai[NPL][4];
#pragma unroll
    for (int q = 0; q < NPL; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) { ar[q][k] = 0.f; ai[q][k] = 0.f; }

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
        const int nbatches = (k1 - k0 + g.batch - 1) / g.batch;

        for (int c = 0; c < C; ++c) {
            // ---- producers: warp `role` evaluates dimension `role` of the batch, one thread per point ------------
            // register prefetch: xq = coordinate of this lane's point in the NEXT batch to produce, vq = its value
            // (z warp), nq = permutation entry of the batch after that
            float xq = 0.f;
            Cell vq = cell_zero((Cell *)nullptr);
            int32_t nq = 0;
            auto load_v = [&](int32_t n) -> Cell {
                Cell v = load_value<float, CPLX>(vp.p[c], n);
                if (nu_weights) v = cmul(v, nu_weights[n]);
                return v;
            };
            if (!consumer) {
                const int k = k0 + lane;
                if (k < k1) xq = xs_r[k];
                if (role == 2) {
                    if (k < k1) vq = load_v(a.perm[k]);
                    if (k + 32 < k1) nq = a.perm[k + 32];
                }
            }
            auto produce = [&](int bi) {
                const int kb = k0 + bi * g.batch;          // g.batch == 32
                const int nb = min(g.batch, k1 - kb);
                float *r = rec_s + ((bi & 1) * g.batch + lane) * REC_F;
                const float x = xq;
                const Cell v = vq;
                {   // issue the loads of the following batches before evaluating this one
                    const int kn = kb + 32 + lane;
                    if (kn < k1) {
                        xq = xs_r[kn];
                        if (role == 2) vq = load_v(nq);
                    }
                    if (role == 2 && kn + 32 < k1) nq = a.perm[kn + 32];
                }
                if (lane < nb) {
                    float w[W], pw[P];
                    unsigned char *mb = reinterpret_cast<unsigned char *>(r + OFF_META);
                    if (role == 0) {
                        const int tx = eval_kernel_values<float, M>(kp, cs_s, 0, x, w) - org0;
                        pad_shift(w, tx & 3, pw);
                        float4 *q = reinterpret_cast<float4 *>(r + OFF_WX);
                        q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
                        q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
                        *reinterpret_cast<float2 *>(r + OFF_WX + 8) = make_float2(pw[8], pw[9]);
                        r[OFF_WX + 10] = pw[10];
                        mb[0] = (unsigned char)(tx >> 2);
                    } else if (role == 1) {
                        const int ty = eval_kernel_values<float, M>(kp, cs_s + kp.cs_stride, 1, x, w) - org1;
                        pad_shift(w, ty & 3, pw);
                        store_y(r, pw);
                        mb[1] = (unsigned char)(ty >> 2);
                    } else {
                        const int tz = eval_kernel_values<float, M>(kp, cs_s + 2 * kp.cs_stride, 2, x, w) - org2;
                        // z weights by residue class of the tile plane: planes tz + j, j = 0..7; class (tz + j) & 3
                        float vr, vi;
                        if constexpr (CPLX) { vr = v.x; vi = v.y; } else { vr = v; vi = 0.f; }
                        float4 *s4 = reinterpret_cast<float4 *>(r + OFF_S);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            s4[(tz + j) & 3] = make_float4(vr * w[j], vi * w[j], vr * w[j + 4], vi * w[j + 4]);
                        mb[2] = (unsigned char)tz;
                        mb[3] = 0;
                    }
                }
            };
            // ---- consumer: add the register tile of column `col` to the shared-memory tile, clear it --------
            auto flush_col = [&](int col) {
                Cell *base = tile + (4 * (col >> 8)) * Sx + 4 * (col & 0xff) + warp * S2;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    if (warp + 4 * q < Tz) {
                        Cell *pl = base + 4 * q * S2;
                        Cell t0 = pl[off0], t1 = pl[off0 + 3 * Sx], t2 = pl[off0 + 6 * Sx], t3 = pl[off3];
                        if constexpr (CPLX) {
                            t0.x += ar[q][0]; t0.y += ai[q][0]; t1.x += ar[q][1]; t1.y += ai[q][1];
                            t2.x += ar[q][2]; t2.y += ai[q][2]; t3.x += ar[q][3]; t3.y += ai[q][3];
                        } else {
                            t0 += ar[q][0]; t1 += ar[q][1]; t2 += ar[q][2]; t3 += ar[q][3];
                        }
                        pl[off0] = t0; pl[off0 + 3 * Sx] = t1; pl[off0 + 6 * Sx] = t2;
                        if (ls.has3) pl[off3] = t3;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) { ar[q][k] = 0.f; ai[q][k] = 0.f; }
                }
            };

            if (!consumer) produce(0);
            __syncthreads();
            int cur_col = -1;
            for (int bi = 0; bi < nbatches; ++bi) {
                if (!consumer) {
                    if (bi + 1 < nbatches) produce(bi + 1);
                } else {
                    const int nb = min(g.batch, k1 - (k0 + bi * g.batch));
                    const float *r = rec_s + (bi & 1) * g.batch * REC_F;
                    // software pipeline: the record of point p + 1 is loaded while point p is accumulated
                    float4 s = *reinterpret_cast<const float4 *>(r + OFF_S + 4 * warp);
                    float4 wy = *reinterpret_cast<const float4 *>(r + OFF_WY + 4 * ls.row);
                    float wx = r[OFF_WX + ls.x], wx3 = r[OFF_WX + ls.x3];
                    int meta = __float_as_int(r[OFF_META]);
                    for (int p = 0; p < nb; ++p) {
                        const float *rn = r + min(p + 1, nb - 1) * REC_F;
                        const float4 s_n = *reinterpret_cast<const float4 *>(rn + OFF_S + 4 * warp);
                        const float4 wy_n = *reinterpret_cast<const float4 *>(rn + OFF_WY + 4 * ls.row);
                        const float wx_n = rn[OFF_WX + ls.x], wx3_n = rn[OFF_WX + ls.x3];
                        const int meta_n = __float_as_int(rn[OFF_META]);

                        const int col = meta & 0xffff, tz = (meta >> 16) & 0xff;
                        if (col != cur_col) {
                            if (cur_col >= 0) flush_col(cur_col);
                            cur_col = col;
                        }
                        const float w0 = wx * wy.x, w1 = wx * wy.y, w2 = wx * wy.z, w3 = wx3 * wy.w;
                        const int qa = (tz - warp + 3) >> 2;          // first owned plane of the support = warp + 4 qa
#define NUFFT_RT_ACC(Q)                                                                                      \
    ar[Q][0] = fmaf(w0, s.x, ar[Q][0]); ar[Q][1] = fmaf(w1, s.x, ar[Q][1]);                                  \
    ar[Q][2] = fmaf(w2, s.x, ar[Q][2]); ar[Q][3] = fmaf(w3, s.x, ar[Q][3]);                                  \
    ar[Q + 1][0] = fmaf(w0, s.z, ar[Q + 1][0]); ar[Q + 1][1] = fmaf(w1, s.z, ar[Q + 1][1]);                  \
    ar[Q + 1][2] = fmaf(w2, s.z, ar[Q + 1][2]); ar[Q + 1][3] = fmaf(w3, s.z, ar[Q + 1][3]);                  \
    if constexpr (CPLX) {                                                                                    \
        ai[Q][0] = fmaf(w0, s.y, ai[Q][0]); ai[Q][1] = fmaf(w1, s.y, ai[Q][1]);                              \
        ai[Q][2] = fmaf(w2, s.y, ai[Q][2]); ai[Q][3] = fmaf(w3, s.y, ai[Q][3]);                              \
        ai[Q + 1][0] = fmaf(w0, s.w, ai[Q + 1][0]); ai[Q + 1][1] = fmaf(w1, s.w, ai[Q + 1][1]);              \
        ai[Q + 1][2] = fmaf(w2, s.w, ai[Q + 1][2]); ai[Q + 1][3] = fmaf(w3, s.w, ai[Q + 1][3]);              \
    }
                        switch (qa) {
                        case 0: { NUFFT_RT_ACC(0) } break;
                        case 1: { NUFFT_RT_ACC(1) } break;
                        case 2: { NUFFT_RT_ACC(2) } break;
                        case 3: { NUFFT_RT_ACC(3) } break;
                        default: { NUFFT_RT_ACC(4) } break;
                        }
#undef NUFFT_RT_ACC
                        s = s_n; wy = wy_n; wx = wx_n; wx3 = wx3_n; meta = meta_n;
                    }
                }
                __syncthreads();
            }
            if (consumer && cur_col >= 0) flush_col(cur_col);
            __syncthreads();
            // ---- flush: tile -> global grid (periodic), vector reductions; re-zero the tile ------------
            {
                Cell *u = us + (int64_t)c * ncells;
                const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];
                const int x0 = org0 - (M - 1), y0 = org1 - (M - 1), z0 = org2 - (M - 1);
                const Cell zero = cell_zero((Cell *)nullptr);
                const bool vec_ok = (Nx % VEC) == 0;
                const int hl = lane & 15, hrow = lane >> 4;
                if (vec_ok) {
                    const int a0 = pmod(x0, VEC);                  // tile x of vector q starts at VEC * q - a0
                    const int nvec = (Tx + a0 + VEC - 1) / VEC;
                    for (int qb = 0; qb < nvec; qb += 16) {
                        const int q = qb + hl;
                        const int xt = VEC * q - a0;
                        const int gx = wrap1(x0 + xt, Nx);
                        bool in[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) in[e] = q < nvec && xt + e >= 0 && xt + e < Tx;
                        for (int z = 0; z < Tz; ++z) {
                            const int gz = wrap1(z0 + z, Nz);
                            Cell *gplane = u + (int64_t)gz * Ny * Nx + gx;
                            Cell *tplane = tile + z * S2 + xt;
                            for (int y = 2 * warp + hrow; y < Ty; y += 2 * NWARP) {
                                const int gy = wrap1(y0 + y, Ny);
                                Cell *trow = tplane + y * Sx;
                                Cell val[VEC];
                                bool nz = false;
#pragma unroll
                                for (int e = 0; e < VEC; ++e) {
                                    val[e] = zero;
                                    if (in[e]) { val[e] = trow[e]; trow[e] = zero; }
                                    nz = nz || cnonzero(val[e]);
                                }
                                if (nz) red_vec(gplane + (int64_t)gy * Nx, val);
                            }
                        }
                    }
                } else {
                    for (int z = 0; z < Tz; ++z) {
                        const int gz = wrap1(z0 + z, Nz);
                        for (int y = warp; y < Ty; y += NWARP) {
                            const int gy = wrap1(y0 + y, Ny);
                            Cell *grow = u + ((int64_t)gz * Ny + gy) * Nx;
                            Cell *trow = tile + z * S2 + y * Sx;
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
    b += (size_t)2 * g.batch * REC_F * sizeof(float);
    b += (size_t)(3 * cs_stride + 4) * sizeof(float);
    return b + 16;
}

}  // namespace rt
}  // namespace nufft
