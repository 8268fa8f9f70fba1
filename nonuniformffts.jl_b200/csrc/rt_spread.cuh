// rt_spread.cuh — K-spread, register-window variant (3-D, HalfSupport(4), Float32).  See rt_common.cuh for the idea.
//
// Replaces src/spreading/gpu.jl:237-434 for this configuration class (same sums, different order).
//
// CTA = 4 warps, persistent, pulling (bin, chunk) work items from a device counter.  set_points has ordered the points
// of a bin by (4 x 4-cell column, z cell), so consecutive points share their padded (x, y) footprint of 11 x 11 cells
// and move monotonically along z.
//   evaluate  : every warp first evaluates ONE ingredient of the next 32-point batch, one thread per point —
//               warp 0 the x weights, warp 1 the y weights (both zero-padded to the column footprint), warp 2 the z
//               weights times the point's value, warp 3 the column / plane keys — into a double-buffered record
//               (one CTA barrier per batch; coordinates, permutation and values are prefetched one batch ahead);
//   accumulate: warp w owns the tile planes z = w (mod 4).  A point's 8 planes meet every class exactly twice, so
//               every warp does the same work for every point (no imbalance, one owner per cell, no shared-memory
//               atomics).  The two planes live in REGISTERS (plane q of the class in the even / odd accumulator set by
//               parity of q): per point and warp 5 shared-memory loads, 2 FMUL2 and 8 FFMA2 (packed re/im) and no
//               tile traffic.  Because z only grows inside a column, a plane is read-modified-written to the tile
//               exactly once per column, when the window moves past it;
//   flush     : tile -> oversampled grid with red.global.add.v4.f32 (2 complex cells), as spread_sm_kernel.
#pragma once
#include "rt_common.cuh"
#include "spread.cuh"

namespace nufft {
namespace rt {

constexpr int SPREAD_NWARP = 4;

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

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_bytes = (g.tile_cells * (int)sizeof(Cell) + 15) & ~15;
    Cell *tile = (Cell *)smem_raw;
    float *rec_s = (float *)(smem_raw + tile_bytes);                  // [2][BATCH][REC_F]
    int *key_s = (int *)(rec_s + 2 * BATCH * REC_F);                  // [2][BATCH][4]
    float *cs_s = (float *)(key_s + 2 * BATCH * 4);                   // [3][cs_stride]
    __shared__ int s_item[2][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

    // register window of this warp: even / odd plane of its class; CPLX: (re, im) per slot, real: (even, odd) per slot
    u64 accE[4], accO[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { accE[k] = 0ull; accO[k] = 0ull; }

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
            // ---- evaluation role of this warp (one thread per point of the batch) ------------------------------
            // register prefetch: coordinates of this lane's point in the NEXT batch to evaluate; warp 2 also its value
            // (vp[perm[k]] is two dependent loads: the permutation entry is fetched one batch earlier still)
            float xq = 0.f, yq = 0.f, zq = 0.f;
            Cell vq = cell_zero((Cell *)nullptr);
            int32_t nq = 0;
            auto load_v = [&](int32_t n) -> Cell {
                Cell v = load_value<float, CPLX>(vp.p[c], n);
                if (nu_weights) v = cmul(v, nu_weights[n]);
                return v;
            };
            auto prefetch = [&](int k) {
                if (k < k1) {
                    if (warp == 0 || warp == 3) xq = xs0[k];
                    if (warp == 1 || warp == 3) yq = xs1[k];
                    if (warp >= 2) zq = xs2[k];
                    if (warp == 2) vq = load_v(nq);
                }
                if (warp == 2 && k + BATCH < k1) nq = a.perm[k + BATCH];
            };
            if (warp == 2 && k0 + lane < k1) nq = a.perm[k0 + lane];
            prefetch(k0 + lane);

            auto produce = [&](int bi) {
                const int kb = k0 + bi * BATCH;
                const int nb = min(BATCH, k1 - kb);
                float *r = rec_s + ((bi & 1) * BATCH + lane) * REC_F;
                const float x = xq, y = yq, z = zq;
                const Cell v = vq;
                prefetch(kb + BATCH + lane);          // loads of the following batch fly during this evaluation
                if (lane < nb) {
                    float w[W], pw[P];
                    if (warp == 0) {
                        const int tx = eval_kernel_values<float, M>(kp, cs_s, 0, x, w) - org0;
                        pad_shift(w, tx & 3, pw);
                        float4 *q = reinterpret_cast<float4 *>(r + OFF_WX);
                        q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
                        q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
                        q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
                    } else if (warp == 1) {
                        const int ty = eval_kernel_values<float, M>(kp, cs_s + kp.cs_stride, 1, y, w) - org1;
                        pad_shift(w, ty & 3, pw);
                        store_y(r, pw);
                    } else if (warp == 2) {
                        const int tz = eval_kernel_values<float, M>(kp, cs_s + 2 * kp.cs_stride, 2, z, w) - org2;
                        // planes tz + j, j = 0..7: class (tz + j) & 3 owns j and j + 4; the one whose class-plane
                        // index q = (tz + j) >> 2 is even goes to the even accumulator
                        float4 *s4 = reinterpret_cast<float4 *>(r + OFF_S);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int t = tz + j;
                            const bool odd = ((t >> 2) & 1) != 0;
                            const float we = odd ? w[j + 4] : w[j], wo = odd ? w[j] : w[j + 4];
                            if constexpr (CPLX) s4[t & 3] = make_float4(v.x * we, v.y * we, v.x * wo, v.y * wo);
                            else s4[t & 3] = make_float4(v * we, v * wo, 0.f, 0.f);
                        }
                    } else {
                        float rr;
                        const int tx = point_to_cell0<float>(x, kp.N[0], rr) - org0;
                        const int ty = point_to_cell0<float>(y, kp.N[1], rr) - org1;
                        const int tz = point_to_cell0<float>(z, kp.N[2], rr) - org2;
                        const int col = (((ty >> 2) << 4) | (tx >> 2)) << 8;
                        // first plane index of class c touched by the support: q = ceil((tz - c) / 4)
                        reinterpret_cast<int4 *>(key_s)[(bi & 1) * BATCH + lane] =
                            make_int4(col | ((tz + 3) >> 2), col | ((tz + 2) >> 2), col | ((tz + 1) >> 2), col | (tz >> 2));
                    }
                }
            };

            // ---- accumulate: add the even / odd register plane to the tile, clear it ---------------------------
            auto retire = [&](u64 (&acc)[4], int col, int q) {
                const int t = warp + 4 * q;
                if (t < Tz) {
                    Cell *pl = tile + (4 * ((col >> 4) & 15)) * Sx + 4 * (col & 15) + t * S2;
                    if constexpr (CPLX) {
                        float2 t0 = pl[off0], t1 = pl[off0 + 3 * Sx], t2 = pl[off0 + 6 * Sx], t3 = pl[off3];
                        const float2 a0 = unpk2(acc[0]), a1 = unpk2(acc[1]), a2 = unpk2(acc[2]), a3 = unpk2(acc[3]);
                        t0.x += a0.x; t0.y += a0.y; t1.x += a1.x; t1.y += a1.y;
                        t2.x += a2.x; t2.y += a2.y; t3.x += a3.x; t3.y += a3.y;
                        pl[off0] = t0; pl[off0 + 3 * Sx] = t1; pl[off0 + 6 * Sx] = t2;
                        if (ls.has3) pl[off3] = t3;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) zero_inplace(acc[k]);
            };
            // real data: both planes share the accumulators (even, odd) -> one combined retire
            auto retire_real = [&](int col, int cqa, bool doE, bool doO) {
                if constexpr (!CPLX) {
                    const int qe = (cqa + 1) & ~1, qo = cqa | 1;
                    Cell *base = tile + (4 * ((col >> 4) & 15)) * Sx + 4 * (col & 15);
                    float e[4], o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const float2 t = unpk2(accE[k]); e[k] = t.x; o[k] = t.y; }
                    const int te = warp + 4 * qe, to = warp + 4 * qo;
                    if (doE && te < Tz) {
                        Cell *pl = base + te * S2;
                        pl[off0] += e[0]; pl[off0 + 3 * Sx] += e[1]; pl[off0 + 6 * Sx] += e[2];
                        if (ls.has3) pl[off3] += e[3];
                    }
                    if (doO && to < Tz) {
                        Cell *pl = base + to * S2;
                        pl[off0] += o[0]; pl[off0 + 3 * Sx] += o[1]; pl[off0 + 6 * Sx] += o[2];
                        if (ls.has3) pl[off3] += o[3];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) accE[k] = pk2(doE ? 0.f : e[k], doO ? 0.f : o[k]);
                }
            };
            int cur_key = -1;
            auto window_move = [&](int key) {          // key < 0: retire everything
                if (cur_key >= 0) {
                    const int ccol = cur_key >> 8, cqa = cur_key & 0xff;
                    const bool all = key < 0 || (key >> 8) != ccol || (key & 0xff) - cqa >= 2;
                    const bool doE = all || !(cqa & 1), doO = all || (cqa & 1);
                    if constexpr (CPLX) {
                        if (doE) retire(accE, ccol, (cqa + 1) & ~1);
                        if (doO) retire(accO, ccol, cqa | 1);
                    } else {
                        retire_real(ccol, cqa, doE, doO);
                    }
                }
                cur_key = key;
            };

            produce(0);
            __syncthreads();
            for (int bi = 0; bi < nbatches; ++bi) {
                if (bi + 1 < nbatches) produce(bi + 1);
                {
                    const int nb = min(BATCH, k1 - (k0 + bi * BATCH));
                    const float *r = rec_s + (bi & 1) * BATCH * REC_F;
                    const float *pS = r + OFF_S + 4 * warp, *pY = r + OFF_WY + 4 * ls.row;
                    const float *pX = r + OFF_WX + ls.x, *pX3 = r + OFF_WX + ls.x3;
                    const int *pK = key_s + (bi & 1) * BATCH * 4 + warp;
                    // software pipeline: the record of point p + 1 is loaded while point p is accumulated
                    float4 s = *reinterpret_cast<const float4 *>(pS);
                    float4 wy = *reinterpret_cast<const float4 *>(pY);
                    float wx = *pX, wx3 = *pX3;
                    int key = *pK;
#pragma unroll 2
                    for (int p = 0; p < nb; ++p) {
                        const int on = min(p + 1, nb - 1);
                        const float4 s_n = *reinterpret_cast<const float4 *>(pS + on * REC_F);
                        const float4 wy_n = *reinterpret_cast<const float4 *>(pY + on * REC_F);
                        const float wx_n = pX[on * REC_F], wx3_n = pX3[on * REC_F];
                        const int key_n = pK[on * 4];

                        if (key != cur_key) window_move(key);
                        const u64 w01 = fmul2(pk2(wx, wx), pk2(wy.x, wy.y));
                        const u64 w23 = fmul2(pk2(wx, wx3), pk2(wy.z, wy.w));
                        const float2 wa = unpk2(w01), wb = unpk2(w23);
                        if constexpr (CPLX) {
                            const u64 sE = pk2(s.x, s.y), sO = pk2(s.z, s.w);
                            accE[0] = ffma2(pk2(wa.x, wa.x), sE, accE[0]);
                            accE[1] = ffma2(pk2(wa.y, wa.y), sE, accE[1]);
                            accE[2] = ffma2(pk2(wb.x, wb.x), sE, accE[2]);
                            accE[3] = ffma2(pk2(wb.y, wb.y), sE, accE[3]);
                            accO[0] = ffma2(pk2(wa.x, wa.x), sO, accO[0]);
                            accO[1] = ffma2(pk2(wa.y, wa.y), sO, accO[1]);
                            accO[2] = ffma2(pk2(wb.x, wb.x), sO, accO[2]);
                            accO[3] = ffma2(pk2(wb.y, wb.y), sO, accO[3]);
                        } else {
                            const u64 sEO = pk2(s.x, s.y);
                            accE[0] = ffma2(pk2(wa.x, wa.x), sEO, accE[0]);
                            accE[1] = ffma2(pk2(wa.y, wa.y), sEO, accE[1]);
                            accE[2] = ffma2(pk2(wb.x, wb.x), sEO, accE[2]);
                            accE[3] = ffma2(pk2(wb.y, wb.y), sEO, accE[3]);
                        }
                        s = s_n; wy = wy_n; wx = wx_n; wx3 = wx3_n; key = key_n;
                    }
                }
                __syncthreads();
            }
            window_move(-1);
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
    b += (size_t)2 * BATCH * REC_F * sizeof(float);
    b += (size_t)2 * BATCH * 4 * sizeof(int);
    b += (size_t)(3 * cs_stride + 4) * sizeof(float);
    return b + 16;
}

}  // namespace rt
}  // namespace nufft
