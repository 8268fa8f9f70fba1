// rt_interp.cuh — K-interp, register-tile variant (3-D, HalfSupport(4), Float32).  See rt_common.cuh for the idea.
//
// Replaces src/interpolation/gpu.jl:211-395 for this configuration class (same sums, different order).
//
// Persistent CTA of NW independent warps; (bin, chunk) work items from a device counter.  The bin's padded tile is
// staged in shared memory with cp.async, DOUBLE BUFFERED: the tile of the next work item is in flight while the
// current one is processed.  Each warp then takes sub-bins of 4 x 4 x ZB cells (ZB = 8: a column's z blocks 2h, 2h+1;
// set_points sorted the points by (column, z block)), loads the sub-bin's padded footprint 11 x 11 x (ZB + 7) into
// REGISTERS once (4 slots per lane and plane), and every point of the sub-bin is a register dot product:
//   * all 32 lanes evaluate one zero-padded kernel value each (12 for x, 12 for y, 8 for z: one Horner pass with a
//     per-lane coefficient column) and exchange them through a per-warp scratch record;
//   * 8 planes x (4 complex FMA + 1) per lane, selected by the local z start (switch -> static register indices);
//   * 5-step shuffle reduction, lane 0 scatters through the permutation.
// Shared memory is read once per sub-bin (60 LDS.64 per lane for ~16 points) instead of 16 LDS.64 per point.
#pragma once
#include "rt_common.cuh"
#include "interp.cuh"

namespace nufft {
namespace rt {

constexpr int INTERP_NW = 12;             // warps per CTA
constexpr int ZB = 8;                     // z cells per register sub-bin
constexpr int ZP = ZB + W - 1;            // register planes = 15
constexpr int SCR_F = 48;                 // per-warp scratch record: wz[8] | wyT[6][4] | wx_pad[12] (+pad)
constexpr int SCR_WZ = 0, SCR_WY = 8, SCR_WX = 32;

template <bool CPLX>
__global__ void __launch_bounds__(32 * INTERP_NW, 1)
rt_interp_kernel(KernelParams<float> kp, TileGeom g, SmArgs a, const int32_t *__restrict__ fine_offsets,
                 const float *__restrict__ xs0, const float *__restrict__ xs1, const float *__restrict__ xs2,
                 MutPtrPack vp, int C, const typename CellOf<float, CPLX>::type *__restrict__ us, int64_t ncells,
                 float prefactor, const float *__restrict__ nu_weights)
{
    using Cell = typename CellOf<float, CPLX>::type;
    constexpr int NT = 32 * INTERP_NW;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_bytes = (g.tile_cells * (int)sizeof(Cell) + 15) & ~15;
    Cell *tiles[2] = {(Cell *)smem_raw, (Cell *)(smem_raw + tile_bytes)};
    float *scr_s = (float *)(smem_raw + 2 * tile_bytes);              // [NW][2][SCR_F]
    float *cs_s = scr_s + INTERP_NW * 2 * SCR_F;                      // [3][cs_stride]
    __shared__ int s_item[2][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Tx = g.T[0], Ty = g.T[1], Tz = g.T[2], Sx = g.S[0], S2 = g.S[2];
    const int total_items = a.item_start[a.nbins];
    const LaneSlots ls = lane_slots(lane);
    const int off0 = ls.g * Sx + ls.x, off3 = ls.y3 * Sx + ls.x3;

    // role of this lane in the kernel evaluation: padded entry e of dimension rd
    const int rd = lane < 12 ? 0 : (lane < 24 ? 1 : 2);
    const int re = lane - 12 * rd;                                    // x, y: padded position 0..11; z: j = 0..7
    const float *xs_r = rd == 0 ? xs0 : (rd == 1 ? xs1 : xs2);
    const float *cs_r = cs_s + rd * kp.cs_stride;
    const bool poly = kp.mode == NUFFT_EVAL_FAST && (kp.kind == NUFFT_KERNEL_KAISER_BESSEL || kp.kind == NUFFT_KERNEL_BACKWARDS_KAISER_BESSEL);
    float *scr = scr_s + warp * 2 * SCR_F;
    // scratch positions this lane writes (relative to the record)
    int st_main;                                                      // main position
    if (rd == 0) st_main = SCR_WX + re;
    else if (rd == 1) st_main = SCR_WY + 4 * (re % 3) + re / 3;       // row y % 3, column y / 3 (y = 11 -> row2[3] = 0)
    else st_main = SCR_WZ + re;
    const bool y258 = rd == 1 && (re % 3) == 2 && re < 9;             // y in {2, 5, 8}: copies into rows 3..5

    for (int i = tid; i < 3 * kp.cs_stride; i += NT) cs_s[i] = kp.cs[i];

    // ---- tile staging (cp.async, one cell per lane; periodic wrap per row / column) ---------------------------
    auto stage = [&](Cell *tile, int bin, int c) {
        int b = bin;
        const int bx = b % g.nb[0]; b /= g.nb[0];
        const int by = b % g.nb[1]; b /= g.nb[1];
        const int bz = b;
        const Cell *u = us + (int64_t)c * ncells;
        const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];
        const int x0 = bx * g.B[0] - (M - 1), y0 = by * g.B[1] - (M - 1), z0 = bz * g.B[2] - (M - 1);
        for (int xb = 0; xb < Tx; xb += 32) {
            const int x = xb + lane;
            const bool in = x < Tx;
            const int gx = wrap1(x0 + (in ? x : 0), Nx);
            for (int z = 0; z < Tz; ++z) {
                const int gz = wrap1(z0 + z, Nz);
                const Cell *gplane = u + (int64_t)gz * Ny * Nx + gx;
                Cell *tplane = tile + z * S2 + x;
                for (int y = warp; y < Ty; y += INTERP_NW) {
                    const int gy = wrap1(y0 + y, Ny);
                    if (in) cp_async_cell<(int)sizeof(Cell)>(tplane + y * Sx, gplane + (int64_t)gy * Nx);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (tid == 0) {
        const int item = atomicAdd(a.work_counter, 1);
        s_item[0][0] = item;
        if (item < total_items) decode_item(a, item, g.chunk, s_item[0][1], s_item[0][2], s_item[0][3]);
    }
    __syncthreads();
    // flattened sequence of (item, component) steps; step s uses tile buffer s & 1
    if (s_item[0][0] < total_items) stage(tiles[0], s_item[0][1], 0);

    int step = 0;
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
        const int org_r = rd == 0 ? org0 : (rd == 1 ? org1 : org2);
        const int32_t *foff = fine_offsets + (int64_t)bin * g.nsub;
        const int nzh = (g.sub[2] + 1) / 2;                           // register sub-bins per column
        const int nrs = g.sub[0] * g.sub[1] * nzh;

        for (int c = 0; c < C; ++c, ++step) {
            __syncthreads();          // s_item[next] visible; everybody is done with tile buffer (step + 1) & 1
            // prefetch the tile of the next step, then wait for the current one
            {
                const int *nxt = s_item[(it + 1) & 1];
                if (c + 1 < C) stage(tiles[(step + 1) & 1], bin, c + 1);
                else if (nxt[0] < total_items) stage(tiles[(step + 1) & 1], nxt[1], 0);
                else asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            }
            __syncthreads();
            const Cell *tile = tiles[step & 1];

            for (int rs = warp; rs < nrs; rs += INTERP_NW) {
                const int col = rs / nzh, zh = rs - col * nzh;
                const int f0 = col * g.sub[2] + 2 * zh, f1 = min(f0 + 2, (col + 1) * g.sub[2]);
                const int p0 = max(foff[f0], k0), p1 = min(foff[f1], k1);
                if (p0 >= p1) continue;
                const int cy = col / g.sub[0], cx = col - cy * g.sub[0];
                // ---- register tile: planes 8 zh .. 8 zh + 14 (clamped to the tile), 4 slots per lane --------
                float ur[ZP][4], ui[ZP][4];
                {
                    const Cell *base = tile + (4 * cy) * Sx + 4 * cx;
#pragma unroll
                    for (int q = 0; q < ZP; ++q) {
                        const int z = min(ZB * zh + q, Tz - 1);
                        const Cell *pl = base + z * S2;
                        const Cell t0 = pl[off0], t1 = pl[off0 + 3 * Sx], t2 = pl[off0 + 6 * Sx], t3 = pl[off3];
                        if constexpr (CPLX) {
                            ur[q][0] = t0.x; ui[q][0] = t0.y; ur[q][1] = t1.x; ui[q][1] = t1.y;
                            ur[q][2] = t2.x; ui[q][2] = t2.y; ur[q][3] = t3.x; ui[q][3] = t3.y;
                        } else {
                            ur[q][0] = t0; ur[q][1] = t1; ur[q][2] = t2; ur[q][3] = t3;
                            ui[q][0] = ui[q][1] = ui[q][2] = ui[q][3] = 0.f;
                        }
                    }
                }
                float xr = xs_r[p0];
                int n_next = a.perm[p0];
                for (int p = p0; p < p1; ++p) {
                    const float x = xr;
                    const int n = n_next;
                    if (p + 1 < p1) { xr = xs_r[p + 1]; n_next = a.perm[p + 1]; }
                    // ---- kernel value of this lane's padded entry -------------------------------------------
                    float r;
                    const int i0 = point_to_cell0<float>(x, kp.N[rd], r);
                    const int t = i0 - org_r;                         // local cell of the point in the bin
                    const int j = rd == 2 ? re : re - (t & 3);        // kernel value index of this entry
                    const bool on = j >= 0 && j < W;
                    float wv;
                    if (poly) {
                        // piecewise polynomial (KB / BKB fast mode): Horner on this lane's coefficient column,
                        // same operation order as eval_kernel_values
                        const int jc = on ? j : 0;
                        const float xt = 2.f * (r - (float)i0) - 1.f;
                        wv = cs_r[(M + 3) * W + jc];
#pragma unroll
                        for (int q = M + 2; q >= 0; --q) wv = fmaf(xt, wv, cs_r[q * W + jc]);
                    } else {
                        float wall[W];
                        eval_kernel_values<float, M>(kp, cs_r, rd, x, wall);   // other kernels / Direct mode
                        wv = 0.f;
#pragma unroll
                        for (int jj = 0; jj < W; ++jj) wv = (jj == j) ? wall[jj] : wv;
                    }
                    if (!on) wv = 0.f;
                    float *sc = scr + (p & 1) * SCR_F;
                    sc[st_main] = wv;
                    if (y258) {
                        const int cc = re / 3;
                        sc[SCR_WY + 12 + cc] = wv; sc[SCR_WY + 16 + cc] = wv; sc[SCR_WY + 20 + cc] = wv;
                        sc[SCR_WY + 12 + 4 * cc + 3] = wv;
                    }
                    const int tzl = __shfl_sync(0xffffffffu, t, 24) - ZB * zh;     // local z start in the register tile
                    __syncwarp();
                    const float4 wza = *reinterpret_cast<const float4 *>(sc + SCR_WZ);
                    const float4 wzb = *reinterpret_cast<const float4 *>(sc + SCR_WZ + 4);
                    const float4 wy = *reinterpret_cast<const float4 *>(sc + SCR_WY + 4 * ls.row);
                    const float wx = sc[SCR_WX + ls.x], wx3 = sc[SCR_WX + ls.x3];
                    const float w0 = wx * wy.x, w1 = wx * wy.y, w2 = wx * wy.z, w3 = wx3 * wy.w;
                    float accr = 0.f, acci = 0.f;
#define NUFFT_RT_PL(Q, WZ)                                                                                  \
    {                                                                                                       \
        float tr = ur[Q][0] * w0;                                                                           \
        tr = fmaf(ur[Q][1], w1, tr); tr = fmaf(ur[Q][2], w2, tr); tr = fmaf(ur[Q][3], w3, tr);              \
        accr = fmaf(tr, WZ, accr);                                                                          \
        if constexpr (CPLX) {                                                                               \
            float ti = ui[Q][0] * w0;                                                                       \
            ti = fmaf(ui[Q][1], w1, ti); ti = fmaf(ui[Q][2], w2, ti); ti = fmaf(ui[Q][3], w3, ti);          \
            acci = fmaf(ti, WZ, acci);                                                                      \
        }                                                                                                   \
    }
#define NUFFT_RT_CASE(T0)                                                                                   \
    case T0:                                                                                                \
        NUFFT_RT_PL(T0 + 0, wza.x) NUFFT_RT_PL(T0 + 1, wza.y) NUFFT_RT_PL(T0 + 2, wza.z) NUFFT_RT_PL(T0 + 3, wza.w) \
        NUFFT_RT_PL(T0 + 4, wzb.x) NUFFT_RT_PL(T0 + 5, wzb.y) NUFFT_RT_PL(T0 + 6, wzb.z) NUFFT_RT_PL(T0 + 7, wzb.w) \
        break;
                    switch (tzl) {
                        NUFFT_RT_CASE(0) NUFFT_RT_CASE(1) NUFFT_RT_CASE(2) NUFFT_RT_CASE(3)
                        NUFFT_RT_CASE(4) NUFFT_RT_CASE(5) NUFFT_RT_CASE(6)
                    default:
                        NUFFT_RT_PL(7, wza.x) NUFFT_RT_PL(8, wza.y) NUFFT_RT_PL(9, wza.z) NUFFT_RT_PL(10, wza.w)
                        NUFFT_RT_PL(11, wzb.x) NUFFT_RT_PL(12, wzb.y) NUFFT_RT_PL(13, wzb.z) NUFFT_RT_PL(14, wzb.w)
                        break;
                    }
#undef NUFFT_RT_CASE
#undef NUFFT_RT_PL
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        accr += __shfl_xor_sync(0xffffffffu, accr, o);
                        if constexpr (CPLX) acci += __shfl_xor_sync(0xffffffffu, acci, o);
                    }
                    if (lane == 0) {
                        const float scale = prefactor * (nu_weights ? nu_weights[n] : 1.f);
                        if constexpr (CPLX) store_value<float, true>(vp.p[c], n, make_float2(accr * scale, acci * scale));
                        else store_value<float, false>(vp.p[c], n, accr * scale);
                    }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

inline size_t interp_smem_bytes(const TileGeom &g, int cs_stride, size_t cell_bytes)
{
    size_t b = 2 * (((size_t)g.tile_cells * cell_bytes + 15) & ~(size_t)15);
    b += (size_t)INTERP_NW * 2 * SCR_F * sizeof(float);
    b += (size_t)(3 * cs_stride + 4) * sizeof(float);
    return b + 16;
}

}  // namespace rt
}  // namespace nufft
