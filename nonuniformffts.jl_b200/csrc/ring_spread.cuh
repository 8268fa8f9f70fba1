// ring_spread.cuh — K-spread, ring-window variant (3-D, HalfSupport(4), ComplexF32): exact support along z, no register moves.
// Replaces src/spreading/gpu.jl:237-434 for this configuration class (same sums, different order).
//
// set_points orders the points by (z segment, column of 4 x 4 cells in (x, y), z CELL).  A warp walks through a chunk of
// that order.  All points of a column touch the same padded (x, y) footprint of 11 x 11 cells (lane L owns the four
// (x, y) columns rt::lane_slots(L)); along z a point touches exactly the 8 planes cz - 3 .. cz + 4, so the warp keeps a
// RING of 8 planes in registers — 4 columns x 8 planes = 32 packed (re, im) accumulators per lane — whose base plane follows
// the cell of the current point:
//   per point   5 shared-memory loads of the point's record (value x wx_pad, wy rows, 8 z weights), 4 FMUL2 and 32 FFMA2
//               per lane: G[k][(rot + j) & 7] += (v wx wy)_k x wz[j].  The rotation `rot` of the ring is warp-uniform, the
//               body exists once per rotation (8 copies selected by a switch), so ring slots are compile-time registers;
//   next cell   the planes that fall out of the window are final for this warp: 4 red.global.add.v2.f32 per lane and plane
//               (rows of 11 consecutive cells), the slot restarts at zero — nothing moves in the register file;
//   new column / end of chunk   the 8 planes are added to the grid and the ring restarts.
// Against the layer-window kernels (cs_spread.cuh): 32 instead of 44 FFMA2 per point (no zero-padded z weights), no plane
// moves (28 64-bit moves per layer there), 64 instead of 88 window registers -> 16 instead of 12 resident warps per SM.
#pragma once
#include <cuda.h>            // CUtensorMap (type only: the encoder is fetched through the runtime, ring_inst.cu)
#include "window_common.cuh"
#include "cs_spread.cuh"

namespace nufft {
namespace ring {

using rt::u64;
using rt::pk2;
using rt::unpk2;
using rt::fmul2;
using rt::ffma2;
using rt::P;
using cs::M;
using cs::W;
using cs::COL;
using cs::BATCH;

constexpr int RING = 8;                   // planes in the window = support along z
#ifndef NUFFT_RING_NWARP
#define NUFFT_RING_NWARP 16
#endif
#ifndef NUFFT_RING_FRESH
#define NUFFT_RING_FRESH 1                // 1: a second set of 8 bodies restarts the retired plane with a product (no zeroing)
#endif
constexpr int NWARP = NUFFT_RING_NWARP;   // warps per CTA (one CTA per SM): <= 128 registers per thread
constexpr int REC_F = 60;                 // floats per point record (240 bytes: 16-byte stores of 8 lanes are conflict-free)
constexpr int OFF_VX = 0;                 // [0..23]  spreading: value x wx_pad[0..10] as 11 (re, im) pairs + one zero pair
                                          //          interpolation: wx_pad[0..10], 0 in [0..11]
constexpr int OFF_WY = 24;                // [24..47] wyT rows (rt::store_y layout)
constexpr int OFF_WZ = 48;                // [48..55] wz[0..7] (not padded: plane cz - 3 + j)
constexpr int STAGE_F = 7 * 32;           // as cs:: (folded record, value, index)
// TMA plane retirement: a leaving plane of the window (11 x 11 cells) is written to a warp-private shared-memory tile of
// 11 rows x 12 cells (one zero pad cell to the left: the box starts at an even grid x, 16-byte aligned) and leaves as ONE
// cp.reduce.async.bulk.tensor.3d ... .add (UTMAREDG): the reduction is done by the TMA unit / L2, not by per-lane RED
// instructions (whose issue rate, about two lanes per clock and SM, bounded the kernel: 1.8 of 2.8 ms at C3).
constexpr int TMA_ROW = 12;               // cells per tile row
constexpr int TMA_TILE_B = 1152;          // bytes per tile (11 x 12 x 8 = 1056, padded to a multiple of 128)
constexpr int TMA_NBUF = 3;               // tiles in flight per warp

struct Rec {                              // what one lane keeps of a point record
    u64 vx, vx3;                          // value x wx of the lane's columns 0..2 / column 3
    float4 wy;
    float4 z0, z1;
};
__device__ __forceinline__ u64 lds64(const float *p)
{
    return *reinterpret_cast<const u64 *>(p);
}

// spreading body for ring rotation R: slot of plane j of the point = (R + j) & 7.  The record registers are reloaded with the
// NEXT point's record as soon as their last use has been issued (vx / wy right after the four products, z0 after the first
// half of the planes, z1 at the end), so the shared-memory latency is covered by this point's FFMA2s.
struct LaneOffs {                         // float offsets of the lane's operands inside a point record
    int vx, vx3, wy;
};
// FRESH: the newest plane of the window (j = 7, ring slot (R + 7) & 7) has just been retired to the grid and restarts with
// this point's contribution: a product instead of a fused multiply-add, so the slot never has to be zeroed.
template <int R, bool FRESH>
__device__ __forceinline__ void spread_body(u64 (&G)[4][RING], Rec &A, const float *nxt, const LaneOffs &lo)
{
    const u64 a0 = fmul2(A.vx, pk2(A.wy.x, A.wy.x));
    const u64 a1 = fmul2(A.vx, pk2(A.wy.y, A.wy.y));
    const u64 a2 = fmul2(A.vx, pk2(A.wy.z, A.wy.z));
    const u64 a3 = fmul2(A.vx3, pk2(A.wy.w, A.wy.w));
    A.vx = lds64(nxt + lo.vx);
    A.vx3 = lds64(nxt + lo.vx3);
    A.wy = *reinterpret_cast<const float4 *>(nxt + lo.wy);
    {
        const float wz[4] = {A.z0.x, A.z0.y, A.z0.z, A.z0.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const u64 w = pk2(wz[j], wz[j]);
            G[0][(R + j) & 7] = ffma2(a0, w, G[0][(R + j) & 7]);
            G[1][(R + j) & 7] = ffma2(a1, w, G[1][(R + j) & 7]);
            G[2][(R + j) & 7] = ffma2(a2, w, G[2][(R + j) & 7]);
            G[3][(R + j) & 7] = ffma2(a3, w, G[3][(R + j) & 7]);
        }
    }
    A.z0 = *reinterpret_cast<const float4 *>(nxt + OFF_WZ);
    {
        const float wz[4] = {A.z1.x, A.z1.y, A.z1.z, A.z1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const u64 w = pk2(wz[j], wz[j]);
            if (FRESH && j == 3) {
                G[0][(R + 7) & 7] = fmul2(a0, w);
                G[1][(R + 7) & 7] = fmul2(a1, w);
                G[2][(R + 7) & 7] = fmul2(a2, w);
                G[3][(R + 7) & 7] = fmul2(a3, w);
            } else {
                G[0][(R + 4 + j) & 7] = ffma2(a0, w, G[0][(R + 4 + j) & 7]);
                G[1][(R + 4 + j) & 7] = ffma2(a1, w, G[1][(R + 4 + j) & 7]);
                G[2][(R + 4 + j) & 7] = ffma2(a2, w, G[2][(R + 4 + j) & 7]);
                G[3][(R + 4 + j) & 7] = ffma2(a3, w, G[3][(R + 4 + j) & 7]);
            }
        }
    }
    A.z1 = *reinterpret_cast<const float4 *>(nxt + OFF_WZ + 4);
}

// one lane evaluates the three 1-D kernels of its point into the point's record; returns the cells.
// SPREAD: the x weights are stored multiplied by the point's value.
template <bool SPREAD>
__device__ __forceinline__ void evaluate_point(const KernelParams<float> &kp, const float *cs_s, float x, float y, float z, float2 v,
                                               float *r, int &cx, int &cy, int &cz)
{
    float w[W], pw[P];
    float4 *q;
    cx = cs::eval_m4<0>(kp, cs_s, x, w);
    rt::pad_shift(w, cx & 3, pw);
    q = reinterpret_cast<float4 *>(r + OFF_VX);
    if constexpr (SPREAD) {
        q[0] = make_float4(v.x * pw[0], v.y * pw[0], v.x * pw[1], v.y * pw[1]);
        q[1] = make_float4(v.x * pw[2], v.y * pw[2], v.x * pw[3], v.y * pw[3]);
        q[2] = make_float4(v.x * pw[4], v.y * pw[4], v.x * pw[5], v.y * pw[5]);
        q[3] = make_float4(v.x * pw[6], v.y * pw[6], v.x * pw[7], v.y * pw[7]);
        q[4] = make_float4(v.x * pw[8], v.y * pw[8], v.x * pw[9], v.y * pw[9]);
        q[5] = make_float4(v.x * pw[10], v.y * pw[10], 0.f, 0.f);
    } else {
        q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
        q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
        q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
    }
    cy = cs::eval_m4<1>(kp, cs_s + kp.cs_stride, y, w);
    rt::pad_shift(w, cy & 3, pw);
    {
        float4 *rr = reinterpret_cast<float4 *>(r + OFF_WY);
        rr[0] = make_float4(pw[0], pw[3], pw[6], pw[9]);
        rr[1] = make_float4(pw[1], pw[4], pw[7], pw[10]);
        rr[2] = make_float4(pw[2], pw[5], pw[8], 0.f);
        rr[3] = make_float4(pw[2], pw[5], pw[8], pw[2]);
        rr[4] = make_float4(pw[2], pw[5], pw[8], pw[5]);
        rr[5] = make_float4(pw[2], pw[5], pw[8], pw[8]);
    }
    cz = cs::eval_m4<2>(kp, cs_s + 2 * kp.cs_stride, z, w);
    q = reinterpret_cast<float4 *>(r + OFF_WZ);
    q[0] = make_float4(w[0], w[1], w[2], w[3]);
    q[1] = make_float4(w[4], w[5], w[6], w[7]);
}

// plane index of global z plane `z` inside the (possibly slab-local) grid: periodic wrap on a full grid (zlo = 0,
// nzwrap = Nz), plain offset on a z slab with halo (zlo = first stored plane, nzwrap large: never wraps)
__device__ __forceinline__ int plane_of(int z, int zlo, int nzwrap)
{
    return wrap1(z - zlo, nzwrap);
}

__device__ __forceinline__ u64 zero64()
{
    u64 r;
    asm("mov.b64 %0, 0;" : "=l"(r));
    return r;
}

#define NUFFT_RING_SWITCH(rot_, F_)                                                                                               \
    switch (rot_) {                                                                                                               \
    case 0: F_(0) break;                                                                                                          \
    case 1: F_(1) break;                                                                                                          \
    case 2: F_(2) break;                                                                                                          \
    case 3: F_(3) break;                                                                                                          \
    case 4: F_(4) break;                                                                                                          \
    case 5: F_(5) break;                                                                                                          \
    case 6: F_(6) break;                                                                                                          \
    default: F_(7) break;                                                                                                         \
    }

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *tm, unsigned smem, int c0, int c1, int c2)
{
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(c2), "r"(smem)
                 : "memory");
}

template <int NW, bool TMA>               // NW == NWARP (a template so that every translation unit may include it)
__global__ void __launch_bounds__(32 * NW)
ring_spread_kernel(const __grid_constant__ CUtensorMap tmap, KernelParams<float> kp, TileGeom g, int np, int chunk, const int32_t *__restrict__ perm, int32_t *work_counter,
                   const float4 *__restrict__ prec, PtrPack vp, int C, float2 *__restrict__ us, int64_t ncells,
                   const float *__restrict__ nu_weights, int zlo, int nzwrap, int nzloc)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *tile_all = smem_raw;                                      // [NWARP][TMA_NBUF][TMA_TILE_B] (TMA only)
    float *rec_all = (float *)(smem_raw + (TMA ? NWARP * TMA_NBUF * TMA_TILE_B : 0));      // [NWARP][BATCH + 1][REC_F]
    float *stage_all = rec_all + NWARP * (BATCH + 1) * REC_F;                // [NWARP][STAGE_F]
    int2 *key_all = (int2 *)(stage_all + NWARP * STAGE_F);                   // [NWARP][BATCH + 1] (column id, z cell)
    float *cs_s = (float *)(key_all + NWARP * (BATCH + 1));                  // [3][cs_stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    float *rec_w = rec_all + warp * (BATCH + 1) * REC_F;
    int2 *key_w = key_all + warp * (BATCH + 1);
    float4 *st_x = reinterpret_cast<float4 *>(stage_all + warp * STAGE_F) + lane;
    float2 *st_v = reinterpret_cast<float2 *>(stage_all + warp * STAGE_F + 128) + lane;
    int32_t *st_n = reinterpret_cast<int32_t *>(stage_all + warp * STAGE_F + 192) + lane;

    for (int i = tid; i < 3 * kp.cs_stride; i += 32 * NWARP) cs_s[i] = kp.cs[i];
    if (TMA) for (int i = tid; i < NWARP * TMA_NBUF * TMA_TILE_B / 4; i += 32 * NWARP) ((float *)tile_all)[i] = 0.f;    // pad cells stay zero
    // record BATCH (one past the last point) is only ever prefetched, never used: keep it finite
    for (int i = lane; i < REC_F; i += 32) rec_w[BATCH * REC_F + i] = 0.f;
    if (lane == 0) key_w[BATCH] = make_int2(-2, 0);
    __syncthreads();                                   // the only CTA barrier: coefficient tables

    const rt::LaneSlots ls = rt::lane_slots(lane);
    LaneOffs lo;
    lo.vx = OFF_VX + 2 * ls.x;
    lo.vx3 = OFF_VX + 2 * ls.x3;
    lo.wy = OFF_WY + 4 * ls.row;
    asm volatile("" : "+r"(lo.vx), "+r"(lo.vx3), "+r"(lo.wy));      // opaque: kept in registers, not recomputed per point
    // cells of the lane's four columns inside a TMA tile (row-major, TMA_ROW cells per row, one pad cell to the left)
    const int so0 = (ls.g) * TMA_ROW + ls.x + 1, so1 = (ls.g + 3) * TMA_ROW + ls.x + 1, so2 = (ls.g + 6) * TMA_ROW + ls.x + 1;
    const int so3 = ls.y3 * TMA_ROW + ls.x3 + 1;
    float2 *tile_w = reinterpret_cast<float2 *>(tile_all + (size_t)warp * TMA_NBUF * TMA_TILE_B);
    const int Nx = g.N[0], Ny = g.N[1];
    const int plane = Nx * Ny;                         // cells per z plane (the whole grid has < 2^31 cells)
    const unsigned long long pol = cs::l2_evict_first_policy();

    u64 G[4][RING];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < RING; ++i) G[k][i] = 0ull;

    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1);
        item = __shfl_sync(FULL, item, 0);
        const int64_t k0l = (int64_t)item * chunk;
        if (k0l >= np) break;
        const int k0 = (int)k0l, k1 = min(k0 + chunk, np);
        const int nbatches = (k1 - k0 + BATCH - 1) / BATCH;

        for (int c = 0; c < C; ++c) {
            const float2 *vc = (const float2 *)vp.p[c];
            float2 *u = us + (int64_t)c * ncells;
            // ---- window state: column, base plane zb (ring slot `rot` holds plane zb, slot (rot + j) & 7 plane zb + j) ------
            int wcol = -1, zb = 0, rot = 0;
            int zpl = 0;                                   // index of plane zb inside the stored grid (plane_of(zb))
            int tc0 = 0, tc1 = 0, tbuf = 0;                // TMA: box origin (floats along x, rows), tile in use
            bool edge = false;                             // TMA: the column's box sticks out of the grid (periodic images; boxes with
                                                           // negative coordinates are illegal for reductions) -> per-lane reductions
            const int tcz = c * nzloc;                     // plane coordinate of component c in the tensor map
            float2 *q0 = u, *q1 = u, *q2 = u, *q3 = u;     // the lane's four cells in plane zb (a lane without a 4th column
                                                           // accumulates zeros there and adds them to its first)
            // `dz` planes leave the window (1 <= dz <= 8): ring slot -> grid plane zb, the slot restarts at zero; the cell
            // pointers follow the base plane incrementally.  ONE call site (the 8-way switch is inlined once: the hot code must
            // stay inside the instruction cache)
            auto next_plane = [&]() {
                if constexpr (!TMA) { q0 += plane; q1 += plane; q2 += plane; q3 += plane; }
                if (++zpl == nzwrap) {                     // periodic wrap of a full grid (never on a slab)
                    zpl = 0;
                    if constexpr (!TMA) {
                        const int64_t back = (int64_t)nzwrap * plane;
                        q0 -= back; q1 -= back; q2 -= back; q3 -= back;
                    }
                }
                rot = (rot + 1) & 7;
                ++zb;
            };
            // one plane (the four values of this lane) leaves the window for grid plane zb
            auto retire_plane = [&](u64 g0, u64 g1, u64 g2, u64 g3) {
                if (TMA && !edge) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(TMA_NBUF - 1) : "memory");     // tile tbuf is free again
                    __syncwarp();
                    float2 *sb = tile_w + tbuf * (TMA_TILE_B / 8);
                    *reinterpret_cast<u64 *>(sb + so0) = g0;
                    *reinterpret_cast<u64 *>(sb + so1) = g1;
                    *reinterpret_cast<u64 *>(sb + so2) = g2;
                    if (ls.has3) *reinterpret_cast<u64 *>(sb + so3) = g3;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the TMA unit
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_3d(&tmap, (unsigned)__cvta_generic_to_shared(sb), tc0, tc1, tcz + zpl);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    tbuf = (tbuf + 1 == TMA_NBUF) ? 0 : tbuf + 1;
                } else {
                    if constexpr (TMA) {                   // (the cell pointers are not kept in registers on the TMA path)
                        const int cx = wcol & 0xffff, cy = wcol >> 16;
                        const unsigned pb = (unsigned)(zpl * plane);
                        const int e0 = wrap1(COL * cy - (M - 1) + ls.g, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                        const int e1 = wrap1(COL * cy - (M - 1) + ls.g + 3, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                        const int e2 = wrap1(COL * cy - (M - 1) + ls.g + 6, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                        const int e3 = ls.has3 ? wrap1(COL * cy - (M - 1) + ls.y3, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x3, Nx) : e0;
                        cs::red_cell(u + (pb + (unsigned)e0), g0); cs::red_cell(u + (pb + (unsigned)e1), g1);
                        cs::red_cell(u + (pb + (unsigned)e2), g2); cs::red_cell(u + (pb + (unsigned)e3), g3);
                    } else {
                        cs::red_cell(q0, g0); cs::red_cell(q1, g1); cs::red_cell(q2, g2); cs::red_cell(q3, g3);
                    }
                }
            };
            auto advance = [&](int dz) {
#pragma unroll 1
                for (int s = 0; s < dz; ++s) {
#define NUFFT_RING_RETIRE(R_)                                                                                                     \
    {                                                                                                                             \
        retire_plane(G[0][R_], G[1][R_], G[2][R_], G[3][R_]);                                                                     \
        G[0][R_] = zero64(); G[1][R_] = zero64(); G[2][R_] = zero64(); G[3][R_] = zero64();                                       \
    }
                    NUFFT_RING_SWITCH(rot, NUFFT_RING_RETIRE)
#undef NUFFT_RING_RETIRE
                    next_plane();
                }
            };

            // ---- global loads: one lane per point, staged through cp.async one batch ahead (the index two steps ahead) --------
            auto issue_n = [&](int bi) {
                const int k = k0 + bi * BATCH + lane;
                if (k < k1) cs::cp_async_stream<4>(st_n, perm + k, pol);
            };
            auto issue_xv = [&](int bi, int32_t n) {
                const int k = k0 + bi * BATCH + lane;
                if (k < k1) {
                    cs::cp_async_stream<16>(st_x, prec + n, pol);
                    cs::cp_async_stream<8>(st_v, vc + n, pol);
                }
            };
            issue_n(0);
            cs::cp_async_commit();
            cs::cp_async_wait0();
            int32_t n_cur = *st_n;
            float wgt = 1.f;
            issue_xv(0, n_cur);
            issue_n(1);
            cs::cp_async_commit();

            for (int bi = 0; bi < nbatches; ++bi) {
                const int nb = min(BATCH, k1 - (k0 + bi * BATCH));
                cs::cp_async_wait0();
                const float4 xyz = *st_x;
                float2 v = *st_v;
                if (nu_weights && lane < nb) wgt = nu_weights[n_cur];
                n_cur = *st_n;
                issue_xv(bi + 1, n_cur);
                issue_n(bi + 2);
                cs::cp_async_commit();

                // ---- evaluate: one lane per point; the same lanes work out what the WINDOW has to do before their point ------------
                // cmd: 0 = same column and cell as the previous point (no move), 1..8 = the window moves up by that many planes,
                // CMD_NEW = another column (or a jump along z): flush and restart.  Lane 0 compares with the window the previous
                // batch left behind.
                constexpr int CMD_NEW = 64;
                int mycol = -3, myz = 0, cmd = 0;
                if (lane < nb) {
                    int cx, cy, cz;
                    if (nu_weights) v = cmul(v, wgt);
                    evaluate_point<true>(kp, cs_s, xyz.x, xyz.y, xyz.z, v, rec_w + lane * REC_F, cx, cy, cz);
                    mycol = ((cy >> 2) << 16) | (cx >> 2);
                    myz = cz - (M - 1);                                     // base plane of the point's window
                    key_w[lane] = make_int2(mycol, myz);
                }
                {
                    int pc = __shfl_up_sync(FULL, mycol, 1), pz = __shfl_up_sync(FULL, myz, 1);
                    if (lane == 0) { pc = wcol; pz = zb; }
                    const int dz = myz - pz;
                    cmd = (mycol != pc || dz < 0 || dz > RING) ? CMD_NEW : dz;
                }
                const unsigned starts = __ballot_sync(FULL, lane < nb && cmd != 0);     // points that move the window
                const bool last = bi == nbatches - 1;                                   // the chunk ends: flush the window
                __syncwarp();

                // ---- accumulate: run after run (points sharing column and cell), the ring follows the z cell ------------------------
                Rec A;
                A.vx = lds64(rec_w + lo.vx);
                A.vx3 = lds64(rec_w + lo.vx3);
                A.wy = *reinterpret_cast<const float4 *>(rec_w + lo.wy);
                A.z0 = *reinterpret_cast<const float4 *>(rec_w + OFF_WZ);
                A.z1 = *reinterpret_cast<const float4 *>(rec_w + OFF_WZ + 4);
                const float *nxt = rec_w + REC_F;
                int p = 0;
#pragma unroll 1
                while (true) {
                    int c = 0;
                    if (p < nb) c = ((starts >> p) & 1u) ? __shfl_sync(FULL, cmd, p) : 0;
                    else if (last) c = CMD_NEW;
                    else break;
                    if (c != 0) {
                        const bool move = c != CMD_NEW;                        // the window moves up inside the column
#if NUFFT_RING_FRESH
                        advance(move ? c - 1 : (wcol >= 0 ? RING : 0));
#else
                        advance(move ? c : (wcol >= 0 ? RING : 0));
#endif
                        if (!move) {                                           // new column (or a jump along z): restart the ring
                            if (p >= nb) { wcol = -2; break; }                 // end of the chunk: window flushed
                            const int2 key = key_w[p];
                            wcol = key.x;
                            zb = key.y;
                            rot = 0;
                            const int cx = key.x & 0xffff, cy = key.x >> 16;
                            zpl = plane_of(zb, zlo, nzwrap);
                            if (nzwrap != nzloc) zpl = min(max(zpl, 0), nzloc - RING);     // z slab: a point outside the slab (caller's
                                                                                           // contract) must not write outside the grid
                            const unsigned pb = (unsigned)(zpl * plane);       // first cell of the plane (the grid has < 2^31 cells)
                            const int g0 = wrap1(COL * cy - (M - 1) + ls.g, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                            const int g1 = wrap1(COL * cy - (M - 1) + ls.g + 3, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                            const int g2 = wrap1(COL * cy - (M - 1) + ls.g + 6, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                            const int g3 = ls.has3 ? wrap1(COL * cy - (M - 1) + ls.y3, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x3, Nx) : g0;
                            q0 = u + (pb + (unsigned)g0); q1 = u + (pb + (unsigned)g1);
                            q2 = u + (pb + (unsigned)g2); q3 = u + (pb + (unsigned)g3);
                            // TMA box: cells 4 cx - 4 .. 4 cx + 7 (x, in floats: two per cell), rows 4 cy - 3 .. 4 cy + 7
                            tc0 = 8 * cx - 8;
                            tc1 = COL * cy - (M - 1);
                            edge = cx == 0 || COL * cx + 8 > Nx || tc1 < 0 || COL * cy + 8 > Ny;
                        }
#if NUFFT_RING_FRESH
                        else {
                            // the last leaving plane is retired here and restarted by the first point of the new cell
#define NUFFT_RING_FRESH_BODY(R_)                                                                                                 \
    {                                                                                                                             \
        retire_plane(G[0][R_], G[1][R_], G[2][R_], G[3][R_]);                                                                     \
        spread_body<(R_ + 1) & 7, true>(G, A, nxt, lo);                                                                           \
    }
                            NUFFT_RING_SWITCH(rot, NUFFT_RING_FRESH_BODY)
#undef NUFFT_RING_FRESH_BODY
                            next_plane();
                            nxt += REC_F;
                            ++p;
                            if (p >= nb || ((starts >> p) & 1u)) continue;     // the run had one point
                        }
#endif
                    }
                    // the rest of the run: all consecutive points with this (column, cell) go through the body of this rotation
                    const unsigned after = p + 1 < 32 ? (starts >> (p + 1)) << (p + 1) : 0u;
                    const int pend = after ? min(__ffs(after) - 1, nb) : nb;
                    int len = pend - p;
                    p = pend;
#define NUFFT_RING_BODY(R_)                                                                                                       \
    _Pragma("unroll 1") do {                                                                                                      \
        spread_body<R_, false>(G, A, nxt, lo);                                                                                    \
        nxt += REC_F;                                                                                                             \
    } while (--len);
                    NUFFT_RING_SWITCH(rot, NUFFT_RING_BODY)
#undef NUFFT_RING_BODY
                }
                __syncwarp();
            }
        }
    }
    if (TMA && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every reduction has been performed
}

inline size_t spread_smem_bytes(int cs_stride, bool tma)
{
    return (tma ? (size_t)NWARP * TMA_NBUF * TMA_TILE_B : 0) + (size_t)NWARP * ((BATCH + 1) * REC_F + STAGE_F) * sizeof(float) + (size_t)NWARP * (BATCH + 1) * sizeof(int2) +
           (size_t)(3 * cs_stride + 4) * sizeof(float) + 16;
}

}  // namespace ring
}  // namespace nufft
