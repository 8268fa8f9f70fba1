// ring_interp.cuh — K-interp, ring-window variant (3-D, HalfSupport(4), ComplexF32).  Mirror image of ring_spread.cuh;
// replaces src/interpolation/gpu.jl:211-395 for this configuration class (same sums, different order).
//
// A warp walks through a chunk of the points ordered by (z segment, column of 4 x 4 cells, z cell) and keeps the 8 planes
// cz - 3 .. cz + 4 of the column's padded 11 x 11 footprint in a register ring (lane L holds the four (x, y) columns
// rt::lane_slots(L) x 8 planes = 32 packed (re, im) grid values):
//   per point   5 shared-memory loads of the point's record, 32 FFMA2 (t_k = sum_j G[k][(rot + j) & 7] wz[j]) + 2 FMUL2 +
//               1 FMUL2 + 3 FFMA2 (sum_k t_k wx wy); the lane partial goes to a warp-private shared-memory row (STS.64);
//   per 8 pts   4 lanes per point add up the 32 partials of that point (8 LDS.64 each, conflict-free rows) + 2 shuffles;
//               prefactor, non-uniform callback, scatter through the permutation;
//   next cell   the plane that enters the window was requested from global memory (L2) PF cells earlier (4 cp.async per
//               lane into a warp-private staging ring: no registers, no scoreboards held); 4 LDS.64 bring it into the ring
//               slot that the leaving plane frees — nothing moves in the register file;
//   new column  32 loads per lane (rows of 11 consecutive cells).
#pragma once
#include "ring_spread.cuh"
#include "cs_interp.cuh"

namespace nufft {
namespace ring {

constexpr int REC_I = 44;                 // floats per point record of the interpolation kernel (176 bytes, conflict-free)
constexpr int IOFF_WX = 0;                // [0..11]  wx_pad[0..10], 0
constexpr int IOFF_WY = 12;               // [12..35] wyT rows
constexpr int IOFF_WZ = 36;               // [36..43] wz[0..7]
constexpr int HALF = 8;                   // points per reduction round
constexpr int PART_LD = 36;               // row length (u64) of the lane-partial buffer: 36 = 4 mod 16 -> conflict-free column sums
constexpr int PF = 3;                     // planes requested ahead of the window

struct IRec {
    float wx, wx3;
    float4 wy, z0, z1;
};

__device__ __forceinline__ void evaluate_point_interp(const KernelParams<float> &kp, const float *cs_s, float x, float y, float z,
                                                      float *r, int &cx, int &cy, int &cz)
{
    float w[W], pw[P];
    float4 *q;
    cx = cs::eval_m4<0>(kp, cs_s, x, w);
    rt::pad_shift(w, cx & 3, pw);
    q = reinterpret_cast<float4 *>(r + IOFF_WX);
    q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
    q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
    q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
    cy = cs::eval_m4<1>(kp, cs_s + kp.cs_stride, y, w);
    rt::pad_shift(w, cy & 3, pw);
    q = reinterpret_cast<float4 *>(r + IOFF_WY);
    q[0] = make_float4(pw[0], pw[3], pw[6], pw[9]);
    q[1] = make_float4(pw[1], pw[4], pw[7], pw[10]);
    q[2] = make_float4(pw[2], pw[5], pw[8], 0.f);
    q[3] = make_float4(pw[2], pw[5], pw[8], pw[2]);
    q[4] = make_float4(pw[2], pw[5], pw[8], pw[5]);
    q[5] = make_float4(pw[2], pw[5], pw[8], pw[8]);
    cz = cs::eval_m4<2>(kp, cs_s + 2 * kp.cs_stride, z, w);
    q = reinterpret_cast<float4 *>(r + IOFF_WZ);
    q[0] = make_float4(w[0], w[1], w[2], w[3]);
    q[1] = make_float4(w[4], w[5], w[6], w[7]);
}

// interpolation body for ring rotation R; the record registers are reloaded with the next point's record after their last use
template <int R>
__device__ __forceinline__ void interp_body(const u64 (&G)[4][RING], IRec &A, const float *nxt, const LaneOffs &lo, u64 *part)
{
    u64 t0, t1, t2, t3;
    {
        const u64 w = pk2(A.z0.x, A.z0.x);
        t0 = fmul2(G[0][R & 7], w); t1 = fmul2(G[1][R & 7], w); t2 = fmul2(G[2][R & 7], w); t3 = fmul2(G[3][R & 7], w);
    }
    {
        const float wz[3] = {A.z0.y, A.z0.z, A.z0.w};
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            const u64 w = pk2(wz[j - 1], wz[j - 1]);
            t0 = ffma2(G[0][(R + j) & 7], w, t0); t1 = ffma2(G[1][(R + j) & 7], w, t1);
            t2 = ffma2(G[2][(R + j) & 7], w, t2); t3 = ffma2(G[3][(R + j) & 7], w, t3);
        }
    }
    A.z0 = *reinterpret_cast<const float4 *>(nxt + IOFF_WZ);
    {
        const float wz[4] = {A.z1.x, A.z1.y, A.z1.z, A.z1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const u64 w = pk2(wz[j], wz[j]);
            t0 = ffma2(G[0][(R + 4 + j) & 7], w, t0); t1 = ffma2(G[1][(R + 4 + j) & 7], w, t1);
            t2 = ffma2(G[2][(R + 4 + j) & 7], w, t2); t3 = ffma2(G[3][(R + 4 + j) & 7], w, t3);
        }
    }
    A.z1 = *reinterpret_cast<const float4 *>(nxt + IOFF_WZ + 4);
    const float2 wa = unpk2(fmul2(pk2(A.wx, A.wx), pk2(A.wy.x, A.wy.y)));
    const float2 wb = unpk2(fmul2(pk2(A.wx, A.wx3), pk2(A.wy.z, A.wy.w)));
    A.wx = nxt[lo.vx];
    A.wx3 = nxt[lo.vx3];
    A.wy = *reinterpret_cast<const float4 *>(nxt + lo.wy);
    u64 acc = fmul2(t0, pk2(wa.x, wa.x));
    acc = ffma2(t1, pk2(wa.y, wa.y), acc);
    acc = ffma2(t2, pk2(wb.x, wb.x), acc);
    acc = ffma2(t3, pk2(wb.y, wb.y), acc);
    *part = acc;
}

// TMA plane loads (UTMALDG): the plane that enters the window arrives as ONE cp.async.bulk.tensor.3d box (11 rows x 12 cells,
// ring_spread.cuh) in a warp-private tile, signalled through an mbarrier — no per-lane LDGSTS.  Columns whose box sticks out of
// the grid (periodic images: the first / last column along x or y) read their planes with plain loads instead.
__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra MBAR_WAIT_%=;\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned smem, const CUtensorMap *tm, int c0, int c1, int c2, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}

template <int NW, bool TMA>               // NW == NWARP (a template so that every translation unit may include it)
__global__ void __launch_bounds__(32 * NW)
ring_interp_kernel(const __grid_constant__ CUtensorMap tmap, KernelParams<float> kp, TileGeom g, int np, int chunk, const int32_t *__restrict__ perm, int32_t *work_counter,
                   const float4 *__restrict__ prec, MutPtrPack vp, int C, const float2 *__restrict__ us, int64_t ncells,
                   float prefactor, const float *__restrict__ nu_weights, int zlo, int nzwrap, int nzloc)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // plane staging: TMA tiles [NWARP][PF][TMA_TILE_B] (128-byte aligned) or per-lane slots [NWARP][PF][4 columns][32 lanes] u64
    constexpr int HST_B = TMA ? PF * TMA_TILE_B : PF * 4 * 32 * 8;
    unsigned char *hst_all = smem_raw;
    float *rec_all = (float *)(smem_raw + NWARP * HST_B);                    // [NWARP][BATCH + 1][REC_I]
    float *stage_all = rec_all + NWARP * (BATCH + 1) * REC_I;                // [NWARP][5 * 32]: folded record (16 B), index
    u64 *part_all = (u64 *)(stage_all + NWARP * 5 * 32);                     // [NWARP][HALF][PART_LD] lane partials
    int2 *key_all = (int2 *)(part_all + NWARP * HALF * PART_LD);             // [NWARP][BATCH + 1]
    u64 *bar_all = (u64 *)(key_all + NWARP * (BATCH + 1));                   // [NWARP][PF] mbarriers (TMA)
    float *cs_s = (float *)(bar_all + NWARP * PF);                           // [3][cs_stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    float *rec_w = rec_all + warp * (BATCH + 1) * REC_I;
    u64 *part_w = part_all + warp * HALF * PART_LD;
    float2 *hst_w = reinterpret_cast<float2 *>(hst_all + (size_t)warp * HST_B) + (TMA ? 0 : lane);
    const unsigned bar_w = (unsigned)__cvta_generic_to_shared(bar_all + warp * PF);
    if (TMA && lane == 0) {
        for (int i = 0; i < PF; ++i) mbar_init(bar_w + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int2 *key_w = key_all + warp * (BATCH + 1);
    float4 *st_x = reinterpret_cast<float4 *>(stage_all + warp * 5 * 32) + lane;
    int32_t *st_n = reinterpret_cast<int32_t *>(stage_all + warp * 5 * 32 + 128) + lane;

    for (int i = tid; i < 3 * kp.cs_stride; i += 32 * NWARP) cs_s[i] = kp.cs[i];
    for (int i = lane; i < REC_I; i += 32) rec_w[BATCH * REC_I + i] = 0.f;   // only ever prefetched: keep it finite
    if (lane == 0) key_w[BATCH] = make_int2(-2, 0);
    __syncthreads();                                   // the only CTA barrier: coefficient tables

    const rt::LaneSlots ls = rt::lane_slots(lane);
    LaneOffs lo;
    lo.vx = IOFF_WX + ls.x;
    lo.vx3 = IOFF_WX + ls.x3;
    lo.wy = IOFF_WY + 4 * ls.row;
    asm volatile("" : "+r"(lo.vx), "+r"(lo.vx3), "+r"(lo.wy));
    // byte offsets of the lane's four cells inside a TMA tile (row-major, TMA_ROW cells per row, one pad cell to the left)
    const int so0 = 8 * ((ls.g) * TMA_ROW + ls.x + 1), so1 = 8 * ((ls.g + 3) * TMA_ROW + ls.x + 1), so2 = 8 * ((ls.g + 6) * TMA_ROW + ls.x + 1);
    const int so3 = 8 * ((ls.has3 ? ls.y3 : ls.g) * TMA_ROW + (ls.has3 ? ls.x3 : ls.x) + 1);
    const int Nx = g.N[0], Ny = g.N[1];
    const int plane = Nx * Ny;
    const unsigned long long pol = cs::l2_evict_first_policy();

    u64 G[4][RING];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < RING; ++i) G[k][i] = 0ull;
    unsigned phase = 0;                                // TMA: parity of the next completion of each tile's mbarrier — the barriers live
                                                       // as long as the kernel, so does this (NOT per chunk)

    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1);
        item = __shfl_sync(FULL, item, 0);
        const int64_t k0l = (int64_t)item * chunk;
        if (k0l >= np) break;
        const int k0 = (int)k0l, k1 = min(k0 + chunk, np);
        const int nbatches = (k1 - k0 + BATCH - 1) / BATCH;

        for (int c = 0; c < C; ++c) {
            float2 *vc = (float2 *)vp.p[c];
            const float2 *u = us + (int64_t)c * ncells;
            int wcol = -1, zb = 0, rot = 0, si = 0;        // column, base plane, ring rotation, staging slot of plane zb + 8
            int zr = 0;                                    // index inside the stored grid of the next plane to request
            const float2 *r0 = u, *r1 = u, *r2 = u, *r3 = u;     // the lane's four cells in that plane (per-lane paths)
            int tc0 = 0, tc1 = 0;                          // TMA: box origin of the column (floats along x, rows)
            bool edge = false;                             // TMA: the column's box sticks out of the grid -> plain loads
            unsigned pending = 0;                          // TMA: a load is in flight, per tile
            const int tcz = c * nzloc;

            auto next_request_plane = [&]() {
                if constexpr (!TMA) { r0 += plane; r1 += plane; r2 += plane; r3 += plane; }
                if (++zr == nzwrap) {                      // periodic wrap of a full grid (never on a slab)
                    zr = 0;
                    if constexpr (!TMA) {
                        const int64_t back = (int64_t)nzwrap * plane;
                        r0 -= back; r1 -= back; r2 -= back; r3 -= back;
                    }
                }
            };
            auto lane_cells = [&](int cx, int cy, int (&e)[4]) {      // cell offsets of the lane's four columns inside a z plane
                e[0] = wrap1(COL * cy - (M - 1) + ls.g, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                e[1] = wrap1(COL * cy - (M - 1) + ls.g + 3, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                e[2] = wrap1(COL * cy - (M - 1) + ls.g + 6, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                e[3] = ls.has3 ? wrap1(COL * cy - (M - 1) + ls.y3, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x3, Nx) : e[0];
            };
            auto request = [&](int slot) {                 // next plane -> staging slot
                if constexpr (TMA) {
                    if (lane == 0) {                       // (a plane past a slab's halo is filled with zeros and never used)
                        mbar_expect_tx(bar_w + 8 * slot, 2 * TMA_ROW * 11 * 4);
                        tma_load_3d((unsigned)__cvta_generic_to_shared((unsigned char *)hst_w + slot * TMA_TILE_B), &tmap, tc0, tc1, tcz + zr,
                                    bar_w + 8 * slot);
                    }
                    pending |= 1u << slot;
                } else {
                    if (zr < nzloc) {                      // (planes past a slab's halo are never used)
                        cp_async_cell<8>(hst_w + (4 * slot + 0) * 32, r0);
                        cp_async_cell<8>(hst_w + (4 * slot + 1) * 32, r1);
                        cp_async_cell<8>(hst_w + (4 * slot + 2) * 32, r2);
                        cp_async_cell<8>(hst_w + (4 * slot + 3) * 32, r3);
                    }
                    cs::cp_async_commit();
                }
                next_request_plane();
            };
            auto advance = [&](int dz) {                   // the window moves up by dz planes (1 <= dz <= 8)
#pragma unroll 1
                for (int s = 0; s < dz; ++s) {
                    // (volatile loads inside the switch: the compiler must keep eight separate branches instead of turning the
                    //  choice of the ring slot into selects over all 32 window registers)
                    if (TMA && edge) {                     // plain loads of the entering plane (zr: nothing was requested ahead)
                        int e[4];
                        lane_cells(wcol & 0xffff, wcol >> 16, e);
                        const unsigned pb = (unsigned)(min(zr, nzloc - 1) * plane);
                        const float2 *a0 = u + (pb + (unsigned)e[0]), *a1 = u + (pb + (unsigned)e[1]);
                        const float2 *a2 = u + (pb + (unsigned)e[2]), *a3 = u + (pb + (unsigned)e[3]);
#define NUFFT_RING_ENTER(R_)                                                                                                      \
    {                                                                                                                             \
        asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(G[0][R_]) : "l"(a0));                                                    \
        asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(G[1][R_]) : "l"(a1));                                                    \
        asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(G[2][R_]) : "l"(a2));                                                    \
        asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(G[3][R_]) : "l"(a3));                                                    \
    }
                        NUFFT_RING_SWITCH(rot, NUFFT_RING_ENTER)
#undef NUFFT_RING_ENTER
                        next_request_plane();
                    } else {
                        unsigned ha, o0, o1, o2, o3;
                        if constexpr (TMA) {
                            mbar_wait(bar_w + 8 * si, (phase >> si) & 1u);
                            phase ^= 1u << si;
                            ha = (unsigned)__cvta_generic_to_shared((unsigned char *)hst_w + si * TMA_TILE_B);
                            o0 = ha + so0; o1 = ha + so1; o2 = ha + so2; o3 = ha + so3;
                        } else {
                            asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
                            ha = (unsigned)__cvta_generic_to_shared(hst_w + 4 * si * 32);
                            o0 = ha; o1 = ha + 256; o2 = ha + 512; o3 = ha + 768;
                        }
#define NUFFT_RING_ENTER(R_)                                                                                                      \
    {                                                                                                                             \
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(G[0][R_]) : "r"(o0));                                                       \
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(G[1][R_]) : "r"(o1));                                                       \
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(G[2][R_]) : "r"(o2));                                                       \
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(G[3][R_]) : "r"(o3));                                                       \
    }
                        NUFFT_RING_SWITCH(rot, NUFFT_RING_ENTER)
#undef NUFFT_RING_ENTER
                        if constexpr (TMA) __syncwarp();   // every lane has read the tile before the next box lands in it
                        request(si);
                    }
                    si = (si + 1 == PF) ? 0 : si + 1;
                    rot = (rot + 1) & 7;
                    ++zb;
                }
            };
            auto drain = [&]() {                           // requests of the previous window may be in flight
                if constexpr (TMA) {
#pragma unroll
                    for (int i = 0; i < PF; ++i)
                        if ((pending >> i) & 1u) { mbar_wait(bar_w + 8 * i, (phase >> i) & 1u); phase ^= 1u << i; }
                    pending = 0;
                    __syncwarp();
                } else {
                    cs::cp_async_wait0();
                }
            };
            auto load_all = [&](int cx, int cy) {          // new column: planes zb .. zb + 7 into slots 0 .. 7
                drain();
                rot = 0;
                si = 0;
                int e[4];
                lane_cells(cx, cy, e);
                zr = plane_of(zb, zlo, nzwrap);
                if (nzwrap != nzloc) zr = min(max(zr, 0), nzloc - RING);       // z slab: never read outside the stored planes
                const unsigned pb = (unsigned)(zr * plane);
                const float2 *a0 = u + (pb + (unsigned)e[0]), *a1 = u + (pb + (unsigned)e[1]);
                const float2 *a2 = u + (pb + (unsigned)e[2]), *a3 = u + (pb + (unsigned)e[3]);
                tc0 = 8 * cx - 8;
                tc1 = COL * cy - (M - 1);
                edge = cx == 0 || COL * cx + 8 > Nx || tc1 < 0 || COL * cy + 8 > Ny;
#pragma unroll
                for (int i = 0; i < RING; ++i) {
                    G[0][i] = cs::ldg_cell(a0); G[1][i] = cs::ldg_cell(a1); G[2][i] = cs::ldg_cell(a2); G[3][i] = cs::ldg_cell(a3);
                    a0 += plane; a1 += plane; a2 += plane; a3 += plane;
                    if (++zr == nzwrap) {
                        zr = 0;
                        const int64_t back = (int64_t)nzwrap * plane;
                        a0 -= back; a1 -= back; a2 -= back; a3 -= back;
                    }
                }
                if constexpr (!TMA) { r0 = a0; r1 = a1; r2 = a2; r3 = a3; }
                if (!(TMA && edge)) {
#pragma unroll
                    for (int i = 0; i < PF; ++i) request(i);
                }
            };

            auto issue_n = [&](int bi) {
                const int k = k0 + bi * BATCH + lane;
                if (k < k1) cs::cp_async_stream<4>(st_n, perm + k, pol);
            };
            auto issue_x = [&](int bi, int32_t n) {
                const int k = k0 + bi * BATCH + lane;
                if (k < k1) cs::cp_async_stream<16>(st_x, prec + n, pol);
            };
            cs::cp_async_wait0();
            issue_n(0);
            cs::cp_async_commit();
            cs::cp_async_wait0();
            int32_t n_nxt = *st_n;
            issue_x(0, n_nxt);
            issue_n(1);
            cs::cp_async_commit();

            for (int bi = 0; bi < nbatches; ++bi) {
                const int nb = min(BATCH, k1 - (k0 + bi * BATCH));
                cs::cp_async_wait0();
                const float4 xyz = *st_x;
                const int32_t n_mine = n_nxt;              // original index of point `lane` of the batch
                n_nxt = *st_n;
                issue_x(bi + 1, n_nxt);
                issue_n(bi + 2);
                cs::cp_async_commit();

                constexpr int CMD_NEW = 64;                // see ring_spread.cuh: what the window does before a point
                int mycol = -3, myz = 0, cmd = 0;
                if (lane < nb) {
                    int cx, cy, cz;
                    evaluate_point_interp(kp, cs_s, xyz.x, xyz.y, xyz.z, rec_w + lane * REC_I, cx, cy, cz);
                    mycol = ((cy >> 2) << 16) | (cx >> 2);
                    myz = cz - (M - 1);
                    key_w[lane] = make_int2(mycol, myz);
                }
                {
                    int pc = __shfl_up_sync(FULL, mycol, 1), pz = __shfl_up_sync(FULL, myz, 1);
                    if (lane == 0) { pc = wcol; pz = zb; }
                    const int dz = myz - pz;
                    cmd = (mycol != pc || dz < 0 || dz > RING) ? CMD_NEW : dz;
                }
                // rounds of HALF points share the lane-partial buffer: a run ends at a round boundary too
                const unsigned starts = __ballot_sync(FULL, lane < nb && (cmd != 0 || (lane & (HALF - 1)) == 0));
                __syncwarp();
                // the staged planes of the current window were waited for together with the batch staging (wait0 above):
                // from here on only plane groups (and one batch group) are outstanding

                IRec A;
                A.wx = rec_w[lo.vx];
                A.wx3 = rec_w[lo.vx3];
                A.wy = *reinterpret_cast<const float4 *>(rec_w + lo.wy);
                A.z0 = *reinterpret_cast<const float4 *>(rec_w + IOFF_WZ);
                A.z1 = *reinterpret_cast<const float4 *>(rec_w + IOFF_WZ + 4);
                const float *nxt = rec_w + REC_I;
                int p = 0;
                for (int h0 = 0; h0 < nb; h0 += HALF) {    // rounds of HALF points share the lane-partial buffer
                    const int hend = min(h0 + HALF, nb);
                    u64 *part = part_w + lane;
#pragma unroll 1
                    while (p < hend) {
                        const int c = __shfl_sync(FULL, cmd, p);
                        if (c == CMD_NEW) {
                            const int2 key = key_w[p];
                            wcol = key.x;
                            zb = key.y;
                            load_all(key.x & 0xffff, key.x >> 16);
                        } else if (c != 0) {
                            advance(c);
                        }
                        const unsigned after = p + 1 < 32 ? (starts >> (p + 1)) << (p + 1) : 0u;
                        const int pend = after ? min(__ffs(after) - 1, hend) : hend;
                        int len = pend - p;
                        p = pend;
#define NUFFT_RING_BODY(R_)                                                                                                       \
    _Pragma("unroll 1") do {                                                                                                      \
        interp_body<R_>(G, A, nxt, lo, part);                                                                                     \
        nxt += REC_I;                                                                                                             \
        part += PART_LD;                                                                                                          \
    } while (--len);
                        NUFFT_RING_SWITCH(rot, NUFFT_RING_BODY)
#undef NUFFT_RING_BODY
                    }
                    __syncwarp();
                    // ---- lanes 4q .. 4q + 3 add up the 32 lane partials of point h0 + q ------------------------------------------
                    {
                        const int q = lane >> 2, hh = lane & 3;
                        u64 sum = 0ull;
                        if (h0 + q < hend) {
                            const u64 *row = part_w + q * PART_LD + hh;
#pragma unroll
                            for (int j = 0; j < 8; ++j) sum = rt::fadd2(sum, row[4 * j]);
                        }
                        sum = rt::fadd2(sum, cs::shfl_xor_u64(sum, 1));
                        sum = rt::fadd2(sum, cs::shfl_xor_u64(sum, 2));
                        // lane h0 + q' stores point h0 + q': fetch its sum from lane 4 q'
                        const u64 res = cs::shfl_idx_u64(sum, (4 * (lane - h0)) & 31);
                        if (lane >= h0 && lane < hend) {
                            const float2 rv = unpk2(res);
                            const float scale = prefactor * (nu_weights ? nu_weights[n_mine] : 1.f);
                            __stcs(vc + n_mine, make_float2(rv.x * scale, rv.y * scale));          // streaming store: written once
                        }
                    }
                    __syncwarp();
                }
            }
            drain();                                       // no load may land in the tiles after the window is gone
        }
    }
    cs::cp_async_wait0();
}

inline size_t interp_smem_bytes(int cs_stride, bool tma)
{
    return (size_t)NWARP * (tma ? PF * TMA_TILE_B : PF * 4 * 32 * 8) + (size_t)NWARP * ((BATCH + 1) * REC_I + 5 * 32) * sizeof(float) +
           (size_t)NWARP * (HALF * PART_LD) * sizeof(u64) + (size_t)NWARP * (BATCH + 1) * sizeof(int2) + (size_t)NWARP * PF * sizeof(u64) +
           (size_t)(3 * cs_stride + 4) * sizeof(float) + 16;
}

}  // namespace ring
}  // namespace nufft
