// cs_spread.cuh — K-spread, column-streaming variant (3-D, HalfSupport(4), Float32 complex or real data): no shared-memory
// tile at all.
// Replaces src/spreading/gpu.jl:237-434 for this configuration class (same sums, different order).
//
// set_points orders the points by (z segment of 64 cells, column of 4 x 4 cells in (x, y), layer of 4 cells in z).  A
// warp walks through a chunk of that order.  All points of a column touch the same padded (x, y) footprint of 11 x 11
// cells, and consecutive layers overlap in 7 of their 11 z planes, so the warp keeps the footprint of the CURRENT layer
// (11 x 11 x 11 cells) in registers — lane L owns the columns rt::lane_slots(L), 4 columns x 11 planes = 44 packed
// (re, im) accumulators — and slides it upwards:
//   per point   8 shared-memory loads of the point's record (zero-padded 1-D weights, value), 6 FMUL2 and 44 FFMA2 per
//               lane: a_k = value x wx x wy for the lane's 4 columns, G[k][i] += a_k x wz[i].  Branch-free;
//   next layer  the 4 lowest planes are final for this warp: 16 red.global.add.v2.f32 per lane (rows of 11 consecutive
//               cells), the other 7 planes move down in the register file, 4 fresh planes start at zero.  About once per
//               8 points at one point per 8 fine cells;
//   new column / end of chunk   the whole window is added to the grid (44 reductions per lane) and restarts at zero.
// Partial sums of different warps meet only in global memory (L2 reductions); the chunk order keeps the set of
// concurrently active columns inside a slab of a few MB, so they meet in L2.  The kernel values are evaluated by one lane
// per point (batches of 32 points) into a warp-private record; permutation, folded coordinates (gathered through the
// permutation from the input-order records set_points keeps: no sorted copy is made) and values of the next batch arrive
// through cp.async (no registers or scoreboards held across a batch).  Shared memory holds only records, staging and
// the coefficient tables, so residency is bounded by registers (12 warps per SM).
#pragma once
#include "window_common.cuh"
#include "spread.cuh"

namespace nufft {
namespace cs {

using rt::u64;
using rt::pk2;
using rt::unpk2;
using rt::fmul2;
using rt::ffma2;
using rt::P;

constexpr int M = 4, W = 8;
constexpr int COL = 4;                    // column edge in x, y and layer thickness in z (cells)
constexpr int SEG = 256;                  // z segment (cells): bins are 4 x 4 x SEG cells, SEG / 4 layers each
constexpr int NWARP = 12;                 // warps per CTA (one CTA per SM): 168 registers per thread
constexpr int BATCH = 32;                 // points per evaluation batch: one lane per point
constexpr int CHUNK = 512;                // default points per work item (NUFFT_B200_CS_CHUNK overrides)
inline int chunk_points()
{
    if (const char *e = getenv("NUFFT_B200_CS_CHUNK")) { const int v = atoi(e); if (v >= 32 && v <= (1 << 20)) return v / 32 * 32; }
    return CHUNK;
}
constexpr int REC_F = 52;                 // floats per point record (208 bytes: 16-byte stores of 8 lanes are conflict-free)
constexpr int STAGE_F = 7 * 32;            // floats per warp of the global-load staging buffer: (x, y, z, -) record, value (2), index
constexpr int OFF_WX = 0;                 // [0..11]  wx_pad[0..10], 0
constexpr int OFF_HV = 12;                // [12..13] value (re, im) [spreading]
constexpr int OFF_WY = rt::OFF_WY;        // [16..39] wyT rows (rt::store_y)
constexpr int OFF_WZ = 40;                // [40..51] wz_pad[0..10], 0
static_assert(OFF_WY == 16, "record layout shared with window_common.cuh");

// 2M = 8 kernel values of dimension D around x; the piecewise-polynomial fast path reads the coefficient rows as two
// 16-byte vectors (same fmaf sequence as eval_kernel_values, bit-identical values).  Returns the 0-based cell.
template <int D>
__device__ __forceinline__ int eval_m4(const KernelParams<float> &kp, const float *cs, float x, float (&w)[W])
{
    if (kp.mode == NUFFT_EVAL_FAST && kp.kind != NUFFT_KERNEL_BSPLINE && kp.kind != NUFFT_KERNEL_GAUSSIAN) {
        float r;
        const int i0 = point_to_cell0<float>(x, kp.N[D], r);
        const float xt = 2.f * (r - (float)i0) - 1.f;
        const float4 *c4 = reinterpret_cast<const float4 *>(cs);
        float4 a = c4[2 * (M + 3)], b = c4[2 * (M + 3) + 1];
#pragma unroll
        for (int p = M + 2; p >= 0; --p) {
            const float4 ca = c4[2 * p], cb = c4[2 * p + 1];
            a.x = fmaf(xt, a.x, ca.x); a.y = fmaf(xt, a.y, ca.y); a.z = fmaf(xt, a.z, ca.z); a.w = fmaf(xt, a.w, ca.w);
            b.x = fmaf(xt, b.x, cb.x); b.y = fmaf(xt, b.y, cb.y); b.z = fmaf(xt, b.z, cb.z); b.w = fmaf(xt, b.w, cb.w);
        }
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        return i0;
    }
    return eval_kernel_values<float, M>(kp, cs, D, x, w);
}

// one lane evaluates the three 1-D kernels of its point into the point's record; returns the cells
__device__ __forceinline__ void evaluate_point(const KernelParams<float> &kp, const float *cs_s, float x, float y, float z,
                                               float *r, int &cx, int &cy, int &cz)
{
    float w[W], pw[P];
    float4 *q;
    cx = eval_m4<0>(kp, cs_s, x, w);
    rt::pad_shift(w, cx & 3, pw);
    q = reinterpret_cast<float4 *>(r + OFF_WX);
    q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
    q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
    q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
    cy = eval_m4<1>(kp, cs_s + kp.cs_stride, y, w);
    rt::pad_shift(w, cy & 3, pw);
    rt::store_y(r, pw);
    cz = eval_m4<2>(kp, cs_s + 2 * kp.cs_stride, z, w);
    rt::pad_shift(w, cz & 3, pw);
    q = reinterpret_cast<float4 *>(r + OFF_WZ);
    q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
    q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
    q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
}

// streamed / gathered operands (permutation, point records, values) are used once: fetch them with an evict-first L2
// policy so that they do not displace the grid lines that the reductions (spreading) and window loads (interpolation) reuse
__device__ __forceinline__ unsigned long long l2_evict_first_policy()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <int BYTES> __device__ __forceinline__ void cp_async_stream(void *smem, const void *gmem, unsigned long long pol)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "l"(pol) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem), "l"(pol) : "memory");
    else
        asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(sa), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void red_cell(float2 *p, u64 v)
{
    const float2 f = unpk2(v);
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(f.x), "f"(f.y) : "memory");
}
__device__ __forceinline__ void red_cell(float *p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// u64 registers per (x, y) column of the window: complex data one packed (re, im) cell per plane; real data two consecutive
// planes per register (planes 2j, 2j + 1; the 12th plane is padding and stays zero)
constexpr int REAL_LAUNCH_BOUND = 640;    // register cap (96) of the real-data spreading kernel, see cs_spread_kernel
template <bool CPLX> struct Win {
    static constexpr int NR = CPLX ? P : (P + 1) / 2;
};

struct PointRec {
    float2 hv;
    float4 wy;
    float wx, wx3;
    float4 z0, z1, z2;
};
__device__ __forceinline__ PointRec load_rec(const float *r, const rt::LaneSlots &ls)
{
    PointRec q;
    q.hv = *reinterpret_cast<const float2 *>(r + OFF_HV);
    q.wy = *reinterpret_cast<const float4 *>(r + OFF_WY + 4 * ls.row);
    q.wx = r[OFF_WX + ls.x];
    q.wx3 = r[OFF_WX + ls.x3];
    const float4 *zq = reinterpret_cast<const float4 *>(r + OFF_WZ);
    q.z0 = zq[0]; q.z1 = zq[1]; q.z2 = zq[2];
    return q;
}
// complex data: G[k][i] += value x (wx wy)_k x wz_i : 6 FMUL2 + 44 FFMA2
// real data:    G[k][j] += (value wx wy)_k x (wz_2j, wz_2j+1) : 4 FMUL2 + 24 FFMA2
template <bool CPLX>
__device__ __forceinline__ void accumulate(u64 (&G)[4][Win<CPLX>::NR], const PointRec &q)
{
    const u64 w01 = fmul2(pk2(q.wx, q.wx), pk2(q.wy.x, q.wy.y));
    const u64 w23 = fmul2(pk2(q.wx, q.wx3), pk2(q.wy.z, q.wy.w));
    if constexpr (CPLX) {
        const float2 wa = unpk2(w01), wb = unpk2(w23);
        const u64 v2 = pk2(q.hv.x, q.hv.y);
        const u64 a0 = fmul2(v2, pk2(wa.x, wa.x)), a1 = fmul2(v2, pk2(wa.y, wa.y));
        const u64 a2 = fmul2(v2, pk2(wb.x, wb.x)), a3 = fmul2(v2, pk2(wb.y, wb.y));
        const float wz[P] = {q.z0.x, q.z0.y, q.z0.z, q.z0.w, q.z1.x, q.z1.y, q.z1.z, q.z1.w, q.z2.x, q.z2.y, q.z2.z};
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const u64 wzz = pk2(wz[i], wz[i]);
            G[0][i] = ffma2(a0, wzz, G[0][i]);
            G[1][i] = ffma2(a1, wzz, G[1][i]);
            G[2][i] = ffma2(a2, wzz, G[2][i]);
            G[3][i] = ffma2(a3, wzz, G[3][i]);
        }
    } else {
        const u64 vv = pk2(q.hv.x, q.hv.x);
        const float2 wa = unpk2(fmul2(w01, vv)), wb = unpk2(fmul2(w23, vv));
        const u64 a0 = pk2(wa.x, wa.x), a1 = pk2(wa.y, wa.y), a2 = pk2(wb.x, wb.x), a3 = pk2(wb.y, wb.y);
        const u64 wz[6] = {pk2(q.z0.x, q.z0.y), pk2(q.z0.z, q.z0.w), pk2(q.z1.x, q.z1.y),
                           pk2(q.z1.z, q.z1.w), pk2(q.z2.x, q.z2.y), pk2(q.z2.z, q.z2.w)};      // z2.w = 0 (record padding)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            G[0][j] = ffma2(a0, wz[j], G[0][j]);
            G[1][j] = ffma2(a1, wz[j], G[1][j]);
            G[2][j] = ffma2(a2, wz[j], G[2][j]);
            G[3][j] = ffma2(a3, wz[j], G[3][j]);
        }
    }
}

template <bool CPLX>                       // complex / real Float32 data (one instantiation per translation unit)
// (no min-blocks hint: with it ptxas renames the accumulators and adds ~30 MOVs per point.  Real data: with the 168 registers
//  that 12 warps allow, ptxas writes 18 of the 24 accumulators to fresh registers and copies them back — 48 MOV per point, 38 %
//  of all instructions in the ncu capture; declaring a bound of 640 threads caps it at 96 registers, the loop is then 24 FFMA2
//  + 4 FMUL2 + 8 LDS + 4 MOV, no local memory inside it.  The launch still uses 12 warps.)
__global__ void __launch_bounds__(CPLX ? 32 * NWARP : REAL_LAUNCH_BOUND)
cs_spread_kernel(KernelParams<float> kp, TileGeom g, int np, int chunk, const int32_t *__restrict__ perm, int32_t *work_counter,
                 const float4 *__restrict__ prec, PtrPack vp, int C,
                 typename CellOf<float, CPLX>::type *__restrict__ us, int64_t ncells, const float *__restrict__ nu_weights)
{
    using Cell = typename CellOf<float, CPLX>::type;
    constexpr int NR = Win<CPLX>::NR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *rec_all = (float *)smem_raw;                                      // [NWARP][BATCH][REC_F]
    float *stage_all = rec_all + NWARP * BATCH * REC_F;                      // [NWARP][STAGE_F]
    float *cs_s = stage_all + NWARP * STAGE_F;                               // [3][cs_stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    float *rec_w = rec_all + warp * BATCH * REC_F;
    // staging slots of this lane (global loads land here through cp.async: no registers, no scoreboards held across a batch)
    float4 *st_x = reinterpret_cast<float4 *>(stage_all + warp * STAGE_F) + lane;      // folded (x, y, z, -) of the point
    float2 *st_v = reinterpret_cast<float2 *>(stage_all + warp * STAGE_F + 128) + lane;
    int32_t *st_n = reinterpret_cast<int32_t *>(stage_all + warp * STAGE_F + 192) + lane;

    for (int i = tid; i < 3 * kp.cs_stride; i += 32 * NWARP) cs_s[i] = kp.cs[i];
    __syncthreads();                                   // the only CTA barrier: coefficient tables

    const rt::LaneSlots ls = rt::lane_slots(lane);
    const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];
    const int plane = Nx * Ny;
    const unsigned long long pol = l2_evict_first_policy();

    u64 G[4][NR];
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
            const Cell *vc = (const Cell *)vp.p[c];
            Cell *u = us + (int64_t)c * ncells;
            // ---- window state ---------------------------------------------------------------------------------------
            int wcol = -1, wl = 0;                         // column id (cy << 16 | cx), layer
            int goff[4] = {0, 0, 0, 0};                    // cell offsets of the lane's 4 columns inside a z plane

            auto flush_planes = [&](auto i0c, auto i1c) {  // planes [i0, i1) of the window -> grid
                constexpr int I0 = decltype(i0c)::value, I1 = decltype(i1c)::value;
#pragma unroll
                for (int i = I0; i < I1; ++i) {
                    const int gz = wrap1(COL * wl - (M - 1) + i, Nz);
                    Cell *pl = u + (int64_t)gz * plane;
                    if constexpr (CPLX) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) red_cell(pl + goff[k], G[k][i]);
                        if (ls.has3) red_cell(pl + goff[3], G[3][i]);
                    } else {                               // plane i = half (i & 1) of register i / 2
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const float2 f = unpk2(G[k][i / 2]);
                            red_cell(pl + goff[k], (i & 1) ? f.y : f.x);
                        }
                        if (ls.has3) {
                            const float2 f = unpk2(G[3][i / 2]);
                            red_cell(pl + goff[3], (i & 1) ? f.y : f.x);
                        }
                    }
                }
            };
            auto flush_all = [&]() {
                flush_planes(std::integral_constant<int, 0>{}, std::integral_constant<int, P>{});
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int i = 0; i < NR; ++i) G[k][i] = 0ull;
            };
            auto shift_one = [&]() {                       // next layer: retire 4 planes, slide the other 7 down
                flush_planes(std::integral_constant<int, 0>{}, std::integral_constant<int, COL>{});
                constexpr int SH = CPLX ? COL : COL / 2;   // registers per layer
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int i = 0; i < NR - SH; ++i) G[k][i] = G[k][i + SH];
#pragma unroll
                    for (int i = NR - SH; i < NR; ++i) G[k][i] = 0ull;
                }
                ++wl;
            };

            // ---- global loads: one lane per point, staged through cp.async one batch ahead (the index two steps ahead:
            //      it is read back and used as the gather address of the value in the next step) --------------------------
            auto issue_n = [&](int bi) {
                const int k = k0 + bi * BATCH + lane;
                if (k < k1) cp_async_stream<4>(st_n, perm + k, pol);
            };
            auto issue_xv = [&](int bi, int32_t n) {
                const int k = k0 + bi * BATCH + lane;
                if (k < k1) {
                    cp_async_stream<16>(st_x, prec + n, pol);   // set_points keeps the folded coordinates in input order
                    cp_async_stream<(int)sizeof(Cell)>(st_v, vc + n, pol);
                }
            };
            issue_n(0);
            cp_async_commit();
            cp_async_wait0();
            int32_t n_cur = *st_n;                         // original index of this lane's point of the batch being staged
            float wgt = 1.f;
            issue_xv(0, n_cur);
            issue_n(1);
            cp_async_commit();

            for (int bi = 0; bi < nbatches; ++bi) {
                const int nb = min(BATCH, k1 - (k0 + bi * BATCH));
                cp_async_wait0();
                const float4 xyz = *st_x;
                const float x = xyz.x, y = xyz.y, z = xyz.z;
                float2 v = *st_v;                          // real data: .x only
                if constexpr (!CPLX) v.y = 0.f;
                if (nu_weights && lane < nb) wgt = nu_weights[n_cur];
                n_cur = *st_n;
                issue_xv(bi + 1, n_cur);
                issue_n(bi + 2);
                cp_async_commit();

                // ---- evaluate: one lane per point ----------------------------------------------------------------------
                int mycol = -1, mylay = 0;                 // lane p < nb: column id and layer of point p of the batch
                if (lane < nb) {
                    float *r = rec_w + lane * REC_F;
                    int cx, cy, cz;
                    evaluate_point(kp, cs_s, x, y, z, r, cx, cy, cz);
                    if (nu_weights) v = cmul(v, wgt);
                    *reinterpret_cast<float2 *>(r + OFF_HV) = v;
                    mycol = ((cy >> 2) << 16) | (cx >> 2);
                    mylay = cz >> 2;
                }
                // runs of points sharing the window: bit p of `starts` is set when point p opens a new (column, layer)
                unsigned starts;
                {
                    const int pc = __shfl_up_sync(FULL, mycol, 1), pl = __shfl_up_sync(FULL, mylay, 1);
                    starts = __ballot_sync(FULL, lane < nb && (lane == 0 || mycol != pc || mylay != pl));
                }
                __syncwarp();

                // ---- accumulate --------------------------------------------------------------------------------------
                for (int p0 = 0; p0 < nb;) {
                    const unsigned rest = starts & ~((2u << p0) - 1u);       // run starts after p0
                    const int p1 = rest ? __ffs(rest) - 1 : nb;
                    const int col = __shfl_sync(FULL, mycol, p0), lay = __shfl_sync(FULL, mylay, p0);
                    if (col != wcol || lay != wl) {                           // move the window (cold path)
                        const int d = lay - wl;
                        if (col == wcol && d > 0 && d < 3) {
                            shift_one();
                            if (d == 2) shift_one();
                        } else {
                            if (wcol >= 0) flush_all();
                            wcol = col;
                            wl = lay;
                            const int cx = col & 0xffff, cy = col >> 16;
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                goff[k] = wrap1(COL * cy - (M - 1) + ls.g + 3 * k, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x, Nx);
                            goff[3] = wrap1(COL * cy - (M - 1) + ls.y3, Ny) * Nx + wrap1(COL * cx - (M - 1) + ls.x3, Nx);
                        }
                    }
                    // hot loop: the points of the run only touch registers.  Rotated: the record of the next point is
                    // requested at the end of the body, into the registers the current one has just released
                    PointRec A = load_rec(rec_w + p0 * REC_F, ls);
#pragma unroll 1
                    for (int p = p0; p < p1; ++p) {
                        accumulate<CPLX>(G, A);
                        A = load_rec(rec_w + min(p + 1, p1 - 1) * REC_F, ls);
                    }
                    p0 = p1;
                }
                __syncwarp();
            }
            if (wcol >= 0) flush_all();
        }
    }
}

inline size_t spread_smem_bytes(int cs_stride)
{
    return (size_t)NWARP * (BATCH * REC_F + STAGE_F) * sizeof(float) + (size_t)(3 * cs_stride + 4) * sizeof(float) + 16;
}

}  // namespace cs
}  // namespace nufft
