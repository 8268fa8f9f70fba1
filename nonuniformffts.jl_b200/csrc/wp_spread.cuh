// wp_spread.cuh — K-spread, warp-private-tile variant (3-D, HalfSupport(4), ComplexF32): the default for the headline
// configuration class.  Replaces src/spreading/gpu.jl:237-434 (same sums, different order).
//
// Three-level accumulation, every level owned by exactly ONE warp, so nothing on the path needs an atomic or a CTA
// barrier until the final vector reduction to the oversampled grid:
//   registers   a 4 x 4 x 4-cell SUB-BIN's padded footprint (11 x 11 x 11 cells) lives in the warp's registers: lane L
//               owns the (x, y) columns rt::lane_slots(L) (4 columns x 11 planes = 44 packed (re, im) accumulators).
//               Points arrive sorted by (bin, sub-bin), so a run of points of the same sub-bin costs 9 shared-memory
//               loads + 2 FMUL2 + 44 FFMA2 per point and touches no tile memory (branch-free: the weights are
//               zero-padded to the footprint; a warp-uniform switch on the z offset would save 12 FFMA2 but ptxas
//               then renames the accumulators per case and pays it back in ~50 MOVs per point);
//   shared      the BIN's (8 x 8 x 8 cells) padded tile (15 x 15 x 15 cells) is private to the warp.  When the sub-bin
//               changes the window is stored back and the next one loaded (44 STS.64 + 44 LDS.64 per lane, about once
//               per 8 points at one point per 8 cells).  Rows are 16 cells wide and rotated by 11 y (mod 16): with the
//               lane map above every half-warp hits 16 distinct 8-byte bank pairs -> conflict-free;
//   global      at the end of the bin the tile is added to the oversampled grid with red.global.add.v4.f32 (2 complex
//               cells, periodic wrap per row) and re-zeroed.
// Warps pull (bin, chunk) work items from the device counter independently; a CTA is just 7 warps sharing an SM
// (7 tiles + records = 219 KiB of shared memory).  Kernel values are evaluated by 3 lanes per point (one per dimension,
// batches of 10 points) into a warp-private record: value x wz (8 complex), the lane-transposed zero-padded wy and the
// zero-padded wx of rt_common.cuh.
#pragma once
#include "rt_common.cuh"
#include "spread.cuh"

namespace nufft {
namespace wp {

using rt::u64;
using rt::pk2;
using rt::unpk2;
using rt::fmul2;
using rt::P;

constexpr int M = 4, W = 8;
constexpr int BIN = 8;                    // bin edge (cells): 2 x 2 x 2 sub-bins of 4 x 4 x 4 cells
constexpr int TE = BIN + W - 1;           // tile edge = 15
constexpr int ROW = 16;                   // physical row length (cells)
constexpr int PLANE = TE * ROW;           // 240
constexpr int TILE_CELLS = TE * PLANE;    // 3600 cells = 28 800 bytes (complex f32)
constexpr int NWARP = 7;
constexpr int BATCH = 10;                 // points per evaluation batch: 3 lanes per point
constexpr int REC_F = 64;                 // floats per point record
constexpr int OFF_WY = rt::OFF_WY;        // [16..39] wyT rows (rt::store_y)
constexpr int OFF_WX = rt::OFF_WX;        // [40..51] wx_pad[0..10], 0
constexpr int OFF_KEY = 12;               // [12]     sub-bin index
constexpr int OFF_S = 40;                 // [40..63] value x wz_pad[0..10]  (re, im), 0, 0     (wx_pad moves to [0..11])
constexpr int OFF_WXP = 0;
static_assert(OFF_WY == 16, "record layout shared with rt_common.cuh");

__device__ __forceinline__ void ffma2_acc(u64 &acc, u64 a, u64 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void lds64_to(u64 &v, unsigned saddr)
{
    asm volatile("ld.shared.b64 %0, [%1];" : "+l"(v) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void sts64(unsigned saddr, u64 v)
{
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(saddr), "l"(v) : "memory");
}

// cell index of tile coordinates (tx, ty) inside a plane: rows rotated by 11 ty
__device__ __forceinline__ int plane_cell(int tx, int ty) { return ty * ROW + ((tx + 11 * ty) & (ROW - 1)); }

struct PointRecord {
    float4 wy;
    float wx, wx3;
    float4 s[6];
    int key;
};

__device__ __forceinline__ PointRecord load_record(const float *r, const rt::LaneSlots &ls)
{
    PointRecord q;
    q.wy = *reinterpret_cast<const float4 *>(r + OFF_WY + 4 * ls.row);
    q.wx = r[OFF_WXP + ls.x];
    q.wx3 = r[OFF_WXP + ls.x3];
    const float4 *s = reinterpret_cast<const float4 *>(r + OFF_S);
#pragma unroll
    for (int h = 0; h < 6; ++h) q.s[h] = s[h];
    q.key = __float_as_int(r[OFF_KEY]);
    return q;
}

// per-lane kernel parameters of ONE dimension (lane = 3 * point + dimension)
__device__ __forceinline__ KernelParams<float> lane_kernel_params(const KernelParams<float> &kp, int d)
{
    KernelParams<float> kl = kp;
    kl.N[0] = d == 0 ? kp.N[0] : (d == 1 ? kp.N[1] : kp.N[2]);
    kl.beta[0] = d == 0 ? kp.beta[0] : (d == 1 ? kp.beta[1] : kp.beta[2]);
    kl.tau[0] = d == 0 ? kp.tau[0] : (d == 1 ? kp.tau[1] : kp.tau[2]);
    kl.dx[0] = d == 0 ? kp.dx[0] : (d == 1 ? kp.dx[1] : kp.dx[2]);
    return kl;
}

template <typename Inst>                   // instantiated only by the ComplexF32 translation unit
__global__ void __launch_bounds__(32 * NWARP, 1)
wp_spread_kernel(KernelParams<float> kp, TileGeom g, SmArgs a, const float *__restrict__ xs0, const float *__restrict__ xs1,
                 const float *__restrict__ xs2, PtrPack vp, int C, float2 *__restrict__ us, int64_t ncells,
                 const float *__restrict__ nu_weights)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tiles = (float2 *)smem_raw;                                      // [NWARP][TILE_CELLS]
    float *rec_all = (float *)(smem_raw + (size_t)NWARP * TILE_CELLS * sizeof(float2));   // [NWARP][BATCH][REC_F]
    float *cs_s = rec_all + NWARP * BATCH * REC_F;                           // [3][cs_stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const int total_items = a.item_start[a.nbins];
    float2 *tile = tiles + warp * TILE_CELLS;
    const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
    float *rec_w = rec_all + warp * BATCH * REC_F;

    for (int i = tid; i < 3 * kp.cs_stride; i += 32 * NWARP) cs_s[i] = kp.cs[i];
    for (int i = lane; i < TILE_CELLS; i += 32) tile[i] = make_float2(0.f, 0.f);
    __syncthreads();                                   // the only CTA barrier: coefficient tables

    const rt::LaneSlots ls = rt::lane_slots(lane);
    // evaluation role: lane = 3 * point + dimension
    const int ep = lane / 3, ed = lane - 3 * ep;
    const float *xs_d = ed == 0 ? xs0 : (ed == 1 ? xs1 : xs2);
    const KernelParams<float> kl = lane_kernel_params(kp, ed);
    const float *cs_d = cs_s + ed * kp.cs_stride;
    // flush role: 8 lanes per tile row (one 16-byte vector = cells 2j - 1, 2j), 4 rows per pass
    const int fj = lane & 7, fr = lane >> 3;
    const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];

    u64 G[4][P];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < P; ++i) G[k][i] = 0ull;

    // ---- work items: fetched one ahead ---------------------------------------------------------------------------
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
            const float2 *vc = (const float2 *)vp.p[c];
            // coordinates / values of the next batch are prefetched
            float xq = 0.f;
            float2 vq = make_float2(0.f, 0.f);
            auto prefetch = [&](int bi) {
                const int k = k0 + bi * BATCH + ep;
                if (bi < nbatches && lane < 3 * BATCH && k < k1) {
                    xq = xs_d[k];
                    if (ed == 2) {
                        const int32_t n = a.perm[k];
                        vq = vc[n];
                        if (nu_weights) vq = cmul(vq, nu_weights[n]);
                    }
                }
            };
            prefetch(0);
            int cur_sub = -1;
            unsigned win_s = tile_s;                       // shared address of the window's first plane
            int off[4] = {0, 0, 0, 0};                     // byte offsets of the lane's 4 columns inside a plane

            auto store_window = [&]() {
#pragma unroll
                for (int i = 0; i < P; ++i) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) sts64(win_s + off[k] + i * (PLANE * 8), G[k][i]);
                    if (ls.has3) sts64(win_s + off[3] + i * (PLANE * 8), G[3][i]);
                }
            };
            auto load_window = [&](int sub) {
                const int sz = sub & 1, sx = (sub >> 1) & 1, sy = sub >> 2;
#pragma unroll
                for (int k = 0; k < 3; ++k) off[k] = 8 * plane_cell(4 * sx + ls.x, 4 * sy + ls.g + 3 * k);
                off[3] = 8 * plane_cell(4 * sx + ls.x3, 4 * sy + ls.y3);
                win_s = tile_s + 4 * sz * (PLANE * 8);
#pragma unroll
                for (int i = 0; i < P; ++i) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) lds64_to(G[k][i], win_s + off[k] + i * (PLANE * 8));
                    if (ls.has3) lds64_to(G[3][i], win_s + off[3] + i * (PLANE * 8));
                }
            };

            for (int bi = 0; bi < nbatches; ++bi) {
                const int kb = k0 + bi * BATCH;
                const int nb = min(BATCH, k1 - kb);
                const float x = xq;
                const float2 v = vq;
                prefetch(bi + 1);

                // ---- evaluate: 3 lanes per point -----------------------------------------------------------------
                const bool act = lane < 3 * BATCH && ep < nb;
                int t = 0;
                if (act) {
                    float *r = rec_w + ep * REC_F;
                    float w[W];
                    t = eval_kernel_values<float, M>(kl, cs_d, 0, x, w) - org_d;
                    float pw[P];
                    rt::pad_shift(w, t & 3, pw);
                    if (ed == 2) {
                        float4 *q = reinterpret_cast<float4 *>(r + OFF_S);
#pragma unroll
                        for (int i = 0; i < 5; ++i)
                            q[i] = make_float4(v.x * pw[2 * i], v.y * pw[2 * i], v.x * pw[2 * i + 1], v.y * pw[2 * i + 1]);
                        q[5] = make_float4(v.x * pw[10], v.y * pw[10], 0.f, 0.f);
                    } else {
                        if (ed == 0) {
                            float4 *q = reinterpret_cast<float4 *>(r + OFF_WXP);
                            q[0] = make_float4(pw[0], pw[1], pw[2], pw[3]);
                            q[1] = make_float4(pw[4], pw[5], pw[6], pw[7]);
                            q[2] = make_float4(pw[8], pw[9], pw[10], 0.f);
                        } else {
                            rt::store_y(r, pw);
                        }
                    }
                }
                {
                    const int src = min(3 * ep, 27);
                    const int t0 = __shfl_sync(FULL, t, src), t1 = __shfl_sync(FULL, t, src + 1), t2 = __shfl_sync(FULL, t, src + 2);
                    if (act && ed == 0) {
                        const int key = ((((t1 >> 2) << 1) | (t0 >> 2)) << 1) | (t2 >> 2);
                        rec_w[ep * REC_F + OFF_KEY] = __int_as_float(key);
                    }
                }
                __syncwarp();

                // ---- accumulate: record of point p + 1 is loaded while point p is accumulated -------------------
                PointRecord cur = load_record(rec_w, ls);
                for (int p = 0; p < nb; ++p) {
                    const PointRecord nxt = load_record(rec_w + min(p + 1, nb - 1) * REC_F, ls);
                    const int sub = cur.key;
                    if (sub != cur_sub) {
                        if (cur_sub >= 0) { store_window(); __syncwarp(); }
                        load_window(sub);
                        cur_sub = sub;
                    }
                    const u64 w01 = fmul2(pk2(cur.wx, cur.wx), pk2(cur.wy.x, cur.wy.y));
                    const u64 w23 = fmul2(pk2(cur.wx, cur.wx3), pk2(cur.wy.z, cur.wy.w));
                    const float2 wa = unpk2(w01), wb = unpk2(w23);
                    const u64 a0 = pk2(wa.x, wa.x), a1 = pk2(wa.y, wa.y), a2 = pk2(wb.x, wb.x), a3 = pk2(wb.y, wb.y);
#pragma unroll
                    for (int h = 0; h < 6; ++h) {
                        const u64 se = pk2(cur.s[h].x, cur.s[h].y), so = pk2(cur.s[h].z, cur.s[h].w);
                        ffma2_acc(G[0][2 * h], a0, se);
                        ffma2_acc(G[1][2 * h], a1, se);
                        ffma2_acc(G[2][2 * h], a2, se);
                        ffma2_acc(G[3][2 * h], a3, se);
                        if (2 * h + 1 < P) {
                            ffma2_acc(G[0][2 * h + 1], a0, so);
                            ffma2_acc(G[1][2 * h + 1], a1, so);
                            ffma2_acc(G[2][2 * h + 1], a2, so);
                            ffma2_acc(G[3][2 * h + 1], a3, so);
                        }
                    }
                    cur = nxt;
                }
                __syncwarp();
            }
            if (cur_sub >= 0) store_window();
            __syncwarp();

            // ---- flush: tile -> oversampled grid (periodic), 16-byte vector reductions; re-zero the tile -------------
            {
                float2 *u = us + (int64_t)c * ncells;
                int gx0 = org0 - (M - 1) - 1 + 2 * fj;                 // even: 16-byte aligned pair (cells 2j - 1, 2j)
                gx0 = wrap1(gx0, Nx);
                for (int rb = 0; rb < TE * TE; rb += 4) {
                    const int row = rb + fr;
                    if (row < TE * TE) {
                        const int tz = row / TE, ty = row - TE * tz;
                        const int rot = 11 * ty + 2 * fj;
                        const unsigned pl = tile_s + 8 * (row * ROW + ((rot - 1) & (ROW - 1)));
                        const unsigned ph = tile_s + 8 * (row * ROW + (rot & (ROW - 1)));
                        u64 lo = 0ull, hi = 0ull;
                        lds64_to(lo, pl);
                        lds64_to(hi, ph);
                        sts64(pl, 0ull);
                        sts64(ph, 0ull);
                        const int gy = wrap1(org1 - (M - 1) + ty, Ny), gz = wrap1(org2 - (M - 1) + tz, Nz);
                        const float2 l2 = unpk2(lo), h2 = unpk2(hi);
                        atomicAdd(reinterpret_cast<float4 *>(u + ((int64_t)gz * Ny + gy) * Nx + gx0),
                                  make_float4(l2.x, l2.y, h2.x, h2.y));
                    }
                }
            }
            __syncwarp();
        }
    }
}

inline size_t spread_smem_bytes(int cs_stride)
{
    return (size_t)NWARP * (TILE_CELLS * sizeof(float2) + BATCH * REC_F * sizeof(float)) + (size_t)(3 * cs_stride + 4) * sizeof(float) + 16;
}

}  // namespace wp
}  // namespace nufft
