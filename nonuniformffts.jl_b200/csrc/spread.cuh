// spread.cuh — K-spread: type-1 spreading kernels (templates; instantiated in spread_inst.cu).
//
// Replaces src/spreading/gpu.jl:2-127 (global-memory kernel) and :237-434 (shared-memory kernel),
// plus the preceding fill_with_zeros_kernel! (src/NonuniformFFTs.jl:116-122,161-167).
//
// Shared-memory kernel (sm_100a), v2:
//   * persistent CTAs pull work items (bin, chunk of <= `chunk` sorted points) from a device counter; the next
//     item is prefetched while the current one is processed;
//   * the bin's padded subgrid tile (B + 2M - 1 per dim) lives in dynamic shared memory (up to 227 KiB);
//   * warp specialisation: NPROD producer warps evaluate the 1-D kernel values of batch b+1 into a
//     double-buffered per-point record while the M consumer warps accumulate batch b (one CTA barrier per batch,
//     consumers never wait on global memory);
//   * consumers: the tile is partitioned among the M consumer warps by residue class of the slowest tile
//     coordinate modulo M.  A point's support spans 2M consecutive planes, so it meets every residue class
//     exactly twice: perfect balance, exactly one owning warp per tile cell -> plain read-modify-write, no
//     shared-memory floating-point atomics (CAS loops on this architecture) and no CTA barrier per point as in
//     the reference.  Lanes cover (2M consecutive x cells) x (32/2M rows): conflict-free row segments;
//   * flush: tile -> oversampled grid with 16-byte vector reductions (red.global.add.v4.f32 = 2 complex cells),
//     periodic wrap applied per row/column, the tile is re-zeroed in the same pass.
#pragma once
#include "tile_common.cuh"

namespace nufft {

// ---------------------------------------------------------------------------------------------
// Global-memory method: one thread per (sorted) point, (2M)^D vector atomics per component.
// ---------------------------------------------------------------------------------------------
template <typename T, bool CPLX, int D, int M>
__global__ void __launch_bounds__(128)
spread_gm_kernel(KernelParams<T> kp, int64_t np, const T *__restrict__ xs0, const T *__restrict__ xs1,
                 const T *__restrict__ xs2, const int32_t *__restrict__ perm, PtrPack vp, int C,
                 typename CellOf<T, CPLX>::type *__restrict__ us, int64_t ncells, const T *__restrict__ nu_weights)
{
    using Cell = typename CellOf<T, CPLX>::type;
    constexpr int W = 2 * M;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np) return;
    T wx[W], wy[D > 1 ? W : 1], wz[D > 2 ? W : 1];
    int ix, iy = 0, iz = 0;
    ix = eval_kernel_values<T, M>(kp, kp.cs, 0, xs0[k], wx) - (M - 1);
    if (D > 1) iy = eval_kernel_values<T, M>(kp, kp.cs + kp.cs_stride, 1, xs1[k], wy) - (M - 1);
    if (D > 2) iz = eval_kernel_values<T, M>(kp, kp.cs + 2 * kp.cs_stride, 2, xs2[k], wz) - (M - 1);
    const int Nx = kp.N[0], Ny = kp.N[1], Nz = kp.N[2];
    if (ix < 0) ix += Nx;
    if (D > 1 && iy < 0) iy += Ny;
    if (D > 2 && iz < 0) iz += Nz;
    const int32_t n = perm[k];
    const T wgt = nu_weights ? nu_weights[n] : (T)1;
    for (int c = 0; c < C; ++c) {
        const Cell v = cmul(load_value<T, CPLX>(vp.p[c], n), wgt);
        Cell *u = us + (int64_t)c * ncells;
        int gz = iz;
        for (int jz = 0; jz < (D > 2 ? W : 1); ++jz) {
            int gy = iy;
            const T wzv = D > 2 ? wz[jz] : (T)1;
            for (int jy = 0; jy < (D > 1 ? W : 1); ++jy) {
                const T wyz = D > 1 ? wy[jy] * wzv : (T)1;
                Cell *row = u + ((int64_t)gz * Ny + gy) * Nx;
                int gx = ix;
#pragma unroll
                for (int jx = 0; jx < W; ++jx) {
                    catomic_add(row + gx, cmul(v, wx[jx] * wyz));
                    gx = (gx + 1 == Nx) ? 0 : gx + 1;
                }
                gy = (gy + 1 == Ny) ? 0 : gy + 1;
            }
            gz = (gz + 1 == Nz) ? 0 : gz + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory method
// ---------------------------------------------------------------------------------------------
constexpr int SPREAD_NPROD = 2;     // producer warps per CTA

// 16-byte vector reduction to global memory where the hardware has one (f32 only, sm_90+)
template <typename Cell> struct FlushVec { static constexpr int VEC = 1; };
template <> struct FlushVec<float> { static constexpr int VEC = 4; };
template <> struct FlushVec<float2> { static constexpr int VEC = 2; };

__device__ __forceinline__ void red_vec(float *p, const float *v)   // 4 consecutive f32 cells, 16-byte aligned
{
    atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
}
__device__ __forceinline__ void red_vec(float2 *p, const float2 *v)  // 2 consecutive complex cells
{
    atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v[0].x, v[0].y, v[1].x, v[1].y));
}
__device__ __forceinline__ void red_vec(double *, const double *) {}
__device__ __forceinline__ void red_vec(double2 *, const double2 *) {}

template <typename T, bool CPLX, int D, int M>
__global__ void __launch_bounds__(32 * (M + SPREAD_NPROD))
spread_sm_kernel(KernelParams<T> kp, TileGeom g, SmArgs a, const T *__restrict__ xs0, const T *__restrict__ xs1,
                 const T *__restrict__ xs2, PtrPack vp, int C, typename CellOf<T, CPLX>::type *__restrict__ us,
                 int64_t ncells, const T *__restrict__ nu_weights)
{
    using Cell = typename CellOf<T, CPLX>::type;
    using LM = LaneMap<M>;
    using REC = WRecord<D, M, true>;
    constexpr int W = 2 * M;
    constexpr int NT = 32 * (M + SPREAD_NPROD);
    constexpr int NWARP = M + SPREAD_NPROD;
    constexpr int NPT = 32 * SPREAD_NPROD;       // producer threads
    constexpr int G = LM::G, NI = LM::NI;
    constexpr int RS = REC::SIZE;
    constexpr int VEC = FlushVec<Cell>::VEC;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_bytes = (g.tile_cells * (int)sizeof(Cell) + 15) & ~15;
    Cell *tile = (Cell *)smem_raw;
    int4 *st_s = (int4 *)(smem_raw + tile_bytes);                     // [2][batch]
    Cell *v_s = (Cell *)(st_s + 2 * g.batch);                         // [2][batch]
    T *rec_s = (T *)(v_s + 2 * g.batch);                              // [2][batch][RS]
    T *cs_s = rec_s + 2 * g.batch * RS;                               // [D][cs_stride]
    __shared__ int s_item[2][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool consumer = warp < M;
    const int ptid = tid - 32 * M;               // producer thread id (>= 0 for producers)
    const int lx = lane % W, lg = lane / W;
    const bool lane_on = lg < G;
    const int Tx = g.T[0], Ty = g.T[1], Tz = g.T[2], Sx = g.S[0], S2 = g.S[2];
    const int total_items = a.item_start[a.nbins];

    for (int i = tid; i < D * kp.cs_stride; i += NT) cs_s[i] = kp.cs[i];
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
            // ---- producer: one thread per point: kernel values, local indices and value --------------
            auto produce = [&](int bi) {
                const int kb = k0 + bi * g.batch;
                const int nb = min(g.batch, k1 - kb);
                int4 *st_b = st_s + (bi & 1) * g.batch;
                Cell *v_b = v_s + (bi & 1) * g.batch;
                T *rec_b = rec_s + (bi & 1) * g.batch * RS;
                for (int p = ptid; p < nb; p += NPT) {
                    // all global loads first (independent), then the dependent value gather
                    const int32_t n = a.perm[kb + p];
                    const T x0 = xs0[kb + p];
                    const T x1 = D > 1 ? xs1[kb + p] : (T)0;
                    const T x2 = D > 2 ? xs2[kb + p] : (T)0;
                    Cell v = load_value<T, CPLX>(vp.p[c], n);
                    if (nu_weights) v = cmul(v, nu_weights[n]);
                    int4 st = make_int4(0, 0, 0, n);
                    T w[W];
                    st.x = eval_kernel_values<T, M>(kp, cs_s, 0, x0, w) - org0;
                    REC::store(rec_b + p * RS, 0, w, st.x);
                    if (D > 1) {
                        st.y = eval_kernel_values<T, M>(kp, cs_s + kp.cs_stride, 1, x1, w) - org1;
                        REC::store(rec_b + p * RS, 1, w, st.y);
                    }
                    if (D > 2) {
                        st.z = eval_kernel_values<T, M>(kp, cs_s + 2 * kp.cs_stride, 2, x2, w) - org2;
                        REC::store(rec_b + p * RS, 2, w, st.z);
                    }
                    st_b[p] = st;
                    v_b[p] = v;
                }
            };
            if (!consumer) produce(0);
            __syncthreads();
            for (int bi = 0; bi < nbatches; ++bi) {
                if (!consumer) {
                    if (bi + 1 < nbatches) produce(bi + 1);
                } else {
                    // ---- consumer: accumulate the batch into the tile (software pipelined: the record of
                    //      point p+1 is loaded while point p is accumulated) --------------------------------
                    const int nb = min(g.batch, k1 - (k0 + bi * g.batch));
                    const int4 *st_b = st_s + (bi & 1) * g.batch;
                    const Cell *v_b = v_s + (bi & 1) * g.batch;
                    const T *rec_b = rec_s + (bi & 1) * g.batch * RS;
                    if constexpr (D == 3) {
                        const int dS2 = M * S2, dSx = G * Sx;
                        const int lane_off = lg * Sx + lx;
                        int4 st = st_b[0];
                        Cell v = v_b[0];
                        T wx = rec_b[lx];
                        T wy[NI], wz[2];
                        VecLoad<T, NI>::load(rec_b + REC::OFF_Y + lg * NI, wy);
                        VecLoad<T, 2>::load(rec_b + REC::OFF_Z + warp * 2, wz);
                        for (int p = 0; p < nb; ++p) {
                            // prefetch the next record (clamped: the last one is re-read, harmless)
                            const int pn = min(p + 1, nb - 1);
                            const T *wn = rec_b + pn * RS;
                            const int4 st_n = st_b[pn];
                            const Cell v_n = v_b[pn];
                            const T wx_n = wn[lx];
                            T wy_n[NI], wz_n[2];
                            VecLoad<T, NI>::load(wn + REC::OFF_Y + lg * NI, wy_n);
                            VecLoad<T, 2>::load(wn + REC::OFF_Z + warp * 2, wz_n);
                            // accumulate point p
                            const int r = pmod(warp - st.z, M);             // first owned plane of the support
                            const Cell vx = cmul(v, wx);
                            Cell *c0 = tile + (st.z + r) * S2 + st.y * Sx + st.x + lane_off;
                            if (lane_on) {
                                constexpr int CH = (2 * NI <= 8) ? NI : (NI < 8 ? NI : 8);   // rows per register chunk
                                constexpr int TP = (2 * NI <= 8) ? 2 : 1;                    // planes per chunk
#pragma unroll
                                for (int tb = 0; tb < 2; tb += TP) {
#pragma unroll
                                    for (int ib = 0; ib < NI; ib += CH) {
                                        Cell acc[TP][CH];
#pragma unroll
                                        for (int t2 = 0; t2 < TP; ++t2)
#pragma unroll
                                            for (int i = 0; i < CH; ++i)
                                                if (ib + i < NI && lg + (ib + i) * G < W)
                                                    acc[t2][i] = c0[(tb + t2) * dS2 + (ib + i) * dSx];
#pragma unroll
                                        for (int t2 = 0; t2 < TP; ++t2)
#pragma unroll
                                            for (int i = 0; i < CH; ++i)
                                                if (ib + i < NI && lg + (ib + i) * G < W)
                                                    cfma(acc[t2][i], vx, wy[ib + i] * wz[tb + t2]);
#pragma unroll
                                        for (int t2 = 0; t2 < TP; ++t2)
#pragma unroll
                                            for (int i = 0; i < CH; ++i)
                                                if (ib + i < NI && lg + (ib + i) * G < W)
                                                    c0[(tb + t2) * dS2 + (ib + i) * dSx] = acc[t2][i];
                                    }
                                }
                            }
                            __syncwarp();
                            st = st_n; v = v_n; wx = wx_n;
#pragma unroll
                            for (int i = 0; i < NI; ++i) wy[i] = wy_n[i];
                            wz[0] = wz_n[0]; wz[1] = wz_n[1];
                        }
                    } else if constexpr (D == 2) {
                        for (int p = 0; p < nb; ++p) {
                            const int4 st = st_b[p];
                            const T *wp = rec_b + p * RS;
                            const Cell v = v_b[p];
                            const int r = pmod(warp - st.y, M);
                            T wy[2];
                            VecLoad<T, 2>::load(wp + REC::OFF_Y + warp * 2, wy);
                            const Cell vx = cmul(v, wp[lx]);
                            Cell *c0 = tile + (st.y + r) * Sx + st.x + lx;
                            for (int t2 = lg; t2 < 2; t2 += G) {
                                if (lane_on) {
                                    Cell acc = c0[t2 * M * Sx];
                                    cfma(acc, vx, wy[t2 & 1]);
                                    c0[t2 * M * Sx] = acc;
                                }
                            }
                            __syncwarp();
                        }
                    } else {
                        for (int p = 0; p < nb; ++p) {
                            const int4 st = st_b[p];
                            const T *wp = rec_b + p * RS;
                            const Cell v = v_b[p];
                            const int r = pmod(warp - st.x, M);
                            if (lane < 2) {
                                Cell acc = tile[st.x + r + lane * M];
                                cfma(acc, v, wp[warp * 2 + lane]);
                                tile[st.x + r + lane * M] = acc;
                            }
                            __syncwarp();
                        }
                    }
                }
                __syncthreads();
            }
            // ---- flush: tile -> global grid (periodic), vector reductions; re-zero the tile ------------
            // (host guarantees T_d <= N_d, so a single conditional wrap per coordinate suffices).
            // Half-warps take one tile row each; a lane owns one 16-byte vector slot of the row, so everything
            // that depends on x is computed once per work item.
            {
                Cell *u = us + (int64_t)c * ncells;
                const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];
                const int x0 = org0 - (M - 1), y0 = D > 1 ? org1 - (M - 1) : 0, z0 = D > 2 ? org2 - (M - 1) : 0;
                const Cell zero = cell_zero((Cell *)nullptr);
                const bool vec_ok = VEC > 1 && (Nx % VEC) == 0;
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
                            const int gz = D > 2 ? wrap1(z0 + z, Nz) : 0;
                            Cell *gplane = u + (int64_t)gz * Ny * Nx + gx;
                            Cell *tplane = tile + z * S2 + xt;
                            for (int y = 2 * warp + hrow; y < Ty; y += 2 * NWARP) {
                                const int gy = D > 1 ? wrap1(y0 + y, Ny) : 0;
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
                        const int gz = D > 2 ? wrap1(z0 + z, Nz) : 0;
                        for (int y = warp; y < Ty; y += NWARP) {
                            const int gy = D > 1 ? wrap1(y0 + y, Ny) : 0;
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

// host-side launchers (defined per (T, CPLX) translation unit)
template <typename T, bool CPLX> int spread_dispatch(Plan &p, const void *const vp[], const nufft_callbacks *cb);

}  // namespace nufft
