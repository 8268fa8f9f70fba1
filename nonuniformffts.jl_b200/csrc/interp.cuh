// interp.cuh — K-interp: type-2 interpolation kernels (templates; instantiated in interp_inst.cu).
//
// Replaces src/interpolation/gpu.jl:4-38,124-193 (global-memory kernel) and :211-395 (shared-memory
// kernel).  Values are scaled by prod_d dx_d (src/interpolation/gpu.jl:55-56).
//
// Shared-memory kernel (sm_100a), v2: persistent CTAs pull (bin, chunk) work items; the bin's padded subgrid
// tile is staged in dynamic shared memory (16-byte vector loads, periodic wrap applied per row/column);
// producer warps evaluate the kernel values of batch b+1 while the consumer warps process batch b; one WARP per
// point computes the (2M)^D dot product with lanes laid out as (2M consecutive x cells) x (32/2M rows) —
// conflict-free row segments — and reduces with warp shuffles; lane 0 scatters the result through the
// permutation (original point index carried in the per-point record).
#pragma once
#include "tile_common.cuh"

namespace nufft {

constexpr int INTERP_THREADS = 256;
constexpr int INTERP_NPROD = 2;

template <typename T, bool CPLX, int D, int M>
__global__ void __launch_bounds__(128)
interp_gm_kernel(KernelParams<T> kp, int64_t np, const T *__restrict__ xs0, const T *__restrict__ xs1,
                 const T *__restrict__ xs2, const int32_t *__restrict__ perm, MutPtrPack vp, int C,
                 const typename CellOf<T, CPLX>::type *__restrict__ us, int64_t ncells, T prefactor,
                 const T *__restrict__ nu_weights)
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
    const T scale = prefactor * (nu_weights ? nu_weights[n] : (T)1);
    for (int c = 0; c < C; ++c) {
        const Cell *u = us + (int64_t)c * ncells;
        Cell acc = cell_zero((Cell *)nullptr);
        int gz = iz;
        for (int jz = 0; jz < (D > 2 ? W : 1); ++jz) {
            int gy = iy;
            const T wzv = D > 2 ? wz[jz] : (T)1;
            for (int jy = 0; jy < (D > 1 ? W : 1); ++jy) {
                const T wyz = D > 1 ? wy[jy] * wzv : (T)1;
                const Cell *row = u + ((int64_t)gz * Ny + gy) * Nx;
                int gx = ix;
#pragma unroll
                for (int jx = 0; jx < W; ++jx) {
                    cfma(acc, row[gx], wx[jx] * wyz);
                    gx = (gx + 1 == Nx) ? 0 : gx + 1;
                }
                gy = (gy + 1 == Ny) ? 0 : gy + 1;
            }
            gz = (gz + 1 == Nz) ? 0 : gz + 1;
        }
        store_value<T, CPLX>(vp.p[c], n, cmul(acc, scale));
    }
}

// 16-byte vector load of VEC consecutive cells (f32 types); scalar otherwise
template <typename Cell> struct TileVec { static constexpr int VEC = 1; };
template <> struct TileVec<float> { static constexpr int VEC = 4; };
template <> struct TileVec<float2> { static constexpr int VEC = 2; };
template <> struct TileVec<double> { static constexpr int VEC = 2; };

__device__ __forceinline__ void ldg_vec(const float *p, float *v)
{
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ldg_vec(const float2 *p, float2 *v)
{
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = make_float2(t.x, t.y); v[1] = make_float2(t.z, t.w);
}
__device__ __forceinline__ void ldg_vec(const double *p, double *v)
{
    const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
    v[0] = t.x; v[1] = t.y;
}
__device__ __forceinline__ void ldg_vec(const double2 *p, double2 *v) { v[0] = __ldg(p); }

template <typename T, bool CPLX, int D, int M>
__global__ void __launch_bounds__(INTERP_THREADS)
interp_sm_kernel(KernelParams<T> kp, TileGeom g, SmArgs a, const T *__restrict__ xs0, const T *__restrict__ xs1,
                 const T *__restrict__ xs2, MutPtrPack vp, int C, const typename CellOf<T, CPLX>::type *__restrict__ us,
                 int64_t ncells, T prefactor, const T *__restrict__ nu_weights)
{
    using Cell = typename CellOf<T, CPLX>::type;
    using LM = LaneMap<M>;
    using REC = WRecord<D, M, false>;
    constexpr int W = 2 * M;
    constexpr int NT = INTERP_THREADS;
    constexpr int NWARP = NT / 32;
    constexpr int NCONS = NWARP - INTERP_NPROD;
    constexpr int NPT = 32 * INTERP_NPROD;
    constexpr int G = LM::G, NI = LM::NI;
    constexpr int RS = REC::SIZE;
    constexpr int VEC = TileVec<Cell>::VEC;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_bytes = (g.tile_cells * (int)sizeof(Cell) + 15) & ~15;
    Cell *tile = (Cell *)smem_raw;
    int4 *st_s = (int4 *)(smem_raw + tile_bytes);                     // [2][batch]: local starts + original index
    Cell *v_s = (Cell *)(st_s + 2 * g.batch);                         // [2][batch]: .x = output scale of the point
    T *rec_s = (T *)(v_s + 2 * g.batch);                              // [2][batch][RS]
    T *cs_s = rec_s + 2 * g.batch * RS;
    __shared__ int s_item[2][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool consumer = warp < NCONS;
    const int ptid = tid - 32 * NCONS;
    const int lx = lane % W, lg = lane / W;
    const bool lane_on = lg < G;
    const int Tx = g.T[0], Ty = g.T[1], Tz = g.T[2], Sx = g.S[0], S2 = g.S[2];
    const int total_items = a.item_start[a.nbins];

    for (int i = tid; i < D * kp.cs_stride; i += NT) cs_s[i] = kp.cs[i];
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
        const int nbatches = (k1 - k0 + g.batch - 1) / g.batch;

        for (int c = 0; c < C; ++c) {
            // ---- producer: one thread per point ---------------------------------------------------------
            auto produce = [&](int bi) {
                const int kb = k0 + bi * g.batch;
                const int nb = min(g.batch, k1 - kb);
                int4 *st_b = st_s + (bi & 1) * g.batch;
                Cell *v_b = v_s + (bi & 1) * g.batch;
                T *rec_b = rec_s + (bi & 1) * g.batch * RS;
                for (int p = ptid; p < nb; p += NPT) {
                    const int32_t n = a.perm[kb + p];
                    const T x0 = xs0[kb + p];
                    const T x1 = D > 1 ? xs1[kb + p] : (T)0;
                    const T x2 = D > 2 ? xs2[kb + p] : (T)0;
                    const T scale = prefactor * (nu_weights ? nu_weights[n] : (T)1);
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
                    reinterpret_cast<T *>(v_b + p)[0] = scale;
                }
            };
            if (!consumer) produce(0);
            // ---- stage the padded tile with asynchronous copies (cp.async / LDGSTS, one cell per lane; periodic
            //      wrap per row/column; host guarantees T_d <= N_d).  All copies of a thread are in flight at once;
            //      the producers' kernel evaluation above overlaps with them. -------------------------------------
            {
                const Cell *u = us + (int64_t)c * ncells;
                const int Nx = g.N[0], Ny = g.N[1], Nz = g.N[2];
                const int x0 = org0 - (M - 1), y0 = D > 1 ? org1 - (M - 1) : 0, z0 = D > 2 ? org2 - (M - 1) : 0;
                for (int xb = 0; xb < Tx; xb += 32) {
                    const int x = xb + lane;
                    const bool in = x < Tx;
                    const int gx = wrap1(x0 + (in ? x : 0), Nx);
                    for (int z = 0; z < Tz; ++z) {
                        const int gz = D > 2 ? wrap1(z0 + z, Nz) : 0;
                        const Cell *gplane = u + (int64_t)gz * Ny * Nx + gx;
                        Cell *tplane = tile + z * S2 + x;
                        for (int y = warp; y < Ty; y += NWARP) {
                            const int gy = D > 1 ? wrap1(y0 + y, Ny) : 0;
                            if (in) cp_async_cell<(int)sizeof(Cell)>(tplane + y * Sx, gplane + (int64_t)gy * Nx);
                        }
                    }
                }
                cp_async_wait_all();
            }
            __syncthreads();
            for (int bi = 0; bi < nbatches; ++bi) {
                if (!consumer) {
                    if (bi + 1 < nbatches) produce(bi + 1);
                } else {
                    const int nb = min(g.batch, k1 - (k0 + bi * g.batch));
                    const int4 *st_b = st_s + (bi & 1) * g.batch;
                    const Cell *v_b = v_s + (bi & 1) * g.batch;
                    const T *rec_b = rec_s + (bi & 1) * g.batch * RS;
                    for (int p = warp; p < nb; p += NCONS) {
                        const int4 st = st_b[p];
                        const T *wp = rec_b + p * RS;
                        Cell acc = cell_zero((Cell *)nullptr);
                        if constexpr (D == 3) {
                            constexpr int ZV = (W % 4 == 0) ? 4 : 2;
                            T wy[NI], wz[W];
                            VecLoad<T, NI>::load(wp + REC::OFF_Y + lg * NI, wy);
#pragma unroll
                            for (int j = 0; j < W; j += ZV) VecLoad<T, ZV>::load(wp + REC::OFF_Z + j, wz + j);
                            const T wx = wp[lx];
                            const Cell *c0 = tile + st.z * S2 + (st.y + lg) * Sx + st.x + lx;
                            const int dSx = G * Sx;
                            if (lane_on) {
                                constexpr int ZC = (NI * W <= 16) ? W : ((NI * 4 <= 16) ? 4 : (NI <= 8 ? 2 : 1));  // planes per chunk
#pragma unroll
                                for (int zb = 0; zb < W; zb += ZC) {
                                    Cell cell[ZC][NI];
#pragma unroll
                                    for (int jz = 0; jz < ZC; ++jz)
#pragma unroll
                                        for (int i = 0; i < NI; ++i)
                                            if (zb + jz < W && lg + i * G < W) cell[jz][i] = c0[(zb + jz) * S2 + i * dSx];
#pragma unroll
                                    for (int jz = 0; jz < ZC; ++jz) {
                                        if (zb + jz < W) {
                                            Cell pl = cell_zero((Cell *)nullptr);
#pragma unroll
                                            for (int i = 0; i < NI; ++i)
                                                if (lg + i * G < W) cfma(pl, cell[jz][i], wy[i]);
                                            cfma(acc, pl, wz[zb + jz]);
                                        }
                                    }
                                }
                                acc = cmul(acc, wx);
                            }
                        } else if constexpr (D == 2) {
                            const Cell *c0 = tile + (st.y + lg) * Sx + st.x + lx;
                            if (lane_on) {
#pragma unroll
                                for (int i = 0; i < NI; ++i) {
                                    const int jy = lg + i * G;
                                    if (jy < W) cfma(acc, c0[i * G * Sx], wp[REC::OFF_Y + jy]);
                                }
                                acc = cmul(acc, wp[lx]);
                            }
                        } else {
                            if (lane < W) acc = cmul(tile[st.x + lane], wp[lane]);
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) acc = cadd(acc, shfl_xor(acc, o));
                        if (lane == 0) {
                            const T scale = reinterpret_cast<const T *>(v_b + p)[0];
                            store_value<T, CPLX>(vp.p[c], st.w, cmul(acc, scale));
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
}

template <typename T, bool CPLX> int interp_dispatch(Plan &p, void *const vp[], const nufft_callbacks *cb);

}  // namespace nufft
