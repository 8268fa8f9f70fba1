// interp.cuh — K-interp: type-2 interpolation kernels (templates; instantiated in interp_inst.cu).
//
// Replaces src/interpolation/gpu.jl:4-38,124-193 (global-memory kernel) and :211-395 (shared-memory
// kernel).  Values are scaled by prod_d dx_d (src/interpolation/gpu.jl:55-56).
//
// Shared-memory kernel: persistent CTAs pull (bin, chunk) work items; the bin's padded subgrid tile is
// staged in dynamic shared memory (periodic wrap applied while loading); threads evaluate the kernel
// values of a batch of points into shared memory; then one WARP per point computes the (2M)^D dot
// product with lanes laid out as (2M consecutive x cells) x (32/2M rows) — conflict-free 64-byte row
// segments — and reduces with warp shuffles; lane 0 scatters the result through the permutation.
#pragma once
#include "spread.cuh"

namespace nufft {

constexpr int INTERP_THREADS = 256;

template <typename T, bool CPLX> __device__ __forceinline__ void store_value(void *vp, int64_t i, typename CellOf<T, CPLX>::type v)
{
    using Cell = typename CellOf<T, CPLX>::type;
    ((Cell *)vp)[i] = v;
}

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

template <typename T, bool CPLX, int D, int M>
__global__ void __launch_bounds__(INTERP_THREADS)
interp_sm_kernel(KernelParams<T> kp, TileGeom g, SmArgs a, const T *__restrict__ xs0, const T *__restrict__ xs1,
                 const T *__restrict__ xs2, MutPtrPack vp, int C, const typename CellOf<T, CPLX>::type *__restrict__ us,
                 int64_t ncells, T prefactor, const T *__restrict__ nu_weights)
{
    using Cell = typename CellOf<T, CPLX>::type;
    constexpr int W = 2 * M;
    constexpr int NT = INTERP_THREADS;
    constexpr int NWARP = NT / 32;
    constexpr int WS = SmLayout<D, M>::WS;
    constexpr int G = 32 / W;
    constexpr int NI = (W + G - 1) / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cell *tile = (Cell *)smem_raw;
    Cell *v_s = tile + g.tile_cells;          // unused here; keeps the layout of the spreading kernel
    T *cs_s = (T *)(v_s + g.batch);
    T *w_s = cs_s + D * kp.cs_stride;
    int *st_s = (int *)(w_s + g.batch * WS);
    __shared__ int s_item[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lx = lane % W, lg = lane / W;
    const bool lane_on = lg < G;
    const int Tx = g.T[0], Ty = g.T[1], Tz = g.T[2], Sx = g.S[0];
    const int total_items = a.item_start[a.nbins];

    for (int i = tid; i < D * kp.cs_stride; i += NT) cs_s[i] = kp.cs[i];

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const int item = atomicAdd(a.work_counter, 1);
            s_item[0] = item;
            if (item < total_items) decode_item(a, item, g.chunk, s_item[1], s_item[2], s_item[3]);
        }
        __syncthreads();
        if (s_item[0] >= total_items) break;
        const int bin = s_item[1], k0 = s_item[2], k1 = s_item[3];
        int b = bin;
        const int bx = b % g.nb[0]; b /= g.nb[0];
        const int by = b % g.nb[1]; b /= g.nb[1];
        const int bz = b;
        const int org[3] = {bx * g.B[0], by * g.B[1], bz * g.B[2]};

        for (int c = 0; c < C; ++c) {
            // ---- stage the padded tile (periodic wrap) ----------------------------------------
            {
                const Cell *u = us + (int64_t)c * ncells;
                const int x0 = org[0] - (M - 1), y0 = org[1] - (M - 1), z0 = org[2] - (M - 1);
                const int rows = Ty * Tz;
                for (int row = warp; row < rows; row += NWARP) {
                    const int z = row / Ty, y = row - z * Ty;
                    const int gy = D > 1 ? wrap_index(y0 + y, g.N[1]) : 0;
                    const int gz = D > 2 ? wrap_index(z0 + z, g.N[2]) : 0;
                    const Cell *grow = u + ((int64_t)gz * g.N[1] + gy) * g.N[0];
                    Cell *trow = tile + (size_t)z * g.S[2] + (size_t)y * Sx;
                    for (int x = lane; x < Tx; x += 32) trow[x] = grow[wrap_index(x0 + x, g.N[0])];
                }
            }
            for (int kb = k0; kb < k1; kb += g.batch) {
                const int nb = min(g.batch, k1 - kb);
                __syncthreads();
                for (int t = tid; t < g.batch * D; t += NT) {
                    const int d = t / g.batch, p = t - d * g.batch;
                    if (p >= nb) continue;
                    const T *xs = d == 0 ? xs0 : (d == 1 ? xs1 : xs2);
                    T w[W];
                    const int i0 = eval_kernel_values<T, M>(kp, cs_s + d * kp.cs_stride, d, xs[kb + p], w);
                    st_s[p * 4 + d] = i0 - org[d];
                    T *dst = w_s + p * WS + d * W;
#pragma unroll
                    for (int j = 0; j < W; ++j) dst[j] = w[j];
                }
                __syncthreads();
                for (int p = warp; p < nb; p += NWARP) {
                    const int *st = st_s + p * 4;
                    const T *wp = w_s + p * WS;
                    Cell acc = cell_zero((Cell *)nullptr);
                    if constexpr (D == 3) {
                        const int sx = st[0], sy = st[1], sz = st[2];
                        const Cell *base = tile + (size_t)sz * g.S[2] + (size_t)sy * Sx + sx + lx;
                        T wyr[NI];
#pragma unroll
                        for (int i = 0; i < NI; ++i) wyr[i] = (lane_on && lg + i * G < W) ? wp[W + lg + i * G] : (T)0;
#pragma unroll
                        for (int jz = 0; jz < W; ++jz) {
                            const T wz = wp[2 * W + jz];
#pragma unroll
                            for (int i = 0; i < NI; ++i) {
                                const int jy = lg + i * G;
                                if (lane_on && jy < W) cfma(acc, base[(size_t)jz * g.S[2] + jy * Sx], wyr[i] * wz);
                            }
                        }
                        acc = cmul(acc, lane_on ? wp[lx] : (T)0);
                    } else if constexpr (D == 2) {
                        const int sx = st[0], sy = st[1];
                        const Cell *base = tile + (size_t)sy * Sx + sx + lx;
#pragma unroll
                        for (int i = 0; i < NI; ++i) {
                            const int jy = lg + i * G;
                            if (lane_on && jy < W) cfma(acc, base[jy * Sx], wp[W + jy]);
                        }
                        acc = cmul(acc, lane_on ? wp[lx] : (T)0);
                    } else {
                        const int sx = st[0];
                        if (lane < W) acc = cmul(tile[sx + lane], wp[lane]);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc = cadd(acc, shfl_xor(acc, o));
                    if (lane == 0) {
                        const int32_t n = a.perm[kb + p];
                        const T scale = prefactor * (nu_weights ? nu_weights[n] : (T)1);
                        store_value<T, CPLX>(vp.p[c], n, cmul(acc, scale));
                    }
                }
            }
            __syncthreads();
        }
    }
}

template <typename T, bool CPLX> int interp_dispatch(Plan &p, void *const vp[], const nufft_callbacks *cb);

}  // namespace nufft
