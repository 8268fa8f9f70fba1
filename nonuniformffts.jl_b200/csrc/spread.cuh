// spread.cuh — K-spread: type-1 spreading kernels (templates; instantiated in spread_inst_*.cu).
//
// Replaces src/spreading/gpu.jl:2-127 (global-memory kernel) and :237-434 (shared-memory kernel),
// plus the preceding fill_with_zeros_kernel! (src/NonuniformFFTs.jl:116-122,161-167).
//
// Shared-memory kernel design (sm_100a):
//   * persistent CTAs pull work items (bin, chunk of <= `chunk` sorted points) from a device counter;
//   * the bin's padded subgrid tile (B + 2M - 1 per dim) lives in dynamic shared memory (up to 227 KiB);
//   * phase 1: threads evaluate the 1-D kernel values of a batch of points into shared memory;
//   * phase 2: the tile is partitioned among the M warps of the CTA by residue class of the slowest
//     tile coordinate modulo M.  A point's support spans 2M consecutive planes, so it meets every
//     residue class exactly twice: every warp performs the same amount of work for every point
//     (perfect balance), each tile cell has exactly one owning warp (plain read-modify-write, no
//     shared-memory atomics, which are CAS loops for floating point on this architecture, and no
//     CTA barrier per point as in the reference);
//   * flush: tile -> oversampled grid with vector red.global.add (REDG.F32x2 for complex f32),
//     periodic wrap applied per row/column.
#pragma once
#include "common.cuh"
#include "kernel_eval.cuh"

namespace nufft {

constexpr int MAX_PACK = 8;
struct PtrPack {
    const void *p[MAX_PACK];
};
struct MutPtrPack {
    void *p[MAX_PACK];
};

template <typename T, bool CPLX> __device__ __forceinline__ typename CellOf<T, CPLX>::type load_value(const void *vp, int64_t i)
{
    using Cell = typename CellOf<T, CPLX>::type;
    return ((const Cell *)vp)[i];
}

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
struct SmArgs {
    const int32_t *perm;
    const int32_t *bin_offsets;   // nbins + 1
    const int32_t *item_start;    // nbins + 1 (inclusive-scan form)
    int32_t *work_counter;
    int nbins;
};

__device__ __forceinline__ int wrap_index(int g, int N)
{
    while (g < 0) g += N;
    while (g >= N) g -= N;
    return g;
}

// Decode a work item into (bin, [k0, k1)).  Executed by one thread.
__device__ __forceinline__ void decode_item(const SmArgs &a, int item, int chunk, int &bin, int &k0, int &k1)
{
    int lo = 0, hi = a.nbins;          // find largest b with item_start[b] <= item
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.item_start[mid] <= item) lo = mid; else hi = mid;
    }
    bin = lo;
    const int off = a.bin_offsets[bin], end = a.bin_offsets[bin + 1];
    k0 = off + (item - a.item_start[bin]) * chunk;
    k1 = min(k0 + chunk, end);
}

template <int D, int M> struct SmLayout {
    static constexpr int W = 2 * M;
    static constexpr int WS = (D * W) | 1;     // per-point stride of the weight rows (odd: conflict-free writes)
};

template <typename T, bool CPLX, int D, int M>
__host__ __device__ inline size_t sm_dynamic_bytes(const TileGeom &g, int cs_stride)
{
    using Cell = typename CellOf<T, CPLX>::type;
    size_t b = (size_t)g.tile_cells * sizeof(Cell);
    b += (size_t)g.batch * sizeof(Cell);                          // values
    b += (size_t)D * cs_stride * sizeof(T);                       // kernel coefficient tables
    b += (size_t)g.batch * SmLayout<D, M>::WS * sizeof(T);        // weights
    b += (size_t)g.batch * 4 * sizeof(int);                       // local start indices
    return b + 16;
}

template <typename T, bool CPLX, int D, int M>
__global__ void __launch_bounds__(32 * M)
spread_sm_kernel(KernelParams<T> kp, TileGeom g, SmArgs a, const T *__restrict__ xs0, const T *__restrict__ xs1,
                 const T *__restrict__ xs2, PtrPack vp, int C, typename CellOf<T, CPLX>::type *__restrict__ us,
                 int64_t ncells, const T *__restrict__ nu_weights)
{
    using Cell = typename CellOf<T, CPLX>::type;
    constexpr int W = 2 * M;
    constexpr int NT = 32 * M;
    constexpr int WS = SmLayout<D, M>::WS;
    constexpr int G = 32 / W;                 // rows handled concurrently by a warp (W <= 24 -> G >= 1)
    constexpr int NI = (W + G - 1) / G;       // row iterations per plane
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cell *tile = (Cell *)smem_raw;
    Cell *v_s = tile + g.tile_cells;
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
        const int org[3] = {bx * g.B[0], by * g.B[1], bz * g.B[2]};   // first cell of the bin

        for (int c = 0; c < C; ++c) {
            const Cell zero = cell_zero((Cell *)nullptr);
            for (int i = tid; i < g.tile_cells; i += NT) tile[i] = zero;
            for (int kb = k0; kb < k1; kb += g.batch) {
                const int nb = min(g.batch, k1 - kb);
                __syncthreads();   // previous batch fully consumed (and tile zeroed)
                // ---- phase 1: kernel values of the batch ------------------------------------
                for (int t = tid; t < g.batch * (D + 1); t += NT) {
                    const int d = t / g.batch, p = t - d * g.batch;
                    if (p >= nb) continue;
                    if (d == D) {
                        const int32_t n = a.perm[kb + p];
                        Cell v = load_value<T, CPLX>(vp.p[c], n);
                        if (nu_weights) v = cmul(v, nu_weights[n]);
                        v_s[p] = v;
                    } else {
                        const T *xs = d == 0 ? xs0 : (d == 1 ? xs1 : xs2);
                        T w[W];
                        const int i0 = eval_kernel_values<T, M>(kp, cs_s + d * kp.cs_stride, d, xs[kb + p], w);
                        st_s[p * 4 + d] = i0 - org[d];         // local index of the first support cell
                        T *dst = w_s + p * WS + d * W;
#pragma unroll
                        for (int j = 0; j < W; ++j) dst[j] = w[j];
                    }
                }
                __syncthreads();
                // ---- phase 2: accumulate into the tile, ownership by residue class mod M ------
                for (int p = 0; p < nb; ++p) {
                    const int *st = st_s + p * 4;
                    const T *wp = w_s + p * WS;
                    const Cell v = v_s[p];
                    if constexpr (D == 3) {
                        const int sx = st[0], sy = st[1], sz = st[2];
                        int r = (warp - sz) % M; if (r < 0) r += M;      // first owned plane offset in [0, M)
                        const Cell vx = cmul(v, lane_on ? wp[lx] : (T)0);
#pragma unroll
                        for (int t2 = 0; t2 < 2; ++t2) {
                            const int jz = r + t2 * M;
                            const T wz = wp[2 * W + jz];
                            Cell *plane = tile + (size_t)(sz + jz) * g.S[2] + sx + lx;
#pragma unroll
                            for (int i = 0; i < NI; ++i) {
                                const int jy = lg + i * G;
                                if (lane_on && jy < W) {
                                    Cell *cell = plane + (sy + jy) * Sx;
                                    Cell acc = *cell;
                                    cfma(acc, vx, wp[W + jy] * wz);
                                    *cell = acc;
                                }
                            }
                        }
                    } else if constexpr (D == 2) {
                        const int sx = st[0], sy = st[1];
                        int r = (warp - sy) % M; if (r < 0) r += M;
                        const Cell vx = cmul(v, lane_on ? wp[lx] : (T)0);
                        for (int t2 = lg; t2 < 2; t2 += G) {
                            if (lane_on) {
                                const int jy = r + t2 * M;
                                Cell *cell = tile + (sy + jy) * Sx + sx + lx;
                                Cell acc = *cell;
                                cfma(acc, vx, wp[W + jy]);
                                *cell = acc;
                            }
                        }
                    } else {
                        const int sx = st[0];
                        if (lane < W) {
                            int r = (sx + lane) % M;
                            if (r == warp) {
                                Cell acc = tile[sx + lane];
                                cfma(acc, v, wp[lane]);
                                tile[sx + lane] = acc;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            // ---- flush: tile -> global grid (periodic), vector atomics ------------------------
            Cell *u = us + (int64_t)c * ncells;
            const int x0 = org[0] - (M - 1), y0 = org[1] - (M - 1), z0 = org[2] - (M - 1);
            const int rows = Ty * Tz;
            for (int row = warp; row < rows; row += M) {
                const int z = row / Ty, y = row - z * Ty;
                const int gy = D > 1 ? wrap_index(y0 + y, g.N[1]) : 0;
                const int gz = D > 2 ? wrap_index(z0 + z, g.N[2]) : 0;
                Cell *grow = u + ((int64_t)gz * g.N[1] + gy) * g.N[0];
                const Cell *trow = tile + (size_t)z * g.S[2] + (size_t)y * Sx;
                for (int x = lane; x < Tx; x += 32) {
                    const Cell val = trow[x];
                    if (cnonzero(val)) catomic_add(grow + wrap_index(x0 + x, g.N[0]), val);
                }
            }
            __syncthreads();
        }
    }
}

// host-side launchers (defined per (T, CPLX) translation unit)
template <typename T, bool CPLX> int spread_dispatch(Plan &p, const void *const vp[], const nufft_callbacks *cb);

}  // namespace nufft
