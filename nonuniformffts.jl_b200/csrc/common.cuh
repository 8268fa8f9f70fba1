// common.cuh — shared declarations of the B200 NUFFT backend (internal; the public ABI is include/nufft_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/nufft_b200.h"

namespace nufft {

constexpr int I0_MAX_TERMS = 96;   // per-dimension table of the I0 power series behind the polynomial coefficients (Direct KB)
constexpr int MAX_M = 12;          // largest half support instantiated
constexpr int MIN_M = 2;
constexpr int MAX_W = 2 * MAX_M;
constexpr int MAX_NPOLY = MAX_M + 4;

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern thread_local int64_t g_launch_count;
#define NUFFT_COUNT_LAUNCH() (++::nufft::g_launch_count)

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::nufft::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                               cudaGetErrorString(_e));                                       \
            return NUFFT_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

#define CUFFT_TRY(expr)                                                                       \
    do {                                                                                      \
        cufftResult _r = (expr);                                                              \
        if (_r != CUFFT_SUCCESS) {                                                            \
            ::nufft::set_error("cuFFT error %d at %s:%d", (int)_r, __FILE__, __LINE__);       \
            return NUFFT_ERR_CUFFT;                                                           \
        }                                                                                     \
    } while (0)

#define NUFFT_TRY(expr)            \
    do {                           \
        int _s = (expr);           \
        if (_s != NUFFT_SUCCESS) return _s; \
    } while (0)

// ---- cell (grid element) types --------------------------------------------------------------
template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T, bool CPLX> struct CellOf { using type = T; };
template <typename T> struct CellOf<T, true> { using type = typename Vec2<T>::type; };

template <typename T> __host__ __device__ inline T cell_zero(T *) { return T(0); }
__host__ __device__ inline float2 cell_zero(float2 *) { return make_float2(0.f, 0.f); }
__host__ __device__ inline double2 cell_zero(double2 *) { return make_double2(0., 0.); }

// scale: cell * real
__device__ __forceinline__ float cmul(float v, float s) { return v * s; }
__device__ __forceinline__ double cmul(double v, double s) { return v * s; }
__device__ __forceinline__ float2 cmul(float2 v, float s) { return make_float2(v.x * s, v.y * s); }
__device__ __forceinline__ double2 cmul(double2 v, double s) { return make_double2(v.x * s, v.y * s); }
// fused acc += v * s
__device__ __forceinline__ void cfma(float &a, float v, float s) { a = fmaf(v, s, a); }
__device__ __forceinline__ void cfma(double &a, double v, double s) { a = fma(v, s, a); }
__device__ __forceinline__ void cfma(float2 &a, float2 v, float s) { a.x = fmaf(v.x, s, a.x); a.y = fmaf(v.y, s, a.y); }
__device__ __forceinline__ void cfma(double2 &a, double2 v, double s) { a.x = fma(v.x, s, a.x); a.y = fma(v.y, s, a.y); }
__device__ __forceinline__ bool cnonzero(float v) { return v != 0.f; }
__device__ __forceinline__ bool cnonzero(double v) { return v != 0.; }
__device__ __forceinline__ bool cnonzero(float2 v) { return v.x != 0.f || v.y != 0.f; }
__device__ __forceinline__ bool cnonzero(double2 v) { return v.x != 0. || v.y != 0.; }

// global atomic accumulate of one cell (complex f32 -> one REDG.F32x2 on sm_90+)
__device__ __forceinline__ void catomic_add(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void catomic_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void catomic_add(float2 *p, float2 v) { atomicAdd(p, v); }
__device__ __forceinline__ void catomic_add(double2 *p, double2 v)
{
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}

__device__ __forceinline__ float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ float2 shfl_xor(float2 v, int m)
{
    return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ double2 shfl_xor(double2 v, int m)
{
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ float cadd(float a, float b) { return a + b; }
__device__ __forceinline__ double cadd(double a, double b) { return a + b; }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// ---- kernel (window function) parameters handed to device code by value ----------------------
template <typename T> struct KernelParams {
    int kind;        // NUFFT_KERNEL_*
    int mode;        // NUFFT_EVAL_*
    int M;
    int N[3];        // oversampled grid size per dim
    T beta[3];       // KB/BKB shape
    T tau[3];        // Gaussian 2 sigma^2
    T dx[3];         // 2pi / N
    const T *cs;     // device: per dim d a block of CS_STRIDE values:
                     //   [ (M+4)*2M polynomial coefs, layout [p][j] | M Gaussian exponentials | I0_MAX_TERMS series coefs ]
    int cs_stride;   // = (M+4)*2M + M + I0_MAX_TERMS
    int i0_terms;    // Direct KB: terms of the I0 power series to use (kernel_eval.cuh), 0 = library cyl_bessel_i0
};

// ---- geometry of bins / tiles handed to spreading & interpolation kernels ---------------------
struct TileGeom {
    int D;
    int N[3];        // oversampled dims (1 for unused dims)
    int B[3];        // bin dims
    int nb[3];       // bins per dim
    int T[3];        // tile dims = B + 2M - 1 (1 for unused dims)
    int S[3];        // tile strides in cells: S[0]=1 implied; S[1]=row stride, S[2]=plane stride ; S[0] holds padded row length
    int tile_cells;  // allocated cells per tile
    int chunk;       // max points per work item
    int batch;       // points per evaluation batch (shared-memory staging)
    // column-streaming fast path (cs_spread.cuh / cs_interp.cuh; D = 3, M = 4, ComplexF32): bins of 4 x 4 x Bz cells are
    // refined into layers of 4 cells along z; the sort key is bin * nsub + layer
    int rt;          // 3 when the plan is on that path (0 otherwise)
    int sub[3];      // sub-bins per bin along each dimension (1 when !rt)
    int nsub;        // sub[0] * sub[1] * sub[2]
};

// ---- the plan ------------------------------------------------------------------------------------
struct Plan {
    nufft_opts opts{};
    int D = 0, M = 0, C = 1;
    bool cplx = false, f64 = false;
    int method = NUFFT_METHOD_GLOBAL_MEMORY;   // resolved method
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t Ns[3] = {1, 1, 1};       // modes
    int64_t Nos[3] = {1, 1, 1};      // oversampled physical dims
    int64_t Nspec[3] = {1, 1, 1};    // oversampled spectral dims (r2c: first halved)
    int64_t nk[3] = {1, 1, 1};       // size(plan)
    int64_t ncells = 1, nspec = 1, nkept = 1;
    size_t real_bytes = 4;

    // host copies of kernel data (double) for introspection
    double h_shape[3] = {0, 0, 0}, h_dx[3] = {0, 0, 0};
    std::string h_cs[3];             // raw bytes of T cs block per dim
    std::string h_phihat[3];         // raw bytes of T phihat per dim

    // device tables
    void *d_cs = nullptr;            // T[3*cs_stride]
    int cs_stride = 0;
    void *d_phihat[3] = {nullptr, nullptr, nullptr};   // T[nk[d]]
    int32_t *d_imap[3] = {nullptr, nullptr, nullptr};  // kept index -> oversampled spectral index
    int32_t *d_invmap[3] = {nullptr, nullptr, nullptr};// oversampled spectral index -> kept index or -1
    double kp_beta[3] = {0, 0, 0}, kp_tau[3] = {0, 0, 0};

    // grids
    void *d_us = nullptr;            // C * ncells * sizeof(Z)
    void *d_uhat = nullptr;          // real plans: C * nspec * sizeof(complex T); complex plans: alias of d_us
    cufftHandle fft_fw = 0, fft_bw = 0;
    bool fft_ok = false;
    void *d_fft_work = nullptr;
    size_t fft_work_bytes = 0;
    // pruned FFT fused with deconvolution (pfft.cu): twiddle tables per dimension and the first intermediate array
    bool pfft = false;
    void *d_pf_tw[3] = {nullptr, nullptr, nullptr};
    void *d_pf_iph[3] = {nullptr, nullptr, nullptr};   // 1 / phihat per dimension
    void *d_pf_a = nullptr;

    // binning state
    TileGeom geom{};
    int64_t nbins = 1;
    int key_bits = 1;
    // plans eligible for the column-streaming kernels keep BOTH geometries and pick per set_points by point density
    // (cs below one point per cs_min_cells fine cells loses to the shared-memory tiles: windows see too few points)
    bool dual_geom = false;
    TileGeom geom_alt[2]{};          // [0] column-streaming, [1] shared-memory tiles
    int64_t nbins_alt[2] = {1, 1};
    int key_bits_alt[2] = {1, 1};
    double cs_min_cells = 16.0;
    int64_t Np = -1;
    int64_t cap = 0;                 // capacity (points) of the buffers below
    const void *user_x[3] = {nullptr, nullptr, nullptr};
    int x_stride = 1;                // element stride of the user's point arrays (D for a (D, Np) matrix)
    uint32_t *d_keys[2] = {nullptr, nullptr};
    int32_t *d_vals[2] = {nullptr, nullptr};
    int32_t *d_perm = nullptr;       // alias into d_vals
    void *d_xs[3] = {nullptr, nullptr, nullptr};       // sorted, folded coordinates (T)
    void *d_rec = nullptr;                             // folded coordinates in input order, one 4 x T record per point (D > 1)
    int32_t *d_bin_offsets = nullptr;                  // nbins + 1
    bool offsets_valid = true;                         // false: build on demand (binning_ensure_offsets)
    int32_t *d_perm_coarse = nullptr;                  // refined plans: bin-stable permutation, built on demand (introspection)
    int64_t perm_coarse_cap = 0;
    const int32_t *perm_coarse_ptr = nullptr;          // valid result of the last on-demand build (reset by set_points)
    int sort_cur = 0;                                  // which of d_vals[] holds the permutation
    uint32_t *d_hist = nullptr;      // radix histograms
    size_t hist_cap = 0;
    uint32_t *d_scan_tmp = nullptr;
    size_t scan_tmp_cap = 0;
    int32_t *d_item_start = nullptr; // nbins + 1: inclusive-scan form of work items per bin
    int2 *d_item_table = nullptr;    // per work item: (bin, chunk index)
    size_t item_cap = 0;
    int32_t *d_counters = nullptr;   // small block of device counters
    uint32_t *d_os = nullptr;        // single-sweep sort scratch: digit histograms, tile tickets, look-back status (binning.cu)
    size_t os_cap = 0;               // bytes
    int num_sms = 148;

    // z-slab plans (multi-GPU, mgpu.cu): the plan's grid holds the planes slab_z0 - (M - 1) .. slab_z0 + slab_nz + M - 1 of the
    // global oversampled grid (owned planes + halo), no periodic wrap along z; points must lie in the owned planes.
    int slab_z0 = 0;                 // first owned plane (global index)
    int slab_nz = 0;                 // owned planes; 0 = not a slab plan (full periodic grid)
    int nz_local = 0;                // planes stored (= Nos[2] for a full grid, slab_nz + 2M - 1 for a slab)

    // timings
    cudaEvent_t ev[32] = {};
    bool ev_ok = false;
    bool ev_rec[32] = {};
    float ms[16] = {};

    void *jit = nullptr;             // run-time compiled callbacks (callbacks_jit.cu)

    bool sticky_error = false;
};

// size (in T) of the per-point weight record of the shared-memory kernels (must match WRecord in tile_common.cuh)
static inline int record_size(int D, int M)
{
    const int W = 2 * M, G = 32 / W, NI = (W + G - 1) / G;
    const int wslot = (W + 3) / 4 * 4, yslot = (G * NI + 3) / 4 * 4;
    return D == 1 ? wslot : (D == 2 ? 2 * wslot : 2 * wslot + yslot);
}


// kernel-launch helper: ceil-div
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- stage implementations (one .cu each) ---------------------------------------------------
int host_plan_init(Plan &p);                 // host_plan.cu: sizes, kernel data, tables, geometry, FFT plans
void host_plan_free(Plan &p);
int binning_set_points(Plan &p, int64_t np, const void *const x[], int xstride = 1);        // binning.cu
int spread_run(Plan &p, const void *const vp[], const nufft_callbacks *cb); // spread_*.cu
int interp_run(Plan &p, void *const vp[], const nufft_callbacks *cb);       // interp_*.cu
int deconv_type1_run(Plan &p, void *const uhat[], const nufft_callbacks *cb); // deconv.cu
int deconv_type2_run(Plan &p, const void *const uhat[], const nufft_callbacks *cb);
bool pfft_eligible(const Plan &p);            // pfft.cu
int pfft_init(Plan &p);
void pfft_free(Plan &p);
int pfft_type1_run(Plan &p, void *const uhat[], const nufft_callbacks *cb);        // FFT + truncation + deconvolution
int pfft_type2_run(Plan &p, const void *const uhat[], const nufft_callbacks *cb);  // deconvolution + padding + FFT
int fft_forward(Plan &p);
int fft_backward(Plan &p);
struct JitCallbacks;                                                       // callbacks_jit.cu
int jit_callbacks_get(Plan &p, const nufft_callbacks *cb, JitCallbacks **out);
bool jit_has_nonuniform(const JitCallbacks *j);
bool jit_has_uniform(const JitCallbacks *j);
int jit_apply_nonuniform(Plan &p, JitCallbacks *j, const nufft_callbacks *cb, const void *const in[], void *out[], bool tmp);
int jit_apply_uniform(Plan &p, JitCallbacks *j, const nufft_callbacks *cb, const void *const in[], void *out[], bool tmp);
void jit_callbacks_free(Plan &p);
int scan_u32(Plan &p, uint32_t *data, int64_t n, bool inclusive);          // binning.cu (in place prefix sum)
int binning_ensure_offsets(Plan &p);                                       // binning.cu (column-streaming plans: on demand)
int binning_coarse_perm(Plan &p, const int32_t **perm);                    // binning.cu (rt plans: reference-order permutation)

template <typename T> KernelParams<T> make_kernel_params(const Plan &p)
{
    KernelParams<T> kp;
    kp.kind = p.opts.kernel;
    kp.mode = p.opts.eval_mode;
    kp.M = p.M;
    for (int d = 0; d < 3; ++d) {
        kp.N[d] = (int)p.Nos[d];
        kp.beta[d] = (T)p.kp_beta[d];
        kp.tau[d] = (T)p.kp_tau[d];
        kp.dx[d] = (T)2 * (T)3.14159265358979323846 / (T)p.Nos[d];
    }
    kp.cs = (const T *)p.d_cs;
    kp.cs_stride = p.cs_stride;
    kp.i0_terms = 0;
    if (p.opts.kernel == NUFFT_KERNEL_KAISER_BESSEL) {
        double bmax = 0;
        for (int d = 0; d < 3; ++d) bmax = p.kp_beta[d] > bmax ? p.kp_beta[d] : bmax;
        const double q = 0.25 * bmax * bmax, eps = sizeof(T) == 8 ? 1e-19 : 1e-10;
        double term = 1, sum = 1;
        int k = 0;
        while (k < 400) {
            ++k;
            term *= q / ((double)k * (double)k);
            sum += term;
            if ((double)k * k > q && term < eps * sum) break;
        }
        kp.i0_terms = (k + 1 <= I0_MAX_TERMS) ? k + 1 : 0;
    }
    return kp;
}

}  // namespace nufft
