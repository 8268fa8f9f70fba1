/*
 * nufft_b200.h — C ABI of the B200-native NUFFT backend (libnufft_b200.so).
 *
 * This is the drop-in boundary for the hot path of NonuniformFFTs.jl v0.9.6
 *     PlanNUFFT(...)  ->  set_points!  ->  exec_type1! / exec_type2!
 * The reference has no FFI: its GPU path is Julia multiple dispatch on
 * `backend::KA.GPU`.  Each entry point below names the Julia method(s) whose GPU
 * specialisation it replaces (paths relative to the reference repository).
 * A Julia host reaches these through `ccall` (see INTEGRATION.md and
 * nonuniformffts.jl_b200/julia/NonuniformFFTsB200.jl); the test-suite and bench
 * reach them through Python ctypes.
 *
 * Conventions
 *   - All array arguments are DEVICE pointers unless the name ends in `_host`.
 *   - Arrays are dense, column-major (Julia layout): dimension 1 is contiguous.
 *   - Uniform data: C = ntransforms separate arrays of complex(T) with dims size(plan)
 *     (real-data plans: first dim N1/2+1, src/plan.jl:558-562).
 *   - Non-uniform data: C separate arrays of length Np of Z (T or complex T).
 *   - Points: D separate arrays of T, length Np (SoA), any real value (folded to [0,2pi)).
 *   - Work is enqueued on the plan's stream; entry points do not synchronise the device
 *     (the reference is stream-ordered too unless synchronise=true, src/plan.jl:453-454).
 *   - Every function returns NUFFT_SUCCESS (0) or a negative error code and never aborts;
 *     nufft_last_error() returns a thread-local description.
 *   - Indices exposed through the ABI are 0-based int32 (reference: 1-based Int64).
 */
#ifndef NUFFT_B200_H
#define NUFFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NUFFT_B200_ABI_VERSION 2

/* ---- error codes (Julia shim maps them back to the reference's exception types) ---- */
enum {
    NUFFT_SUCCESS = 0,
    NUFFT_ERR_ARG = -1,         /* ArgumentError   (src/plan.jl:545-556, src/set_points.jl:35, src/NonuniformFFTs.jl:154,243) */
    NUFFT_ERR_DIM = -2,         /* DimensionMismatch (src/NonuniformFFTs.jl:92-114, src/blocking/gpu.jl:86) */
    NUFFT_ERR_UNSUPPORTED = -3, /* parameter combination not instantiated in this build */
    NUFFT_ERR_CUDA = -4,        /* CUDA runtime error (sticky) */
    NUFFT_ERR_CUFFT = -5,       /* cuFFT error */
    NUFFT_ERR_ALLOC = -6,       /* device allocation failed */
    NUFFT_ERR_STATE = -7        /* exec before set_points, destroyed plan, ... */
};

/* ---- enumerations ---- */
enum { NUFFT_F32 = 0, NUFFT_F64 = 1 };
/* src/NonuniformFFTs.jl:23-35: the four reference kernels */
enum { NUFFT_KERNEL_KAISER_BESSEL = 0, NUFFT_KERNEL_BACKWARDS_KAISER_BESSEL = 1,
       NUFFT_KERNEL_GAUSSIAN = 2, NUFFT_KERNEL_BSPLINE = 3,
       /* NOT in the reference (src/Kernels/ has the four above): "exponential of semicircle" exp(beta (sqrt(1 - y^2) - 1)) of
        * Barnett, Magland & af Klinteberg (2019), beta = 0.976 pi M (2 - 1/sigma), NUFFT_EVAL_FAST only (the same piecewise
        * polynomials as KB / BKB), Fourier transform by Gauss-Legendre quadrature.  PARITY UNPINNED: there is no reference
        * output to compare with; the oracle restates the same construction and both are checked against exact NUDFT sums. */
       NUFFT_KERNEL_ES = 4 };
/* src/Kernels/Kernels.jl:14-46 */
enum { NUFFT_EVAL_FAST = 0, NUFFT_EVAL_DIRECT = 1 };
/* gpu_method keyword, src/plan.jl:479, src/blocking/gpu.jl:26 */
enum { NUFFT_METHOD_AUTO = 0, NUFFT_METHOD_GLOBAL_MEMORY = 1, NUFFT_METHOD_SHARED_MEMORY = 2 };

typedef struct nufft_plan_s *nufft_plan;   /* opaque; owns all device scratch */

/*
 * Options of a plan = keyword arguments of `_PlanNUFFT` (src/plan.jl:467-482) and
 * `PlanNUFFT` (src/plan.jl:568-599).  Zero-initialise, set struct_size = sizeof(nufft_opts),
 * then fill; nufft_opts_default() does that with the reference defaults.
 */
typedef struct {
    uint32_t struct_size;     /* = sizeof(nufft_opts) (versioning) */
    int32_t  dim;             /* D in 1..3 */
    int64_t  n_modes[3];      /* Ns: number of Fourier modes per dimension (src/plan.jl:467) */
    int32_t  is_complex;      /* Z <: Complex ? 1 : 0   (non-uniform data type) */
    int32_t  dtype;           /* NUFFT_F32 / NUFFT_F64 = real(Z) */
    int32_t  half_support;    /* M of HalfSupport(M), default 4 (src/plan.jl:583) */
    double   sigma;           /* oversampling factor, default 2 (src/plan.jl:573) */
    int32_t  kernel;          /* NUFFT_KERNEL_*; reference CUDA default: KAISER_BESSEL (ext/NonuniformFFTsCUDAExt.jl:19) */
    double   kernel_param;    /* beta (KB/BKB) or ell/dx (Gaussian); NaN = default shape rule */
    int32_t  eval_mode;       /* NUFFT_EVAL_*; reference CUDA default: DIRECT (ext/...CUDAExt.jl:23) */
    int32_t  ntransforms;     /* C >= 1 (src/plan.jl:570) */
    int32_t  fftshift;        /* 1: increasing wavenumber order (complex plans only, src/plan.jl:273-280) */
    int32_t  sort_points;     /* accepted for API parity; this backend always keeps a sorted, folded
                                 copy of the points, results are identical either way */
    int32_t  gpu_method;      /* NUFFT_METHOD_* */
    int64_t  block_dims[3];   /* bin (block) dims in oversampled-grid cells; 0 = automatic */
    int32_t  point_convention;/* 0: x in [0,2pi), e^{-ikx} type-1;  1: AbstractNFFTs convention, x in [-1/2,1/2),
                                 opposite sign (src/abstractNFFTs.jl:150-158) */
    int32_t  device;          /* CUDA device ordinal; -1 = current device */
    void    *stream;          /* cudaStream_t; NULL = default stream */
    int32_t  record_timings;  /* 1: bracket every stage with CUDA events (reference: TimerOutputs + synchronise) */
    int32_t  spread_chunk;    /* max points per work item (0 = default); bins holding more are split */
} nufft_opts;

/*
 * Callbacks (NUFFTCallbacks, src/plan.jl:146-164).  A compiled library cannot inline Julia
 * closures; this is the menu that covers every callback the reference's tests use
 * (test/callbacks.jl:17-25, test/pseudo_gpu.jl:176-223).  NULL pointer / NULL fields = default_callback.
 */
typedef struct {
    uint32_t struct_size;
    const void *nu_weights;          /* T[Np] device: v[c] *= w[n], n = ORIGINAL point index           */
    const void *const *u_factor_sep; /* D device tables of T, len size(p)[d]: w[c] *= prod_d f_d[i_d]  */
    const void *u_factor_dense;      /* T array of dims size(p): w[c] *= f[I]                           */
    /* general callbacks (optional; read only when struct_size covers them): CUDA C++ source compiled with NVRTC once per
     * plan, defining nufft_cb_nonuniform(nufft_cell (&v)[NUFFT_C], long long n, const void *user) after
     * `#define NUFFT_HAS_NONUNIFORM 1` and / or nufft_cb_uniform(nufft_cplx (&w)[NUFFT_C], const int (&idx)[3],
     * const void *user) after `#define NUFFT_HAS_UNIFORM 1` — the arbitrary closures of src/plan.jl:146-164; applied
     * where the reference applies them (csrc/callbacks_jit.cu), in addition to the menu above. */
    const char *nvrtc_src;
    const void *user_data;           /* device pointer handed to the callbacks                          */
} nufft_callbacks;

/* ---- plan lifetime: PlanNUFFT constructors, src/plan.jl:467-599 ---- */
int  nufft_opts_default(nufft_opts *opts);
int  nufft_plan_create(nufft_plan *plan, const nufft_opts *opts);
int  nufft_plan_destroy(nufft_plan plan);

/* Base.size(p) (src/plan.jl:426), oversampled dims (src/plan.jl:485-498), ntransforms(p) (:435) */
int  nufft_plan_shape(nufft_plan plan, int64_t size_out[3], int64_t os_dims[3], int32_t *ntransforms);

/* kernel data for parity tests: shape parameter (beta or tau), dx, the (M+4)x2M polynomial
 * coefficients [p][j] and the Fourier coefficients phihat_d (src/Kernels/<kernel>.jl). Host output. */
int  nufft_plan_kernel_info(nufft_plan plan, int32_t d, double *shape_param, double *dx,
                            double *cs_host /* (M+4)*2M or NULL */, double *phihat_host /* size(p)[d] or NULL */);

/* The same kernel data WITHOUT a plan and without touching a device: what a plan created from `opts` would hold for dimension d
 * (optimal_kernel + the *KernelData constructors, src/Kernels/<kernel>.jl; phihat over the kept wavenumbers, src/plan.jl:503-512).
 * cs_len >= (M+4)*2M, phihat_len >= size(p)[d]; any output pointer may be NULL.  Lets CPU-only tests compare the library's tables
 * with the oracle's. */
int  nufft_kernel_tables(const nufft_opts *opts, int32_t d, double *shape_param, double *dx, int64_t *os_dim,
                         double *cs_host, size_t cs_len, double *phihat_host, size_t phihat_len);

/* ---- set_points!(p, (xs, ys, zs)): src/set_points.jl:33-52 + src/blocking/gpu.jl:73-142 ---- */
int  nufft_set_points(nufft_plan plan, int64_t np, const void *const x[/*dim*/]);
/* set_points!(p, xp::AbstractMatrix) with size(xp) == (D, Np) and set_points!(p, xp::AbstractVector{<:SVector{D}})
 * (src/set_points.jl:62-88; the reference copies both into D separate vectors on the host side of Julia): xmat is a
 * dense column-major (D, Np) device matrix of the plan's real type, i.e. an array of Np D-vectors.  Read in place by
 * the binning kernel — no transposed copy.  Same retention rules as nufft_set_points. */
int  nufft_set_points_matrix(nufft_plan plan, int64_t np, const void *xmat);

/* binning result (BlockDataGPU.pointperm / cumulative_npoints_per_block, src/blocking/gpu.jl:2-21).
 * Device pointers owned by the plan, valid until the next set_points / destroy. */
int  nufft_get_binning(nufft_plan plan, const int32_t **perm, const int32_t **bin_offsets,
                       int64_t *nbins, int64_t bin_dims[3]);

/* The order the kernels actually use.  Plans on the column-streaming fast path (3-D, Float32 data, HalfSupport(4), at least
 * one point per 16 fine cells) use bins of 4 x 4 x Bz cells refined by the z cell inside the bin and sort by
 * key = bin * Bz + z cell (stable), a refinement of the reference order; nufft_get_binning then rebuilds the bin-stable
 * permutation (and the bin offsets) on demand.
 * Other plans: nsub = 1 and both getters agree.  fine_offsets has nfine + 1 entries. */
int  nufft_get_binning_fine(nufft_plan plan, const int32_t **perm, const int32_t **fine_offsets,
                            int64_t *nfine, int64_t sub_dims[3]);

/* ---- exec_type1!(us_k, p, vp; callbacks): src/NonuniformFFTs.jl:148-195 ---- */
int  nufft_exec_type1(nufft_plan plan, void *const uhat[/*C*/], const void *const vp[/*C*/],
                      const nufft_callbacks *cb);
/* ---- exec_type2!(vp, p, us_k; callbacks): src/NonuniformFFTs.jl:237-291 ---- */
int  nufft_exec_type2(nufft_plan plan, void *const vp[/*C*/], const void *const uhat[/*C*/],
                      const nufft_callbacks *cb);

/*
 * Stage-level entry points (used by the multi-GPU host layer to place a collective between
 * stages; same stage split as the reference's timer sections, src/NonuniformFFTs.jl:157-186,246-283).
 *   type 1:  spread  = (0) fill with zeros + (1) spreading          -> plan grid
 *            finish  = (2) forward FFT + (3) deconvolution           -> uhat
 *   type 2:  prepare = (0)+(1) zero-pad/deconvolve + (2) backward FFT -> plan grid
 *            interp  = (3) interpolation                              -> vp
 */
int  nufft_type1_spread(nufft_plan plan, const void *const vp[], const nufft_callbacks *cb);
int  nufft_type1_finish(nufft_plan plan, void *const uhat[], const nufft_callbacks *cb);
int  nufft_type2_prepare(nufft_plan plan, const void *const uhat[], const nufft_callbacks *cb);
int  nufft_type2_interp(nufft_plan plan, void *const vp[], const nufft_callbacks *cb);

/* plan-owned oversampled physical grid `us` (PlanNUFFT.data.us, src/plan.jl:3-31): C contiguous
 * arrays of Z with dims os_dims.  bytes_per_transform = prod(os_dims) * sizeof(Z). */
int  nufft_get_grid(nufft_plan plan, void **grid, size_t *bytes_per_transform);

/* per-stage device times in ms of the most recent calls (record_timings=1), reference stage names:
 *  [0] set_points  [1] t1 fill zeros  [2] t1 spreading  [3] t1 forward FFT  [4] t1 deconvolution
 *  [5] t2 zero-pad+deconvolution  [6] t2 backward FFT  [7] t2 interpolation  [8..15] reserved */
int  nufft_get_timings(nufft_plan plan, float ms[16]);

/* counts kernel launches issued by this library on the calling thread since the last reset */
int64_t nufft_launch_count(int reset);

/* Base.show(::PlanNUFFT) (src/plan.jl:362-392) + chosen tile/CTA shapes and bin statistics */
int  nufft_describe(nufft_plan plan, char *buf, size_t buflen);

const char *nufft_last_error(void);
int  nufft_abi_version(void);

/* =====================================================================================================================
 * Multi-GPU transforms (one B200 per rank, NCCL over NVLink / NVSwitch).  The reference has no distributed code
 * (SURVEY.md §2c); these entry points shard the same hot path — set_points! (src/set_points.jl:33-52),
 * exec_type1! / exec_type2! (src/NonuniformFFTs.jl:148-291) — the three ways it shards naturally:
 *   NUFFT_MGPU_SLAB        z-slab spatial decomposition of ONE large transform: points are routed to the rank that owns
 *                          their z planes, each rank spreads / interpolates on its slab (+ 2M - 1 halo planes exchanged with
 *                          the two neighbours), the pruned FFT runs its x and y passes on the slab and its z pass after an
 *                          all-to-all transpose.  3-D, ComplexF32, HalfSupport(4), ntransforms = 1, power-of-two oversampled
 *                          sizes; oversampled z size and kept y size divisible by the number of ranks.  Uniform data is
 *                          DISTRIBUTED: rank r holds uhat[:, r Ky/G : (r + 1) Ky/G, :] (nufft_mgpu_local_block).
 *   NUFFT_MGPU_POINTS      points partitioned, full grid per rank; type 1 sums the partial outputs (all-reduce of
 *                          prod(size(p)) values), type 2 broadcasts rank 0's coefficients (written into the uhat buffers of
 *                          the other ranks).  Every plan the single-GPU entry points support.
 *   NUFFT_MGPU_TRANSFORMS  the C = ntransforms independent transforms are dealt to ranks (c mod nranks == rank): every rank
 *                          passes the same points and the full tuples of arrays; only its own components are touched.
 *   NUFFT_MGPU_AUTO        SLAB when eligible, else TRANSFORMS when ntransforms >= nranks, else POINTS.
 * One handle drives `nlocal` of the `nranks` ranks: nlocal == nranks for a single host process that owns all GPUs (what a
 * Julia host does through ccall), nlocal == 1 for one process per GPU (the id of nufft_mgpu_unique_id travels through the
 * host's own channel, e.g. MPI or torch.distributed).  Array arguments are per LOCAL rank l, flattened:
 * x[3 l + d], vp[C l + c], uhat[C l + c].  All calls are collective over the ranks and stream-ordered on each rank's
 * stream (nufft_mgpu_get_stream; opts.stream when nlocal == 1).  NCCL is loaded at run time (dlopen of libnccl.so.2).
 * ===================================================================================================================== */
#define NUFFT_MGPU_ID_BYTES 128
enum { NUFFT_MGPU_AUTO = 0, NUFFT_MGPU_SLAB = 1, NUFFT_MGPU_POINTS = 2, NUFFT_MGPU_TRANSFORMS = 3 };
typedef struct nufft_mgpu_s *nufft_mgpu;

/* ncclGetUniqueId: call on one rank, hand the 128 bytes to every nufft_mgpu_create of the job */
int  nufft_mgpu_unique_id(void *id128);
/* PlanNUFFT(...) on every rank (src/plan.jl:467-599); opts.device is ignored (devices[l] is used) */
int  nufft_mgpu_create(nufft_mgpu *h, const nufft_opts *opts, int32_t nranks, int32_t nlocal, const int32_t *local_ranks,
                       const int32_t *devices, const void *id128, int32_t strategy);
int  nufft_mgpu_destroy(nufft_mgpu h);
/* resolved strategy, rank counts, size(p) and oversampled dims of the GLOBAL problem */
int  nufft_mgpu_info(nufft_mgpu h, int32_t *strategy, int32_t *nranks, int32_t *nlocal, int64_t size_out[3], int64_t os_dims[3]);
/* block of the uniform array that local rank l holds: uhat[offset : offset + size) per dimension */
int  nufft_mgpu_local_block(nufft_mgpu h, int32_t l, int64_t offset[3], int64_t size[3]);
/* set_points!: local rank l passes ITS np[l] points (any partition of the global set; TRANSFORMS: all points on every rank) */
int  nufft_mgpu_set_points(nufft_mgpu h, const int64_t np[/*nlocal*/], const void *const x[/*3 nlocal*/]);
/* exec_type1! / exec_type2!: values belong to the points the same local rank passed to set_points, in the same order */
int  nufft_mgpu_exec_type1(nufft_mgpu h, void *const uhat[/*C nlocal*/], const void *const vp[/*C nlocal*/], const nufft_callbacks *cb);
int  nufft_mgpu_exec_type2(nufft_mgpu h, void *const vp[/*C nlocal*/], const void *const uhat[/*C nlocal*/], const nufft_callbacks *cb);
/* SLAB: all-gather of the distributed type-1 output into a full size(p) array on every local rank (checks, small problems) */
int  nufft_mgpu_gather_output(nufft_mgpu h, void *const full[/*nlocal*/], const void *const local[/*nlocal*/]);
int  nufft_mgpu_synchronize(nufft_mgpu h);
int  nufft_mgpu_get_stream(nufft_mgpu h, int32_t l, void **stream);
/* per-stage device times (ms) of local rank l, record_timings = 1:  [0] point exchange  [1] local set_points
 * [2] type-1 value exchange  [3] zero fill + spreading  [4] halo exchange + add  [5] FFT passes x, y  [6] transpose
 * [7] FFT pass z  [8] type-2 FFT pass z  [9] transpose  [10] FFT passes y, x  [11] halo exchange  [12] interpolation
 * [13] value return */
int  nufft_mgpu_get_timings(nufft_mgpu h, int32_t l, float ms[16]);
/* how the z-slab exchanges travel: 1 = peer windows (the sender copies straight into the receiver's buffer over NVLink:
 * CUDA IPC mappings between processes, peer access inside one; stream-ordered barriers around each exchange),
 * 0 = NCCL send / recv (the fallback when a mapping fails on any rank, or NUFFT_B200_MGPU_P2P=0) */
int  nufft_mgpu_exchange_mode(nufft_mgpu h, int32_t *mode);

#ifdef __cplusplus
}
#endif
#endif /* NUFFT_B200_H */
