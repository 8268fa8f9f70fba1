// api.cu — the C ABI (include/nufft_b200.h): argument validation, stage sequencing, timing events.
//
// Stage order is the reference's (src/NonuniformFFTs.jl:148-189 type 1, :237-286 type 2):
//   type 1: (0) zero grid -> (1) spread -> (2) forward FFT -> (3) deconvolve + truncate
//   type 2: (0)+(1) deconvolve + zero-pad (one pass) -> (2) backward FFT -> (3) interpolate
#include <cstdarg>
#include <cmath>
#include <new>
#include <vector>

#include "common.cuh"
#include "spread.cuh"
#include "interp.cuh"

namespace nufft {

thread_local int64_t g_launch_count = 0;
static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int spread_run(Plan &p, const void *const vp[], const nufft_callbacks *cb)
{
    if (p.f64) return p.cplx ? spread_dispatch<double, true>(p, vp, cb) : spread_dispatch<double, false>(p, vp, cb);
    return p.cplx ? spread_dispatch<float, true>(p, vp, cb) : spread_dispatch<float, false>(p, vp, cb);
}

int interp_run(Plan &p, void *const vp[], const nufft_callbacks *cb)
{
    if (p.f64) return p.cplx ? interp_dispatch<double, true>(p, vp, cb) : interp_dispatch<double, false>(p, vp, cb);
    return p.cplx ? interp_dispatch<float, true>(p, vp, cb) : interp_dispatch<float, false>(p, vp, cb);
}

static inline void rec(Plan &p, int i)
{
    if (p.ev_ok) { cudaEventRecord(p.ev[i], p.stream); }
}

// Selects the plan's device for the duration of an entry point and restores the caller's device on every return path
// (a host that drives several GPUs from one thread must not find its current device changed by a library call).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
        else if (err == cudaSuccess) prev = -1;       // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define NUFFT_DEVICE_GUARD(p_) DeviceGuard _guard((p_).device); CUDA_TRY(_guard.err)

static int check_plan(nufft_plan h)
{
    if (!h) { set_error("null plan handle"); return NUFFT_ERR_STATE; }
    return NUFFT_SUCCESS;
}

static int check_exec(Plan &p, const void *const a[], const void *const b[])
{
    if (p.Np < 0) { set_error("set_points must be called before exec_type1 / exec_type2"); return NUFFT_ERR_STATE; }
    if (!a || !b) { set_error("null array-of-pointers argument"); return NUFFT_ERR_ARG; }
    for (int c = 0; c < p.C; ++c) {
        if (!a[c] && p.nkept > 0) { set_error("null uniform array for transform %d (expected a tuple of %d arrays)", c, p.C); return NUFFT_ERR_DIM; }
        if (!b[c] && p.Np > 0) { set_error("null non-uniform data vector for transform %d (expected a tuple of %d vectors)", c, p.C); return NUFFT_ERR_DIM; }
    }
    return NUFFT_SUCCESS;
}

}  // namespace nufft

namespace nufft {
int host_kernel_tables(const nufft_opts &o, int d, double *shape, double *dx, int64_t *os_dim, double *cs, size_t cs_len, double *phihat,
                       size_t phihat_len);          // host_plan.cu
}

using namespace nufft;

extern "C" {

int nufft_abi_version(void) { return NUFFT_B200_ABI_VERSION; }

const char *nufft_last_error(void) { return g_err; }

int64_t nufft_launch_count(int reset)
{
    const int64_t v = g_launch_count;
    if (reset) g_launch_count = 0;
    return v;
}

int nufft_opts_default(nufft_opts *o)
{
    if (!o) { set_error("null opts"); return NUFFT_ERR_ARG; }
    memset(o, 0, sizeof(*o));
    o->struct_size = (uint32_t)sizeof(nufft_opts);
    o->dim = 1;
    o->n_modes[0] = o->n_modes[1] = o->n_modes[2] = 1;
    o->is_complex = 1;                               // PlanNUFFT(N) defaults to ComplexF64 (src/plan.jl:597-599)
    o->dtype = NUFFT_F64;
    o->half_support = 4;                             // src/plan.jl:583
    o->sigma = 2.0;                                  // src/plan.jl:573
    o->kernel = NUFFT_KERNEL_KAISER_BESSEL;          // ext/NonuniformFFTsCUDAExt.jl:19
    o->kernel_param = NAN;
    o->eval_mode = NUFFT_EVAL_DIRECT;                // ext/NonuniformFFTsCUDAExt.jl:23
    o->ntransforms = 1;
    o->gpu_method = NUFFT_METHOD_AUTO;
    o->device = -1;
    return NUFFT_SUCCESS;
}

int nufft_plan_create(nufft_plan *out, const nufft_opts *opts)
{
    if (!out || !opts) { set_error("null argument"); return NUFFT_ERR_ARG; }
    *out = nullptr;
    if (opts->struct_size != sizeof(nufft_opts)) {
        set_error("nufft_opts.struct_size = %u does not match this library (%zu): ABI mismatch", opts->struct_size, sizeof(nufft_opts));
        return NUFFT_ERR_ARG;
    }
    Plan *p = new (std::nothrow) Plan();
    if (!p) { set_error("out of host memory"); return NUFFT_ERR_ALLOC; }
    p->opts = *opts;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    const int rc = host_plan_init(*p);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    if (rc != NUFFT_SUCCESS) {
        host_plan_free(*p);
        delete p;
        return rc;
    }
    *out = reinterpret_cast<nufft_plan>(p);
    return NUFFT_SUCCESS;
}

int nufft_plan_destroy(nufft_plan h)
{
    if (!h) return NUFFT_SUCCESS;
    Plan *p = reinterpret_cast<Plan *>(h);
    DeviceGuard guard(p->device);
    cudaStreamSynchronize(p->stream);
    host_plan_free(*p);
    delete p;
    return NUFFT_SUCCESS;
}

int nufft_plan_shape(nufft_plan h, int64_t size_out[3], int64_t os_dims[3], int32_t *ntransforms)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    for (int d = 0; d < 3; ++d) {
        if (size_out) size_out[d] = p.nk[d];
        if (os_dims) os_dims[d] = p.Nos[d];
    }
    if (ntransforms) *ntransforms = p.C;
    return NUFFT_SUCCESS;
}

int nufft_plan_kernel_info(nufft_plan h, int32_t d, double *shape_param, double *dx, double *cs_host, double *phihat_host)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (d < 0 || d >= p.D) { set_error("dimension %d out of range", d); return NUFFT_ERR_ARG; }
    if (shape_param) *shape_param = p.h_shape[d];
    if (dx) *dx = p.h_dx[d];
    const int ncs = (p.M + 4) * 2 * p.M;
    if (cs_host) {
        if (p.f64) memcpy(cs_host, p.h_cs[d].data(), (size_t)ncs * sizeof(double));
        else for (int i = 0; i < ncs; ++i) cs_host[i] = (double)((const float *)p.h_cs[d].data())[i];
    }
    if (phihat_host) {
        if (p.f64) memcpy(phihat_host, p.h_phihat[d].data(), (size_t)p.nk[d] * sizeof(double));
        else for (int64_t i = 0; i < p.nk[d]; ++i) phihat_host[i] = (double)((const float *)p.h_phihat[d].data())[i];
    }
    return NUFFT_SUCCESS;
}

int nufft_kernel_tables(const nufft_opts *opts, int32_t d, double *shape_param, double *dx, int64_t *os_dim, double *cs_host, size_t cs_len,
                        double *phihat_host, size_t phihat_len)
{
    if (!opts) { set_error("null options"); return NUFFT_ERR_ARG; }
    if (opts->struct_size != sizeof(nufft_opts)) { set_error("nufft_opts.struct_size does not match this library: ABI mismatch"); return NUFFT_ERR_ARG; }
    return nufft::host_kernel_tables(*opts, d, shape_param, dx, os_dim, cs_host, cs_len, phihat_host, phihat_len);
}

int nufft_set_points(nufft_plan h, int64_t np, const void *const x[])
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (!x) { set_error("null point array list"); return NUFFT_ERR_ARG; }
    NUFFT_DEVICE_GUARD(p);
    return binning_set_points(p, np, x);
}

int nufft_set_points_matrix(nufft_plan h, int64_t np, const void *xmat)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (!xmat && np > 0) { set_error("null point matrix"); return NUFFT_ERR_ARG; }
    NUFFT_DEVICE_GUARD(p);
    // column-major (D, Np) matrix == array of D-vectors: coordinate d of point i at xmat[i * D + d]; read in place by K-bin
    const void *x[3] = {nullptr, nullptr, nullptr};
    for (int d = 0; d < p.D; ++d) x[d] = (const char *)xmat + (size_t)d * p.real_bytes;
    return binning_set_points(p, np, x, p.D);
}

int nufft_get_binning(nufft_plan h, const int32_t **perm, const int32_t **bin_offsets, int64_t *nbins, int64_t bin_dims[3])
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (p.Np < 0) { set_error("set_points has not been called"); return NUFFT_ERR_STATE; }
    if (perm) NUFFT_TRY(binning_coarse_perm(p, perm));     // rt plans: reference-order permutation rebuilt on demand
    if (bin_offsets) { NUFFT_TRY(binning_ensure_offsets(p)); *bin_offsets = p.d_bin_offsets; }
    if (nbins) *nbins = p.nbins;
    if (bin_dims) for (int d = 0; d < 3; ++d) bin_dims[d] = p.geom.B[d];
    return NUFFT_SUCCESS;
}

int nufft_get_binning_fine(nufft_plan h, const int32_t **perm, const int32_t **fine_offsets, int64_t *nfine, int64_t sub_dims[3])
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (p.Np < 0) { set_error("set_points has not been called"); return NUFFT_ERR_STATE; }
    if (perm) *perm = p.d_perm;
    if (fine_offsets) *fine_offsets = p.geom.nsub > 1 ? nullptr : p.d_bin_offsets;   // sub-bin offsets are not materialised
    if (nfine) *nfine = p.geom.nsub > 1 ? 0 : p.nbins;
    if (sub_dims) for (int d = 0; d < 3; ++d) sub_dims[d] = p.geom.sub[d];
    return NUFFT_SUCCESS;
}

int nufft_type1_spread(nufft_plan h, const void *const vp[], const nufft_callbacks *cb)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (p.Np < 0) { set_error("set_points must be called before exec_type1"); return NUFFT_ERR_STATE; }
    if (!vp) { set_error("null vp"); return NUFFT_ERR_ARG; }
    NUFFT_DEVICE_GUARD(p);
    rec(p, 2);
    const size_t zbytes = p.real_bytes * (p.cplx ? 2 : 1);
    CUDA_TRY(cudaMemsetAsync(p.d_us, 0, (size_t)p.C * p.ncells * zbytes, p.stream));
    NUFFT_COUNT_LAUNCH();
    rec(p, 3);
    {
        JitCallbacks *j = nullptr;                 // general (run-time compiled) nonuniform callback: v -> cb(v, n) on a copy
        NUFFT_TRY(jit_callbacks_get(p, cb, &j));
        if (jit_has_nonuniform(j)) {
            std::vector<void *> tmp(p.C);
            NUFFT_TRY(jit_apply_nonuniform(p, j, cb, vp, tmp.data(), true));
            NUFFT_TRY(spread_run(p, (const void *const *)tmp.data(), cb));
        } else {
            NUFFT_TRY(spread_run(p, vp, cb));
        }
    }
    rec(p, 4);
    p.ev_rec[1] = true;
    return NUFFT_SUCCESS;
}

int nufft_type1_finish(nufft_plan h, void *const uhat[], const nufft_callbacks *cb)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (!uhat) { set_error("null uhat"); return NUFFT_ERR_ARG; }
    NUFFT_DEVICE_GUARD(p);
    rec(p, 5);
    if (p.pfft) {
        NUFFT_TRY(pfft_type1_run(p, uhat, cb));     // truncating FFT passes with the deconvolution fused in
        rec(p, 6);
    } else {
        NUFFT_TRY(fft_forward(p));
        rec(p, 6);
        NUFFT_TRY(deconv_type1_run(p, uhat, cb));
    }
    {
        JitCallbacks *j = nullptr;                 // general uniform callback on the output coefficients, in place
        NUFFT_TRY(jit_callbacks_get(p, cb, &j));
        if (jit_has_uniform(j)) NUFFT_TRY(jit_apply_uniform(p, j, cb, nullptr, (void **)uhat, false));
    }
    rec(p, 7);
    p.ev_rec[2] = true;
    return NUFFT_SUCCESS;
}

int nufft_exec_type1(nufft_plan h, void *const uhat[], const void *const vp[], const nufft_callbacks *cb)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    NUFFT_TRY(check_exec(p, (const void *const *)uhat, vp));
    NUFFT_TRY(nufft_type1_spread(h, vp, cb));
    return nufft_type1_finish(h, uhat, cb);
}

int nufft_type2_prepare(nufft_plan h, const void *const uhat[], const nufft_callbacks *cb)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (!uhat) { set_error("null uhat"); return NUFFT_ERR_ARG; }
    NUFFT_DEVICE_GUARD(p);
    rec(p, 8);
    std::vector<void *> tmp_u(p.C);
    {
        JitCallbacks *j = nullptr;                 // general uniform callback on a copy of the input coefficients
        NUFFT_TRY(jit_callbacks_get(p, cb, &j));
        if (jit_has_uniform(j)) {
            NUFFT_TRY(jit_apply_uniform(p, j, cb, uhat, tmp_u.data(), true));
            uhat = (const void *const *)tmp_u.data();
        }
    }
    if (p.pfft) {
        rec(p, 9);
        NUFFT_TRY(pfft_type2_run(p, uhat, cb));     // zero-padding FFT passes with the deconvolution fused in
    } else {
        NUFFT_TRY(deconv_type2_run(p, uhat, cb));
        rec(p, 9);
        NUFFT_TRY(fft_backward(p));
    }
    rec(p, 10);
    p.ev_rec[3] = true;
    return NUFFT_SUCCESS;
}

int nufft_type2_interp(nufft_plan h, void *const vp[], const nufft_callbacks *cb)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (p.Np < 0) { set_error("set_points must be called before exec_type2"); return NUFFT_ERR_STATE; }
    if (!vp) { set_error("null vp"); return NUFFT_ERR_ARG; }
    NUFFT_DEVICE_GUARD(p);
    rec(p, 11);
    NUFFT_TRY(interp_run(p, vp, cb));
    {
        JitCallbacks *j = nullptr;                 // general nonuniform callback on the interpolated values, in place
        NUFFT_TRY(jit_callbacks_get(p, cb, &j));
        if (jit_has_nonuniform(j)) NUFFT_TRY(jit_apply_nonuniform(p, j, cb, nullptr, (void **)vp, false));
    }
    rec(p, 12);
    p.ev_rec[4] = true;
    return NUFFT_SUCCESS;
}

int nufft_exec_type2(nufft_plan h, void *const vp[], const void *const uhat[], const nufft_callbacks *cb)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    NUFFT_TRY(check_exec(p, uhat, (const void *const *)vp));
    NUFFT_TRY(nufft_type2_prepare(h, uhat, cb));
    return nufft_type2_interp(h, vp, cb);
}

int nufft_get_grid(nufft_plan h, void **grid, size_t *bytes_per_transform)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (grid) *grid = p.d_us;
    if (bytes_per_transform) *bytes_per_transform = (size_t)p.ncells * p.real_bytes * (p.cplx ? 2 : 1);
    return NUFFT_SUCCESS;
}

int nufft_get_timings(nufft_plan h, float ms[16])
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (!p.ev_ok) { set_error("plan was created with record_timings = 0"); return NUFFT_ERR_STATE; }
    CUDA_TRY(cudaStreamSynchronize(p.stream));
    auto el = [&](int a, int b, float &dst) { float t = 0; if (cudaEventElapsedTime(&t, p.ev[a], p.ev[b]) == cudaSuccess) dst = t; else cudaGetLastError(); };
    if (p.ev_rec[0]) el(0, 1, p.ms[0]);
    if (p.ev_rec[1]) { el(2, 3, p.ms[1]); el(3, 4, p.ms[2]); }
    if (p.ev_rec[2]) { el(5, 6, p.ms[3]); el(6, 7, p.ms[4]); }
    if (p.ev_rec[3]) { el(8, 9, p.ms[5]); el(9, 10, p.ms[6]); }
    if (p.ev_rec[4]) el(11, 12, p.ms[7]);
    for (int i = 0; i < 16; ++i) ms[i] = p.ms[i];
    return NUFFT_SUCCESS;
}

int nufft_describe(nufft_plan h, char *buf, size_t buflen)
{
    NUFFT_TRY(check_plan(h));
    Plan &p = *reinterpret_cast<Plan *>(h);
    if (!buf || buflen == 0) { set_error("null buffer"); return NUFFT_ERR_ARG; }
    static const char *knames[] = {"KaiserBesselKernel", "BackwardsKaiserBesselKernel", "GaussianKernel", "BSplineKernel", "ESKernel"};
    const char *shape = (p.opts.kernel <= 1 || p.opts.kernel == NUFFT_KERNEL_ES) ? "beta" : (p.opts.kernel == 2 ? "tau" : "-");
    double sigma = 0;
    for (int d = 0; d < p.D; ++d) sigma = std::fmax(sigma, (double)p.Nos[d] / (double)p.Ns[d]);
    const TileGeom &g = p.geom;
    snprintf(buf, buflen,
             "%d-dimensional PlanNUFFT with input type %s%s:\n"
             "  - backend: B200 (sm_100a) native CUDA\n"
             "  - kernel: %s(%s = %.17g) with half-support M = %d\n"
             "  - kernel evaluation mode: %s\n"
             "  - oversampling factor: sigma = %.6g\n"
             "  - uniform dimensions: (%lld, %lld, %lld)\n"
             "  - oversampled dimensions: (%lld, %lld, %lld)\n"
             "  - simultaneous transforms: %d\n"
             "  - frequency order: %s (fftshift = %s)\n"
             "  - block size: (%d, %d, %d) (excluding 2M - 1 = %d ghost cells in each direction), %lld bins\n"
             "  - GPU method: :%s\n"
             "  - FFT: %s\n"
             "  - tile (shared memory): (%d, %d, %d) cells, row stride %d, batch %d points, chunk %d points%s\n",
             p.D, p.cplx ? "Complex" : "", p.f64 ? "Float64" : "Float32", knames[p.opts.kernel], shape, p.h_shape[0], p.M,
             p.opts.eval_mode == NUFFT_EVAL_FAST ? "FastApproximation" : "Direct", sigma,
             (long long)p.nk[0], (long long)p.nk[1], (long long)p.nk[2],
             (long long)p.Nos[0], (long long)p.Nos[1], (long long)p.Nos[2], p.C,
             p.opts.fftshift ? "increasing" : "FFTW", p.opts.fftshift ? "true" : "false",
             g.B[0], g.B[1], g.B[2], 2 * p.M - 1, (long long)p.nbins,
             p.method == NUFFT_METHOD_SHARED_MEMORY ? "shared_memory" : "global_memory",
             p.pfft ? "pruned 1-D passes fused with deconvolution (pfft.cu)" : "cuFFT + separate deconvolution",
             g.T[0], g.T[1], g.T[2], g.S[0], g.batch, g.chunk,
             g.rt == 3 ? (p.cplx ? ", ring-window register kernels" : ", column-streaming register kernels") : "");
    return NUFFT_SUCCESS;
}

}  // extern "C"
