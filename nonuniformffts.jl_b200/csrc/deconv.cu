// deconv.cu — K-deconv1 / K-deconv2: deconvolution fused with mode truncation / zero padding.
//
// Replaces (reference):
//   src/NonuniformFFTs.jl:387-414  copy_deconvolve_to_non_oversampled_kernel!   (type-1 epilogue)
//   src/NonuniformFFTs.jl:453-480  copy_deconvolve_to_oversampled_kernel!        (type-2 prologue)
//   src/NonuniformFFTs.jl:116-122,260-266  fill_with_zeros_kernel! before the type-2 prologue: fused here —
//   K-deconv2 writes the WHOLE oversampled spectrum in one coalesced pass (zeros outside the kept modes),
//   removing the reference's separate full-grid zero-fill pass.
// Uniform callbacks (src/plan.jl:146-164): multiply by separable tables and/or a dense factor array.
#include "common.cuh"

namespace nufft {

struct DeconvArgs {
    int D;
    int nk[3];         // size(plan)
    int nos[3];        // oversampled spectral dims
    const int32_t *imap[3];
    const int32_t *invmap[3];
    const void *phihat[3];
    const void *fsep[3];     // separable uniform factors or null
    const void *fdense;      // dense uniform factor or null
    int C;
    int64_t nkept, nspec;
};

constexpr int MAXC_PACK = 8;
struct CPack { void *p[MAXC_PACK]; };
struct CCPack { const void *p[MAXC_PACK]; };

// type 1: one thread per kept mode.  out[c][I] = cb( normfactor / prod(phihat) * uhat[c][map(I)] )
template <typename T>
__global__ void __launch_bounds__(256)
deconv_type1_kernel(DeconvArgs a, T normfactor, const typename Vec2<T>::type *__restrict__ uhat, CPack out)
{
    using C2 = typename Vec2<T>::type;
    const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= a.nkept) return;
    int64_t t = I;
    const int i1 = (int)(t % a.nk[0]); t /= a.nk[0];
    const int i2 = (int)(t % a.nk[1]); t /= a.nk[1];
    const int i3 = (int)t;
    T ph = ((const T *)a.phihat[0])[i1];
    int64_t J = a.imap[0][i1];
    T f = (T)1;
    if (a.fsep[0]) f = ((const T *)a.fsep[0])[i1];
    if (a.D > 1) {
        ph *= ((const T *)a.phihat[1])[i2];
        J += (int64_t)a.imap[1][i2] * a.nos[0];
        if (a.fsep[1]) f *= ((const T *)a.fsep[1])[i2];
    }
    if (a.D > 2) {
        ph *= ((const T *)a.phihat[2])[i3];
        J += (int64_t)a.imap[2][i3] * a.nos[0] * a.nos[1];
        if (a.fsep[2]) f *= ((const T *)a.fsep[2])[i3];
    }
    if (a.fdense) f *= ((const T *)a.fdense)[I];
    const T beta = normfactor / ph;
    const bool has_cb = a.fdense || a.fsep[0] || a.fsep[1] || a.fsep[2];
    for (int c = 0; c < a.C; ++c) {
        C2 v = uhat[(int64_t)c * a.nspec + J];
        v.x *= beta; v.y *= beta;
        if (has_cb) { v.x *= f; v.y *= f; }
        ((C2 *)out.p[c])[I] = v;
    }
}

// type 2: one thread per oversampled spectral cell (coalesced full-grid write).
template <typename T>
__global__ void __launch_bounds__(256)
deconv_type2_kernel(DeconvArgs a, typename Vec2<T>::type *__restrict__ uhat, CCPack in)
{
    using C2 = typename Vec2<T>::type;
    const int64_t J = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (J >= a.nspec) return;
    int64_t t = J;
    const int j1 = (int)(t % a.nos[0]); t /= a.nos[0];
    const int j2 = (int)(t % a.nos[1]); t /= a.nos[1];
    const int j3 = (int)t;
    const int i1 = a.invmap[0][j1];
    const int i2 = a.D > 1 ? a.invmap[1][j2] : 0;
    const int i3 = a.D > 2 ? a.invmap[2][j3] : 0;
    const bool kept = (i1 >= 0) && (i2 >= 0) && (i3 >= 0);
    C2 zero; zero.x = 0; zero.y = 0;
    if (!kept) {
        for (int c = 0; c < a.C; ++c) uhat[(int64_t)c * a.nspec + J] = zero;
        return;
    }
    T ph = ((const T *)a.phihat[0])[i1];
    T f = (T)1;
    if (a.fsep[0]) f = ((const T *)a.fsep[0])[i1];
    int64_t I = i1;
    if (a.D > 1) {
        ph *= ((const T *)a.phihat[1])[i2];
        I += (int64_t)i2 * a.nk[0];
        if (a.fsep[1]) f *= ((const T *)a.fsep[1])[i2];
    }
    if (a.D > 2) {
        ph *= ((const T *)a.phihat[2])[i3];
        I += (int64_t)i3 * a.nk[0] * a.nk[1];
        if (a.fsep[2]) f *= ((const T *)a.fsep[2])[i3];
    }
    if (a.fdense) f *= ((const T *)a.fdense)[I];
    const T beta = (T)1 / ph;
    const bool has_cb = a.fdense || a.fsep[0] || a.fsep[1] || a.fsep[2];
    for (int c = 0; c < a.C; ++c) {
        C2 v = ((const C2 *)in.p[c])[I];
        v.x *= beta; v.y *= beta;
        if (has_cb) { v.x *= f; v.y *= f; }
        uhat[(int64_t)c * a.nspec + J] = v;
    }
}

static DeconvArgs make_args(const Plan &p, const nufft_callbacks *cb, int C)
{
    DeconvArgs a{};
    a.D = p.D;
    for (int d = 0; d < 3; ++d) {
        a.nk[d] = (int)p.nk[d];
        a.nos[d] = (int)p.Nspec[d];
        a.imap[d] = p.d_imap[d];
        a.invmap[d] = p.d_invmap[d];
        a.phihat[d] = p.d_phihat[d];
        a.fsep[d] = (cb && cb->u_factor_sep && d < p.D) ? cb->u_factor_sep[d] : nullptr;
    }
    a.fdense = cb ? cb->u_factor_dense : nullptr;
    a.C = C;
    a.nkept = p.nkept;
    a.nspec = p.nspec;
    return a;
}

template <typename T> static int run_type1(Plan &p, void *const uhat[], const nufft_callbacks *cb)
{
    using C2 = typename Vec2<T>::type;
    double nf = 1.0;
    for (int d = 0; d < p.D; ++d) nf *= 2.0 * M_PI / (double)p.Nos[d];     // src/NonuniformFFTs.jl:181
    for (int c0 = 0; c0 < p.C; c0 += MAXC_PACK) {
        const int cn = p.C - c0 < MAXC_PACK ? p.C - c0 : MAXC_PACK;
        CPack out{};
        for (int c = 0; c < cn; ++c) out.p[c] = uhat[c0 + c];
        DeconvArgs a = make_args(p, cb, cn);
        deconv_type1_kernel<T><<<(unsigned)cdiv(p.nkept, 256), 256, 0, p.stream>>>(
            a, (T)nf, (const C2 *)p.d_uhat + (int64_t)c0 * p.nspec, out);
        NUFFT_COUNT_LAUNCH();
    }
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

template <typename T> static int run_type2(Plan &p, const void *const uhat[], const nufft_callbacks *cb)
{
    using C2 = typename Vec2<T>::type;
    for (int c0 = 0; c0 < p.C; c0 += MAXC_PACK) {
        const int cn = p.C - c0 < MAXC_PACK ? p.C - c0 : MAXC_PACK;
        CCPack in{};
        for (int c = 0; c < cn; ++c) in.p[c] = uhat[c0 + c];
        DeconvArgs a = make_args(p, cb, cn);
        deconv_type2_kernel<T><<<(unsigned)cdiv(p.nspec, 256), 256, 0, p.stream>>>(
            a, (C2 *)p.d_uhat + (int64_t)c0 * p.nspec, in);
        NUFFT_COUNT_LAUNCH();
    }
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

int deconv_type1_run(Plan &p, void *const uhat[], const nufft_callbacks *cb)
{
    return p.f64 ? run_type1<double>(p, uhat, cb) : run_type1<float>(p, uhat, cb);
}

int deconv_type2_run(Plan &p, const void *const uhat[], const nufft_callbacks *cb)
{
    return p.f64 ? run_type2<double>(p, uhat, cb) : run_type2<float>(p, uhat, cb);
}

}  // namespace nufft
