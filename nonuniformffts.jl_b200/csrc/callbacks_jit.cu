// callbacks_jit.cu — general NUFFTCallbacks compiled at run time with NVRTC.
//
// The reference lets the user pass arbitrary Julia closures as callbacks (src/plan.jl:146-164): `nonuniform(v, n)` on the
// tuple of C values of point n (type 1: before spreading, src/spreading/gpu.jl:34,326-340; type 2: after interpolation,
// src/interpolation/gpu.jl:31,288-325) and `uniform(w, idx)` on the tuple of C coefficients of mode idx (type 1: after
// deconvolution, src/NonuniformFFTs.jl:398; type 2: before it, :464).  A compiled library cannot inline closures; beyond the
// menu of nufft_callbacks (weights / factors, fused into the kernels) the caller may hand CUDA C++ source that defines
//
//     #define NUFFT_HAS_NONUNIFORM 1
//     __device__ void nufft_cb_nonuniform(nufft_cell (&v)[NUFFT_C], long long n, const void *user);     // n: 0-based ORIGINAL index
//     #define NUFFT_HAS_UNIFORM 1
//     __device__ void nufft_cb_uniform(nufft_cplx (&w)[NUFFT_C], const int (&idx)[3], const void *user); // idx: 0-based, idx[0] fastest
//
// (either or both; nufft_real = float | double, nufft_cplx = float2 | double2, nufft_cell = nufft_real for real data and
// nufft_cplx for complex data, NUFFT_C = ntransforms, NUFFT_D = dimensions).  The source is compiled once per plan for
// sm_100a (cached by content) and applied as one coalesced elementwise pass per callback: input-side callbacks write a
// plan-owned copy (inputs of exec are never modified, src/plan.jl:72-73), output-side callbacks run in place.
#include <nvrtc.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace nufft {

struct JitPtrs {
    void *p[8];
};

struct JitCallbacks {
    std::string src;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k_nu = nullptr, k_u = nullptr;
    bool has_nu = false, has_u = false;
    void *d_tmp_nu = nullptr, *d_tmp_u = nullptr;
    size_t tmp_nu_bytes = 0, tmp_u_bytes = 0;
};

static const char *kWrapper = R"SRC(
struct NufftPtrs { void *p[8]; };
extern "C" __global__ void nufft_cb_nu_kernel(long long np, NufftPtrs in, NufftPtrs out, const void *user)
{
#ifdef NUFFT_HAS_NONUNIFORM
    const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (n >= np) return;
    nufft_cell v[NUFFT_C];
#pragma unroll
    for (int c = 0; c < NUFFT_C; ++c) v[c] = ((const nufft_cell *)in.p[c])[n];
    nufft_cb_nonuniform(v, n, user);
#pragma unroll
    for (int c = 0; c < NUFFT_C; ++c) ((nufft_cell *)out.p[c])[n] = v[c];
#endif
}
extern "C" __global__ void nufft_cb_u_kernel(long long nk, int n0, int n1, NufftPtrs in, NufftPtrs out, const void *user)
{
#ifdef NUFFT_HAS_UNIFORM
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nk) return;
    const long long r = i / n0;
    const int idx[3] = {(int)(i - r * n0), (int)(r % n1), (int)(r / n1)};
    nufft_cplx w[NUFFT_C];
#pragma unroll
    for (int c = 0; c < NUFFT_C; ++c) w[c] = ((const nufft_cplx *)in.p[c])[i];
    nufft_cb_uniform(w, idx, user);
#pragma unroll
    for (int c = 0; c < NUFFT_C; ++c) ((nufft_cplx *)out.p[c])[i] = w[c];
#endif
}
)SRC";

// fields beyond the original three are optional: read them only when the caller's struct is large enough
const char *callbacks_source(const nufft_callbacks *cb)
{
    if (!cb || cb->struct_size < offsetof(nufft_callbacks, nvrtc_src) + sizeof(const char *)) return nullptr;
    return (cb->nvrtc_src && cb->nvrtc_src[0]) ? cb->nvrtc_src : nullptr;
}
static const void *callbacks_user(const nufft_callbacks *cb)
{
    if (!cb || cb->struct_size < offsetof(nufft_callbacks, user_data) + sizeof(const void *)) return nullptr;
    return cb->user_data;
}

void jit_callbacks_free(Plan &p)
{
    JitCallbacks *j = (JitCallbacks *)p.jit;
    if (!j) return;
    if (j->lib) cudaLibraryUnload(j->lib);
    if (j->d_tmp_nu) cudaFree(j->d_tmp_nu);
    if (j->d_tmp_u) cudaFree(j->d_tmp_u);
    delete j;
    p.jit = nullptr;
}

// compile (or fetch from the plan's cache) the callbacks of `cb`; *out = nullptr when cb carries no source
int jit_callbacks_get(Plan &p, const nufft_callbacks *cb, JitCallbacks **out)
{
    *out = nullptr;
    const char *src = callbacks_source(cb);
    if (!src) return NUFFT_SUCCESS;
    JitCallbacks *j = (JitCallbacks *)p.jit;
    if (j && j->src == src) { *out = j; return NUFFT_SUCCESS; }
    if (p.C > 8) { set_error("run-time compiled callbacks support ntransforms <= 8 (got %d)", p.C); return NUFFT_ERR_UNSUPPORTED; }
    if (j) jit_callbacks_free(p);

    std::string full;
    char buf[512];
    snprintf(buf, sizeof buf,
             "#define NUFFT_C %d\n#define NUFFT_D %d\n#define NUFFT_IS_COMPLEX %d\ntypedef %s nufft_real;\ntypedef %s nufft_cplx;\n"
             "typedef %s nufft_cell;\n",
             p.C, p.D, p.cplx ? 1 : 0, p.f64 ? "double" : "float", p.f64 ? "double2" : "float2",
             p.cplx ? (p.f64 ? "double2" : "float2") : (p.f64 ? "double" : "float"));
    full = buf;
    full += src;
    full += "\n";
    full += kWrapper;

    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, full.c_str(), "nufft_callbacks.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        set_error("nvrtcCreateProgram failed");
        return NUFFT_ERR_CUDA;
    }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device"};
    const nvrtcResult rc = nvrtcCompileProgram(prog, 3, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        set_error("callback source does not compile (%s):\n%.1500s", nvrtcGetErrorString(rc), log.c_str());
        return NUFFT_ERR_ARG;
    }
    size_t nbin = 0;
    nvrtcGetCUBINSize(prog, &nbin);
    std::vector<char> cubin(nbin);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);

    // the plan only ever caches a fully resolved object: a failure below leaves p.jit empty, so the next call recompiles
    j = new JitCallbacks();
    auto fail = [&](cudaError_t e, const char *what) {
        set_error("CUDA error %s in %s: %s", cudaGetErrorName(e), what, cudaGetErrorString(e));
        if (j->lib) cudaLibraryUnload(j->lib);
        delete j;
        cudaGetLastError();
        return NUFFT_ERR_CUDA;
    };
    cudaError_t e = cudaLibraryLoadData(&j->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) { j->lib = nullptr; return fail(e, "cudaLibraryLoadData"); }
    if ((e = cudaLibraryGetKernel(&j->k_nu, j->lib, "nufft_cb_nu_kernel")) != cudaSuccess) return fail(e, "cudaLibraryGetKernel");
    if ((e = cudaLibraryGetKernel(&j->k_u, j->lib, "nufft_cb_u_kernel")) != cudaSuccess) return fail(e, "cudaLibraryGetKernel");
    j->has_nu = strstr(src, "NUFFT_HAS_NONUNIFORM") != nullptr;
    j->has_u = strstr(src, "NUFFT_HAS_UNIFORM") != nullptr;
    j->src = src;
    p.jit = j;
    *out = j;
    return NUFFT_SUCCESS;
}

bool jit_has_nonuniform(const JitCallbacks *j) { return j && j->has_nu; }
bool jit_has_uniform(const JitCallbacks *j) { return j && j->has_u; }

// v[c][n] -> cb(v, n); in == nullptr: in place on out.  tmp = true: out are plan-owned copies (returned in `out`)
int jit_apply_nonuniform(Plan &p, JitCallbacks *j, const nufft_callbacks *cb, const void *const in[], void *out[], bool tmp)
{
    const size_t zbytes = p.real_bytes * (p.cplx ? 2 : 1);
    if (tmp) {
        const size_t need = (size_t)p.C * (size_t)std::max<int64_t>(p.Np, 1) * zbytes;
        if (need > j->tmp_nu_bytes) {
            if (j->d_tmp_nu) cudaFree(j->d_tmp_nu);
            j->tmp_nu_bytes = 0;
            CUDA_TRY(cudaMalloc(&j->d_tmp_nu, need));
            j->tmp_nu_bytes = need;
        }
        for (int c = 0; c < p.C; ++c) out[c] = (char *)j->d_tmp_nu + (size_t)c * p.Np * zbytes;
    }
    if (p.Np == 0) return NUFFT_SUCCESS;
    JitPtrs pi{}, po{};
    for (int c = 0; c < p.C; ++c) { pi.p[c] = (void *)(in ? in[c] : out[c]); po.p[c] = out[c]; }
    long long np = p.Np;
    const void *user = callbacks_user(cb);
    void *args[] = {&np, &pi, &po, &user};
    CUDA_TRY(cudaLaunchKernel((const void *)j->k_nu, dim3((unsigned)cdiv(np, 256)), dim3(256), args, 0, p.stream));
    NUFFT_COUNT_LAUNCH();
    return NUFFT_SUCCESS;
}

int jit_apply_uniform(Plan &p, JitCallbacks *j, const nufft_callbacks *cb, const void *const in[], void *out[], bool tmp)
{
    const size_t cbytes = 2 * p.real_bytes;
    if (tmp) {
        const size_t need = (size_t)p.C * (size_t)p.nkept * cbytes;
        if (need > j->tmp_u_bytes) {
            if (j->d_tmp_u) cudaFree(j->d_tmp_u);
            j->tmp_u_bytes = 0;
            CUDA_TRY(cudaMalloc(&j->d_tmp_u, need));
            j->tmp_u_bytes = need;
        }
        for (int c = 0; c < p.C; ++c) out[c] = (char *)j->d_tmp_u + (size_t)c * p.nkept * cbytes;
    }
    JitPtrs pi{}, po{};
    for (int c = 0; c < p.C; ++c) { pi.p[c] = (void *)(in ? in[c] : out[c]); po.p[c] = out[c]; }
    long long nk = p.nkept;
    int n0 = (int)p.nk[0], n1 = (int)p.nk[1];
    const void *user = callbacks_user(cb);
    void *args[] = {&nk, &n0, &n1, &pi, &po, &user};
    CUDA_TRY(cudaLaunchKernel((const void *)j->k_u, dim3((unsigned)cdiv(nk, 256)), dim3(256), args, 0, p.stream));
    NUFFT_COUNT_LAUNCH();
    return NUFFT_SUCCESS;
}

}  // namespace nufft
