// kernel_eval.cuh — device-side evaluation of the spreading kernels (window functions).
//
// Replaces (reference, relative to /root/reference):
//   src/Kernels/Kernels.jl:121-139            point_to_cell, evaluate_kernel dispatch
//   src/Kernels/piecewise_polynomial.jl:76-92 evaluate_piecewise / evaluate_horner   (KB, BKB fast mode)
//   src/Kernels/kaiser_bessel.jl:177-210      KB fast / direct (I0)
//   src/Kernels/kaiser_bessel_backwards.jl:147-175  BKB fast / direct (sinh)
//   src/Kernels/gaussian.jl:125-192           fast Gaussian gridding / direct
//   src/Kernels/bspline.jl:99-119,143-193     de Boor recursion (same for both modes)
//   src/blocking/blocking.jl:12-21            fold to [0, 2pi)
//   src/abstractNFFTs.jl:150-158              AbstractNFFTs point convention
#pragma once
#include "common.cuh"

namespace nufft {

template <typename T> __device__ __forceinline__ T two_pi() { return (T)2 * (T)3.14159265358979323846; }

__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float exp_t(float x) { return expf(x); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ float sinh_t(float x) { return sinhf(x); }
__device__ __forceinline__ double sinh_t(double x) { return sinh(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float i0_t(float x) { return cyl_bessel_i0f(x); }
__device__ __forceinline__ double i0_t(double x) { return cyl_bessel_i0(x); }
__device__ __forceinline__ float fmod_t(float x, float y) { return fmodf(x, y); }
__device__ __forceinline__ double fmod_t(double x, double y) { return fmod(x, y); }   // (a float-only overload would fold far-away Float64 points at float precision)

// Fold x onto [0, 2pi) with the reference CPU loop (bit-identical bins); points further than
// 64 periods away are first reduced with fmod so the loop is bounded.
template <typename T> __device__ __forceinline__ T fold_point(T x, int convention)
{
    const T L = two_pi<T>();
    if (convention == 1) {           // x in [-1/2, 1/2): 2pi*x, flip sign, shift negative values
        x = mul_rn(L, x);
        x = -x;
        if (x < 0) x = add_rn(x, L);
    }
    if (!(x > -(T)64 * L && x < (T)64 * L)) {
        if (!isfinite(x)) return (T)0;
        x = fmod_t(x, L);
    }
    while (x < 0) x = add_rn(x, L);
    while (x >= L) x = add_rn(x, -L);
    return x;
}

// point_to_cell: r = (x / L) * N in this order (test/near_2pi.jl:19-46); 0-based cell.
template <typename T> __device__ __forceinline__ int point_to_cell0(T x, int N, T &r)
{
    r = mul_rn(div_rn(x, two_pi<T>()), (T)N);
    int i = (int)r;               // trunc; r >= 0
    return i < N ? i : N - 1;     // guard (never taken for x < L, see the reference test)
}

// B-spline values by de Boor recursion, order K = 2M (bspline.jl:143-193, @generated branch).
template <typename T, int K> __device__ __forceinline__ void bspline_all(T x, T *out)
{
    T bp[K], bq[K];
    bp[0] = (T)1;
#pragma unroll
    for (int q = 2; q <= K; ++q) {
        const T alpha = (T)1 / (T)(q - 1);
        // ds[j] = alpha * (x + j), j = 0..q-2
        bq[0] = (alpha * x) * bp[0];
#pragma unroll
        for (int j = 2; j <= q - 1; ++j) {
            T dsm = alpha * (x + (T)(j - 2));
            T dsj = alpha * (x + (T)(j - 1));
            bq[j - 1] = ((T)1 - dsm) * bp[j - 2] + dsj * bp[j - 1];
        }
        bq[q - 1] = ((T)1 - alpha * (x + (T)(q - 2))) * bp[q - 2];
#pragma unroll
        for (int j = 0; j < q; ++j) bp[j] = bq[j];
    }
#pragma unroll
    for (int j = 0; j < K; ++j) out[j] = bp[j];
}

// Evaluate the 2M kernel values of dimension d around folded point x.
// cs points to this dimension's coefficient block (global or shared memory).
// Returns the 0-based cell index i0; the values belong to grid cells i0-M+1 ... i0+M.
template <typename T, int M>
__device__ __forceinline__ int eval_kernel_values(const KernelParams<T> &kp, const T *cs, int d, T x, T *w)
{
    constexpr int W = 2 * M;
    T r;
    const int i0 = point_to_cell0<T>(x, kp.N[d], r);
    const T X = r - (T)i0;
    const int kind = kp.kind;
    if (kind == NUFFT_KERNEL_BSPLINE) {
        // x' = i - r with 1-based i (bspline.jl:103)
        bspline_all<T, W>((T)(i0 + 1) - r, w);
        return i0;
    }
    if (kp.mode == NUFFT_EVAL_FAST) {
        if (kind == NUFFT_KERNEL_GAUSSIAN) {
            const T tau = kp.tau[d], dx = kp.dx[d];
            const T *gcs = cs + (M + 4) * W;
            const T Xp = x - (T)i0 * dx;
            const T a = exp_t(-(Xp * Xp) / tau);
            const T b = exp_t((T)2 * Xp * dx / tau);
            T bpow = (T)1;
            w[M - 1] = a;
#pragma unroll
            for (int m = 1; m <= M - 1; ++m) {
                bpow *= b;
                w[M - m - 1] = a * gcs[m - 1] / bpow;
                w[M + m - 1] = a * gcs[m - 1] * bpow;
            }
            w[W - 1] = a * gcs[M - 1] * bpow * b;
        } else {
            // piecewise polynomial of degree M+3 on each of the 2M sub-intervals, Horner
            const T xt = (T)2 * X - (T)1;
#pragma unroll
            for (int j = 0; j < W; ++j) w[j] = cs[(M + 3) * W + j];
#pragma unroll
            for (int p = M + 2; p >= 0; --p) {
#pragma unroll
                for (int j = 0; j < W; ++j) w[j] = fma_t(xt, w[j], cs[p * W + j]);
            }
        }
        return i0;
    }
    // Direct evaluation from the definition
    if (kind == NUFFT_KERNEL_KAISER_BESSEL && kp.i0_terms > 0) {
        // KB: w = I0(beta sqrt(1 - y^2)) by the power series of I0 in t = 1 - y^2, I0(beta sqrt(t)) = sum_k c_k t^k with
        // c_k = (beta^2 / 4)^k / (k!)^2 tabulated per dimension at plan creation (behind the polynomial block of `cs`).  All
        // terms are positive: relative accuracy ~ nterms * eps everywhere.  The 2M arguments advance in lockstep: 2M
        // independent Horner chains instead of 2M calls of the library cyl_bessel_i0, each one long dependent chain — with
        // the two producer warps of the tile kernels (one lane per point in the column-streaming kernels) the library version
        // was latency-bound and cost more than the spreading itself (Float64: 32.9 -> 17.1 ms, Float32: 10.7 -> 5.6 ms per
        // transform on the reference's benchmark at rho = 1).  No square root either.
        const T *ic = cs + (M + 4) * W + M;
        const int nt = kp.i0_terms;
        T t[W], acc[W];
        const T ctop = ic[nt - 1];
#pragma unroll
        for (int j = 1; j <= W; ++j) {
            const T y = ((T)(M - j) + X) / (T)M;
            const T z = (T)1 - y * y;
            t[j - 1] = z < (T)0 ? (T)0 : z;
            acc[j - 1] = ctop;
        }
        for (int k = nt - 2; k >= 0; --k) {
            const T c = ic[k];
#pragma unroll
            for (int j = 0; j < W; ++j) acc[j] = fma_t(acc[j], t[j], c);
        }
#pragma unroll
        for (int j = 0; j < W; ++j) w[j] = acc[j];
        return i0;
    }
#pragma unroll
    for (int j = 1; j <= W; ++j) {
        if (kind == NUFFT_KERNEL_GAUSSIAN) {
            const T y = ((T)(M - j) + X) * kp.dx[d];
            w[j - 1] = exp_t(-(y * y) / kp.tau[d]);
        } else {
            const T y = ((T)(M - j) + X) / (T)M;
            T z = (T)1 - y * y;
            z = z < (T)0 ? (T)0 : z;
            const T s = sqrt_t(z);
            const T beta = kp.beta[d];
            if (kind == NUFFT_KERNEL_KAISER_BESSEL) {
                w[j - 1] = i0_t(beta * s);
            } else {
                const T bs = beta * s;
                const T f = (s == (T)0) ? (T)1 : sinh_t(bs) / bs;
                w[j - 1] = f * (beta / (T)3.14159265358979323846);
            }
        }
    }
    return i0;
}

}  // namespace nufft
