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

// I0 by its power series, I0(z) = sum_k (z^2 / 4)^k / (k!)^2 (all terms positive: relative accuracy ~ nterms * eps everywhere),
// for W arguments in lockstep.  The library cyl_bessel_i0 is a call with one long dependent chain per argument: with the
// two producer warps of the tile kernels (and one lane per point elsewhere) Float64 Direct evaluation was latency-bound and
// cost more than the spreading itself; W independent Horner chains keep the FP64 pipe busy instead.  nterms is chosen on the
// host from the plan's beta (make_kernel_params); 0 = beta too large for the table, use the library function.
constexpr int I0_TABLE = 96;
static __constant__ double c_i0_coef[I0_TABLE] = {
    1.0, 1.0, 0.25, 0.027777777777777776,
    0.001736111111111111, 6.944444444444444e-05, 1.9290123456790124e-06, 3.936759889140842e-08,
    6.151187326782565e-10, 7.594058428126624e-12, 7.594058428126623e-14, 6.276081345559193e-16,
    4.358389823304995e-18, 2.5789288895295828e-20, 1.3157800456783586e-22, 5.8479113141260385e-25,
    2.2843403570804838e-27, 7.904291893012054e-30, 2.4395962632753253e-32, 6.757884385804225e-35,
    1.6894710964510564e-37, 3.8310002187098785e-40, 7.915289708078262e-43, 1.4962740468957016e-45,
    2.5976979980828152e-48, 4.156316796932504e-51, 6.14839762859838e-54, 8.434015951438106e-57,
    1.0757673407446564e-59, 1.2791526049282477e-62, 1.4212806721424974e-65, 1.4789601166935458e-68,
    1.4442969889585408e-71, 1.3262598613026087e-74, 1.147283617043779e-77, 9.365580547296156e-81,
    7.226528200074194e-84, 5.278691161485898e-87, 3.6556032974279072e-90, 2.4034209713529963e-93,
    1.5021381070956227e-96, 8.935979221270807e-100, 5.065747857863269e-103, 2.739723016691871e-106,
    1.4151461863077845e-109, 6.988376228680417e-113, 3.3026352687525603e-116, 1.4950816064973112e-119,
    6.48906947264458e-123, 2.7026528415845812e-126, 1.0810611366338324e-129, 4.156328860568368e-133,
    1.5371038685533903e-136, 5.472067883778535e-140, 1.876566489635986e-143, 6.203525585573507e-147,
    1.9781650464201234e-150, 6.088535076700903e-154, 1.8099093569265468e-157, 5.199394877697635e-161,
    1.4442763549160097e-164, 3.8814199272131406e-168, 1.0097346324695995e-171, 2.544052991860921e-175,
    6.211066874660451e-179, 1.4700749999196336e-182, 3.374827823506964e-186, 7.517994705963386e-190,
    1.6258639069990022e-193, 3.414963047676963e-197, 6.969312342197885e-201, 1.3825257572302885e-204,
    2.6669092539164517e-208, 5.004521024425692e-212, 9.139008444897175e-216, 1.6247126124261642e-219,
    2.812868096305686e-223, 4.744253830841096e-227, 7.797918854110941e-231, 1.2494662480549497e-234,
    1.952291012585859e-238, 2.975599775317572e-242, 4.425341724148679e-246, 6.423779538610364e-250,
    9.103995944742579e-254, 1.2600686428709451e-257, 1.7037163911180977e-261, 2.2509134510742472e-265,
    2.9066547663665382e-269, 3.6695553167106908e-273, 4.530315205815668e-277, 5.470734459383731e-281,
    6.463533151445807e-285, 7.473156609371958e-289, 8.457624048632819e-293, 9.371328585742735e-297};
template <int W> __device__ __forceinline__ void i0_series(const double (&z)[W], double (&out)[W], int nterms)
{
    double q[W], acc[W];
#pragma unroll
    for (int j = 0; j < W; ++j) { q[j] = 0.25 * z[j] * z[j]; acc[j] = c_i0_coef[nterms - 1]; }
    for (int k = nterms - 2; k >= 0; --k) {
        const double c = c_i0_coef[k];
#pragma unroll
        for (int j = 0; j < W; ++j) acc[j] = fma(acc[j], q[j], c);
    }
#pragma unroll
    for (int j = 0; j < W; ++j) out[j] = acc[j];
}

__device__ __forceinline__ double fmod_t(double x, double y) { return fmod(x, y); }

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
    if constexpr (sizeof(T) == 8) {
        if (kind == NUFFT_KERNEL_KAISER_BESSEL && kp.i0_terms > 0) {       // Float64 KB: batched power series (see i0_series)
            double zz[W];
#pragma unroll
            for (int j = 1; j <= W; ++j) {
                const T y = ((T)(M - j) + X) / (T)M;
                T z = (T)1 - y * y;
                z = z < (T)0 ? (T)0 : z;
                zz[j - 1] = kp.beta[d] * sqrt_t(z);
            }
            double ww[W];
            i0_series<W>(zz, ww, kp.i0_terms);
#pragma unroll
            for (int j = 0; j < W; ++j) w[j] = (T)ww[j];
            return i0;
        }
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
