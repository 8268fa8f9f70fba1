// host_plan.cu — plan-time host logic: everything `_PlanNUFFT` does (reference src/plan.jl:467-541)
// except the Julia type machinery.  Runs once per plan on the host, uploads small tables.
//
//   oversampled sizes            src/plan.jl:485-498  (nextprod((2,3,5), floor(sigma*N)))
//   check_nufft_size             src/plan.jl:545-556
//   kernel shape rules           src/Kernels/kaiser_bessel.jl:152-166, kaiser_bessel_backwards.jl:123-136,
//                                gaussian.jl:106-115, bspline.jl:87-88
//   piecewise-polynomial fit     src/Kernels/piecewise_polynomial.jl:23-74  (LU in precision T)
//   Fourier coefficients         src/Kernels/Kernels.jl:108-117 + evaluate_fourier_func of each kernel
//   wavenumbers                  src/plan.jl:558-566
//   index maps                   src/NonuniformFFTs.jl:318-348
//   FFT plans                    src/plan.jl:37-60   (cuFFT here; FFTW / cuFFT.jl there)
//   tile geometry                replaces src/gpu_common.jl:19-92 (48 KiB static smem) with a
//                                227 KiB dynamic shared-memory budget on sm_100a
#include <cmath>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "common.cuh"

namespace nufft {

static int64_t nextprod235(int64_t n)
{
    if (n < 1) n = 1;
    int64_t best = -1;
    for (int64_t p5 = 1; p5 < 5 * n; p5 *= 5)
        for (int64_t p35 = p5; p35 < 3 * n; p35 *= 3) {
            int64_t v = p35;
            while (v < n) v *= 2;
            if (best < 0 || v < best) best = v;
        }
    return best;
}

// Solve the monomial Vandermonde system V c = y (nodes xs) by LU with partial pivoting in precision T.
template <typename T> static void solve_vandermonde(int n, const T *xs, T *ys)
{
    std::vector<T> A((size_t)n * n), xp(n, (T)1);
    std::vector<int> piv(n);
    auto a = [&](int i, int j) -> T & { return A[(size_t)j * n + i]; };   // column-major
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) { a(i, j) = xp[i]; xp[i] *= xs[i]; }
    for (int k = 0; k < n; ++k) {
        int p = k;
        T amax = std::fabs(a(k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(a(i, k)) > amax) { amax = std::fabs(a(i, k)); p = i; }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < n; ++j) std::swap(a(k, j), a(p, j));
        const T inv = (T)1 / a(k, k);
        for (int i = k + 1; i < n; ++i) a(i, k) *= inv;
        for (int j = k + 1; j < n; ++j) {
            const T akj = a(k, j);
            for (int i = k + 1; i < n; ++i) a(i, j) -= a(i, k) * akj;
        }
    }
    for (int k = 0; k < n; ++k)
        if (piv[k] != k) std::swap(ys[k], ys[piv[k]]);
    for (int i = 0; i < n; ++i) {
        T s = ys[i];
        for (int j = 0; j < i; ++j) s -= a(i, j) * ys[j];
        ys[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
        T s = ys[i];
        for (int j = i + 1; j < n; ++j) s -= a(i, j) * ys[j];
        ys[i] = s / a(i, i);
    }
}

// (NUFFT_KERNEL_ES is not in the reference: "exponential of semicircle" phi(y) = exp(beta (sqrt(1 - y^2) - 1)) of Barnett, Magland &
// af Klinteberg, SIAM J. Sci. Comput. 41 (2019), beta = 0.976 pi M (2 - 1/sigma) = 2.30 x 2M at sigma = 2; evaluated through the
// same piecewise polynomials as KB / BKB, Fourier transform by Gauss-Legendre quadrature.  Parity unpinned: see include/nufft_b200.h)
static double kernel_fit_func(int kind, double beta, double y)
{
    double z = 1.0 - y * y;
    double s = std::sqrt(z < 0 ? 0.0 : z);
    if (kind == NUFFT_KERNEL_KAISER_BESSEL) return std::cyl_bessel_i(0.0, beta * s);
    if (kind == NUFFT_KERNEL_ES) return std::exp(beta * (s - 1.0));
    if (s == 0.0) return beta / M_PI;
    return std::sinh(beta * s) / (s * M_PI);
}

// Gauss-Legendre nodes and weights on [0, 1]: Newton iteration on the Legendre polynomial P_n
static void gauss_legendre_unit(int n, std::vector<double> &x, std::vector<double> &w)
{
    x.assign(n, 0.0); w.assign(n, 0.0);
    auto legendre = [n](double t, double &pn, double &pn1) {
        double a = 1.0, b = t;
        for (int k = 2; k <= n; ++k) { const double c = ((2.0 * k - 1.0) * t * b - (k - 1.0) * a) / k; a = b; b = c; }
        pn = b; pn1 = a;
    };
    for (int i = 0; i < n; ++i) {
        double t = std::cos(M_PI * (i + 0.75) / (n + 0.5)), pn, pn1;
        for (int it = 0; it < 100; ++it) {
            legendre(t, pn, pn1);
            const double dt = pn / (n * (t * pn - pn1) / (t * t - 1.0));
            t -= dt;
            if (std::fabs(dt) < 1e-16) break;
        }
        legendre(t, pn, pn1);
        const double dp = n * (t * pn - pn1) / (t * t - 1.0);
        x[i] = 0.5 * (t + 1.0);
        w[i] = 1.0 / ((1.0 - t * t) * dp * dp);
    }
}

// Kernel data of one dimension in precision T: shape parameter, cs block, phihat table.
template <typename T>
static void build_kernel_dim(Plan &p, int d, std::vector<T> &cs_block, std::vector<T> &phihat)
{
    const int M = p.M, W = 2 * M, np = M + 4;
    const int kind = p.opts.kernel;
    const int64_t Nos = p.Nos[d];
    const T L = (T)2 * (T)M_PI;
    const T dx = L / (T)Nos;
    const T w = (T)M * dx;
    const T sigma_d = (T)((T)Nos / (T)p.Ns[d]);            // src/plan.jl:503
    const bool user_param = !std::isnan(p.opts.kernel_param);
    T beta = 0, tau = 0;
    cs_block.assign((size_t)p.cs_stride, (T)0);

    if (kind == NUFFT_KERNEL_KAISER_BESSEL || kind == NUFFT_KERNEL_BACKWARDS_KAISER_BESSEL || kind == NUFFT_KERNEL_ES) {
        if (user_param) beta = (T)p.opts.kernel_param;
        else {
            const T a = (T)M * ((T)2 - (T)1 / sigma_d);
            const double a2 = (double)(a * a);
            const double gamma = (kind == NUFFT_KERNEL_KAISER_BESSEL) ? std::sqrt(1.0 - 0.8 / a2)
                                 : (kind == NUFFT_KERNEL_BACKWARDS_KAISER_BESSEL) ? std::max(0.995, std::sqrt(1.0 - 0.3 / a2)) : 0.976;
            const T pa = (T)M_PI * a;
            beta = (T)((double)pa * gamma);
        }
        // Chebyshev nodes cospi((i - 1/2)/np); samples in double, rounded to T; LU in T
        std::vector<T> xs(np), ys(np);
        for (int i = 1; i <= np; ++i) {
            const T arg = (T)(i - 0.5) / (T)np;
            xs[i - 1] = (T)std::cos(M_PI * (double)arg);
        }
        for (int j = 1; j <= W; ++j) {
            const double h = 1.0 - 2.0 * (j - 0.5) / W, delta = 1.0 / W;
            for (int i = 0; i < np; ++i) ys[i] = (T)kernel_fit_func(kind, (double)beta, h + (double)xs[i] * delta);
            solve_vandermonde<T>(np, xs.data(), ys.data());
            for (int q = 0; q < np; ++q) cs_block[(size_t)q * W + (j - 1)] = ys[q];
        }
        p.h_shape[d] = (double)beta;
        if (kind == NUFFT_KERNEL_KAISER_BESSEL) {
            // I0(beta sqrt(t)) = sum_k c_k t^k, c_k = (beta^2 / 4)^k / (k!)^2 (Direct evaluation, kernel_eval.cuh)
            const double q = 0.25 * (double)beta * (double)beta;
            double term = 1.0;
            for (int k = 0; k < I0_MAX_TERMS; ++k) {
                if (k > 0) term *= q / ((double)k * (double)k);
                cs_block[(size_t)np * W + M + k] = (T)term;
            }
        }
    } else if (kind == NUFFT_KERNEL_GAUSSIAN) {
        T ell = user_param ? (T)p.opts.kernel_param
                           : (T)std::sqrt((double)(sigma_d * (T)M / ((T)2 * sigma_d - (T)1)) / M_PI);
        const T sg = ell * dx;
        tau = (T)2 * sg * sg;
        for (int i = 1; i <= M; ++i) {
            const T x = (T)i * dx;
            cs_block[(size_t)np * W + (i - 1)] = (T)std::exp((double)(-(x * x) / tau));
        }
        p.h_shape[d] = (double)tau;
    } else {
        p.h_shape[d] = 0.0;
    }
    p.kp_beta[d] = (double)beta;
    p.kp_tau[d] = (double)tau;
    p.h_dx[d] = (double)dx;

    // wavenumbers (integers) in output order, then phihat(k)
    const int64_t nk = p.nk[d];
    const bool r2c = (!p.cplx) && d == 0;
    phihat.resize((size_t)nk);
    // ES: phihat(k) = 2 w int_0^1 phi(y) cos(k w y) dy; with y = sin(theta) the integrand exp(beta (cos(theta) - 1)) cos(k w sin(theta))
    // cos(theta) is entire (no square-root end point): 24 + 8 M Gauss-Legendre nodes on [0, pi/2] reach double precision
    std::vector<double> qx, qw;
    if (kind == NUFFT_KERNEL_ES) {
        gauss_legendre_unit(24 + 8 * M, qx, qw);
        for (size_t i = 0; i < qx.size(); ++i) {
            const double th = 0.5 * M_PI * qx[i];
            qw[i] *= 0.5 * M_PI * std::exp((double)beta * (std::cos(th) - 1.0)) * std::cos(th);
            qx[i] = std::sin(th);
        }
    }
    for (int64_t a = 0; a < nk; ++a) {
        int64_t ki;
        if (r2c) ki = a;
        else {
            const int64_t N = p.Ns[d];
            int64_t idx = a;
            if (p.opts.fftshift) idx = (a + (N + 1) / 2) % N;   // fftshift(fftfreq): position a holds original index (a + ceil(N/2)) mod N
            ki = (idx < (N + 1) / 2) ? idx : idx - N;
        }
        const T k = (T)ki;
        T val;
        if (kind == NUFFT_KERNEL_ES) {
            const double kw = (double)k * (double)w;
            double acc = 0.0;
            for (size_t i = 0; i < qx.size(); ++i) acc += qw[i] * std::cos(kw * qx[i]);
            val = (T)(2.0 * (double)w * acc);
        } else if (kind == NUFFT_KERNEL_KAISER_BESSEL) {
            const T q = w * k;
            const T s = (T)std::sqrt((double)(beta * beta - q * q));
            val = (T)2 * w * (T)std::sinh((double)s) / s;
        } else if (kind == NUFFT_KERNEL_BACKWARDS_KAISER_BESSEL) {
            const T q = w * k;
            const T s = (T)std::sqrt((double)(beta * beta - q * q));
            val = w * (T)std::cyl_bessel_i(0.0, (double)s);
        } else if (kind == NUFFT_KERNEL_GAUSSIAN) {
            val = (T)std::exp((double)(-tau * k * k / (T)4)) * (T)std::sqrt(M_PI * (double)tau);
        } else {
            const T kh = k * dx / (T)2;
            const T s = (T)std::sin((double)kh) / kh;
            T pw = 1;
            for (int e = 0; e < W; ++e) pw *= s;
            val = ((ki == 0) ? (T)1 : pw) * dx;
        }
        phihat[(size_t)a] = val;
    }
}

// kept-mode -> oversampled spectral index (0-based), src/NonuniformFFTs.jl:318-348
static std::vector<int32_t> build_index_map(int64_t Nk, int64_t Nos, bool r2c, bool fftshift)
{
    std::vector<int32_t> m((size_t)Nk);
    if (r2c) {
        for (int64_t i = 0; i < Nk; ++i) m[i] = (int32_t)i;
    } else if (Nk % 2 == 0) {
        const int64_t h = Nk / 2;
        for (int64_t i = 0; i < h; ++i) {
            if (fftshift) { m[i] = (int32_t)(Nos - h + i); m[h + i] = (int32_t)i; }
            else { m[i] = (int32_t)i; m[h + i] = (int32_t)(Nos - h + i); }
        }
    } else {
        const int64_t h = (Nk - 1) / 2;
        if (fftshift) {
            for (int64_t i = 0; i < h; ++i) m[i] = (int32_t)(Nos - h + i);
            for (int64_t i = 0; i <= h; ++i) m[h + i] = (int32_t)i;
        } else {
            for (int64_t i = 0; i <= h; ++i) m[i] = (int32_t)i;
            for (int64_t i = 0; i < h; ++i) m[h + 1 + i] = (int32_t)(Nos - h + i);
        }
    }
    return m;
}

// ---- tile geometry ---------------------------------------------------------------------------
// Shared-memory budget per CTA.  sm_100a: 228 KiB per SM, 227 KiB max per CTA, 1 KiB reserved per CTA.
static constexpr int SMEM_PER_SM = 228 * 1024;
static constexpr int SMEM_MAX_CTA = 227 * 1024;

// per-CTA shared memory besides the tile: double-buffered batch (values, int4 starts, weight records) + tables
size_t sm_batch_bytes(int D, int M, size_t real_bytes, size_t cell_bytes, int batch)
{
    return (size_t)2 * batch * (cell_bytes + 16 + (size_t)record_size(D, M) * real_bytes);
}

// Shared-memory wavefronts needed by one warp-wide tile access of the spreading / interpolation kernels:
// lane -> (lx = lane % W, lg = lane / W), cell = base + lg * Sx + lx; lanes >= (32 / W) * W idle.
// 4-byte cells are served 32 lanes per wavefront, 8-byte cells 16 lanes, 16-byte cells 8 lanes.
static int tile_access_wavefronts(int Sx, int W, int cell_bytes)
{
    const int G = 32 / W, lanes_per_phase = 128 / cell_bytes, groups = 128 / cell_bytes;
    int worst = 0;
    for (int base = 0; base < groups; ++base) {
        int total = 0;
        for (int ph = 0; ph * lanes_per_phase < 32; ++ph) {
            int cnt[32] = {0};
            int mx = 0;
            for (int l = ph * lanes_per_phase; l < (ph + 1) * lanes_per_phase; ++l) {
                const int lx = l % W, lg = l / W;
                if (lg >= G) continue;
                const int c = (base + lg * Sx + lx) % groups;
                mx = std::max(mx, ++cnt[c]);
            }
            total += mx;
        }
        worst = std::max(worst, total);
    }
    return worst;
}

// row stride (cells) of the shared-memory tile: smallest padding that minimises bank conflicts
static int padded_row(int Tx, int W, int cell_bytes)
{
    int best = Tx, best_wf = 1 << 30;
    for (int s = Tx; s <= Tx + 32 && s <= Tx + Tx / 4 + 1; ++s) {
        const int wf = tile_access_wavefronts(s, W, cell_bytes);
        if (wf < best_wf) { best_wf = wf; best = s; }
    }
    return best;
}

static bool choose_geometry(Plan &p, bool allow_cs = true)
{
    TileGeom &g = p.geom;
    const int D = p.D, M = p.M, W = 2 * M;
    const size_t cell_bytes = p.real_bytes * (p.cplx ? 2 : 1);
    g.D = D;
    g.batch = 64;
    if (const char *e = getenv("NUFFT_B200_BATCH")) {          // tuning knob (multiple of 32, >= 32)
        const int v = atoi(e);
        if (v >= 32 && v <= 1024) g.batch = v / 32 * 32;
    }
    g.chunk = p.opts.spread_chunk > 0 ? p.opts.spread_chunk : 4096;
    for (int d = 0; d < 3; ++d) { g.N[d] = (int)p.Nos[d]; g.B[d] = 1; g.nb[d] = 1; g.T[d] = 1; g.S[d] = 1; }

    auto tile_bytes = [&](const int *B, int *T, int *S) -> size_t {
        for (int d = 0; d < 3; ++d) T[d] = (d < D) ? B[d] + W - 1 : 1;
        S[0] = padded_row(T[0], W, (int)cell_bytes);
        S[1] = S[0];
        S[2] = S[0] * T[1];
        return (size_t)S[0] * T[1] * T[2] * cell_bytes;
    };
    const size_t fixed = sm_batch_bytes(D, M, p.real_bytes, cell_bytes, g.batch) + ((size_t)D * p.cs_stride + 4) * p.real_bytes + 128;

    // bins never exceed N - (2M - 1) cells so that the padded tile fits in one period (T_d <= N_d): the kernels
    // then wrap coordinates with a single conditional add/subtract (N >= 2M is guaranteed by check_nufft_size)
    auto bcap = [&](int d) -> int64_t { return std::max<int64_t>(1, p.Nos[d] - (W - 1)); };
    bool user = false;
    for (int d = 0; d < D; ++d) if (p.opts.block_dims[d] > 0) user = true;
    int B[3] = {1, 1, 1}, T[3], S[3];
    bool ok = false;
    g.rt = 0;
    g.sub[0] = g.sub[1] = g.sub[2] = 1;
    g.nsub = 1;

    // ---- column-streaming fast path (cs_spread.cuh / cs_interp.cuh): D = 3, M = 4, Float32 data -----------------------
    // bins = columns of 4 x 4 cells in (x, y), segments of up to 256 cells in z; the sort key is refined by the z cell
    // inside the segment, so the points of a column arrive bottom to top, cell by cell
    {
        bool cs = allow_cs && D == 3 && M == 4 && !p.f64 && p.opts.gpu_method != NUFFT_METHOD_GLOBAL_MEMORY && !user;
        if (const char *e = getenv("NUFFT_B200_CS")) cs = cs && atoi(e) != 0;
        for (int d = 0; d < D && cs; ++d)
            if (p.Nos[d] < 16 || p.Nos[d] > 65536 * 4) cs = false;
        if (p.ncells >= ((int64_t)1 << 31)) cs = false;        // 32-bit cell offsets
        if (cs) {
            g.rt = 3;
            // (a z slab bins its owned planes only)
            const int64_t nzown = p.slab_nz > 0 ? p.slab_nz : p.Nos[2];
            int64_t seg = 256;                 // planes per z segment (tuning knob: NUFFT_B200_SEG)
            if (const char *e = getenv("NUFFT_B200_SEG")) { const int v = atoi(e); if (v >= 8 && v <= 4096) seg = v; }
            int Bc[3] = {4, 4, (int)std::max<int64_t>(4, std::min<int64_t>(seg, p.slab_nz > 0 ? nzown : bcap(2)) / 4 * 4)};
            tile_bytes(Bc, T, S);               // strides of the generic shared-memory kernels (unused on this path)
            int64_t nbins = 1;
            for (int d = 0; d < 3; ++d) {
                g.B[d] = Bc[d]; g.T[d] = T[d]; g.S[d] = S[d];
                g.sub[d] = d < 2 ? 1 : Bc[2];          // the sort key is refined by the z CELL inside the segment
                g.nb[d] = (int)cdiv(d == 2 ? nzown : p.Nos[d], Bc[d]);
                nbins *= g.nb[d];
            }
            g.nsub = g.sub[2];
            g.tile_cells = S[0] * T[1] * T[2];
            p.nbins = nbins;
            p.key_bits = 1;
            while (((int64_t)1 << p.key_bits) < nbins * g.nsub) ++p.key_bits;
            return true;
        }
    }

    if (user) {
        for (int d = 0; d < D; ++d) {
            int64_t b = p.opts.block_dims[d] > 0 ? p.opts.block_dims[d] : 16;
            B[d] = (int)std::min<int64_t>(b, bcap(d));
        }
        ok = tile_bytes(B, T, S) + fixed <= (size_t)SMEM_MAX_CTA;
    } else {
        // cubic bins; prefer >= 2 resident CTAs per SM, fall back to 1
        const size_t budgets[2] = {(size_t)(SMEM_PER_SM / 2 - 1024), (size_t)SMEM_MAX_CTA};
        const int bmax = (D == 1) ? 4096 : (D == 2 ? 64 : 32);
        for (int pass = 0; pass < 2 && !ok; ++pass) {
            int best = 0;
            for (int b = 1; b <= bmax; ++b) {
                for (int d = 0; d < D; ++d) B[d] = (int)std::min<int64_t>(b, bcap(d));
                if (tile_bytes(B, T, S) + fixed <= budgets[pass]) best = b;
            }
            const int bmin = (D == 3) ? std::max(4, M) : 8;
            if (best >= bmin || (pass == 1 && best >= 2)) {
                for (int d = 0; d < D; ++d) B[d] = (int)std::min<int64_t>(best, bcap(d));
                ok = true;
            }
        }
    }
    if (!ok) {
        // shared-memory tiles unusable: bins only serve locality of the global-memory kernels
        for (int d = 0; d < D; ++d) B[d] = (int)std::min<int64_t>(D == 1 ? 1024 : (D == 2 ? 32 : 8), bcap(d));
    }
    tile_bytes(B, T, S);
    int64_t nbins = 1;
    for (int d = 0; d < 3; ++d) {
        g.B[d] = B[d]; g.T[d] = T[d]; g.S[d] = S[d];
        g.nb[d] = (d < D) ? (int)cdiv(p.Nos[d], B[d]) : 1;
        nbins *= g.nb[d];
    }
    g.tile_cells = S[0] * T[1] * T[2];
    p.nbins = nbins;
    p.key_bits = 1;
    while (((int64_t)1 << p.key_bits) < nbins) ++p.key_bits;
    return ok;
}

template <typename T> static int upload_tables(Plan &p)
{
    std::vector<T> all_cs((size_t)3 * p.cs_stride, (T)0);
    for (int d = 0; d < p.D; ++d) {
        std::vector<T> cs, ph;
        build_kernel_dim<T>(p, d, cs, ph);
        std::copy(cs.begin(), cs.end(), all_cs.begin() + (size_t)d * p.cs_stride);
        p.h_cs[d].assign((const char *)cs.data(), cs.size() * sizeof(T));
        p.h_phihat[d].assign((const char *)ph.data(), ph.size() * sizeof(T));
        CUDA_TRY(cudaMalloc(&p.d_phihat[d], ph.size() * sizeof(T)));
        CUDA_TRY(cudaMemcpyAsync(p.d_phihat[d], ph.data(), ph.size() * sizeof(T), cudaMemcpyHostToDevice, p.stream));
    }
    CUDA_TRY(cudaMalloc(&p.d_cs, all_cs.size() * sizeof(T)));
    CUDA_TRY(cudaMemcpyAsync(p.d_cs, all_cs.data(), all_cs.size() * sizeof(T), cudaMemcpyHostToDevice, p.stream));
    CUDA_TRY(cudaStreamSynchronize(p.stream));
    return NUFFT_SUCCESS;
}

// validation of the options and every size that follows from them (no device is touched)
static int plan_sizes(Plan &p)
{
    const nufft_opts &o = p.opts;
    if (o.dim < 1 || o.dim > 3) { set_error("dim must be 1, 2 or 3 (got %d)", o.dim); return NUFFT_ERR_ARG; }
    if (o.dtype != NUFFT_F32 && o.dtype != NUFFT_F64) { set_error("dtype must be NUFFT_F32 or NUFFT_F64"); return NUFFT_ERR_ARG; }
    if (o.kernel < 0 || o.kernel > NUFFT_KERNEL_ES) { set_error("unknown kernel %d", o.kernel); return NUFFT_ERR_ARG; }
    if (o.eval_mode != NUFFT_EVAL_FAST && o.eval_mode != NUFFT_EVAL_DIRECT) { set_error("unknown eval_mode %d", o.eval_mode); return NUFFT_ERR_ARG; }
    if (o.kernel == NUFFT_KERNEL_ES && o.eval_mode != NUFFT_EVAL_FAST) {
        set_error("the ES kernel is evaluated through its piecewise polynomials only: use FastApproximation");
        return NUFFT_ERR_UNSUPPORTED;
    }
    if (o.ntransforms < 1) { set_error("ntransforms must be >= 1"); return NUFFT_ERR_ARG; }
    if (o.gpu_method < 0 || o.gpu_method > 2) { set_error("expected gpu_method in (auto, global_memory, shared_memory)"); return NUFFT_ERR_ARG; }
    if (o.half_support < MIN_M || o.half_support > MAX_M) {
        set_error("HalfSupport(%d) is not instantiated in this build (supported: %d..%d)", o.half_support, MIN_M, MAX_M);
        return NUFFT_ERR_UNSUPPORTED;
    }
    if (!(o.sigma >= 1.0)) { set_error("oversampling factor sigma must be >= 1 (got %g)", o.sigma); return NUFFT_ERR_ARG; }
    if (o.fftshift && !o.is_complex) { set_error("fftshift = true requires complex non-uniform data"); return NUFFT_ERR_ARG; }
    p.D = o.dim; p.M = o.half_support; p.C = o.ntransforms;
    p.cplx = o.is_complex != 0; p.f64 = o.dtype == NUFFT_F64;
    p.real_bytes = p.f64 ? 8 : 4;

    // oversampled sizes (sigma converted to T first, src/plan.jl:575-576)
    for (int d = 0; d < p.D; ++d) {
        const int64_t N = o.n_modes[d];
        if (N < 1) { set_error("n_modes[%d] must be >= 1", d); return NUFFT_ERR_ARG; }
        p.Ns[d] = N;
        int64_t base;
        const bool half = (!p.cplx) && d == 0;
        const int64_t Nin = half ? (N + 1) / 2 : N;
        if (p.f64) base = (int64_t)std::floor(o.sigma * (double)Nin);
        else base = (int64_t)std::floor((double)((float)o.sigma * (float)Nin));
        int64_t Nt = nextprod235(base);
        if (half) Nt *= 2;
        if (Nt < 2 * p.M) {
            set_error("data size is too small: sigma*N = %lld < %d = 2M. Try either: 1. increasing the number of data points N "
                      "2. increasing the oversampling factor sigma 3. decreasing the kernel half-support M", (long long)Nt, 2 * p.M);
            return NUFFT_ERR_ARG;
        }
        if (Nt > (int64_t)1 << 30) { set_error("oversampled dimension too large"); return NUFFT_ERR_ARG; }
        p.Nos[d] = Nt;
        p.Nspec[d] = half ? Nt / 2 + 1 : Nt;
        p.nk[d] = half ? N / 2 + 1 : N;
    }
    p.nz_local = (int)p.Nos[2];
    if (p.slab_nz > 0) {
        if (p.D != 3 || p.slab_z0 < 0 || p.slab_z0 + p.slab_nz > p.Nos[2] || p.slab_nz < p.M) {
            set_error("invalid z slab [%d, %d) of %lld planes", p.slab_z0, p.slab_z0 + p.slab_nz, (long long)p.Nos[2]);
            return NUFFT_ERR_ARG;
        }
        p.nz_local = p.slab_nz + 2 * p.M - 1;
    }
    p.ncells = p.Nos[0] * p.Nos[1] * (int64_t)p.nz_local;
    p.nspec = p.Nspec[0] * p.Nspec[1] * p.Nspec[2];
    p.nkept = p.nk[0] * p.nk[1] * p.nk[2];
    if (p.ncells >= ((int64_t)1 << 31)) { set_error("oversampled grid has >= 2^31 cells (32-bit cell indices)"); return NUFFT_ERR_UNSUPPORTED; }
    p.cs_stride = (p.M + 4) * 2 * p.M + p.M + I0_MAX_TERMS;
    return NUFFT_SUCCESS;
}

// host-only: the kernel data of dimension d that a plan with these options would hold (nufft_kernel_tables)
int host_kernel_tables(const nufft_opts &o, int d, double *shape, double *dx, int64_t *os_dim, double *cs, size_t cs_len, double *phihat,
                       size_t phihat_len)
{
    Plan p;
    p.opts = o;
    NUFFT_TRY(plan_sizes(p));
    if (d < 0 || d >= p.D) { set_error("dimension %d out of range", d); return NUFFT_ERR_ARG; }
    const size_t ncs = (size_t)(p.M + 4) * 2 * p.M;
    if ((cs && cs_len < ncs) || (phihat && phihat_len < (size_t)p.nk[d])) { set_error("output buffer too small"); return NUFFT_ERR_ARG; }
    if (p.f64) {
        std::vector<double> c, ph;
        build_kernel_dim<double>(p, d, c, ph);
        if (cs) std::copy(c.begin(), c.begin() + ncs, cs);
        if (phihat) std::copy(ph.begin(), ph.end(), phihat);
    } else {
        std::vector<float> c, ph;
        build_kernel_dim<float>(p, d, c, ph);
        if (cs) for (size_t i = 0; i < ncs; ++i) cs[i] = (double)c[i];
        if (phihat) for (size_t i = 0; i < ph.size(); ++i) phihat[i] = (double)ph[i];
    }
    if (shape) *shape = p.h_shape[d];
    if (dx) *dx = p.h_dx[d];
    if (os_dim) *os_dim = p.Nos[d];
    return NUFFT_SUCCESS;
}

int host_plan_init(Plan &p)
{
    NUFFT_TRY(plan_sizes(p));
    const nufft_opts &o = p.opts;
    p.stream = (cudaStream_t)o.stream;
    if (o.device >= 0) { CUDA_TRY(cudaSetDevice(o.device)); }
    CUDA_TRY(cudaGetDevice(&p.device));
    CUDA_TRY(cudaDeviceGetAttribute(&p.num_sms, cudaDevAttrMultiProcessorCount, p.device));

    NUFFT_TRY(p.f64 ? upload_tables<double>(p) : upload_tables<float>(p));

    for (int d = 0; d < p.D; ++d) {
        const bool r2c = (!p.cplx) && d == 0;
        std::vector<int32_t> m = build_index_map(p.nk[d], p.Nspec[d], r2c, o.fftshift != 0);
        std::vector<int32_t> inv((size_t)p.Nspec[d], -1);
        for (size_t i = 0; i < m.size(); ++i) inv[(size_t)m[i]] = (int32_t)i;
        CUDA_TRY(cudaMalloc(&p.d_imap[d], m.size() * sizeof(int32_t)));
        CUDA_TRY(cudaMalloc(&p.d_invmap[d], inv.size() * sizeof(int32_t)));
        CUDA_TRY(cudaMemcpy(p.d_imap[d], m.data(), m.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p.d_invmap[d], inv.data(), inv.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }

    // bins / tiles and method
    bool sm_ok = choose_geometry(p);
    if (p.slab_nz > 0 && p.geom.rt != 3) {
        set_error("z-slab plans need the column-streaming kernels (3-D, HalfSupport(4), Float32)");
        return NUFFT_ERR_UNSUPPORTED;
    }
    if (p.geom.rt == 3 && p.slab_nz > 0) {
        sm_ok = true;
    } else if (p.geom.rt == 3) {
        // keep the tile geometry too: set_points picks by density (NUFFT_B200_CS_DENSITY = cells per point below which the
        // column-streaming kernels are used; 0 = always)
        p.geom_alt[0] = p.geom; p.nbins_alt[0] = p.nbins; p.key_bits_alt[0] = p.key_bits;
        const bool tile_ok = choose_geometry(p, false);
        p.geom_alt[1] = p.geom; p.nbins_alt[1] = p.nbins; p.key_bits_alt[1] = p.key_bits;
        p.dual_geom = tile_ok;
        if (const char *e = getenv("NUFFT_B200_CS_DENSITY")) p.cs_min_cells = atof(e);
        p.geom = p.geom_alt[0]; p.nbins = p.nbins_alt[0]; p.key_bits = p.key_bits_alt[0];
        sm_ok = true;
    }
    if (o.gpu_method == NUFFT_METHOD_SHARED_MEMORY && !sm_ok) {
        set_error("GPU shared memory is too small for the chosen problem (element bytes %zu, M = %d, D = %d); "
                  "reduce some of these parameters or switch to gpu_method = global_memory",
                  p.real_bytes * (p.cplx ? 2 : 1), p.M, p.D);
        return NUFFT_ERR_ARG;
    }
    if (o.gpu_method == NUFFT_METHOD_AUTO) p.method = (sm_ok && p.D >= 2) ? NUFFT_METHOD_SHARED_MEMORY : NUFFT_METHOD_GLOBAL_MEMORY;
    else p.method = o.gpu_method;

    // grids
    const size_t zbytes = p.real_bytes * (p.cplx ? 2 : 1);
    if (cudaMalloc(&p.d_us, (size_t)p.C * p.ncells * zbytes) != cudaSuccess) {
        cudaGetLastError();
        set_error("cannot allocate the oversampled grid (%zu bytes)", (size_t)p.C * p.ncells * zbytes);
        return NUFFT_ERR_ALLOC;
    }
    if (p.cplx) p.d_uhat = p.d_us;
    else if (cudaMalloc(&p.d_uhat, (size_t)p.C * p.nspec * 2 * p.real_bytes) != cudaSuccess) {
        cudaGetLastError();
        set_error("cannot allocate the oversampled spectrum");
        return NUFFT_ERR_ALLOC;
    }

    // pruned FFT fused with deconvolution (pfft.cu) when the plan is eligible; cuFFT + K-deconv otherwise
    p.pfft = pfft_eligible(p);
    if (p.slab_nz > 0 && !p.pfft) { set_error("z-slab plans need the pruned FFT (complex data, power-of-two oversampled sizes)"); return NUFFT_ERR_UNSUPPORTED; }
    if (p.pfft) NUFFT_TRY(pfft_init(p));

    // cuFFT plans: Julia dims are column-major -> reversed for cuFFT; batch = ntransforms
    int n[3];
    for (int d = 0; d < p.D; ++d) n[d] = (int)p.Nos[p.D - 1 - d];
    size_t ws_fw = 0, ws_bw = 0;
    if (!p.pfft) {
    CUFFT_TRY(cufftCreate(&p.fft_fw));
    CUFFT_TRY(cufftCreate(&p.fft_bw));
    p.fft_ok = true;
    CUFFT_TRY(cufftSetAutoAllocation(p.fft_fw, 0));
    CUFFT_TRY(cufftSetAutoAllocation(p.fft_bw, 0));
    if (p.cplx) {
        const cufftType t = p.f64 ? CUFFT_Z2Z : CUFFT_C2C;
        CUFFT_TRY(cufftMakePlanMany(p.fft_fw, p.D, n, nullptr, 1, 0, nullptr, 1, 0, t, p.C, &ws_fw));
        CUFFT_TRY(cufftMakePlanMany(p.fft_bw, p.D, n, nullptr, 1, 0, nullptr, 1, 0, t, p.C, &ws_bw));
    } else {
        CUFFT_TRY(cufftMakePlanMany(p.fft_fw, p.D, n, nullptr, 1, 0, nullptr, 1, 0, p.f64 ? CUFFT_D2Z : CUFFT_R2C, p.C, &ws_fw));
        CUFFT_TRY(cufftMakePlanMany(p.fft_bw, p.D, n, nullptr, 1, 0, nullptr, 1, 0, p.f64 ? CUFFT_Z2D : CUFFT_C2R, p.C, &ws_bw));
    }
    p.fft_work_bytes = std::max(ws_fw, ws_bw);
    if (p.fft_work_bytes > 0) {
        if (cudaMalloc(&p.d_fft_work, p.fft_work_bytes) != cudaSuccess) {
            cudaGetLastError();
            set_error("cannot allocate the cuFFT work area (%zu bytes)", p.fft_work_bytes);
            return NUFFT_ERR_ALLOC;
        }
        CUFFT_TRY(cufftSetWorkArea(p.fft_fw, p.d_fft_work));
        CUFFT_TRY(cufftSetWorkArea(p.fft_bw, p.d_fft_work));
    }
    CUFFT_TRY(cufftSetStream(p.fft_fw, p.stream));
    CUFFT_TRY(cufftSetStream(p.fft_bw, p.stream));
    }

    // binning tables that depend only on the plan
    const int64_t nbins_max = p.dual_geom ? std::max(p.nbins_alt[0], p.nbins_alt[1]) : p.nbins;
    CUDA_TRY(cudaMalloc(&p.d_bin_offsets, (size_t)(nbins_max + 1) * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&p.d_item_start, (size_t)(nbins_max + 1) * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&p.d_counters, 64 * sizeof(int32_t)));
    CUDA_TRY(cudaMemset(p.d_counters, 0, 64 * sizeof(int32_t)));
    if (o.record_timings) {
        for (int i = 0; i < 32; ++i) CUDA_TRY(cudaEventCreate(&p.ev[i]));
        p.ev_ok = true;
    }
    return NUFFT_SUCCESS;
}

void host_plan_free(Plan &p)
{
    jit_callbacks_free(p);
    if (p.fft_ok) { cufftDestroy(p.fft_fw); cufftDestroy(p.fft_bw); p.fft_ok = false; }
    pfft_free(p);
    auto f = [](auto *&ptr) { if (ptr) { cudaFree((void *)ptr); ptr = nullptr; } };
    f(p.d_fft_work);
    if (!p.cplx) f(p.d_uhat);
    p.d_uhat = nullptr;
    f(p.d_us); f(p.d_cs);
    for (int d = 0; d < 3; ++d) { f(p.d_phihat[d]); f(p.d_imap[d]); f(p.d_invmap[d]); f(p.d_xs[d]); }
    f(p.d_keys[0]); f(p.d_keys[1]); f(p.d_vals[0]); f(p.d_vals[1]); f(p.d_rec);
    f(p.d_perm_coarse);
    f(p.d_os);
    f(p.d_bin_offsets); f(p.d_hist); f(p.d_scan_tmp); f(p.d_item_start); f(p.d_item_table); f(p.d_counters);
    if (p.ev_ok) { for (int i = 0; i < 32; ++i) cudaEventDestroy(p.ev[i]); p.ev_ok = false; }
}

int fft_forward(Plan &p)
{
    NUFFT_COUNT_LAUNCH();
    if (p.cplx) {
        if (p.f64) CUFFT_TRY(cufftExecZ2Z(p.fft_fw, (cufftDoubleComplex *)p.d_us, (cufftDoubleComplex *)p.d_us, CUFFT_FORWARD));
        else CUFFT_TRY(cufftExecC2C(p.fft_fw, (cufftComplex *)p.d_us, (cufftComplex *)p.d_us, CUFFT_FORWARD));
    } else {
        if (p.f64) CUFFT_TRY(cufftExecD2Z(p.fft_fw, (cufftDoubleReal *)p.d_us, (cufftDoubleComplex *)p.d_uhat));
        else CUFFT_TRY(cufftExecR2C(p.fft_fw, (cufftReal *)p.d_us, (cufftComplex *)p.d_uhat));
    }
    return NUFFT_SUCCESS;
}

int fft_backward(Plan &p)
{
    NUFFT_COUNT_LAUNCH();
    if (p.cplx) {
        if (p.f64) CUFFT_TRY(cufftExecZ2Z(p.fft_bw, (cufftDoubleComplex *)p.d_us, (cufftDoubleComplex *)p.d_us, CUFFT_INVERSE));
        else CUFFT_TRY(cufftExecC2C(p.fft_bw, (cufftComplex *)p.d_us, (cufftComplex *)p.d_us, CUFFT_INVERSE));
    } else {
        // C2R overwrites its input: uhat is plan-owned scratch (same as ext/NonuniformFFTsCUDAExt.jl:53-64)
        if (p.f64) CUFFT_TRY(cufftExecZ2D(p.fft_bw, (cufftDoubleComplex *)p.d_uhat, (cufftDoubleReal *)p.d_us));
        else CUFFT_TRY(cufftExecC2R(p.fft_bw, (cufftComplex *)p.d_uhat, (cufftReal *)p.d_us));
    }
    return NUFFT_SUCCESS;
}

}  // namespace nufft
