/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C, compiled twice: REAL=float and REAL=double) of the
 * NonuniformFFTs.jl v0.9.6 CPU algorithm for the path
 *     set_points! -> exec_type1! / exec_type2!
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * Pinning: the reference is pure Julia and cannot run in this environment (no julia
 * binary); it ships no binary golden vectors.  The oracle is pinned against every
 * known-answer test the reference's own suite holds for this path (exact-NUDFT
 * error thresholds of test/accuracy.jl and test/multidimensional.jl, FFT
 * equivalence of test/uniform_points.jl, the cell-index edge cases of
 * test/near_2pi.jl, fast-vs-direct agreement of test/approx_window_functions.jl,
 * callback equivalence of test/callbacks.jl, the error case of test/errors.jl):
 * see tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line (relative to /root/reference) it
 * restates.  This file is included by nufft_oracle.c with
 *     #define REAL float  / double      and     #define SUF(x) x##_f32 / x##_f64
 */

#define ORC_MAXM 16            /* max half support handled by the oracle */
#define ORC_MAXW (2 * ORC_MAXM)
#define ORC_MAXP (ORC_MAXM + 4)

typedef struct {
    int kind;      /* 0 KB, 1 BKB, 2 Gaussian, 3 B-spline, 4 ES (not in the reference: parity unpinned, see orc_fit_func) */
    int M;         /* half support */
    int N;         /* oversampled grid size in this dimension */
    REAL beta;     /* KB / BKB shape parameter */
    REAL beta2;    /* beta*beta (kaiser_bessel.jl:112) */
    REAL w;        /* M * dx */
    REAL dx;       /* 2pi / N */
    REAL tau;      /* Gaussian: 2 sigma^2 */
    REAL gcs[ORC_MAXM];            /* Gaussian precomputed exponentials (gaussian.jl:80-83) */
    REAL cs[ORC_MAXP][ORC_MAXW];   /* piecewise polynomial coefficients cs[p][j] */
} SUF(orc_kernel);

/* src/Kernels/Kernels.jl:87 — domain period 2*T(pi) */
static inline REAL SUF(orc_period)(void) { return (REAL)2 * (REAL)M_PI; }

/* src/blocking/blocking.jl:12-21 — to_unit_cell_cpu (while loops, not fmod) */
REAL SUF(orc_fold)(REAL x)
{
    const REAL L = SUF(orc_period)();
    while (x < 0) x += L;
    while (x >= L) x -= L;
    return x;
}

/* src/abstractNFFTs.jl:150-158 — _transform_point_convention */
static inline REAL SUF(orc_nfft_convention)(REAL x)
{
    const REAL twopi = SUF(orc_period)();
    x = twopi * x;
    x = -x;
    return (x < 0) ? x + twopi : x;
}

/* src/plan.jl:459-464 — point_transform then fold */
static inline REAL SUF(orc_transform_fold)(REAL x, int convention)
{
    if (convention == 1) x = SUF(orc_nfft_convention)(x);
    return SUF(orc_fold)(x);
}

/* src/Kernels/Kernels.jl:121-126 — point_to_cell; returns 1-based cell, r through pointer.
 * The order (x / L) * N is normative (test/near_2pi.jl:19-46). */
int64_t SUF(orc_point_to_cell)(REAL x, int64_t N, REAL *r_out)
{
    const REAL L = SUF(orc_period)();
    volatile REAL q = x / L;   /* volatile: forbid reassociation/contraction */
    REAL r = q * (REAL)N;
    if (r_out) *r_out = r;
    return (int64_t)r + 1;
}

/* ---- modified Bessel I0 (Bessels.jl besseli0 is an un-vendored dependency; this is the
 * defining power series sum_k (x^2/4)^k / (k!)^2, all terms positive, evaluated in long double) */
#ifndef ORC_BESSELI0_DEFINED
#define ORC_BESSELI0_DEFINED
static double orc_besseli0(double x)
{
    long double q = (long double)x * (long double)x / 4.0L;
    long double term = 1.0L, sum = 1.0L;
    for (int k = 1; k < 2000; ++k) {
        term *= q / ((long double)k * (long double)k);
        sum += term;
        if (term < sum * 1e-22L) break;
    }
    return (double)sum;
}
#endif

/* src/Kernels/piecewise_polynomial.jl:23-41 — Vandermonde solve by partial-pivot LU in precision T */
static void SUF(orc_solve_vandermonde)(int n, const REAL *xs, REAL *ys /* in: samples, out: coefs */)
{
    REAL A[ORC_MAXP][ORC_MAXP];
    REAL xp[ORC_MAXP];
    int piv[ORC_MAXP];
    for (int i = 0; i < n; ++i) xp[i] = 1;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) { A[i][j] = xp[i]; xp[i] *= xs[i]; }
    /* LU, partial pivoting (LAPACK getrf semantics: first max |a_ik| in the column) */
    for (int k = 0; k < n; ++k) {
        int p = k; REAL amax = (REAL)fabs((double)A[k][k]);
        for (int i = k + 1; i < n; ++i) {
            REAL a = (REAL)fabs((double)A[i][k]);
            if (a > amax) { amax = a; p = i; }
        }
        piv[k] = p;
        if (p != k) for (int j = 0; j < n; ++j) { REAL t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
        REAL inv = (REAL)1 / A[k][k];
        for (int i = k + 1; i < n; ++i) A[i][k] *= inv;
        for (int j = k + 1; j < n; ++j) {
            REAL akj = A[k][j];
            for (int i = k + 1; i < n; ++i) A[i][j] -= A[i][k] * akj;
        }
    }
    for (int k = 0; k < n; ++k) if (piv[k] != k) { REAL t = ys[k]; ys[k] = ys[piv[k]]; ys[piv[k]] = t; }
    for (int i = 0; i < n; ++i) { REAL s = ys[i]; for (int j = 0; j < i; ++j) s -= A[i][j] * ys[j]; ys[i] = s; }
    for (int i = n - 1; i >= 0; --i) { REAL s = ys[i]; for (int j = i + 1; j < n; ++j) s -= A[i][j] * ys[j]; ys[i] = s / A[i][i]; }
}

/* kernel function on [-1,1] used for the fit.
 * KB  : kaiser_bessel.jl:127-131  besseli0(beta*sqrt(1-x^2))
 * BKB : kaiser_bessel_backwards.jl:98-102  sinh(beta*s)/(s*pi)
 * The fit samples are evaluated in Float64 (y = h + x*delta is Float64 in the reference
 * because h, delta are Float64 literals, piecewise_polynomial.jl:63-68) then rounded to T.
 * ES  : NOT in the reference (the north star of this project names it; PARITY UNPINNED, checked against exact NUDFT sums only).
 *       "Exponential of semicircle" of Barnett, Magland & af Klinteberg, SIAM J. Sci. Comput. 41 (2019):
 *       phi(y) = exp(beta (sqrt(1 - y^2) - 1)),  beta = 0.976 pi M (2 - 1/sigma)  (= 2.30 x 2M at sigma = 2, their rule);
 *       evaluated through the same piecewise polynomials as KB / BKB; its Fourier transform has no closed form and is
 *       computed by Gauss-Legendre quadrature (orc_es_fourier). */
static double SUF(orc_fit_func)(int kind, double beta, double y)
{
    double z = 1.0 - y * y;
    double s = sqrt(z < 0 ? 0 : z);
    if (kind == 0) return orc_besseli0(beta * s);
    if (kind == 4) return exp(beta * (s - 1.0));
    if (s == 0) return beta / M_PI;
    return sinh(beta * s) / (s * M_PI);
}

/* src/Kernels/piecewise_polynomial.jl:50-74 — solve_piecewise_polynomial_coefficients */
static void SUF(orc_fit_piecewise)(SUF(orc_kernel) *g)
{
    const int M = g->M, W = 2 * M, n = M + 4;
    REAL xs[ORC_MAXP], ys[ORC_MAXP];
    for (int i = 1; i <= n; ++i) {
        /* cospi(T(i - 1/2) / N) */
        REAL a = (REAL)(i - 0.5) / (REAL)n;
        xs[i - 1] = (REAL)cos(M_PI * (double)a);
    }
    for (int j = 1; j <= W; ++j) {
        double h = 1.0 - 2.0 * (j - 0.5) / W;
        double delta = 1.0 / W;
        for (int i = 0; i < n; ++i) {
            double y = h + (double)xs[i] * delta;
            ys[i] = (REAL)SUF(orc_fit_func)(g->kind, (double)g->beta, y);
        }
        SUF(orc_solve_vandermonde)(n, xs, ys);
        for (int p = 0; p < n; ++p) g->cs[p][j - 1] = ys[p];   /* transposed layout (:43-47) */
    }
}

/*
 * optimal_kernel + *KernelData constructors:
 *   KB   kaiser_bessel.jl:105-120,152-166     BKB  kaiser_bessel_backwards.jl:91-104,123-136
 *   Gauss gaussian.jl:75-86,106-115           B-spline bspline.jl:63-69,87-88
 * param = NaN selects the default shape rule; otherwise it is beta (KB/BKB) or ell/dx (Gaussian).
 */
int SUF(orc_kernel_init)(SUF(orc_kernel) *g, int kind, int M, int64_t N, REAL sigma, double param)
{
    if (M < 1 || M > ORC_MAXM || kind < 0 || kind > 4) return -1;
    memset(g, 0, sizeof(*g));
    g->kind = kind; g->M = M; g->N = (int)N;
    const REAL L = SUF(orc_period)();
    g->dx = L / (REAL)N;
    g->w = (REAL)M * g->dx;
    if (kind == 0 || kind == 1 || kind == 4) {
        if (isnan(param)) {
            REAL a = (REAL)M * ((REAL)2 - (REAL)1 / sigma);
            double a2 = (double)(a * a);
            double gamma = (kind == 0) ? sqrt(1.0 - 0.8 / a2) : (kind == 1) ? fmax(0.995, sqrt(1.0 - 0.3 / a2)) : 0.976;
            REAL pa = (REAL)M_PI * a;
            g->beta = (REAL)((double)pa * gamma);
        } else {
            g->beta = (REAL)param;
        }
        g->beta2 = g->beta * g->beta;
        SUF(orc_fit_piecewise)(g);
    } else if (kind == 2) {
        REAL ell;
        if (isnan(param)) ell = (REAL)sqrt((double)(sigma * (REAL)M / ((REAL)2 * sigma - (REAL)1)) / M_PI);
        else ell = (REAL)param;
        REAL sg = ell * g->dx;
        g->tau = (REAL)2 * sg * sg;
        for (int i = 1; i <= M; ++i) {
            REAL x = (REAL)i * g->dx;
            g->gcs[i - 1] = (REAL)exp((double)(-(x * x) / g->tau));
        }
    }
    return 0;
}

/* evaluate_fourier_func: KB kaiser_bessel.jl:168-175; BKB kaiser_bessel_backwards.jl:138-145;
 * Gaussian gaussian.jl:117-122; B-spline bspline.jl:121-129.  (Kernels.jl:108-117 applies it to ks.) */
/* Gauss-Legendre nodes and weights on [0, 1] (Newton iteration on P_n, double precision) */
#ifndef ORC_GAUSS_LEGENDRE_DEFINED
#define ORC_GAUSS_LEGENDRE_DEFINED
static void orc_gauss_legendre01(int n, double *x, double *w)
{
    for (int i = 0; i < n; ++i) {
        double t = cos(M_PI * (i + 0.75) / (n + 0.5)), dp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p0 = 1.0, p1 = t;
            for (int k = 2; k <= n; ++k) { const double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = p2; }
            dp = n * (t * p1 - p0) / (t * t - 1.0);
            const double dt = p1 / dp;
            t -= dt;
            if (fabs(dt) < 1e-16) break;
        }
        {   /* derivative at the converged node */
            double p0 = 1.0, p1 = t;
            for (int k = 2; k <= n; ++k) { const double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = p2; }
            dp = n * (t * p1 - p0) / (t * t - 1.0);
        }
        x[i] = 0.5 * (t + 1.0);
        w[i] = 1.0 / ((1.0 - t * t) * dp * dp);          /* = (2 / ((1 - t^2) P_n'(t)^2)) / 2 for the interval [0, 1] */
    }
}
#endif

/* ES: phihat(k) = int_{-w}^{w} phi(x / w) exp(-i k x) dx = 2 w int_0^1 phi(y) cos(k w y) dy; with y = sin(theta) the integrand
 * exp(beta (cos(theta) - 1)) cos(k w sin(theta)) cos(theta) is entire (no square-root end point): Gauss-Legendre on [0, pi/2]
 * with 24 + 8 M nodes converges to double precision */
static void SUF(orc_es_fourier)(const SUF(orc_kernel) *g, int64_t n, const REAL *ks, REAL *out)
{
    enum { QMAX = 24 + 8 * ORC_MAXM };
    double x[QMAX], wq[QMAX], f[QMAX], sn[QMAX];
    const int q = 24 + 8 * g->M;
    orc_gauss_legendre01(q, x, wq);
    for (int i = 0; i < q; ++i) {
        const double th = 0.5 * M_PI * x[i];
        sn[i] = sin(th);
        f[i] = 0.5 * M_PI * wq[i] * exp((double)g->beta * (cos(th) - 1.0)) * cos(th);
    }
    for (int64_t a = 0; a < n; ++a) {
        const double kw = (double)ks[a] * (double)g->w;
        double acc = 0.0;
        for (int i = 0; i < q; ++i) acc += f[i] * cos(kw * sn[i]);
        out[a] = (REAL)(2.0 * (double)g->w * acc);
    }
}

void SUF(orc_kernel_fourier)(const SUF(orc_kernel) *g, int64_t n, const REAL *ks, REAL *out)
{
    if (g->kind == 4) { SUF(orc_es_fourier)(g, n, ks, out); return; }
    for (int64_t a = 0; a < n; ++a) {
        REAL k = ks[a];
        if (g->kind == 0) {
            REAL q = g->w * k;
            REAL s = (REAL)sqrt((double)(g->beta2 - q * q));
            out[a] = (REAL)2 * g->w * (REAL)sinh((double)s) / s;
        } else if (g->kind == 1) {
            REAL q = g->w * k;
            REAL s = (REAL)sqrt((double)(g->beta * g->beta - q * q));
            out[a] = g->w * (REAL)orc_besseli0((double)s);
        } else if (g->kind == 2) {
            out[a] = (REAL)exp((double)(-g->tau * k * k / (REAL)4)) * (REAL)sqrt(M_PI * (double)g->tau);
        } else {
            REAL kh = k * g->dx / (REAL)2;
            REAL s = (REAL)sin((double)kh) / kh;
            REAL p = 1;
            for (int e = 0; e < 2 * g->M; ++e) p *= s;
            out[a] = ((k == 0) ? (REAL)1 : p) * g->dx;
        }
    }
}

/* src/Kernels/bspline.jl:143-193 — bsplines_evaluate_all (@generated branch) + evaluate_step */
static void SUF(orc_bsplines_all)(REAL x, int k, REAL *out)
{
    REAL bp[ORC_MAXW], bq[ORC_MAXW], ds[ORC_MAXW];
    bp[0] = 1;
    for (int q = 2; q <= k; ++q) {
        REAL alpha = (REAL)1 / (REAL)(q - 1);
        REAL xx = x;
        for (int j = 0; j < q - 1; ++j) { ds[j] = alpha * xx; xx += 1; }
        bq[0] = ds[0] * bp[0];
        for (int j = 2; j <= q - 1; ++j)
            bq[j - 1] = ((REAL)1 - ds[j - 2]) * bp[j - 2] + ds[j - 1] * bp[j - 1];
        bq[q - 1] = ((REAL)1 - ds[q - 2]) * bp[q - 2];
        for (int j = 0; j < q; ++j) bp[j] = bq[j];
    }
    for (int j = 0; j < k; ++j) out[j] = bp[j];
}

/*
 * evaluate_kernel(evalmode, g, x): returns the 1-based cell i and the 2M values.
 *   mode 0 = FastApproximation, 1 = Direct  (Kernels.jl:129-139)
 *   fast:   KB kaiser_bessel.jl:177-186, BKB kaiser_bessel_backwards.jl:147-156 (piecewise_polynomial.jl:76-92),
 *           Gaussian gaussian.jl:125-138,155-192, B-spline bspline.jl:99-111
 *   direct: KB :198-210, BKB :158-175, Gaussian :141-153, B-spline :113-119
 */
int64_t SUF(orc_kernel_eval)(const SUF(orc_kernel) *g, int mode, REAL x, REAL *vals)
{
    const int M = g->M, W = 2 * M;
    REAL r;
    int64_t i = SUF(orc_point_to_cell)(x, g->N, &r);
    REAL X = r - (REAL)(i - 1);
    if (g->kind == 3) {
        REAL xp = (REAL)i - r;
        SUF(orc_bsplines_all)(xp, W, vals);
        return i;
    }
    if (mode == 0) {
        if (g->kind == 0 || g->kind == 1 || g->kind == 4) {
            REAL xt = (REAL)2 * X - (REAL)1;
            const int n = M + 4;
            for (int j = 0; j < W; ++j) {
                REAL y = g->cs[n - 1][j];
                for (int p = n - 2; p >= 0; --p) y = xt * y + g->cs[p][j];
                vals[j] = y;
            }
        } else { /* fast Gaussian gridding */
            REAL Xp = x - (REAL)(i - 1) * g->dx;
            REAL a = (REAL)exp((double)(-(Xp * Xp) / g->tau));
            REAL b = (REAL)exp((double)((REAL)2 * Xp * g->dx / g->tau));
            REAL bpow = 1;
            vals[M - 1] = a;
            for (int m = 1; m <= M - 1; ++m) {
                bpow *= b;
                vals[M - m - 1] = a * g->gcs[m - 1] / bpow;
                vals[M + m - 1] = a * g->gcs[m - 1] * bpow;
            }
            vals[W - 1] = a * g->gcs[M - 1] * bpow * b;
        }
        return i;
    }
    for (int j = 1; j <= W; ++j) {
        if (g->kind == 2) {
            REAL y = ((REAL)(M - j) + X) * g->dx;
            vals[j - 1] = (REAL)exp((double)(-(y * y) / g->tau));
        } else {
            REAL y = ((REAL)(M - j) + X) / (REAL)M;
            REAL z = (REAL)1 - y * y;
            REAL s = (REAL)sqrt((double)(z < 0 ? 0 : z));
            if (g->kind == 0) vals[j - 1] = (REAL)orc_besseli0((double)(g->beta * s));
            else if (g->kind == 4) vals[j - 1] = (REAL)exp((double)(g->beta * (s - (REAL)1)));
            else {
                REAL bs = g->beta * s;
                REAL f = (s == 0) ? (REAL)1 : (REAL)sinh((double)bs) / bs;
                vals[j - 1] = f * (g->beta / (REAL)M_PI);
            }
        }
    }
    return i;
}

/* ------------------------------------------------------------------------------------------
 * set_points!: stable counting sort into blocks.
 * src/blocking/gpu.jl:145-160 (block_index), src/blocking/cpu.jl:73-111,113-185.
 * The reference's rank is an atomic counter (cpu.jl:88) so its intra-block order is only
 * deterministic with one thread; this restates the 1-thread (stable) order.
 * Outputs are 0-based: blockid[Np], cum[nblocks+1], perm[Np].
 * ------------------------------------------------------------------------------------------ */
int SUF(orc_set_points)(int D, int64_t Np, const REAL *const *xs, const int64_t *Ns,
                        const int64_t *block_dims, int convention,
                        int32_t *blockid, int32_t *cum, int32_t *perm)
{
    int64_t nb[3] = {1, 1, 1}, nblocks = 1;
    for (int d = 0; d < D; ++d) { nb[d] = (Ns[d] + block_dims[d] - 1) / block_dims[d]; nblocks *= nb[d]; }
    for (int64_t b = 0; b <= nblocks; ++b) cum[b] = 0;
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Np > 0 ? Np : 1));
    if (!rank) return -1;
    for (int64_t I = 0; I < Np; ++I) {
        int64_t n = 0, stride = 1;
        for (int d = 0; d < D; ++d) {
            REAL y = SUF(orc_transform_fold)(xs[d][I], convention);
            int64_t i = SUF(orc_point_to_cell)(y, Ns[d], NULL);     /* 1-based cell */
            int64_t b = (i + block_dims[d] - 1) / block_dims[d];    /* cld(i, B), 1-based */
            n += (b - 1) * stride;
            stride *= nb[d];
        }
        blockid[I] = (int32_t)n;
        rank[I] = cum[n + 1]++;             /* 0-based rank inside the block */
    }
    for (int64_t b = 1; b <= nblocks; ++b) cum[b] += cum[b - 1];
    for (int64_t I = 0; I < Np; ++I) perm[cum[blockid[I]] + rank[I]] = (int32_t)I;
    free(rank);
    return 0;
}

/* src/Kernels/Kernels.jl:146-158 — kernel_indices with periodic wrapping, 0-based output */
static inline void SUF(orc_wrapped_indices)(int64_t i /*1-based cell*/, int M, int64_t N, int64_t *idx)
{
    int64_t j = i - M;                   /* 1-based index before the first one */
    if (j < 0) j += N;
    for (int a = 0; a < 2 * M; ++a) { j = (j == N) ? 1 : j + 1; idx[a] = j - 1; }
}

/* ------------------------------------------------------------------------------------------
 * Type-1 spreading, serial non-blocked form:
 * src/spreading/cpu_nonblocked.jl:16-39,41-65,68-93.
 * us[c]: column-major oversampled grid, ncomp = 1 (real) or 2 (complex, interleaved) reals per cell.
 * vp[c]: values (same interleaving).  nu_weights (may be NULL): the non-uniform callback
 * v .* weights[n] of test/callbacks.jl:17.  `us` must be zeroed by the caller.
 * ------------------------------------------------------------------------------------------ */
void SUF(orc_spread)(int D, const int64_t *Ns, const SUF(orc_kernel) *gs, int mode,
                     int64_t Np, const REAL *const *xs, int convention,
                     int C, int ncomp, const REAL *const *vp, REAL *const *us,
                     const REAL *nu_weights)
{
    const int M = gs[0].M, W = 2 * M;
    for (int64_t n = 0; n < Np; ++n) {
        REAL vals[3][ORC_MAXW];
        int64_t idx[3][ORC_MAXW];
        int wd[3] = {1, 1, 1};
        for (int d = 0; d < 3; ++d) { vals[d][0] = 1; idx[d][0] = 0; }
        for (int d = 0; d < D; ++d) {
            REAL y = SUF(orc_transform_fold)(xs[d][n], convention);
            int64_t i = SUF(orc_kernel_eval)(&gs[d], mode, y, vals[d]);
            SUF(orc_wrapped_indices)(i, M, Ns[d], idx[d]);
            wd[d] = W;
        }
        const int64_t s1 = Ns[0], s2 = (D > 1) ? Ns[0] * Ns[1] : 0;
        for (int c = 0; c < C; ++c) {
            REAL vr = vp[c][ncomp * n], vi = (ncomp == 2) ? vp[c][2 * n + 1] : 0;
            if (nu_weights) { vr *= nu_weights[n]; vi *= nu_weights[n]; }
            REAL *u = us[c];
            for (int jz = 0; jz < wd[2]; ++jz)
                for (int jy = 0; jy < wd[1]; ++jy) {
                    REAL gt = vals[2][jz] * vals[1][jy];
                    if (D == 1) gt = 1;
                    else if (D == 2) gt = vals[1][jy];
                    int64_t base = idx[1][jy] * s1 + idx[2][jz] * s2;
                    for (int jx = 0; jx < W; ++jx) {
                        REAL gp = gt * vals[0][jx];
                        int64_t o = base + idx[0][jx];
                        if (ncomp == 2) { u[2 * o] += vr * gp; u[2 * o + 1] += vi * gp; }
                        else u[o] += vr * gp;
                    }
                }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Type-1 spreading, blocked + threaded form (the reference's default CPU path, used as the
 * timed CPU baseline):  src/spreading/cpu_blocked.jl:38-64 (spread into padded block buffer),
 * :94-168 (driver over blocks; per-thread buffers), :216-266 (add block to the periodic grid,
 * here with atomics = the `use_atomics=true` variant, :201-214), :16-36 (split_periodic).
 * Blocks are scheduled dynamically over OpenMP threads (the reference balances blocks over
 * threads by point count, src/blocking/cpu.jl:187-220).
 * ------------------------------------------------------------------------------------------ */
void SUF(orc_spread_blocked)(int D, const int64_t *Ns, const SUF(orc_kernel) *gs, int mode,
                             int64_t Np, const REAL *const *xs, int convention,
                             int C, int ncomp, const REAL *const *vp, REAL *const *us,
                             const REAL *nu_weights,
                             const int64_t *block_dims, const int32_t *cum, const int32_t *perm)
{
    const int M = gs[0].M, W = 2 * M;
    int64_t nb[3] = {1, 1, 1}, B[3] = {1, 1, 1}, Nd[3] = {1, 1, 1}, pad[3] = {1, 1, 1}, nblocks = 1;
    for (int d = 0; d < D; ++d) {
        B[d] = block_dims[d]; Nd[d] = Ns[d];
        nb[d] = (Ns[d] + B[d] - 1) / B[d]; nblocks *= nb[d];
        pad[d] = B[d] + 2 * M;                       /* cpu.jl:54: padding 2M per dimension */
    }
    const int64_t bufcells = pad[0] * pad[1] * pad[2];
    (void)Np;
    /* Merge strategy: the reference merges block buffers under a lock (:156-163) or with atomics (:201-214).
     * Here blocks are processed in "colours": two blocks whose coordinates differ by >= ncol[d] blocks in some
     * dimension have disjoint padded regions, so blocks of one colour are merged concurrently with plain adds
     * (same sums; no lock contention, no CAS loops).  ncol[d] = number of blocks a padded buffer can overlap. */
    int64_t ncol[3] = {1, 1, 1};
    for (int d = 0; d < D; ++d) {
        const int64_t need = (pad[d] + B[d] - 1) / B[d] + 1;   /* +1: the last block may stick out past N */
        ncol[d] = need;
        if (ncol[d] > nb[d]) ncol[d] = nb[d];
        /* periodic wrap: colours must also separate the first and last blocks */
        while (ncol[d] < nb[d] && (nb[d] % ncol[d]) != 0 && (nb[d] % ncol[d]) < need) ++ncol[d];
    }
    const int64_t ncolors = ncol[0] * ncol[1] * ncol[2];
#pragma omp parallel
    {
        REAL *buf = (REAL *)malloc(sizeof(REAL) * (size_t)(bufcells * ncomp));
        for (int64_t color = 0; color < ncolors; ++color) {
            const int64_t c0 = color % ncol[0], c1 = (color / ncol[0]) % ncol[1], c2 = color / (ncol[0] * ncol[1]);
            const int64_t m0 = (nb[0] - c0 + ncol[0] - 1) / ncol[0], m1 = (nb[1] - c1 + ncol[1] - 1) / ncol[1],
                          m2 = (nb[2] - c2 + ncol[2] - 1) / ncol[2];
            const int64_t nblk_color = m0 * m1 * m2;
#pragma omp for schedule(dynamic, 1)
        for (int64_t jj = 0; jj < nblk_color; ++jj) {
            const int64_t b0 = c0 + (jj % m0) * ncol[0], b1 = c1 + ((jj / m0) % m1) * ncol[1], b2 = c2 + (jj / (m0 * m1)) * ncol[2];
            const int64_t j = (b2 * nb[1] + b1) * nb[0] + b0;
            const int32_t a = cum[j], b = cum[j + 1];
            if (a == b) continue;
            int64_t I0[3];                           /* 0-based first cell of the block */
            { int64_t t = j; for (int d = 0; d < 3; ++d) { I0[d] = (t % nb[d]) * B[d]; t /= nb[d]; } }
            for (int c = 0; c < C; ++c) {
                memset(buf, 0, sizeof(REAL) * (size_t)(bufcells * ncomp));
                for (int32_t k = a; k < b; ++k) {
                    const int64_t l = perm[k];
                    REAL vals[3][ORC_MAXW]; int64_t st[3] = {0, 0, 0}; int wd[3] = {1, 1, 1};
                    vals[1][0] = 1; vals[2][0] = 1;
                    for (int d = 0; d < D; ++d) {
                        REAL y = SUF(orc_transform_fold)(xs[d][l], convention);
                        int64_t i = SUF(orc_kernel_eval)(&gs[d], mode, y, vals[d]);   /* 1-based */
                        st[d] = (i - 1) - I0[d] + 1;   /* cpu_blocked.jl:49-56; buffer index 0 is unused, as in the reference */
                        wd[d] = W;
                    }
                    REAL vr = vp[c][ncomp * l], vi = (ncomp == 2) ? vp[c][2 * l + 1] : 0;
                    if (nu_weights) { vr *= nu_weights[l]; vi *= nu_weights[l]; }
                    for (int jz = 0; jz < wd[2]; ++jz)
                        for (int jy = 0; jy < wd[1]; ++jy) {
                            REAL gt = (D == 1) ? (REAL)1 : (D == 2 ? vals[1][jy] : vals[2][jz] * vals[1][jy]);
                            REAL gr = vr * gt, gi = vi * gt;
                            int64_t o = ((st[2] + jz) * pad[1] + (st[1] + jy)) * pad[0] + st[0];
                            if (ncomp == 2) {
                                REAL *p = buf + 2 * o;
                                for (int jx = 0; jx < W; ++jx) { p[2 * jx] += gr * vals[0][jx]; p[2 * jx + 1] += gi * vals[0][jx]; }
                            } else {
                                REAL *p = buf + o;
                                for (int jx = 0; jx < W; ++jx) p[jx] += gr * vals[0][jx];
                            }
                        }
                }
                /* add_from_block!: buffer cell q (0-based) <-> global cell I0 - M + q (periodic) */
                REAL *u = us[c];
                for (int64_t qz = 0; qz < pad[2]; ++qz) {
                    int64_t gz = (D > 2) ? ((I0[2] - M + qz) % Nd[2] + Nd[2]) % Nd[2] : 0;
                    for (int64_t qy = 0; qy < pad[1]; ++qy) {
                        int64_t gy = (D > 1) ? ((I0[1] - M + qy) % Nd[1] + Nd[1]) % Nd[1] : 0;
                        const REAL *src = buf + ncomp * ((qz * pad[1] + qy) * pad[0]);
                        REAL *dst = u + ncomp * ((gz * Nd[1] + gy) * Nd[0]);
                        int64_t gx = ((I0[0] - M) % Nd[0] + Nd[0]) % Nd[0];
                        for (int64_t qx = 0; qx < pad[0]; ++qx) {
                            for (int e = 0; e < ncomp; ++e) dst[ncomp * gx + e] += src[ncomp * qx + e];
                            if (++gx == Nd[0]) gx = 0;
                        }
                    }
                }
            }
        }   /* omp for: implicit barrier between colours */
        }
        free(buf);
    }
}

/* ------------------------------------------------------------------------------------------
 * Type-2 interpolation, non-blocked form (threaded over points; per-point arithmetic as
 * src/interpolation/cpu_nonblocked.jl:1-79: values scaled by dx per dimension (:44-47), then
 * sum_j prod_d g_d[j_d] * u[idx]).  nu_weights: non-uniform callback applied to the result.
 * ------------------------------------------------------------------------------------------ */
void SUF(orc_interp)(int D, const int64_t *Ns, const SUF(orc_kernel) *gs, int mode,
                     int64_t Np, const REAL *const *xs, int convention,
                     int C, int ncomp, REAL *const *vp, const REAL *const *us,
                     const REAL *nu_weights)
{
    const int M = gs[0].M, W = 2 * M;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < Np; ++n) {
        REAL vals[3][ORC_MAXW];
        int64_t idx[3][ORC_MAXW];
        int wd[3] = {1, 1, 1};
        for (int d = 0; d < 3; ++d) { vals[d][0] = 1; idx[d][0] = 0; }
        for (int d = 0; d < D; ++d) {
            REAL y = SUF(orc_transform_fold)(xs[d][n], convention);
            int64_t i = SUF(orc_kernel_eval)(&gs[d], mode, y, vals[d]);
            SUF(orc_wrapped_indices)(i, M, Ns[d], idx[d]);
            for (int j = 0; j < W; ++j) vals[d][j] *= gs[d].dx;
            wd[d] = W;
        }
        const int64_t s1 = Ns[0], s2 = (D > 1) ? Ns[0] * Ns[1] : 0;
        for (int c = 0; c < C; ++c) {
            const REAL *u = us[c];
            REAL ar = 0, ai = 0;
            for (int jz = 0; jz < wd[2]; ++jz)
                for (int jy = 0; jy < wd[1]; ++jy) {
                    REAL gt = (D == 1) ? (REAL)1 : (D == 2 ? vals[1][jy] : vals[2][jz] * vals[1][jy]);
                    int64_t base = idx[1][jy] * s1 + idx[2][jz] * s2;
                    for (int jx = 0; jx < W; ++jx) {
                        REAL gp = gt * vals[0][jx];
                        int64_t o = base + idx[0][jx];
                        if (ncomp == 2) { ar += gp * u[2 * o]; ai += gp * u[2 * o + 1]; }
                        else ar += gp * u[o];
                    }
                }
            if (nu_weights) { ar *= nu_weights[n]; ai *= nu_weights[n]; }
            if (ncomp == 2) { vp[c][2 * n] = ar; vp[c][2 * n + 1] = ai; }
            else vp[c][n] = ar;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Type-2 interpolation, blocked + threaded form (the reference's default CPU path; timed CPU baseline):
 * src/interpolation/cpu_blocked.jl:95-153 (driver over blocks), :156-206 (copy_to_block!),
 * :1-35 (interpolate_blocked; values scaled by dx :29-32), :38-93 (dot products).
 * ------------------------------------------------------------------------------------------ */
void SUF(orc_interp_blocked)(int D, const int64_t *Ns, const SUF(orc_kernel) *gs, int mode,
                             int64_t Np, const REAL *const *xs, int convention,
                             int C, int ncomp, REAL *const *vp, const REAL *const *us,
                             const REAL *nu_weights,
                             const int64_t *block_dims, const int32_t *cum, const int32_t *perm)
{
    const int M = gs[0].M, W = 2 * M;
    int64_t nb[3] = {1, 1, 1}, B[3] = {1, 1, 1}, Nd[3] = {1, 1, 1}, pad[3] = {1, 1, 1}, nblocks = 1;
    for (int d = 0; d < D; ++d) {
        B[d] = block_dims[d]; Nd[d] = Ns[d];
        nb[d] = (Ns[d] + B[d] - 1) / B[d]; nblocks *= nb[d];
        pad[d] = B[d] + 2 * M;
    }
    const int64_t bufcells = pad[0] * pad[1] * pad[2];
    (void)Np;
#pragma omp parallel
    {
        REAL *buf = (REAL *)malloc(sizeof(REAL) * (size_t)(bufcells * ncomp));
#pragma omp for schedule(dynamic, 1)
        for (int64_t j = 0; j < nblocks; ++j) {
            const int32_t a = cum[j], b = cum[j + 1];
            if (a == b) continue;
            int64_t I0[3];
            { int64_t t = j; for (int d = 0; d < 3; ++d) { I0[d] = (t % nb[d]) * B[d]; t /= nb[d]; } }
            for (int c = 0; c < C; ++c) {
                const REAL *u = us[c];
                for (int64_t qz = 0; qz < pad[2]; ++qz) {
                    int64_t gz = (D > 2) ? ((I0[2] - M + qz) % Nd[2] + Nd[2]) % Nd[2] : 0;
                    for (int64_t qy = 0; qy < pad[1]; ++qy) {
                        int64_t gy = (D > 1) ? ((I0[1] - M + qy) % Nd[1] + Nd[1]) % Nd[1] : 0;
                        REAL *dst = buf + ncomp * ((qz * pad[1] + qy) * pad[0]);
                        const REAL *src = u + ncomp * ((gz * Nd[1] + gy) * Nd[0]);
                        int64_t gx = ((I0[0] - M) % Nd[0] + Nd[0]) % Nd[0];
                        for (int64_t qx = 0; qx < pad[0]; ++qx) {
                            for (int e = 0; e < ncomp; ++e) dst[ncomp * qx + e] = src[ncomp * gx + e];
                            if (++gx == Nd[0]) gx = 0;
                        }
                    }
                }
                for (int32_t k = a; k < b; ++k) {
                    const int64_t l = perm[k];
                    REAL vals[3][ORC_MAXW]; int64_t st[3] = {0, 0, 0}; int wd[3] = {1, 1, 1};
                    vals[1][0] = 1; vals[2][0] = 1;
                    for (int d = 0; d < D; ++d) {
                        REAL y = SUF(orc_transform_fold)(xs[d][l], convention);
                        int64_t i = SUF(orc_kernel_eval)(&gs[d], mode, y, vals[d]);
                        for (int q = 0; q < W; ++q) vals[d][q] *= gs[d].dx;
                        st[d] = (i - 1) - I0[d] + 1;
                        wd[d] = W;
                    }
                    REAL ar = 0, ai = 0;
                    for (int jz = 0; jz < wd[2]; ++jz)
                        for (int jy = 0; jy < wd[1]; ++jy) {
                            REAL gt = (D == 1) ? (REAL)1 : (D == 2 ? vals[1][jy] : vals[2][jz] * vals[1][jy]);
                            int64_t o = ((st[2] + jz) * pad[1] + (st[1] + jy)) * pad[0] + st[0];
                            if (ncomp == 2) {
                                const REAL *p = buf + 2 * o;
                                REAL sr = 0, si = 0;
                                for (int jx = 0; jx < W; ++jx) { sr += vals[0][jx] * p[2 * jx]; si += vals[0][jx] * p[2 * jx + 1]; }
                                ar += gt * sr; ai += gt * si;
                            } else {
                                const REAL *p = buf + o;
                                REAL sr = 0;
                                for (int jx = 0; jx < W; ++jx) sr += vals[0][jx] * p[jx];
                                ar += gt * sr;
                            }
                        }
                    if (nu_weights) { ar *= nu_weights[l]; ai *= nu_weights[l]; }
                    if (ncomp == 2) { vp[c][2 * l] = ar; vp[c][2 * l + 1] = ai; }
                    else vp[c][l] = ar;
                }
            }
        }
        free(buf);
    }
}

/* ------------------------------------------------------------------------------------------
 * Deconvolution.  index_map is 0-based here (reference: 1-based, NonuniformFFTs.jl:318-348).
 * type 1: src/NonuniformFFTs.jl:350-385   w[I] = cb( normfactor / prod(phihat_d[I_d]) * uhat[map(I)] )
 * type 2: src/NonuniformFFTs.jl:416-451   uhat[map(I)] = cb( w[I] / prod(phihat_d[I_d]) )   (after zero fill, :260-266)
 * Uniform callback: multiply by dense factor[I] (covers the 1/k^2 callback of test/callbacks.jl:18-23).
 * All arrays complex interleaved, column-major.  nk = size(plan), nos = spectral oversampled dims.
 * ------------------------------------------------------------------------------------------ */
void SUF(orc_deconv_type1)(int D, const int64_t *nk, const int64_t *nos, const int64_t *const *imap,
                           const REAL *const *phihat, REAL normfactor,
                           int C, REAL *const *wout, const REAL *const *uhat, const REAL *u_factor)
{
    int64_t n1 = nk[0], n2 = (D > 1) ? nk[1] : 1, n3 = (D > 2) ? nk[2] : 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t i3 = 0; i3 < n3; ++i3)
        for (int64_t i2 = 0; i2 < n2; ++i2) {
            int64_t j3 = (D > 2) ? imap[2][i3] : 0, j2 = (D > 1) ? imap[1][i2] : 0;
            for (int64_t i1 = 0; i1 < n1; ++i1) {
                REAL ph = phihat[0][i1];
                if (D > 1) ph *= phihat[1][i2];
                /* prod(phihat_front) * phihat_last, :374 */
                if (D > 2) ph = (phihat[0][i1] * phihat[1][i2]) * phihat[2][i3];
                REAL beta = normfactor / ph;
                int64_t I = (i3 * n2 + i2) * n1 + i1;
                int64_t J = (j3 * ((D > 1) ? nos[1] : 1) + j2) * nos[0] + imap[0][i1];
                REAL f = u_factor ? u_factor[I] : (REAL)1;
                for (int c = 0; c < C; ++c) {
                    REAL re = beta * uhat[c][2 * J], im = beta * uhat[c][2 * J + 1];
                    if (u_factor) { re *= f; im *= f; }
                    wout[c][2 * I] = re; wout[c][2 * I + 1] = im;
                }
            }
        }
}

void SUF(orc_deconv_type2)(int D, const int64_t *nk, const int64_t *nos, const int64_t *const *imap,
                           const REAL *const *phihat,
                           int C, REAL *const *uhat /* zeroed here */, const REAL *const *win, const REAL *u_factor)
{
    int64_t n1 = nk[0], n2 = (D > 1) ? nk[1] : 1, n3 = (D > 2) ? nk[2] : 1;
    int64_t tot = nos[0] * ((D > 1) ? nos[1] : 1) * ((D > 2) ? nos[2] : 1);
    for (int c = 0; c < C; ++c) memset(uhat[c], 0, sizeof(REAL) * 2 * (size_t)tot);
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t i3 = 0; i3 < n3; ++i3)
        for (int64_t i2 = 0; i2 < n2; ++i2) {
            int64_t j3 = (D > 2) ? imap[2][i3] : 0, j2 = (D > 1) ? imap[1][i2] : 0;
            for (int64_t i1 = 0; i1 < n1; ++i1) {
                REAL ph = phihat[0][i1];
                if (D > 1) ph *= phihat[1][i2];
                if (D > 2) ph = (phihat[0][i1] * phihat[1][i2]) * phihat[2][i3];
                REAL beta = (REAL)1 / ph;
                int64_t I = (i3 * n2 + i2) * n1 + i1;
                int64_t J = (j3 * ((D > 1) ? nos[1] : 1) + j2) * nos[0] + imap[0][i1];
                REAL f = u_factor ? u_factor[I] : (REAL)1;
                for (int c = 0; c < C; ++c) {
                    REAL re = beta * win[c][2 * I], im = beta * win[c][2 * I + 1];
                    if (u_factor) { re *= f; im *= f; }
                    uhat[c][2 * J] = re; uhat[c][2 * J + 1] = im;
                }
            }
        }
}

size_t SUF(orc_kernel_sizeof)(void) { return sizeof(SUF(orc_kernel)); }
void SUF(orc_kernel_get)(const SUF(orc_kernel) *g, REAL *beta, REAL *w, REAL *dx, REAL *tau, REAL *cs /* (M+4)*2M, [p][j] */, REAL *gcs)
{
    *beta = g->beta; *w = g->w; *dx = g->dx; *tau = g->tau;
    for (int p = 0; p < g->M + 4; ++p) for (int j = 0; j < 2 * g->M; ++j) cs[p * 2 * g->M + j] = g->cs[p][j];
    for (int m = 0; m < g->M; ++m) gcs[m] = g->gcs[m];
}
