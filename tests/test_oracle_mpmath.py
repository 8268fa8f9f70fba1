"""
CPU tests (no GPU): pin the ORACLE below the reference's NUDFT thresholds.

The reference cannot run here (no Julia), and its own tests bound errors at the 1e-7 class for HalfSupport(4).  A wrong
but accurate table would pass those.  Here the reference's FORMULAE (paths relative to /root/reference) are restated
independently in 50-digit arithmetic (mpmath) and the oracle's plan data must agree to a few ulp of its working precision T:

  shape rules      beta (KB: src/Kernels/kaiser_bessel.jl:152-166, BKB: kaiser_bessel_backwards.jl:123-136),
                   ell / tau (Gaussian: gaussian.jl:106-115)
  phihat tables    evaluate_fourier_func (KB :168-175, BKB :138-145, Gaussian :117-122, B-spline bspline.jl:121-129)
  fast evaluation  the piecewise polynomials = interpolants of the kernel at the Chebyshev nodes cospi((i - 1/2) / (M + 4)) of
                   the 2M sub-intervals, right to left (piecewise_polynomial.jl:23-92): the oracle's Horner values must equal the
                   EXACT interpolant (Lagrange form, no linear solve) up to the conditioning of the reference's Vandermonde solve
                   in T; fast Gaussian gridding (gaussian.jl:125-192); de Boor recursion (bspline.jl:143-193) against the
                   closed-form cardinal B-spline
  direct evaluation  KB :198-210, BKB :158-175, Gaussian :141-153 from the definitions

The ES kernel ("es": exponential of semicircle, NOT in the reference — parity unpinned) goes through the same checks against its
own definition: beta = 0.976 pi M (2 - 1/sigma), phi(y) = exp(beta (sqrt(1 - y^2) - 1)), phihat by 50-digit quadrature.
"""
import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
from oracle import OraclePlan  # noqa: E402

mp.mp.dps = 50
TWO_PI = 2 * mp.pi


def ulp(T, x):
    return float(np.spacing(np.abs(T(x))))


def shape_params(kernel, M, sigma_T):
    """(beta or None, ell_over_dx or None) by the reference's shape rules, real arithmetic."""
    s = mp.mpf(float(sigma_T))
    a = M * (2 - 1 / s)
    if kernel == "kaiser_bessel":
        return mp.pi * a * mp.sqrt(1 - mp.mpf("0.8") / a ** 2), None
    if kernel == "backwards_kaiser_bessel":
        return mp.pi * a * max(mp.mpf("0.995"), mp.sqrt(1 - mp.mpf("0.3") / a ** 2)), None
    if kernel == "es":
        return mp.mpf("0.976") * mp.pi * a, None
    if kernel == "gaussian":
        return None, mp.sqrt(s * M / (2 * s - 1) / mp.pi)
    return None, None


def kernel_func(kernel, beta, tau, M, dx):
    """phi(y), y in [-1, 1] the offset normalised by the half width w = M dx."""
    if kernel == "kaiser_bessel":
        return lambda y: mp.besseli(0, beta * mp.sqrt(max(mp.mpf(0), 1 - y * y)))
    if kernel == "backwards_kaiser_bessel":
        def f(y):
            s = mp.sqrt(max(mp.mpf(0), 1 - y * y))
            return beta / mp.pi if s == 0 else mp.sinh(beta * s) / (s * mp.pi)
        return f
    if kernel == "gaussian":
        return lambda y: mp.exp(-((y * M * dx) ** 2) / tau)
    if kernel == "es":
        return lambda y: mp.exp(beta * (mp.sqrt(max(mp.mpf(0), 1 - y * y)) - 1))
    raise ValueError(kernel)


def cardinal_bspline(n, t):
    """Centred cardinal B-spline of order n (degree n - 1, knots at integers shifted by n / 2), integral 1."""
    t = t + mp.mpf(n) / 2
    if t <= 0 or t >= n:
        return mp.mpf(0)
    s = mp.mpf(0)
    for k in range(n + 1):
        if t - k > 0:
            s += (-1) ** k * mp.binomial(n, k) * (t - k) ** (n - 1)
    return s / mp.factorial(n - 1)


CASES = [(T, k, M, sig) for T in (np.float64, np.float32) for k in ("kaiser_bessel", "backwards_kaiser_bessel", "gaussian", "bspline", "es")
         for M, sig in ((4, 2.0), (4, 1.5), (6, 1.25), (8, 2.0), (2, 2.0))]


@pytest.mark.parametrize("T,kernel,M,sigma", CASES)
def test_shape_and_fourier_tables(T, kernel, M, sigma):
    N = 48
    p = OraclePlan(T, N, m=M, sigma=sigma, kernel=kernel)
    kd = p.kernel_data(0)
    Nos = p.Nos[0]
    sig_T = T(T(Nos) / T(N))                                       # src/plan.jl:503
    beta, ell = shape_params(kernel, M, sig_T)
    dx = TWO_PI / Nos
    w = M * dx
    eps = float(np.finfo(T).eps)
    if beta is not None:
        assert abs(kd["beta"] - float(beta)) <= 4 * ulp(T, float(beta)), (kd["beta"], float(beta))
        beta = mp.mpf(kd["beta"])                                  # downstream formulae use the ROUNDED parameter, as the reference
    tau = None
    if kernel == "gaussian":
        sig_phys = mp.mpf(float(T(float(ell)))) * dx                # ell is stored in T (gaussian.jl:110), tau = 2 sigma^2
        tau = 2 * sig_phys ** 2
        assert abs(kd["tau"] - float(tau)) <= 8 * ulp(T, float(tau))
        tau = mp.mpf(kd["tau"])
    ks = p.ks[0]
    ref = []
    for k in ks:
        k = mp.mpf(int(k))
        if kernel == "kaiser_bessel":
            s = mp.sqrt(beta ** 2 - (w * k) ** 2)
            ref.append(2 * w * mp.sinh(s) / s)
        elif kernel == "backwards_kaiser_bessel":
            ref.append(w * mp.besseli(0, mp.sqrt(beta ** 2 - (w * k) ** 2)))
        elif kernel == "gaussian":
            ref.append(mp.exp(-tau * k ** 2 / 4) * mp.sqrt(mp.pi * tau))
        elif kernel == "es":
            if int(k) % 5 not in (0, 1) and abs(int(k)) != int(max(abs(ks))):
                ref.append(None)                                  # (50-digit quadrature is slow: a subset of the wavenumbers)
                continue
            f = lambda y: mp.exp(beta * (mp.sqrt(1 - y * y) - 1)) * mp.cos(k * w * y)
            ref.append(2 * w * mp.quad(f, mp.linspace(0, 1, 5)))
        else:
            kh = k * dx / 2
            ref.append(dx if k == 0 else (mp.sin(kh) / kh) ** (2 * M) * dx)
    sel = np.array([r is not None for r in ref])
    ref = np.array([float(r) for r in ref if r is not None])
    # relative error of phihat: a few ulp of T times the conditioning of sinh / I0 / exp at arguments of size beta (~ 2.3 M pi)
    rel = np.abs(p.phihat[0].astype(np.float64)[sel] - ref) / np.abs(ref)
    if kernel == "es":
        # no closed form: the quadrature sum of an oscillatory integrand is accurate to a few ulp of its LARGEST value phihat(0),
        # the cancellation at high k (phihat falls by 1e-3 ... 1e-6 across the kept band) is inherent in T
        rel = np.abs(p.phihat[0].astype(np.float64)[sel] - ref) / np.abs(ref).max()
    assert rel.max() <= (60 + 8 * M * np.pi) * eps, rel.max() / eps


@pytest.mark.parametrize("T,kernel,M,sigma", [c for c in CASES if c[1] in ("kaiser_bessel", "backwards_kaiser_bessel", "es")])
def test_piecewise_polynomial_is_the_chebyshev_interpolant(T, kernel, M, sigma):
    N = 48
    p = OraclePlan(T, N, m=M, sigma=sigma, kernel=kernel)
    kd = p.kernel_data(0)
    Nos = p.Nos[0]
    dx = TWO_PI / Nos
    beta = mp.mpf(kd["beta"])
    f = kernel_func(kernel, beta, None, M, dx)
    npoly, L = M + 4, 2 * M
    nodes = [mp.cospi((mp.mpf(i) - mp.mpf(1) / 2) / npoly) for i in range(1, npoly + 1)]

    def interpolant(j, t):          # j = 1..2M (right to left), t in [-1, 1]: Lagrange form, exact up to 50 digits
        h, d = 1 - 2 * (mp.mpf(j) - mp.mpf(1) / 2) / L, mp.mpf(1) / L
        tot = mp.mpf(0)
        for a, xa in enumerate(nodes):
            la = mp.mpf(1)
            for b, xb in enumerate(nodes):
                if a != b:
                    la *= (t - xb) / (xa - xb)
            tot += la * f(h + xa * d)
        return tot

    eps = float(np.finfo(T).eps)
    fmax = float(f(mp.mpf(0)))
    worst_i = worst_f = 0.0
    for X in (0.001953125, 0.03125, 0.25, 0.5, 0.671875, 0.96875):
        cell = 7
        x = T((cell + X) * float(dx))
        i, vals = p.evaluate_kernel(x, 0, mode="fast")
        r = (mp.mpf(float(x)) / TWO_PI) * Nos                      # the point the oracle saw (x rounded to T)
        Xe = r - (i - 1)
        assert i == cell + 1 and 0 <= Xe < 1
        for j in range(1, L + 1):
            t = 2 * Xe - 1
            worst_i = max(worst_i, abs(float(vals[j - 1]) - float(interpolant(j, t))) / fmax)
            y = (M - j + Xe) / M                                   # offset of grid point i - M + j from the point, in half widths
            worst_f = max(worst_f, abs(float(vals[j - 1]) - float(f(y))) / fmax)
    # (a) the Horner values ARE the Chebyshev interpolant, up to the conditioning (~ 2^(M + 3)) of the monomial Vandermonde
    #     solve that the reference performs in T (piecewise_polynomial.jl:39-40)
    assert worst_i <= 2.0 ** (M + 5) * eps, (worst_i / eps)
    # (b) and the interpolant approximates the kernel to the accuracy the reference's own test demands at HalfSupport(4),
    #     sigma = 1.5 (test/approx_window_functions.jl:9-24, rtol = 1e-7 on the value vector); degree M + 3 = 5 at
    #     HalfSupport(2) reaches 1.2e-5, which is all a 1e-3-accurate M = 2 transform needs
    bound = max(2e-7, 2.0 ** (M + 5) * eps) if M >= 4 else 2e-5
    if kernel == "es":
        # the ES kernel ends in a square root: on the outermost sub-intervals a polynomial misses it by ~ exp(-beta) sqrt(1/M)
        # of the peak (4.9e-5 at M = 2, 4e-9 at M = 4) — the size of the kernel's own truncation error, as in FINUFFT
        bound += 0.6 * float(mp.exp(-beta))
    assert worst_f <= bound, worst_f


@pytest.mark.parametrize("T,kernel,M,sigma", CASES)
def test_direct_and_gridding_evaluation(T, kernel, M, sigma):
    N = 48
    p = OraclePlan(T, N, m=M, sigma=sigma, kernel=kernel)
    kd = p.kernel_data(0)
    Nos = p.Nos[0]
    dx = TWO_PI / Nos
    eps = float(np.finfo(T).eps)
    beta = mp.mpf(kd["beta"]) if kernel.endswith("bessel") or kernel == "es" else None
    tau = mp.mpf(kd["tau"]) if kernel == "gaussian" else None
    L = 2 * M
    for X in (0.001953125, 0.0625, 0.4375, 0.90625):
        cell = 11
        x = T((cell + X) * float(dx))
        r = (mp.mpf(float(x)) / TWO_PI) * Nos
        for mode in ("direct", "fast"):
            if mode == "fast" and (kernel.endswith("bessel") or kernel == "es"):
                continue                                           # covered by the interpolant test
            i, vals = p.evaluate_kernel(x, 0, mode=mode)
            Xe = r - (i - 1)
            ref = []
            for j in range(1, L + 1):
                if kernel == "bspline":
                    # value of the B-spline centred on grid point i - M + j at the point: offset (M - j + X) cells
                    ref.append(cardinal_bspline(L, M - j + Xe))
                else:
                    ref.append(kernel_func(kernel, beta, tau, M, dx)((M - j + Xe) / M))
            ref = np.array([float(v) for v in ref])
            err = np.abs(vals.astype(np.float64) - ref).max() / np.abs(ref).max()
            # conditioning: exp / sinh / I0 of arguments up to beta ~ 2.3 M pi, or 2M - 1 recursion levels
            assert err <= (40 + 10 * M * np.pi) * eps, (kernel, mode, X, err / eps)
