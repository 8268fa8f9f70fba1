"""
CPU tests (no GPU): the ES kernel ("exponential of semicircle", Barnett, Magland & af Klinteberg 2019) of the ORACLE.

The ES kernel is NOT in the reference (src/Kernels/ holds KB, BKB, Gaussian and B-spline): there is nothing to pin parity to
("parity unpinned").  What can be checked is that the construction is a valid NUFFT window — type-1 / type-2 errors against
exact NUDFT sums, with the protocol of the reference's own accuracy test (test/accuracy.jl:219-250: N = 256, Np = 512, points
shifted by random multiples of 2 pi) — and that it behaves like its siblings: thresholds are the reference's KB bounds relaxed by
the factor the ES kernel is known to lose against KB (about 2-3 x at equal support).  The tables themselves are pinned to 50-digit
evaluations of the definition in tests/test_oracle_mpmath.py.
"""
import numpy as np
import pytest

from oracle import OraclePlan, nudft_type1, nudft_type2
from helpers import l2_error, make_values
from test_oracle_reference_tests import _points_1d, _threshold_ref

ES_LOSS = 4.0          # ES against the reference's KB bound at the same (M, sigma)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("M", [4, 5, 6, 7, 8])
@pytest.mark.parametrize("sigma", [2.0, 1.25])
def test_es_accuracy_1d(dtype, M, sigma):
    N, Np = 256, 512
    rng = np.random.default_rng(42)
    p = OraclePlan(dtype, N, m=M, sigma=sigma, kernel="es")
    x = _points_1d(rng, Np, p.T)
    v = make_values(rng, Np, dtype)
    p.set_points(x)
    thr = ES_LOSS * _threshold_ref(dtype, "kaiser_bessel", M, sigma)
    e1 = l2_error(p.exec_type1(v), nudft_type1(p.ks, [x], v))
    assert e1 < thr, f"type-1 error {e1:.3e} >= {thr:.3e}"
    uk = make_values(rng, p.size[0], p.CT)
    e2 = l2_error(p.exec_type2(uk), nudft_type2(p.ks, [x], uk, not p.is_complex))
    assert e2 < thr, f"type-2 error {e2:.3e} >= {thr:.3e}"


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_es_accuracy_float32(dtype):
    """HalfSupport(4), sigma = 2: the headline class (1e-5 relative L2 in Float32; points in [0, 2 pi) — shifted by multiples of
    2 pi the Float32 rounding of the coordinates alone costs 1.5e-5 at k = 128, with every kernel)."""
    N, Np = 256, 512
    rng = np.random.default_rng(42)
    p = OraclePlan(dtype, N, m=4, sigma=2.0, kernel="es")
    x = (rng.random(Np) * 2 * np.pi).astype(p.T)
    v = make_values(rng, Np, dtype)
    p.set_points(x)
    assert l2_error(p.exec_type1(v), nudft_type1(p.ks, [x], v)) < 1e-5
    uk = make_values(rng, p.size[0], p.CT)
    assert l2_error(p.exec_type2(uk), nudft_type2(p.ks, [x], uk, not p.is_complex)) < 1e-5


def test_es_multidimensional_and_modes():
    """2-D, several blocks; FastApproximation (the polynomials, what the GPU evaluates) against Direct (exp / sqrt)."""
    Ns, Np, M = (48, 40), 800, 6
    rng = np.random.default_rng(5)
    xs = [(rng.random(Np) * 2 * np.pi) for _ in Ns]
    v = make_values(rng, Np, np.complex128)
    out = {}
    for mode in ("fast", "direct"):
        p = OraclePlan(np.complex128, Ns, m=M, sigma=2.0, kernel="es", evalmode=mode, block_size=16, use_blocked_spreading=True)
        p.set_points(xs)
        out[mode] = p.exec_type1(v)
        assert l2_error(out[mode], nudft_type1(p.ks, xs, v)) < ES_LOSS * 6 * 10.0 ** (-1.9 * M)
    assert l2_error(out["fast"], out["direct"]) < 1e-10


def test_es_shape_rule_and_explicit_parameter():
    """beta = 0.976 pi M (2 - 1/sigma) — 2.30 x (2M) at sigma = 2, the rule of the FINUFFT paper; an explicit beta is honoured."""
    p = OraclePlan(np.float64, 64, m=4, sigma=2.0, kernel="es")
    assert abs(p.kernel_data(0)["beta"] - 0.976 * np.pi * 4 * 1.5) < 1e-12
    assert abs(p.kernel_data(0)["beta"] / 8 - 2.30) < 2e-3
    q = OraclePlan(np.float64, 64, m=4, sigma=2.0, kernel="es", kernel_param=17.0)
    assert q.kernel_data(0)["beta"] == 17.0
