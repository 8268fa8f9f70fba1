"""
CPU tests (no GPU): the kernel tables the PRODUCT library computes on the host (nufft_kernel_tables: shape parameter, piecewise
polynomial coefficients, phihat over the kept wavenumbers — what every plan uploads to the device) against the oracle's, for every
kernel, precision and a spread of (M, sigma).  The oracle's tables are pinned to the reference's formulae at 50 digits in
tests/test_oracle_mpmath.py; this closes the chain library -> oracle -> formulae without a GPU.
"""
import numpy as np
import pytest

import nufft_b200 as nb
from oracle import OraclePlan
from helpers import KERNEL_CLASSES

CASES = [(T, k, M, sig) for T in (np.float64, np.float32, np.complex64, np.complex128)
         for k in ("kaiser_bessel", "backwards_kaiser_bessel", "gaussian", "bspline", "es")
         for M, sig in ((4, 2.0), (4, 1.5), (6, 1.25), (8, 2.0), (2, 2.0))]


@pytest.mark.parametrize("T,kernel,M,sigma", CASES)
def test_library_tables_equal_the_oracle(T, kernel, M, sigma):
    import torch
    tdt = {np.float32: torch.float32, np.float64: torch.float64, np.complex64: torch.complex64, np.complex128: torch.complex128}[T]
    dims = (48, 40)
    op = OraclePlan(T, dims, m=M, sigma=sigma, kernel=kernel)
    rt = np.dtype(op.T).type
    eps = float(np.finfo(rt).eps)
    for d in range(2):
        kt = nb.kernel_tables(tdt, dims, d, m=M, sigma=sigma, kernel=getattr(nb, KERNEL_CLASSES[kernel])(),
                              kernel_evalmode=nb.FastApproximation())
        kd = op.kernel_data(d)
        assert kt["os_dim"] == op.Nos[d]
        ph_o = op.phihat[d].astype(np.float64)
        assert kt["phihat"].shape == ph_o.shape
        if kernel == "es":      # a quadrature sum: accurate to an ulp of its largest term, phihat(0)
            np.testing.assert_allclose(kt["phihat"], ph_o, rtol=0, atol=50 * eps * np.abs(ph_o).max())
        else:
            np.testing.assert_allclose(kt["phihat"], ph_o, rtol=50 * eps)
        if kernel in ("kaiser_bessel", "backwards_kaiser_bessel", "es"):
            assert abs(kt["shape"] - kd["beta"]) <= 4 * np.spacing(rt(kd["beta"]))
            cs_o = kd["cs"].astype(np.float64)
            scale = np.abs(cs_o).max()
            # two LU solves of the same ill-conditioned monomial Vandermonde system in T, different elimination order
            assert np.abs(kt["cs"] - cs_o).max() <= 2.0 ** (M + 6) * eps * scale
        elif kernel == "gaussian":
            assert abs(kt["shape"] - kd["tau"]) <= 8 * np.spacing(rt(kd["tau"]))


def test_real_data_plans_use_the_half_spectrum_along_the_first_dimension():
    import torch
    kt = nb.kernel_tables(torch.float64, (48, 40), 0, m=4, sigma=2.0, kernel=nb.BackwardsKaiserBesselKernel(),
                          kernel_evalmode=nb.FastApproximation())
    op = OraclePlan(np.float64, (48, 40), m=4, sigma=2.0, kernel="backwards_kaiser_bessel")
    assert kt["phihat"].shape == (25,) and kt["os_dim"] == op.Nos[0]


def test_es_kernel_refuses_direct_evaluation_without_a_gpu():
    import torch
    with pytest.raises(nb.ArgumentError):
        nb.kernel_tables(torch.complex64, (32, 32), 0, kernel=nb.ESKernel(), kernel_evalmode=nb.Direct())


def test_argument_errors_of_the_reference_without_a_gpu():
    """The option checks run before any device is touched (host_plan.cu: plan_sizes): the reference's ArgumentError cases —
    oversampled size below the kernel support (src/plan.jl:545-556, test/errors.jl), sigma < 1, fftshift with real data —
    are raised as ArgumentError by the same code path that creates plans."""
    import torch
    with pytest.raises(nb.ArgumentError, match="too small"):
        nb.kernel_tables(torch.float64, (6,), 0, m=8, sigma=1.25, kernel=nb.KaiserBesselKernel())
    with pytest.raises(nb.ArgumentError, match="sigma"):
        nb.kernel_tables(torch.complex64, (32, 32), 0, m=4, sigma=0.9, kernel=nb.KaiserBesselKernel())
    with pytest.raises(nb.ArgumentError, match="fftshift"):
        nb.kernel_tables(torch.float32, (32, 32), 0, m=4, sigma=2.0, kernel=nb.KaiserBesselKernel(), fftshift=True)
    with pytest.raises(nb.ArgumentError, match="out of range"):
        nb.kernel_tables(torch.complex64, (32, 32), 2, m=4, sigma=2.0, kernel=nb.KaiserBesselKernel())


def test_oversampled_sizes_follow_the_reference_rule_without_a_gpu():
    """Ñ = nextprod((2, 3, 5), floor(sigma N)) with sigma converted to T first (src/plan.jl:575-576, 485-498); real data: the first
    dimension is rounded on N/2 and doubled."""
    import torch
    for T, tdt in ((np.float32, torch.float32), (np.float64, torch.float64), (np.complex64, torch.complex64)):
        for dims, sigma in (((48, 40), 1.5), ((97, 33), 1.25), ((256,), 2.0), ((30, 30, 30), 1.1)):
            op = OraclePlan(T, dims, m=2, sigma=sigma)
            for d in range(len(dims)):
                kt = nb.kernel_tables(tdt, dims, d, m=2, sigma=sigma, kernel=nb.BackwardsKaiserBesselKernel(),
                                      kernel_evalmode=nb.FastApproximation())
                assert kt["os_dim"] == op.Nos[d], (T, dims, sigma, d)
