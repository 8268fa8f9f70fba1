"""Shared helpers of the parity tests: seeded inputs, oracle/GPU plan pairs, error norms."""
from __future__ import annotations

import numpy as np

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}   # north_star: rel. L2 vs the reference CPU path


def l2_error(a, b) -> float:
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    den = np.sqrt(np.sum(np.abs(b) ** 2))
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2)) / (den if den > 0 else 1.0))


def real_of(dtype):
    dtype = np.dtype(dtype)
    return np.dtype(np.float32) if dtype in (np.dtype(np.float32), np.dtype(np.complex64)) else np.dtype(np.float64)


def complex_of(dtype):
    return np.dtype(np.complex64) if real_of(dtype) == np.float32 else np.dtype(np.complex128)


def make_points(rng, D, Np, rtype, dist="uniform", shift=True):
    """Synthetic point sets.  uniform: 2pi*U(0,1) (+ random multiples of 2pi, exercising the fold,
    as test/accuracy.jl:114-117); clustered: wrapped normal N(0,1) rad as the reference's benchmark draws
    (benchmark/CPU+CUDA/run_benchmarks.jl:47-51); blobs: 8 narrow Gaussian blobs; onecell: all points in one cell."""
    xs = []
    centres = rng.random((8, D)) * 2 * np.pi
    for d in range(D):
        if dist == "uniform":
            x = rng.random(Np) * 2 * np.pi
        elif dist == "clustered":
            x = rng.standard_normal(Np)
        elif dist == "blobs":
            x = centres[rng.integers(0, 8, Np), d] + 0.05 * rng.standard_normal(Np)
        elif dist == "onecell":
            x = 1.2345 + 1e-4 * rng.random(Np)
        else:
            raise ValueError(dist)
        if shift and dist == "uniform":
            x = x + rng.integers(-1, 2, Np) * 2 * np.pi
        xs.append(np.ascontiguousarray(x.astype(rtype)))
    return xs


def make_values(rng, n, dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(dtype)
    return rng.standard_normal(n).astype(dtype)


KERNEL_CLASSES = {"kaiser_bessel": "KaiserBesselKernel", "backwards_kaiser_bessel": "BackwardsKaiserBesselKernel",
                  "gaussian": "GaussianKernel", "bspline": "BSplineKernel", "es": "ESKernel"}


def gpu_plan(nufft, dtype, dims, *, m=4, sigma=2.0, kernel="backwards_kaiser_bessel", evalmode="fast",
             ntransforms=1, **kw):
    import torch
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
           np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}[np.dtype(dtype)]
    kern = getattr(nufft, KERNEL_CLASSES[kernel])()
    mode = nufft.FastApproximation() if evalmode == "fast" else nufft.Direct()
    return nufft.PlanNUFFT(tdt, dims, m=m, sigma=sigma, kernel=kern, kernel_evalmode=mode, ntransforms=ntransforms, **kw)


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()
