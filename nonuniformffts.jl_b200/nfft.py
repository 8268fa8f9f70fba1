"""
AbstractNFFTs-compatible front-end: host-side mirror of /root/reference `src/abstractNFFTs.jl:115-245`.

    p = NFFTPlan(xp, Ns; m=4, sigma=2.0)     # xp: (Np, D) tensor == Julia (D, Np) matrix, nodes in [-1/2, 1/2)
    mul(vp, p, fhat)                         # mul!(vp, p, f^)          : uniform -> non-uniform  (exec_type2!)
    mul_adjoint(fhat, p, vp)                 # mul!(f^, adjoint(p), vp) : non-uniform -> uniform  (exec_type1!)
    nodes(p, xp)                             # nodes!(p, xp)

Only the two hooks of the reference's wrapper touch the kernels, and both are flags of the C ABI: the NFFT point
convention (`point_convention = 1`: x in [-1/2, 1/2) -> -2 pi x folded to [0, 2 pi), `src/abstractNFFTs.jl:150-158`,
applied inside K-bin) and `fftshift = True` (NFFT.jl's increasing frequency order, `:194`).  Everything else is glue.
Only complex-to-complex plans exist in this interface (`:188`).
"""
from __future__ import annotations

import math

import torch

from .plan import (ArgumentError, BackwardsKaiserBesselKernel, BSplineKernel, DimensionMismatch, GaussianKernel,
                   KaiserBesselKernel, PlanNUFFT)

# convert_window_function (src/abstractNFFTs.jl:170-184): NFFT.jl's :kaiser_bessel_rev is KaiserBesselKernel here and
# its :kaiser_bessel is BackwardsKaiserBesselKernel; anything else -> the backend's default kernel (KB on CUDA,
# ext/NonuniformFFTsCUDAExt.jl:19)
_WINDOWS = {"gauss": GaussianKernel, "spline": BSplineKernel, "kaiser_bessel_rev": KaiserBesselKernel,
            "kaiser_bessel": BackwardsKaiserBesselKernel}


def accuracy_params(m=None, sigma=None, reltol=None):
    """AbstractNFFTs.accuracyParams restated (AbstractNFFTs.jl is an un-vendored dependency; compat "0.8, 0.9" in the
    reference's Project.toml): with `reltol` the window width is w = ceil(log10(1 / reltol)) + 1, m = w / 2 (integer
    division) and sigma = 2; otherwise m (default 4) and sigma (default 2) are taken as given and
    reltol = 10^-(2m - 1)."""
    if reltol is not None:
        w = int(math.ceil(math.log10(1.0 / reltol))) + 1
        return max(w // 2, 1), 2.0, float(reltol)
    m = 4 if m is None else int(m)
    sigma = 2.0 if sigma is None else float(sigma)
    return m, sigma, 10.0 ** (-(2 * m - 1))


class NFFTPlan:
    """NonuniformFFTs.NFFTPlan: wraps a complex PlanNUFFT with ntransforms = 1 (src/abstractNFFTs.jl:115-119)."""

    def __init__(self, xp: torch.Tensor, Ns, *, m=None, sigma=None, reltol=None, window=None, fftshift: bool = True,
                 blocking: bool = True, sortNodes: bool = False, precompute=None, **kws):
        if not isinstance(xp, torch.Tensor) or xp.ndim != 2 or not xp.dtype.is_floating_point:
            raise ArgumentError("NFFTPlan expects the non-uniform points as a real (Np, D) tensor [Julia: (D, Np) matrix]")
        Ns = tuple(int(n) for n in (Ns if hasattr(Ns, "__len__") else (Ns,)))
        if xp.shape[1] != len(Ns):
            raise DimensionMismatch(f"expected input matrix to have dimensions ({len(Ns)}, Np)")
        m_actual, sigma_actual, self.reltol = accuracy_params(m, sigma, reltol)
        if window is None or isinstance(window, str):
            kernel = _WINDOWS.get(window, KaiserBesselKernel)()
        else:
            kernel = window
        cdtype = torch.complex64 if xp.dtype == torch.float32 else torch.complex128
        if not blocking:                       # block_size = nothing -> NullBlockData -> the naive (global-memory) kernels
            kws.setdefault("gpu_method", "global_memory")
        self.p = PlanNUFFT(cdtype, Ns, m=m_actual, sigma=sigma_actual, kernel=kernel, fftshift=fftshift,
                           sort_points=bool(sortNodes), point_convention=1, device=xp.device, **kws)
        nodes(self, xp)

    # AbstractNFFTs.size_in / size_out (src/abstractNFFTs.jl:126-127)
    def size_in(self):
        return self.p.size

    def size_out(self):
        return (self.p.Np,)

    def __repr__(self):
        return f"NonuniformFFTs.NFFTPlan{{{self.p.real_dtype}, {len(self.p.size)}}} wrapping a PlanNUFFT:\n{self.p!r}"

    def close(self):
        self.p.close()


def nodes(p: NFFTPlan, xp: torch.Tensor) -> NFFTPlan:
    """AbstractNFFTs.nodes!(p, xp): locations in [-1/2, 1/2)^D (src/abstractNFFTs.jl:160-163)."""
    p.p.set_points(xp)
    return p


def mul(vp: torch.Tensor, p: NFFTPlan, us: torch.Tensor) -> torch.Tensor:
    """LinearAlgebra.mul!(vp, p, us): uniform to non-uniform (src/abstractNFFTs.jl:130-135)."""
    p.p.exec_type2(vp, us)
    return vp


def mul_adjoint(us: torch.Tensor, p: NFFTPlan, vp: torch.Tensor) -> torch.Tensor:
    """LinearAlgebra.mul!(us, adjoint(p), vp): non-uniform to uniform (src/abstractNFFTs.jl:138-145)."""
    p.p.exec_type1(us, vp)
    return us


def plan_nfft(xp: torch.Tensor, Ns, **kwargs) -> NFFTPlan:
    """AbstractNFFTs.plan_nfft(NonuniformFFTsBackend(), Q, xp, Ns; kwargs...) (src/abstractNFFTs.jl:236-243): the
    array type Q is the tensor's device here."""
    return NFFTPlan(xp, Ns, **kwargs)
