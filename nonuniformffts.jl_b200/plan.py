"""
Host-side mirror of the reference's user API for the NUFFT hot path, on top of the C ABI.

    PlanNUFFT(dtype, dims; m, sigma, kernel, ntransforms, fftshift, ...)   src/plan.jl:467-599
    set_points!(p, points)                                                  src/set_points.jl:33-88
    exec_type1!(us, p, vp; callbacks)  /  exec_type2!(vp, p, us; callbacks) src/NonuniformFFTs.jl:148-291
    size(p), ndims(p), ntransforms(p), eltype(p)                            src/plan.jl:360-435

Same names, argument meaning and error behaviour (ArgumentError -> ``ArgumentError(ValueError)``,
DimensionMismatch -> ``DimensionMismatch(ValueError)``).  Arrays are torch CUDA tensors (torch is used
for device memory and streams only).  Layout: Julia arrays are column-major, so a uniform array of
Julia dims ``size(p) = (N1, N2, N3)`` is a C-contiguous tensor of shape ``(N3, N2, N1)``
(``plan.shape``); dimension 1 of the reference is the last (contiguous) torch dimension.

All compute happens in libnufft_b200.so; nothing here touches the data.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import nufft_callbacks, nufft_opts


class ArgumentError(ValueError):
    """Julia's ArgumentError."""


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch."""


class NUFFTError(RuntimeError):
    """CUDA / cuFFT / allocation / state errors reported by the library."""


def _check(rc: int) -> None:
    if rc == _lib.NUFFT_SUCCESS:
        return
    msg = _lib.last_error()
    if rc in (_lib.NUFFT_ERR_ARG, _lib.NUFFT_ERR_UNSUPPORTED):
        raise ArgumentError(msg)
    if rc == _lib.NUFFT_ERR_DIM:
        raise DimensionMismatch(msg)
    raise NUFFTError(f"[{rc}] {msg}")


# ---- kernels (src/NonuniformFFTs.jl:23-35) and HalfSupport (src/Kernels/Kernels.jl:8-10) --------
@dataclass(frozen=True)
class HalfSupport:
    M: int


@dataclass(frozen=True)
class KaiserBesselKernel:
    beta: Optional[float] = None
    name = "kaiser_bessel"

    @property
    def param(self):
        return self.beta


@dataclass(frozen=True)
class BackwardsKaiserBesselKernel:
    beta: Optional[float] = None
    name = "backwards_kaiser_bessel"

    @property
    def param(self):
        return self.beta


@dataclass(frozen=True)
class GaussianKernel:
    ell: Optional[float] = None      # l / dx
    name = "gaussian"

    @property
    def param(self):
        return self.ell


@dataclass(frozen=True)
class BSplineKernel:
    name = "bspline"
    param = None


@dataclass(frozen=True)
class ESKernel:
    """"Exponential of semicircle" exp(beta (sqrt(1 - y^2) - 1)) (Barnett, Magland & af Klinteberg 2019).  NOT in the reference
    (src/Kernels/ holds the four kernels above): parity unpinned, checked against exact sums.  FastApproximation only."""
    beta: Optional[float] = None
    name = "es"

    @property
    def param(self):
        return self.beta


_KERNEL_BY_NAME = {
    "kaiser_bessel": KaiserBesselKernel, "backwards_kaiser_bessel": BackwardsKaiserBesselKernel,
    "gaussian": GaussianKernel, "bspline": BSplineKernel, "es": ESKernel,
}


class Direct:
    name = "direct"


class FastApproximation:
    name = "fast"


@dataclass
class NUFFTCallbacks:
    """NUFFTCallbacks (src/plan.jl:146-164) restricted to the menu the C ABI offers:
    nonuniform: tensor of weights w[n] (v -> v * w[n], n = original point index);
    uniform: dense tensor f[I] of shape plan.shape, or a tuple of D 1-D tensors (separable factor)."""
    nonuniform: Optional[torch.Tensor] = None
    uniform: object = None
    # general callbacks (the reference's arbitrary closures): CUDA C++ source compiled with NVRTC once per plan, defining
    # nufft_cb_nonuniform(nufft_cell (&v)[NUFFT_C], long long n, const void *user) after `#define NUFFT_HAS_NONUNIFORM 1`
    # and / or nufft_cb_uniform(nufft_cplx (&w)[NUFFT_C], const int (&idx)[3], const void *user) after
    # `#define NUFFT_HAS_UNIFORM 1` (include/nufft_b200.h); user_data: optional device tensor handed to them
    source: Optional[str] = None
    user_data: Optional[torch.Tensor] = None


_REAL = {torch.float32: torch.float32, torch.float64: torch.float64,
         torch.complex64: torch.float32, torch.complex128: torch.float64}
_CPLX = {torch.float32: torch.complex64, torch.float64: torch.complex128}


def _to_torch_dtype(dt) -> torch.dtype:
    if isinstance(dt, torch.dtype):
        return dt
    import numpy as np
    m = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
         np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}
    return m[np.dtype(dt)]


def build_opts(lib, dims, is_complex, real_dtype, M, sigma, kernel, mode_name, ntransforms, fftshift, sort_points, gpu_method,
               block_size, point_convention, device_index, stream, timer, spread_chunk) -> nufft_opts:
    """nufft_opts of a plan = the keyword arguments of PlanNUFFT (src/plan.jl:467-599)."""
    o = nufft_opts()
    _check(lib.nufft_opts_default(C.byref(o)))
    o.dim = len(dims)
    for d, n in enumerate(dims):
        o.n_modes[d] = n
    o.is_complex = 1 if is_complex else 0
    o.dtype = _lib.NUFFT_F64 if real_dtype == torch.float64 else _lib.NUFFT_F32
    o.half_support = M
    o.sigma = float(sigma)
    o.kernel = _lib.KERNEL_IDS[kernel.name]
    o.kernel_param = float("nan") if kernel.param is None else float(kernel.param)
    o.eval_mode = _lib.EVAL_IDS[mode_name]
    o.ntransforms = int(ntransforms)
    o.fftshift = 1 if fftshift else 0
    o.sort_points = 1 if sort_points else 0
    o.gpu_method = _lib.METHOD_IDS[gpu_method]
    if block_size is not None:
        if isinstance(block_size, int):
            # the reference turns a linear block size into power-of-two block dims (src/plan.jl:437-449) for its CPU path and
            # ignores it in the GPU shared-memory method (src/plan.jl:216,236-238); this backend sizes its bins itself
            if block_size < 1:
                raise ArgumentError("block_size must be positive")
        else:
            for d, b in enumerate(block_size):
                o.block_dims[d] = int(b)
    o.point_convention = int(point_convention)
    o.device = -1 if device_index is None else int(device_index)
    if stream is not None:
        o.stream = C.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
    o.record_timings = 1 if timer else 0
    o.spread_chunk = int(spread_chunk)
    return o


def kernel_tables(dtype, dims, d: int = 0, m=4, sigma: float = 2.0, kernel=None, kernel_evalmode=None, fftshift: bool = False):
    """Kernel data of dimension ``d`` that ``PlanNUFFT(dtype, dims, ...)`` would hold — computed by the library's host code without
    a plan and without a GPU (``nufft_kernel_tables``): dict(shape, dx, os_dim, cs[(M+4), 2M], phihat[size(p)[d]])."""
    import numpy as np
    lib = _lib.load()
    dtype = _to_torch_dtype(dtype)
    real_dtype = _REAL[dtype]
    dims = (int(dims),) if isinstance(dims, int) else tuple(int(n) for n in dims)
    M = m.M if isinstance(m, HalfSupport) else int(m)
    kernel = KaiserBesselKernel() if kernel is None else (_KERNEL_BY_NAME[kernel]() if isinstance(kernel, str) else kernel)
    mode = kernel_evalmode if kernel_evalmode is not None else (FastApproximation() if kernel.name == "es" else Direct())
    o = build_opts(lib, dims, dtype.is_complex, real_dtype, M, sigma, kernel, mode if isinstance(mode, str) else mode.name, 1, fftshift,
                   False, "auto", None, 0, None, None, False, 0)
    if not 0 <= d < len(dims):
        raise ArgumentError(f"dimension {d} out of range")
    nk = dims[d] // 2 + 1 if (not dtype.is_complex and d == 0) else dims[d]
    shape, dx, os_dim = C.c_double(), C.c_double(), C.c_int64()
    cs = np.zeros((M + 4, 2 * M), dtype=np.float64)
    ph = np.zeros(nk, dtype=np.float64)
    _check(lib.nufft_kernel_tables(C.byref(o), d, C.byref(shape), C.byref(dx), C.byref(os_dim), cs.ctypes.data_as(C.c_void_p), cs.size,
                                   ph.ctypes.data_as(C.c_void_p), ph.size))
    return dict(shape=shape.value, dx=dx.value, os_dim=os_dim.value, cs=cs, phihat=ph)


class PlanNUFFT:
    """PlanNUFFT([T = ComplexF64], dims; m = 4, sigma = 2, kernel, ntransforms = 1, fftshift = false, ...).

    Defaults follow the reference's CUDA backend: KaiserBesselKernel + Direct evaluation
    (ext/NonuniformFFTsCUDAExt.jl:19-23); pass ``kernel=BackwardsKaiserBesselKernel(),
    kernel_evalmode=FastApproximation()`` to reproduce the reference CPU path exactly.
    """

    def __init__(self, dtype=torch.complex128, dims=None, m=4, sigma: float = 2.0, kernel=None,
                 ntransforms: int = 1, fftshift: bool = False, sort_points: bool = False,
                 block_size=None, kernel_evalmode=None, gpu_method: str = "auto", synchronise: bool = False,
                 timer: bool = False, device=None, stream=None, point_convention: int = 0,
                 spread_chunk: int = 0, backend=None):
        if dims is None:                      # PlanNUFFT(N or dims) with the ComplexF64 default (src/plan.jl:597-599)
            dims, dtype = dtype, torch.complex128
        self._h = C.c_void_p(None)
        self._lib = _lib.load()
        dtype = _to_torch_dtype(dtype)
        if dtype not in _REAL:
            raise ArgumentError(f"unsupported data type {dtype}")
        dims = (int(dims),) if isinstance(dims, int) else tuple(int(d) for d in dims)
        if not 1 <= len(dims) <= 3:
            raise ArgumentError("only 1, 2 and 3 dimensions are supported")
        M = m.M if isinstance(m, HalfSupport) else int(m)
        if kernel is None:
            kernel = KaiserBesselKernel()
        if isinstance(kernel, str):
            kernel = _KERNEL_BY_NAME[kernel]()
        mode = kernel_evalmode if kernel_evalmode is not None else (FastApproximation() if kernel.name == "es" else Direct())
        mode_name = mode if isinstance(mode, str) else mode.name
        if gpu_method not in _lib.METHOD_IDS:
            raise ArgumentError("expected gpu_method in (auto, global_memory, shared_memory)")
        self.dtype = dtype
        self.real_dtype = _REAL[dtype]
        self.complex_dtype = _CPLX[self.real_dtype]
        self.is_complex = dtype.is_complex
        self.M = M
        self.kernel = kernel
        self.synchronise = bool(synchronise)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())

        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        self.stream = stream
        o = build_opts(self._lib, dims, self.is_complex, self.real_dtype, M, sigma, kernel, mode_name, ntransforms, fftshift,
                       sort_points, gpu_method, block_size, point_convention, self.device.index, stream, timer, spread_chunk)
        self._opts = o
        h = C.c_void_p()
        _check(self._lib.nufft_plan_create(C.byref(h), C.byref(o)))
        self._h = h
        sz = (C.c_int64 * 3)()
        osz = (C.c_int64 * 3)()
        nt = C.c_int32()
        _check(self._lib.nufft_plan_shape(self._h, sz, osz, C.byref(nt)))
        D = len(dims)
        self._ndims = D
        self.size = tuple(int(sz[d]) for d in range(D))            # Base.size(p): Julia order
        self.shape = self.size[::-1]                                # torch (C-order) shape of uniform arrays
        self.oversampled_dims = tuple(int(osz[d]) for d in range(D))
        self._ntransforms = int(nt.value)
        self.points = None
        self.Np = None
        self._keep = None

    # -- Base methods ------------------------------------------------------------------------
    def ndims(self) -> int:
        return self._ndims

    def ntransforms(self) -> int:
        return self._ntransforms

    def eltype(self):
        return self.complex_dtype

    def __repr__(self) -> str:
        buf = C.create_string_buffer(4096)
        _check(self._lib.nufft_describe(self._h, buf, 4096))
        return buf.value.decode()

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nufft_plan_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------------
    def _ptr_array(self, tensors: Sequence[torch.Tensor]):
        arr = (C.c_void_p * len(tensors))()
        for i, t in enumerate(tensors):
            arr[i] = t.data_ptr()
        return arr

    def _check_dev(self, t: torch.Tensor, what: str) -> None:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise ArgumentError(f"{what} must be a CUDA tensor (no CPU fallback in this backend)")
        if t.device != self.device:
            raise ArgumentError(f"{what} lives on {t.device}, the plan on {self.device}")
        if not t.is_contiguous():
            raise ArgumentError(f"{what} must be contiguous")

    def _callbacks(self, cb: Optional[NUFFTCallbacks]):
        if cb is None or (cb.nonuniform is None and cb.uniform is None and not cb.source):
            return None, None
        s = nufft_callbacks()
        s.struct_size = C.sizeof(nufft_callbacks)
        keep = []
        if cb.source:
            src = cb.source.encode()
            s.nvrtc_src = src
            keep.append(src)
            if cb.user_data is not None:
                self._check_dev(cb.user_data, "callback user data")
                s.user_data = cb.user_data.data_ptr()
                keep.append(cb.user_data)
        if cb.nonuniform is not None:
            w = cb.nonuniform
            self._check_dev(w, "nonuniform callback weights")
            if w.dtype != self.real_dtype or w.numel() != self.Np:
                raise ArgumentError("nonuniform callback weights must be a real tensor of length Np with the plan's precision")
            s.nu_weights = w.data_ptr()
            keep.append(w)
        if cb.uniform is not None:
            if isinstance(cb.uniform, torch.Tensor):
                f = cb.uniform
                self._check_dev(f, "uniform callback factor")
                if f.dtype != self.real_dtype or tuple(f.shape) != self.shape:
                    raise ArgumentError(f"dense uniform callback factor must have shape {self.shape} and dtype {self.real_dtype}")
                s.u_factor_dense = f.data_ptr()
                keep.append(f)
            else:
                fs = list(cb.uniform)
                if len(fs) != self._ndims:
                    raise ArgumentError("separable uniform callback needs one table per dimension")
                for d, f in enumerate(fs):
                    self._check_dev(f, "uniform callback table")
                    if f.dtype != self.real_dtype or f.numel() != self.size[d]:
                        raise ArgumentError("separable uniform callback table d must have length size(p)[d]")
                arr = self._ptr_array(fs)
                s.u_factor_sep = C.cast(arr, C.POINTER(C.c_void_p))
                keep += fs + [arr]
        return s, keep

    # -- set_points! ----------------------------------------------------------------------------
    def set_points(self, xp) -> "PlanNUFFT":
        """set_points!(p, xp): tuple of D vectors (preferred), a single vector in 1-D, or a (Np, D)
        tensor == Julia (d, Np) matrix / Vector{SVector{D}} (src/set_points.jl:62-88; read in place by the
        binning kernel through nufft_set_points_matrix when contiguous)."""
        D = self._ndims
        if isinstance(xp, torch.Tensor):
            if xp.ndim == 1:
                if D != 1:
                    raise DimensionMismatch(f"expected {D}-dimensional points")
                xp = (xp,)
            elif xp.ndim == 2:
                if xp.shape[1] != D:
                    raise DimensionMismatch(f"expected input matrix to have dimensions ({D}, Np)")
                if xp.is_contiguous():
                    if xp.dtype != self.real_dtype:
                        raise ArgumentError(
                            f"input points must have the same accuracy as the created plan (got {xp.dtype} points for a {self.dtype} plan)")
                    self._check_dev(xp, "points")
                    self.points = (xp,)
                    self.Np = xp.shape[0]
                    _check(self._lib.nufft_set_points_matrix(self._h, self.Np, C.c_void_p(xp.data_ptr())))
                    if self.synchronise:
                        self.stream.synchronize()
                    return self
                xp = tuple(xp[:, d].contiguous() for d in range(D))
            else:
                raise ArgumentError("unexpected point container")
        xp = tuple(xp)
        if len(xp) != D:
            raise DimensionMismatch(f"expected {D}-dimensional points")
        for x in xp:
            if not isinstance(x, torch.Tensor):
                raise ArgumentError("unexpected point container: expected CUDA tensors")
            if x.dtype != self.real_dtype:
                raise ArgumentError(
                    f"input points must have the same accuracy as the created plan (got {x.dtype} points for a {self.dtype} plan)")
            self._check_dev(x, "points")
            if x.ndim != 1:
                raise ArgumentError("unexpected point container: expected 1-D tensors")
        Np = xp[0].numel()
        if any(x.numel() != Np for x in xp):
            raise DimensionMismatch("input points must have the same length along all dimensions")
        self.points = xp            # keep a reference, like points_ref[] (src/set_points.jl:45)
        self.Np = Np
        _check(self._lib.nufft_set_points(self._h, Np, self._ptr_array(xp)))
        if self.synchronise:
            self.stream.synchronize()
        return self

    # -- exec ---------------------------------------------------------------------------------
    def _uniform_list(self, us) -> list:
        us = list(us) if isinstance(us, (tuple, list)) else [us]
        if len(us) != self._ntransforms:
            raise DimensionMismatch(f"wrong amount of arrays (expected a tuple of {self._ntransforms} arrays)")
        for u in us:
            if not isinstance(u, torch.Tensor) or u.dtype != self.complex_dtype:
                raise ArgumentError(
                    f"uniform data must have the same accuracy as the created plan (got {getattr(u, 'dtype', type(u))} values for a {self.dtype} plan)")
            if u.ndim != self._ndims:
                raise DimensionMismatch(f"wrong dimensions of array (expected {self._ndims}-dimensional array)")
            if tuple(u.shape) != self.shape:
                raise DimensionMismatch(f"wrong dimensions of array (expected dimensions {self.shape})")
            self._check_dev(u, "uniform data")
        return us

    def _nonuniform_list(self, vp) -> list:
        vp = list(vp) if isinstance(vp, (tuple, list)) else [vp]
        if self.Np is None:
            raise NUFFTError("set_points must be called before exec_type1 / exec_type2")
        if len(vp) != self._ntransforms:
            raise DimensionMismatch(f"wrong amount of data vectors (expected a tuple of {self._ntransforms} vectors)")
        for v in vp:
            if not isinstance(v, torch.Tensor) or v.dtype != self.dtype:
                raise ArgumentError(f"non-uniform data must have element type {self.dtype}")
            if v.numel() != self.Np or v.ndim != 1:
                raise DimensionMismatch(
                    f"wrong length of data vector (it should match the number of points {self.Np}, got length {v.numel()})")
            self._check_dev(v, "non-uniform data")
        return vp

    def exec_type1(self, us, vp, callbacks: Optional[NUFFTCallbacks] = None):
        """exec_type1!(us, p, vp; callbacks): non-uniform -> uniform.  Returns ``us``."""
        ul, vl = self._uniform_list(us), self._nonuniform_list(vp)
        cb, keep = self._callbacks(callbacks)
        _check(self._lib.nufft_exec_type1(self._h, self._ptr_array(ul), self._ptr_array(vl),
                                          C.byref(cb) if cb is not None else None))
        if self.synchronise:
            self.stream.synchronize()
        return us

    def exec_type2(self, vp, us, callbacks: Optional[NUFFTCallbacks] = None):
        """exec_type2!(vp, p, us; callbacks): uniform -> non-uniform.  Returns ``vp``."""
        ul, vl = self._uniform_list(us), self._nonuniform_list(vp)
        cb, keep = self._callbacks(callbacks)
        _check(self._lib.nufft_exec_type2(self._h, self._ptr_array(vl), self._ptr_array(ul),
                                          C.byref(cb) if cb is not None else None))
        if self.synchronise:
            self.stream.synchronize()
        return vp

    # -- stage-level calls (multi-GPU layer) ---------------------------------------------------
    def type1_spread(self, vp, callbacks=None):
        vl = self._nonuniform_list(vp)
        cb, keep = self._callbacks(callbacks)
        _check(self._lib.nufft_type1_spread(self._h, self._ptr_array(vl), C.byref(cb) if cb is not None else None))

    def type1_finish(self, us, callbacks=None):
        ul = self._uniform_list(us)
        cb, keep = self._callbacks(callbacks)
        _check(self._lib.nufft_type1_finish(self._h, self._ptr_array(ul), C.byref(cb) if cb is not None else None))
        return us

    def type2_prepare(self, us, callbacks=None):
        ul = self._uniform_list(us)
        cb, keep = self._callbacks(callbacks)
        _check(self._lib.nufft_type2_prepare(self._h, self._ptr_array(ul), C.byref(cb) if cb is not None else None))

    def type2_interp(self, vp, callbacks=None):
        vl = self._nonuniform_list(vp)
        cb, keep = self._callbacks(callbacks)
        _check(self._lib.nufft_type2_interp(self._h, self._ptr_array(vl), C.byref(cb) if cb is not None else None))
        return vp

    def grid(self) -> torch.Tensor:
        """The plan-owned oversampled grid ``us`` as a torch view, shape (C, *oversampled_dims[::-1])."""
        ptr = C.c_void_p()
        nbytes = C.c_size_t()
        _check(self._lib.nufft_get_grid(self._h, C.byref(ptr), C.byref(nbytes)))
        shape = (self._ntransforms,) + self.oversampled_dims[::-1]
        return _tensor_from_ptr(ptr.value, shape, self.dtype, self.device, self)

    # -- introspection ------------------------------------------------------------------------
    def binning(self):
        """(perm, bin_offsets, bin_dims): 0-based int32 device tensors (views of plan memory)."""
        perm, off = C.c_void_p(), C.c_void_p()
        nb = C.c_int64()
        bd = (C.c_int64 * 3)()
        _check(self._lib.nufft_get_binning(self._h, C.byref(perm), C.byref(off), C.byref(nb), bd))
        p = _tensor_from_ptr(perm.value, (self.Np,), torch.int32, self.device, self) if self.Np else \
            torch.empty(0, dtype=torch.int32, device=self.device)
        o = _tensor_from_ptr(off.value, (nb.value + 1,), torch.int32, self.device, self)
        return p, o, tuple(int(bd[d]) for d in range(self._ndims))

    def binning_fine(self):
        """(perm, fine_offsets, sub_dims): the (bin, sub-bin) order the kernels use.  Column-streaming plans refine every
        bin (a column of 4 x 4 cells, up to 256 cells along z) into sub_dims single cells along z and do not materialise
        sub-bin offsets (fine_offsets is None); other plans: identical to `binning()`, sub_dims (1, 1, 1)."""
        perm, off = C.c_void_p(), C.c_void_p()
        nf = C.c_int64()
        sd = (C.c_int64 * 3)()
        _check(self._lib.nufft_get_binning_fine(self._h, C.byref(perm), C.byref(off), C.byref(nf), sd))
        p = _tensor_from_ptr(perm.value, (self.Np,), torch.int32, self.device, self) if self.Np else \
            torch.empty(0, dtype=torch.int32, device=self.device)
        o = _tensor_from_ptr(off.value, (nf.value + 1,), torch.int32, self.device, self) if off.value else None
        return p, o, tuple(int(sd[d]) for d in range(3))

    def kernel_info(self, d: int = 0):
        import numpy as np
        shape, dx = C.c_double(), C.c_double()
        cs = np.zeros(((self.M + 4), 2 * self.M), dtype=np.float64)
        ph = np.zeros(self.size[d], dtype=np.float64)
        _check(self._lib.nufft_plan_kernel_info(self._h, d, C.byref(shape), C.byref(dx),
                                                cs.ctypes.data_as(C.c_void_p), ph.ctypes.data_as(C.c_void_p)))
        return dict(shape=shape.value, dx=dx.value, cs=cs, phihat=ph)

    @property
    def timer(self):
        """Per-stage device times (ms) with the reference's stage names (needs timer=True)."""
        ms = (C.c_float * 16)()
        _check(self._lib.nufft_get_timings(self._h, ms))
        names = ["Set points", "T1 (0) Fill with zeros", "T1 (1) Spreading", "T1 (2) Forward FFT", "T1 (3) Deconvolution",
                 "T2 (0+1) Zero-pad + deconvolution", "T2 (2) Backward FFT", "T2 (3) Interpolation"]
        return {n: float(ms[i]) for i, n in enumerate(names)}


class _CudaArrayView:
    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}
        self._owner = owner


_TYPESTR = {torch.float32: "<f4", torch.float64: "<f8", torch.complex64: "<c8", torch.complex128: "<c16", torch.int32: "<i4"}


def _tensor_from_ptr(ptr, shape, dtype, device, owner) -> torch.Tensor:
    view = _CudaArrayView(ptr, shape, _TYPESTR[dtype], owner)
    with torch.cuda.device(device):
        return torch.as_tensor(view, device=device)


def launch_count(reset: bool = False) -> int:
    """Number of kernel launches issued by libnufft_b200 on this thread since the last reset."""
    return int(_lib.load().nufft_launch_count(1 if reset else 0))


# ---- functional aliases in the reference's spelling ------------------------------------------------
def set_points(p: PlanNUFFT, xp):
    return p.set_points(xp)


def exec_type1(us, p: PlanNUFFT, vp, callbacks=None):
    return p.exec_type1(us, vp, callbacks)


def exec_type2(vp, p: PlanNUFFT, us, callbacks=None):
    return p.exec_type2(vp, us, callbacks)
