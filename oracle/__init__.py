"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement of NonuniformFFTs.jl v0.9.6's CPU path (set_points! -> exec_type1! /
exec_type2!).  Per-point arithmetic lives in C (``nufft_oracle_impl.h``, one function per
reference function, each citing reference file:line); plan-level integer logic is restated
here in numpy/Python; the FFT (FFTW.jl in the reference, an un-vendored dependency) is
pocketfft through ``scipy.fft`` with the same sign/normalisation conventions
(unnormalised forward e^{-i}, unnormalised backward e^{+i}; pinned by
tests/test_oracle_reference_tests.py::test_uniform_points, which restates
/root/reference/test/uniform_points.jl).

Who may import this: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and
``--impl reference``).  Pinning status: see the header of nufft_oracle_impl.h.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np
import scipy.fft as _sfft

_HERE = Path(__file__).resolve().parent
_LIB = None

KERNELS = {"kaiser_bessel": 0, "backwards_kaiser_bessel": 1, "gaussian": 2, "bspline": 3,
           "es": 4}          # "es": not in the reference (exponential of semicircle; parity unpinned, checked against exact sums)
EVALMODES = {"fast": 0, "direct": 1}


def build(force: bool = False) -> Path:
    """Compile the C oracle (gcc) into oracle/libnufft_oracle.so."""
    so = _HERE / "libnufft_oracle.so"
    srcs = [_HERE / "nufft_oracle.c", _HERE / "nufft_oracle_impl.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs if s.exists()):
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "CC=gcc"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_num_threads.restype = C.c_int
        for suf, ct in (("_f32", C.c_float), ("_f64", C.c_double)):
            getattr(_LIB, "orc_kernel_sizeof" + suf).restype = C.c_size_t
            getattr(_LIB, "orc_fold" + suf).restype = ct
            getattr(_LIB, "orc_fold" + suf).argtypes = [ct]
            f = getattr(_LIB, "orc_point_to_cell" + suf)
            f.restype = C.c_int64
            f.argtypes = [ct, C.c_int64, C.c_void_p]
            f = getattr(_LIB, "orc_kernel_init" + suf)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, ct, C.c_double]
            f = getattr(_LIB, "orc_kernel_eval" + suf)
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p, C.c_int, ct, C.c_void_p]
    return _LIB


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))


# ------------------------------------------------------------------------------------------
# plan-level logic (numpy / Python)
# ------------------------------------------------------------------------------------------

def nextprod235(n: int) -> int:
    """Base.nextprod((2,3,5), n): smallest 2^a 3^b 5^c >= n (used at src/plan.jl:492-494)."""
    n = max(int(n), 1)
    best = None
    p5 = 1
    while p5 < 5 * n:
        p35 = p5
        while p35 < 3 * n:
            v = p35
            while v < n:
                v *= 2
            if best is None or v < best:
                best = v
            p35 *= 3
        p5 *= 5
    return best


def oversampled_dims(Ns, sigma, real_data: bool, rtype) -> tuple:
    """src/plan.jl:485-498.  sigma is converted to the real type T first (plan.jl:575-576)."""
    out = []
    s = rtype(sigma)
    for d, N in enumerate(Ns):
        if real_data and d == 0:
            out.append(2 * nextprod235(int(math.floor(float(s * rtype((N + 1) // 2))))))
        else:
            out.append(nextprod235(int(math.floor(float(s * rtype(N))))))
    return tuple(out)


def wavenumbers(Ns, real_data: bool):
    """src/plan.jl:558-566 — fftfreq(N, N) (rfftfreq for the first dim of real-data plans); integers."""
    ks = []
    for d, N in enumerate(Ns):
        if real_data and d == 0:
            ks.append(np.arange(N // 2 + 1, dtype=np.int64))
        else:
            k = np.arange(N, dtype=np.int64)
            k[k >= (N + 1) // 2] -= N
            ks.append(k)
    return ks


def non_oversampled_indices(Nk: int, Nos: int, r2c: bool, fftshift: bool) -> np.ndarray:
    """src/NonuniformFFTs.jl:318-348, 0-based."""
    ax = np.arange(Nos, dtype=np.int64)
    m = np.empty(Nk, dtype=np.int64)
    if r2c:
        m[:] = ax[:Nk]
    elif Nk % 2 == 0:
        h = Nk // 2
        if fftshift:
            m[:h] = ax[Nos - h:]
            m[h:] = ax[:h]
        else:
            m[:h] = ax[:h]
            m[h:] = ax[Nos - h:]
    else:
        h = (Nk - 1) // 2
        if fftshift:
            m[:h] = ax[Nos - h:]
            m[h:] = ax[: h + 1]
        else:
            m[: h + 1] = ax[: h + 1]
            m[h + 1:] = ax[Nos - h:]
    return m


def get_block_dims(Nos, bsize) -> tuple:
    """src/plan.jl:437-451 — linear block size -> power-of-two dims by round-robin doubling."""
    if not isinstance(bsize, (int, np.integer)):
        return tuple(int(b) for b in bsize)
    d = len(Nos)
    bd = [1] * d
    prod, i = 1, 0
    while prod < bsize:
        bd[i] <<= 1
        prod <<= 1
        i = 0 if i == d - 1 else i + 1
    return tuple(bd)


class OraclePlan:
    """
    Restatement of ``PlanNUFFT`` (src/plan.jl:467-541) + ``set_points!`` (src/set_points.jl:33-52,
    src/blocking/cpu.jl:113-185) + ``exec_type1!`` / ``exec_type2!`` (src/NonuniformFFTs.jl:148-189,
    237-286) for the CPU backend.

    dtype: np.float32/np.float64 (real data) or np.complex64/np.complex128.
    kernel: one of KERNELS; evalmode: 'fast' (reference CPU default) or 'direct'.
    block_size: int | tuple | None  (None = NullBlockData, src/blocking/no_blocking.jl).
    """

    def __init__(self, dtype, Ns, m: int = 4, sigma: float = 2.0, kernel: str = "backwards_kaiser_bessel",
                 kernel_param: float | None = None, ntransforms: int = 1, fftshift: bool = False,
                 block_size=4096, evalmode: str = "fast", point_convention: int = 0,
                 use_blocked_spreading: bool = False):
        dtype = np.dtype(dtype)
        self.Z = dtype
        self.is_complex = dtype.kind == "c"
        self.T = np.dtype(np.float32 if dtype in (np.float32, np.complex64) else np.float64)
        self.CT = np.dtype(np.complex64 if self.T == np.float32 else np.complex128)
        self.suf = "_f32" if self.T == np.float32 else "_f64"
        self.ct = C.c_float if self.T == np.float32 else C.c_double
        Ns = (int(Ns),) if np.isscalar(Ns) else tuple(int(n) for n in Ns)
        self.Ns = Ns
        self.D = len(Ns)
        self.M = int(m)
        self.C = int(ntransforms)
        self.fftshift = bool(fftshift)
        self.kernel = kernel
        self.mode = EVALMODES[evalmode]
        self.convention = int(point_convention)
        self.use_blocked = bool(use_blocked_spreading)
        if self.fftshift and not self.is_complex:
            raise ValueError("fftshift=true requires complex data (src/plan.jl:273-280)")
        rt = self.T.type
        self.Nos = oversampled_dims(Ns, sigma, not self.is_complex, rt)
        for n in self.Nos:
            if n < 2 * self.M:                                  # src/plan.jl:545-556
                raise ValueError(f"data size is too small: sigma*N = {n} < {2 * self.M} = 2M")
        self.ks = wavenumbers(Ns, not self.is_complex)
        self.size = tuple(len(k) for k in self.ks)               # Base.size(p), src/plan.jl:426
        L = lib()
        ksz = getattr(L, "orc_kernel_sizeof" + self.suf)()
        self._ksz = ksz
        self._kbuf = C.create_string_buffer(ksz * 3)
        self.phihat = []
        for d in range(self.D):
            sig_d = rt(rt(self.Nos[d]) / rt(Ns[d]))              # src/plan.jl:503
            param = float("nan") if kernel_param is None else float(kernel_param)
            rc = getattr(L, "orc_kernel_init" + self.suf)(
                C.addressof(self._kbuf) + d * ksz, KERNELS[kernel], self.M, self.Nos[d], self.ct(sig_d), param)
            if rc != 0:
                raise ValueError("bad kernel parameters")
            kk = self.ks[d]
            if self.fftshift:
                kk = np.fft.fftshift(kk)                         # src/plan.jl:509-512
            kk = np.ascontiguousarray(kk.astype(self.T))
            out = np.empty(len(kk), dtype=self.T)
            getattr(L, "orc_kernel_fourier" + self.suf)(
                C.c_void_p(C.addressof(self._kbuf) + d * ksz), C.c_int64(len(kk)),
                kk.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
            self.phihat.append(out)
        # spectral oversampled dims (r2c halves the first dim: src/plan.jl:43)
        self.Nos_spec = ((self.Nos[0] // 2 + 1,) + self.Nos[1:]) if not self.is_complex else self.Nos
        self.index_map = [
            non_oversampled_indices(self.size[d], self.Nos_spec[d], (not self.is_complex) and d == 0, self.fftshift)
            for d in range(self.D)
        ]
        if block_size is None:
            self.block_dims = None
        else:
            bd = get_block_dims(self.Nos, block_size)
            self.block_dims = tuple(min(b, n - self.M) for b, n in zip(bd, self.Nos))   # cpu.jl:49-51
        self.points = None
        self.Np = 0
        self.perm = self.cum = self.blockid = None

    # -- helpers -----------------------------------------------------------------------
    def _kptr(self, d=0):
        return C.c_void_p(C.addressof(self._kbuf) + d * self._ksz)

    def kernel_data(self, d=0):
        """(beta, w, dx, tau, cs[(M+4),2M], gcs[M]) of dimension d."""
        M = self.M
        vals = [self.ct() for _ in range(4)]
        cs = np.zeros((M + 4, 2 * M), dtype=self.T)
        gcs = np.zeros(M, dtype=self.T)
        getattr(lib(), "orc_kernel_get" + self.suf)(
            self._kptr(d), *[C.byref(v) for v in vals], cs.ctypes.data_as(C.c_void_p), gcs.ctypes.data_as(C.c_void_p))
        return dict(beta=vals[0].value, w=vals[1].value, dx=vals[2].value, tau=vals[3].value, cs=cs, gcs=gcs)

    def evaluate_kernel(self, x, d=0, mode=None):
        """Kernels.evaluate_kernel(evalmode, g, x) -> (i (1-based), values[2M])."""
        vals = np.empty(2 * self.M, dtype=self.T)
        i = getattr(lib(), "orc_kernel_eval" + self.suf)(
            self._kptr(d), self.mode if mode is None else EVALMODES[mode], self.ct(x), vals.ctypes.data_as(C.c_void_p))
        return int(i), vals

    def _ptrs(self, arrs):
        return (C.c_void_p * len(arrs))(*[a.ctypes.data_as(C.c_void_p) for a in arrs])

    def _i64(self, seq):
        return (C.c_int64 * len(seq))(*[int(s) for s in seq])

    # -- set_points! --------------------------------------------------------------------
    def set_points(self, xp):
        if isinstance(xp, np.ndarray) and xp.ndim == 1:
            xp = (xp,)
        if len(xp) != self.D:
            raise ValueError(f"expected {self.D}-dimensional points")
        pts = []
        for x in xp:
            x = np.asarray(x)
            if x.dtype != self.T:                                 # src/set_points.jl:35
                raise TypeError("input points must have the same accuracy as the created plan")
            pts.append(np.ascontiguousarray(x))
        Np = len(pts[0])
        if any(len(p) != Np for p in pts):                         # src/blocking/cpu.jl:128
            raise ValueError("input points must have the same length along all dimensions")
        self.points, self.Np = pts, Np
        if self.block_dims is not None:
            self.blockid, self.cum, self.perm = self.sort_points(pts, self.block_dims)
        return self

    def sort_points(self, pts, block_dims):
        """Stable counting sort (single-thread reference order).  0-based outputs."""
        Np = len(pts[0])
        nb = [-(-n // b) for n, b in zip(self.Nos, block_dims)]
        nblocks = int(np.prod(nb))
        blockid = np.empty(Np, dtype=np.int32)
        cum = np.empty(nblocks + 1, dtype=np.int32)
        perm = np.empty(Np, dtype=np.int32)
        rc = getattr(lib(), "orc_set_points" + self.suf)(
            C.c_int(self.D), C.c_int64(Np), self._ptrs(pts), self._i64(self.Nos), self._i64(block_dims),
            C.c_int(self.convention), blockid.ctypes.data_as(C.c_void_p), cum.ctypes.data_as(C.c_void_p),
            perm.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return blockid, cum, perm

    # -- exec ---------------------------------------------------------------------------
    def _as_list(self, a):
        return list(a) if isinstance(a, (list, tuple)) else [a]

    def spread(self, vps, nu_weights=None):
        """zero-fill + spread_from_points! -> list of oversampled grids (column-major, dtype Z)."""
        vps = [np.ascontiguousarray(v, dtype=self.Z) for v in self._as_list(vps)]
        assert len(vps) == self.C and all(len(v) == self.Np for v in vps)
        ncells = int(np.prod(self.Nos))
        us = [np.zeros(ncells, dtype=self.Z) for _ in range(self.C)]
        w = None if nu_weights is None else np.ascontiguousarray(nu_weights, dtype=self.T)
        wp = None if w is None else w.ctypes.data_as(C.c_void_p)
        ncomp = 2 if self.is_complex else 1
        L = lib()
        if self.use_blocked and self.block_dims is not None:
            getattr(L, "orc_spread_blocked" + self.suf)(
                C.c_int(self.D), self._i64(self.Nos), self._kptr(0), C.c_int(self.mode), C.c_int64(self.Np),
                self._ptrs(self.points), C.c_int(self.convention), C.c_int(self.C), C.c_int(ncomp),
                self._ptrs(vps), self._ptrs(us), wp, self._i64(self.block_dims),
                self.cum.ctypes.data_as(C.c_void_p), self.perm.ctypes.data_as(C.c_void_p))
        else:
            getattr(L, "orc_spread" + self.suf)(
                C.c_int(self.D), self._i64(self.Nos), self._kptr(0), C.c_int(self.mode), C.c_int64(self.Np),
                self._ptrs(self.points), C.c_int(self.convention), C.c_int(self.C), C.c_int(ncomp),
                self._ptrs(vps), self._ptrs(us), wp)
        return us

    def _fft_forward(self, u):
        """_type1_fft! (src/NonuniformFFTs.jl:197-211): rfft along dim 1 for real data, fft otherwise."""
        a = u.reshape(self.Nos[::-1])          # C-order view of the column-major array
        nw = num_threads()
        if self.is_complex:
            return _sfft.fftn(a, workers=nw)
        return _sfft.rfftn(a, workers=nw)      # halves the last C axis == first Julia dim

    def _fft_backward(self, uh):
        """_type2_fft! (src/NonuniformFFTs.jl:293-314): unnormalised backward transforms."""
        a = uh.reshape(self.Nos_spec[::-1])
        nw = num_threads()
        tot = int(np.prod(self.Nos))
        if self.is_complex:
            return _sfft.ifftn(a, workers=nw, norm="forward")
        return _sfft.irfftn(a, s=self.Nos[::-1], workers=nw, norm="forward")

    def exec_type1(self, vps, nu_weights=None, u_factor=None):
        """exec_type1! — returns list of C arrays shaped size(plan) in Julia (column-major) order,
        i.e. numpy arrays of shape size[::-1] C-contiguous."""
        single = not isinstance(vps, (list, tuple))
        us = self.spread(vps, nu_weights)
        uh = [np.ascontiguousarray(self._fft_forward(u)).reshape(-1) for u in us]
        normfactor = self.T.type(np.prod([2 * np.pi / n for n in self.Nos]))       # :181
        nk = int(np.prod(self.size))
        outs = [np.empty(nk, dtype=self.CT) for _ in range(self.C)]
        f = None if u_factor is None else np.ascontiguousarray(u_factor, dtype=self.T).reshape(-1)
        getattr(lib(), "orc_deconv_type1" + self.suf)(
            C.c_int(self.D), self._i64(self.size), self._i64(self.Nos_spec), self._ptrs(self.index_map),
            self._ptrs(self.phihat), self.ct(normfactor), C.c_int(self.C), self._ptrs(outs), self._ptrs(uh),
            None if f is None else f.ctypes.data_as(C.c_void_p))
        outs = [o.reshape(self.size[::-1]) for o in outs]
        return outs[0] if single else outs

    def deconv_pad(self, uks, u_factor=None):
        uks = [np.ascontiguousarray(u, dtype=self.CT).reshape(-1) for u in self._as_list(uks)]
        assert len(uks) == self.C and all(u.size == int(np.prod(self.size)) for u in uks)
        nspec = int(np.prod(self.Nos_spec))
        uh = [np.empty(nspec, dtype=self.CT) for _ in range(self.C)]
        f = None if u_factor is None else np.ascontiguousarray(u_factor, dtype=self.T).reshape(-1)
        getattr(lib(), "orc_deconv_type2" + self.suf)(
            C.c_int(self.D), self._i64(self.size), self._i64(self.Nos_spec), self._ptrs(self.index_map),
            self._ptrs(self.phihat), C.c_int(self.C), self._ptrs(uh), self._ptrs(uks),
            None if f is None else f.ctypes.data_as(C.c_void_p))
        return uh

    def interp(self, us, nu_weights=None):
        us = [np.ascontiguousarray(u, dtype=self.Z).reshape(-1) for u in us]
        vps = [np.empty(self.Np, dtype=self.Z) for _ in range(self.C)]
        w = None if nu_weights is None else np.ascontiguousarray(nu_weights, dtype=self.T)
        ncomp = 2 if self.is_complex else 1
        wp = None if w is None else w.ctypes.data_as(C.c_void_p)
        if self.use_blocked and self.block_dims is not None:
            getattr(lib(), "orc_interp_blocked" + self.suf)(
                C.c_int(self.D), self._i64(self.Nos), self._kptr(0), C.c_int(self.mode), C.c_int64(self.Np),
                self._ptrs(self.points), C.c_int(self.convention), C.c_int(self.C), C.c_int(ncomp),
                self._ptrs(vps), self._ptrs(us), wp, self._i64(self.block_dims),
                self.cum.ctypes.data_as(C.c_void_p), self.perm.ctypes.data_as(C.c_void_p))
        else:
            getattr(lib(), "orc_interp" + self.suf)(
                C.c_int(self.D), self._i64(self.Nos), self._kptr(0), C.c_int(self.mode), C.c_int64(self.Np),
                self._ptrs(self.points), C.c_int(self.convention), C.c_int(self.C), C.c_int(ncomp),
                self._ptrs(vps), self._ptrs(us), wp)
        return vps

    def exec_type2(self, uks, nu_weights=None, u_factor=None):
        single = not isinstance(uks, (list, tuple))
        uh = self.deconv_pad(uks, u_factor)
        us = [np.ascontiguousarray(self._fft_backward(u)).astype(self.Z, copy=False) for u in uh]
        vps = self.interp(us, nu_weights)
        return vps[0] if single else vps


# ------------------------------------------------------------------------------------------
# exact NUDFT (the "ground truth" the reference's own tests use: test/accuracy.jl:119-125,179-194)
# ------------------------------------------------------------------------------------------

def nudft_type1(ks_list, xs, vp):
    """u(k) = sum_j v_j exp(-i k.x_j), float64/complex128; output in Julia column-major order
    (numpy shape reversed).  Direct summation: small cases only."""
    xs = [np.asarray(x, dtype=np.float64) for x in xs]
    v = np.asarray(vp).astype(np.complex128)
    fac = [np.exp(-1j * np.outer(np.asarray(k, dtype=np.float64), x)) for k, x in zip(ks_list, xs)]
    D = len(xs)
    if D == 1:
        return fac[0] @ v
    if D == 2:
        return np.einsum("bj,aj,j->ba", fac[1], fac[0], v)
    return np.einsum("cj,bj,aj,j->cba", fac[2], fac[1], fac[0], v)


def nudft_type2(ks_list, xs, uk, real_data: bool):
    """v_j = sum_k u(k) exp(+i k.x_j); Hermitian completion (factor 2 for k_1 != 0) for real data."""
    xs = [np.asarray(x, dtype=np.float64) for x in xs]
    D = len(xs)
    u = np.asarray(uk).astype(np.complex128)
    fac = [np.exp(1j * np.outer(np.asarray(k, dtype=np.float64), x)) for k, x in zip(ks_list, xs)]
    if real_data:
        w = np.where(np.asarray(ks_list[0]) == 0, 1.0, 2.0)
        fac[0] = fac[0] * w[:, None]
    if D == 1:
        out = u @ fac[0]
    elif D == 2:
        out = np.einsum("ba,bj,aj->j", u, fac[1], fac[0])
    else:
        out = np.einsum("cba,cj,bj,aj->j", u, fac[2], fac[1], fac[0])
    return out.real if real_data else out
