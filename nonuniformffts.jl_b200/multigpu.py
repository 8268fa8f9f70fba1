"""
Multi-GPU transforms through the C ABI (``nufft_mgpu_*``, csrc/mgpu.cu): one B200 per rank, NCCL over NVLink / NVSwitch.

    MultiGPUPlan(dtype, dims; devices=[0, 1, ...] | group=<torch.distributed group>, strategy="auto", m, sigma, kernel, ...)
    set_points / exec_type1 / exec_type2 as for PlanNUFFT (src/set_points.jl:33-52, src/NonuniformFFTs.jl:148-291)

Two ways to drive it:

* ``devices=[...]``  — ONE process owns all ranks (what a Julia host does through ccall): every argument is a list with one entry
  per device.
* ``group=...``      — one process per GPU (torchrun): each process owns one rank, arguments are that rank's tensors.  The
  NCCL id is created on rank 0 and broadcast through ``torch.distributed`` (any backend).

Strategies (include/nufft_b200.h): ``"slab"`` — z-slab spatial decomposition, uniform data distributed over ranks in y
(`local_block`); ``"points"`` — points partitioned, full grid per rank, partial outputs all-reduced / spectrum broadcast;
``"transforms"`` — independent ntransforms dealt to ranks.  No data is touched on this side.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from .plan import (ArgumentError, Direct, FastApproximation, KaiserBesselKernel, _CPLX, _KERNEL_BY_NAME, _REAL, _check, _to_torch_dtype, build_opts,
                   HalfSupport)


class MultiGPUPlan:
    def __init__(self, dtype, dims, *, devices: Optional[Sequence[int]] = None, group=None, strategy: str = "auto", m=4,
                 sigma: float = 2.0, kernel=None, ntransforms: int = 1, fftshift: bool = False, kernel_evalmode=None,
                 gpu_method: str = "auto", timer: bool = False, point_convention: int = 0):
        self._h = C.c_void_p(None)
        self._lib = _lib.load()
        dtype = _to_torch_dtype(dtype)
        if dtype not in _REAL:
            raise ArgumentError(f"unsupported data type {dtype}")
        dims = (int(dims),) if isinstance(dims, int) else tuple(int(d) for d in dims)
        M = m.M if isinstance(m, HalfSupport) else int(m)
        if kernel is None:
            kernel = KaiserBesselKernel()
        if isinstance(kernel, str):
            kernel = _KERNEL_BY_NAME[kernel]()
        mode = kernel_evalmode if kernel_evalmode is not None else (FastApproximation() if kernel.name == "es" else Direct())
        mode_name = mode if isinstance(mode, str) else mode.name
        if strategy not in _lib.MGPU_STRATEGIES:
            raise ArgumentError("expected strategy in (auto, slab, points, transforms)")
        self.dtype, self.real_dtype = dtype, _REAL[dtype]
        self.complex_dtype = _CPLX[self.real_dtype]
        self.C = int(ntransforms)

        idbuf = (C.c_ubyte * _lib.MGPU_ID_BYTES)()
        if devices is not None:                      # single process, all ranks
            self.devices = [int(d) for d in devices]
            self.nranks = len(self.devices)
            self.ranks = list(range(self.nranks))
            if self.nranks > 1:
                _check(self._lib.nufft_mgpu_unique_id(idbuf))
        else:                                        # one process per GPU
            import torch.distributed as dist
            if not dist.is_initialized():
                raise ArgumentError("pass devices=[...] (single process) or initialise torch.distributed (one process per GPU)")
            self.nranks = dist.get_world_size(group)
            rank = dist.get_rank(group)
            self.ranks = [rank]
            self.devices = [torch.cuda.current_device()]
            if self.nranks > 1:
                if rank == 0:
                    _check(self._lib.nufft_mgpu_unique_id(idbuf))
                backend = dist.get_backend(group)
                t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda" if backend == "nccl" else "cpu")
                dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                idbuf = (C.c_ubyte * _lib.MGPU_ID_BYTES)(*t.cpu().tolist())
        self.nlocal = len(self.ranks)
        stream = torch.cuda.current_stream(torch.device("cuda", self.devices[0])) if self.nlocal == 1 else None
        o = build_opts(self._lib, dims, dtype.is_complex, self.real_dtype, M, sigma, kernel, mode_name, ntransforms, fftshift,
                       False, gpu_method, None, point_convention, None, stream, timer, 0)
        h = C.c_void_p()
        lr = (C.c_int32 * self.nlocal)(*self.ranks)
        dv = (C.c_int32 * self.nlocal)(*self.devices)
        _check(self._lib.nufft_mgpu_create(C.byref(h), C.byref(o), self.nranks, self.nlocal, lr, dv, C.cast(idbuf, C.c_void_p),
                                           _lib.MGPU_STRATEGIES[strategy]))
        self._h = h
        st, nr, nl = C.c_int32(), C.c_int32(), C.c_int32()
        sz, osz = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        _check(self._lib.nufft_mgpu_info(self._h, C.byref(st), C.byref(nr), C.byref(nl), sz, osz))
        self.strategy = {v: k for k, v in _lib.MGPU_STRATEGIES.items()}[int(st.value)]
        D = len(dims)
        self.size = tuple(int(sz[d]) for d in range(D))
        self.oversampled_dims = tuple(int(osz[d]) for d in range(D))
        self._D = D
        self._keep = None

    # ---- geometry --------------------------------------------------------------------------------------------------
    def local_block(self, l: int = 0):
        """(offset, size) in Julia order of the block of the uniform array held by local rank l."""
        off, sz = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        _check(self._lib.nufft_mgpu_local_block(self._h, l, off, sz))
        return tuple(int(off[d]) for d in range(self._D)), tuple(int(sz[d]) for d in range(self._D))

    def local_shape(self, l: int = 0):
        """torch (C-order) shape of the uniform array of local rank l."""
        return self.local_block(l)[1][::-1]

    def stream(self, l: int = 0) -> torch.cuda.ExternalStream:
        s = C.c_void_p()
        _check(self._lib.nufft_mgpu_get_stream(self._h, l, C.byref(s)))
        return torch.cuda.ExternalStream(s.value or 0, device=torch.device("cuda", self.devices[l]))

    # ---- the transforms ------------------------------------------------------------------------------------------------
    def _per_rank(self, a):
        return list(a) if self.nlocal > 1 else [a]

    def _flat(self, per_rank, n: int):
        out = []
        for l, t in enumerate(per_rank):
            ts = list(t) if isinstance(t, (tuple, list)) else [t]
            if len(ts) != n:
                raise ArgumentError(f"expected {n} arrays per rank, got {len(ts)}")
            for x in ts:
                if not x.is_cuda or x.device.index != self.devices[l] or not x.is_contiguous():
                    raise ArgumentError(f"arrays of local rank {l} must be contiguous CUDA tensors on device {self.devices[l]}")
            out += ts
        arr = (C.c_void_p * len(out))(*[t.data_ptr() for t in out])
        return arr, out

    def set_points(self, points) -> "MultiGPUPlan":
        """points: tuple of D vectors of THIS rank (one process per GPU), or a list of such tuples, one per device."""
        pr = self._per_rank(points)
        arr, keep = self._flat(pr, self._D)
        for t in keep:
            if t.dtype != self.real_dtype:
                raise ArgumentError("input points must have the same accuracy as the created plan")
        # the ABI takes 3 pointers per rank
        if self._D != 3:
            padded = []
            for l in range(self.nlocal):
                padded += [keep[l * self._D + d].data_ptr() for d in range(self._D)] + [None] * (3 - self._D)
            arr = (C.c_void_p * len(padded))(*padded)
        nps = (C.c_int64 * self.nlocal)(*[int(keep[l * self._D].numel()) for l in range(self.nlocal)])
        self.Np = [int(n) for n in nps]
        self._keep = keep
        self._sync_inputs()
        _check(self._lib.nufft_mgpu_set_points(self._h, nps, arr))
        return self

    def _sync_inputs(self):
        # single process, several devices: the library runs on its own streams; make them wait for the producers of the inputs
        if self.nlocal > 1:
            for d in self.devices:
                torch.cuda.synchronize(d)

    def exec_type1(self, uhat, vp):
        ua, uk = self._flat(self._per_rank(uhat), self.C)
        va, vk = self._flat(self._per_rank(vp), self.C)
        self._sync_inputs()
        _check(self._lib.nufft_mgpu_exec_type1(self._h, ua, va, None))
        return uhat

    def exec_type2(self, vp, uhat):
        ua, uk = self._flat(self._per_rank(uhat), self.C)
        va, vk = self._flat(self._per_rank(vp), self.C)
        self._sync_inputs()
        _check(self._lib.nufft_mgpu_exec_type2(self._h, va, ua, None))
        return vp

    def gather_output(self, local):
        """Full size(p) arrays (one per local rank) from the distributed output of the slab strategy."""
        pr = self._per_rank(local)
        full = [torch.empty(self.size[::-1], dtype=self.complex_dtype, device=torch.device("cuda", self.devices[l])) for l in range(self.nlocal)]
        la, _ = self._flat(pr, 1)
        fa, _ = self._flat(full, 1)
        _check(self._lib.nufft_mgpu_gather_output(self._h, fa, la))
        self.synchronize()
        return full if self.nlocal > 1 else full[0]

    def synchronize(self) -> None:
        _check(self._lib.nufft_mgpu_synchronize(self._h))

    @property
    def exchange(self) -> str:
        """How the z-slab exchanges travel: "peer-windows" (direct copies into the receiver's buffers) or "nccl"."""
        mode = C.c_int32(0)
        _check(self._lib.nufft_mgpu_exchange_mode(self._h, C.byref(mode)))
        return "peer-windows" if mode.value == 1 else "nccl"

    def timings(self, l: int = 0) -> dict:
        ms = (C.c_float * 16)()
        _check(self._lib.nufft_mgpu_get_timings(self._h, l, ms))
        names = ["point exchange", "local set_points", "T1 value exchange", "T1 zero fill + spreading", "T1 halo exchange + add",
                 "T1 FFT passes x, y", "T1 transpose", "T1 FFT pass z", "T2 FFT pass z", "T2 transpose", "T2 FFT passes y, x",
                 "T2 halo exchange", "T2 interpolation", "T2 value return"]
        return {n: float(ms[i]) for i, n in enumerate(names)}

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nufft_mgpu_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
