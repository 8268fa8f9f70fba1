"""nonuniformffts.jl_b200 — B200-native NUFFT backend behind the NonuniformFFTs.jl API.

The directory name contains a dot, so import it through the repo-root shim: ``import nufft_b200``.
Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + C ABI), the ctypes binding
(``_lib``), the host-side mirror of the reference interface (``plan``), the multi-GPU plans (``multigpu``: the C ABI's
``nufft_mgpu_*``; ``distributed``: the same sharding over ``torch.distributed``) and the Julia ccall shim (``julia/``).
"""
from .plan import (ArgumentError, BackwardsKaiserBesselKernel, BSplineKernel, DimensionMismatch, Direct, ESKernel,
                   FastApproximation, GaussianKernel, HalfSupport, KaiserBesselKernel, NUFFTCallbacks, NUFFTError,
                   PlanNUFFT, exec_type1, exec_type2, kernel_tables, launch_count, set_points)

from .nfft import NFFTPlan, accuracy_params, mul, mul_adjoint, nodes, plan_nfft
from .distributed import PointPartitionedNUFFT, TransformShardedNUFFT, partition_points, shard_transforms
from .multigpu import MultiGPUPlan

__all__ = [
    "PointPartitionedNUFFT", "TransformShardedNUFFT", "partition_points", "shard_transforms", "MultiGPUPlan",
    "NFFTPlan", "plan_nfft", "nodes", "mul", "mul_adjoint", "accuracy_params",
    "PlanNUFFT", "set_points", "exec_type1", "exec_type2", "HalfSupport", "NUFFTCallbacks",
    "KaiserBesselKernel", "BackwardsKaiserBesselKernel", "GaussianKernel", "BSplineKernel", "ESKernel",
    "Direct", "FastApproximation", "ArgumentError", "DimensionMismatch", "NUFFTError", "launch_count", "kernel_tables",
]
