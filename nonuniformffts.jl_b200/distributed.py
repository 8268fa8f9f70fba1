"""
Multi-GPU host layer: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch) for the plumbing.

The reference has no distributed code; this is the sharding the hot path offers naturally
(DESIGN.md §6):

* ``shard_transforms``       independent `ntransforms` components are dealt round-robin to ranks (no data-path collective);
* ``PointPartitionedNUFFT``  type-1: every rank transforms its own points and the partial NON-OVERSAMPLED outputs are
                             summed (all-reduce / reduce of prod(size(p)) complex values — FFT and deconvolution are
                             linear, so the reduction commutes with them and moves 1/sigma^D of the grid bytes);
                             type-2: the spectrum is broadcast, every rank interpolates at its own points;
* ``reduce_grid=True``       the north-star variant: partial OVERSAMPLED grids are reduced onto one rank
                             (`nufft_type1_spread` -> reduce -> `nufft_type1_finish` on the root).

The executor only needs ``set_points / exec_type1 / exec_type2`` (and, for ``reduce_grid``, ``type1_spread /
type1_finish / grid``): `PlanNUFFT` on GPUs; the gloo CPU tests plug in an oracle-backed stand-in to check the
host logic.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def partition_points(np_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of np_total points: rank r gets [start, stop)."""
    base, rem = divmod(int(np_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_transforms(ntransforms: int, world: int, rank: int) -> List[int]:
    """Components {c : c mod world == rank} (independent transforms: no collective on the data path)."""
    return [c for c in range(int(ntransforms)) if c % world == rank]


class PointPartitionedNUFFT:
    """Type-1 / type-2 transforms whose points are partitioned over the ranks of `group`."""

    def __init__(self, plan, group=None, reduce_grid: bool = False):
        self.plan = plan
        self.group = group
        self.reduce_grid = bool(reduce_grid)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def set_points(self, local_points):
        """Each rank passes ITS points (any partition; use `partition_points` for a balanced contiguous one)."""
        self.plan.set_points(local_points)
        return self

    def exec_type1(self, us, vp_local, callbacks=None, dst: Optional[int] = None):
        """us <- sum over ranks of type-1(local points).  dst=None: all ranks get the result (all-reduce);
        otherwise only rank `dst` does (reduce)."""
        single = not isinstance(us, (tuple, list))
        if self.reduce_grid and self.world > 1:
            root = 0 if dst is None else dst
            self.plan.type1_spread(vp_local, callbacks)
            grid = self.plan.grid()
            dist.reduce(grid, dst=root, op=dist.ReduceOp.SUM, group=self.group)
            if self.rank == root:
                self.plan.type1_finish(us, callbacks)
            if dst is None:
                for u in ([us] if single else us):
                    dist.broadcast(u, src=root, group=self.group)
            return us
        self.plan.exec_type1(us, vp_local, callbacks)
        if self.world > 1:
            for u in ([us] if single else us):
                if dst is None:
                    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=self.group)
                else:
                    dist.reduce(u, dst=dst, op=dist.ReduceOp.SUM, group=self.group)
        return us

    def exec_type2(self, vp_local, us, callbacks=None, src: int = 0):
        """Broadcast the spectrum held by rank `src`, interpolate at the local points."""
        if self.world > 1:
            for u in ([us] if not isinstance(us, (tuple, list)) else us):
                dist.broadcast(u, src=src, group=self.group)
        return self.plan.exec_type2(vp_local, us, callbacks)


class TransformShardedNUFFT:
    """`ntransforms` independent transforms dealt to ranks: rank r owns components c with c % world == r and runs the
    single-GPU pipeline on them with a plan created for ``len(shard_transforms(C, world, rank))`` transforms."""

    def __init__(self, make_plan, ntransforms: int, group=None):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.components = shard_transforms(ntransforms, self.world, self.rank)
        self.plan = make_plan(len(self.components)) if self.components else None

    def set_points(self, points):
        if self.plan is not None:
            self.plan.set_points(points)
        return self

    def exec_type1(self, us_all: Sequence, vp_all: Sequence, callbacks=None):
        """us_all / vp_all: the full tuples (length ntransforms); only this rank's components are touched."""
        if self.plan is not None:
            us = [us_all[c] for c in self.components]
            vp = [vp_all[c] for c in self.components]
            self.plan.exec_type1(us if len(us) > 1 else us[0], vp if len(vp) > 1 else vp[0], callbacks)
        return self.components

    def exec_type2(self, vp_all: Sequence, us_all: Sequence, callbacks=None):
        if self.plan is not None:
            us = [us_all[c] for c in self.components]
            vp = [vp_all[c] for c in self.components]
            self.plan.exec_type2(vp if len(vp) > 1 else vp[0], us if len(us) > 1 else us[0], callbacks)
        return self.components
