"""CPU tests of the multi-GPU HOST logic with world_size = 2 over gloo.

The product executor needs a GPU, so the ranks plug an oracle-backed stand-in (same set_points / exec_type1 /
exec_type2 surface, CPU tensors) into the distributed layer; what is verified is the sharding arithmetic and that
partition + collective reproduces the single-process transform."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


class OracleExecutor:
    """Stand-in for PlanNUFFT on CPU tensors, backed by the oracle (tests only)."""

    def __init__(self, dtype, dims, ntransforms=1, **kw):
        import oracle
        self.p = oracle.OraclePlan(dtype, dims, ntransforms=ntransforms, block_size=None, **kw)
        self.C = ntransforms

    def set_points(self, xs):
        self.p.set_points([x.numpy() for x in xs])

    def exec_type1(self, us, vp, callbacks=None):
        multi = isinstance(us, (tuple, list))
        out = self.p.exec_type1([v.numpy() for v in vp] if multi else vp.numpy())
        for u, o in zip(us if multi else [us], out if multi else [out]):
            u.copy_(torch.from_numpy(np.ascontiguousarray(o)))
        return us

    def exec_type2(self, vp, us, callbacks=None):
        multi = isinstance(us, (tuple, list))
        out = self.p.exec_type2([u.numpy() for u in us] if multi else us.numpy())
        for v, o in zip(vp if multi else [vp], out if multi else [out]):
            v.copy_(torch.from_numpy(np.ascontiguousarray(o)))
        return vp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("nufft_dist", ROOT / "nonuniformffts.jl_b200" / "distributed.py")
        nd = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(nd)
        rng = np.random.default_rng(123)                     # same stream on every rank
        dims, Np = (12, 10, 8), 1001
        xs = [rng.random(Np) * 2 * np.pi for _ in dims]
        vp = rng.standard_normal(Np) + 1j * rng.standard_normal(Np)
        full = OracleExecutor(np.complex128, dims)
        full.set_points([torch.from_numpy(x) for x in xs])
        ref1 = torch.empty(dims[::-1], dtype=torch.complex128)
        full.exec_type1(ref1, torch.from_numpy(vp))
        ref2 = torch.empty(Np, dtype=torch.complex128)
        full.exec_type2(ref2, ref1.clone())
        # --- point partition: type 1 all-reduce, type 2 broadcast + split
        s, e = nd.partition_points(Np, world, rank)
        pp = nd.PointPartitionedNUFFT(OracleExecutor(np.complex128, dims))
        pp.set_points([torch.from_numpy(x[s:e].copy()) for x in xs])
        out1 = torch.empty(dims[::-1], dtype=torch.complex128)
        pp.exec_type1(out1, torch.from_numpy(vp[s:e].copy()))
        err1 = float((out1 - ref1).abs().max() / ref1.abs().max())
        spec_in = ref1.clone() if rank == 0 else torch.zeros_like(ref1)    # only rank 0 holds the spectrum
        out2 = torch.empty(e - s, dtype=torch.complex128)
        pp.exec_type2(out2, spec_in, src=0)
        err2 = float((out2 - ref2[s:e]).abs().max() / ref2.abs().max())
        # --- reduce to one rank only
        out1b = torch.zeros(dims[::-1], dtype=torch.complex128)
        pp.exec_type1(out1b, torch.from_numpy(vp[s:e].copy()), dst=1)
        err1b = float((out1b - ref1).abs().max() / ref1.abs().max()) if rank == 1 else 0.0
        # --- ntransforms sharding (C = 3 over 2 ranks): no collective
        C = 3
        vps = [torch.from_numpy(rng.standard_normal(Np) + 1j * rng.standard_normal(Np)) for _ in range(C)]
        ts = nd.TransformShardedNUFFT(lambda c: OracleExecutor(np.complex128, dims, ntransforms=c), C)
        ts.set_points([torch.from_numpy(x) for x in xs])
        outs = [torch.zeros(dims[::-1], dtype=torch.complex128) for _ in range(C)]
        mine = ts.exec_type1(outs, vps)
        errs = []
        for c in mine:
            r = torch.empty(dims[::-1], dtype=torch.complex128)
            full.exec_type1(r, vps[c])
            errs.append(float((outs[c] - r).abs().max() / r.abs().max()))
        q.put((rank, err1, err2, err1b, mine, max(errs) if errs else 0.0))
    finally:
        dist.destroy_process_group()


def test_partition_helpers():
    import importlib.util
    spec = importlib.util.spec_from_file_location("nufft_dist", ROOT / "nonuniformffts.jl_b200" / "distributed.py")
    nd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nd)
    for n in (0, 1, 7, 1000, 2 ** 24 + 3):
        for w in (1, 2, 3, 8):
            parts = [nd.partition_points(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert nd.shard_transforms(3, 2, 0) == [0, 2] and nd.shard_transforms(3, 2, 1) == [1]
    assert nd.shard_transforms(1, 4, 3) == []


def test_point_partition_and_transform_sharding_world2():
    import oracle
    oracle.build()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    comps = []
    for rank, e1, e2, e1b, mine, ec in res:
        assert e1 < 1e-12, f"rank {rank}: type-1 all-reduce mismatch {e1}"
        assert e2 < 1e-12, f"rank {rank}: type-2 broadcast/split mismatch {e2}"
        assert e1b < 1e-12 and ec < 1e-12
        comps += mine
    assert sorted(comps) == [0, 1, 2]
