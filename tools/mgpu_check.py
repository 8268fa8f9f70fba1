"""Multi-GPU transforms (nufft_mgpu_*, csrc/mgpu.cu) against the single-GPU plan on the same inputs, in ONE process that owns all
GPUs of the box; optional timing of a large problem.

    python tools/mgpu_check.py --gpus 2 [--modes 64] [--np 200000] [--time-modes 512 --time-np 134217728]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import nufft_b200 as nb  # noqa: E402


def l2(a, b):
    a = a.astype(np.complex128).ravel(); b = b.astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def check(G, modes, npts, strategy, dist="uniform", C=1, dtype=torch.complex64):
    rng = np.random.default_rng(7)
    rt = np.float32 if dtype in (torch.complex64, torch.float32) else np.float64
    if dist == "uniform":
        xs = [(rng.random(npts) * 2 * np.pi).astype(rt) for _ in range(3)]
    else:
        xs = [rng.standard_normal(npts).astype(rt) for _ in range(3)]
    vps = [(rng.standard_normal(npts) + 1j * rng.standard_normal(npts)).astype(np.complex64 if rt == np.float32 else np.complex128) for _ in range(C)]
    uks = [(rng.standard_normal((modes,) * 3) + 1j * rng.standard_normal((modes,) * 3)).astype(vps[0].dtype) for _ in range(C)]
    kw = dict(m=4, sigma=2.0, kernel=nb.BackwardsKaiserBesselKernel(), kernel_evalmode=nb.FastApproximation(), ntransforms=C)
    # single-GPU reference
    d0 = torch.device("cuda", 0)
    p1 = nb.PlanNUFFT(dtype, (modes,) * 3, device=d0, **kw)
    p1.set_points(tuple(torch.from_numpy(x).to(d0) for x in xs))
    ref1 = [torch.empty(p1.shape, dtype=p1.complex_dtype, device=d0) for _ in range(C)]
    p1.exec_type1(ref1 if C > 1 else ref1[0], [torch.from_numpy(v).to(d0) for v in vps] if C > 1 else torch.from_numpy(vps[0]).to(d0))
    ref2 = [torch.empty(npts, dtype=dtype, device=d0) for _ in range(C)]
    p1.exec_type2(ref2 if C > 1 else ref2[0], [torch.from_numpy(u).to(d0) for u in uks] if C > 1 else torch.from_numpy(uks[0]).to(d0))
    torch.cuda.synchronize()
    ref1 = [r.cpu().numpy() for r in ref1]; ref2 = [r.cpu().numpy() for r in ref2]
    p1.close()

    mp = nb.MultiGPUPlan(dtype, (modes,) * 3, devices=list(range(G)), strategy=strategy, **kw)
    devs = [torch.device("cuda", g) for g in range(G)]
    if mp.strategy == "transforms":
        parts = [(0, npts)] * G
    else:
        parts = [nb.partition_points(npts, G, g) for g in range(G)]
    pts = [tuple(torch.from_numpy(x[a:b]).to(devs[g]) for x in xs) for g, (a, b) in enumerate(parts)]
    mp.set_points(pts if G > 1 else pts[0])
    vloc = [[torch.from_numpy(v[a:b]).to(devs[g]) for v in vps] for g, (a, b) in enumerate(parts)]
    out = [[torch.zeros(mp.local_shape(g), dtype=mp.complex_dtype, device=devs[g]) for _ in range(C)] for g in range(G)]
    mp.exec_type1([o if C > 1 else o[0] for o in out] if G > 1 else (out[0] if C > 1 else out[0][0]),
                  [v if C > 1 else v[0] for v in vloc] if G > 1 else (vloc[0] if C > 1 else vloc[0][0]))
    mp.synchronize()
    res = {"strategy": mp.strategy, "G": G, "modes": modes, "np": npts, "dist": dist, "C": C, "exchange": mp.exchange}
    if mp.strategy == "slab":
        full = mp.gather_output([o[0] for o in out] if G > 1 else out[0][0])
        full = full if G > 1 else [full]
        res["type1_err"] = max(l2(f.cpu().numpy(), ref1[0]) for f in full)
    elif mp.strategy == "points":
        res["type1_err"] = max(l2(out[g][c].cpu().numpy(), ref1[c]) for g in range(G) for c in range(C))
    else:
        res["type1_err"] = max(l2(out[c % G][c].cpu().numpy(), ref1[c]) for c in range(C))
    # type 2
    if mp.strategy == "slab":
        u_loc = []
        for g in range(G):
            off, sz = mp.local_block(g)
            u_loc.append([torch.from_numpy(np.ascontiguousarray(uks[0][:, off[1]:off[1] + sz[1], :])).to(devs[g])])
    elif mp.strategy == "points":       # rank 0 holds the spectrum, the others receive it
        u_loc = [[torch.from_numpy(u).to(devs[g]) if g == 0 else torch.zeros((modes,) * 3, dtype=mp.complex_dtype, device=devs[g]) for u in uks] for g in range(G)]
    else:
        u_loc = [[torch.from_numpy(u).to(devs[g]) for u in uks] for g in range(G)]
    v_out = [[torch.zeros(b - a, dtype=dtype, device=devs[g]) for _ in range(C)] for g, (a, b) in enumerate(parts)]
    mp.exec_type2([v if C > 1 else v[0] for v in v_out] if G > 1 else (v_out[0] if C > 1 else v_out[0][0]),
                  [u if C > 1 else u[0] for u in u_loc] if G > 1 else (u_loc[0] if C > 1 else u_loc[0][0]))
    mp.synchronize()
    if mp.strategy == "transforms":
        res["type2_err"] = max(l2(v_out[c % G][c].cpu().numpy(), ref2[c]) for c in range(C))
    else:
        res["type2_err"] = max(l2(np.concatenate([v_out[g][c].cpu().numpy() for g in range(G)]), ref2[c]) for c in range(C))
    mp.close()
    tol = 2e-5 if rt == np.float32 else 1e-11
    res["ok"] = bool(res["type1_err"] <= tol and res["type2_err"] <= tol)
    print(json.dumps(res), flush=True)
    return res["ok"]


def timing(G, modes, npts, iters):
    """Strong scaling: ONE problem (modes^3, npts points in total), points dealt evenly and randomly to the G ranks."""
    kw = dict(m=4, sigma=2.0, kernel=nb.BackwardsKaiserBesselKernel(), kernel_evalmode=nb.FastApproximation())
    devs = [torch.device("cuda", g) for g in range(G)]
    n_loc = npts // G
    pts, vps, outs, v_out = [], [], [], []
    if G == 1:
        plan = nb.PlanNUFFT(torch.complex64, (modes,) * 3, device=devs[0], timer=True, **kw)
    else:
        plan = nb.MultiGPUPlan(torch.complex64, (modes,) * 3, devices=list(range(G)), strategy="slab", timer=True, **kw)
    for g in range(G):
        with torch.cuda.device(devs[g]):
            gen = torch.Generator(device=devs[g]); gen.manual_seed(100 + g)
            pts.append(tuple(torch.rand(n_loc, device=devs[g], generator=gen) * (2 * np.pi) for _ in range(3)))
            vps.append(torch.randn(n_loc, 2, device=devs[g], generator=gen).view(torch.complex64).reshape(n_loc) if False else
                       torch.view_as_complex(torch.randn(n_loc, 2, device=devs[g], generator=gen)))
            shp = plan.local_shape(g) if G > 1 else plan.shape
            outs.append(torch.zeros(shp, dtype=torch.complex64, device=devs[g]))
            v_out.append(torch.zeros(n_loc, dtype=torch.complex64, device=devs[g]))
    for d in devs:
        torch.cuda.synchronize(d)

    def step():
        if G == 1:
            plan.set_points(pts[0]); plan.exec_type1(outs[0], vps[0]); plan.set_points(pts[0]); plan.exec_type2(v_out[0], outs[0])
            torch.cuda.synchronize()
        else:
            plan.set_points(pts); plan.exec_type1(outs, vps); plan.set_points(pts); plan.exec_type2(v_out, outs)
            plan.synchronize()
    for _ in range(2):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    ms = (time.perf_counter() - t0) / iters * 1e3
    stages = plan.timings(0) if G > 1 else dict(plan.timer)
    print(json.dumps({"G": G, "modes": modes, "np_total": n_loc * G, "ms_per_step": ms, "points_per_s": 2.0 * n_loc * G / (ms * 1e-3),
                      "exchange": plan.exchange if G > 1 else None,
                      "stage_ms_rank0": {k: round(v, 3) for k, v in stages.items()}}), flush=True)
    plan.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--modes", type=int, default=64)
    ap.add_argument("--np", type=int, default=300000)
    ap.add_argument("--time-modes", type=int, default=0)
    ap.add_argument("--time-np", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--skip-single", action="store_true", help="time the multi-GPU plan only")
    a = ap.parse_args()
    ok = True
    if a.modes > 0:
        ok &= check(a.gpus, a.modes, a.np, "slab")
        ok &= check(a.gpus, a.modes, a.np, "slab", dist="clustered")
        ok &= check(a.gpus, a.modes, a.np // 4, "points")
        ok &= check(a.gpus, 32, 20000, "points", dtype=torch.complex128)
        ok &= check(a.gpus, 32, 20000, "transforms", C=3, dtype=torch.complex128)
    if a.time_modes:
        for g in ([a.gpus] if a.skip_single else sorted({1, a.gpus})):
            timing(g, a.time_modes, a.time_np, a.iters)
    print("MGPU_CHECK", "ok" if ok else "FAILED")
    sys.exit(0 if ok else 1)
