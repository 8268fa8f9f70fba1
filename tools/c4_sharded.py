"""BASELINE config C4 (3-D 256^3 modes, ntransforms = 3, Float64 real data, HalfSupport(8), KaiserBessel + Direct, clustered
points) on one GPU and with its three transforms dealt to the GPUs of the box (NUFFT_MGPU_TRANSFORMS: component c on rank
c mod G, no data-path collective).  Prints ms per step (set_points + type 1 + type 2, all three components) and the deviation of
the sharded results from the single-GPU ones."""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import nufft_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--np", type=int, default=1 << 24)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
N, C, npts, G = 256, 3, a.np, a.gpus
kw = dict(m=8, sigma=2.0, kernel=nb.KaiserBesselKernel(), ntransforms=C, kernel_evalmode=nb.Direct())
g = torch.Generator(device="cuda:0").manual_seed(4)
xs0 = [torch.randn(npts, generator=g, device="cuda:0", dtype=torch.float64) for _ in range(3)]
vp0 = [torch.randn(npts, generator=g, device="cuda:0", dtype=torch.float64) for _ in range(C)]
shape = (N, N, N // 2 + 1)


def run(step, sync, iters):
    step(); sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    sync()
    return (time.perf_counter() - t0) / iters * 1e3


plan = nb.PlanNUFFT(torch.float64, (N,) * 3, **kw)
o1 = [torch.empty(shape, dtype=torch.complex128, device="cuda:0") for _ in range(C)]
w1 = [torch.empty(npts, dtype=torch.float64, device="cuda:0") for _ in range(C)]
def step1():
    plan.set_points(tuple(xs0)); plan.exec_type1(o1, vp0); plan.exec_type2(w1, o1)
ms1 = run(step1, torch.cuda.synchronize, a.iters)
plan.close()

mp = nb.MultiGPUPlan(torch.float64, (N,) * 3, devices=list(range(G)), strategy="transforms", **kw)
xs = [tuple(x.to(f"cuda:{d}") for x in xs0) for d in range(G)]
vp = [[v.to(f"cuda:{d}") for v in vp0] for d in range(G)]
oG = [[torch.zeros(shape, dtype=torch.complex128, device=f"cuda:{d}") for _ in range(C)] for d in range(G)]
wG = [[torch.zeros(npts, dtype=torch.float64, device=f"cuda:{d}") for _ in range(C)] for d in range(G)]
for d in range(G):
    torch.cuda.synchronize(d)
def stepG():
    mp.set_points(xs); mp.exec_type1(oG, vp); mp.exec_type2(wG, oG)
msG = run(stepG, mp.synchronize, a.iters)
rel = lambda x, y: float(torch.linalg.vector_norm(x - y) / torch.linalg.vector_norm(y))
e1 = max(rel(oG[c % G][c].to("cuda:0"), o1[c]) for c in range(C))
e2 = max(rel(wG[c % G][c].to("cuda:0"), w1[c]) for c in range(C))
mp.close()
print(json.dumps({"config": "C4: 256^3 modes, ntransforms=3, Float64 real, HalfSupport(8), KaiserBessel+Direct, clustered (wrapped normal)",
                  "np": npts, "one_gpu_ms_per_step": ms1, "gpus": G, "strategy": "transforms", "sharded_ms_per_step": msG,
                  "speedup": ms1 / msG, "type1_rel_l2_vs_one_gpu": e1, "type2_rel_l2_vs_one_gpu": e2}), flush=True)
