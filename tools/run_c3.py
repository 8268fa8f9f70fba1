"""Run the C3 workload (256^3 modes, Np = 2^24, ComplexF32, M = 4, sigma = 2) a few times — used under ncu."""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import nufft_b200 as nb  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--np", type=int, default=bench.NP_FULL)
ap.add_argument("--modes", type=int, default=256)
ap.add_argument("--m", type=int, default=4)
ap.add_argument("--sigma", type=float, default=2.0)
ap.add_argument("--method", default="auto")
ap.add_argument("--dist", default="uniform")
ap.add_argument("--block", default="")
ap.add_argument("--real", action="store_true", help="Float32 real non-uniform data (r2c plan) instead of ComplexF32")
a = ap.parse_args()
bench.N_MODES = a.modes
xs, vp, uk = bench.make_inputs(3, a.np)
if a.dist == "clustered":
    rng = np.random.default_rng(4)
    xs = [rng.standard_normal(a.np).astype(np.float32) for _ in range(3)]
dev = torch.device("cuda", 0)
xs_d = [torch.from_numpy(x).to(dev) for x in xs]
vp_d, uk_d = torch.from_numpy(vp).to(dev), torch.from_numpy(uk).to(dev)
vdt = torch.complex64
if a.real:
    vdt = torch.float32
    vp_d = vp_d.real.contiguous()
plan = nb.PlanNUFFT(vdt, (a.modes,) * 3, m=a.m, sigma=a.sigma, kernel=nb.BackwardsKaiserBesselKernel(),
                    kernel_evalmode=nb.FastApproximation(), timer=True, gpu_method=a.method,
                    block_size=tuple(int(b) for b in a.block.split(',')) if a.block else None)
print(repr(plan).splitlines()[-3:])
out1 = torch.empty(plan.shape, dtype=torch.complex64, device=dev)
out2 = torch.empty(a.np, dtype=vdt, device=dev)
if a.real:
    uk_d = (torch.randn(out1.shape, device=dev) + 1j * torch.randn(out1.shape, device=dev)).to(torch.complex64)
for it in range(a.iters):
    plan.set_points(tuple(xs_d))
    plan.exec_type1(out1, vp_d)
    plan.exec_type2(out2, uk_d)
    torch.cuda.synchronize()
    print({k: round(v, 3) for k, v in plan.timer.items()})
