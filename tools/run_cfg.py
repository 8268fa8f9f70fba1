"""Stage timings of the other BASELINE.json configurations on one GPU:
   C4: 3-D 256^3 modes, ntransforms = 3, Float64 real data, HalfSupport(8), KaiserBessel, clustered points (wrapped normal);
   C5 (one GPU's share and the full problem): 3-D 512^3 modes, ComplexF32, HalfSupport(4)."""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import nufft_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="c4")
ap.add_argument("--np", type=int, default=1 << 24)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--fast", action="store_true", help="FastApproximation instead of the reference's GPU default (Direct)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(4)
if a.cfg == "c4":
    N, C, npts = 256, 3, a.np
    xs = [torch.randn(npts, generator=g, device=dev, dtype=torch.float64) for _ in range(3)]
    vp = [torch.randn(npts, generator=g, device=dev, dtype=torch.float64) for _ in range(C)]
    plan = nb.PlanNUFFT(torch.float64, (N,) * 3, m=8, sigma=2.0, kernel=nb.KaiserBesselKernel(), ntransforms=C, timer=True,
                        kernel_evalmode=nb.FastApproximation() if a.fast else nb.Direct())
    out1 = [torch.empty((N, N, N // 2 + 1), dtype=torch.complex128, device=dev) for _ in range(C)]
    out2 = [torch.empty(npts, dtype=torch.float64, device=dev) for _ in range(C)]
else:
    N, C, npts = 512, 1, a.np
    xs = [torch.rand(npts, generator=g, device=dev, dtype=torch.float32) * np.float32(2 * np.pi) for _ in range(3)]
    for x in xs:
        x[x >= np.float32(2 * np.pi)] = 0.0
    vp = [torch.complex(torch.randn(npts, generator=g, device=dev), torch.randn(npts, generator=g, device=dev))]
    plan = nb.PlanNUFFT(torch.complex64, (N,) * 3, m=4, sigma=2.0, kernel=nb.BackwardsKaiserBesselKernel(),
                        kernel_evalmode=nb.FastApproximation(), timer=True)
    out1 = [torch.empty((N, N, N), dtype=torch.complex64, device=dev)]
    out2 = [torch.empty(npts, dtype=torch.complex64, device=dev)]
print(repr(plan).splitlines()[-3:])
for it in range(a.iters):
    plan.set_points(tuple(xs))
    plan.exec_type1(out1 if C > 1 else out1[0], vp if C > 1 else vp[0])
    plan.exec_type2(out2 if C > 1 else out2[0], out1 if C > 1 else out1[0])
    torch.cuda.synchronize()
    print(a.cfg, npts, {k: round(v, 3) for k, v in plan.timer.items()})
