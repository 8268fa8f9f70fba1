"""The reference's own benchmark (benchmark/CPU+CUDA/run_benchmarks.jl) on this backend, written in the reference's
5-column .dat format (run_benchmarks.jl:279-297) so that the curves can be overlaid on its committed H100 / MI300A data:

    3-D, 256^3 modes, sigma = 1.5, HalfSupport(4), KaiserBessel kernel, Direct evaluation (the reference's GPU default),
    points ~ randn (wrapped normal, run_benchmarks.jl:47-51), Np = rho * 256^3 for rho = 10^(-4 : 0.5 : 1),
    one sample = set_points! + exec_typeX! + synchronise (run_benchmarks.jl:82-92), median over the samples,
    relative errors against a HalfSupport(8), sigma = 2 plan (run_benchmarks.jl:63-76).

    python tools/run_benchmarks.py [--types ComplexF64 Float64 ComplexF32] [--methods shared_memory global_memory] [--out DIR]
"""
import argparse
import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import nufft_b200 as nb  # noqa: E402

TYPES = {"ComplexF64": torch.complex128, "Float64": torch.float64, "ComplexF32": torch.complex64, "Float32": torch.float32}
N = 256


def bench_one(Z, Np, method, sigma, m, samples, fast):
    dev = torch.device("cuda", 0)
    real = Z in (torch.float64, torch.float32)
    T = {torch.complex128: torch.float64, torch.float64: torch.float64, torch.complex64: torch.float32, torch.float32: torch.float32}[Z]
    CZ = torch.complex128 if T == torch.float64 else torch.complex64
    g = torch.Generator(device="cuda").manual_seed(42)
    xp = tuple(torch.randn(Np, generator=g, device=dev, dtype=T) for _ in range(3))
    vp = torch.randn(Np, generator=g, device=dev, dtype=T) if real else \
        torch.complex(torch.randn(Np, generator=g, device=dev, dtype=T), torch.randn(Np, generator=g, device=dev, dtype=T))
    wp = torch.empty_like(vp)
    kern = nb.BackwardsKaiserBesselKernel() if fast else nb.KaiserBesselKernel()
    mode = nb.FastApproximation() if fast else nb.Direct()
    p = nb.PlanNUFFT(Z, (N, N, N), m=m, sigma=sigma, kernel=kern, kernel_evalmode=mode, gpu_method=method)
    us = torch.empty(p.shape, dtype=CZ, device=dev)
    p.set_points(xp); p.exec_type1(us, vp); p.exec_type2(wp, us)
    # accuracy against a high-accuracy plan (run_benchmarks.jl:63-76)
    # (always in Float64: unnormalised Float32 KB values overflow at HalfSupport(8) in 3-D, in the reference too)
    Zr = torch.float64 if real else torch.complex128
    pr = nb.PlanNUFFT(Zr, (N, N, N), m=8, sigma=2.0, kernel=nb.KaiserBesselKernel(), kernel_evalmode=nb.Direct(), gpu_method="global_memory")
    ur = torch.empty(pr.shape, dtype=torch.complex128, device=dev)
    wr = torch.empty(Np, dtype=Zr, device=dev)
    pr.set_points(tuple(x.double() for x in xp)); pr.exec_type1(ur, vp.to(Zr)); pr.exec_type2(wr, ur)
    e1 = float(torch.linalg.vector_norm(ur - us.to(torch.complex128)) / torch.linalg.vector_norm(ur))
    pr.exec_type2(wr, us.to(torch.complex128))            # same input as the timed plan's type 2
    e2 = float(torch.linalg.vector_norm(wr - wp.to(Zr)) / torch.linalg.vector_norm(wr))
    pr.close()
    del ur, wr

    def sample(fn):
        ts = []
        for _ in range(samples):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p.set_points(xp); fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts)

    t1 = sample(lambda: p.exec_type1(us, vp))
    t2 = sample(lambda: p.exec_type2(wp, us))
    desc = repr(p)
    p.close()
    return t1, t2, e1, e2, desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--types", nargs="+", default=["ComplexF64", "Float64", "ComplexF32"])
    ap.add_argument("--methods", nargs="+", default=["shared_memory"])
    ap.add_argument("--out", default="profiles/bench_dat")
    ap.add_argument("--sigma", type=float, default=1.5)
    ap.add_argument("--m", type=int, default=4)
    ap.add_argument("--samples", type=int, default=9)
    ap.add_argument("--max-rho", type=float, default=10.0)
    ap.add_argument("--fast", action="store_true", help="BackwardsKaiserBessel + FastApproximation (the reference's CPU default)")
    a = ap.parse_args()
    out = Path(a.out); out.mkdir(parents=True, exist_ok=True)
    rhos = [10.0 ** e for e in np.arange(-4, 1.01, 0.5) if 10.0 ** e <= a.max_rho * 1.0001]
    nps = [int(round(r * N ** 3)) for r in rhos]
    devname = torch.cuda.get_device_name(0)
    for tname in a.types:
        for method in a.methods:
            tag = "fast" if a.fast else "direct"
            fn = out / f"NonuniformFFTs_b200_{N}_{tname}_CUDA_{method}_sigma{a.sigma:g}_{tag}.dat"
            with open(fn, "w") as io:
                io.write("# nonuniformffts.jl_b200 (B200-native backend behind the NonuniformFFTs.jl API)\n# Benchmark: NUFFT of scalar data\n")
                io.write(f"#  - Backend: CUDA (sm_100a, C ABI)\n#  - Device: {devname}\n#  - Element type: {tname}\n")
                io.write(f"#  - Grid size: ({N}, {N}, {N})\n#  - Oversampling factor: {a.sigma:g}\n#  - Half support: HalfSupport({a.m})\n")
                io.write(f"#  - Kernel: {'BackwardsKaiserBesselKernel' if a.fast else 'KaiserBesselKernel'}\n")
                io.write(f"#  - Kernel evaluation: {'FastApproximation()' if a.fast else 'Direct()'}\n#  - GPU method: {method}\n")
                io.write("# (1) Number of points  (2) Type 1 (median, s)  (3) Type 2 (median, s)  (4) Relative error type 1  (5) Relative error type 2\n")
                for Np in nps:
                    t1, t2, e1, e2, _ = bench_one(TYPES[tname], Np, method, a.sigma, a.m, a.samples, a.fast)
                    io.write("\t".join(str(v) for v in (Np, t1, t2, e1, e2)) + "\n")
                    io.flush()
                    print(tname, method, Np, f"{t1 * 1e3:.3f} ms", f"{t2 * 1e3:.3f} ms", f"{e1:.2e}", f"{e2:.2e}", flush=True)
            print("wrote", fn, flush=True)


if __name__ == "__main__":
    main()
