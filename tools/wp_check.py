"""GPU check of the warp-private-tile kernels (3-D, M = 4, ComplexF32) vs the oracle, then C3 stage timings."""
import os
import sys
from pathlib import Path
import numpy as np
pass  # kernel family chosen by NUFFT_B200_CS / NUFFT_B200_WP in the environment
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import nufft_b200 as nb, oracle  # noqa: E402
from test_gpu_parity import run_case  # noqa: E402
ok = True
cases = [
    dict(dims=(35, 64, 40), Np=20000, sigma=1.5),
    dict(dims=(32, 32, 32), Np=50000, sigma=2.0, dist="clustered"),
    dict(dims=(16, 24, 20), Np=3000, sigma=2.0, C=2),
    dict(dims=(40, 12, 30), Np=7000, sigma=1.25, callbacks=True, f32_relaxed=True),
    dict(dims=(24, 24, 24), Np=9000, sigma=2.0, fftshift=True, kernel="gaussian"),
    dict(dims=(20, 20, 20), Np=1, sigma=2.0),
    dict(dims=(64, 64, 64), Np=300000, sigma=2.0),
]
for kw in cases:
    kw = dict(kw)
    dims, Np = kw.pop("dims"), kw.pop("Np")
    try:
        run_case(nb, oracle, np.complex64, dims, Np, method="shared_memory", seed=3, **kw)
        print("ok  ", dims, Np, kw, flush=True)
    except Exception as e:
        ok = False
        print("FAIL", dims, Np, kw, str(e)[:300], flush=True)
print("ALL OK" if ok else "FAILURES", flush=True)
