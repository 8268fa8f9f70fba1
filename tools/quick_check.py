"""Quick GPU correctness check of the M=4 paths vs the oracle (dev builds)."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import nufft_b200 as nb, oracle
from helpers import *
from test_gpu_parity import run_case
ok = True
for dt in (np.complex64, np.float32, np.float64, np.complex128):
    for dims, Np in (((35, 64, 40), 20000), ((64, 81), 5000), ((256,), 1000)):
        for method in ("shared_memory", "global_memory"):
            for dist in ("uniform", "clustered"):
                try:
                    run_case(nb, oracle, dt, dims, Np, sigma=1.5, method=method, seed=1, dist=dist, C=2 if dims == (64, 81) else 1)
                    print("ok  ", np.dtype(dt).name, dims, method, dist)
                except Exception as e:
                    ok = False
                    print("FAIL", np.dtype(dt).name, dims, method, dist, str(e)[:200])
print("ALL OK" if ok else "FAILURES")
