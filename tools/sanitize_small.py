"""Small transform pairs through every kernel family — run under compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
3-D ComplexF32 HalfSupport(4) (ring-window kernels, single-sweep sort, pruned FFT), the same with real data (column-streaming
kernels), a 2-D Float64 plan (shared-memory tile kernels, cuFFT).  Results are checked against the oracle by the test helper."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import nufft_b200 as nb  # noqa: E402
import oracle  # noqa: E402
from test_gpu_parity import run_case  # noqa: E402

for dt, dims, Np in ((np.complex64, (24, 16, 32), 20000), (np.float32, (24, 16, 32), 20000), (np.complex128, (20, 24), 3000)):
    run_case(nb, oracle, dt, dims, Np, sigma=2.0, seed=5)
    print("ok", np.dtype(dt).name, dims, flush=True)
print("SANITIZE_SMALL done")
