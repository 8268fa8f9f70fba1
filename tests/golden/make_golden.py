"""
Generates tests/golden/golden_cases.npz: small seeded inputs and the ORACLE's outputs for them.

The reference is Julia and cannot run in this container, so these fixtures are outputs of the CPU oracle
(which is itself pinned against the reference's known-answer tests, tests/test_oracle_reference_tests.py).
They freeze the oracle's behaviour (CPU regression test) and travel to the GPU box, where the CUDA path is
compared against them without rebuilding anything.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from helpers import make_points, make_values, real_of, complex_of  # noqa: E402

# name: (dtype, dims, Np, M, sigma, kernel, evalmode, C, dist)
CASES = {
    "c1_1d_f64_readme": (np.float64, (256,), 100, 4, 2.0, "backwards_kaiser_bessel", "fast", 1, "uniform"),
    "c2_2d_f64_real": (np.float64, (32, 32), 1000, 4, 2.0, "backwards_kaiser_bessel", "fast", 1, "uniform"),
    "c3_3d_c64_m4": (np.complex64, (16, 16, 16), 2000, 4, 2.0, "backwards_kaiser_bessel", "fast", 1, "uniform"),
    "c4_3d_f64_m8_kb_clustered_nt3": (np.float64, (16, 16, 16), 1500, 8, 2.0, "kaiser_bessel", "fast", 3, "clustered"),
    "kb_direct_c128_2d": (np.complex128, (24, 20), 800, 6, 1.5, "kaiser_bessel", "direct", 1, "uniform"),
    "gauss_f32_3d": (np.float32, (12, 16, 10), 900, 3, 2.0, "gaussian", "fast", 1, "blobs"),
    "bspline_c128_1d": (np.complex128, (100,), 300, 5, 2.0, "bspline", "fast", 2, "uniform"),
}


def main():
    out = {}
    for name, (dtype, dims, Np, M, sigma, kernel, mode, C, dist) in CASES.items():
        seed = abs(hash(name)) % (2 ** 31) if False else sum(ord(ch) for ch in name)   # deterministic
        rng = np.random.default_rng(seed)
        rt, ct = real_of(dtype), complex_of(dtype)
        xs = make_points(rng, len(dims), Np, rt, dist)
        vps = [make_values(rng, Np, dtype) for _ in range(C)]
        p = oracle.OraclePlan(dtype, dims, m=M, sigma=sigma, kernel=kernel, evalmode=mode, ntransforms=C, block_size=None)
        p.set_points(xs)
        u = p.exec_type1(vps if C > 1 else vps[0])
        u = u if C > 1 else [u]
        uks = [make_values(rng, int(np.prod(p.size)), ct).reshape(p.size[::-1]) for _ in range(C)]
        w = p.exec_type2(uks if C > 1 else uks[0])
        w = w if C > 1 else [w]
        bdims = tuple(min(8, n) for n in p.Nos)
        _, cum, perm = p.sort_points(xs, bdims)
        out[name + "/xs"] = np.stack(xs)
        out[name + "/vp"] = np.stack(vps)
        out[name + "/type1"] = np.stack(u)
        out[name + "/uk"] = np.stack(uks)
        out[name + "/type2"] = np.stack(w)
        out[name + "/bdims"] = np.array(bdims)
        out[name + "/cum"] = cum
        out[name + "/perm"] = perm
    path = Path(__file__).resolve().parent / "golden_cases.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
