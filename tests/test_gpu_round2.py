"""GPU parity tests added in round 2: far-away Float64 points, the reference's GPU test matrix at its own
number of points, C2 / C4 at their stated sizes, the AbstractNFFTs case at reltol = 1e-9, matrix-shaped points against the
oracle, polynomial tables against the oracle, and the multi-GPU strategies of the C ABI (skipped below two devices)."""
import numpy as np
import pytest

from helpers import TOL, complex_of, gpu_plan, l2_error, make_points, make_values, real_of, to_dev
from test_gpu_parity import run_case

pytestmark = pytest.mark.gpu


def test_far_away_points_float64(nufft, oracle_mod):
    """Points hundreds / hundreds of thousands of periods away fold at FULL Float64 precision (the reference's
    to_unit_cell_gpu uses a full-precision remainder, src/blocking/blocking.jl:23-33): the transform of x + 2 pi k must equal
    the transform of x up to the rounding of the shifted coordinate itself (ulp(|x|) * max wavenumber)."""
    import torch
    rng = np.random.default_rng(60)
    dims, Np = (64, 48), 4000
    x0 = [rng.random(Np) * 2 * np.pi for _ in dims]
    v = make_values(rng, Np, np.complex128)
    outs = []
    for shift in (0.0, 1.0e3, -1.0e6):
        k = np.round(shift / (2 * np.pi))
        xs = [np.ascontiguousarray(x + 2 * np.pi * k) for x in x0]
        gp = gpu_plan(nufft, np.complex128, dims, m=6, sigma=2.0)
        gp.set_points(tuple(to_dev(x) for x in xs))
        u = torch.empty(gp.shape, dtype=torch.complex128, device="cuda")
        gp.exec_type1(u, to_dev(v))
        outs.append((u.cpu().numpy(), max(abs(shift), 1.0)))
        gp.close()
    for u, mag in outs[1:]:
        # coordinate rounding: ulp(mag) ~ 2.2e-16 * mag, times the largest wavenumber (32), times a safety factor
        assert l2_error(u, outs[0][0]) <= 64 * 32 * 2.3e-16 * mag, (mag, l2_error(u, outs[0][0]))
    # and the oracle agrees on the unshifted points
    op = oracle_mod.OraclePlan(np.complex128, dims, m=6, sigma=2.0, kernel="backwards_kaiser_bessel", evalmode="fast", block_size=None)
    op.set_points(x0)
    assert l2_error(outs[0][0], op.exec_type1(v)) <= 1e-12


@pytest.mark.parametrize("method", ["global_memory", "shared_memory"])
@pytest.mark.parametrize("dtype,C", [(np.float32, 1), (np.complex64, 1), (np.float64, 1), (np.complex128, 1), (np.float32, 2)])
def test_reference_gpu_matrix_at_its_own_size(nufft, oracle_mod, dtype, C, method):
    """test/pseudo_gpu.jl:184-226 as the reference runs it: dims (35, 64, 40), Np = prod(dims) = 89 600, HalfSupport(4),
    KaiserBesselKernel, sigma = 1.5; Direct evaluation on both sides (the reference GPU default)."""
    dims = (35, 64, 40)
    run_case(nufft, oracle_mod, dtype, dims, int(np.prod(dims)), m=4, sigma=1.5, kernel="kaiser_bessel", evalmode="direct", C=C,
             method=method, seed=42, tol=1e-11 if real_of(dtype) == np.float64 else None, f32_relaxed=True)


def test_reference_gpu_matrix_callbacks(nufft, oracle_mod):
    """test/pseudo_gpu.jl:204-222: Np = prod(dims) / 2 with non-uniform weights and a dense uniform factor, ntransforms 1 and 2."""
    dims = (35, 64, 40)
    for C in (1, 2):
        for method in ("global_memory", "shared_memory"):
            run_case(nufft, oracle_mod, np.complex64, dims, int(np.prod(dims)) // 2, m=4, sigma=1.5, kernel="kaiser_bessel", evalmode="direct",
                     C=C, method=method, seed=43, callbacks=True, f32_relaxed=True)


def test_c2_full_size(nufft, oracle_mod):
    """BASELINE config C2: 2-D 256 x 256 modes, Np = 1000, Float64 real data, type 1 then type 2 of its output."""
    import torch
    rng = np.random.default_rng(2)
    dims, Np = (256, 256), 1000
    xs = [(rng.random(Np) * 2 * np.pi) for _ in dims]
    v = rng.standard_normal(Np)
    op = oracle_mod.OraclePlan(np.float64, dims, m=4, sigma=2.0, kernel="backwards_kaiser_bessel", evalmode="fast", block_size=None)
    op.set_points(xs)
    u_ref = op.exec_type1(v)
    v_ref = op.exec_type2(u_ref)
    gp = gpu_plan(nufft, np.float64, dims, m=4, sigma=2.0)
    gp.set_points(tuple(to_dev(x) for x in xs))
    perm, off, bdims = gp.binning()
    _, cum_o, perm_o = op.sort_points(xs, bdims)
    assert np.array_equal(off.cpu().numpy(), cum_o) and np.array_equal(perm.cpu().numpy(), perm_o)
    u = torch.empty(gp.shape, dtype=torch.complex128, device="cuda")
    gp.exec_type1(u, to_dev(v))
    assert l2_error(u.cpu().numpy(), u_ref) <= 1e-12
    w = torch.empty(Np, dtype=torch.float64, device="cuda")
    gp.exec_type2(w, u)
    assert l2_error(w.cpu().numpy(), v_ref) <= 1e-12
    gp.close()


def test_c4_full_size(nufft, oracle_mod):
    """BASELINE config C4 at its stated mode count: 3-D 256^3 modes, ntransforms = 3, Float64 real data, HalfSupport(8),
    KaiserBesselKernel (Direct, the GPU default), clustered points (wrapped normal, as the reference's benchmark draws).
    Against the oracle with 2^18 points (the CPU finishes in seconds); linearity in the points at 2^22 points."""
    import torch
    rng = np.random.default_rng(4)
    dims, Np, C = (256, 256, 256), 1 << 18, 3
    xs = [rng.standard_normal(Np) for _ in dims]
    vps = [rng.standard_normal(Np) for _ in range(C)]
    op = oracle_mod.OraclePlan(np.float64, dims, m=8, sigma=2.0, kernel="kaiser_bessel", evalmode="direct", ntransforms=C, block_size=4096,
                               use_blocked_spreading=True)
    op.set_points(xs)
    ref1 = op.exec_type1(vps)
    gp = gpu_plan(nufft, np.float64, dims, m=8, sigma=2.0, kernel="kaiser_bessel", evalmode="direct", ntransforms=C)
    gp.set_points(tuple(to_dev(x) for x in xs))
    perm, off, bdims = gp.binning()
    _, cum_o, perm_o = op.sort_points(xs, bdims)
    assert np.array_equal(off.cpu().numpy(), cum_o) and np.array_equal(perm.cpu().numpy(), perm_o)
    us = [torch.empty(gp.shape, dtype=torch.complex128, device="cuda") for _ in range(C)]
    gp.exec_type1(us, [to_dev(v) for v in vps])
    for c in range(C):
        assert l2_error(us[c].cpu().numpy(), ref1[c]) <= 1e-11, (c, l2_error(us[c].cpu().numpy(), ref1[c]))
    ref2 = op.exec_type2(ref1)
    ws = [torch.empty(Np, dtype=torch.float64, device="cuda") for _ in range(C)]
    gp.exec_type2(ws, us)
    for c in range(C):
        assert l2_error(ws[c].cpu().numpy(), ref2[c]) <= 1e-11
    # linearity in the POINTS at 2^22 clustered points (the binning stress of the config): T1(A u B) = T1(A) + T1(B)
    Np2 = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(44)
    xs2 = tuple(torch.randn(Np2, generator=g, device="cuda", dtype=torch.float64) for _ in dims)
    v2 = [torch.randn(Np2, generator=g, device="cuda", dtype=torch.float64) for _ in range(C)]
    gp.set_points(xs2)
    gp.exec_type1(us, v2)
    half = Np2 // 2
    acc = [torch.zeros_like(u) for u in us]
    tmp = [torch.empty_like(u) for u in us]
    for a, b in ((0, half), (half, Np2)):
        gp.set_points(tuple(x[a:b].contiguous() for x in xs2))
        gp.exec_type1(tmp, [v[a:b].contiguous() for v in v2])
        for c in range(C):
            acc[c] += tmp[c]
    for c in range(C):
        e = float(torch.linalg.vector_norm(acc[c] - us[c]) / torch.linalg.vector_norm(us[c]))
        assert e <= 1e-12, (c, e)
    gp.close()


@pytest.mark.parametrize("dims", [(512,), (64, 81)])
def test_nfft_frontend_reltol_1e9(nufft, dims):
    """test/abstractNFFTs.jl:9-70: Float64, Np = 1000, reltol = 1e-9, window = :kaiser_bessel, dims (512,) and (64, 81);
    the reference compares with NFFT.jl at sqrt(eps) ~ 1.5e-8 — here against the exact sums."""
    import torch
    rng = np.random.default_rng(43)
    D, Np = len(dims), 1000
    x = (rng.random((Np, D)) - 0.5)
    f = (rng.standard_normal(Np) + 1j * rng.standard_normal(Np))
    p = nufft.plan_nfft(to_dev(x), dims, reltol=1e-9, window="kaiser_bessel")
    assert p.size_in() == tuple(dims) and p.size_out() == (Np,)
    ks = [np.arange(-(n // 2), (n + 1) // 2) for n in dims]
    ph = np.zeros((Np,) + tuple(dims[::-1]))
    for d in range(D):
        shape = [1] * (D + 1)
        shape[D - d] = dims[d]
        ph = ph + x[:, d].reshape((Np,) + (1,) * D) * ks[d].reshape(shape)
    E = np.exp(-2j * np.pi * ph)
    ref_adj = (np.conj(E) * f.reshape((Np,) + (1,) * D)).sum(axis=0)
    adj = torch.empty(dims[::-1], dtype=torch.complex128, device="cuda")
    nufft.mul_adjoint(adj, p, to_dev(f))
    assert l2_error(adj.cpu().numpy(), ref_adj) <= 1.5e-8
    ref_fwd = (E * ref_adj[None]).reshape(Np, -1).sum(axis=1)
    out = torch.empty(Np, dtype=torch.complex128, device="cuda")
    nufft.mul(out, p, adj)
    assert l2_error(out.cpu().numpy(), ref_fwd) <= 1.5e-8
    p.close()


def test_matrix_points_against_the_oracle(nufft, oracle_mod):
    """set_points!(p, xp::Matrix (D, Np)) (src/set_points.jl:76-88): the in-place matrix path against the ORACLE (binning exact,
    results within the parity bars), not only against the tuple path."""
    import torch
    rng = np.random.default_rng(52)
    for dtype, dims in ((np.complex64, (24, 20, 16)), (np.float64, (30, 18)), (np.complex128, (64,))):
        rt = real_of(dtype)
        Np, D = 6000, len(dims)
        xs = make_points(rng, D, Np, rt, "uniform")
        v = make_values(rng, Np, dtype)
        op = oracle_mod.OraclePlan(dtype, dims, m=4, sigma=2.0, kernel="backwards_kaiser_bessel", evalmode="fast", block_size=None)
        op.set_points(xs)
        gp = gpu_plan(nufft, dtype, dims, m=4, sigma=2.0)
        gp.set_points(to_dev(np.ascontiguousarray(np.stack(xs, axis=1))))
        perm, off, bdims = gp.binning()
        _, cum_o, perm_o = op.sort_points(xs, bdims)
        assert np.array_equal(off.cpu().numpy(), cum_o) and np.array_equal(perm.cpu().numpy(), perm_o)
        u = torch.empty(gp.shape, dtype=gp.complex_dtype, device="cuda")
        gp.exec_type1(u, to_dev(v))
        u_ref = op.exec_type1(v)
        assert l2_error(u.cpu().numpy(), u_ref) <= TOL[rt]
        w = torch.empty(Np, dtype=gp.dtype, device="cuda")
        gp.exec_type2(w, to_dev(u_ref))
        assert l2_error(w.cpu().numpy(), op.exec_type2(u_ref)) <= TOL[rt]
        gp.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kernel", ["kaiser_bessel", "backwards_kaiser_bessel"])
def test_polynomial_tables_equal_the_oracle(nufft, oracle_mod, dtype, kernel):
    """The plan's own piecewise-polynomial coefficients (host_plan.cu: Chebyshev nodes, Vandermonde LU in T) against the
    oracle's (pinned to the reference's formulae by tests/test_oracle_mpmath.py): same arithmetic, same table."""
    for M, sigma in ((4, 2.0), (4, 1.5), (6, 1.25), (8, 2.0)):
        dims = (48, 40)
        cdt = complex_of(dtype)
        gp = gpu_plan(nufft, cdt, dims, m=M, sigma=sigma, kernel=kernel)
        op = oracle_mod.OraclePlan(cdt, dims, m=M, sigma=sigma, kernel=kernel)
        for d in range(2):
            ki = gp.kernel_info(d)
            kd = op.kernel_data(d)
            cs_g = np.asarray(ki["cs"], dtype=np.float64).reshape(M + 4, 2 * M)
            cs_o = kd["cs"].astype(np.float64)
            scale = np.abs(cs_o).max()
            # (two LU solves of the same ill-conditioned monomial Vandermonde system in T, different elimination order)
            assert np.abs(cs_g - cs_o).max() <= 2.0 ** (M + 6) * np.finfo(dtype).eps * scale, (M, sigma, d, np.abs(cs_g - cs_o).max() / scale)
            assert abs(ki["shape"] - kd["beta"]) <= 4 * np.spacing(dtype(kd["beta"]))
        gp.close()


# ---- the ES kernel (not in the reference: parity unpinned; the oracle restates the same construction) ----------------------------
@pytest.mark.parametrize("dtype,dims,Np,M,sigma", [
    (np.complex64, (35, 64, 40), 20000, 4, 2.0),          # ring-window kernels
    (np.float32, (35, 64, 40), 20000, 4, 2.0),            # column-streaming kernels (real data)
    (np.complex128, (64, 81), 5000, 4, 1.5),              # shared-memory tile kernels, cuFFT
    (np.float64, (256,), 1000, 6, 2.0),
])
def test_es_kernel_against_oracle(nufft, oracle_mod, dtype, dims, Np, M, sigma):
    """Tables (phihat by quadrature, polynomial coefficients), binning, type 1 and type 2 of an ES plan against the oracle."""
    from test_gpu_parity import run_case
    run_case(nufft, oracle_mod, dtype, dims, Np, m=M, sigma=sigma, kernel="es", evalmode="fast", seed=11)


def test_es_kernel_exact_sums_and_errors(nufft, oracle_mod):
    """Against exact NUDFT sums (the only absolute check an unpinned kernel has), and Direct evaluation is refused."""
    import torch
    rng = np.random.default_rng(12)
    N, Np = 64, 400
    x = (rng.random(Np) * 2 * np.pi)
    v = make_values(rng, Np, np.complex128)
    gp = gpu_plan(nufft, np.complex128, (N,), m=6, sigma=2.0, kernel="es")
    gp.set_points((to_dev(x),))
    u = torch.empty(gp.shape, dtype=torch.complex128, device="cuda")
    gp.exec_type1(u, to_dev(v))
    ks = [np.fft.fftfreq(N, 1 / N).astype(np.int64)]
    assert l2_error(u.cpu().numpy(), oracle_mod.nudft_type1(ks, [x], v)) < 4 * 6 * 10.0 ** (-1.9 * 6)
    gp.close()
    with pytest.raises(nufft.ArgumentError):
        nufft.PlanNUFFT(torch.complex64, (32, 32), kernel=nufft.ESKernel(), kernel_evalmode=nufft.Direct())
    p = nufft.PlanNUFFT(torch.complex64, (32, 32), kernel=nufft.ESKernel())           # default mode of ESKernel: FastApproximation
    assert "ESKernel" in repr(p)
    p.close()


# ---- multi-GPU strategies of the C ABI (nufft_mgpu_*), one process driving all devices ------------------------------------------
def _need_gpus(n):
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} CUDA devices")


@pytest.mark.parametrize("strategy,dist", [("slab", "uniform"), ("slab", "clustered"), ("points", "uniform")])
def test_multi_gpu_against_single_gpu(nufft, strategy, dist):
    _need_gpus(2)
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import mgpu_check
    import torch
    G = min(torch.cuda.device_count(), 8)
    G = G if G in (2, 4, 8) else 2
    assert mgpu_check.check(G, 64, 200000, strategy, dist=dist)


def test_multi_gpu_transform_sharding_float64(nufft):
    _need_gpus(2)
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import mgpu_check
    import torch
    assert mgpu_check.check(2, 32, 20000, "transforms", C=3, dtype=torch.complex128)
    assert mgpu_check.check(2, 32, 20000, "points", dtype=torch.complex128)


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_multi_gpu_slab_exchange_modes_and_growing_point_sets(nufft, p2p, monkeypatch):
    """The z-slab exchanges through peer windows (default) and through NCCL send / recv (NUFFT_B200_MGPU_P2P=0) give the
    single-GPU result; point sets that GROW between set_points calls force the collective reallocation (windows closed,
    buffers reallocated, windows republished), repeated transforms on one point set reuse the mappings."""
    _need_gpus(2)
    import torch
    monkeypatch.setenv("NUFFT_B200_MGPU_P2P", p2p)
    G = 2
    N = 64
    kw = dict(kernel=nufft.BackwardsKaiserBesselKernel(), kernel_evalmode=nufft.FastApproximation())
    mp = nufft.MultiGPUPlan(torch.complex64, (N, N, N), devices=list(range(G)), strategy="slab", **kw)
    plan1 = nufft.PlanNUFFT(torch.complex64, (N, N, N), device="cuda:0", **kw)
    try:
        assert mp.exchange == ("peer-windows" if p2p == "1" else "nccl")
        for it, n_loc in enumerate((3000, 40000, 40000, 90000, 500)):
            pts, vals, outs, back = [], [], [], []
            for g in range(G):
                dev = f"cuda:{g}"
                gen = torch.Generator(device=dev); gen.manual_seed(7 + 13 * g + it)
                n = n_loc + 17 * g
                pts.append(tuple(torch.rand(n, device=dev, generator=gen) * (2 * np.pi) for _ in range(3)))
                vals.append(torch.view_as_complex(torch.randn(n, 2, device=dev, generator=gen)))
                outs.append(torch.zeros(mp.local_shape(g), dtype=torch.complex64, device=dev))
                back.append(torch.zeros(n, dtype=torch.complex64, device=dev))
            for g in range(G):
                torch.cuda.synchronize(g)
            mp.set_points(pts)
            for rep in range(2):
                mp.exec_type1(outs, vals)
            mp.synchronize()
            full = mp.gather_output(outs)[0]
            mp.synchronize()
            plan1.set_points(tuple(torch.cat([p[d].to("cuda:0") for p in pts]) for d in range(3)))
            ref = torch.empty(plan1.shape, dtype=torch.complex64, device="cuda:0")
            plan1.exec_type1(ref, torch.cat([v.to("cuda:0") for v in vals]))
            torch.cuda.synchronize(0)
            assert l2_error(full.cpu().numpy(), ref.cpu().numpy()) <= 1e-5, (it, p2p)
            for rep in range(2):
                mp.exec_type2(back, outs)
            mp.synchronize()
            ref2 = torch.empty(sum(v.numel() for v in vals), dtype=torch.complex64, device="cuda:0")
            plan1.exec_type2(ref2, full.to("cuda:0"))
            torch.cuda.synchronize(0)
            got = torch.cat([b.to("cuda:0") for b in back])
            assert l2_error(got.cpu().numpy(), ref2.cpu().numpy()) <= 1e-5, (it, p2p)
    finally:
        mp.close(); plan1.close()


def test_multi_gpu_single_rank_handle(nufft):
    """nranks = 1 needs neither NCCL nor a second device: the handle degenerates to the single-GPU plan."""
    import torch
    rng = np.random.default_rng(70)
    Np, dims = 5000, (16, 16, 16)
    xs = [to_dev((rng.random(Np) * 2 * np.pi).astype(np.float32)) for _ in dims]
    v = to_dev(make_values(rng, Np, np.complex64))
    mp = nufft.MultiGPUPlan(torch.complex64, dims, devices=[0], strategy="points", kernel=nufft.BackwardsKaiserBesselKernel(),
                            kernel_evalmode=nufft.FastApproximation())
    mp.set_points(tuple(xs))
    u = torch.zeros(mp.local_shape(0), dtype=torch.complex64, device="cuda")
    mp.exec_type1(u, v)
    mp.synchronize()
    p1 = gpu_plan(nufft, np.complex64, dims)
    p1.set_points(tuple(xs))
    r = torch.empty(p1.shape, dtype=torch.complex64, device="cuda")
    p1.exec_type1(r, v)
    torch.cuda.synchronize()
    assert l2_error(u.cpu().numpy(), r.cpu().numpy()) <= 1e-6
    mp.close(); p1.close()
