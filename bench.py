#!/usr/bin/env python
"""
bench.py — headline benchmark of the B200-native NUFFT backend (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): 3-D NUFFT type-1/type-2 points/s at 256^3 modes, Np = 256^3 (config C3:
ComplexF32, HalfSupport(4), default sigma = 2 -> 512^3 oversampled grid, uniform-random points).
A "step" follows the reference's protocol (benchmark/CPU+CUDA/run_benchmarks.jl:82-92) for both
transform types back to back:   set_points! + exec_type1!   then   set_points! + exec_type2!
so one step processes 2*Np points per GPU.  `value` = points / s with inputs resident in HBM;
`e2e` = the same through the public API with pinned HOST buffers (H2D of points+values / spectrum and
D2H of the result inside the timed region).

N > 1 (torchrun, one rank per GPU): STRONG scaling of config C5 — ONE 512^3-mode problem (1024^3 oversampled grid),
Np = 2^27 points in total, dealt evenly to the ranks; z-slab strategy of the multi-GPU C ABI (nufft_mgpu_*, csrc/mgpu.cu):
point / value exchange, slab spreading + halo exchange, slab-decomposed pruned FFT with one all-to-all transpose.
The line also carries the same problem timed on ONE GPU in the same run (`single_gpu_same_problem`), a multi-rank result
check against a single-GPU transform (`check`), and the weak-scaling number of round 1 (`weak_c3`: 2^24 own points of a
256^3 problem per rank, partial outputs all-reduced) as extra keys.  N = 1 is the headline C3 line (+ `c5_single_gpu`).

`--impl reference` times the reference's CPU algorithm (the oracle port, all host threads) on a
bounded sample of the same workload.  The oracle is only ever used there and in `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_MODES = 256
NP_FULL = 256 ** 3
HALF_SUPPORT = 4
SIGMA = 2.0
KERNEL = "backwards_kaiser_bessel"      # + FastApproximation == the reference CPU path's defaults
METRIC = "3D NUFFT type-1/type-2 points/sec at 256^3 modes, Np=256^3, vs HBM roofline"
UNIT = "points/s"


def measured_peak_gbs():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(seed: int, npts: int):
    rng = np.random.default_rng(seed)
    xs = [(rng.random(npts, dtype=np.float32) * np.float32(2 * np.pi)) for _ in range(3)]
    two_pi = np.float32(2) * np.float32(np.pi)
    for x in xs:
        x[x >= two_pi] = 0.0
    vp = (rng.standard_normal(npts, dtype=np.float32) + 1j * rng.standard_normal(npts, dtype=np.float32)).astype(np.complex64)
    nk = N_MODES ** 3
    uk = (rng.standard_normal(nk, dtype=np.float32) + 1j * rng.standard_normal(nk, dtype=np.float32)).astype(np.complex64)
    return xs, vp, uk.reshape(N_MODES, N_MODES, N_MODES)


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, sample_points: int):
    import oracle
    oracle.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    cores = oracle.num_threads()
    xs, vp, uk = make_inputs(3, sample_points)
    plan = oracle.OraclePlan(np.complex64, (N_MODES,) * 3, m=HALF_SUPPORT, sigma=SIGMA, kernel=KERNEL, evalmode="fast",
                             block_size=4096, use_blocked_spreading=True)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        plan.set_points(xs)
        plan.exec_type1(vp)
        plan.set_points(xs)
        plan.exec_type2(uk)
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    value = 2.0 * sample_points * len(times) / total
    return value, cores, total / len(times) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = NP_FULL
    steps = max(1, min(args.steps, 3))          # each full-size CPU step takes ~10-20 s: bound the run to a few minutes
    warm = 1 if args.warmup > 0 else 0
    value, cores, ms = cpu_reference_run(steps, warm, sample)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample_desc = (("N > 1: the repo arm runs C5 (512^3 modes, Np = 2^27); the CPU arm times a C3-sized sample of the same per-point "
                    "work (the 1024^3 grid of C5 does not fit the time bound on host cores): " if world > 1 else "") +
                   f"full C3 workload per step (256^3 modes, 512^3 oversampled grid, Np = 2^24 uniform-random points), "
                   f"{steps} timed step(s) after {warm} warm-up (step count bounded to keep the run within minutes); "
                   "restatement of the reference CPU algorithm (oracle port: blocked spreading/interpolation + OpenMP, "
                   "pocketfft), NOT NonuniformFFTs.jl itself (no Julia in this image)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3: 3D 256^3 modes, ComplexF32, HalfSupport(4), sigma=2, uniform-random points "
                               "Np=2^24; step = set_points+type1, set_points+type2 (2*Np points per step)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0



# ----------------------------------------------------------------------------------------------
# N > 1: strong scaling of C5 through the multi-GPU C ABI
# ----------------------------------------------------------------------------------------------
C5_MODES = int(os.environ.get("BENCH_C5_MODES", "512"))          # (development: smaller problems under compute-sanitizer)
C5_NP = int(os.environ.get("BENCH_C5_NP", str(512 ** 3)))
BENCH_QUICK = os.environ.get("BENCH_QUICK", "") == "1"              # skip the single-GPU baseline and the weak-scaling extra


def bind_to_gpu_numa(local_rank: int):
    """Pin this process to the host cores of the GPU's NUMA node before allocating pinned buffers (first touch places them there)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return node
    except Exception:
        return None


def multi_rank_check(nb, torch, dist, world, rank, dev):
    """64^3-mode problem: N-rank slab transform against a single-GPU transform of the same inputs (rank 0 computes it)."""
    modes, npts = 64, 400000
    kw = dict(m=HALF_SUPPORT, sigma=SIGMA, kernel=nb.BackwardsKaiserBesselKernel(), kernel_evalmode=nb.FastApproximation())
    rng = np.random.default_rng(11)
    xs = [(rng.random(npts) * 2 * np.pi).astype(np.float32) for _ in range(3)]
    vp = (rng.standard_normal(npts) + 1j * rng.standard_normal(npts)).astype(np.complex64)
    uk = (rng.standard_normal((modes,) * 3) + 1j * rng.standard_normal((modes,) * 3)).astype(np.complex64)
    a, b = nb.partition_points(npts, world, rank)
    mp = nb.MultiGPUPlan(torch.complex64, (modes,) * 3, strategy="slab", **kw)
    mp.set_points(tuple(torch.from_numpy(x[a:b]).to(dev) for x in xs))
    out = torch.zeros(mp.local_shape(0), dtype=torch.complex64, device=dev)
    mp.exec_type1(out, torch.from_numpy(vp[a:b]).to(dev))
    full = mp.gather_output(out)
    off, sz = mp.local_block(0)
    v = torch.zeros(b - a, dtype=torch.complex64, device=dev)
    mp.exec_type2(v, torch.from_numpy(np.ascontiguousarray(uk[:, off[1]:off[1] + sz[1], :])).to(dev))
    mp.synchronize()
    err = torch.zeros(2, device=dev)
    p1 = nb.PlanNUFFT(torch.complex64, (modes,) * 3, device=dev, **kw)      # every rank checks its own share against a local single-GPU plan
    p1.set_points(tuple(torch.from_numpy(x).to(dev) for x in xs))
    r1 = torch.empty(p1.shape, dtype=torch.complex64, device=dev)
    p1.exec_type1(r1, torch.from_numpy(vp).to(dev))
    r2 = torch.empty(npts, dtype=torch.complex64, device=dev)
    p1.exec_type2(r2, torch.from_numpy(uk).to(dev))
    torch.cuda.synchronize()
    err[0] = (torch.linalg.vector_norm(full - r1) / torch.linalg.vector_norm(r1)).float()
    err[1] = (torch.linalg.vector_norm(v - r2[a:b]) / torch.linalg.vector_norm(r2[a:b])).float()
    p1.close(); mp.close()
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    e1, e2 = float(err[0]), float(err[1])
    return {"result": "ok" if (e1 <= 2e-5 and e2 <= 2e-5) else "FAILED", "type1_rel_l2_vs_single_gpu": e1, "type2_rel_l2_vs_single_gpu": e2,
            "problem": "64^3 modes, 4e5 points, slab strategy"}


def run_ours_multi(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    import nufft_b200 as nb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank)
    dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    kw = dict(m=HALF_SUPPORT, sigma=SIGMA, kernel=nb.BackwardsKaiserBesselKernel(), kernel_evalmode=nb.FastApproximation())

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        if drain is not None:
            drain()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    check = multi_rank_check(nb, torch, dist, world, rank, dev)

    # ---- C5, strong scaling: one problem, points dealt evenly (and randomly) to the ranks -----------------------------------
    n_loc = C5_NP // world
    gen = torch.Generator(device=dev); gen.manual_seed(5 + rank)
    xs_d = [torch.rand(n_loc, device=dev, generator=gen) * (2 * np.pi) for _ in range(3)]
    for x in xs_d:
        x[x >= 2 * np.pi] = 0.0
    vp_d = torch.view_as_complex(torch.randn(n_loc, 2, device=dev, generator=gen))
    plan = nb.MultiGPUPlan(torch.complex64, (C5_MODES,) * 3, strategy="slab", timer=True, **kw)
    shp = plan.local_shape(0)
    uk_d = torch.view_as_complex(torch.randn(shp + (2,), device=dev, generator=gen))
    out1_d = torch.zeros(shp, dtype=torch.complex64, device=dev)
    out2_d = torch.zeros(n_loc, dtype=torch.complex64, device=dev)

    def step_device():
        plan.set_points(tuple(xs_d)); plan.exec_type1(out1_d, vp_d)
        plan.set_points(tuple(xs_d)); plan.exec_type2(out2_d, uk_d)

    for _ in range(W):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    nb.launch_count(reset=True)
    ms_total = timed(step_device, K)
    launches = nb.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    stages = plan.timings(0)
    exchange = plan.exchange

    def t1_only():
        plan.set_points(tuple(xs_d)); plan.exec_type1(out1_d, vp_d)
    def t2_only():
        plan.set_points(tuple(xs_d)); plan.exec_type2(out2_d, uk_d)
    ms_t1 = timed(t1_only, K) / K
    ms_t2 = timed(t2_only, K) / K

    # ---- end to end: this rank's inputs from pinned host memory every step, results back (double-buffered copy streams) -------
    pin = lambda t: t.cpu().pin_memory()
    xs_pin, vp_pin, uk_pin = [pin(x) for x in xs_d], pin(vp_d), pin(uk_d)
    out1_pin, out2_pin = torch.empty_like(out1_d, device="cpu").pin_memory(), torch.empty_like(out2_d, device="cpu").pin_memory()
    main = torch.cuda.current_stream(dev)
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    NB = 2
    xs_b = [[torch.empty_like(x) for x in xs_d] for _ in range(NB)]
    vp_b = [torch.empty_like(vp_d) for _ in range(NB)]
    uk_b = [torch.empty_like(uk_d) for _ in range(NB)]
    o1_b = [torch.empty_like(out1_d) for _ in range(NB)]
    o2_b = [torch.empty_like(out2_d) for _ in range(NB)]
    ev_in = [torch.cuda.Event() for _ in range(NB)]
    ev_free = [torch.cuda.Event() for _ in range(NB)]
    ev_o1 = [torch.cuda.Event() for _ in range(NB)]
    ev_o2 = [torch.cuda.Event() for _ in range(NB)]
    ev_out = [torch.cuda.Event() for _ in range(NB)]
    st = {"k": 0}

    def step_e2e():
        k = st["k"]; st["k"] = k + 1
        b = k % NB
        with torch.cuda.stream(s_h2d):
            if k >= NB:
                s_h2d.wait_event(ev_free[b])
            for dst, src in zip(xs_b[b], xs_pin):
                dst.copy_(src, non_blocking=True)
            vp_b[b].copy_(vp_pin, non_blocking=True)
            uk_b[b].copy_(uk_pin, non_blocking=True)
            ev_in[b].record(s_h2d)
        main.wait_event(ev_in[b])
        if k >= NB:
            main.wait_event(ev_out[b])
        plan.set_points(tuple(xs_b[b])); plan.exec_type1(o1_b[b], vp_b[b])
        ev_o1[b].record(main)
        plan.set_points(tuple(xs_b[b])); plan.exec_type2(o2_b[b], uk_b[b])
        ev_o2[b].record(main)
        ev_free[b].record(main)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_o1[b])
            out1_pin.copy_(o1_b[b], non_blocking=True)
            s_d2h.wait_event(ev_o2[b])
            out2_pin.copy_(o2_b[b], non_blocking=True)
            ev_out[b].record(s_d2h)

    def e2e_drain():
        main.wait_stream(s_d2h)
        main.wait_stream(s_h2d)

    for _ in range(2):
        step_e2e()
    e2e_drain()
    ms_e2e = timed(step_e2e, K, drain=e2e_drain)
    h2d = (sum(x.numel() * 4 for x in xs_pin) + vp_pin.numel() * 8 + uk_pin.numel() * 8) * world
    d2h = (out1_pin.numel() * 8 + out2_pin.numel() * 8) * world
    plan.close()
    del xs_b, vp_b, uk_b, o1_b, o2_b, uk_d, out1_d, out2_d
    torch.cuda.empty_cache()

    # ---- the same problem on ONE GPU (rank 0), same run: the strong-scaling baseline -----------------------------------------
    single = None
    barrier()
    if rank == 0 and not BENCH_QUICK:
        try:
            gen0 = torch.Generator(device=dev); gen0.manual_seed(99)
            xs1 = [torch.rand(C5_NP, device=dev, generator=gen0) * (2 * np.pi) for _ in range(3)]
            for x in xs1:
                x[x >= 2 * np.pi] = 0.0
            vp1 = torch.view_as_complex(torch.randn(C5_NP, 2, device=dev, generator=gen0))
            p1 = nb.PlanNUFFT(torch.complex64, (C5_MODES,) * 3, device=dev, timer=True, **kw)
            o1 = torch.zeros(p1.shape, dtype=torch.complex64, device=dev)
            o2 = torch.zeros(C5_NP, dtype=torch.complex64, device=dev)
            def s1():
                p1.set_points(tuple(xs1)); p1.exec_type1(o1, vp1)
                p1.set_points(tuple(xs1)); p1.exec_type2(o2, o1)
            for _ in range(2):
                s1()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n1 = 3
            e0.record()
            for _ in range(n1):
                s1()
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / n1
            single = {"ms_per_step": ms1, "value": 2.0 * C5_NP / (ms1 * 1e-3), "unit": UNIT, "steps": n1, "stage_ms": dict(p1.timer),
                      "note": "C5 on one GPU (8 GiB grid), single-GPU plan, timed on rank 0 in this run"}
            p1.close()
            del xs1, vp1, o1, o2
        except Exception as e:      # pragma: no cover
            single = {"error": str(e)[:200]}
        torch.cuda.empty_cache()
    barrier()

    # ---- weak scaling as in round 1: every rank its own 2^24 points of a 256^3 problem, points strategy ------------------------
    npts = NP_FULL if not BENCH_QUICK else 1 << 18
    xs_w = [torch.rand(npts, device=dev, generator=gen) * (2 * np.pi) for _ in range(3)]
    vp_w = torch.view_as_complex(torch.randn(npts, 2, device=dev, generator=gen))
    uk_w = torch.view_as_complex(torch.randn((N_MODES,) * 3 + (2,), device=dev, generator=gen))
    pw = nb.MultiGPUPlan(torch.complex64, (N_MODES,) * 3, strategy="points", **kw)
    ow1 = torch.zeros((N_MODES,) * 3, dtype=torch.complex64, device=dev)
    ow2 = torch.zeros(npts, dtype=torch.complex64, device=dev)
    def step_weak():
        pw.set_points(tuple(xs_w)); pw.exec_type1(ow1, vp_w)
        pw.set_points(tuple(xs_w)); pw.exec_type2(ow2, uk_w)
    for _ in range(3):
        step_weak()
    ms_weak = timed(step_weak, K)
    pw.close()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        pts_per_step = 2.0 * C5_NP
        value = pts_per_step * K / (ms_total * 1e-3)
        # roofline of the dominant per-rank kernel: slab spreading / interpolation of n_loc points on 1/world of the grid
        G_x = (2 * C5_MODES) ** 3 * 8 / world
        cand = {"spread": (stages["T1 zero fill + spreading"], n_loc * 24 + G_x), "interp": (stages["T2 interpolation"], G_x + n_loc * 24)}
        dom = max(cand, key=lambda k: cand[k][0])
        dom_ms, dom_bytes = cand[dom]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5: 3D 512^3 modes (1024^3 oversampled grid), Np=2^27 uniform-random points IN TOTAL dealt evenly to the "
                                   "ranks, ComplexF32, HalfSupport(4), sigma=2, BackwardsKaiserBessel+FastApproximation; step = "
                                   "set_points+type1, set_points+type2 (2*Np points per step); N=1 line of this bench is C3 (the headline)",
                       "l2": "inputs larger than L2 (>= 1 GiB of grid per rank)",
                       "multi_gpu": "z-slab strategy of the C ABI (nufft_mgpu_*): point/value all-to-all, slab spreading + halo exchange, "
                                    "slab pruned FFT with one all-to-all transpose; uniform data distributed in y",
                       "exchange": exchange},
            "type1_points_per_s": C5_NP / (ms_t1 * 1e-3), "type2_points_per_s": C5_NP / (ms_t2 * 1e-3),
            "type1_ms": ms_t1, "type2_ms": ms_t2,
            "stage_ms_rank0": stages,
            "check": check["result"], "check_detail": check,
            "single_gpu_same_problem": single,
            "speedup_vs_single_gpu": (single["ms_per_step"] / (ms_total / K)) if single and "ms_per_step" in single else None,
            "weak_c3": {"value": 2.0 * npts * world * K / (ms_weak * 1e-3), "unit": UNIT, "ms_per_step": ms_weak / K,
                        "workload": "C3 per rank (2^24 own points of one 256^3 problem), points strategy: outputs all-reduced, spectrum broadcast"},
            "roofline": {"bound": "hbm", "kernel": f"K-{dom} (per rank)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "algorithmic_bytes": dom_bytes, "peak_source": peak_src,
                         "launch_ms": dom_ms},
            "cpu_baseline": None, "clocks": clocks,
            "e2e": {"value": pts_per_step * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "numa_node_rank0": numa,
                    "h2d_gb_per_s": h2d * K / (ms_e2e * 1e-3) / 1e9, "d2h_gb_per_s": d2h * K / (ms_e2e * 1e-3) / 1e9,
                    "note": "public API (MultiGPUPlan) with pinned host buffers allocated on each GPU's NUMA node; per step every rank "
                            "copies its points+values+spectrum block in and both results out on copy streams, double-buffered"},
            "gpu_launches": int(launches),
        }
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import nufft_b200 as nb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if world > 1:
        try:
            return run_ours_multi(args, world, rank, local_rank)
        except BaseException:
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            os._exit(1)          # no destructors: the other ranks are inside collectives, the launcher tears them down
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    npts = NP_FULL
    K, W = args.steps, max(args.warmup, 3)

    xs_h, vp_h, uk_h = make_inputs(3 + rank, npts)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    xs_pin, vp_pin, uk_pin = [pin(x) for x in xs_h], pin(vp_h), pin(uk_h)
    xs_d = [x.to(dev) for x in xs_pin]
    vp_d, uk_d = vp_pin.to(dev), uk_pin.to(dev)
    out1_d = torch.empty((N_MODES,) * 3, dtype=torch.complex64, device=dev)
    out2_d = torch.empty(npts, dtype=torch.complex64, device=dev)
    out1_pin = torch.empty((N_MODES,) * 3, dtype=torch.complex64).pin_memory()
    out2_pin = torch.empty(npts, dtype=torch.complex64).pin_memory()

    plan = nb.PlanNUFFT(torch.complex64, (N_MODES,) * 3, m=HALF_SUPPORT, sigma=SIGMA,
                        kernel=nb.BackwardsKaiserBesselKernel(), kernel_evalmode=nb.FastApproximation(),
                        timer=True, device=dev)

    pp = nb.PointPartitionedNUFFT(plan)     # N > 1: type-1 partial outputs all-reduced, type-2 spectrum broadcast

    def step_device():
        pp.set_points(tuple(xs_d))
        pp.exec_type1(out1_d, vp_d)
        pp.set_points(tuple(xs_d))
        pp.exec_type2(out2_d, uk_d, src=0)

    # End to end through the public API with HOST buffers.  Every step copies its inputs (points, values, spectrum) from
    # pinned host memory and reads both results back; the copies run on two copy streams and are double-buffered, so the
    # H2D of step k + 1 and the D2H of step k overlap the transforms of the neighbouring steps (what a user of a
    # stream-ordered API does).  Nothing is reused between steps: each step's inputs cross PCIe again.
    main = torch.cuda.current_stream(dev)
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    NB = 2
    xs_b = [[torch.empty_like(x) for x in xs_d] for _ in range(NB)]
    vp_b = [torch.empty_like(vp_d) for _ in range(NB)]
    uk_b = [torch.empty_like(uk_d) for _ in range(NB)]
    o1_b = [torch.empty_like(out1_d) for _ in range(NB)]
    o2_b = [torch.empty_like(out2_d) for _ in range(NB)]
    ev_in = [torch.cuda.Event() for _ in range(NB)]
    ev_free = [torch.cuda.Event() for _ in range(NB)]       # inputs of buffer b consumed
    ev_o1 = [torch.cuda.Event() for _ in range(NB)]
    ev_o2 = [torch.cuda.Event() for _ in range(NB)]
    ev_out = [torch.cuda.Event() for _ in range(NB)]        # results of buffer b are on the host
    e2e_state = {"k": 0}

    def step_e2e():
        k = e2e_state["k"]; e2e_state["k"] = k + 1
        b = k % NB
        with torch.cuda.stream(s_h2d):
            if k >= NB:
                s_h2d.wait_event(ev_free[b])
            for dst, src in zip(xs_b[b], xs_pin):
                dst.copy_(src, non_blocking=True)
            vp_b[b].copy_(vp_pin, non_blocking=True)
            uk_b[b].copy_(uk_pin, non_blocking=True)
            ev_in[b].record(s_h2d)
        main.wait_event(ev_in[b])
        if k >= NB:
            main.wait_event(ev_out[b])                      # previous results of this buffer have left the device
        pp.set_points(tuple(xs_b[b]))
        pp.exec_type1(o1_b[b], vp_b[b])
        ev_o1[b].record(main)
        pp.set_points(tuple(xs_b[b]))
        pp.exec_type2(o2_b[b], uk_b[b], src=0)
        ev_o2[b].record(main)
        ev_free[b].record(main)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_o1[b])
            out1_pin.copy_(o1_b[b], non_blocking=True)
            s_d2h.wait_event(ev_o2[b])
            out2_pin.copy_(o2_b[b], non_blocking=True)
            ev_out[b].record(s_d2h)

    def e2e_drain():
        main.wait_stream(s_d2h)                             # the timed region ends when the last result is on the host
        main.wait_stream(s_h2d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        if drain is not None:
            drain()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(W):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    nb.launch_count(reset=True)
    ms_total = timed(step_device, K)
    launches = nb.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    timer = plan.timer                                    # per-stage device times of the last step
    # per-type timings (separate loops, same protocol)
    def t1_only():
        plan.set_points(tuple(xs_d)); plan.exec_type1(out1_d, vp_d)
    def t2_only():
        plan.set_points(tuple(xs_d)); plan.exec_type2(out2_d, uk_d)
    ms_t1 = timed(t1_only, K) / K
    ms_t2 = timed(t2_only, K) / K

    # clustered points (north_star: "uniform-random and clustered"): Gaussian cloud, sigma = 1 rad around x = 0 (folded by
    # set_points), same values / spectrum; same protocol, reported beside the headline numbers
    rng_c = np.random.default_rng(1000 + rank)
    xs_c = [torch.from_numpy(rng_c.standard_normal(npts, dtype=np.float32)).to(dev) for _ in range(3)]
    def t1_clustered():
        plan.set_points(tuple(xs_c)); plan.exec_type1(out1_d, vp_d)
    def t2_clustered():
        plan.set_points(tuple(xs_c)); plan.exec_type2(out2_d, uk_d)
    t1_clustered(); t2_clustered()
    ms_t1c = timed(t1_clustered, K) / K
    ms_t2c = timed(t2_clustered, K) / K
    timer_c = plan.timer
    del xs_c

    for _ in range(2):
        step_e2e()
    e2e_drain()
    ms_e2e = timed(step_e2e, K, drain=e2e_drain)

    pts_per_step = 2.0 * npts * world
    value = pts_per_step * K / (ms_total * 1e-3)
    e2e_value = pts_per_step * K / (ms_e2e * 1e-3)
    h2d = sum(x.numel() * 4 for x in xs_pin) + vp_pin.numel() * 8 + uk_pin.numel() * 8
    d2h = out1_pin.numel() * 8 + out2_pin.numel() * 8

    # roofline of the dominant kernel (K-spread): algorithmic bytes = Np*(D*s_T + C*s_Z + 4) + C*G_x
    peak, peak_src = measured_peak_gbs()
    G_x = (2 * N_MODES) ** 3 * 8
    stages = {"spread": (timer["T1 (1) Spreading"], npts * (3 * 4 + 8 + 4) + G_x),
              "interp": (timer["T2 (3) Interpolation"], G_x + npts * (3 * 4 + 4 + 8))}
    dom = max(stages, key=lambda k: stages[k][0])
    dom_ms, dom_bytes = stages[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # whole-transform fractions (B_1 + B_sp, B_2 + B_sp of SURVEY §8d)
    B_sp = npts * 16
    B_1 = npts * 24 + 3 * G_x + 2 * N_MODES ** 3 * 8
    B_2 = N_MODES ** 3 * 8 + 4 * G_x + npts * 24
    traffic = None                  # measured DRAM bytes per launch of that kernel (ncu --set full capture, profiles/)
    try:
        t = json.loads((ROOT / "profiles" / "r2_dominant_kernel_traffic.json").read_text())[f"K-{dom}"]
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": f"K-{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": dom_bytes, "peak_source": peak_src,
            "note": "ring-window register kernels (ring_spread.cuh / ring_interp.cuh): not HBM-bound at one point per 8 fine cells — "
                    "interpolation: L1/LSU busy 77 %, issue slots 63 %, FMA pipe 43 %, DRAM 23 % of peak; spreading: bound by the "
                    "per-lane L2 reductions (0.99 ms without them, 2.79 ms with), issue slots 64 %, DRAM 32 % "
                    "(profiles/r2_ring_interp_ncu_details.txt, r2b_ring_spread_ncu_details.txt, r2_spread_red_experiments.txt); "
                    "measured DRAM traffic is 3.8 x (interpolation) / 5.3 x (spreading) the algorithmic bytes: neighbouring 4x4-cell "
                    "columns re-read / re-reduce the overlapping part of their 11x11-cell windows (7.6 x the grid in total) and L2 "
                    "catches about half of it (hit rate 43-50 %); see DESIGN.md section 3",
            "launch_ms": dom_ms,
            "type1_incl_set_points_frac": (B_1 + B_sp) / (ms_t1 * 1e-3) / 1e9 / peak,
            "type2_incl_set_points_frac": (B_2 + B_sp) / (ms_t2 * 1e-3) / 1e9 / peak,
            "stage_ms": timer}

    # C5 (512^3 modes, Np = 2^27) on this one GPU: the strong-scaling baseline of the N > 1 lines, reported beside the headline
    c5 = None
    if not args.no_c5:
        try:
            plan.close()
            del xs_b, vp_b, uk_b, o1_b, o2_b
            torch.cuda.empty_cache()
            gen0 = torch.Generator(device=dev); gen0.manual_seed(99)
            xs5 = [torch.rand(C5_NP, device=dev, generator=gen0) * (2 * np.pi) for _ in range(3)]
            for x in xs5:
                x[x >= 2 * np.pi] = 0.0
            vp5 = torch.view_as_complex(torch.randn(C5_NP, 2, device=dev, generator=gen0))
            p5 = nb.PlanNUFFT(torch.complex64, (C5_MODES,) * 3, m=HALF_SUPPORT, sigma=SIGMA, kernel=nb.BackwardsKaiserBesselKernel(),
                              kernel_evalmode=nb.FastApproximation(), timer=True, device=dev)
            o5 = torch.zeros(p5.shape, dtype=torch.complex64, device=dev)
            w5 = torch.zeros(C5_NP, dtype=torch.complex64, device=dev)
            def s5():
                p5.set_points(tuple(xs5)); p5.exec_type1(o5, vp5)
                p5.set_points(tuple(xs5)); p5.exec_type2(w5, o5)
            for _ in range(2):
                s5()
            n5 = 3
            ms5 = timed(s5, n5) / n5
            c5 = {"workload": "C5: 512^3 modes (1024^3 grid), Np = 2^27, ComplexF32, one GPU", "ms_per_step": ms5,
                  "value": 2.0 * C5_NP / (ms5 * 1e-3), "unit": UNIT, "steps": n5, "stage_ms": dict(p5.timer)}
            p5.close()
            del xs5, vp5, o5, w5
        except Exception as e:      # pragma: no cover
            c5 = {"error": str(e)[:200]}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, ms = cpu_reference_run(1, 0, NP_FULL)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": ms,
                   "sample": "ONE full-size step (set_points+type1, set_points+type2; 256^3 modes, 512^3 oversampled grid, "
                             "Np = 2^24), no warm-up; oracle port of the reference CPU algorithm (blocked, OpenMP + pocketfft)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: 3D 256^3 modes, Np=2^24 uniform-random points per GPU, ComplexF32, HalfSupport(4), "
                                   "sigma=2 (512^3 grid), BackwardsKaiserBessel+FastApproximation; step = set_points+type1, "
                                   "set_points+type2 (2*Np points per GPU per step)",
                       "l2": "inputs larger than L2 (1 GiB grid + 320 MiB points/values per transform)",
                       "multi_gpu": "points partitioned; type-1 partial outputs all-reduced (NCCL), type-2 spectrum broadcast"
                       if world > 1 else "single GPU"},
            "type1_points_per_s": npts * world / (ms_t1 * 1e-3), "type2_points_per_s": npts * world / (ms_t2 * 1e-3),
            "type1_ms": ms_t1, "type2_ms": ms_t2,
            "clustered": {"points": "Gaussian cloud, sigma = 1 rad (folded), Np = 2^24 per GPU", "type1_ms": ms_t1c, "type2_ms": ms_t2c,
                          "type1_points_per_s": npts * world / (ms_t1c * 1e-3), "type2_points_per_s": npts * world / (ms_t2c * 1e-3),
                          "stage_ms": timer_c},
            "c5_single_gpu": c5,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K,
                    "h2d_gb_per_s": h2d * K / (ms_e2e * 1e-3) / 1e9, "d2h_gb_per_s": d2h * K / (ms_e2e * 1e-3) / 1e9,
                    "note": "public API with pinned host buffers; per-step H2D of points+values+spectrum and D2H of both "
                            "results on copy streams, double-buffered and overlapped with the transforms; when ms_per_step "
                            "exceeds the device-resident step the sustained H2D rate (h2d_gb_per_s) is the bound: the host link"},
            "gpu_launches": int(launches),
        }
        print(json.dumps(line))
    try:
        plan.close()
    except Exception:
        pass
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="N = 1: skip the C5 single-GPU extra key")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
