"""
Build recipe of libnufft_b200.so (hand-written sm_100a CUDA kernels + C ABI).

    python nonuniformffts.jl_b200/build.py [--force] [--jobs N]

nvcc cross-compiles without a GPU.  The library is built IN-TREE
(nonuniformffts.jl_b200/libnufft_b200.so) so that it travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
# development: NUFFT_BUILD_TAG=<tag> builds a second library (libnufft_b200_<tag>.so, objects in build_<tag>/) with the extra
# -D flags of NUFFT_EXTRA_DEFS; NUFFT_B200_LIB selects it at run time (_lib.py)
TAG = os.environ.get("NUFFT_BUILD_TAG", "")
OBJ = HERE / ("build" + ("_" + TAG if TAG else ""))
LIB = HERE / ("libnufft_b200" + ("_" + TAG if TAG else "") + ".so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
DEV_M = os.environ.get("NUFFT_DEV_M")          # development only: instantiate a single half support
CFLAGS = ([f"-DNUFFT_DEV_M={DEV_M}"] if DEV_M else []) + os.environ.get("NUFFT_EXTRA_DEFS", "").split() + ["-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-O2",
          "-ccbin", shutil.which("g++") or "g++", "-Xptxas", "-v", "-Xfatbin", "-compress-all"]

# (source, object suffix, extra defines)
UNITS = [("api.cu", "", []), ("host_plan.cu", "", []), ("binning.cu", "", []), ("deconv.cu", "", []), ("pfft.cu", "", []),
         ("callbacks_jit.cu", "", []), ("mgpu.cu", "", []), ("ring_inst.cu", "", [])]
for t in ("float", "double"):
    for c in (0, 1):
        tag = f"_{'f32' if t == 'float' else 'f64'}_{'c' if c else 'r'}"
        UNITS.append(("spread_inst.cu", tag, [f"-DINST_T={t}", f"-DINST_CPLX={c}"]))
        UNITS.append(("interp_inst.cu", tag, [f"-DINST_T={t}", f"-DINST_CPLX={c}"]))


import re
import time

_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _unit_deps(src: Path, seen=None) -> set:
    """Files a translation unit depends on: the quoted includes, followed recursively."""
    seen = seen if seen is not None else set()
    if src in seen or not src.exists():
        return seen
    seen.add(src)
    for inc in _INC.findall(src.read_text()):
        _unit_deps((src.parent / inc).resolve(), seen)
    return seen


def _flags_stamp() -> str:
    return " ".join(ARCH + CFLAGS)


def _deps_mtime() -> float:
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "nufft_b200.h"]
    return max(f.stat().st_mtime for f in files)


def _compile(unit) -> tuple[str, str]:
    src, tag, defs = unit
    obj = OBJ / (Path(src).stem + tag + ".o")
    cmd = [NVCC, *ARCH, *CFLAGS, *defs, "-c", str(CSRC / src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}{tag}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return str(obj), r.stderr


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= _deps_mtime():
        return LIB
    OBJ.mkdir(exist_ok=True)
    started = time.time()
    jobs = jobs or min(len(UNITS), os.cpu_count() or 4)
    stamp = OBJ / "flags.stamp"
    if not stamp.exists() or stamp.read_text() != _flags_stamp():
        force = True
    todo, objs = [], []
    for u in UNITS:
        obj = OBJ / (Path(u[0]).stem + u[1] + ".o")
        objs.append(str(obj))
        dep_m = max(f.stat().st_mtime for f in _unit_deps((CSRC / u[0]).resolve()))
        if force or not obj.exists() or obj.stat().st_mtime < dep_m:
            todo.append(u)
    logs = []
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        for obj, log in ex.map(_compile, todo):
            logs.append(log)
            if verbose:
                print(f"[nvcc] {obj}", file=sys.stderr)
    with open(OBJ / "ptxas.log", "a" if len(todo) < len(UNITS) else "w") as f:
        f.write("\n".join(logs))
    stamp.write_text(_flags_stamp())
    cuda_lib = str(Path(NVCC).resolve().parent.parent / "lib64")
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *objs, "-L" + cuda_lib, "-lcufft", "-lnvrtc", "-ldl",
           "-Xlinker", "-rpath," + cuda_lib, "-ccbin", shutil.which("g++") or "g++"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if _deps_mtime() > started:          # a source changed while this build ran: the next call must look again
        os.utime(LIB, (started, started))
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    a = ap.parse_args()
    print(build(force=a.force, jobs=a.jobs, verbose=True))
