"""CPU test: libnufft_b200.so loads and exports every symbol include/nufft_b200.h declares (no compute calls),
and the ctypes structs mirror the C structs."""
import ctypes as C
import re

import pytest
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "nufft_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nufft_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_boundary():
    names = declared_functions()
    for required in ("nufft_plan_create", "nufft_plan_destroy", "nufft_set_points", "nufft_exec_type1", "nufft_exec_type2",
                     "nufft_plan_shape", "nufft_get_binning", "nufft_get_timings", "nufft_describe", "nufft_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol():
    import nufft_b200  # noqa: F401
    from nufft_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding and header disagree"
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/nufft_b200.h but not exported"
    assert lib.nufft_abi_version() == 2


def test_opts_struct_layout_and_defaults():
    import nufft_b200  # noqa: F401
    from nufft_b200 import _lib
    lib = _lib.load()
    o = _lib.nufft_opts()
    assert lib.nufft_opts_default(C.byref(o)) == 0
    # the library writes sizeof(nufft_opts) as it sees it: must equal the ctypes mirror
    assert o.struct_size == C.sizeof(_lib.nufft_opts)
    assert (o.half_support, o.sigma, o.ntransforms) == (4, 2.0, 1)          # src/plan.jl:573,583
    assert o.kernel == _lib.KERNEL_IDS["kaiser_bessel"] and o.eval_mode == _lib.EVAL_IDS["direct"]  # CUDA ext defaults
    # a wrong struct_size is rejected without touching the GPU
    o.struct_size = 8
    h = C.c_void_p()
    assert lib.nufft_plan_create(C.byref(h), C.byref(o)) == _lib.NUFFT_ERR_ARG
    assert b"ABI mismatch" in lib.nufft_last_error()


def test_no_cpu_fallback_in_product_package():
    """The product package must not import the oracle (or any CPU compute path)."""
    pkg = ROOT / "nonuniformffts.jl_b200"
    for f in pkg.glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_nfft_accuracy_params():
    """AbstractNFFTs.accuracyParams restated in the host mirror (no GPU needed)."""
    import nufft_b200 as nb
    assert nb.accuracy_params() == (4, 2.0, 1e-7)
    m, sigma, reltol = nb.accuracy_params(reltol=1e-9)
    assert (m, sigma) == (5, 2.0) and reltol == 1e-9
    assert nb.accuracy_params(m=6, sigma=1.5)[:2] == (6, 1.5)


def test_mgpu_argument_validation_without_a_gpu():
    """nufft_mgpu_create rejects bad rank counts / strategies / struct sizes before it touches CUDA or NCCL."""
    import nufft_b200  # noqa: F401
    from nufft_b200 import _lib
    lib = _lib.load()
    o = _lib.nufft_opts()
    assert lib.nufft_opts_default(C.byref(o)) == 0
    h = C.c_void_p()
    one = (C.c_int32 * 1)(0)
    idbuf = (C.c_ubyte * _lib.MGPU_ID_BYTES)()
    assert lib.nufft_mgpu_create(C.byref(h), C.byref(o), 0, 1, one, one, idbuf, 0) == _lib.NUFFT_ERR_ARG        # nranks < 1
    assert lib.nufft_mgpu_create(C.byref(h), C.byref(o), 64, 1, one, one, idbuf, 0) == _lib.NUFFT_ERR_ARG       # nranks > 16
    assert lib.nufft_mgpu_create(C.byref(h), C.byref(o), 2, 3, one, one, idbuf, 0) == _lib.NUFFT_ERR_ARG        # nlocal > nranks
    assert lib.nufft_mgpu_create(C.byref(h), C.byref(o), 1, 1, one, one, idbuf, 9) == _lib.NUFFT_ERR_ARG        # unknown strategy
    o.struct_size = 8
    assert lib.nufft_mgpu_create(C.byref(h), C.byref(o), 1, 1, one, one, idbuf, 0) == _lib.NUFFT_ERR_ARG
    assert h.value is None
    assert lib.nufft_mgpu_destroy(None) == 0
    assert lib.nufft_mgpu_synchronize(None) == _lib.NUFFT_ERR_STATE
    assert _lib.MGPU_STRATEGIES == {"auto": 0, "slab": 1, "points": 2, "transforms": 3}


def test_header_is_plain_c99(tmp_path):
    """The drop-in boundary is a C ABI: include/nufft_b200.h must compile as C99 without warnings (no C++, no torch types)."""
    import shutil
    import subprocess
    from pathlib import Path
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = Path(__file__).resolve().parent.parent
    src = tmp_path / "hdr.c"
    src.write_text('#include "nufft_b200.h"\nint main(void) { nufft_opts o; nufft_callbacks c; (void)o; (void)c; '
                   'return NUFFT_KERNEL_ES == 4 && NUFFT_MGPU_SLAB == 1 ? 0 : 1; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(root / "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
