"""ctypes binding of libnufft_b200.so (the C ABI declared in include/nufft_b200.h).

There is NO fallback: if the shared library is missing or cannot be loaded this module raises
ImportError (build it with ``python nonuniformffts.jl_b200/build.py`` or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
import os as _os
LIB_PATH = Path(_os.environ["NUFFT_B200_LIB"]) if _os.environ.get("NUFFT_B200_LIB") else _HERE / "libnufft_b200.so"

NUFFT_SUCCESS = 0
NUFFT_ERR_ARG, NUFFT_ERR_DIM, NUFFT_ERR_UNSUPPORTED, NUFFT_ERR_CUDA = -1, -2, -3, -4
NUFFT_ERR_CUFFT, NUFFT_ERR_ALLOC, NUFFT_ERR_STATE = -5, -6, -7
NUFFT_F32, NUFFT_F64 = 0, 1
KERNEL_IDS = {"kaiser_bessel": 0, "backwards_kaiser_bessel": 1, "gaussian": 2, "bspline": 3, "es": 4}
EVAL_IDS = {"fast": 0, "direct": 1}
METHOD_IDS = {"auto": 0, "global_memory": 1, "shared_memory": 2}


class nufft_opts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("dim", C.c_int32),
        ("n_modes", C.c_int64 * 3),
        ("is_complex", C.c_int32),
        ("dtype", C.c_int32),
        ("half_support", C.c_int32),
        ("sigma", C.c_double),
        ("kernel", C.c_int32),
        ("kernel_param", C.c_double),
        ("eval_mode", C.c_int32),
        ("ntransforms", C.c_int32),
        ("fftshift", C.c_int32),
        ("sort_points", C.c_int32),
        ("gpu_method", C.c_int32),
        ("block_dims", C.c_int64 * 3),
        ("point_convention", C.c_int32),
        ("device", C.c_int32),
        ("stream", C.c_void_p),
        ("record_timings", C.c_int32),
        ("spread_chunk", C.c_int32),
    ]


class nufft_callbacks(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("nu_weights", C.c_void_p),
        ("u_factor_sep", C.POINTER(C.c_void_p)),
        ("u_factor_dense", C.c_void_p),
        ("nvrtc_src", C.c_char_p),
        ("user_data", C.c_void_p),
    ]


# every symbol include/nufft_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_PP = C.POINTER(C.c_void_p)
SYMBOLS = {
    "nufft_opts_default": (C.c_int, [C.POINTER(nufft_opts)]),
    "nufft_plan_create": (C.c_int, [C.POINTER(_VP), C.POINTER(nufft_opts)]),
    "nufft_plan_destroy": (C.c_int, [_VP]),
    "nufft_plan_shape": (C.c_int, [_VP, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "nufft_plan_kernel_info": (C.c_int, [_VP, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), _VP, _VP]),
    "nufft_kernel_tables": (C.c_int, [C.POINTER(nufft_opts), C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                      _VP, C.c_size_t, _VP, C.c_size_t]),
    "nufft_set_points": (C.c_int, [_VP, C.c_int64, _PP]),
    "nufft_set_points_matrix": (C.c_int, [_VP, C.c_int64, _VP]),
    "nufft_get_binning": (C.c_int, [_VP, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nufft_get_binning_fine": (C.c_int, [_VP, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nufft_exec_type1": (C.c_int, [_VP, _PP, _PP, C.POINTER(nufft_callbacks)]),
    "nufft_exec_type2": (C.c_int, [_VP, _PP, _PP, C.POINTER(nufft_callbacks)]),
    "nufft_type1_spread": (C.c_int, [_VP, _PP, C.POINTER(nufft_callbacks)]),
    "nufft_type1_finish": (C.c_int, [_VP, _PP, C.POINTER(nufft_callbacks)]),
    "nufft_type2_prepare": (C.c_int, [_VP, _PP, C.POINTER(nufft_callbacks)]),
    "nufft_type2_interp": (C.c_int, [_VP, _PP, C.POINTER(nufft_callbacks)]),
    "nufft_get_grid": (C.c_int, [_VP, C.POINTER(_VP), C.POINTER(C.c_size_t)]),
    "nufft_get_timings": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "nufft_launch_count": (C.c_int64, [C.c_int]),
    "nufft_describe": (C.c_int, [_VP, C.c_char_p, C.c_size_t]),
    "nufft_last_error": (C.c_char_p, []),
    "nufft_abi_version": (C.c_int, []),
    # multi-GPU transforms (csrc/mgpu.cu)
    "nufft_mgpu_unique_id": (C.c_int, [_VP]),
    "nufft_mgpu_create": (C.c_int, [C.POINTER(_VP), C.POINTER(nufft_opts), C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int32), _VP, C.c_int32]),
    "nufft_mgpu_destroy": (C.c_int, [_VP]),
    "nufft_mgpu_info": (C.c_int, [_VP, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64)]),
    "nufft_mgpu_local_block": (C.c_int, [_VP, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nufft_mgpu_set_points": (C.c_int, [_VP, C.POINTER(C.c_int64), _PP]),
    "nufft_mgpu_exec_type1": (C.c_int, [_VP, _PP, _PP, C.POINTER(nufft_callbacks)]),
    "nufft_mgpu_exec_type2": (C.c_int, [_VP, _PP, _PP, C.POINTER(nufft_callbacks)]),
    "nufft_mgpu_gather_output": (C.c_int, [_VP, _PP, _PP]),
    "nufft_mgpu_synchronize": (C.c_int, [_VP]),
    "nufft_mgpu_get_stream": (C.c_int, [_VP, C.c_int32, C.POINTER(_VP)]),
    "nufft_mgpu_get_timings": (C.c_int, [_VP, C.c_int32, C.POINTER(C.c_float)]),
    "nufft_mgpu_exchange_mode": (C.c_int, [_VP, C.POINTER(C.c_int32)]),
}
MGPU_STRATEGIES = {"auto": 0, "slab": 1, "points": 2, "transforms": 3}
MGPU_ID_BYTES = 128

_lib = None


def load() -> C.CDLL:
    """Load libnufft_b200.so and bind every declared symbol.  Raises ImportError when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python nonuniformffts.jl_b200/build.py`). There is no CPU fallback.")
    try:
        lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    except OSError as e:  # pragma: no cover
        raise ImportError(f"cannot load {LIB_PATH}: {e}. There is no CPU fallback.") from e
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return (load().nufft_last_error() or b"").decode("utf-8", "replace")
