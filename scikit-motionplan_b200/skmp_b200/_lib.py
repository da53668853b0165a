"""ctypes binding of the C-ABI shared library (include/skmp_b200.h).

The library is built in-tree (csrc/Makefile -> skmp_b200/libskmp_b200.so).  There is no CPU
fallback: if the library is missing, or no CUDA device is visible when a compute entry point is
called, a RuntimeError is raised.
"""
import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libskmp_b200.so"

F32, F64 = 0, 1
MODE_ALL, MODE_CLOSEST = 0, 1
COM_POINT, COM_STABILITY = 0, 1

_lib = None


class SkbError(RuntimeError):
    pass


def _declare(lib):
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pi32, pdbl, pvp = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_void_p)
    sig = {
        "skb_version": ([], C.c_int),
        "skb_last_error": ([], C.c_char_p),
        "skb_device_count": ([pi32], C.c_int),
        "skb_launch_count": ([], i64),
        "skb_model_create": ([i32, pi32, pi32, pdbl, pdbl, pvp], C.c_int),
        "skb_model_add_link": ([vp, i32, pdbl, pdbl, pi32], C.c_int),
        "skb_model_set_q": ([vp, pi32, i32, pdbl, i32], C.c_int),
        "skb_model_get_state": ([vp, pdbl, pdbl], C.c_int),
        "skb_model_num_links": ([vp, pi32], C.c_int),
        "skb_model_clone": ([vp, pvp], C.c_int),
        "skb_model_destroy": ([vp], None),
        "skb_sdf_create": ([pvp], C.c_int),
        "skb_sdf_add_box": ([vp, pdbl, pdbl, pdbl], C.c_int),
        "skb_sdf_add_cylinder": ([vp, pdbl, pdbl, dbl, dbl], C.c_int),
        "skb_sdf_add_sphere": ([vp, pdbl, dbl], C.c_int),
        "skb_sdf_add_grid": ([vp, pdbl, pdbl, pdbl, pdbl, pi32, dbl, vp, i32], C.c_int),
        "skb_sdf_num_prims": ([vp, pi32], C.c_int),
        "skb_sdf_destroy": ([vp], None),
        "skb_sdf_eval": ([vp, i32, vp, i64, vp, vp, vp], C.c_int),
        "skb_plan_create": ([vp, pi32, i32, pi32, i32, i32, i32, pvp], C.c_int),
        "skb_plan_dims": ([vp, pi32, pi32, pi32], C.c_int),
        "skb_plan_set_radii": ([vp, pdbl], C.c_int),
        "skb_plan_set_pairs": ([vp, pi32, i32, pdbl], C.c_int),
        "skb_plan_set_weights": ([vp, pdbl], C.c_int),
        "skb_plan_set_tuning": ([vp, i32, i32, i32, i32], C.c_int),
        "skb_plan_tune": ([vp, vp, i32, vp, i64, dbl, i32, i32, vp, pi32], C.c_int),
        "skb_plan_get_tuned": ([vp, vp, i32, i32, i32, pi32], C.c_int),
        "skb_plan_destroy": ([vp], None),
        "skb_fk": ([vp, i32, vp, i64, i32, vp, vp, vp], C.c_int),
        "skb_collfree": ([vp, vp, i32, vp, i64, dbl, i32, vp, vp, vp, vp], C.c_int),
        "skb_collfree_csc": ([vp, vp, i32, vp, i64, dbl, vp, vp, vp], C.c_int),
        "skb_pairwise": ([vp, i32, vp, i64, i32, vp, vp, vp, vp], C.c_int),
        "skb_com": ([vp, vp, i32, vp, i64, i32, vp, vp, vp], C.c_int),
        "skb_com_host": ([vp, vp, i32, vp, i64, i32, vp, vp], C.c_int),
        "skb_edge_counts": ([vp, vp, vp, i64, i32, i32, vp, vp], C.c_int),
        "skb_edge_samples": ([vp, vp, vp, i64, i32, i32, vp, i32, vp, vp], C.c_int),
        "skb_segment_all": ([vp, vp, i64, vp, vp], C.c_int),
        "skb_fma_peak": ([i32, i32, vp], C.c_int),
        "skb_store_peak": ([i32, i64, i32, vp], C.c_int),
        "skb_lm_step": ([i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, vp], C.c_int),
        "skb_fk_host": ([vp, i32, vp, i64, i32, vp, vp], C.c_int),
        "skb_collfree_host": ([vp, vp, i32, vp, i64, dbl, i32, vp, vp, vp], C.c_int),
        "skb_pairwise_host": ([vp, i32, vp, i64, i32, vp, vp, vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    return sig


EXPORTED_SYMBOLS = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib, EXPORTED_SYMBOLS
    if _lib is None:
        path = os.environ.get("SKMP_B200_LIB", str(LIB_PATH))
        if not Path(path).exists():
            raise SkbError(
                f"{path} not found: build the CUDA extension first (python __graft_entry__.py build). "
                "skmp_b200 has no CPU fallback."
            )
        handle = C.CDLL(path)
        EXPORTED_SYMBOLS = sorted(_declare(handle))
        _lib = handle
    return _lib


def check(status: int):
    if status != 0:
        raise SkbError(lib().skb_last_error().decode())


def require_cuda() -> int:
    n = C.c_int32(0)
    if lib().skb_device_count(C.byref(n)) != 0:
        raise SkbError("skmp_b200 needs a CUDA device (no CPU fallback): " + lib().skb_last_error().decode())
    return n.value


def dptr(arr, ctype=C.c_double):
    """numpy array -> typed ctypes pointer (array must stay alive during the call)."""
    return arr.ctypes.data_as(C.POINTER(ctype))
