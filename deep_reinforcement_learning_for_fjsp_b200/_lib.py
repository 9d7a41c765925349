"""ctypes binding of libfjsp_b200.so (include/fjsp_b200.h).  Fails loudly when the CUDA
extension is missing: there is no CPU path in the product."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FJSP_B200_LIB", os.path.join(HERE, "libfjsp_b200.so"))   # override: tuning builds
_lib = None

SYMBOLS = ["fjsp_last_error", "fjsp_abi_version", "fjsp_vec_create", "fjsp_vec_destroy", "fjsp_vec_query",
           "fjsp_vec_reset", "fjsp_vec_step", "fjsp_vec_step_host", "fjsp_vec_step_host_begin", "fjsp_vec_step_host_wait", "fjsp_vec_reset_host", "fjsp_vec_info", "fjsp_vec_trace", "fjsp_vec_slots"]


class ExtensionMissing(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissing(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  The FJSP vector environment has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    L.fjsp_last_error.restype = ctypes.c_char_p
    L.fjsp_abi_version.restype = i
    L.fjsp_vec_create.argtypes = [vp, vp, i, vp, i, i, i, i, vp]
    L.fjsp_vec_destroy.argtypes = [vp]
    L.fjsp_vec_query.argtypes = [vp, vp]
    L.fjsp_vec_reset.argtypes = [vp, vp, vp, vp]
    L.fjsp_vec_step.argtypes = [vp, vp, i, vp, vp, i, d, d, d, i, vp, vp, vp, vp, vp]
    L.fjsp_vec_step_host.argtypes = [vp, i, vp, vp, i, d, d, d, i, vp, vp, vp, vp, vp]
    L.fjsp_vec_step_host_begin.argtypes = [vp, i, vp, vp, i, d, d, d, i, vp, vp, vp, vp, vp]
    L.fjsp_vec_step_host_wait.argtypes = [vp]
    L.fjsp_vec_reset_host.argtypes = [vp, vp, vp]
    L.fjsp_vec_info.argtypes = [vp, vp]
    L.fjsp_vec_trace.argtypes = [vp, vp, i]
    L.fjsp_vec_slots.argtypes = [vp, vp, i]
    L.fjsp_vec_slots.restype = i
    L.fjsp_vec_trace.restype = i
    for f in ("fjsp_vec_create", "fjsp_vec_destroy", "fjsp_vec_query", "fjsp_vec_reset", "fjsp_vec_step",
              "fjsp_vec_step_host", "fjsp_vec_step_host_begin", "fjsp_vec_step_host_wait", "fjsp_vec_reset_host", "fjsp_vec_info"):
        getattr(L, f).restype = i
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError("fjsp_b200: %s (code %d)" % (load().fjsp_last_error().decode(), rc))
