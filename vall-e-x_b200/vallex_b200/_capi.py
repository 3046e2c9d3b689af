"""ctypes binding of libvallex_b200.so (include/vallex_b200.h).  No torch types cross this boundary:
only raw pointers and sizes.  There is no CPU fallback: if the library is missing or the device is not
a B200 the constructors raise."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvallex_b200.so")

VX_F32, VX_BF16 = 0, 1
VX_HOST, VX_DEVICE = 0, 1

_lib = None

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "vx_version": (C.c_char_p, []),
    "vx_last_error": (C.c_char_p, []),
    "vx_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vx_destroy": (None, [C.c_void_p]),
    "vx_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "vx_load_weight_mem": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "vx_finalize_weights": (C.c_int, [C.c_void_p]),
    "vx_set_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vx_ar_prefill": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vx_ar_fork": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vx_ar_select_beam": (C.c_int, [C.c_void_p, C.c_float, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "vx_ar_budget": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "vx_ar_decode": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_uint64, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vx_ar_decode_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_uint64, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "vx_slot_prefill": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vx_nar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "vx_vocos_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "vx_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vx_last_decode_kernel": (C.c_char_p, [C.c_void_p]),
    "vx_ar_ctx_clamped": (C.c_int, [C.c_void_p]),
    "vx_debug_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "vx_debug_gemm_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "vx_debug_attn_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vx_debug_attn_t5": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vx_encodec_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vx_encodec_destroy": (None, [C.c_void_p]),
    "vx_encodec_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "vx_encodec_finalize": (C.c_int, [C.c_void_p]),
    "vx_encodec_num_codebooks": (C.c_int, [C.c_void_p]),
    "vx_encodec_frames": (C.c_int64, [C.c_int64]),
    "vx_encodec_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "vx_encodec_embeddings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "vx_encodec_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


class VallexB200Error(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built: build with
    `python -c "import __graft_entry__ as g; g.build()"` or `make -C vall-e-x_b200/csrc`."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise VallexB200Error(f"{LIB_PATH} not found: the CUDA library must be built (no CPU fallback exists)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().vx_last_error()
        raise VallexB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Raw address of a torch tensor / numpy array (or None)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)
