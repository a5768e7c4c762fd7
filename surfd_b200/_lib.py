"""ctypes binding of the C-ABI in include/surfd_b200.h.

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing any
compute entry point raises.  (The oracle under oracle/ is test infrastructure and is never imported here.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SURFD_B200_LIB selects an alternative in-tree build of the same sources (diagnostic builds, e.g. -DMC_PROFILE)
LIB_PATH = os.environ.get("SURFD_B200_LIB") or os.path.join(_HERE, "_surfd_b200.so")

_lib = None

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/surfd_b200.h one to one
PROTOTYPES = {
    "surfd_version": (ctypes.c_int, []),
    "surfd_last_error": (ctypes.c_char_p, []),
    "surfd_launch_count": (c_i64, [ctypes.c_int]),
    "surfd_dec_create": (ctypes.c_int, [c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "surfd_dec_destroy": (None, [c_vp]),
    "surfd_dec_packed_floats": (ctypes.c_size_t, [ctypes.c_int]),
    "surfd_dec_set_latent": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "surfd_dec_set_precision": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "surfd_dec_set_sm_budget": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "surfd_dec_set_chain": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "surfd_dec_num_sms": (ctypes.c_int, [c_vp]),
    "surfd_dec_chunk_points": (ctypes.c_int, [c_vp]),
    "surfd_dec_profile": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_double)]),
    "surfd_dec_time_layer": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), c_vp]),
    "surfd_dec_debug_layer": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "surfd_udf_query": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "surfd_dec_logits": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "surfd_udf_lattice": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_vp, c_vp, ctypes.POINTER(c_i64), c_vp]),
    "surfd_mc_create": (ctypes.c_int, [ctypes.POINTER(c_vp)]),
    "surfd_mc_destroy": (None, [c_vp]),
    "surfd_mc_udf": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_vp]),
    "surfd_mc_launch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, c_vp]),
    "surfd_mc_profile": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "surfd_mc_finish": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "surfd_mc_fetch": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "surfd_mc_classify": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_vp, ctypes.POINTER(c_i64), c_vp]),
    "surfd_mc_time_classify": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), c_vp]),
    "surfd_face_filter": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, ctypes.c_int, c_vp, c_vp]),
    "surfd_unet_create": (ctypes.c_int, [c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "surfd_unet_destroy": (None, [c_vp]),
    "surfd_unet_set_lanes": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "surfd_unet_set_sampler": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int]),
    "surfd_unet_status": (ctypes.c_int, [c_vp]),
    "surfd_unet_profile": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_i64)]),
    "surfd_unet_set_precision": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "surfd_unet_packed_floats": (ctypes.c_size_t, []),
    "surfd_unet_forward": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "surfd_sample": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_float, c_vp, c_vp]),
    "surfd_obj_write": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, c_vp, c_i64, ctypes.c_int, ctypes.c_char_p, c_vp, c_i64, ctypes.c_char_p]),
    "surfd_obj_read": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_vp, c_vp]),
}

SURFD_OK, SURFD_EMPTY_SURFACE, SURFD_CAPACITY, SURFD_QUEUE_OVERFLOW, SURFD_BAD_ARGUMENT = 0, 1, 2, 3, 4


class SurfdError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"surfd_b200 status {status}: {message}")
        self.status = status


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m surfd_b200.build` "
            "(or __graft_entry__.build()).  surfd_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, allow=()):
    """Map a C status to an exception (ValueError / RuntimeError like the reference's wrappers)."""
    if status == 0 or status in allow:
        return status
    msg = load().surfd_last_error().decode("utf-8", "replace")
    if status == SURFD_BAD_ARGUMENT:
        raise ValueError(msg)
    if status == SURFD_EMPTY_SURFACE:
        raise RuntimeError("No surface found at the given iso value.")
    raise SurfdError(status, msg)


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
