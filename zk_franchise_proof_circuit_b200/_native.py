"""ctypes binding of libzkcensus_b200.so (the C ABI declared in include/zkcensus_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` / `csrc/Makefile`.  There is no
Python or CPU fallback: if the library is missing, or no sm_100 GPU is visible when a compute entry
point is called, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkcensus_b200.so")
_lib = None


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"zkcensus_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built - run __graft_entry__.build() (no CPU fallback exists)")
        L = ctypes.CDLL(LIB_PATH)
        vp, sz, i32, fp = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_float)
        L.zkb_last_error.restype = ctypes.c_char_p
        L.zkb_raw_field_op.argtypes = [i32, i32, vp, vp, vp, sz]
        L.zkb_bench_modmul.argtypes = [i32, i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.zkb_raw_ntt.argtypes = [vp, i32, i32, i32, fp]
        L.zkb_raw_coset_ntt.argtypes = [vp, i32, i32]
        L.zkb_raw_msm_g1.argtypes = [vp, sz, vp, i32, vp, fp, fp]
        L.zkb_raw_msm_g2.argtypes = [vp, sz, vp, i32, vp, fp, fp]
        u64p, pp = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_void_p)
        L.zkb_msm_session_create.argtypes = [i32, i32, i32, i32, ctypes.c_uint64, i32, pp]
        L.zkb_msm_session_destroy.argtypes = [vp]
        L.zkb_msm_session_destroy.restype = None
        L.zkb_msm_session_info.argtypes = [vp, u64p]
        L.zkb_msm_session_export.argtypes = [vp, vp]
        L.zkb_msm_session_attach.argtypes = [vp, vp]
        L.zkb_msm_session_attach_local.argtypes = [vp, vp]
        L.zkb_msm_session_run.argtypes = [vp, fp]
        L.zkb_msm_session_combine.argtypes = [vp, i32, vp, fp]
        L.zkb_msm_session_madds.argtypes = [vp, u64p]
        L.zkb_msm_session_read.argtypes = [vp, vp, vp]
        L.zkb_ntt_bench.argtypes = [i32, i32, i32, i32, fp, fp]
        L.zkb_ntt_dist_create.argtypes = [i32, i32, i32, i32, ctypes.c_uint64, pp]
        L.zkb_ntt_dist_destroy.argtypes = [vp]
        L.zkb_ntt_dist_destroy.restype = None
        L.zkb_ntt_dist_export.argtypes = [vp, vp]
        L.zkb_ntt_dist_attach.argtypes = [vp, i32, vp]
        L.zkb_ntt_dist_attach_local.argtypes = [vp, i32, vp]
        L.zkb_ntt_dist_fill.argtypes = [vp]
        L.zkb_ntt_dist_run.argtypes = [vp]
        L.zkb_ntt_dist_sync.argtypes = [vp, fp]
        L.zkb_ntt_dist_read.argtypes = [vp, i32, vp]
        L.zkb_raw_ntt_dif_forward.argtypes = [vp, i32]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise NativeError(rc, lib().zkb_last_error().decode(errors="replace"))
