"""ctypes binding of include/oc_b200.h.  There is NO fallback: if the CUDA
library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB

OC_F32, OC_F64, OC_U8 = 0, 1, 2
OC_TRUNC_INLINE = 0
OC_TRUNC_SWEEP = 1
TRUNC = {"inline": OC_TRUNC_INLINE, "sweep": OC_TRUNC_SWEEP}
OC_MAX_L = 10
OC_MAX_SITES = 64

EXPORTS = [
    "oc_version", "oc_last_error", "oc_set_option", "oc_mps_layout", "oc_svals_width",
    "oc_encode_batch", "oc_overlap_batch", "oc_encode_classify", "oc_count_correct",
    "oc_pack_classifier", "oc_classify_host", "oc_fma_peak_launch", "oc_encode_step_ms",
    "oc_classifier_cache", "oc_release", "oc_overlap_batch_signed", "oc_label_mix", "oc_ensemble_vote",
    "oc_class_encode", "oc_add_compress_centre",
]


class OcError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB):
        raise ImportError(
            f"{LIB} is missing: the CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    p32 = C.POINTER(C.c_int32)
    p64 = C.POINTER(C.c_int64)
    lib.oc_version.restype = i32
    lib.oc_last_error.restype = C.c_char_p
    lib.oc_set_option.argtypes = [C.c_char_p, i32]
    lib.oc_mps_layout.argtypes = [i32, i32, p32, p64]
    lib.oc_svals_width.argtypes = [i32, i32]
    lib.oc_encode_batch.argtypes = [vp, i32, i64, i32, i64, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.oc_overlap_batch.argtypes = [vp, p32, vp, p32, p32, i32, i64, i32, vp, vp, vp]
    lib.oc_encode_classify.argtypes = [vp, i32, i64, i32, i64, i32, i32, i32, i32, vp, p32, p32,
                                       vp, vp, vp, vp]
    lib.oc_count_correct.argtypes = [vp, vp, i64, vp, vp]
    lib.oc_pack_classifier.argtypes = [C.POINTER(vp), p32, i32, i32, i32, vp, i64, p64, p32, p32]
    lib.oc_classify_host.argtypes = [vp, i32, i64, i32, i64, i32, i32, i32, i32, vp, i64, p32, p32,
                                     vp, vp, vp]
    lib.oc_overlap_batch_signed.argtypes = [vp, p32, vp, p32, p32, i32, i64, i32, vp, vp, vp]
    lib.oc_label_mix.argtypes = [vp, i64, i32, vp, i32, i32, i32, vp, vp]
    lib.oc_ensemble_vote.argtypes = [vp, i32, i64, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.oc_class_encode.argtypes = [vp, i32, i64, i32, p32, vp, i32, i32, vp, vp]
    lib.oc_add_compress_centre.argtypes = [vp, i64, i32, i32, i32, p32, i32, i32, i32, vp, p32, vp]
    lib.oc_classifier_cache.argtypes = [vp, i32, vp]
    lib.oc_release.argtypes = [vp]
    lib.oc_fma_peak_launch.argtypes = [i64, vp, C.POINTER(C.c_double), vp]
    lib.oc_encode_step_ms.argtypes = [C.POINTER(C.c_float), C.c_int]
    for name in EXPORTS:
        if name not in ("oc_last_error",):
            getattr(lib, name).restype = i32
    lib.oc_last_error.restype = C.c_char_p
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise OcError(f"oc_b200 error {rc}: {lib.oc_last_error().decode()}")


def i32_array(values):
    arr = (C.c_int32 * len(values))(*[int(v) for v in values])
    return arr
