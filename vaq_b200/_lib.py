"""ctypes binding of include/vaqgpu.h (vaq_b200/libvaqgpu.so).

There is no CPU path: if the shared library is missing this module raises at import of
the first symbol, and every compute entry point fails with VAQGPU_ECUDA on a machine
without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
# VAQGPU_LIB: development only (e.g. a -DVAQGPU_STATS build next to the release library)
LIB_PATH = Path(os.environ["VAQGPU_LIB"]) if os.environ.get("VAQGPU_LIB") else PKG / "libvaqgpu.so"

VAQGPU_OK, VAQGPU_EINVAL, VAQGPU_ECUDA, VAQGPU_ENOMEM, VAQGPU_ESTATE = 0, -1, -2, -3, -4
# search flags (include/vaqgpu.h; low byte == VAQ::NNMethod, reference VAQ.hpp:38-49)
EA, TI, HEAP, PROJECTED, SQRT, SCAN_V1, SCAN_F32 = 0x02, 0x04, 0x80, 0x100, 0x200, 0x1000, 0x2000


class VaqGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vaqgpu error {code}: {msg}")
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [("D", C.c_int32), ("M", C.c_int32), ("L", C.c_int32), ("bits", C.POINTER(C.c_int32)),
                ("centroids", C.POINTER(C.c_float)), ("eig_real", C.POINTER(C.c_float))]


_p = C.c_void_p
_i32, _i64, _u32, _u64, _f = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float

# name -> (restype, argtypes); every symbol include/vaqgpu.h declares
PROTOTYPES = {
    "vaqgpu_last_error": (C.c_char_p, []),
    "vaqgpu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "vaqgpu_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int, C.POINTER(_p)]),
    "vaqgpu_destroy": (None, [_p]),
    "vaqgpu_set_id_base": (C.c_int, [_p, _i64]),
    "vaqgpu_add_codes_u16": (C.c_int, [_p, _p, _i64]),
    "vaqgpu_encode_add": (C.c_int, [_p, _p, _i64]),
    "vaqgpu_add_codes_synthetic": (C.c_int, [_p, _i64, _u64, _p]),
    "vaqgpu_reserve": (C.c_int, [_p, _i64]),
    "vaqgpu_num_rows": (C.c_int, [_p, C.POINTER(_i64)]),
    "vaqgpu_row_bytes": (C.c_int, [_p, C.POINTER(_i32)]),
    "vaqgpu_get_codes_u16": (C.c_int, [_p, _i64, _i64, _p]),
    "vaqgpu_get_row_order": (C.c_int, [_p, _i64, _i64, _p]),
    "vaqgpu_build_lut": (C.c_int, [_p, _p, _i32, _p]),
    "vaqgpu_search": (C.c_int, [_p, _p, _i32, _i32, _u32, _p, _p]),
    "vaqgpu_search_device": (C.c_int, [_p, _p, _i32, _i32, _u32, _p, _p, _p]),
    "vaqgpu_search_keys_device": (C.c_int, [_p, _p, _i32, _i32, _u32, _p, _p]),
    "vaqgpu_merge_keys_device": (C.c_int, [_p, _i32, _i32, _i32, _u32, _p, _p, _p]),
    "vaqgpu_bounds_export": (C.c_int, [_p, _i32, _p, C.POINTER(_p)]),
    "vaqgpu_bounds_attach_ipc": (C.c_int, [_p, _i32, _p]),
    "vaqgpu_bounds_attach_ptr": (C.c_int, [_p, _i32, C.POINTER(_p)]),
    "vaqgpu_set_clusters": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p]),
    "vaqgpu_cluster_ti": (C.c_int, [_p, _i32, _i32, _i32]),
    "vaqgpu_get_clusters": (C.c_int, [_p, C.POINTER(_i32), C.POINTER(_i32), _p, _p, _p, _p]),
    "vaqgpu_set_visit": (C.c_int, [_p, _f]),
    "vaqgpu_set_cluster_rule_sizes": (C.c_int, [_p, _p]),
    "vaqgpu_set_raw_vectors": (C.c_int, [_p, _p, _i64, _i32]),
    "vaqgpu_refine": (C.c_int, [_p, _p, _i32, _p, _i32, _i32, _p, _p]),
    "vaqgpu_last_timings": (C.c_int, [_p, C.POINTER(_f)]),
    "vaqgpu_last_config": (C.c_int, [_p, C.POINTER(_i32)]),
    "hamgpu_create": (C.c_int, [_i32, C.c_int, C.POINTER(_p)]),
    "hamgpu_destroy": (None, [_p]),
    "hamgpu_set_id_base": (C.c_int, [_p, _i64]),
    "hamgpu_add": (C.c_int, [_p, _p, _i64]),
    "hamgpu_add_synthetic": (C.c_int, [_p, _i64, _u64]),
    "hamgpu_num_rows": (C.c_int, [_p, C.POINTER(_i64)]),
    "hamgpu_query": (C.c_int, [_p, _p, _i32, _i32, _p, _p]),
    "hamgpu_query_device": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p]),
    "hamgpu_query_keys_device": (C.c_int, [_p, _p, _i32, _i32, _p, _p]),
    "hamgpu_merge_keys_device": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p]),
    "hamgpu_last_timings": (C.c_int, [_p, C.POINTER(_f)]),
    "hamgpu_last_config": (C.c_int, [_p, C.POINTER(_i32)]),
    "vaqgpu_sharded_create": (C.c_int, [C.POINTER(ModelDesc), _i32, C.POINTER(C.c_int), _i64, C.POINTER(_p)]),
    "vaqgpu_sharded_destroy": (None, [_p]),
    "vaqgpu_sharded_add_codes_u16": (C.c_int, [_p, _p, _i64]),
    "vaqgpu_sharded_encode_add": (C.c_int, [_p, _p, _i64]),
    "vaqgpu_sharded_add_codes_synthetic": (C.c_int, [_p, _i64, _u64, _p]),
    "vaqgpu_sharded_num_shards": (C.c_int, [_p, C.POINTER(_i32)]),
    "vaqgpu_sharded_shard": (C.c_int, [_p, _i32, C.POINTER(_p)]),
    "vaqgpu_sharded_search": (C.c_int, [_p, _p, _i32, _i32, _u32, _p, _p]),
    "hamgpu_sharded_create": (C.c_int, [_i32, _i32, C.POINTER(C.c_int), _i64, C.POINTER(_p)]),
    "hamgpu_sharded_destroy": (None, [_p]),
    "hamgpu_sharded_add": (C.c_int, [_p, _p, _i64]),
    "hamgpu_sharded_add_synthetic": (C.c_int, [_p, _i64, _u64]),
    "hamgpu_sharded_query": (C.c_int, [_p, _p, _i32, _i32, _p, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libvaqgpu.so and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} is not built — run `python -m vaq_b200.build` (or __graft_entry__.build()). "
            "vaq_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise VaqGpuError(rc, load().vaqgpu_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    rc = load().vaqgpu_device_count(C.byref(n))
    return n.value if rc == 0 else 0
