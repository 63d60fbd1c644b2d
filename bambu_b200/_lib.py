"""ctypes binding of ``libbambu_em.so`` (the C-ABI of ``include/bambu_em.h``).

The product path has no CPU fallback: if the shared library is missing, or no
sm_100 device is usable, the calls raise."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BAMBU_EM_LIB") or os.path.join(_HERE, "libbambu_em.so")   # override: kernel experiments
_lib = None

OK, EINVAL, ENODEVICE, ECUDA, ENOMEM, ESTATE = 0, -1, -2, -3, -4, -5

# every symbol include/bambu_em.h declares: name -> (restype, argtypes)
_p = ctypes.c_void_p
_i32, _i64, _f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_double
SYMBOLS = {
    "bambu_em_version": (ctypes.c_char_p, []),
    "bambu_em_last_error": (ctypes.c_char_p, []),
    "bambu_em_device_count": (ctypes.c_int, []),
    "bambu_em_set_device": (ctypes.c_int, [ctypes.c_int]),
    "bambu_em_shutdown": (ctypes.c_int, []),
    "bambu_em_batched": (ctypes.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _f64, _f64, _p, _p, _p]),
    "bambu_em_batched_v2": (ctypes.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _i32, _f64, _f64, _p, _p, _p]),
    "bambu_wire_last_error": (ctypes.c_char_p, []),
    "bambu_wire_create": (ctypes.c_int, [ctypes.POINTER(_p), _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "bambu_wire_free": (ctypes.c_int, [_p]),
    "bambu_wire_sizes": (ctypes.c_int, [_p] + [ctypes.POINTER(_i64)] * 5 + [ctypes.POINTER(_i32), ctypes.POINTER(_i64)]),
    "bambu_wire_arrays": (ctypes.c_int, [_p] + [ctypes.POINTER(_p)] * 12),
    "bambu_em_plan_create_v2": (ctypes.c_int, [ctypes.POINTER(_p), _i64, _p, _p, _p]),
    "bambu_em_plan_upload_v2": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "bambu_emWithL1": (ctypes.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _f64, _f64, _p, _p]),
    "bambu_em_plan_create": (ctypes.c_int, [ctypes.POINTER(_p), _i64, _p, _p, _p]),
    "bambu_em_plan_reset": (ctypes.c_int, [_p, _i64, _p, _p, _p]),
    "bambu_em_plan_destroy": (ctypes.c_int, [_p]),
    "bambu_em_plan_set_stream": (ctypes.c_int, [_p, _p]),
    "bambu_em_plan_upload": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _p]),
    "bambu_em_plan_bind_device": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _p]),
    "bambu_em_plan_run": (ctypes.c_int, [_p, _i32, _f64, _f64]),
    "bambu_em_plan_download": (ctypes.c_int, [_p, _p, _p, _p, _p]),
    "bambu_em_plan_download_live": (ctypes.c_int, [_p, _p, _p, _p, _p]),
    "bambu_em_plan_sync": (ctypes.c_int, [_p]),
    "bambu_em_plan_device_outputs": (ctypes.c_int, [_p, ctypes.POINTER(_p), ctypes.POINTER(_p), ctypes.POINTER(_p)]),
    "bambu_em_plan_last_run_ms": (ctypes.c_int, [_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                                                  ctypes.POINTER(ctypes.c_float)]),
    "bambu_em_plan_last_run_launches": (ctypes.c_int, [_p, ctypes.POINTER(_i32)]),
    "bambu_em_plan_device_bytes": (ctypes.c_int, [_p, ctypes.POINTER(_i64)]),
    "bambu_quant_last_error": (ctypes.c_char_p, []),
    "bambu_quant_pack_create": (ctypes.c_int, [ctypes.POINTER(_p), _i64, _p, _p, _p, _p, _p, _p, _p, _i32]),
    "bambu_quant_pack_free": (ctypes.c_int, [_p]),
    "bambu_quant_pack_sizes": (ctypes.c_int, [_p] + [ctypes.POINTER(_i64)] * 6 + [ctypes.POINTER(_f64)]),
    "bambu_quant_pack_arena": (ctypes.c_int, [_p] + [ctypes.POINTER(_p)] * 10),
    "bambu_quant_pack_bookkeeping": (ctypes.c_int, [_p] + [ctypes.POINTER(_p)] * 3),
    "bambu_quant_merge": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, ctypes.POINTER(_i64)]),
    "bambu_quantDT": (ctypes.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _f64, _f64, _p, _p, _p, _p,
                                     ctypes.POINTER(_i64), ctypes.POINTER(_f64)]),
    "bambu_quant_last_used_device": (ctypes.c_int, []),
    "bambu_quantDT_hostpack": (ctypes.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _f64, _f64, _p, _p, _p, _p,
                                              ctypes.POINTER(_i64), ctypes.POINTER(_f64)]),
    "bambu_equirc_last_error": (ctypes.c_char_p, []),
    "bambu_equirc_create": (ctypes.c_int, [ctypes.POINTER(_p), _i64, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p]),
    "bambu_equirc_free": (ctypes.c_int, [_p]),
    "bambu_equirc_size": (ctypes.c_int, [_p, ctypes.POINTER(_i64)]),
    "bambu_equirc_table": (ctypes.c_int, [_p] + [ctypes.POINTER(_p)] * 8),
    "bambu_quantDT_cohort": (ctypes.c_int, [_i32, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _f64, _f64, _i64, _p, _p, _i32,
                                            _p, _p, _p, _p]),
    "bambu_quantDT_dotC": (None, [_p] * 19),
    "bambu_quant_pack_create_gpu": (ctypes.c_int, [ctypes.POINTER(_p), _i64, _p, _p, _p, _p, _p, _p, _p, _i32,
                                                   ctypes.POINTER(_i32)]),
}


class BambuEmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbambu_em error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile ``csrc/`` for sm_100a with nvcc (in-tree; seconds)."""
    csrc = os.path.join(_HERE, "csrc")
    subprocess.check_call(["make", "-C", csrc, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BambuEmError(ENODEVICE, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                          "g.build()'` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().bambu_em_last_error().decode() or lib().bambu_quant_last_error().decode()
        raise BambuEmError(rc, msg)


def check_quant(rc: int) -> None:
    if rc != 0:
        raise BambuEmError(rc, lib().bambu_quant_last_error().decode() or lib().bambu_em_last_error().decode())
