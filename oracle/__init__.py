"""CPU oracle for the bambu quantification hot path -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
as the timed CPU baseline.  Nothing under ``bambu_b200/`` imports it.

* ``em_oracle.c``  - C restatement of ``src/em.cpp`` (dense literal + sparse).
* ``quant_ref.py`` - loop-level restatement of the R packing around it
  (``R/bambu-quantify.R:53-74``, ``R/bambu-quantify_utilityFunctions.R:196-455``).

Parity is pinned by ``tests/test_oracle_golden.py`` against the reference's
golden tables (``R/sysdata.rda``) and the bundled chr9 end-to-end assays.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libem_oracle.so")
_lib = None

MODE_DENSE = 0        # literal dense sweeps (B1 without the theta_trace matrix)
MODE_DENSE_TRACE = 1  # + M x maxiter theta_trace allocation and copy (src/em.cpp:23,68)
MODE_SPARSE = 2       # same arithmetic on the non-zeros (B2)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "em_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        p = ctypes.c_void_p
        L.oracle_em_batched.restype = ctypes.c_int
        L.oracle_em_batched.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                        p, p, p, p, p, p, p, p, p,
                                        ctypes.c_int, ctypes.c_double, ctypes.c_double, p, p, p]
        L.oracle_emWithL1_dense.restype = ctypes.c_int
        L.oracle_emWithL1_dense.argtypes = [p, p, p, p, p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_double, ctypes.c_double, ctypes.c_int, p, p]
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def emWithL1(A, A_full, A_unique, Y, K, maxiter=10000, minvalue=1e-8, conv=1e-2, keep_trace=False):
    """Dense ``emWithL1`` (src/em.cpp:77-110) on explicit M x N matrices.

    Returns ``(est[3, M], iters)``."""
    A = np.asfortranarray(A, dtype=np.float64)
    Af = np.asfortranarray(A_full, dtype=np.float64)
    Au = np.asfortranarray(A_unique, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    K = np.ascontiguousarray(K, dtype=np.float64)
    M, N = A.shape
    est = np.empty((3, M), dtype=np.float64)
    fd = ctypes.c_double(0.0)
    it = lib().oracle_emWithL1_dense(_ptr(A), _ptr(Af), _ptr(Au), _ptr(Y), _ptr(K), M, N, int(maxiter),
                                     float(minvalue), float(conv), int(keep_trace), _ptr(est),
                                     ctypes.cast(ctypes.pointer(fd), ctypes.c_void_p))
    return est, int(it)


def em_batched(arena, maxiter=10000, minvalue=1e-8, conv=1e-2, mode=MODE_DENSE, n_threads=0):
    """``abundance_quantification`` over an arena (any object with the arena
    attributes of ``include/bambu_em.h``).  Returns ``(est[3, total_rows],
    iters[n_groups], status[n_groups])``."""
    g = np.ascontiguousarray(arena.grp_row_ptr, dtype=np.int64)
    c = np.ascontiguousarray(arena.grp_col_ptr, dtype=np.int64)
    r = np.ascontiguousarray(arena.row_nnz_ptr, dtype=np.int64)
    ci = np.ascontiguousarray(arena.col_idx, dtype=np.int32)
    av = np.ascontiguousarray(arena.aval, dtype=np.float64)
    nf = np.ascontiguousarray(arena.nz_flags, dtype=np.uint8)
    Y = np.ascontiguousarray(arena.Y, dtype=np.float64)
    K = np.ascontiguousarray(arena.K, dtype=np.float64)
    cf = np.ascontiguousarray(arena.col_flags, dtype=np.uint8)
    n_groups = len(g) - 1
    rows = int(g[-1])
    est = np.zeros((3, rows), dtype=np.float64)
    iters = np.zeros(n_groups, dtype=np.int32)
    status = np.zeros(n_groups, dtype=np.int32)
    rc = lib().oracle_em_batched(int(mode), int(n_threads), n_groups, _ptr(g), _ptr(c), _ptr(r), _ptr(ci),
                                 _ptr(av), _ptr(nf), _ptr(Y), _ptr(K), _ptr(cf), int(maxiter),
                                 float(minvalue), float(conv), _ptr(est), _ptr(iters), _ptr(status))
    assert rc == 0
    return est, iters, status
