"""Host-side mirror of the reference's EM entry points, on top of the C-ABI.

    abundance_quantification(arena, maxiter, conv, minvalue)
        = R/bambu-quantify_utilityFunctions.R:379-391 for ALL gene groups at once
    emWithL1(A, A_full, A_unique, Y, K, maxiter, minvalue, conv)
        = R/RcppExports.R:12-14 / src/em.cpp:77-110 for one dense group
    EmPlan
        the same work with the arena resident in HBM (repeated runs, timing)

Everything here calls ``libbambu_em.so``; there is no CPU path."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .arena import Arena


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def abundance_quantification(arena: Arena, maxiter: int = 20000, conv: float = 1e-2, minvalue: float = 1e-8):
    """All gene groups of ``arena`` through one ``bambu_em_batched`` call (host
    buffers in, host buffers out).  Defaults are those of the reference's
    signature (R/bambu-quantify_utilityFunctions.R:379-380); ``bambu.quantDT``
    passes ``maxiter=10000`` (R/bambu_utilityFunctions.R:48-53).

    Returns ``(est[3, rows], iters[n_groups], status[n_groups])``; ``est`` rows
    are counts, fullLengthCounts, uniqueCounts in arena row order."""
    L = _lib.lib()
    g, c, r = _c(arena.grp_row_ptr, np.int64), _c(arena.grp_col_ptr, np.int64), _c(arena.row_nnz_ptr, np.int64)
    ci, av, nf = _c(arena.col_idx, np.int32), _c(arena.aval, np.float64), _c(arena.nz_flags, np.uint8)
    Y, K, cf = _c(arena.Y, np.float64), _c(arena.K, np.float64), _c(arena.col_flags, np.uint8)
    n_groups, rows = len(g) - 1, int(g[-1])
    est = np.zeros((3, rows), dtype=np.float64)
    iters = np.zeros(max(n_groups, 1), dtype=np.int32)
    status = np.zeros(max(n_groups, 1), dtype=np.int32)
    _lib.check(L.bambu_em_batched(n_groups, _ptr(g), _ptr(c), _ptr(r), _ptr(ci), _ptr(av), _ptr(nf), _ptr(Y),
                                  _ptr(K), _ptr(cf), int(maxiter), float(minvalue), float(conv), _ptr(est),
                                  _ptr(iters), _ptr(status)))
    return est, iters[:n_groups], status[:n_groups]


def abundance_quantification_into(arena: Arena, est, iters, status, maxiter=20000, conv=1e-2, minvalue=1e-8):
    """Same call with caller-provided HOST output buffers (e.g. pinned memory):
    ``est`` float64[3, rows], ``iters``/``status`` int32[n_groups]."""
    L = _lib.lib()
    g, c, r = _c(arena.grp_row_ptr, np.int64), _c(arena.grp_col_ptr, np.int64), _c(arena.row_nnz_ptr, np.int64)
    ci, av, nf = _c(arena.col_idx, np.int32), _c(arena.aval, np.float64), _c(arena.nz_flags, np.uint8)
    Y, K, cf = _c(arena.Y, np.float64), _c(arena.K, np.float64), _c(arena.col_flags, np.uint8)
    n_groups, rows = len(g) - 1, int(g[-1])
    if est.dtype != np.float64 or est.size < 3 * rows or not est.flags.c_contiguous:
        raise ValueError("est must be a contiguous float64 buffer of 3 x rows")
    if iters.dtype != np.int32 or status.dtype != np.int32 or iters.size < n_groups or status.size < n_groups:
        raise ValueError("iters/status must be int32 buffers of n_groups")
    _lib.check(L.bambu_em_batched(n_groups, _ptr(g), _ptr(c), _ptr(r), _ptr(ci), _ptr(av), _ptr(nf), _ptr(Y),
                                  _ptr(K), _ptr(cf), int(maxiter), float(minvalue), float(conv), _ptr(est),
                                  _ptr(iters), _ptr(status)))


def _wire_args(w):
    mode = int(w.col_mode)
    cdt = np.int32 if mode == 1 else np.float64
    bufs = (_c(w.row_end, np.uint32), _c(w.rs, np.float64), _c(w.aval, np.float64), _c(w.cidx, np.uint8),
            _c(w.eq_bits, np.uint32), _c(w.colY, cdt), _c(w.colK, cdt), _c(w.multi_bits, np.uint32))
    return bufs, mode


def abundance_quantification_wire(wire, maxiter: int = 20000, conv: float = 1e-2, minvalue: float = 1e-8, out=None):
    """``abundance_quantification`` fed with the wire form v2 (``bambu_b200.wire.Wire``): the same estimates,
    iteration counts and status words from about a third of the host->device bytes (``bambu_em_batched_v2``).
    ``out`` = optional ``(est, iters, status)`` HOST buffers (e.g. pinned)."""
    L = _lib.lib()
    g, c, e = _c(wire.grp_row_ptr, np.int64), _c(wire.grp_col_ptr, np.int64), _c(wire.grp_ent_ptr, np.int64)
    (re, rs, av, ci, eq, cy, ck, mb), mode = _wire_args(wire)
    n_groups, rows = len(g) - 1, int(g[-1])
    if out is None:
        est = np.zeros((3, rows), dtype=np.float64)
        iters = np.zeros(max(n_groups, 1), dtype=np.int32)
        status = np.zeros(max(n_groups, 1), dtype=np.int32)
    else:
        est, iters, status = out
        if est.dtype != np.float64 or est.size < 3 * rows or not est.flags.c_contiguous:
            raise ValueError("est must be a contiguous float64 buffer of 3 x rows")
        if iters.dtype != np.int32 or status.dtype != np.int32 or iters.size < n_groups or status.size < n_groups:
            raise ValueError("iters/status must be int32 buffers of n_groups")
    _lib.check(L.bambu_em_batched_v2(n_groups, _ptr(g), _ptr(c), _ptr(e), _ptr(re), _ptr(rs), _ptr(av), _ptr(ci),
                                     _ptr(eq), mode, _ptr(cy), _ptr(ck), _ptr(mb), int(maxiter), float(minvalue),
                                     float(conv), _ptr(est), _ptr(iters), _ptr(status)))
    return est, iters[:n_groups], status[:n_groups]


def emWithL1(A, A_full, A_unique, Y, K, maxiter: int = 10000, minvalue: float = 1e-8, conv: float = 1e-2):
    """One dense group, the signature of the reference's ``emWithL1``
    (R/RcppExports.R:12).  Returns ``{"theta": 3 x M}`` like the reference
    (src/em.cpp:104-107) plus ``"iter"`` (which the reference computes and drops)."""
    L = _lib.lib()
    A = np.asfortranarray(A, dtype=np.float64)
    Af = np.asfortranarray(A_full, dtype=np.float64)
    Au = np.asfortranarray(A_unique, dtype=np.float64)
    if A.ndim != 2 or Af.shape != A.shape or Au.shape != A.shape:
        raise ValueError("A, A_full, A_unique must be M x N matrices of one shape")
    M, N = A.shape
    Y, K = _c(Y, np.float64), _c(K, np.float64)
    if len(Y) != N or len(K) != N:
        raise ValueError("Y and K must have one entry per column of A")
    est = np.zeros((3, M), dtype=np.float64, order="F")
    it = ctypes.c_int32(0)
    _lib.check(L.bambu_emWithL1(_ptr(A), _ptr(Af), _ptr(Au), _ptr(Y), _ptr(K), M, N, int(maxiter), float(minvalue),
                                float(conv), _ptr(est), ctypes.cast(ctypes.pointer(it), ctypes.c_void_p)))
    return {"theta": np.ascontiguousarray(est), "iter": int(it.value)}


class EmPlan:
    """Arena resident in HBM (``bambu_em_plan_*``)."""

    def __init__(self, arena):
        """``arena``: an ``Arena`` (v1) or a ``bambu_b200.wire.Wire`` (v2, ``bambu_em_plan_create_v2``)."""
        self._L = _lib.lib()
        self.arena = arena
        self.is_wire = hasattr(arena, "grp_ent_ptr")
        self._g = _c(arena.grp_row_ptr, np.int64)
        self._c = _c(arena.grp_col_ptr, np.int64)
        self._r = _c(arena.grp_ent_ptr if self.is_wire else arena.row_nnz_ptr, np.int64)
        self.n_groups, self.rows = len(self._g) - 1, int(self._g[-1])
        self._h = ctypes.c_void_p()
        create = self._L.bambu_em_plan_create_v2 if self.is_wire else self._L.bambu_em_plan_create
        _lib.check(create(ctypes.byref(self._h), self.n_groups, _ptr(self._g), _ptr(self._c), _ptr(self._r)))
        self._keep = None

    def reset(self, arena: Arena):
        """Re-analyse another arena in this plan (``bambu_em_plan_reset``): device buffers are kept and only grow.
        The payload has to be uploaded / bound again."""
        self.arena = arena
        self._g = _c(arena.grp_row_ptr, np.int64)
        self._c = _c(arena.grp_col_ptr, np.int64)
        self._r = _c(arena.row_nnz_ptr, np.int64)
        self.n_groups, self.rows = len(self._g) - 1, int(self._g[-1])
        _lib.check(self._L.bambu_em_plan_reset(self._h, self.n_groups, _ptr(self._g), _ptr(self._c), _ptr(self._r)))
        self._keep = None

    def set_stream(self, cuda_stream: int | None):
        _lib.check(self._L.bambu_em_plan_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def upload(self, col_idx=None, aval=None, nz_flags=None, Y=None, K=None, col_flags=None):
        """H2D of the payload (defaults: the arena's arrays).  The arrays are
        kept referenced until the next upload (the copy is asynchronous)."""
        a = self.arena
        if self.is_wire:
            bufs, mode = _wire_args(a)
            self._keep = bufs
            re, rs, av, ci, eq, cy, ck, mb = bufs
            _lib.check(self._L.bambu_em_plan_upload_v2(self._h, _ptr(re), _ptr(rs), _ptr(av), _ptr(ci), _ptr(eq), mode,
                                                       _ptr(cy), _ptr(ck), _ptr(mb)))
            return
        bufs = (_c(a.col_idx if col_idx is None else col_idx, np.int32),
                _c(a.aval if aval is None else aval, np.float64),
                _c(a.nz_flags if nz_flags is None else nz_flags, np.uint8),
                _c(a.Y if Y is None else Y, np.float64), _c(a.K if K is None else K, np.float64),
                _c(a.col_flags if col_flags is None else col_flags, np.uint8))
        self._keep = bufs
        _lib.check(self._L.bambu_em_plan_upload(self._h, *[_ptr(b) for b in bufs]))

    def upload_ptrs(self, col_idx, aval, nz_flags, Y, K, col_flags):
        """H2D from raw host addresses (e.g. pinned torch tensors)."""
        _lib.check(self._L.bambu_em_plan_upload(self._h, *[ctypes.c_void_p(int(x)) for x in
                                                           (col_idx, aval, nz_flags, Y, K, col_flags)]))

    def upload_ptrs_v2(self, row_end, rs, aval, cidx, eq_bits, col_mode, colY, colK, multi_bits):
        """H2D of a wire-form payload from raw host addresses (e.g. pinned torch tensors)."""
        v = ctypes.c_void_p
        _lib.check(self._L.bambu_em_plan_upload_v2(self._h, v(int(row_end)), v(int(rs)), v(int(aval)), v(int(cidx)),
                                                   v(int(eq_bits)), int(col_mode), v(int(colY)), v(int(colK)),
                                                   v(int(multi_bits))))

    def bind_device(self, col_idx, aval, nz_flags, Y, K, col_flags):
        """Use device buffers (raw device addresses) owned by the caller."""
        _lib.check(self._L.bambu_em_plan_bind_device(self._h, *[ctypes.c_void_p(int(x)) for x in
                                                                (col_idx, aval, nz_flags, Y, K, col_flags)]))

    def run(self, maxiter=10000, minvalue=1e-8, conv=1e-2):
        _lib.check(self._L.bambu_em_plan_run(self._h, int(maxiter), float(minvalue), float(conv)))

    def sync(self):
        _lib.check(self._L.bambu_em_plan_sync(self._h))

    def download(self, want_nnz=False, out=None):
        if out is None:
            est = np.zeros((3, self.rows), dtype=np.float64)
            iters = np.zeros(max(self.n_groups, 1), dtype=np.int32)
            status = np.zeros(max(self.n_groups, 1), dtype=np.int32)
        else:
            est, iters, status = out
        nz = np.zeros(max(self.n_groups, 1), dtype=np.int64) if want_nnz else None
        _lib.check(self._L.bambu_em_plan_download(self._h, _ptr(est), _ptr(iters), _ptr(status), _ptr(nz)))
        res = (est, iters[:self.n_groups], status[:self.n_groups])
        return res + (nz[:self.n_groups],) if want_nnz else res

    def download_live(self):
        """(live columns, live entries) per group, as kept by the pack kernels."""
        nl = np.zeros(max(self.n_groups, 1), dtype=np.int32)
        nz = np.zeros(max(self.n_groups, 1), dtype=np.int32)
        _lib.check(self._L.bambu_em_plan_download_live(self._h, _ptr(nl), _ptr(nz), None, None))
        return nl[:self.n_groups], nz[:self.n_groups]

    def download_image_stats(self):
        """(shared-memory bytes of the EM, 32-lane slots) per group."""
        sm = np.zeros(max(self.n_groups, 1), dtype=np.int32)
        sl = np.zeros(max(self.n_groups, 1), dtype=np.int32)
        _lib.check(self._L.bambu_em_plan_download_live(self._h, None, None, _ptr(sm), _ptr(sl)))
        return sm[:self.n_groups], sl[:self.n_groups]

    def download_ptrs(self, est, iters, status):
        _lib.check(self._L.bambu_em_plan_download(self._h, ctypes.c_void_p(int(est)), ctypes.c_void_p(int(iters)),
                                                  ctypes.c_void_p(int(status)), None))

    def last_run_ms(self):
        a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        _lib.check(self._L.bambu_em_plan_last_run_ms(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"pack_ms": a.value, "em_ms": b.value, "total_ms": c.value}

    def last_run_launches(self) -> int:
        n = ctypes.c_int32()
        _lib.check(self._L.bambu_em_plan_last_run_launches(self._h, ctypes.byref(n)))
        return int(n.value)

    def device_bytes(self) -> int:
        n = ctypes.c_int64()
        _lib.check(self._L.bambu_em_plan_device_bytes(self._h, ctypes.byref(n)))
        return int(n.value)

    def close(self):
        if self._h:
            self._L.bambu_em_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
