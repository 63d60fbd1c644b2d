"""Wire form v2 of the arena ("live arena"): what ``bambu_em_batched_v2`` copies to the GPU.

The EM never reads a column with ``Y == 0`` or an entry with ``aval == 0`` (``src/em.cpp:48-49, 98-99``: such
terms are +-0 or NaN->0), and the only thing those entries feed is ``rs_i = sum_j X_ij`` (``src/em.cpp:33``).  The
v2 packer therefore sends the live entries plus ``rs`` -- about 37 % of the v1 bytes on the synthetic cohorts --
with 16-bit column indices, bit-packed flags and integer counts (``Y = nobs / K`` is formed on the device exactly as
``R/bambu-quantify_utilityFunctions.R:221-228`` forms it).  Layout: ``include/bambu_em.h``.

``Wire.from_arena`` is native host code (``csrc/wire.cpp``); ``wire_from_arena_numpy`` is its numpy mirror (tests).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .arena import Arena

PAYLOAD = ("row_end", "rs", "aval", "cidx", "eq_bits", "colY", "colK", "multi_bits")


@dataclass
class Wire:
    grp_row_ptr: np.ndarray     # int64 [G+1]
    grp_col_ptr: np.ndarray     # int64 [G+1]   live columns
    grp_ent_ptr: np.ndarray     # int64 [G+1]   live entries
    row_end: np.ndarray         # uint32 [rows]
    rs: np.ndarray              # float64 [rows]
    aval: np.ndarray            # float64 [entries]
    cidx: np.ndarray            # uint8 blob (u16 / u32 index blocks, 4-byte aligned per group)
    eq_bits: np.ndarray         # uint32
    col_mode: int               # 1: colY/colK int32 (nobs, K); 0: float64 (Y, K)
    colY: np.ndarray
    colK: np.ndarray
    multi_bits: np.ndarray      # uint32
    nnz_nonzero: np.ndarray     # int64 [G]  aval != 0 entries of the ORIGINAL group (metric; not sent)
    meta: dict = field(default_factory=dict)

    @property
    def n_groups(self) -> int:
        return len(self.grp_row_ptr) - 1

    @property
    def n_rows(self) -> int:
        return int(self.grp_row_ptr[-1])

    @property
    def n_cols(self) -> int:
        return int(self.grp_col_ptr[-1])

    @property
    def nnz(self) -> int:
        return int(self.grp_ent_ptr[-1])

    def nbytes(self) -> int:
        """Bytes ``bambu_em_batched_v2`` copies host -> device."""
        return sum(getattr(self, n).nbytes for n in PAYLOAD + ("grp_row_ptr", "grp_col_ptr", "grp_ent_ptr"))

    @staticmethod
    def from_arena(a: Arena) -> "Wire":
        L = _lib.lib()
        c = lambda x, dt: np.ascontiguousarray(x, dtype=dt)     # noqa: E731
        g, cp, r = c(a.grp_row_ptr, np.int64), c(a.grp_col_ptr, np.int64), c(a.row_nnz_ptr, np.int64)
        ci, av, nf = c(a.col_idx, np.int32), c(a.aval, np.float64), c(a.nz_flags, np.uint8)
        Y, K, cf = c(a.Y, np.float64), c(a.K, np.float64), c(a.col_flags, np.uint8)
        p = lambda x: x.ctypes.data_as(ctypes.c_void_p)         # noqa: E731
        h = ctypes.c_void_p()
        rc = L.bambu_wire_create(ctypes.byref(h), len(g) - 1, p(g), p(cp), p(r), p(ci), p(av), p(nf), p(Y), p(K), p(cf))
        if rc:
            raise _lib.BambuEmError(rc, L.bambu_wire_last_error().decode())
        try:
            n = [ctypes.c_int64() for _ in range(5)]
            mode, tot = ctypes.c_int32(), ctypes.c_int64()
            L.bambu_wire_sizes(h, *[ctypes.byref(x) for x in n], ctypes.byref(mode), ctypes.byref(tot))
            G, rows, cols, ents, ib = (int(x.value) for x in n)
            ptr = [ctypes.c_void_p() for _ in range(12)]
            L.bambu_wire_arrays(h, *[ctypes.byref(x) for x in ptr])

            def arr(q, count, dt):
                if count == 0:
                    return np.zeros(0, dt)
                buf = (ctypes.c_char * (count * np.dtype(dt).itemsize)).from_address(q.value)
                return np.frombuffer(buf, dtype=dt).copy()
            cdt = np.int32 if mode.value == 1 else np.float64
            w = Wire(arr(ptr[0], G + 1, np.int64), arr(ptr[1], G + 1, np.int64), arr(ptr[2], G + 1, np.int64),
                     arr(ptr[3], rows, np.uint32), arr(ptr[4], rows, np.float64), arr(ptr[5], ents, np.float64),
                     arr(ptr[6], ib, np.uint8), arr(ptr[7], (ents + 31) // 32 + 1, np.uint32), int(mode.value),
                     arr(ptr[8], cols, cdt), arr(ptr[9], cols, cdt), arr(ptr[10], (cols + 31) // 32 + 1, np.uint32),
                     arr(ptr[11], G, np.int64))
        finally:
            L.bambu_wire_free(h)
        w.meta = dict(getattr(a, "meta", {}) or {})
        return w


def _idx_width(NL):
    return np.where(NL <= 32767, 2, 4)


def wire_from_arena_numpy(a: Arena) -> Wire:
    """numpy mirror of ``csrc/wire.cpp`` (tests pin the native converter on it, array for array)."""
    G = a.n_groups
    M, N, _ = a.group_shapes()
    rows_of = np.repeat(np.arange(a.n_rows), np.diff(a.row_nnz_ptr))
    grp_of_row = np.repeat(np.arange(G), M)
    grp_of_ent = grp_of_row[rows_of]
    col_abs = a.grp_col_ptr[grp_of_ent] + a.col_idx
    live_col = a.Y != 0                                  # NaN != 0 is True: live, like the device
    grp_of_col = np.repeat(np.arange(G), N)
    nl_per_grp = np.bincount(grp_of_col, weights=live_col, minlength=G).astype(np.int64)
    grp_col_ptr = np.concatenate(([0], np.cumsum(nl_per_grp))).astype(np.int64)
    live_rank = np.cumsum(live_col) - 1                  # global live index
    live_local = live_rank - grp_col_ptr[grp_of_col]     # local to the group
    keep = live_col[col_abs] & (a.aval != 0)
    # rs: two accumulators by column parity, each in ascending column order (row order of the arena)
    rs = np.zeros(a.n_rows)
    for par in (0, 1):
        sel = (a.col_idx & 1) == par
        acc = np.zeros(a.n_rows)
        # sequential adds in entry order; np.add.at applies them in index order
        np.add.at(acc, rows_of[sel], a.aval[sel])
        rs = acc if par == 0 else rs + acc
    ent_per_grp = np.bincount(grp_of_ent, weights=keep, minlength=G).astype(np.int64)
    grp_ent_ptr = np.concatenate(([0], np.cumsum(ent_per_grp))).astype(np.int64)
    ent_per_row = np.bincount(rows_of, weights=keep, minlength=a.n_rows).astype(np.int64)
    row_abs_end = np.cumsum(ent_per_row)
    row_end = (row_abs_end - grp_ent_ptr[grp_of_row]).astype(np.uint32)
    idx = ((live_local[col_abs[keep]] << 1) | (a.col_idx[keep] & 1)).astype(np.uint32)
    width = _idx_width(nl_per_grp)
    blob = bytearray()
    for g in range(G):
        while len(blob) % 4:
            blob.append(0)
        seg = idx[grp_ent_ptr[g]:grp_ent_ptr[g + 1]]
        blob += (seg.astype(np.uint16) if width[g] == 2 else seg).tobytes()
    while len(blob) % 4:
        blob.append(0)
    blob += b"\0\0\0\0"
    ne, nl = int(grp_ent_ptr[-1]), int(grp_col_ptr[-1])

    def bits(flags, n):
        out = np.zeros((n + 31) // 32 + 1, np.uint32)
        k = np.nonzero(flags)[0]
        np.bitwise_or.at(out, k >> 5, (np.uint32(1) << (k & 31).astype(np.uint32)))
        return out
    eq_bits = bits((a.nz_flags[keep] & 1) != 0, ne)
    multi_bits = bits((a.col_flags[live_col] & 1) != 0, nl)
    Yl, Kl = a.Y[live_col], a.K[live_col]
    if not np.all(np.isfinite(a.K)):
        raise ValueError("K must be finite")
    with np.errstate(invalid="ignore", divide="ignore"):
        nobs = np.rint(Yl * Kl)
        ints = bool(np.all((Kl >= 1) & (Kl < 2 ** 31) & (Kl == np.floor(Kl)) & (nobs >= 0) & (nobs < 2 ** 31) & (nobs / Kl == Yl)))
    if ints:
        colY, colK, mode = nobs.astype(np.int32), Kl.astype(np.int32), 1
    else:
        colY, colK, mode = Yl.astype(np.float64), Kl.astype(np.float64), 0
    return Wire(a.grp_row_ptr.astype(np.int64), grp_col_ptr, grp_ent_ptr, row_end, rs, a.aval[keep].astype(np.float64),
                np.frombuffer(bytes(blob), np.uint8).copy(), eq_bits, mode, colY, colK, multi_bits,
                a.group_nonzeros().astype(np.int64))
