"""Segmented-CSR arena: every gene group of one or many samples in one set of
flat arrays.  This is the data format of the C-ABI (``include/bambu_em.h``).

One *gene group* is what the reference hands to one ``emWithL1`` call
(``R/bambu-quantify_utilityFunctions.R:401-431``): the M x N matrix whose rows
are the group's transcripts (``txid`` ascending) and whose columns are its
equivalence read classes (``(gene_sid, eqClassId)`` ascending).  Instead of the
three dense matrices ``A``, ``A_full``, ``A_unique`` that ``getAMat`` builds
(``:427-431``) the arena stores the structural entries of ``A`` once, row by
row, plus one flag per entry (``equal`` => ``A_full = aval * equal``) and one
flag per column (``multi_align`` => ``A_unique = aval * !multi_align``), which
is exactly how ``filterTxRc`` defines them (``:336-337``).

Layout (g = group, rows/columns/entries of all groups are concatenated):

    grp_row_ptr[n_groups+1]  int64   rows of group g      = [grp_row_ptr[g], grp_row_ptr[g+1])
    grp_col_ptr[n_groups+1]  int64   columns of group g   = [grp_col_ptr[g], grp_col_ptr[g+1])
    row_nnz_ptr[rows+1]      int64   entries of row r     = [row_nnz_ptr[r], row_nnz_ptr[r+1])
    col_idx[nnz]             int32   column of the entry, local to its group, strictly
                                     increasing inside a row
    aval[nnz]                float64 A value (explicit zeros allowed)
    nz_flags[nnz]            uint8   bit 0: equal (full-length) entry
    Y[cols]                  float64 n.obs of the column
    K[cols]                  float64 K of the column's gene
    col_flags[cols]          uint8   bit 0: multi_align column
    txid[rows]               int32   transcript id of the row (carried, not used by the EM)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BAMBU_NZ_EQUAL = 1
BAMBU_COL_MULTI = 1


@dataclass
class Arena:
    grp_row_ptr: np.ndarray
    grp_col_ptr: np.ndarray
    row_nnz_ptr: np.ndarray
    col_idx: np.ndarray
    aval: np.ndarray
    nz_flags: np.ndarray
    Y: np.ndarray
    K: np.ndarray
    col_flags: np.ndarray
    txid: np.ndarray
    # optional bookkeeping carried by the packers (not part of the C-ABI)
    sample_grp_ptr: np.ndarray | None = None   # groups of sample s = [ptr[s], ptr[s+1])
    meta: dict = field(default_factory=dict)

    # ------------------------------------------------------------------ sizes
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
        return int(self.row_nnz_ptr[-1])

    def nbytes(self) -> int:
        return sum(a.nbytes for a in (self.grp_row_ptr, self.grp_col_ptr, self.row_nnz_ptr, self.col_idx,
                                      self.aval, self.nz_flags, self.Y, self.K, self.col_flags))

    # ------------------------------------------------------------- per group
    def group_shapes(self):
        """(M, N, structural entries) per group."""
        M = np.diff(self.grp_row_ptr)
        N = np.diff(self.grp_col_ptr)
        e = self.row_nnz_ptr[self.grp_row_ptr[1:]] - self.row_nnz_ptr[self.grp_row_ptr[:-1]]
        return M, N, e

    def group_nonzeros(self) -> np.ndarray:
        """Entries with ``aval != 0`` per group -- the ``nnz_g`` of the
        nnz-iterations metric (explicit zeros are no-ops)."""
        nzc = np.concatenate(([0], np.cumsum(self.aval != 0, dtype=np.int64)))
        lo = self.row_nnz_ptr[self.grp_row_ptr[:-1]]
        hi = self.row_nnz_ptr[self.grp_row_ptr[1:]]
        return nzc[hi] - nzc[lo]

    def dense_group(self, g: int):
        """Dense (A, A_full, A_unique, Y, K) of group g, as ``getAMat`` would
        build them -- for tests on small groups only."""
        r0, r1 = int(self.grp_row_ptr[g]), int(self.grp_row_ptr[g + 1])
        c0, c1 = int(self.grp_col_ptr[g]), int(self.grp_col_ptr[g + 1])
        M, N = r1 - r0, c1 - c0
        A = np.zeros((M, N))
        Af = np.zeros((M, N))
        Au = np.zeros((M, N))
        multi = (self.col_flags[c0:c1] & BAMBU_COL_MULTI) != 0
        for i in range(M):
            k0, k1 = int(self.row_nnz_ptr[r0 + i]), int(self.row_nnz_ptr[r0 + i + 1])
            j = self.col_idx[k0:k1]
            a = self.aval[k0:k1]
            A[i, j] = a
            Af[i, j] = a * ((self.nz_flags[k0:k1] & BAMBU_NZ_EQUAL) != 0)
            Au[i, j] = a * (~multi[j])
        return A, Af, Au, self.Y[c0:c1].copy(), self.K[c0:c1].copy()

    def validate(self) -> None:
        """Structural checks the C-ABI also performs (``BAMBU_EM_EINVAL``)."""
        G = self.n_groups
        for name, p in (("grp_row_ptr", self.grp_row_ptr), ("grp_col_ptr", self.grp_col_ptr),
                        ("row_nnz_ptr", self.row_nnz_ptr)):
            if p[0] != 0 or np.any(np.diff(p) < 0):
                raise ValueError(f"{name} must start at 0 and be non-decreasing")
        if len(self.grp_col_ptr) != G + 1 or len(self.row_nnz_ptr) != self.n_rows + 1:
            raise ValueError("pointer arrays have inconsistent lengths")
        if not (len(self.col_idx) == len(self.aval) == len(self.nz_flags) == self.nnz):
            raise ValueError("entry arrays must have nnz elements")
        if not (len(self.Y) == len(self.K) == len(self.col_flags) == self.n_cols):
            raise ValueError("column arrays must have n_cols elements")
        if self.nnz:
            rows_of = np.repeat(np.arange(self.n_rows), np.diff(self.row_nnz_ptr))
            grp_of_row = np.repeat(np.arange(G), np.diff(self.grp_row_ptr))
            ncol = np.diff(self.grp_col_ptr)[grp_of_row[rows_of]]
            if np.any(self.col_idx < 0) or np.any(self.col_idx >= ncol):
                raise ValueError("col_idx out of range for its group")
            same_row = rows_of[1:] == rows_of[:-1]
            if np.any(same_row & (np.diff(self.col_idx.astype(np.int64)) <= 0)):
                raise ValueError("col_idx must be strictly increasing inside a row")

    # ------------------------------------------------------------ assembling
    @staticmethod
    def concatenate(arenas: "list[Arena]") -> "Arena":
        """One arena holding the groups of several arenas (e.g. one per
        sample) back to back; ``sample_grp_ptr`` records the boundaries."""
        def cat_ptr(ps):
            out, off = [np.zeros(1, np.int64)], 0
            for p in ps:
                out.append(p[1:] + off)
                off += int(p[-1])
            return np.concatenate(out)
        a = arenas
        sgp = np.concatenate(([0], np.cumsum([x.n_groups for x in a]))).astype(np.int64)
        return Arena(cat_ptr([x.grp_row_ptr for x in a]), cat_ptr([x.grp_col_ptr for x in a]),
                     cat_ptr([x.row_nnz_ptr for x in a]),
                     np.concatenate([x.col_idx for x in a]), np.concatenate([x.aval for x in a]),
                     np.concatenate([x.nz_flags for x in a]), np.concatenate([x.Y for x in a]),
                     np.concatenate([x.K for x in a]), np.concatenate([x.col_flags for x in a]),
                     np.concatenate([x.txid for x in a]), sample_grp_ptr=sgp)

    def select_groups(self, groups) -> "Arena":
        """Sub-arena with the listed groups, in the listed order (used to shard
        one sample's groups across ranks)."""
        groups = np.asarray(groups, dtype=np.int64)
        M, N, _ = self.group_shapes()

        def ranges(starts, lens):
            tot = int(lens.sum())
            if tot == 0:
                return np.zeros(0, np.int64)
            off = np.repeat(starts - np.concatenate(([0], np.cumsum(lens)[:-1])), lens)
            return off + np.arange(tot, dtype=np.int64)
        rows = ranges(self.grp_row_ptr[groups], M[groups])
        cols = ranges(self.grp_col_ptr[groups], N[groups])
        rl = np.diff(self.row_nnz_ptr)[rows]
        ent = ranges(self.row_nnz_ptr[rows], rl)
        z = np.zeros(1, np.int64)
        return Arena(np.concatenate((z, np.cumsum(M[groups]))), np.concatenate((z, np.cumsum(N[groups]))),
                     np.concatenate((z, np.cumsum(rl))), self.col_idx[ent], self.aval[ent], self.nz_flags[ent],
                     self.Y[cols], self.K[cols], self.col_flags[cols], self.txid[rows])

    # ------------------------------------------------------------ on disk
    FORMAT = "bambu-arena-1"

    def save(self, path) -> None:
        """Write the arena as an uncompressed ``.npz`` (one member per ABI array, plus ``sample_grp_ptr`` and a
        format tag): a cohort packed once can be re-quantified later without R."""
        members = {n: np.ascontiguousarray(getattr(self, n)) for n in
                   ("grp_row_ptr", "grp_col_ptr", "row_nnz_ptr", "col_idx", "aval", "nz_flags", "Y", "K", "col_flags", "txid")}
        if self.sample_grp_ptr is not None:
            members["sample_grp_ptr"] = np.asarray(self.sample_grp_ptr, dtype=np.int64)
        members["format"] = np.array(self.FORMAT)
        with open(path, "wb") as f:
            np.savez(f, **members)

    @staticmethod
    def load(path, validate: bool = True) -> "Arena":
        with np.load(path, allow_pickle=False) as z:
            if "format" not in z.files or str(z["format"]) != Arena.FORMAT:
                raise ValueError(f"{path}: not a {Arena.FORMAT} file")
            a = Arena(z["grp_row_ptr"], z["grp_col_ptr"], z["row_nnz_ptr"], z["col_idx"], z["aval"], z["nz_flags"], z["Y"],
                      z["K"], z["col_flags"], z["txid"],
                      sample_grp_ptr=z["sample_grp_ptr"] if "sample_grp_ptr" in z.files else None)
        if validate:
            a.validate()
        return a

    @staticmethod
    def from_dense_groups(groups) -> "Arena":
        """Build an arena from dense per-group inputs, each a tuple
        ``(A, equal_mask, multi_align[N], Y[N], K[N], txid[M])``; ``A`` entries
        that are structurally absent are given as NaN-free zeros and dropped
        unless ``equal_mask`` is non-zero there (tests only)."""
        grp_row, grp_col, rnp = [0], [0], [0]
        ci, av, nf, Ys, Ks, cf, tx = [], [], [], [], [], [], []
        for (A, eq, multi, Y, K, txid) in groups:
            A = np.asarray(A, dtype=np.float64)
            M, N = A.shape
            eq = np.zeros((M, N), bool) if eq is None else np.asarray(eq, bool)
            for i in range(M):
                j = np.nonzero((A[i] != 0) | eq[i])[0]
                ci.append(j.astype(np.int32))
                av.append(A[i, j])
                nf.append(eq[i, j].astype(np.uint8))
                rnp.append(rnp[-1] + len(j))
            grp_row.append(grp_row[-1] + M)
            grp_col.append(grp_col[-1] + N)
            Ys.append(np.asarray(Y, np.float64))
            Ks.append(np.asarray(K, np.float64))
            cf.append(np.asarray(multi, bool).astype(np.uint8))
            tx.append(np.asarray(txid if txid is not None else np.arange(M), np.int32))
        cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))  # noqa: E731
        return Arena(np.asarray(grp_row, np.int64), np.asarray(grp_col, np.int64), np.asarray(rnp, np.int64),
                     cat(ci, np.int32), cat(av, np.float64), cat(nf, np.uint8), cat(Ys, np.float64),
                     cat(Ks, np.float64), cat(cf, np.uint8), cat(tx, np.int32))
