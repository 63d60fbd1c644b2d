"""Equivalence read classes from the distance table -- ``genEquiRCs`` and helpers
(R/bambu-quantify_utilityFunctions.R:46-192): the producer of the long table
``bambu.quantDT`` takes.  Host code, like the R it mirrors.

Class keys (R: list column ``eqClassById`` of sorted signed txids, negative = full-length match) are interned
into integers once; every later step -- the per-class counts and widths, the join with the annotation's minimal
classes, ``cur_group_id()`` -- works on integer columns.  The annotation side (``processMinEquiRC``: two minimal
classes per transcript) is the same for every sample of a cohort: build :class:`AnnotationClasses` once and pass
it to every :func:`genEquiRCs` call.

Tables are dicts of equal-length lists / arrays, column names as in the R code."""
from __future__ import annotations

import numpy as np


def _split(values, ends):
    out, s = [], 0
    for e in ends:
        out.append(values[s:e])
        s = e
    return out


class _Interner:
    """tuple of signed txids -> dense id; remembers the tuples."""
    def __init__(self):
        self.ids, self.keys = {}, []

    def __call__(self, key):
        k = self.ids.get(key)
        if k is None:
            k = self.ids[key] = len(self.keys)
            self.keys.append(key)
        return k


class AnnotationClasses:
    """``processMinEquiRC`` (:159-181) + transcript lengths (:58-59), kept for a whole cohort.

    For every annotated transcript with minimal class ``{t1..tk}``: the all-partial class (``minRC = 1``) and the
    class in which the transcript itself is matched full length (signed key, ``minRC`` absent)."""

    def __init__(self, annotation):
        self.intern = _Interner()
        txid = [int(t) for t in annotation["txid"]]
        genes = list(annotation["GENEID"])
        members = _split([int(v) for v in annotation["eqClassById_values"]], annotation["eqClassById_end"])
        widths = _split(annotation["exon_width"], annotation["exon_end"])
        self.txlen = {t: float(sum(w)) for t, w in zip(txid, widths)}
        rows = []                                   # (class id, gene, txid, equal, minRC or None), bind_rows order
        for cls, g, _t in zip(members, genes, txid):
            cid = self.intern(tuple(cls))           # annotation keys are stored sorted (unsigned = all partial)
            rows.extend((cid, g, m, False, 1) for m in cls)
        seen = set()
        for cls, g, t in zip(members, genes, txid):
            cid = self.intern(tuple(sorted(-m if m == t else m for m in cls)))
            for m in cls:
                r = (cid, g, m, m == t)
                if r not in seen:                   # distinct()
                    seen.add(r)
                    rows.append(r + (None,))
        self.rows = rows
        self.n_annotation_keys = len(self.intern.keys)


def genEquiRCsBasedOnObservedReads(distTable, readClasses):
    """:74-101 -- compatible rows of the distance table that lie in the read class's own gene, with the read
    class's first-exon and total width and its class key (sorted signed txids of ALL its rows)."""
    names = list(readClasses["names"])
    w = _split(list(readClasses["exon_width"]), readClasses["exon_end"])
    r = _split(list(readClasses["exon_rank"]), readClasses["exon_end"])
    first = {n: next(wi for wi, ri in zip(ws, rs) if ri == 1) for n, ws, rs in zip(names, w, r)}
    total = {n: sum(ws) for n, ws in zip(names, w)}
    gene_of = dict(zip(names, readClasses["GENEID"]))
    by_rc = {}
    for k, rc in enumerate(distTable["readClassId"]):
        if "unidentified" in str(distTable["annotationTxId"][k]):
            continue
        by_rc.setdefault(rc, []).append(k)
    out = {c: [] for c in ("readClassId", "readCount", "GENEID", "equal", "txid", "firstExonWidth", "totalWidth", "key")}
    for rc in names:                                # X[Y] join: order of rowData(readClass)
        ks = [k for k in by_rc.get(rc, ()) if distTable["GENEID"][k] == gene_of[rc]]
        if not ks:
            continue
        key = tuple(sorted(int(distTable["txid"][k]) * (-1 if distTable["equal"][k] else 1) for k in ks))
        for k in ks:
            out["readClassId"].append(rc); out["readCount"].append(distTable["readCount"][k])
            out["GENEID"].append(distTable["GENEID"][k]); out["equal"].append(bool(distTable["equal"][k]))
            out["txid"].append(int(distTable["txid"][k])); out["firstExonWidth"].append(first[rc])
            out["totalWidth"].append(total[rc]); out["key"].append(key)
    return out


def getUniCountPerEquiRC(obs, intern):
    """:118-132 -- per class: ``nobs = sum(readCount)`` and ``rcWidth`` over the DISTINCT
    (class, widths, readCount, gene) rows (two read classes with equal widths and counts collapse: the
    reference's distinct()-before-sum), ``rcWidth = max(totalWidth)`` if any row is a full-length match, else
    ``max(firstExonWidth)``.  Returns rows (class id, gene, nobs, rcWidth), distinct, first-appearance order."""
    cid = [intern(k) for k in obs["key"]]
    any_equal = {}
    for c, e in zip(cid, obs["equal"]):
        any_equal[c] = any_equal.get(c, False) or e
    distinct = dict.fromkeys(zip(cid, obs["firstExonWidth"], obs["totalWidth"], obs["readCount"], obs["GENEID"]))
    nobs, width = {}, {}
    for c, few, tw, cnt, _g in distinct:
        nobs[c] = nobs.get(c, 0) + cnt
        wv = tw if any_equal[c] else few
        width[c] = max(width.get(c, wv), wv)
    return [(c, g, nobs[c], width[c]) for c, g in dict.fromkeys((c, g) for c, _f, _t, _n, g in distinct)]


def genEquiRCs(distTable, readClasses, annotation):
    """:46-64 -- the long table of ``bambu.quantDT``: columns GENEID, txid, eqClassId, nobs, equal, rcWidth,
    txlen, minRC.  ``annotation``: an :class:`AnnotationClasses` (reuse it across samples) or the annotation
    columns (TXNAME, GENEID, txid, eqClassById_values/_end, exon_width/_end)."""
    ann = annotation if isinstance(annotation, AnnotationClasses) else AnnotationClasses(annotation)
    intern = _Interner()                            # per sample, seeded with the annotation's keys
    intern.ids, intern.keys = dict(ann.intern.ids), list(ann.intern.keys)
    counts = getUniCountPerEquiRC(genEquiRCsBasedOnObservedReads(distTable, readClasses), intern)
    # createEqClassToTxMapping (:185-192) then full_join with the minimal classes (addEmptyRC :140-152)
    min_rows = {}
    for c, g, t, e, m in ann.rows:
        min_rows.setdefault((c, g, t, e), []).append(m)
    joined, used = [], set()                        # rows: [class, gene, txid, equal, nobs, rcWidth, minRC]
    for c, g, n, wv in counts:
        for s in intern.keys[c]:
            k = (c, g, abs(s), s < 0)
            ms = min_rows.get(k)
            if ms is None:
                joined.append([c, g, abs(s), s < 0, n, wv, 0])
            else:
                used.add(k)
                joined.extend([c, g, abs(s), s < 0, n, wv, m or 0] for m in ms)
    for c, g, t, e, m in ann.rows:                  # unmatched rows of the right table follow, NA -> 0
        if (c, g, t, e) not in used:
            joined.append([c, g, t, e, 0, 0, m or 0])
    top = {}
    for c, _g, _t, _e, n, wv, m in joined:          # group_by(eqClassById): max of nobs, rcWidth, minRC
        a = top.setdefault(c, [0, 0, 0])
        a[0], a[1], a[2] = max(a[0], n), max(a[1], wv), max(a[2], m)
    final = dict.fromkeys((c, g, t, e) for c, g, t, e, _n, _w, _m in joined)          # distinct()
    # cur_group_id(): dplyr numbers the groups of a list column by first appearance
    eq_id = {}
    for c, _g, _t, _e in final:
        eq_id.setdefault(c, len(eq_id) + 1)
    n = len(final)
    tab = {"GENEID": [None] * n, "txid": np.zeros(n, np.int64), "eqClassId": np.zeros(n, np.int64), "nobs": np.zeros(n),
           "equal": np.zeros(n, bool), "rcWidth": np.zeros(n), "txlen": np.zeros(n), "minRC": np.zeros(n, np.int64)}
    for k, (c, g, t, e) in enumerate(final):
        tab["GENEID"][k] = g; tab["txid"][k] = t; tab["eqClassId"][k] = eq_id[c]; tab["equal"][k] = e
        tab["nobs"][k], tab["rcWidth"][k], tab["minRC"][k] = top[c]
        tab["txlen"][k] = ann.txlen[t]
    return tab


def genEquiRCs_native(distTable, readClasses, annotation):
    """:func:`genEquiRCs` done by the native host code of ``libbambu_em.so`` (``bambu_equirc_create``, csrc/equirc.cpp):
    hash joins on integer columns instead of Python dictionaries of tuples.  Same table, same row order, same class
    numbering (tests/test_equirc_random.py, tests/test_chr9_e2e.py)."""
    import ctypes
    from . import _lib
    L = _lib.lib()
    ann_txid = np.ascontiguousarray(annotation["txid"], dtype=np.int64)
    genes = {}
    code = lambda g: genes.setdefault(g, len(genes))        # noqa: E731
    ann_gene = np.array([code(g) for g in annotation["GENEID"]], dtype=np.int64)
    ptr_ = np.concatenate(([0], np.asarray(annotation["eqClassById_end"], dtype=np.int64))).astype(np.int64)
    val_ = np.ascontiguousarray(annotation["eqClassById_values"], dtype=np.int64)
    ew = np.asarray(annotation["exon_width"], dtype=np.float64)
    ee = np.concatenate(([0], np.asarray(annotation["exon_end"], dtype=np.int64)))
    txlen = np.array([ew[ee[i]:ee[i + 1]].sum() for i in range(len(ann_txid))], dtype=np.float64)
    names = list(readClasses["names"])
    rc_row = {n: k for k, n in enumerate(names)}
    rc_gene = np.array([code(g) for g in readClasses["GENEID"]], dtype=np.int64)
    rw = np.asarray(readClasses["exon_width"], dtype=np.float64)
    rr = np.asarray(readClasses["exon_rank"], dtype=np.float64)
    re = np.concatenate(([0], np.asarray(readClasses["exon_end"], dtype=np.int64)))
    first = np.zeros(len(names))
    total = np.zeros(len(names))
    for k in range(len(names)):
        w, r = rw[re[k]:re[k + 1]], rr[re[k]:re[k + 1]]
        first[k] = w[np.nonzero(r == 1)[0][0]]
        total[k] = w.sum()
    n_dt = len(distTable["txid"])
    dt_rc = np.array([rc_row[r] for r in distTable["readClassId"]], dtype=np.int64)
    dt_tx = np.ascontiguousarray(distTable["txid"], dtype=np.int64)
    dt_cnt = np.ascontiguousarray(distTable["readCount"], dtype=np.float64)
    dt_eq = np.ascontiguousarray(np.asarray(distTable["equal"], dtype=bool), dtype=np.uint8)
    dt_gene = np.array([code(g) for g in distTable["GENEID"]], dtype=np.int64)
    dt_un = np.array(["unidentified" in str(a) for a in distTable["annotationTxId"]], dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    h = ctypes.c_void_p()
    rc = L.bambu_equirc_create(ctypes.byref(h), len(ann_txid), p(ann_txid), p(ann_gene), p(ptr_), p(val_), p(txlen), len(names),
                               p(rc_gene), p(first), p(total), n_dt, p(dt_rc), p(dt_tx), p(dt_cnt), p(dt_eq), p(dt_gene), p(dt_un))
    if rc:
        raise ValueError(L.bambu_equirc_last_error().decode())
    try:
        n = ctypes.c_int64()
        L.bambu_equirc_size(h, ctypes.byref(n))
        n = int(n.value)
        ptrs = [ctypes.c_void_p() for _ in range(8)]
        L.bambu_equirc_table(h, *[ctypes.byref(x) for x in ptrs])

        def arr(q, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.frombuffer((ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(q.value), dtype=dt).copy()
        gene_names = list(genes)
        g = arr(ptrs[0], np.int64)
        tab = {"GENEID": [gene_names[k] for k in g], "txid": arr(ptrs[1], np.int64), "eqClassId": arr(ptrs[2], np.int64),
               "nobs": arr(ptrs[3], np.float64), "equal": arr(ptrs[4], np.uint8).astype(bool), "rcWidth": arr(ptrs[5], np.float64),
               "txlen": arr(ptrs[6], np.float64), "minRC": arr(ptrs[7], np.int64)}
    finally:
        L.bambu_equirc_free(h)
    return tab
