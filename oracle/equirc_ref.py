"""Loop-level restatement of the steps between the distance table and
``bambu.quantDT`` -- TEST INFRASTRUCTURE (used to replay the reference's bundled
chr9 end-to-end fixture without R).

Follows ``genEquiRCs`` and helpers, ``R/bambu-quantify_utilityFunctions.R:46-192``,
and the tail of ``bambu.quantify``, ``R/bambu-quantify.R:19-38`` (bambu 3.3.5).
Input is the dict stored in ``tests/golden/chr9_e2e.json`` (made by
``tests/golden/make_golden.py`` from ``inst/extdata/seOutput_distTable_*.rds`` and
``seReadClassUnstranded_*.rds``).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def _split_by_end(values, ends):
    out, s = [], 0
    for e in ends:
        out.append(list(values[s:e]))
        s = e
    return out


def read_class_widths(rc):
    """``rcWidth`` table of ``genEquiRCsBasedOnObservedReads`` (:76-80)."""
    widths = _split_by_end(rc["exon_width"], rc["exon_end"])
    ranks = _split_by_end(rc["exon_rank"], rc["exon_end"])
    first, total = {}, {}
    for name, w, r in zip(rc["names"], widths, ranks):
        first[name] = [wi for wi, ri in zip(w, r) if ri == 1][0]
        total[name] = sum(w)
    return first, total


def gen_equi_rcs_observed(fx):
    """``genEquiRCsBasedOnObservedReads`` (:74-101): rows of the distance table that are
    compatible and belong to the read class's own gene, plus the signed-txid class key."""
    dt, rc = fx["distTable"], fx["readClasses"]
    first, total = read_class_widths(rc)
    rows = []
    for k in range(len(dt["readClassId"])):
        if "unidentified" in str(dt["annotationTxId"][k]):
            continue
        rows.append({"readClassId": dt["readClassId"][k], "readCount": dt["readCount"][k], "GENEID": dt["GENEID"][k],
                     "equal": bool(dt["equal"][k]), "txid": int(dt["txid"][k])})
    present = {r["readClassId"] for r in rows}
    gene_of_rc = dict(zip(rc["names"], rc["GENEID"]))
    out = []
    for name in rc["names"]:                       # X[Y] join: order of the read-class table
        if name not in present:
            continue
        for r in rows:
            if r["readClassId"] == name and r["GENEID"] == gene_of_rc[name]:
                q = dict(r)
                q["firstExonWidth"], q["totalWidth"] = first[name], total[name]
                out.append(q)
    # createList (:106-112): per read class the sorted signed txids (negative = equal)
    members = {}
    for r in out:
        members.setdefault(r["readClassId"], []).append(r["txid"] * (-1 if r["equal"] else 1))
    for r in out:
        r["eqClassById"] = tuple(sorted(members[r["readClassId"]]))
    return out


def uni_count_per_equirc(rows):
    """``getUniCountPerEquiRC`` (:118-132), including its distinct()-before-sum quirk."""
    any_equal = {}
    for r in rows:
        any_equal[r["eqClassById"]] = any_equal.get(r["eqClassById"], False) or r["equal"]
    seen = OrderedDict()
    for r in rows:
        key = (r["eqClassById"], r["firstExonWidth"], r["totalWidth"], r["readCount"], r["GENEID"])
        seen[key] = None
    nobs, width = {}, {}
    for (cls, few, tw, cnt, gene) in seen:
        nobs[cls] = nobs.get(cls, 0) + cnt
        w = tw if any_equal[cls] else few
        width[cls] = max(width.get(cls, w), w)
    out = OrderedDict()
    for (cls, _few, _tw, _cnt, gene) in seen:
        out[(cls, gene)] = {"eqClassById": cls, "GENEID": gene, "nobs": nobs[cls], "rcWidth": width[cls]}
    return list(out.values())


def process_min_equirc(ann):
    """``processMinEquiRC`` (:159-181): per annotated transcript two minimal classes --
    all members partial (minRC = 1), and the owner equal with the others partial."""
    lists = _split_by_end(ann["eqClassById_values"], ann["eqClassById_end"])
    part, signed = [], OrderedDict()
    for cls, gene, tx in zip(lists, ann["GENEID"], ann["txid"]):
        for member in cls:
            part.append({"eqClassById": tuple(cls), "GENEID": gene, "txid": int(member), "minRC": 1, "equal": False})
        key = tuple(sorted(int(m) * (-1 if int(m) == int(tx) else 1) for m in cls))
        for member in cls:
            row = (key, gene, int(member), int(member) == int(tx))
            signed[row] = None
    return part + [{"eqClassById": k, "GENEID": g, "txid": t, "minRC": None, "equal": e} for (k, g, t, e) in signed]


def gen_equi_rcs(fx):
    """``genEquiRCs`` (:46-64): the long table ``bambu.quantDT`` takes."""
    ann = fx["annotation"]
    counts = uni_count_per_equirc(gen_equi_rcs_observed(fx))
    # createEqClassToTxMapping (:185-192)
    x = []
    for c in counts:
        for t in c["eqClassById"]:
            x.append({"eqClassById": c["eqClassById"], "GENEID": c["GENEID"], "nobs": c["nobs"], "rcWidth": c["rcWidth"],
                      "txid": abs(t), "equal": t < 0, "minRC": None})
    y = process_min_equirc(ann)
    # full_join by (eqClassById, GENEID, txid, equal); NA -> 0   (addEmptyRC :140-152)
    key = lambda r: (r["eqClassById"], r["GENEID"], r["txid"], r["equal"])  # noqa: E731
    ymap = {}
    for r in y:
        ymap.setdefault(key(r), []).append(r)
    joined, used = [], set()
    for r in x:
        ms = ymap.get(key(r))
        if ms:
            used.add(key(r))
            for m in ms:
                joined.append(dict(r, minRC=m["minRC"]))
        else:
            joined.append(dict(r))
    for r in y:
        if key(r) not in used:
            joined.append({"eqClassById": r["eqClassById"], "GENEID": r["GENEID"], "nobs": None, "rcWidth": None,
                           "txid": r["txid"], "equal": r["equal"], "minRC": r["minRC"]})
    for r in joined:
        for c in ("nobs", "rcWidth", "minRC"):
            if r[c] is None:
                r[c] = 0
    mx = {}
    for r in joined:
        m = mx.setdefault(r["eqClassById"], [0, 0, 0])
        m[0], m[1], m[2] = max(m[0], r["nobs"]), max(m[1], r["rcWidth"]), max(m[2], r["minRC"])
    final = OrderedDict()
    for r in joined:
        m = mx[r["eqClassById"]]
        final[(r["eqClassById"], r["GENEID"], r["txid"], r["equal"])] = (m[0], m[1], m[2])
    # eqClassId = group id of eqClassById (list column: groups numbered by first appearance)
    ids = OrderedDict()
    for (cls, _g, _t, _e) in final:
        ids.setdefault(cls, len(ids) + 1)
    widths = _split_by_end(ann["exon_width"], ann["exon_end"])
    txlen = {int(t): float(sum(w)) for t, w in zip(ann["txid"], widths)}
    tab = {k: [] for k in ("GENEID", "txid", "eqClassId", "nobs", "equal", "rcWidth", "txlen", "minRC")}
    for (cls, gene, tx, eq), (nobs, rcw, minrc) in final.items():
        tab["GENEID"].append(gene); tab["txid"].append(tx); tab["eqClassId"].append(ids[cls])
        tab["nobs"].append(float(nobs)); tab["equal"].append(eq); tab["rcWidth"].append(float(rcw))
        tab["txlen"].append(txlen[tx]); tab["minRC"].append(minrc)
    return tab


def r_round(x, digits):
    """R's round(x, digits) for the magnitudes met here (round half even on the decimal
    representation, as R >= 4.0 does)."""
    return np.array([float(np.format_float_positional(v, precision=digits, unique=False, fractional=True))
                     if np.isfinite(v) else v for v in np.asarray(x, dtype=np.float64)])


def assemble_assays(fx, quant, cpm_fn, sig_digit=5):
    """Tail of ``bambu.quantify`` (R/bambu-quantify.R:23-38): CPM, annotation order, rounding."""
    ann = fx["annotation"]
    inc = np.array(fx["incompatibleCounts"]["counts"], dtype=np.float64)
    inc = np.where(np.isnan(inc), 0.0, inc)
    cpm = cpm_fn(quant["counts"], inc)
    pos = {int(t): k for k, t in enumerate(quant["txid"])}
    order = [pos[int(t)] for t in ann["txid"]]
    return {"counts": r_round(np.asarray(quant["counts"])[order], sig_digit),
            "CPM": r_round(cpm[order], sig_digit),
            "fullLengthCounts": r_round(np.asarray(quant["fullLengthCounts"])[order], sig_digit),
            "uniqueCounts": np.asarray(quant["uniqueCounts"])[order]}
