"""Loop-level restatement of ``bambu.quantDT`` -- TEST INFRASTRUCTURE.

Follows, step by step and with plain Python loops and dictionaries,
``R/bambu-quantify.R:53-74`` and
``R/bambu-quantify_utilityFunctions.R:196-455`` of the reference (bambu 3.3.5).
It is deliberately independent of ``bambu_b200/quantify.py`` (the vectorised
product packer) so that the two can be compared in tests.  R's ``sum()`` /
``mean()`` accumulate in ``long double``; ``numpy.longdouble`` (80-bit on
x86-64) restates that.

Input "long table": dict of equal-length lists with the columns
``GENEID, txid, eqClassId, nobs, equal, rcWidth, txlen``
(``R/bambu-quantify_utilityFunctions.R:199-203``).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

from . import emWithL1

LD = np.longdouble


def _r_sum(xs):
    s = LD(0)
    for x in xs:
        s += LD(x)
    return float(s)


def _r_median(xs):
    xs = sorted(x for x in xs if x is not None and not math.isnan(x))
    n = len(xs)
    if n == 0:
        return float("nan")
    if n % 2:
        return xs[n // 2]
    a, b = xs[n // 2 - 1], xs[n // 2]
    # R: mean(c(a, b)) -- long double sum / n, plus one refinement pass
    s = (LD(a) + LD(b)) / LD(2)
    t = ((LD(a) - s) + (LD(b) - s)) / LD(2)
    return float(s + t)


def simplify_names(tab):
    """``simplifyNames`` (:236-241): gene_sid = rank of GENEID by first appearance (1-based)."""
    seen = OrderedDict()
    sid = []
    for g in tab["GENEID"]:
        if g not in seen:
            seen[g] = len(seen) + 1
        sid.append(seen[g])
    return sid


def degradation_rate(rows):
    """``calculateDegradationRate`` (:247-269).  ``rows`` = list of row dicts."""
    rc = OrderedDict()
    rc_par = OrderedDict()
    for r in rows:
        rc[(r["gene_sid"], r["eqClassId"], r["nobs"])] = None
        if not r["equal"]:
            rc_par[(r["gene_sid"], r["eqClassId"], r["nobs"])] = None
    gene_count = OrderedDict()
    for (g, _e, n) in rc:
        gene_count.setdefault(g, []).append(n)
    gene_par = OrderedDict()
    for (g, _e, n) in rc_par:
        gene_par.setdefault(g, []).append(n)
    gene_len = {}
    for r in rows:
        gene_len[r["gene_sid"]] = max(gene_len.get(r["gene_sid"], -math.inf), r["txlen"])
    table = []
    for g, ns in gene_count.items():
        nobs = _r_sum(ns)
        dobs = _r_sum(gene_par[g]) if g in gene_par else None
        rate = None if dobs is None else (dobs / nobs if nobs != 0 else float("nan"))
        table.append((g, nobs, dobs, rate, gene_len[g]))
    keep = [t for t in table if t[2] is not None and t[1] >= 30 and (t[1] - t[2]) >= 5]
    if keep:
        table = keep
    vals = [None if t[3] is None else t[3] * 1000 / t[4] for t in table]
    return _r_median(vals), len(table)


def add_aval(rows, d_rate, d_mode):
    """``modifyAvaluewithDegradation_rate`` (:276-296)."""
    txs = {}
    for r in rows:
        txs.setdefault((r["eqClassId"], r["gene_sid"]), set()).add(r["txid"])
    for r in rows:
        r["multi_align"] = len(txs[(r["eqClassId"], r["gene_sid"])]) > 1
    if not d_mode:
        for r in rows:
            r["aval"] = 1.0
        return
    for r in rows:
        r["aval"] = None
    by_gt = OrderedDict()
    for r in rows:
        if r["multi_align"]:
            by_gt.setdefault((r["gene_sid"], r["txid"]), []).append(r)
    for grp in by_gt.values():
        partial = _r_sum([q["rcWidth"] * d_rate / 1000 for q in grp if not q["equal"]])
        for q in grp:
            q["aval"] = (1 - partial) if q["equal"] else q["rcWidth"] * d_rate / 1000
    if d_rate == 0:
        status = {}
        for r in rows:
            k = (r["eqClassId"], r["gene_sid"])
            status[k] = status.get(k, True) and ((not r["equal"]) and r["multi_align"])
        for r in rows:
            if status[(r["eqClassId"], r["gene_sid"])]:
                r["aval"] = 0.01
    for r in rows:
        if r["aval"] is not None:
            r["aval"] = max(min(r["aval"], 1.0), 0.0)
    for r in rows:
        if r["multi_align"] and r["equal"]:
            r["aval"] = min(1.0, max(r["aval"], r["rcWidth"] * d_rate / 1000))
    for r in rows:
        if not r["multi_align"]:
            r["aval"] = 1.0


def quantDT_ref(tab, degradationBias=True, maxiter=10000, conv=1e-2, minvalue=1e-8, details=None):
    """``bambu.quantDT`` (R/bambu-quantify.R:53-74).

    Returns an ``OrderedDict txid -> (counts, fullLengthCounts, uniqueCounts)`` in
    the row order of the reference's result table.  If ``details`` is a dict it
    receives ``d_rate`` and the per-group dense inputs/outputs (``groups``)."""
    n = len(tab["txid"])
    sid = simplify_names(tab)
    rows = []
    for k in range(n):
        rows.append({"gene_sid": sid[k], "txid": int(tab["txid"][k]), "eqClassId": int(tab["eqClassId"][k]),
                     "nobs": float(tab["nobs"][k]), "equal": bool(tab["equal"][k]),
                     "rcWidth": float(tab["rcWidth"][k]) if "rcWidth" in tab else 0.0,
                     "txlen": float(tab["txlen"][k]) if "txlen" in tab else 1.0})
    d_rate = float("nan")
    if degradationBias:
        d_rate, _ = degradation_rate(rows)
    add_aval(rows, d_rate, degradationBias)

    # removeUnObservedGenes (:301-315)
    gsum = {}
    for r in rows:
        gsum[r["gene_sid"]] = gsum.get(r["gene_sid"], 0.0) + r["nobs"]
    unobserved = OrderedDict()
    for r in rows:
        if gsum[r["gene_sid"]] == 0:
            unobserved[(r["txid"], r["gene_sid"])] = None
    rows = [r for r in rows if gsum[r["gene_sid"]] != 0]

    # K, n.obs (:221-228); right_join puts rows in first-appearance order of (gene_sid, eqClassId)
    uniq = OrderedDict()
    for r in rows:
        uniq[(r["gene_sid"], r["eqClassId"], r["nobs"])] = None
    K = {}
    for (g, _e, nobs) in uniq:
        K.setdefault(g, []).append(nobs)
    K = {g: _r_sum(v) for g, v in K.items()}
    order = OrderedDict()
    for r in rows:
        order.setdefault((r["gene_sid"], r["eqClassId"]), []).append(r)
    rows = [r for grp in order.values() for r in grp]
    for r in rows:
        r["K"] = K[r["gene_sid"]]
        r["n.obs"] = r["nobs"] / r["K"]

    # initialiseOutput (:322-327)
    out = OrderedDict()
    for r in rows:
        out.setdefault(r["txid"], [0.0, 0.0, 0.0])

    # filterTxRc (:332-342)
    tsum = {}
    for r in rows:
        k = (r["gene_sid"], r["txid"])
        tsum[k] = tsum.get(k, 0.0) + r["nobs"]
    rows = [r for r in rows if tsum[(r["gene_sid"], r["txid"])] > 0]
    rows.sort(key=lambda r: (r["gene_sid"], r["txid"], r["eqClassId"]))
    for r in rows:
        r["uniqueAval"] = r["aval"] * (0.0 if r["multi_align"] else 1.0)
        r["fullAval"] = r["aval"] * (1.0 if r["equal"] else 0.0)

    # assignGroups (:347-356)
    gtx, geq = OrderedDict(), {}
    for r in rows:
        gtx.setdefault(r["gene_sid"], set()).add(r["txid"])
        geq.setdefault(r["gene_sid"], set()).add(r["eqClassId"])
    dims = [(g, len(gtx[g]) * len(geq[g])) for g in gtx]           # table order = gene_sid ascending
    cum = 0
    grp_of = {}
    for g, d in sorted(dims, key=lambda t: t[1]):                   # stable, like data.table order()
        cum += d
        grp_of[g] = cum // 1000 + 1
    for r in rows:
        r["grp"] = grp_of[r["gene_sid"]]

    # getInputList + run_parallel (:359-431)
    groups = []
    est_rows = []  # (txid, counts, full, unique) in rbind order
    for gid in sorted(set(grp_of.values())):
        grows = [r for r in rows if r["grp"] == gid]
        cols = sorted({(r["gene_sid"], r["eqClassId"]) for r in grows})
        txids = sorted({r["txid"] for r in grows})
        cpos = {c: j for j, c in enumerate(cols)}
        rpos = {t: i for i, t in enumerate(txids)}
        M, N = len(txids), len(cols)
        A = np.zeros((M, N))
        Af = np.zeros((M, N))
        Au = np.zeros((M, N))
        Yv = np.zeros(N)
        Kv = np.zeros(N)
        for r in grows:
            i, j = rpos[r["txid"]], cpos[(r["gene_sid"], r["eqClassId"])]
            A[i, j] = r["aval"]
            Af[i, j] = r["fullAval"]
            Au[i, j] = r["uniqueAval"]
            Yv[j] = r["n.obs"]
            Kv[j] = r["K"]
        if M == 1:  # :410-414
            est = np.array([[_r_sum(Kv * Yv * A[0])], [_r_sum(Kv * Yv * Af[0])], [_r_sum(Kv * Yv * Au[0])]])
            iters = 0
        else:
            est, iters = emWithL1(A, Af, Au, Yv, Kv, maxiter=maxiter, minvalue=minvalue, conv=conv)
        for i, t in enumerate(txids):
            est_rows.append((t, est[0, i], est[1, i], est[2, i]))
        groups.append({"gene_grp_id": gid, "txids": txids, "cols": cols, "A": A, "A_full": Af, "A_unique": Au,
                       "Y": Yv, "K": Kv, "est": est, "iters": iters})

    # modifyQuantOut (:438-443): positional, duplicates -> last wins
    for (t, c0, c1, c2) in est_rows:
        out[t] = [c0, c1, c2]
    # rbind(unobserved, outEst) then removeDuplicates (:449-455): sum by txid, first-appearance order
    final = OrderedDict()
    for (t, _g) in unobserved:
        final.setdefault(t, []).append((0.0, 0.0, 0.0))
    for t, v in out.items():
        final.setdefault(t, []).append(tuple(v))
    result = OrderedDict()
    for t, vs in final.items():
        result[t] = tuple(_r_sum([v[k] for v in vs]) for k in range(3))
    if details is not None:
        details["d_rate"] = d_rate
        details["groups"] = groups
    return result
