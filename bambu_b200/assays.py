"""What follows ``bambu.quantDT`` in the reference, on the host:

* :func:`quantify_assays` -- the tail of ``bambu.quantify`` (R/bambu-quantify.R:23-38): CPM with the
  incompatible counts, rows in annotation order, ``round(., sig.digit)`` of counts / CPM / fullLengthCounts;
* :func:`combine_assays` -- ``combineCountSes`` (R/bambu_utilityFunctions.R:215-241): samples side by side,
  incompatible counts merged by GENEID;
* :func:`transcriptToGeneExpression` -- R/transcriptToGeneExpression.R:13-67: gene-level counts (transcripts
  plus the gene's incompatible reads) and gene CPM.

Same function names, argument meaning and output layout as the R code; R's ``sum()`` accumulates in long double
and in table order, so do the sums here.  Pinned on the reference's stored outputs (tests/test_assays.py)."""
from __future__ import annotations

import numpy as np

from .quantify import calculateCPM

LD = np.longdouble


def r_round(x, digits):
    """R's ``round(x, digits)`` (R >= 4.0: the closest ``digits``-decimal number to the double, ties to even)."""
    x = np.asarray(x, dtype=np.float64)
    flat = x.ravel()
    out = np.array([float(format(v, ".%df" % digits)) if np.isfinite(v) else v for v in flat], dtype=np.float64)
    return out.reshape(x.shape)


def quantify_assays(quant, annotation_txid, incompatible_counts, sig_digit=5):
    """One sample's four assays, one value per annotation transcript (R/bambu-quantify.R:23-38).

    ``quant``: the result of ``bambu_quantDT`` (txid, counts, fullLengthCounts, uniqueCounts);
    ``incompatible_counts``: per-gene incompatible read counts (NA = 0) -- they enter the CPM denominator
    (``calculateCPM``, R/bambu-quantify_utilityFunctions.R:493-497).  ``uniqueCounts`` is not rounded (:36-37)."""
    inc = np.asarray(incompatible_counts, dtype=np.float64)
    inc = np.where(np.isnan(inc), 0.0, inc)
    counts = np.asarray(quant["counts"], dtype=np.float64)
    cpm = calculateCPM(counts, inc)
    txid = np.asarray(quant["txid"], dtype=np.int64)
    srt = np.argsort(txid, kind="stable")
    want = np.asarray(annotation_txid, dtype=np.int64)
    pos = np.searchsorted(txid[srt], want)
    found = (pos < len(txid)) & (txid[srt][np.minimum(pos, len(txid) - 1)] == want) if len(txid) else np.zeros(len(want), bool)
    order = srt[np.minimum(pos, max(len(txid) - 1, 0))] if len(txid) else np.zeros(len(want), np.int64)

    def take(v, rounded):          # match(): transcripts without a row come out NA
        v = np.asarray(v, dtype=np.float64)
        col = np.where(found, v[order] if len(v) else np.nan, np.nan)
        return r_round(col, sig_digit) if rounded else col
    return {"counts": take(counts, True), "CPM": take(cpm, True),
            "fullLengthCounts": take(quant["fullLengthCounts"], True), "uniqueCounts": take(quant["uniqueCounts"], False)}


def combine_assays(samples, sample_names, incompatible=None, gene_ids=None):
    """``combineCountSes``: ``samples`` = list of :func:`quantify_assays` results on the same annotation.
    Returns (assays dict of n_tx x n_samples matrices, incompatibleCounts dict GENEID -> columns)."""
    assays = {k: np.column_stack([s[k] for s in samples]) for k in ("counts", "CPM", "fullLengthCounts", "uniqueCounts")}
    inc = None
    if incompatible is not None:
        inc = {"GENEID": list(gene_ids), "columns": list(sample_names),
               "values": np.column_stack([np.asarray(v, dtype=np.float64) for v in incompatible])}
    return assays, inc


def rename_duplicatedNames(runnames):
    """R/transcriptToGeneExpression_utilityFunctions.R:6-22: later duplicates get ``...1``, ``...2``, ..."""
    names = list(runnames)

    def dup(v):
        seen, out = set(), []
        for k, n in enumerate(v):
            if n in seen:
                out.append(k)
            seen.add(n)
        return out
    it = 1
    while dup(names):
        for k in dup(names):
            if it == 1:
                names[k] = names[k] + "...1"
            elif names[k].endswith("...%d" % (it - 1)):
                names[k] = names[k][: -len("...%d" % (it - 1))] + "...%d" % it
        it += 1
    return names


def _seq_sum_by(values, seg, n_seg):
    """Per-segment sum in long double, elements added in their order of appearance (R's ``sum()`` by group)."""
    values = np.asarray(values, dtype=np.float64)
    order = np.argsort(seg, kind="stable")
    seg_s, val_s = seg[order], values[order].astype(LD)
    start = np.searchsorted(seg_s, np.arange(n_seg))
    count = np.bincount(seg, minlength=n_seg)
    acc = np.zeros(n_seg, dtype=LD)
    for k in range(int(count.max()) if len(count) else 0):          # k-th member of every segment at once
        live = np.nonzero(count > k)[0]
        acc[live] = acc[live] + val_s[start[live] + k]
    return acc


def transcriptToGeneExpression(counts, GENEID, sample_names, incompatibleCounts=None):
    """Gene-level counts and CPM (R/transcriptToGeneExpression.R:13-67).

    ``counts``: n_tx x n_samples transcript counts; ``GENEID``: gene of every transcript (rowData order);
    ``incompatibleCounts``: ``{"GENEID": [...], "<column name>": [...], ...}`` as in ``metadata(se)`` -- a column is
    used for the sample of the same (de-duplicated) name only (:28).  Returns ``(gene_ids, sample_names, counts_gene,
    CPM_gene)`` with genes in ascending order (``dcast``) and
    ``CPM = valueGene / max(sum(all values of the sample), 1) * 1e6`` (:30-31)."""
    counts = np.asarray(counts, dtype=np.float64)
    if counts.ndim == 1:
        counts = counts[:, None]
    names = rename_duplicatedNames(sample_names)
    genes = sorted(set(GENEID) | (set(incompatibleCounts["GENEID"]) if incompatibleCounts else set()))
    gidx = {g: k for k, g in enumerate(genes)}
    tx_gene = np.array([gidx[g] for g in GENEID], dtype=np.int64)
    G = len(genes)
    out_c = np.full((G, counts.shape[1]), np.nan)
    out_p = np.full((G, counts.shape[1]), np.nan)
    for s, name in enumerate(names):
        vals, seg = counts[:, s], tx_gene
        if incompatibleCounts is not None and name in incompatibleCounts and name != "GENEID":
            inc_gene = np.array([gidx[g] for g in incompatibleCounts["GENEID"]], dtype=np.int64)
            inc_val = np.asarray(incompatibleCounts[name], dtype=np.float64)
            vals, seg = np.concatenate((vals, inc_val)), np.concatenate((seg, inc_gene))      # rbind: after the transcripts
        gene_sum = _seq_sum_by(vals, seg, G).astype(np.float64)
        present = np.bincount(seg, minlength=G) > 0
        total = float(np.cumsum(vals.astype(LD))[-1]) if len(vals) else 0.0
        out_c[present, s] = gene_sum[present]
        out_p[present, s] = gene_sum[present] / max(total, 1.0) * 10 ** 6
    return genes, names, out_c, out_p
