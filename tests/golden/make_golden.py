#!/usr/bin/env python
"""Regenerate the committed golden fixtures from the reference's own test data.

Run in the BUILD container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.

Outputs (all under ``tests/golden/``):

* ``quantDT_unit.json`` - the inputs ``data1..data5`` and the expected
  ``bambu.quantDT`` tables ``estOutput_woBC[[1..5]]`` / ``estOutput_wBC[[1..5]]``
  exactly as stored in ``R/sysdata.rda`` (the tables that
  ``tests/testthat/test_quant.R:12-27`` compares against; inputs defined at
  ``data-raw/DATASET.R:17-75``).
* ``chr9_e2e.json`` - everything needed to replay the quantification stage of
  the bundled chr9:1-1Mb run without R: the distance table, read-class widths
  and the annotation columns stored inside
  ``inst/extdata/seOutput_distTable_*.rds`` / ``seReadClassUnstranded_*.rds``
  (written by ``data-raw/DATASET.R:245-246``), plus the four expected assays
  (``tests/testthat/test_bambu_arguments.R:33-43``).

Floats are written with ``repr`` (shortest round-trip), so the JSON carries
the doubles bit for bit.
"""
from __future__ import annotations

import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rds_reader import read_rda, read_rds  # noqa: E402

REF = "/root/reference"
STEM = "SGNex_A549_directRNA_replicate5_run1_chr9_1_1000000"


def _nan_to_none(xs):
    return [None if (isinstance(x, float) and math.isnan(x)) else x for x in xs]


def unit_goldens():
    items = dict(read_rda(os.path.join(REF, "R/sysdata.rda"), 7))
    out = {"source": "R/sysdata.rda (objects data1..5, estOutput_woBC, estOutput_wBC)",
           "emParameters": {"maxiter": 10000, "conv": 1e-2, "minvalue": 1e-8},
           "cases": []}
    for s in range(1, 6):
        d = items["data%d" % s]
        inp = {}
        for col in ("txid", "equal", "eqClassId", "nobs", "txlen", "rcWidth", "GENEID"):
            inp[col] = d[col].value
        inp["minRC"] = _nan_to_none(d["minRC"].value)
        exp = {}
        for key in ("estOutput_woBC", "estOutput_wBC"):
            t = items[key].value[s - 1]
            exp[key] = {n: t[n].value for n in t.names()}
        out["cases"].append({"name": "data%d" % s, "input": inp, "expected": exp})
    return out


def _glist(grl):
    """CompressedGRangesList -> (names, per-element exon arrays, mcols)."""
    ul = grl.attrs["unlistData"]
    rng = ul.attrs["ranges"]
    start = rng.attrs["start"].value
    width = rng.attrs["width"].value
    part = grl.attrs["partitioning"]
    ends = part.attrs["end"].value
    names = part.attrs["NAMES"].value if part.attrs.get("NAMES") is not None else None
    return names, start, width, ends, ul, grl.attrs["elementMetadata"]


def chr9_e2e():
    se = read_rds(os.path.join(REF, "inst/extdata/seOutput_distTable_%s.rds" % STEM))
    rc = read_rds(os.path.join(REF, "inst/extdata/seReadClassUnstranded_%s.rds" % STEM))

    # --- annotation (rowRanges of the output SE) ---------------------------
    names, start, width, ends, ul, em = _glist(se.attrs["rowRanges"])
    ld = em.attrs["listData"]
    eq_by = ld["eqClassById"]
    eq_ul = eq_by.attrs["unlistData"].value
    eq_end = eq_by.attrs["partitioning"].attrs["end"].value
    ann = {
        "TXNAME": ld["TXNAME"].value,
        "GENEID": ld["GENEID"].value,
        "txid": ld["txid"].value,
        "eqClassById_values": eq_ul,
        "eqClassById_end": eq_end,
        "names": names,
        "exon_width": width,
        "exon_end": ends,
    }

    # --- distance table (already post-modifyIncompatibleAssignment) -------
    dt = se.attrs["metadata"]["distTables"].value[0].attrs["listData"]
    dist = {k: dt[k].value for k in
            ("annotationTxId", "txid", "readClassId", "readCount", "compatible", "equal", "dist", "GENEID")}
    dist["eqClassById"] = [e.value for e in dt["eqClassById"].value]

    inc = se.attrs["metadata"]["incompatibleCounts"]
    inc_names = inc.names()
    incompatible = {"GENEID": inc[inc_names[0]].value, "counts": inc[inc_names[1]].value}

    assays = se.attrs["assays"].attrs["data"].attrs["listData"]
    expected = {n: assays[n].value for n in ("counts", "CPM", "fullLengthCounts", "uniqueCounts")}

    # --- read classes: exon widths + exon_rank + counts -------------------
    rnames, rstart, rwidth, rends, rul, rem = _glist(rc.attrs["rowRanges"])
    rmc = rul.attrs["elementMetadata"].attrs["listData"]
    rcounts = rc.attrs["assays"].attrs["data"].attrs["listData"]["counts"].value
    rld = rem.attrs["listData"]
    strand_rle = rul.attrs["strand"]
    read_classes = {
        "names": rnames,
        "exon_width": rwidth,
        "exon_end": rends,
        "exon_rank": rmc["exon_rank"].value,
        "counts": rcounts,
        "rowData_names": rld.names(),
    }
    for key in ("GENEID", "novelGene", "readCount"):
        if key in (rld.names() or []):
            read_classes[key] = rld[key].value
    return {
        "source": "inst/extdata/seOutput_distTable_%s.rds + seReadClassUnstranded_%s.rds" % (STEM, STEM),
        "annotation": ann, "distTable": dist, "incompatibleCounts": incompatible,
        "readClasses": read_classes, "expected_assays": expected,
    }


def gene_level():
    """Inputs and expected outputs of tests/testthat/test_transcriptToGeneExpression.R:3-28: the four stored
    bambu outputs of inst/extdata and the se*GeneExpected objects of R/sysdata.rda (data-raw/DATASET.R:92-102)."""
    import sys as _sys
    _sys.setrecursionlimit(100000)
    objs = dict(read_rda(os.path.join(REF, "R/sysdata.rda"), 13))
    cases = []
    for fname, key in (("seOutput_%s.rds", "seGeneExpected"), ("seOutputExtended_%s.rds", "seExtendedGeneExpected"),
                       ("seOutputCombined_%s.rds", "seCombinedGeneExpected"),
                       ("seOutputCombinedExtended_%s.rds", "seCombinedExtendedGeneExpected")):
        se = read_rds(os.path.join(REF, "inst/extdata", fname % STEM))
        ld = se.attrs["assays"].attrs["data"].attrs["listData"]
        counts = ld["counts"]
        em = se.attrs["rowRanges"].attrs["elementMetadata"].attrs["listData"]
        inc = se.attrs["metadata"]["incompatibleCounts"]
        exp = objs[key]
        eld = exp.attrs["assays"].attrs["data"].attrs["listData"]
        dn = eld["counts"].attrs["dimnames"].value
        cases.append({
            "name": key, "source": fname % STEM,
            "input": {"counts": _nan_to_none(counts.value), "dim": counts.attrs["dim"].value,
                      "TXNAME": em["TXNAME"].value, "GENEID": em["GENEID"].value,
                      "sample_names": se.attrs["colData"].attrs["rownames"].value,
                      "incompatibleCounts": {n: inc[n].value for n in inc.names()}},
            "expected": {"GENEID": dn[0].value, "sample_names": dn[1].value, "dim": eld["counts"].attrs["dim"].value,
                         "counts": _nan_to_none(eld["counts"].value), "CPM": _nan_to_none(eld["CPM"].value)}})
    return {"cases": cases}


def main():
    with open(os.path.join(HERE, "quantDT_unit.json"), "w") as f:
        json.dump(unit_goldens(), f, indent=1)
    sys.setrecursionlimit(20000)
    with open(os.path.join(HERE, "chr9_e2e.json"), "w") as f:
        json.dump(chr9_e2e(), f)
    with open(os.path.join(HERE, "gene_level.json"), "w") as f:
        json.dump(gene_level(), f)
    print("wrote quantDT_unit.json, chr9_e2e.json, gene_level.json")


if __name__ == "__main__":
    main()
