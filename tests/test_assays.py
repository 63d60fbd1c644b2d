"""Host steps after bambu.quantDT (bambu_b200/assays.py) against the reference's stored outputs:
transcriptToGeneExpression on the four cases of tests/testthat/test_transcriptToGeneExpression.R:3-28
(inputs inst/extdata/seOutput*.rds, expected se*GeneExpected of R/sysdata.rda; tests/golden/gene_level.json),
rename_duplicatedNames, and the tail of bambu.quantify (CPM, annotation order, rounding) on the chr9 run."""
import json
import os

import numpy as np
import pytest

from bambu_b200 import assays as A
from conftest import GOLDEN, assert_same_doubles


@pytest.fixture(scope="module")
def gene_cases():
    with open(os.path.join(GOLDEN, "gene_level.json")) as f:
        return json.load(f)["cases"]


def _mat(values, dim):
    v = np.array([np.nan if x is None else x for x in values], dtype=np.float64)
    return v.reshape(dim[1], dim[0]).T          # R matrices are column-major


def test_transcriptToGeneExpression_goldens(gene_cases):
    assert [c["name"] for c in gene_cases] == ["seGeneExpected", "seExtendedGeneExpected", "seCombinedGeneExpected",
                                               "seCombinedExtendedGeneExpected"]
    for c in gene_cases:
        i, e = c["input"], c["expected"]
        genes, names, cg, cpm = A.transcriptToGeneExpression(_mat(i["counts"], i["dim"]), i["GENEID"], i["sample_names"],
                                                             i["incompatibleCounts"])
        assert genes == e["GENEID"], c["name"]
        assert names == e["sample_names"], c["name"]
        assert_same_doubles(cg, _mat(e["counts"], e["dim"]), c["name"] + " counts")
        assert_same_doubles(cpm, _mat(e["CPM"], e["dim"]), c["name"] + " CPM")


def test_incompatible_reads_count_only_under_a_matching_name(gene_cases):
    """In the combined objects the merged incompatible columns are called *.x / *.y, so :28 drops them."""
    single, combined = gene_cases[0], gene_cases[2]
    inc = single["input"]["incompatibleCounts"]
    assert sum(inc[single["input"]["sample_names"][0]]) > 0
    tx_total = np.nansum(_mat(single["input"]["counts"], single["input"]["dim"]))
    got = np.nansum(_mat(single["expected"]["counts"], single["expected"]["dim"]))
    assert abs(got - (tx_total + sum(inc[single["input"]["sample_names"][0]]))) < 1e-6
    tx_total2 = np.nansum(_mat(combined["input"]["counts"], combined["input"]["dim"]))
    assert abs(np.nansum(_mat(combined["expected"]["counts"], combined["expected"]["dim"])) - tx_total2) < 1e-6


def test_rename_duplicatedNames():
    assert A.rename_duplicatedNames(["a", "b"]) == ["a", "b"]
    assert A.rename_duplicatedNames(["a", "a"]) == ["a", "a...1"]
    assert A.rename_duplicatedNames(["a", "a", "a", "b", "b"]) == ["a", "a...1", "a...2", "b", "b...1"]


def test_sequential_long_double_sums():
    rng = np.random.default_rng(3)
    seg = rng.integers(0, 40, 3000)
    val = np.round(rng.gamma(0.5, 50.0, 3000), 5)
    got = A._seq_sum_by(val, seg, 40)
    for g in range(40):
        s = np.longdouble(0)
        for v in val[seg == g]:
            s += np.longdouble(v)
        assert got[g] == s


def test_quantify_assays_chr9(oracle_mod):
    """tests/testthat/test_bambu_arguments.R:33-43 / test_quant.R:57-66: the four assays of the bundled run."""
    from oracle.equirc_ref import gen_equi_rcs
    from bambu_b200.quantify import bambu_quantDT
    with open(os.path.join(GOLDEN, "chr9_e2e.json")) as f:
        fx = json.load(f)
    quant = bambu_quantDT(gen_equi_rcs(fx), {"degradationBias": True},
                          em=lambda a, maxiter, conv, minvalue: oracle_mod.em_batched(a, maxiter=maxiter, conv=conv,
                                                                                      minvalue=minvalue))
    inc = [np.nan if v is None else v for v in fx["incompatibleCounts"]["counts"]]
    got = A.quantify_assays(quant, fx["annotation"]["txid"], inc)
    for name in ("counts", "CPM", "fullLengthCounts", "uniqueCounts"):
        assert_same_doubles(got[name], np.array(fx["expected_assays"][name], dtype=np.float64), name)
    both, _ = A.combine_assays([got, got], ["s1", "s2"])
    assert both["counts"].shape == (len(fx["annotation"]["txid"]), 2)


def test_r_round():
    assert A.r_round([2.5, 1e-9, 123.456784999], 5).tolist() == [2.5, 0.0, 123.45678]
    assert np.isnan(A.r_round([np.nan], 5)[0])
    x = np.array([0.5, 1.5, 2.5, 0.125, 0.375])
    assert A.r_round(x, 0).tolist() == [0.0, 2.0, 2.0, 0.0, 0.0]
    assert A.r_round(x, 2).tolist() == [0.5, 1.5, 2.5, 0.12, 0.38]


def test_genEquiRCs_chr9(oracle_mod):
    """The product's genEquiRCs (bambu_b200/equirc.py) on the bundled run: the same long table as the loop-level
    restatement (oracle/equirc_ref.py; 491 rows, 222 classes -- SURVEY 8c), and through bambu.quantDT and
    quantify_assays the four assays the reference stores (tests/testthat/test_bambu_arguments.R:33-43)."""
    from oracle.equirc_ref import gen_equi_rcs
    from bambu_b200.equirc import AnnotationClasses, genEquiRCs
    from bambu_b200.quantify import bambu_quantDT
    with open(os.path.join(GOLDEN, "chr9_e2e.json")) as f:
        fx = json.load(f)
    ann = AnnotationClasses(fx["annotation"])
    tab = genEquiRCs(fx["distTable"], fx["readClasses"], ann)
    ref = gen_equi_rcs(fx)
    assert len(tab["txid"]) == 491 and len(set(tab["eqClassId"].tolist())) == 222
    for col in ("GENEID", "txid", "eqClassId", "nobs", "equal", "rcWidth", "txlen", "minRC"):
        assert list(tab[col]) == list(ref[col]), col
    again = genEquiRCs(fx["distTable"], fx["readClasses"], ann)          # the cached annotation side is not consumed
    assert again["eqClassId"].tolist() == tab["eqClassId"].tolist()
    quant = bambu_quantDT(tab, {"degradationBias": True},
                          em=lambda a, maxiter, conv, minvalue: oracle_mod.em_batched(a, maxiter=maxiter, conv=conv,
                                                                                      minvalue=minvalue))
    inc = [np.nan if v is None else v for v in fx["incompatibleCounts"]["counts"]]
    got = A.quantify_assays(quant, fx["annotation"]["txid"], inc)
    for name in ("counts", "CPM", "fullLengthCounts", "uniqueCounts"):
        assert_same_doubles(got[name], np.array(fx["expected_assays"][name], dtype=np.float64), name)


def test_duplicated_sample_chain(gene_cases, oracle_mod):
    """tests/testthat/test_quant.R:63-66 / test_transcriptToGeneExpression.R:18-21: the same sample twice, from the distance
    table to the gene level -- genEquiRCs -> bambu.quantDT -> quantify_assays -> combine_assays reproduces the counts
    stored in seOutputCombined_*.rds, and transcriptToGeneExpression on them gives seCombinedGeneExpected."""
    from bambu_b200.equirc import AnnotationClasses, genEquiRCs
    from bambu_b200.quantify import bambu_quantDT
    with open(os.path.join(GOLDEN, "chr9_e2e.json")) as f:
        fx = json.load(f)
    ann = AnnotationClasses(fx["annotation"])
    inc = [np.nan if v is None else v for v in fx["incompatibleCounts"]["counts"]]
    one = []
    for _ in range(2):
        quant = bambu_quantDT(genEquiRCs(fx["distTable"], fx["readClasses"], ann), {"degradationBias": True},
                              em=lambda a, maxiter, conv, minvalue: oracle_mod.em_batched(a, maxiter=maxiter, conv=conv,
                                                                                          minvalue=minvalue))
        one.append(A.quantify_assays(quant, fx["annotation"]["txid"], inc))
    combined = gene_cases[2]
    assays, _ = A.combine_assays(one, combined["input"]["sample_names"])
    assert_same_doubles(assays["counts"], _mat(combined["input"]["counts"], combined["input"]["dim"]), "combined counts")
    genes, names, cg, cpm = A.transcriptToGeneExpression(assays["counts"], fx["annotation"]["GENEID"],
                                                         combined["input"]["sample_names"], combined["input"]["incompatibleCounts"])
    e = combined["expected"]
    assert genes == e["GENEID"] and names == e["sample_names"]
    assert_same_doubles(cg, _mat(e["counts"], e["dim"]), "gene counts")
    assert_same_doubles(cpm, _mat(e["CPM"], e["dim"]), "gene CPM")
