"""BASELINE.json configs[0]: the reference's bundled chr9:1-1Mb run (SGNex A549 directRNA),
replayed without R from the objects stored in inst/extdata/seOutput_distTable_*.rds and
seReadClassUnstranded_*.rds (tests/golden/chr9_e2e.json; tests/testthat/test_bambu_arguments.R:33-43
and test_quant.R:57-66 compare the same four assays).  The whole chain is product code -- genEquiRCs
(bambu_b200/equirc.py), bambu.quantDT (quantify.py + the EM), the tail of bambu.quantify (assays.py); only the EM
callable changes between the CPU run (oracle) and the -m gpu run (libbambu_em.so).  oracle/equirc_ref.py keeps an
independent loop-level restatement of genEquiRCs."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_same_doubles


@pytest.fixture(scope="module")
def fx():
    with open(os.path.join(GOLDEN, "chr9_e2e.json")) as f:
        return json.load(f)


def _run(fx, em):
    from bambu_b200.assays import quantify_assays
    from bambu_b200.equirc import genEquiRCs
    from bambu_b200.quantify import bambu_quantDT
    tab = genEquiRCs(fx["distTable"], fx["readClasses"], fx["annotation"])
    assert len(tab["txid"]) == 491 and len(set(tab["eqClassId"])) == 222
    det = {}
    quant = bambu_quantDT(tab, {"degradationBias": True}, em=em, details=det)
    inc = [np.nan if v is None else v for v in fx["incompatibleCounts"]["counts"]]
    return quantify_assays(quant, fx["annotation"]["txid"], inc), det


def _check(fx, assays, det):
    assert det["d_rate"] == 0.08893379057159526
    a = det["arena"]
    M, N, _ = a.group_shapes()
    assert M.tolist() == [10, 19] and N.tolist() == [26, 52]
    assert a.group_nonzeros().tolist() == [47, 80]
    assert det["iters"].tolist() == [248, 54]
    for name in ("counts", "CPM", "fullLengthCounts", "uniqueCounts"):
        want = np.array(fx["expected_assays"][name], dtype=np.float64)
        assert_same_doubles(assays[name], want, name)
    assert abs(assays["counts"].sum() - 88) < 1e-4


def test_chr9_with_oracle_em(fx, oracle_mod):
    def em(arena, maxiter, conv, minvalue):
        return oracle_mod.em_batched(arena, maxiter=maxiter, conv=conv, minvalue=minvalue, mode=oracle_mod.MODE_DENSE)
    assays, det = _run(fx, em)
    _check(fx, assays, det)


def test_quant_ref_agrees_on_chr9(fx, oracle_mod):
    from oracle.equirc_ref import gen_equi_rcs
    from oracle.quant_ref import quantDT_ref
    from bambu_b200.quantify import bambu_quantDT
    tab = gen_equi_rcs(fx)
    ref = quantDT_ref(tab, degradationBias=True)
    got = bambu_quantDT(tab, {"degradationBias": True},
                        em=lambda a, maxiter, conv, minvalue: oracle_mod.em_batched(a, maxiter=maxiter, conv=conv,
                                                                                    minvalue=minvalue))
    assert got["txid"].tolist() == list(ref.keys())
    for k, name in enumerate(("counts", "fullLengthCounts", "uniqueCounts")):
        assert_same_doubles(got[name], np.array([v[k] for v in ref.values()]), name)


@pytest.mark.gpu
def test_chr9_through_cuda(fx):
    assays, det = _run(fx, None)      # default EM: libbambu_em.so
    _check(fx, assays, det)


def test_native_genEquiRCs_on_chr9(fx):
    """csrc/equirc.cpp on the bundled run: the 491-row table of the Python mirror, row for row"""
    from bambu_b200.equirc import genEquiRCs, genEquiRCs_native
    a = genEquiRCs(fx["distTable"], fx["readClasses"], fx["annotation"])
    b = genEquiRCs_native(fx["distTable"], fx["readClasses"], fx["annotation"])
    assert len(b["txid"]) == 491 and len(set(b["eqClassId"])) == 222
    for col in ("GENEID", "txid", "eqClassId", "nobs", "equal", "rcWidth", "txlen", "minRC"):
        assert list(a[col]) == list(b[col]), col
