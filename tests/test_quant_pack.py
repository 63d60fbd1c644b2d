"""The native packer / merger of libbambu_em.so (csrc/quant_pack.cpp: addAval ... assignGroups ...
getInputList, modifyQuantOut + removeDuplicates) against the numpy mirror (bambu_b200/quantify.py),
which tests/test_quantify_host.py pins on the reference goldens and the loop-level restatement.
Everything here is host code: no GPU needed.  Bar: identical arrays, identical doubles."""
import ctypes
import json
import os

import numpy as np
import pytest

from bambu_b200 import quantify as Q
from conftest import GOLDEN, assert_same_doubles
from test_quantify_host import random_long_table


@pytest.fixture(scope="module", autouse=True)
def _built():
    from bambu_b200 import _lib
    _lib.build()


def same_pack(a: Q.PackedQuant, b: Q.PackedQuant):
    assert (a.d_rate == b.d_rate) or (np.isnan(a.d_rate) and np.isnan(b.d_rate))
    for name in ("grp_row_ptr", "grp_col_ptr", "row_nnz_ptr", "col_idx", "nz_flags", "col_flags", "txid"):
        assert np.array_equal(getattr(a.arena, name), getattr(b.arena, name)), name
    for name in ("aval", "Y", "K"):
        assert_same_doubles(getattr(a.arena, name), getattr(b.arena, name), name)
    assert np.array_equal(a.out_txid, b.out_txid)
    assert np.array_equal(a.unobserved_txid, b.unobserved_txid)
    assert np.array_equal(a.gene_grp_id, b.gene_grp_id)


@pytest.mark.parametrize("bias", [False, True])
def test_goldens(unit_goldens, bias):
    for case in unit_goldens["cases"]:
        same_pack(Q.pack_quantDT_native(case["input"], {"degradationBias": bias}),
                  Q.pack_quantDT(case["input"], {"degradationBias": bias}))


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("bias", [False, True])
def test_random_long_tables(seed, bias):
    rng = np.random.default_rng(500 + seed)
    tab = random_long_table(rng, n_genes=int(rng.integers(2, 400)))
    same_pack(Q.pack_quantDT_native(tab, {"degradationBias": bias}), Q.pack_quantDT(tab, {"degradationBias": bias}))


def test_zero_degradation_rate_and_no_partial_classes():
    # d_rate == 0 branch (:283-289) and d_rate NA (no partial class anywhere)
    tab = {"GENEID": ["a"] * 6, "txid": [1, 2, 1, 2, 1, 2], "eqClassId": [1, 1, 2, 3, 4, 4], "nobs": [0.0, 0.0, 40.0, 5.0, 0.0, 0.0],
           "equal": [False, False, True, True, True, False], "rcWidth": [500.0, 500.0, 900.0, 800.0, 300.0, 300.0],
           "txlen": [1000.0] * 6}
    a, b = Q.pack_quantDT_native(tab), Q.pack_quantDT(tab)
    assert a.d_rate == 0.0
    same_pack(a, b)
    tab2 = dict(tab, equal=[True] * 6)
    a, b = Q.pack_quantDT_native(tab2), Q.pack_quantDT(tab2)
    assert np.isnan(a.d_rate)
    same_pack(a, b)


def test_chr9_long_table():
    from oracle.equirc_ref import gen_equi_rcs
    with open(os.path.join(GOLDEN, "chr9_e2e.json")) as f:
        tab = gen_equi_rcs(json.load(f))
    a, b = Q.pack_quantDT_native(tab), Q.pack_quantDT(tab)
    same_pack(a, b)
    assert a.d_rate == 0.08893379057159526


def test_errors():
    with pytest.raises(ValueError, match="missing"):
        Q.pack_quantDT_native({"txid": [1], "nobs": [1]})
    bad = {"GENEID": [1, 1], "txid": [1, 2], "eqClassId": [7, 7], "nobs": [3.0, 4.0]}
    with pytest.raises(ValueError, match="more than one nobs"):
        Q.pack_quantDT_native(bad)
    dup = {"GENEID": [1, 1], "txid": [1, 1], "eqClassId": [7, 7], "nobs": [3.0, 3.0]}
    with pytest.raises(ValueError, match="duplicated"):
        Q.pack_quantDT_native(dup)


def test_merge_matches_numpy(oracle_mod):
    """bambu_quant_merge == modifyQuantOut + removeDuplicates of the numpy mirror, incl. a txid
    estimated under two genes (last group wins, then summed with the unobserved rows)."""
    from bambu_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(9)
    for _ in range(5):
        tab = random_long_table(rng, n_genes=int(rng.integers(5, 150)))
        pk = Q.pack_quantDT(tab)
        est = oracle_mod.em_batched(pk.arena)[0]
        want = Q.modifyQuantOut_removeDuplicates(pk, est)
        # native: rebuild the pack, merge with the same estimates
        gene = Q._gene_codes(tab["GENEID"])
        c = lambda a, dt: np.ascontiguousarray(np.asarray(a), dtype=dt)      # noqa: E731
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)                    # noqa: E731
        arrs = [gene, c(tab["txid"], np.int64), c(tab["eqClassId"], np.int64), c(tab["nobs"], np.float64),
                c(np.asarray(tab["equal"], bool), np.uint8), c(tab["rcWidth"], np.float64), c(tab["txlen"], np.float64)]
        h = ctypes.c_void_p()
        assert L.bambu_quant_pack_create(ctypes.byref(h), len(gene), *[ptr(a) for a in arrs], 1) == 0
        cap = len(gene)
        t = np.zeros(cap, np.int64)
        o = [np.zeros(cap) for _ in range(3)]
        k = ctypes.c_int64()
        estc = np.ascontiguousarray(est)
        assert L.bambu_quant_merge(h, ptr(estc), ptr(t), ptr(o[0]), ptr(o[1]), ptr(o[2]), ctypes.byref(k)) == 0
        L.bambu_quant_pack_free(h)
        m = int(k.value)
        assert t[:m].tolist() == want["txid"].tolist()
        for got, name in zip(o, ("counts", "fullLengthCounts", "uniqueCounts")):
            assert_same_doubles(got[:m], want[name], name)


@pytest.mark.gpu
def test_bambu_quantDT_one_call(unit_goldens):
    """tests/testthat/test_quant.R:12-27 through the single native entry point."""
    for bias, key in ((False, "estOutput_woBC"), (True, "estOutput_wBC")):
        for case in unit_goldens["cases"]:
            got = Q.bambu_quantDT_native(case["input"], dict(unit_goldens["emParameters"], degradationBias=bias))
            exp = case["expected"][key]
            assert got["txid"].tolist() == [int(t) for t in exp["txid"]]
            for name in ("counts", "fullLengthCounts", "uniqueCounts"):
                assert_same_doubles(got[name], np.array(exp[name]), f"{case['name']} {name}")
