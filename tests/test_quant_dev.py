"""The device packer's pipeline (csrc/quant_dev.inl: the long table -> arena as sorts, run ids, scans and
order-free reductions) against the host packer (csrc/quant_pack.cpp), which tests/test_quant_pack.py pins
on the numpy mirror and the reference goldens.

CPU part: the same passes run serially by the test-only host executor (csrc/quant_dev_hostexec.cpp ->
libquant_hostexec.so, built by `make hostexec`; parallel loops run in REVERSE order so nothing can lean on
loop order).  GPU part (-m gpu): the CUDA executor through bambu_quant_pack_create_gpu / bambu_quantDT.
Bar: identical arrays, identical doubles."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from bambu_b200 import quantify as Q
from conftest import GOLDEN, assert_same_doubles
from test_quantify_host import random_long_table
from test_quant_pack import same_pack

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bambu_b200", "csrc")


@pytest.fixture(scope="module")
def hx():
    subprocess.run(["make", "-s", "hostexec"], cwd=CSRC, check=True)
    L = ctypes.CDLL(os.path.join(os.path.dirname(CSRC), "libquant_hostexec.so"))
    L.qd_hostexec_last_error.restype = ctypes.c_char_p
    L.qd_hostexec_ext_sum.restype = ctypes.c_double
    L.qd_hostexec_ld_sum.restype = ctypes.c_double
    L.qd_hostexec_ext_sum.argtypes = L.qd_hostexec_ld_sum.argtypes = [ctypes.c_void_p, ctypes.c_int64]
    return L


def _cols(tab):
    c = lambda a, dt: np.ascontiguousarray(np.asarray(a), dtype=dt)      # noqa: E731
    return [Q._gene_codes(tab["GENEID"]), c(tab["txid"], np.int64), c(tab["eqClassId"], np.int64), c(tab["nobs"], np.float64),
            c(np.asarray(tab["equal"], bool), np.uint8) if "equal" in tab else None,
            c(tab["rcWidth"], np.float64) if "rcWidth" in tab else None,
            c(tab["txlen"], np.float64) if "txlen" in tab else None]


def hostexec_pack(L, tab, bias):
    cols = _cols(tab)
    n = len(cols[0])
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None   # noqa: E731
    cap = max(n, 1)
    sizes = np.zeros(6, np.int64)
    bufs = dict(grp_row_ptr=np.zeros(cap + 1, np.int64), grp_col_ptr=np.zeros(cap + 1, np.int64),
                row_nnz_ptr=np.zeros(cap + 1, np.int64), col_idx=np.zeros(cap, np.int32), aval=np.zeros(cap),
                nz_flags=np.zeros(cap, np.uint8), Y=np.zeros(cap), K=np.zeros(cap), col_flags=np.zeros(cap, np.uint8),
                txid=np.zeros(cap, np.int32), out_txid=np.zeros(cap, np.int64), unobserved_txid=np.zeros(cap, np.int64),
                gene_grp_id=np.zeros(cap, np.int64))
    d = ctypes.c_double()
    rc = L.qd_hostexec_pack(ctypes.c_int64(n), *[ptr(a) for a in cols], int(bias), ptr(sizes),
                            *[ptr(b) for b in bufs.values()], ctypes.byref(d))
    if rc == 1:
        return None
    if rc:
        raise ValueError(L.qd_hostexec_last_error().decode())
    ng, rows, ncol, ne, n_out, n_uo = (int(v) for v in sizes)
    from bambu_b200.arena import Arena
    out = Q.PackedQuant()
    out.arena = Arena(bufs["grp_row_ptr"][:ng + 1].copy(), bufs["grp_col_ptr"][:ng + 1].copy(), bufs["row_nnz_ptr"][:rows + 1].copy(),
                      bufs["col_idx"][:ne].copy(), bufs["aval"][:ne].copy(), bufs["nz_flags"][:ne].copy(), bufs["Y"][:ncol].copy(),
                      bufs["K"][:ncol].copy(), bufs["col_flags"][:ncol].copy(), bufs["txid"][:rows].copy())
    out.out_txid = bufs["out_txid"][:n_out].copy()
    out.unobserved_txid = bufs["unobserved_txid"][:n_uo].copy()
    out.gene_grp_id = bufs["gene_grp_id"][:ng].copy()
    out.d_rate = float(d.value)
    return out


def test_r_round_matches_the_host_mirror(hx):
    """csrc/r_round.h (what the device epilogue of bambu_quantDT_cohort rounds with) == assays.r_round (the mirror of
    R's round() that is pinned on the reference's stored assays), incl. exact binary ties, tiny and huge values."""
    from bambu_b200.assays import r_round
    hx.qd_hostexec_r_round.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    hx.qd_hostexec_r_round.restype = None
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(0, 1e6, 60000), rng.uniform(0, 1, 30000) * 10.0 ** rng.integers(-12, 3, 30000),
                        np.round(rng.uniform(0, 100, 20000), 5) + 0.000005, (rng.integers(0, 10 ** 8, 20000) * 2 + 1) / 2e5,
                        np.array([0.0, -0.0, 2.5e-6, 1.5e-5, 0.5, 2.5, 1e300, 5e-324, 1e-20, 123456789.123455, 0.000005, 0.000015,
                                  0.000025, -3.141592653589793, 4.4093676540475e-07, np.nan, np.inf, -np.inf, 2.0 ** 53, 1e11 + 0.5])])
    for digits in (5, 0, 2, 8):
        out = np.zeros_like(x)
        hx.qd_hostexec_r_round(x.ctypes.data_as(ctypes.c_void_p), len(x), digits, out.ctypes.data_as(ctypes.c_void_p))
        assert_same_doubles(out, r_round(x, digits), f"round(x, {digits})")


def test_integer_x87_sum(hx):
    """ext_add / ext_to_double == `long double s; s += x; (double)s` (what R's sum() does on x86-64)."""
    rng = np.random.default_rng(5)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    for trial in range(4000):
        n = int(rng.integers(0, 40))
        kind = trial % 5
        if kind == 0:
            x = rng.uniform(0, 3, n)
        elif kind == 1:
            x = rng.uniform(0, 1, n) * 2.0 ** rng.integers(-80, 80, n)
        elif kind == 2:      # many ties: values with few mantissa bits at distant exponents
            x = rng.integers(1, 8, n).astype(np.float64) * 2.0 ** rng.integers(-70, 5, n)
        elif kind == 3:
            x = np.where(rng.random(n) < 0.3, 0.0, rng.uniform(0, 1e-3, n))
        else:                # includes denormals and huge values
            x = rng.uniform(0, 1, n) * 2.0 ** rng.integers(-1074, 900, n)
        x = np.ascontiguousarray(x)
        a, b = hx.qd_hostexec_ext_sum(ptr(x), n), hx.qd_hostexec_ld_sum(ptr(x), n)
        if np.isnan(a):      # result outside the normal double range: reported, not produced
            assert b == 0 or b < 2.3e-308 or np.isinf(b)
        else:
            assert a == b, (trial, x.tolist())
    # exact half-way cases at the 64-bit and the 53-bit rounding
    for x in ([1.0, 2.0 ** -64], [1.0, 2.0 ** -64, 2.0 ** -64], [1.0 + 2.0 ** -52, 2.0 ** -64], [1.0, 2.0 ** -53],
              [1.0 + 2.0 ** -52, 2.0 ** -53], [1.0, 2.0 ** -53, 2.0 ** -64], [2.0 - 2.0 ** -52, 2.0 ** -53, 2.0 ** -53]):
        x = np.array(x)
        assert hx.qd_hostexec_ext_sum(ptr(x), len(x)) == hx.qd_hostexec_ld_sum(ptr(x), len(x)), x.tolist()


@pytest.mark.parametrize("bias", [False, True])
def test_hostexec_goldens(hx, unit_goldens, bias):
    for case in unit_goldens["cases"]:
        a = hostexec_pack(hx, case["input"], bias)
        assert a is not None
        same_pack(a, Q.pack_quantDT_native(case["input"], {"degradationBias": bias}))


@pytest.mark.parametrize("seed", range(24))
@pytest.mark.parametrize("bias", [False, True])
def test_hostexec_random_long_tables(hx, seed, bias):
    rng = np.random.default_rng(4000 + seed)
    tab = random_long_table(rng, n_genes=int(rng.integers(2, 400)))
    a = hostexec_pack(hx, tab, bias)
    assert a is not None
    same_pack(a, Q.pack_quantDT_native(tab, {"degradationBias": bias}))


def test_hostexec_chr9(hx):
    from oracle.equirc_ref import gen_equi_rcs
    with open(os.path.join(GOLDEN, "chr9_e2e.json")) as f:
        tab = gen_equi_rcs(json.load(f))
    a = hostexec_pack(hx, tab, True)
    same_pack(a, Q.pack_quantDT_native(tab))
    assert a.d_rate == 0.08893379057159526


def test_hostexec_edge_cases(hx):
    # every gene unobserved: empty arena, all txids come back as unobserved
    tab = {"GENEID": [5, 5, 9], "txid": [3, 4, 8], "eqClassId": [1, 1, 2], "nobs": [0.0, 0.0, 0.0],
           "equal": [True, False, True], "rcWidth": [10.0, 20.0, 30.0], "txlen": [100.0, 200.0, 300.0]}
    a = hostexec_pack(hx, tab, False)
    same_pack(a, Q.pack_quantDT_native(tab, {"degradationBias": False}))
    assert len(a.arena.grp_row_ptr) == 1 and a.unobserved_txid.tolist() == [3, 4, 8]
    assert hostexec_pack(hx, tab, True) is None          # degradation rate 0/0: left to the host packer
    # zero degradation rate (no partial rows at all): the 0.01 rule
    rng = np.random.default_rng(77)
    tab = random_long_table(rng, n_genes=40)
    tab["equal"] = [True] * len(tab["txid"])
    a = hostexec_pack(hx, tab, True)
    b = Q.pack_quantDT_native(tab)
    assert b.d_rate == 0 or np.isnan(b.d_rate)
    if a is not None:
        same_pack(a, b)
    # values the device formulation leaves to the host packer
    tab = random_long_table(rng, n_genes=10)
    tab["nobs"] = [v + 0.5 for v in tab["nobs"]]
    assert hostexec_pack(hx, tab, True) is None
    # the reference's input errors
    bad = {"GENEID": [1, 1], "txid": [1, 2], "eqClassId": [7, 7], "nobs": [3.0, 4.0]}
    with pytest.raises(ValueError, match="more than one nobs"):
        hostexec_pack(hx, bad, False)
    dup = {"GENEID": [1, 1], "txid": [1, 1], "eqClassId": [7, 7], "nobs": [3.0, 3.0]}
    with pytest.raises(ValueError, match="duplicated"):
        hostexec_pack(hx, dup, False)
    # wide id ranges (sort keys of up to 64 bits); txid stays inside R's integer range, which is what the arena carries
    tab = random_long_table(rng, n_genes=30)
    tab["txid"] = [t * 1_003 - 2 ** 30 for t in tab["txid"]]
    tab["eqClassId"] = [e * 7_000_000_007 for e in tab["eqClassId"]]
    tab["GENEID"] = [hash(g) % (2 ** 62) - 2 ** 61 for g in tab["GENEID"]]
    a = hostexec_pack(hx, tab, True)
    b = Q.pack_quantDT_native(tab)
    same_pack(a, b)
    # a txid beyond int32 would be truncated in the arena and merged under a wrong id: both packers refuse it
    wide = dict(tab, txid=[t + 2 ** 33 for t in tab["txid"]])
    with pytest.raises(ValueError, match="32-bit"):
        hostexec_pack(hx, wide, True)
    with pytest.raises(ValueError, match="32-bit"):
        Q.pack_quantDT_native(wide)


# ---------------------------------------------------------------------------- the CUDA executor (B200)
@pytest.fixture(scope="module")
def built():
    from bambu_b200 import _lib
    _lib.build()


@pytest.mark.gpu
@pytest.mark.parametrize("bias", [False, True])
def test_gpu_pack_goldens_and_chr9(built, unit_goldens, bias):
    par = {"degradationBias": bias}
    for case in unit_goldens["cases"]:
        a = Q.pack_quantDT_native(case["input"], par, gpu=True)
        assert a.used_device
        same_pack(a, Q.pack_quantDT_native(case["input"], par))
    from oracle.equirc_ref import gen_equi_rcs
    with open(os.path.join(GOLDEN, "chr9_e2e.json")) as f:
        tab = gen_equi_rcs(json.load(f))
    a = Q.pack_quantDT_native(tab, par, gpu=True)
    assert a.used_device
    same_pack(a, Q.pack_quantDT_native(tab, par))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_gpu_pack_random_long_tables(built, seed):
    rng = np.random.default_rng(5000 + seed)
    tab = random_long_table(rng, n_genes=int(rng.integers(2, 1500)))
    for bias in (False, True):
        a = Q.pack_quantDT_native(tab, {"degradationBias": bias}, gpu=True)
        assert a.used_device
        same_pack(a, Q.pack_quantDT_native(tab, {"degradationBias": bias}))


@pytest.mark.gpu
def test_gpu_pack_falls_back_and_reports_errors(built):
    rng = np.random.default_rng(8)
    tab = random_long_table(rng, n_genes=20)
    frac = dict(tab, nobs=[v + 0.25 for v in tab["nobs"]])
    a = Q.pack_quantDT_native(frac, gpu=True)
    assert not a.used_device                       # non-integer counts: host packer, same answer
    same_pack(a, Q.pack_quantDT_native(frac))
    bad = {"GENEID": [1, 1], "txid": [1, 2], "eqClassId": [7, 7], "nobs": [3.0, 4.0]}
    with pytest.raises(ValueError, match="more than one nobs"):
        Q.pack_quantDT_native(bad, gpu=True)
    dup = {"GENEID": [1, 1], "txid": [1, 1], "eqClassId": [7, 7], "nobs": [3.0, 3.0]}
    with pytest.raises(ValueError, match="duplicated"):
        Q.pack_quantDT_native(dup, gpu=True)
    empty = {"GENEID": np.zeros(0, np.int64), "txid": [], "eqClassId": [], "nobs": []}
    a = Q.pack_quantDT_native(empty, gpu=True)
    assert a.arena.n_groups == 0 and len(a.out_txid) == 0


@pytest.mark.gpu
def test_gpu_pack_config2_sample(built):
    """A 2M-row table of BASELINE config 2 (multi-tile sorts, long (gene, tx) runs): identical to the host packer;
    bambu_quantDT with the arena left in HBM == the host-packed path, double for double."""
    from bambu_b200 import synthetic
    ann = synthetic.annotation_for("config2")
    tab = synthetic.sample_arena(ann, 0, long_table=True)
    a = Q.pack_quantDT_native(tab, gpu=True)
    assert a.used_device
    b = Q.pack_quantDT_native(tab)
    same_pack(a, b)
    assert b.arena.n_groups > 3000 and b.arena.nnz > 1_900_000
    got = Q.bambu_quantDT_native(tab)
    want = Q.bambu_quantDT_native(tab, host_pack=True)
    assert got["d_rate"] == want["d_rate"]
    assert np.array_equal(got["txid"], want["txid"])
    for name in ("counts", "fullLengthCounts", "uniqueCounts"):
        assert_same_doubles(got[name], want[name], name)


@pytest.mark.gpu
def test_gpu_quantDT_goldens(built, unit_goldens):
    """tests/testthat/test_quant.R:12-27 through bambu_quantDT with the device packer."""
    for bias, key in ((False, "estOutput_woBC"), (True, "estOutput_wBC")):
        for case in unit_goldens["cases"]:
            got = Q.bambu_quantDT_native(case["input"], dict(unit_goldens["emParameters"], degradationBias=bias))
            exp = case["expected"][key]
            assert got["txid"].tolist() == [int(t) for t in exp["txid"]]
            for name in ("counts", "fullLengthCounts", "uniqueCounts"):
                assert_same_doubles(got[name], np.array(exp[name]), f"{case['name']} {name}")


@pytest.mark.gpu
def test_dotC_entry_point(built, unit_goldens):
    """bambu_quantDT_dotC (R's .C() calling convention: int / double / logical-as-int vectors, rc by pointer)
    returns the goldens of tests/testthat/test_quant.R:12-27."""
    from bambu_b200 import _lib
    L = _lib.lib()
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)      # noqa: E731
    i32 = lambda v: np.array([v], dtype=np.int32)          # noqa: E731
    f64 = lambda v: np.array([v], dtype=np.float64)        # noqa: E731
    par = unit_goldens["emParameters"]
    for case in unit_goldens["cases"]:
        tab = case["input"]
        n = len(tab["txid"])
        cols = [Q._gene_codes(tab["GENEID"]).astype(np.int32), np.asarray(tab["txid"], np.int32), np.asarray(tab["eqClassId"], np.int32),
                np.asarray(tab["nobs"], np.float64), np.asarray(tab["equal"], bool).astype(np.int32),
                np.asarray(tab["rcWidth"], np.float64), np.asarray(tab["txlen"], np.float64)]
        t_out, c0, c1, c2 = np.zeros(n, np.int32), np.zeros(n), np.zeros(n), np.zeros(n)
        k, d, rc = i32(0), f64(0), i32(-99)
        L.bambu_quantDT_dotC(ptr(i32(n)), *[ptr(c) for c in cols], ptr(i32(1)), ptr(i32(par["maxiter"])),
                             ptr(f64(par["minvalue"])), ptr(f64(par["conv"])), ptr(t_out), ptr(c0), ptr(c1), ptr(c2), ptr(k),
                             ptr(d), ptr(rc))
        assert rc[0] == 0
        exp = case["expected"]["estOutput_wBC"]
        m = int(k[0])
        assert t_out[:m].tolist() == [int(t) for t in exp["txid"]]
        assert_same_doubles(c0[:m], np.array(exp["counts"]), "counts")
        assert_same_doubles(c1[:m], np.array(exp["fullLengthCounts"]), "fullLengthCounts")
        assert_same_doubles(c2[:m], np.array(exp["uniqueCounts"]), "uniqueCounts")


def _table_with_giant_gene(rng, n_small=150, m_giant=2500, n_cls_giant=30000):
    """random small genes + one locus with thousands of transcripts and classes (long (gene, tx) runs in the
    packer, a streamed-tier group in the EM)."""
    tab = random_long_table(rng, n_genes=n_small, shared_tx=False)
    tx0 = max(tab["txid"]) + 1
    cls0 = max(tab["eqClassId"]) + 1
    k = 1 + rng.binomial(m_giant - 1, 12.0 / m_giant, n_cls_giant)
    cls = np.repeat(np.arange(n_cls_giant), k)
    member = np.concatenate([rng.choice(m_giant, size=int(kk), replace=False) for kk in k])
    nobs_c = np.where(rng.random(n_cls_giant) < 0.4, 1.0 + rng.negative_binomial(0.5, 0.5 / 40.5, n_cls_giant), 0.0)
    eq_pick = rng.random(len(cls)) < 0.08
    first = np.ones(len(cls), bool)
    first[1:] = cls[1:] != cls[:-1]
    equal = eq_pick & first                                   # at most one full-length member per class
    w_c = rng.integers(1, 3000, n_cls_giant).astype(float)
    txlen = rng.integers(400, 6000, m_giant).astype(float)
    # every transcript also gets its own single-member class so that it keeps an observation
    u_nobs = 1.0 + rng.negative_binomial(0.5, 0.5 / 40.5, m_giant)
    out = {kname: list(v) for kname, v in tab.items()}
    out["GENEID"] += ["GIANT"] * (len(cls) + m_giant)
    out["txid"] += (tx0 + member).tolist() + (tx0 + np.arange(m_giant)).tolist()
    out["eqClassId"] += (cls0 + cls).tolist() + (cls0 + n_cls_giant + np.arange(m_giant)).tolist()
    out["nobs"] += nobs_c[cls].tolist() + u_nobs.tolist()
    out["equal"] += equal.tolist() + [True] * m_giant
    out["rcWidth"] += w_c[cls].tolist() + txlen.tolist()
    out["txlen"] += txlen[member].tolist() + txlen.tolist()
    perm = rng.permutation(len(out["txid"]))
    return {kname: [v[i] for i in perm] for kname, v in out.items()}


@pytest.mark.gpu
def test_gpu_quantDT_with_giant_gene(built):
    rng = np.random.default_rng(20260003)
    tab = _table_with_giant_gene(rng)
    a = Q.pack_quantDT_native(tab, gpu=True)
    assert a.used_device
    b = Q.pack_quantDT_native(tab)
    same_pack(a, b)
    M, N, _ = b.arena.group_shapes()
    assert M.max() >= 2500 and N.max() >= 30000
    got = Q.bambu_quantDT_native(tab)
    want = Q.bambu_quantDT_native(tab, host_pack=True)
    assert np.array_equal(got["txid"], want["txid"])
    for name in ("counts", "fullLengthCounts", "uniqueCounts"):
        assert_same_doubles(got[name], want[name], name)
    # mass conservation inside the giant locus: every observed read is assigned to one of its transcripts
    # (the small random genes may hold zero-width classes, whose NaN handling the parity above already covers)
    seen, total = set(), 0.0
    giant_tx = set()
    for g, c, n, t in zip(tab["GENEID"], tab["eqClassId"], tab["nobs"], tab["txid"]):
        if g != "GIANT":
            continue
        giant_tx.add(t)
        if c not in seen:
            seen.add(c)
            total += n
    mass = got["counts"][np.isin(got["txid"], np.array(sorted(giant_tx)))].sum()
    assert np.isfinite(mass) and abs(mass - total) <= 1e-6 * total


@pytest.mark.gpu
def test_quantDT_reports_which_packer_ran(built):
    """the host-packer path is a ~25x cliff: every entry point says which packer ran (bambu_quant_last_used_device)"""
    rng = np.random.default_rng(5)
    tab = random_long_table(rng, n_genes=50)
    assert Q.bambu_quantDT_native(tab)["used_device"] is True
    odd = dict(tab, nobs=[v + 0.5 if v else v for v in tab["nobs"]])
    assert Q.bambu_quantDT_native(odd)["used_device"] is False
    assert Q.bambu_quantDT_native(tab, host_pack=True)["used_device"] is False
