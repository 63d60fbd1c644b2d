"""bambu_quantDT_cohort: the quantification stage of a whole cohort in one call, with the epilogue
(modifyQuantOut, removeDuplicates, calculateCPM, match to the annotation, round) on the device.
Checked bit for bit against the per-sample host pipeline: bambu_quantDT (native) + assays.quantify_assays +
assays.combine_assays -- the mirror of R/bambu-quantify.R:22-38 and combineCountSes that is pinned on the
reference's stored assays (tests/test_assays.py, tests/test_chr9_e2e.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_same_doubles
from test_quantify_host import random_long_table

pytestmark = pytest.mark.gpu
NAMES = ("counts", "CPM", "fullLengthCounts", "uniqueCounts")


def host_pipeline(tables, annotation_txid, inc_totals, par=None, sig_digit=5):
    from bambu_b200.assays import combine_assays, quantify_assays
    from bambu_b200.quantify import bambu_quantDT_native
    per = []
    for tab, inc in zip(tables, inc_totals):
        q = bambu_quantDT_native(tab, par)
        per.append(quantify_assays(q, annotation_txid, [inc], sig_digit))
    return combine_assays(per, [str(i) for i in range(len(tables))])[0]


def check(got, want):
    for n in NAMES:
        assert got[n].shape == want[n].shape, n
        assert_same_doubles(got[n], want[n], n)


def test_chr9_duplicated_sample():
    """tests/testthat/test_quant.R:63-66: the same sample twice -> two identical columns equal to the stored assays."""
    from bambu_b200.equirc import genEquiRCs
    from bambu_b200.quantify import bambu_quantDT_cohort_native
    fx = json.load(open(os.path.join(GOLDEN, "chr9_e2e.json")))
    tab = genEquiRCs(fx["distTable"], fx["readClasses"], fx["annotation"])
    inc = [0.0 if v is None else v for v in fx["incompatibleCounts"]["counts"]]
    inc_total = float(np.cumsum(np.asarray(inc, dtype=np.longdouble))[-1])
    got = bambu_quantDT_cohort_native([tab, tab], fx["annotation"]["txid"], [inc_total, inc_total], {"degradationBias": True})
    assert got["used_device"].tolist() == [1, 1] and got["d_rate"].tolist() == [0.08893379057159526] * 2
    for n in NAMES:
        want = np.array(fx["expected_assays"][n], dtype=np.float64)
        assert got[n].shape == (len(want), 2)
        assert_same_doubles(got[n][:, 0], want, n)
        assert_same_doubles(got[n][:, 1], want, n)


@pytest.mark.parametrize("seed", range(3))
def test_random_tables_match_host_pipeline(seed, monkeypatch):
    from bambu_b200.quantify import bambu_quantDT_cohort_native
    rng = np.random.default_rng(900 + seed)
    tables = [random_long_table(rng, n_genes=int(rng.integers(5, 300))) for _ in range(7)]
    all_tx = np.unique(np.concatenate([np.asarray(t["txid"]) for t in tables]))
    # annotation: every txid seen, a few never seen (-> NA rows), shuffled
    ann = rng.permutation(np.concatenate((all_tx, all_tx.max() + 1 + np.arange(5))))
    inc = rng.integers(0, 500, len(tables)).astype(np.float64)
    want = host_pipeline(tables, ann, inc)
    for workers in ("1", "3"):
        monkeypatch.setenv("BAMBU_COHORT_WORKERS", workers)
        got = bambu_quantDT_cohort_native(tables, ann, inc)
        check(got, want)
        assert got["used_device"].all()
    assert np.isnan(want["counts"]).any()                     # the never-seen transcripts
    # no bias correction, other rounding
    want = host_pipeline(tables, ann, inc, {"degradationBias": False}, sig_digit=2)
    check(bambu_quantDT_cohort_native(tables, ann, inc, {"degradationBias": False}, sig_digit=2), want)


def test_host_fallback_sample_inside_a_cohort():
    """a table the device packer does not cover (non-integer counts) goes through the host packer; same assays"""
    from bambu_b200.quantify import bambu_quantDT_cohort_native
    rng = np.random.default_rng(77)
    tables = [random_long_table(rng, n_genes=60) for _ in range(4)]
    tables[2] = dict(tables[2], nobs=[v + 0.5 if v else v for v in tables[2]["nobs"]])
    ann = np.unique(np.concatenate([np.asarray(t["txid"]) for t in tables]))
    inc = np.array([3.0, 0.0, 11.0, 5.0])
    got = bambu_quantDT_cohort_native(tables, ann, inc)
    assert got["used_device"].tolist() == [1, 1, 0, 1]
    check(got, host_pipeline(tables, ann, inc))


def test_config4_slab_from_long_tables():
    """three full-size cohort samples (2 M table rows each) -> assays, against the per-sample host pipeline"""
    from bambu_b200 import synthetic
    from bambu_b200.quantify import bambu_quantDT_cohort_native
    ann4 = synthetic.annotation_for("config4")
    tables = [synthetic.sample_arena(ann4, s, d_rate=None, resample_frac=0.2, long_table=True) for s in (0, 1, 2)]
    ann = np.arange(1, 250_001, dtype=np.int64)
    inc = np.array([1000.0, 0.0, 123456.0])
    got = bambu_quantDT_cohort_native(tables, ann, inc)
    assert got["used_device"].all() and len(set(got["d_rate"].tolist())) == 3
    want = host_pipeline(tables, ann, inc)
    check(got, want)
    assert abs(np.nansum(got["CPM"][:, 1]) - 1e6) < 1.0     # no incompatible reads: CPM sums to a million


def test_cohort_edge_cases():
    """an empty cohort, a sample without rows, a sample whose genes are all unobserved, an annotation without any
    transcript of the tables"""
    from bambu_b200.quantify import bambu_quantDT_cohort_native
    rng = np.random.default_rng(3)
    ann = np.arange(1, 40, dtype=np.int64)
    out = bambu_quantDT_cohort_native([], ann)
    assert out["counts"].shape == (39, 0)
    empty = {"GENEID": [], "txid": [], "eqClassId": [], "nobs": [], "equal": [], "rcWidth": [], "txlen": []}
    dead = {"GENEID": [5, 5, 9], "txid": [3, 4, 8], "eqClassId": [1, 1, 2], "nobs": [0.0, 0.0, 0.0],
            "equal": [True, False, True], "rcWidth": [10.0, 20.0, 30.0], "txlen": [100.0, 200.0, 300.0]}
    live = random_long_table(rng, n_genes=8, shared_tx=False)
    tables = [live, empty, dead]
    inc = np.array([4.0, 0.0, 2.0])
    got = bambu_quantDT_cohort_native(tables, ann, inc, {"degradationBias": False})
    want = host_pipeline([live], ann, inc[:1], {"degradationBias": False})
    for n in NAMES:
        assert_same_doubles(got[n][:, 0], want[n][:, 0], n)
        assert np.isnan(got[n][:, 1]).all()                              # no row at all: every transcript NA
    # unobserved genes: their transcripts are zeros (CPM 0 / 2 * 1e6 = 0), the others NA
    seen = np.isin(ann, [3, 4, 8])
    assert (got["counts"][seen, 2] == 0).all() and (got["CPM"][seen, 2] == 0).all() and np.isnan(got["counts"][~seen, 2]).all()
    # an annotation that shares no txid with the tables
    far = bambu_quantDT_cohort_native([live], np.arange(10_000, 10_020, dtype=np.int64), inc[:1], {"degradationBias": False})
    assert np.isnan(far["counts"]).all()
