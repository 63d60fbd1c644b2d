"""Host-side mirror of bambu.quantDT (bambu_b200/quantify.py) against the
reference goldens and against the loop-level restatement (oracle/quant_ref.py).
The EM callable is the CPU oracle here: these tests pin the packing / merging
logic, the CUDA path is pinned by the -m gpu tests."""
import numpy as np
import pytest

from bambu_b200 import quantify as Q


def _oracle_em(oracle_mod):
    def em(arena, maxiter, conv, minvalue):
        return oracle_mod.em_batched(arena, maxiter=maxiter, conv=conv, minvalue=minvalue, mode=oracle_mod.MODE_DENSE)
    return em


@pytest.mark.parametrize("bias,key", [(False, "estOutput_woBC"), (True, "estOutput_wBC")])
def test_quantDT_goldens(unit_goldens, oracle_mod, bias, key):
    par = dict(unit_goldens["emParameters"], degradationBias=bias)
    for case in unit_goldens["cases"]:
        got = Q.bambu_quantDT(case["input"], par, em=_oracle_em(oracle_mod))
        exp = case["expected"][key]
        assert got["txid"].tolist() == [int(t) for t in exp["txid"]]
        for name in ("counts", "fullLengthCounts", "uniqueCounts"):
            assert got[name].tobytes() == np.array(exp[name], dtype=np.float64).tobytes(), (case["name"], name)


def random_long_table(rng, n_genes=30, with_unobserved=True, shared_tx=True):
    """Random long table: several genes, classes with 1..k transcripts, some
    unobserved genes / transcripts, optionally a txid listed under two genes."""
    rows = []
    tx_next, cls_next = 1, 1
    for g in range(n_genes):
        m = int(rng.integers(1, 7))
        txs = list(range(tx_next, tx_next + m))
        tx_next += m
        if shared_tx and g > 0 and rng.random() < 0.15:
            txs.append(int(rng.integers(1, tx_next - m)))       # a transcript of an earlier gene
        txlen = {t: float(rng.integers(400, 4000)) for t in txs}
        ncls = int(rng.integers(1, 2 * len(txs) + 2))
        dead = with_unobserved and rng.random() < 0.15
        for _ in range(ncls):
            k = int(rng.integers(1, len(txs) + 1))
            members = rng.choice(txs, size=k, replace=False)
            nobs = 0.0 if (dead or rng.random() < 0.4) else float(1 + rng.negative_binomial(0.5, 0.5 / 40.5))
            eq_member = members[rng.integers(0, k)] if rng.random() < 0.5 else None
            w = float(rng.integers(0, 3000))
            for t in members:
                rows.append((f"G{g:03d}", int(t), cls_next, nobs, bool(t == eq_member), w, txlen[int(t)]))
            cls_next += 1
    perm = rng.permutation(len(rows))
    rows = [rows[i] for i in perm]
    cols = list(zip(*rows))
    return {"GENEID": list(cols[0]), "txid": list(cols[1]), "eqClassId": list(cols[2]), "nobs": list(cols[3]),
            "equal": list(cols[4]), "rcWidth": list(cols[5]), "txlen": list(cols[6])}


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("bias", [False, True])
def test_quantDT_matches_loop_level_restatement(oracle_mod, seed, bias):
    from oracle.quant_ref import quantDT_ref
    rng = np.random.default_rng(100 + seed)
    tab = random_long_table(rng, n_genes=int(rng.integers(3, 120)))
    det_ref, det = {}, {}
    ref = quantDT_ref(tab, degradationBias=bias, details=det_ref)
    got = Q.bambu_quantDT(tab, {"degradationBias": bias}, em=_oracle_em(oracle_mod), details=det)
    assert got["txid"].tolist() == list(ref.keys())
    if bias:
        assert det["d_rate"] == det_ref["d_rate"] or (np.isnan(det["d_rate"]) and np.isnan(det_ref["d_rate"]))
    # same groups, same matrices (integer/index work: exact)
    a = det["arena"]
    assert a.n_groups == len(det_ref["groups"])
    for g, gr in enumerate(det_ref["groups"]):
        A, Af, Au, Y, K = a.dense_group(g)
        assert a.txid[a.grp_row_ptr[g]:a.grp_row_ptr[g + 1]].tolist() == gr["txids"]
        for mine, theirs in ((A, gr["A"]), (Af, gr["A_full"]), (Au, gr["A_unique"]), (Y, gr["Y"]), (K, gr["K"])):
            assert np.array_equal(mine, theirs, equal_nan=True)
        assert int(det["iters"][g]) == gr["iters"]
    for k, name in enumerate(("counts", "fullLengthCounts", "uniqueCounts")):
        want = np.array([v[k] for v in ref.values()])
        assert np.array_equal(got[name], want, equal_nan=True), name


def test_assign_groups_rule():
    from bambu_b200.synthetic import assign_groups
    dims = np.array([2, 600, 2, 500, 998, 4, 1200])
    # stable ascending order: 2,2,4,500,600,998,1200 -> cumsum 2,4,8,508,1108,2106,3306
    assert assign_groups(dims).tolist() == [1, 2, 1, 1, 3, 1, 4]


def test_missing_columns_error():
    with pytest.raises(ValueError, match="missing"):
        Q.pack_quantDT({"txid": [1], "nobs": [1]})
    with pytest.raises(ValueError, match="missing"):
        Q.pack_quantDT(None)


def test_calculateCPM():
    c = np.array([10.0, 30.0, 0.0])
    assert np.allclose(Q.calculateCPM(c, [60.0]), c / 100 * 1e6)


def test_arena_file_round_trip(tmp_path):
    from bambu_b200 import synthetic
    from bambu_b200.arena import Arena
    a = Arena.concatenate([synthetic.config2(scale=0.004), synthetic.config2(seed=5, scale=0.004)])
    p = tmp_path / "cohort.arena.npz"
    a.save(p)
    b = Arena.load(p)
    for name in ("grp_row_ptr", "grp_col_ptr", "row_nnz_ptr", "col_idx", "aval", "nz_flags", "Y", "K", "col_flags", "txid",
                 "sample_grp_ptr"):
        x, y = getattr(a, name), getattr(b, name)
        assert x.dtype == y.dtype and x.tobytes() == y.tobytes(), name
    np.savez(tmp_path / "other.npz", x=np.zeros(3))
    with pytest.raises(ValueError, match="not a bambu-arena-1"):
        Arena.load(tmp_path / "other.npz")
