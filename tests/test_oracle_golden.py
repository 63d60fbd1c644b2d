"""The oracle (oracle/em_oracle.c + oracle/quant_ref.py) against the reference's
own golden tables: R/sysdata.rda estOutput_woBC / estOutput_wBC, the tables
tests/testthat/test_quant.R:12-27 compares bambu.quantDT with."""
import numpy as np
import pytest

from conftest import assert_same_doubles, random_arena

# EM update counts of the ten golden cases (SURVEY.md section 6 / BASELINE.md section 2)
ITERS_WOBC = [2, 2, 11, 14, 5]
ITERS_WBC = [2, 2, 144, 395, 5]


@pytest.mark.parametrize("bias,key,iters", [(False, "estOutput_woBC", ITERS_WOBC), (True, "estOutput_wBC", ITERS_WBC)])
def test_quant_ref_reproduces_goldens_bit_for_bit(unit_goldens, oracle_mod, bias, key, iters):
    from oracle.quant_ref import quantDT_ref
    par = unit_goldens["emParameters"]
    for case, want_iters in zip(unit_goldens["cases"], iters):
        det = {}
        got = quantDT_ref(case["input"], degradationBias=bias, maxiter=par["maxiter"], conv=par["conv"],
                          minvalue=par["minvalue"], details=det)
        exp = case["expected"][key]
        assert [float(t) for t in got.keys()] == exp["txid"]
        for k, name in enumerate(("counts", "fullLengthCounts", "uniqueCounts")):
            have = np.array([v[k] for v in got.values()])
            assert have.tobytes() == np.array(exp[name], dtype=np.float64).tobytes(), (case["name"], name, have, exp[name])
        assert [g["iters"] for g in det["groups"]] == [want_iters], case["name"]


def test_sparse_oracle_equals_dense_oracle(oracle_mod):
    rng = np.random.default_rng(7)
    a = random_arena(rng, n_groups=60)
    d_est, d_it, d_st = oracle_mod.em_batched(a, mode=oracle_mod.MODE_DENSE)
    t_est, t_it, t_st = oracle_mod.em_batched(a, mode=oracle_mod.MODE_DENSE_TRACE, maxiter=2000)
    s_est, s_it, s_st = oracle_mod.em_batched(a, mode=oracle_mod.MODE_SPARSE)
    assert np.array_equal(d_it, s_it) and np.array_equal(d_st, s_st)
    assert_same_doubles(d_est, s_est, 'dense vs sparse')
    d2_est, d2_it, _ = oracle_mod.em_batched(a, mode=oracle_mod.MODE_DENSE, maxiter=2000)
    assert_same_doubles(t_est, d2_est, 'trace vs plain')
    assert np.array_equal(t_it, d2_it)


def test_sparse_oracle_nan_poisoning_matches_dense(oracle_mod):
    """a kept transcript with rs_i == 0 drives theta_i to
    NaN; the dense code then poisons every column of the group."""
    from bambu_b200.arena import Arena
    A = np.array([[1.0, 0.5, 0.0, 0.0], [0.0, 0.5, 1.0, 0.0], [0.0, 0.0, 0.0, 0.0]])
    eq = np.zeros_like(A, bool)
    eq[2, 3] = True   # structural entry with aval 0 -> rs_2 = 0
    nobs = np.array([5.0, 7.0, 3.0, 2.0])
    a = Arena.from_dense_groups([(A, eq, A.astype(bool).sum(0) > 1, nobs / nobs.sum(), np.full(4, nobs.sum()), None)])
    d_est, d_it, d_st = oracle_mod.em_batched(a, mode=oracle_mod.MODE_DENSE)
    s_est, s_it, s_st = oracle_mod.em_batched(a, mode=oracle_mod.MODE_SPARSE)
    assert d_it[0] == 2 and s_it[0] == 2
    assert np.array_equal(np.isnan(d_est), np.isnan(s_est))
    assert np.array_equal(np.nan_to_num(d_est), np.nan_to_num(s_est))
    assert d_st[0] & 2 and s_st[0] & 2


def test_worked_example_case3(unit_goldens, oracle_mod):
    """SURVEY Appendix B worked example: data3 without bias correction."""
    A = np.array([[1, 0, 0, 1], [1, 1, 1, 1], [1, 0, 0, 0]], dtype=float)
    Y = np.array([500, 10, 0, 50]) / 560
    K = np.full(4, 560.0)
    est, it = oracle_mod.emWithL1(A, np.zeros_like(A), np.zeros_like(A), Y, K)
    assert it == 11
    assert est[0].tolist() == unit_goldens["cases"][2]["expected"]["estOutput_woBC"]["counts"]
