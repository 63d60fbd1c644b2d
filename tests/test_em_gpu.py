"""Parity of the CUDA path (through the C-ABI) with the CPU oracle.

Bar (BASELINE.json north_star): indexing and iteration counts exact; counts
within 1e-6 relative / 1e-9 absolute.  The resident tier reproduces the
reference's summation order, so these tests ask for more: identical doubles."""
import ctypes

import numpy as np
import pytest

from conftest import assert_same_doubles, random_arena

pytestmark = pytest.mark.gpu

REL, ABS = 1e-6, 1e-9     # the tolerance north_star states


@pytest.fixture(autouse=True, params=["latency", "throughput"])
def kernel_mode(request, monkeypatch):
    """Small arenas run on the latency instance of the EM kernel (256 threads, 4 slots in flight), big ones on the
    throughput instances (64 / 128 threads): every test here runs on both."""
    if request.param == "throughput":
        monkeypatch.setenv("BAMBU_EM_LATENCY_GROUPS", "0")
    return request.param


@pytest.fixture(scope="module")
def em():
    from bambu_b200 import _lib
    import bambu_b200.em as em_mod
    L = _lib.lib()
    assert L.bambu_em_device_count() >= 1, "no sm_100 device: the CUDA path cannot be exercised"
    return em_mod


def check_against_oracle(em, oracle_mod, arena, exact=True, **par):
    est, iters, status = em.abundance_quantification(arena, **{"maxiter": 10000, **par})
    o_est, o_it, o_st = oracle_mod.em_batched(arena, mode=oracle_mod.MODE_SPARSE, **{"maxiter": 10000, **par})
    assert np.array_equal(iters, o_it), f"iteration counts differ in {np.count_nonzero(iters != o_it)} groups"
    assert np.array_equal(status, o_st)
    if exact:
        assert_same_doubles(est, o_est, "est")
    else:
        assert np.array_equal(np.isnan(est), np.isnan(o_est))
        ok = np.isclose(est, o_est, rtol=REL, atol=ABS, equal_nan=True)
        assert ok.all()
    return est, iters, status


@pytest.mark.parametrize("bias,key", [(False, "estOutput_woBC"), (True, "estOutput_wBC")])
def test_reference_goldens_through_cuda(em, unit_goldens, bias, key):
    """tests/testthat/test_quant.R:12-27 with the CUDA library as the EM."""
    from bambu_b200.quantify import bambu_quantDT
    par = dict(unit_goldens["emParameters"], degradationBias=bias)
    want_iters = {"estOutput_woBC": [2, 2, 11, 14, 5], "estOutput_wBC": [2, 2, 144, 395, 5]}[key]
    for case, wi in zip(unit_goldens["cases"], want_iters):
        det = {}
        got = bambu_quantDT(case["input"], par, details=det)
        exp = case["expected"][key]
        assert got["txid"].tolist() == [int(t) for t in exp["txid"]]
        for name in ("counts", "fullLengthCounts", "uniqueCounts"):
            assert_same_doubles(got[name], np.array(exp[name]), f"{case['name']} {name}")
        assert det["iters"].tolist() == [wi]


@pytest.mark.parametrize("seed", range(5))
def test_random_groups_exact(em, oracle_mod, seed):
    rng = np.random.default_rng(1000 + seed)
    a = random_arena(rng, n_groups=150, max_m=60, max_n=120)
    check_against_oracle(em, oracle_mod, a)


def test_larger_groups_exact(em, oracle_mod):
    """groups that land in the bigger shared-memory classes (one CTA of 256/512 threads)"""
    rng = np.random.default_rng(5)
    a = random_arena(rng, n_groups=12, max_m=400, max_n=900, density=0.02, single_tx_frac=0.0)
    b = random_arena(rng, n_groups=4, max_m=150, max_n=250, density=0.25, single_tx_frac=0.0)
    from bambu_b200.arena import Arena
    check_against_oracle(em, oracle_mod, Arena.concatenate([a, b]))


def test_batched_pipeline_many_chunks(em, oracle_mod, monkeypatch):
    """bambu_em_batched streams the arena through three slots; tiny chunks force many
    hand-overs (slot reuse, buffer growth, offsets into the caller's output arrays)."""
    from bambu_b200 import synthetic
    a = synthetic.config2(scale=0.05)
    o_est, o_it, o_st = oracle_mod.em_batched(a, mode=oracle_mod.MODE_SPARSE)
    for mb in ("1", "3", "0.2"):
        monkeypatch.setenv("BAMBU_EM_CHUNK_MB", mb)
        est, iters, status = em.abundance_quantification(a, maxiter=10000)
        assert np.array_equal(iters, o_it) and np.array_equal(status, o_st)
        assert_same_doubles(est, o_est, f"chunk {mb} MB")


def test_config2_small_exact(em, oracle_mod):
    from bambu_b200 import synthetic
    a = synthetic.config2(scale=0.05)
    est, iters, _ = check_against_oracle(em, oracle_mod, a)
    assert iters.max() > 50


def test_config2_full_size(em, oracle_mod):
    """BASELINE.json configs[1] at full size: 60k genes / 250k tx / 500k classes / 2M nnz.
    The sparse oracle finishes this in seconds on the host cores, so compare
    directly, then check the size-independent invariants of the EM."""
    from bambu_b200 import synthetic
    a = synthetic.config2()
    est, iters, status = check_against_oracle(em, oracle_mod, a)
    assert (status == 0).all()
    M, N, _ = a.group_shapes()
    # mass conservation per group: sum_i counts_i = sum_j K_j Y_j (every observed column is explained)
    grp_of_row = np.repeat(np.arange(a.n_groups), M)
    grp_of_col = np.repeat(np.arange(a.n_groups), N)
    got = np.bincount(grp_of_row, weights=est[0], minlength=a.n_groups)
    want = np.bincount(grp_of_col, weights=a.K * a.Y, minlength=a.n_groups)
    assert np.allclose(got, want, rtol=1e-9)
    assert np.all(est[1] <= est[0] * (1 + 1e-12) + 1e-12) and np.all(est[2] <= est[0] * (1 + 1e-12) + 1e-12)
    assert np.all(est >= 0)


def test_config4_cohort_slab(em, oracle_mod):
    """BASELINE.json configs[3] -- the workload bench.py times: a slab of cohort samples (shared annotation, per-sample
    counts, 20 % of the multi-transcript classes re-drawn, per-sample degradation rate) at full sample size.  Whole-table
    equality with the sparse oracle, like tests/testthat/test_quant.R:12-27: identical doubles and iteration counts."""
    from bambu_b200 import synthetic
    from bambu_b200.arena import Arena
    ann = synthetic.annotation_for("config4")
    samples = [synthetic.cohort_sample(ann, s) for s in (0, 1, 7)]
    rates = [x.meta["d_rate"] for x in samples]
    assert len(set(rates)) == 3 and max(rates) - min(rates) > 0.02          # different per-sample d_rate
    a = Arena.concatenate(samples)
    assert a.nnz > 5_500_000 and a.n_rows == 3 * 250_000
    est, iters, status = check_against_oracle(em, oracle_mod, a)
    assert (status == 0).all() and iters.max() > 500
    # sample s of the slab == the sample on its own (groups do not interact, R/bambu.R:187)
    one = em.abundance_quantification(samples[1], maxiter=10000)
    r0, r1 = samples[0].n_rows, samples[0].n_rows + samples[1].n_rows
    assert_same_doubles(est[:, r0:r1], one[0], "sample 1 inside the slab vs alone")


def test_config5_cohort_sample(em, oracle_mod):
    """BASELINE.json configs[4] shape: 75k genes / 400k transcripts / 800k read classes / ~3.2M entries, one sample."""
    from bambu_b200 import synthetic
    ann = synthetic.annotation_for("config5")
    a = synthetic.cohort_sample(ann, 3)
    assert a.n_rows == 400_000 and a.n_cols == 800_000 and a.nnz > 3_000_000
    est, iters, status = check_against_oracle(em, oracle_mod, a)
    assert (status == 0).all()
    M, N, _ = a.group_shapes()
    got = np.bincount(np.repeat(np.arange(a.n_groups), M), weights=est[0], minlength=a.n_groups)
    want = np.bincount(np.repeat(np.arange(a.n_groups), N), weights=a.K * a.Y, minlength=a.n_groups)
    assert np.allclose(got, want, rtol=1e-9)


def test_emWithL1_dense_entry(em, oracle_mod):
    rng = np.random.default_rng(3)
    for M, N in ((3, 4), (1, 5), (7, 20), (30, 64)):
        mask = rng.random((M, N)) < 0.4
        mask[np.arange(M), rng.integers(0, N, M)] = True
        A = np.where(mask, rng.uniform(0.05, 1.0, (M, N)), 0.0)
        eq = mask & (rng.random((M, N)) < 0.3)
        multi = mask.sum(0) > 1
        nobs = rng.integers(0, 30, N).astype(float)
        nobs[0] += 1
        Y, K = nobs / nobs.sum(), np.full(N, nobs.sum())
        got = em.emWithL1(A, A * eq, A * ~multi, Y, K)
        want, it = oracle_mod.emWithL1(A, A * eq, A * ~multi, Y, K)
        assert got["iter"] == it
        assert_same_doubles(got["theta"], want, f"emWithL1 {M}x{N}")


def test_stopping_rule_parameters(em, oracle_mod):
    rng = np.random.default_rng(11)
    a = random_arena(rng, n_groups=40, single_tx_frac=0.0)
    for par in ({"maxiter": 1}, {"maxiter": 2}, {"maxiter": 7}, {"conv": 1.0}, {"conv": 2.0}, {"conv": 1e-6},
                {"conv": 0.0, "maxiter": 300}, {"minvalue": 0.5}, {"minvalue": 1e300}):
        est, iters, status = check_against_oracle(em, oracle_mod, a, **par)
        if par.get("maxiter", 10000) <= 7:
            assert iters.max() <= max(par["maxiter"] - 1, 0)
    _, iters, status = check_against_oracle(em, oracle_mod, a, maxiter=3, conv=1e-12)
    assert iters.max() == 2 and (status[iters < 2] & 1 == 0).all() and (status & 1).any()


def test_nan_poisoning_group(em, oracle_mod):
    from bambu_b200.arena import Arena
    A = np.array([[1.0, 0.5, 0.0, 0.0], [0.0, 0.5, 1.0, 0.0], [0.0, 0.0, 0.0, 0.0]])
    eq = np.zeros_like(A, bool)
    eq[2, 3] = True
    nobs = np.array([5.0, 7.0, 3.0, 2.0])
    a = Arena.from_dense_groups([(A, eq, A.astype(bool).sum(0) > 1, nobs / nobs.sum(), np.full(4, nobs.sum()), None)])
    est, iters, status = check_against_oracle(em, oracle_mod, a)
    assert iters[0] == 2 and status[0] & 2 and np.isnan(est[:, 2]).all()


def test_empty_and_degenerate_inputs(em, oracle_mod):
    from bambu_b200.arena import Arena
    z = np.zeros(1, np.int64)
    empty = Arena(z, z, z, np.zeros(0, np.int32), np.zeros(0), np.zeros(0, np.uint8), np.zeros(0), np.zeros(0),
                  np.zeros(0, np.uint8), np.zeros(0, np.int32))
    est, iters, status = em.abundance_quantification(empty)
    assert est.shape == (3, 0) and len(iters) == 0
    # a group without rows between two ordinary groups; a group whose columns are all unobserved
    rng = np.random.default_rng(2)
    a = random_arena(rng, n_groups=3, single_tx_frac=0.0, zero_frac=0.0, no_empty_cols=True)
    a.grp_row_ptr = np.insert(a.grp_row_ptr, 1, a.grp_row_ptr[1])
    a.grp_col_ptr = np.insert(a.grp_col_ptr, 1, a.grp_col_ptr[1])
    est, iters, status = check_against_oracle(em, oracle_mod, a)
    assert iters[1] == 0 and np.isfinite(est).all()
    a.Y[:] = 0.0
    est, iters, status = check_against_oracle(em, oracle_mod, a)
    assert (est == 0).all() and (iters[[0, 2, 3]] == 1).all()
    # an observed column without a non-zero entry has colsum 0 with Y > 0 -> baseSum = inf ->
    # the dense reference turns every row of the group into NaN (0 * inf)
    b = random_arena(np.random.default_rng(2), n_groups=3, single_tx_frac=0.0)
    est, iters, status = check_against_oracle(em, oracle_mod, b)
    assert np.isnan(est).any() and (status & 2).any()


def test_malformed_arena_is_rejected(em):
    from bambu_b200 import _lib
    rng = np.random.default_rng(4)
    a = random_arena(rng, n_groups=5, single_tx_frac=0.0)
    bad = a.col_idx.copy()
    bad[3] = 10_000
    a.col_idx = bad
    with pytest.raises(_lib.BambuEmError) as e:
        em.abundance_quantification(a)
    assert e.value.code == _lib.EINVAL
    b = random_arena(rng, n_groups=5, single_tx_frac=0.0)
    k = int(np.nonzero(np.diff(b.row_nnz_ptr) >= 2)[0][0])
    s = int(b.row_nnz_ptr[k])
    b.col_idx[s], b.col_idx[s + 1] = b.col_idx[s + 1], b.col_idx[s]
    with pytest.raises(_lib.BambuEmError) as e:
        em.abundance_quantification(b)
    assert e.value.code == _lib.EINVAL


def test_plan_api_resident_runs_and_torch_buffers(em, oracle_mod):
    import torch
    from bambu_b200 import synthetic
    a = synthetic.config2(scale=0.03)
    o_est, o_it, o_st = oracle_mod.em_batched(a, mode=oracle_mod.MODE_SPARSE)
    with em.EmPlan(a) as plan:
        plan.upload()
        for _ in range(2):
            plan.run()
            est, iters, status, nz = plan.download(want_nnz=True)
            assert np.array_equal(iters, o_it)
            assert_same_doubles(est, o_est, "plan est")
            assert np.array_equal(nz, a.group_nonzeros())
        ms = plan.last_run_ms()
        assert ms["total_ms"] > 0 and plan.last_run_launches() >= 2 and plan.device_bytes() > a.nbytes()
        # device buffers owned by torch, kernels ordered on torch's current stream
        dev = torch.device("cuda:0")
        t = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in
             (a.col_idx, a.aval, a.nz_flags, a.Y, a.K, a.col_flags)]
        plan.set_stream(torch.cuda.current_stream().cuda_stream)
        plan.bind_device(*[x.data_ptr() for x in t])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run()
        e1.record()
        torch.cuda.synchronize()
        assert e0.elapsed_time(e1) > 0
        est, iters, status = plan.download()
        assert np.array_equal(iters, o_it)
        assert_same_doubles(est, o_est, "bound est")


# ---------------------------------------------------------------- streamed tier (giant loci)
@pytest.mark.parametrize("seed", range(3))
def test_streamed_tier_forced_on_random_groups(em, oracle_mod, monkeypatch, seed):
    """Every group through the multi-CTA cooperative path (normally reserved for giant
    loci): same doubles, same iteration counts as the oracle, edge cases included."""
    monkeypatch.setenv("BAMBU_EM_FORCE_GIANT", "1")
    rng = np.random.default_rng(2000 + seed)
    a = random_arena(rng, n_groups=25, max_m=70, max_n=150)
    check_against_oracle(em, oracle_mod, a)
    for par in ({"maxiter": 2}, {"conv": 1.0}, {"conv": 1e-5}, {"minvalue": 0.3}):
        check_against_oracle(em, oracle_mod, a, **par)


def test_streamed_tier_single_row_blocks(em, oracle_mod, monkeypatch):
    """the M-phase of the streamed tier works on blocks of 2 rows per warp; loci with more than ~24 000 transcripts
    leave shared memory for 1-row blocks only -- the same code path, forced here"""
    monkeypatch.setenv("BAMBU_EM_FORCE_GIANT", "1")
    monkeypatch.setenv("BAMBU_EM_GIANT_RB", "1")
    rng = np.random.default_rng(2100)
    a = random_arena(rng, n_groups=25, max_m=70, max_n=150)
    check_against_oracle(em, oracle_mod, a)
    check_against_oracle(em, oracle_mod, a, conv=1e-5)


def giant_locus_arena(rng, M, N, k_mean, observed=0.4):
    from bambu_b200.arena import Arena
    k = 1 + rng.binomial(M - 1, k_mean / M, N)
    cols = np.repeat(np.arange(N), k)
    rows = rng.integers(0, M, len(cols))
    key = np.unique(rows.astype(np.int64) * N + cols)
    rows, cols = key // N, key % N
    aval = np.where(rng.random(len(key)) < 0.3, 1.0, rng.uniform(0.02, 0.3, len(key)))
    nobs = np.where(rng.random(N) < observed, 1.0 + rng.negative_binomial(0.5, 0.5 / 40.5, N), 0.0)
    K = nobs.sum()
    multi = np.bincount(cols, minlength=N) > 1
    rnp = np.concatenate(([0], np.cumsum(np.bincount(rows, minlength=M)))).astype(np.int64)
    return Arena(np.array([0, M], np.int64), np.array([0, N], np.int64), rnp, cols.astype(np.int32), aval,
                 (rng.random(len(key)) < 0.2).astype(np.uint8), nobs / K, np.full(N, K), multi.astype(np.uint8),
                 np.arange(M, dtype=np.int32))


def test_giant_locus_matches_oracle(em, oracle_mod):
    """One locus far beyond shared memory (400 transcripts x 30k read classes, ~600k entries)
    next to ordinary groups; maxiter bounds the oracle's run time."""
    from bambu_b200.arena import Arena
    rng = np.random.default_rng(77)
    big = giant_locus_arena(rng, 400, 30_000, 19.0)
    small = random_arena(rng, n_groups=30)
    a = Arena.concatenate([small, big, random_arena(rng, n_groups=10)])
    est, iters, status = check_against_oracle(em, oracle_mod, a, maxiter=400)
    assert iters[30] > 20


def test_config3_scaled(em, oracle_mod):
    """BASELINE.json configs[2] shape, scaled: ordinary genes + giant loci in one arena."""
    from bambu_b200 import synthetic
    ann = synthetic.annotation_for("config3", scale=0.04)      # 200 genes incl. 2 giant loci
    a = synthetic.sample_arena(ann, 0)
    M, N, e = a.group_shapes()
    # the two giant loci have their FULL size (2-4k transcripts x 100-200k read classes, millions of entries each);
    # only the number of loci and of ordinary genes is scaled, and maxiter bounds the oracle's run time
    big = np.argsort(e)[-2:]
    assert (M[big] >= 2000).all() and (N[big] >= 100_000).all() and (e[big] > 1_000_000).all()
    est, iters, status = check_against_oracle(em, oracle_mod, a, maxiter=150)
    assert (iters[big] == 149).all()
    got = np.bincount(np.repeat(np.arange(a.n_groups), M), weights=est[0], minlength=a.n_groups)
    want = np.bincount(np.repeat(np.arange(a.n_groups), N), weights=a.K * a.Y, minlength=a.n_groups)
    assert np.allclose(got, want, rtol=1e-9)


def test_image_pool_grows_and_reruns(em, oracle_mod, monkeypatch):
    """The SELL images are bump-allocated on the device; when the pool is too small the kernels flag
    it, the host doubles the pool and reruns the chunk -- results must not change."""
    from bambu_b200 import synthetic
    a = synthetic.config2(scale=0.03)
    monkeypatch.setenv("BAMBU_EM_POOL_BYTES", "65536")
    check_against_oracle(em, oracle_mod, a)
    with em.EmPlan(a) as plan:
        plan.upload()
        plan.run()
        est, iters, status = plan.download()
    o_est, o_it, _ = oracle_mod.em_batched(a, mode=oracle_mod.MODE_SPARSE)
    assert np.array_equal(iters, o_it)
    assert_same_doubles(est, o_est, "after pool growth")


def test_wide_group_routes_to_streamed_tier(em, oracle_mod):
    """A group whose rows + columns exceed what 16-bit byte offsets can address (M + N > 8190) cannot
    be shared-memory resident even though it is sparse; it must take the streamed tier."""
    from bambu_b200.arena import Arena
    rng = np.random.default_rng(123)
    big = giant_locus_arena(rng, 60, 9000, 1.5)
    a = Arena.concatenate([random_arena(rng, n_groups=5), big])
    check_against_oracle(em, oracle_mod, a, maxiter=500)


def test_giant_groups_in_small_chunks(em, oracle_mod, monkeypatch):
    """streamed-tier groups inside the chunked pipeline (each chunk makes its own cooperative launch)"""
    from bambu_b200.arena import Arena
    rng = np.random.default_rng(321)
    parts = []
    for _ in range(3):
        parts += [random_arena(rng, n_groups=8), giant_locus_arena(rng, 200, 6000, 12.0)]
    a = Arena.concatenate(parts)
    monkeypatch.setenv("BAMBU_EM_CHUNK_MB", "0.5")
    check_against_oracle(em, oracle_mod, a, maxiter=300)


def test_long_shared_columns(em, oracle_mod):
    """Read classes compatible with (almost) every transcript of a locus: columns with thousands of
    entries, in the resident tier (one group of 2000 transcripts) and in the streamed tier."""
    from bambu_b200.arena import Arena
    rng = np.random.default_rng(55)

    def locus(M, n_wide, n_unique):
        N = n_wide + n_unique
        rows = np.concatenate([np.repeat(np.arange(M), n_wide), rng.integers(0, M, n_unique)])
        cols = np.concatenate([np.tile(np.arange(n_wide), M), n_wide + np.arange(n_unique)])
        keep = rng.random(len(rows)) < 0.97
        keep[M * n_wide:] = True
        key = np.unique(rows[keep].astype(np.int64) * N + cols[keep])
        r, c = key // N, key % N
        aval = rng.uniform(0.05, 1.0, len(key))
        nobs = 1.0 + rng.negative_binomial(0.5, 0.5 / 40.5, N)
        K = nobs.sum()
        rnp = np.concatenate(([0], np.cumsum(np.bincount(r, minlength=M)))).astype(np.int64)
        return Arena(np.array([0, M], np.int64), np.array([0, N], np.int64), rnp, c.astype(np.int32), aval,
                     (rng.random(len(key)) < 0.2).astype(np.uint8), nobs / K, np.full(N, K),
                     (np.bincount(c, minlength=N) > 1).astype(np.uint8), np.arange(M, dtype=np.int32))
    a = Arena.concatenate([locus(2000, 2, 2100), locus(3000, 12, 4000), random_arena(rng, n_groups=4)])
    check_against_oracle(em, oracle_mod, a, maxiter=120)


@pytest.mark.gpu
def test_plan_reset_reuses_the_plan(em, oracle_mod):
    """bambu_em_plan_reset: differently shaped arenas through one plan (grow-only device buffers), each identical to
    the oracle; running before the payload is there again is a state error."""
    from bambu_b200.em import EmPlan
    from bambu_b200 import synthetic, _lib
    arenas = [synthetic.config2(scale=0.01), synthetic.config2(seed=11, scale=0.03), synthetic.config2(seed=12, scale=0.005)]
    with EmPlan(arenas[0]) as plan:
        sizes = []
        for k, a in enumerate(arenas):
            if k:
                plan.reset(a)
                with pytest.raises(_lib.BambuEmError) as e:
                    plan.run()
                assert e.value.code == _lib.ESTATE
            plan.upload()
            plan.run(maxiter=10000, minvalue=1e-8, conv=1e-2)
            est, iters, status = plan.download()
            o_est, o_it, o_st = oracle_mod.em_batched(a, maxiter=10000, minvalue=1e-8, conv=1e-2, mode=oracle_mod.MODE_SPARSE)
            assert np.array_equal(iters, o_it) and np.array_equal(status, o_st)
            assert_same_doubles(est, o_est, f"arena {k}")
            sizes.append(plan.device_bytes())
        assert sizes[1] >= sizes[0] and sizes[2] == sizes[1]      # buffers grow, never shrink
