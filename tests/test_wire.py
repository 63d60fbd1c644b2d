"""Wire form v2 of the arena (include/bambu_em.h, bambu_em_batched_v2): the native converter against its numpy
mirror on CPU, and -- on the GPU -- the EM fed with it against the oracle run on the ORIGINAL arena: dropping the
dead columns / zero entries before the copy must not change a bit of the result."""
import numpy as np
import pytest

from conftest import assert_same_doubles, random_arena

FIELDS = ("grp_row_ptr", "grp_col_ptr", "grp_ent_ptr", "row_end", "rs", "aval", "cidx", "eq_bits", "colY", "colK",
          "multi_bits", "nnz_nonzero")


def same_wire(w, v):
    assert w.col_mode == v.col_mode
    for n in FIELDS:
        x, y = getattr(w, n), getattr(v, n)
        assert x.dtype == y.dtype and x.shape == y.shape and x.tobytes() == y.tobytes(), n


def wide_arena(rng, M=40, N=70_000, observed=0.6):
    """one group with more than 32767 live columns: its index block is uint32"""
    from bambu_b200.arena import Arena
    cols = np.arange(N)
    rows = rng.integers(0, M, N)
    extra_c = rng.integers(0, N, N // 2)
    extra_r = rng.integers(0, M, N // 2)
    key = np.unique(np.concatenate((rows, extra_r)).astype(np.int64) * N + np.concatenate((cols, extra_c)))
    r, c = key // N, key % N
    aval = np.where(rng.random(len(key)) < 0.3, 1.0, rng.uniform(0.02, 0.3, len(key)))
    nobs = np.where(rng.random(N) < observed, 1.0 + rng.negative_binomial(0.5, 0.5 / 40.5, N), 0.0)
    K = nobs.sum()
    rnp = np.concatenate(([0], np.cumsum(np.bincount(r, minlength=M)))).astype(np.int64)
    return Arena(np.array([0, M], np.int64), np.array([0, N], np.int64), rnp, c.astype(np.int32), aval,
                 (rng.random(len(key)) < 0.2).astype(np.uint8), nobs / K, np.full(N, K),
                 (np.bincount(c, minlength=N) > 1).astype(np.uint8), np.arange(M, dtype=np.int32))


# ------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("seed", range(4))
def test_native_converter_equals_numpy_mirror(seed):
    from bambu_b200.wire import Wire, wire_from_arena_numpy
    rng = np.random.default_rng(300 + seed)
    a = random_arena(rng, n_groups=60, max_m=50, max_n=100)
    w, v = Wire.from_arena(a), wire_from_arena_numpy(a)
    same_wire(w, v)
    assert w.col_mode == 1                      # integer counts: nobs / K travel as int32
    M, N, e = a.group_shapes()
    assert np.array_equal(np.diff(w.grp_ent_ptr) <= e, np.ones(a.n_groups, bool))
    assert w.nbytes() < a.nbytes()
    # Y = nobs / K is reproduced bit for bit
    live = a.Y != 0
    assert (w.colY.astype(np.float64) / w.colK.astype(np.float64)).tobytes() == a.Y[live].tobytes()


def test_converter_modes_and_edges():
    from bambu_b200.arena import Arena
    from bambu_b200.wire import Wire, wire_from_arena_numpy
    rng = np.random.default_rng(9)
    a = random_arena(rng, n_groups=20)
    a.Y = a.Y * (1 + 1e-3 * rng.random(len(a.Y)))            # no longer integer / K: doubles travel
    w = Wire.from_arena(a)
    same_wire(w, wire_from_arena_numpy(a))
    assert w.col_mode == 0 and w.colY.dtype == np.float64 and w.colY.tobytes() == a.Y[a.Y != 0].tobytes()
    z = np.zeros(1, np.int64)
    empty = Arena(z, z, z, np.zeros(0, np.int32), np.zeros(0), np.zeros(0, np.uint8), np.zeros(0), np.zeros(0),
                  np.zeros(0, np.uint8), np.zeros(0, np.int32))
    w = Wire.from_arena(empty)
    assert w.n_groups == 0 and w.nnz == 0
    b = wide_arena(rng)
    w = Wire.from_arena(b)
    same_wire(w, wire_from_arena_numpy(b))
    assert w.n_cols > 32767 and len(w.cidx) >= 4 * w.nnz           # uint32 index block
    # a group whose columns are all dead keeps its rows (rs) and no entries
    c = random_arena(rng, n_groups=3, single_tx_frac=0.0)
    c.Y[:] = 0.0
    w = Wire.from_arena(c)
    assert w.n_cols == 0 and w.nnz == 0 and w.n_rows == c.n_rows and np.all(w.row_end == 0)


def test_converter_rejects_malformed_arena():
    from bambu_b200 import _lib
    from bambu_b200.wire import Wire
    rng = np.random.default_rng(4)
    a = random_arena(rng, n_groups=5, single_tx_frac=0.0)
    bad = a.col_idx.copy()
    bad[3] = 10_000
    a.col_idx = bad
    with pytest.raises(_lib.BambuEmError) as e:
        Wire.from_arena(a)
    assert e.value.code == _lib.EINVAL
    b = random_arena(rng, n_groups=5, single_tx_frac=0.0)
    b.K = b.K.copy()
    b.K[0] = np.inf
    with pytest.raises(_lib.BambuEmError):
        Wire.from_arena(b)


# ------------------------------------------------------------------------------------ GPU
@pytest.fixture(autouse=True, params=["latency", "throughput"])
def kernel_mode(request, monkeypatch):
    """both instances of the EM kernel (see tests/test_em_gpu.py)"""
    if request.param == "throughput":
        monkeypatch.setenv("BAMBU_EM_LATENCY_GROUPS", "0")
    return request.param


def check_wire_against_oracle(oracle_mod, arena, **par):
    import bambu_b200.em as em
    from bambu_b200.wire import Wire
    w = Wire.from_arena(arena)
    est, iters, status = em.abundance_quantification_wire(w, **{"maxiter": 10000, **par})
    o_est, o_it, o_st = oracle_mod.em_batched(arena, mode=oracle_mod.MODE_SPARSE, **{"maxiter": 10000, **par})
    assert np.array_equal(iters, o_it), f"iteration counts differ in {np.count_nonzero(iters != o_it)} groups"
    assert np.array_equal(status, o_st)
    assert_same_doubles(est, o_est, "est (wire v2)")
    return w, est, iters, status


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_wire_random_groups_exact(oracle_mod, seed):
    rng = np.random.default_rng(4000 + seed)
    a = random_arena(rng, n_groups=150, max_m=60, max_n=120)
    check_wire_against_oracle(oracle_mod, a)
    for par in ({"maxiter": 2}, {"conv": 1e-5}, {"minvalue": 0.3}):
        check_wire_against_oracle(oracle_mod, a, **par)


@pytest.mark.gpu
def test_wire_double_columns_and_nan_cases(oracle_mod):
    from bambu_b200.arena import Arena
    rng = np.random.default_rng(41)
    a = random_arena(rng, n_groups=40)
    a.Y = a.Y * (1 + 1e-3 * rng.random(len(a.Y)))            # col_mode 0
    w, *_ = check_wire_against_oracle(oracle_mod, a)
    assert w.col_mode == 0
    # rs = 0 row (NaN poisoning) and an observed column without a non-zero entry (baseSum = inf)
    A = np.array([[1.0, 0.5, 0.0, 0.0], [0.0, 0.5, 1.0, 0.0], [0.0, 0.0, 0.0, 0.0]])
    eq = np.zeros_like(A, bool)
    eq[2, 3] = True
    nobs = np.array([5.0, 7.0, 3.0, 2.0])
    b = Arena.from_dense_groups([(A, eq, A.astype(bool).sum(0) > 1, nobs / nobs.sum(), np.full(4, nobs.sum()), None)])
    _, est, iters, status = check_wire_against_oracle(oracle_mod, b)
    assert iters[0] == 2 and status[0] & 2 and np.isnan(est[:, 2]).all()
    c = random_arena(np.random.default_rng(2), n_groups=3, single_tx_frac=0.0)
    _, est, _, status = check_wire_against_oracle(oracle_mod, c)
    assert np.isnan(est).any() and (status & 2).any()


@pytest.mark.gpu
def test_wire_config2_chunks_and_plan(oracle_mod, monkeypatch):
    import bambu_b200.em as em
    from bambu_b200 import synthetic
    from bambu_b200.wire import Wire
    a = synthetic.config2(scale=0.05)
    w, est, iters, status = check_wire_against_oracle(oracle_mod, a)
    for mb in ("0.3", "0.05"):                      # many pipeline chunks: slot reuse, bit streams cut inside a word
        monkeypatch.setenv("BAMBU_EM_CHUNK_MB", mb)
        e2, i2, s2 = em.abundance_quantification_wire(w, maxiter=10000)
        assert np.array_equal(i2, iters) and np.array_equal(s2, status)
        assert_same_doubles(e2, est, f"chunk {mb} MB")
    monkeypatch.delenv("BAMBU_EM_CHUNK_MB")
    with em.EmPlan(w) as plan:
        plan.upload()
        for _ in range(2):
            plan.run()
            e3, i3, s3, nz = plan.download(want_nnz=True)
            assert np.array_equal(i3, iters)
            assert_same_doubles(e3, est, "plan v2")
        assert np.array_equal(nz, np.diff(w.grp_ent_ptr))


@pytest.mark.gpu
def test_wire_streamed_tier_and_wide_indices(oracle_mod, monkeypatch):
    from bambu_b200.arena import Arena
    from test_em_gpu import giant_locus_arena
    rng = np.random.default_rng(77)
    big = giant_locus_arena(rng, 400, 30_000, 19.0)
    a = Arena.concatenate([random_arena(rng, n_groups=30), big, random_arena(rng, n_groups=10)])
    check_wire_against_oracle(oracle_mod, a, maxiter=300)
    wide = wide_arena(rng)                           # > 32767 live columns: uint32 index block, streamed tier
    w, *_ = check_wire_against_oracle(oracle_mod, Arena.concatenate([random_arena(rng, n_groups=4), wide]), maxiter=200)
    assert np.diff(w.grp_col_ptr).max() > 32767
    monkeypatch.setenv("BAMBU_EM_FORCE_GIANT", "1")
    check_wire_against_oracle(oracle_mod, random_arena(rng, n_groups=25, max_m=70, max_n=150))


@pytest.mark.gpu
def test_wire_config4_sample_full_size(oracle_mod):
    from bambu_b200 import synthetic
    ann = synthetic.annotation_for("config4")
    a = synthetic.cohort_sample(ann, 2)
    w, est, iters, status = check_wire_against_oracle(oracle_mod, a)
    assert w.nbytes() <= 0.40 * a.nbytes()            # the bytes that cross PCIe: at most 40 % of the v1 arena
    assert (status == 0).all()


@pytest.mark.gpu
def test_malformed_wire_is_rejected():
    import bambu_b200.em as em
    from bambu_b200 import _lib
    from bambu_b200.wire import Wire
    rng = np.random.default_rng(4)
    w = Wire.from_arena(random_arena(rng, n_groups=5, single_tx_frac=0.0, dead_col_frac=0.0))
    idx = w.cidx.view(np.uint16).copy()
    idx[2] = 60_000                                  # live column index out of range
    bad = Wire(**{**w.__dict__, "cidx": idx.view(np.uint8)})
    with pytest.raises(_lib.BambuEmError) as e:
        em.abundance_quantification_wire(bad)
    assert e.value.code == _lib.EINVAL
    re = w.row_end.copy()
    re[0] = re[1] + 5 if len(re) > 1 else 10 ** 6   # row_end not monotone
    bad = Wire(**{**w.__dict__, "row_end": re})
    with pytest.raises(_lib.BambuEmError) as e:
        em.abundance_quantification_wire(bad)
    assert e.value.code == _lib.EINVAL
    # and the library is still usable afterwards
    est, iters, status = em.abundance_quantification_wire(w)
    assert np.isfinite(est).any()


@pytest.mark.gpu
def test_wire_spill_to_streamed_tier_and_pool_growth(oracle_mod, monkeypatch):
    """the rarely taken host paths with the wire form: a group the host admits to the resident tier whose padded image
    does not fit shared memory (flagged by sell_kernel, re-packed by giant_pack.cuh, iterated by the streamed tier), and
    an image pool that is too small (grown, chunk rerun)"""
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
    a = Arena.concatenate([locus(2000, 2, 2100), random_arena(rng, n_groups=6), locus(1500, 3, 1800)])
    check_wire_against_oracle(oracle_mod, a, maxiter=120)
    from bambu_b200 import synthetic
    b = synthetic.config2(scale=0.03)
    monkeypatch.setenv("BAMBU_EM_POOL_BYTES", "65536")
    check_wire_against_oracle(oracle_mod, b)
