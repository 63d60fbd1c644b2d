import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def unit_goldens():
    with open(os.path.join(GOLDEN, "quantDT_unit.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def assert_same_doubles(a, b, what=""):
    """Bit-for-bit equality of two float64 arrays, NaNs compared as NaNs (the
    sign/payload of a NaN is not part of the contract)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs"
    ai = np.where(na, 0.0, a).view(np.int64)
    bi = np.where(nb, 0.0, b).view(np.int64)
    bad = np.nonzero(ai != bi)
    assert len(bad[0]) == 0, f"{what}: {len(bad[0])} doubles differ, first at {tuple(x[0] for x in bad)}: " \
                             f"{a[tuple(x[0] for x in bad)]!r} vs {b[tuple(x[0] for x in bad)]!r}"


def random_arena(rng, n_groups=20, max_m=40, max_n=80, density=0.15, zero_frac=0.05, dead_col_frac=0.5,
                 single_tx_frac=0.1, d_like=True, no_empty_cols=False):
    """Random arena with the edge cases the EM must survive: explicit zeros,
    columns with Y == 0, single-transcript groups, rows sharing columns."""
    from bambu_b200.arena import Arena
    groups = []
    for _ in range(n_groups):
        M = 1 if rng.random() < single_tx_frac else int(rng.integers(2, max_m + 1))
        N = int(rng.integers(1, max_n + 1))
        mask = rng.random((M, N)) < density
        mask[np.arange(M), rng.integers(0, N, M)] = True      # every row has an entry
        if no_empty_cols:
            mask[rng.integers(0, M, N), np.arange(N)] = True
        vals = np.where(rng.random((M, N)) < 0.4, 1.0, rng.uniform(0.01, 0.6, (M, N))) if d_like else np.ones((M, N))
        A = np.where(mask, vals, 0.0)
        eq = mask & (rng.random((M, N)) < 0.3)
        explicit_zero = mask & (rng.random((M, N)) < zero_frac)
        A[explicit_zero] = 0.0
        eq |= explicit_zero          # keeps the explicit zero as a structural entry
        nobs = np.where(rng.random(N) < dead_col_frac, 0.0, 1.0 + rng.negative_binomial(0.5, 0.5 / 40.5, N))
        if nobs.sum() == 0:
            nobs[rng.integers(0, N)] = 3.0
        Kval = float(nobs.sum())
        multi = (mask.sum(axis=0) > 1)
        groups.append((A, eq, multi, nobs / Kval, np.full(N, Kval), np.arange(M)))
    return Arena.from_dense_groups(groups)
