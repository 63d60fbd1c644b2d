"""World-size-2 gloo tests of the multi-GPU plumbing (host logic only): sharding is a
partition, shards quantified independently reassemble to the unsharded result, and the
gather returns every rank's table to rank 0.  The EM callable is the CPU oracle here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from bambu_b200 import dist as bdist


def test_lpt_partition_balances():
    rng = np.random.default_rng(0)
    w = rng.pareto(1.5, 500) + 1
    part = bdist.lpt_partition(w, 8)
    loads = np.bincount(part, weights=w, minlength=8)
    assert sorted(np.concatenate([np.nonzero(part == r)[0] for r in range(8)]).tolist()) == list(range(500))
    assert loads.max() <= loads.mean() + w.max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from bambu_b200 import synthetic
    a = synthetic.config2(scale=0.01)
    shards = bdist.shard_groups(a, world)
    mine = a.select_groups(shards[rank])
    est, iters, status = oracle.em_batched(mine, mode=oracle.MODE_SPARSE)
    got = bdist.gather_estimates(est)
    g2 = bdist.EstimateGatherer(est.shape[1], "cpu")
    got2 = g2(np.ascontiguousarray(est))
    # host-side gather: every rank writes its slice of one shared-memory table, rank 0 reads all of them
    tab = bdist.SharedHostTable(est.shape[1])
    for step in range(1, 6):          # double-buffered: a rank may run ahead of rank 0 by at most two steps
        tab.mine[...] = est * step
        got3 = tab.gather()
        if rank == 0:
            assert all(np.array_equal(x * step, y, equal_nan=True) for x, y in zip(got, got3))
        else:
            assert got3 is None
    got3 = None
    tab.close()
    if rank == 0:
        assert all(np.array_equal(a, b.numpy(), equal_nan=True) for a, b in zip(got, got2))
        full_est, full_it, _ = oracle.em_batched(a, mode=oracle.MODE_SPARSE)
        M, _, _ = a.group_shapes()
        ok = True
        for r in range(world):
            rows = np.concatenate([np.arange(a.grp_row_ptr[g], a.grp_row_ptr[g + 1]) for g in shards[r]])
            ok &= np.array_equal(got[r], full_est[:, rows], equal_nan=True)
        q.put(bool(ok) and len(got) == world)
    else:
        q.put(got is None)
    dist.destroy_process_group()


def test_sharded_groups_reassemble_world2():
    world = 2
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(res)


def test_shard_samples_partition():
    sh = bdist.shard_samples([5, 1, 9, 3, 3, 7, 2], 3)
    assert sorted(np.concatenate(sh).tolist()) == list(range(7))


def test_shard_groups_by_weight_places_giants_first():
    w = np.array([1, 2, 3_000_000, 5, 2_000_000, 7, 2_500_000, 1])
    sh = bdist.shard_groups_by_weight(w, 3)
    assert sorted(np.concatenate(sh).tolist()) == list(range(8))
    assert sorted(int(np.argmax(w[x])) >= 0 for x in sh) and len({tuple(x) for x in sh}) == 3
    heavy = [set(x) & {2, 4, 6} for x in sh]
    assert all(len(h) == 1 for h in heavy)           # one giant locus per rank


def _cohort_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total, n_tx = 7, 11
    shards = bdist.shard_samples([5, 1, 9, 3, 3, 7, 2], world)
    mine = [int(s) for s in shards[rank]]

    def fake_quantify(tables, ann, inc, par, sig, out):          # stands in for bambu_quantDT_cohort (needs a GPU)
        for k, t in enumerate(tables):
            for a in range(4):
                out[a, k, :len(ann)] = 1000 * a + 10 * t["sample"] + np.arange(len(ann)) / 100 + inc[k]
    tables = [{"sample": s} for s in mine]
    res = bdist.cohort_assays(tables, mine, n_total, np.arange(1, n_tx + 1), [0.5] * len(mine), quantify=fake_quantify)
    if rank == 0:
        ok = set(res) == {"counts", "CPM", "fullLengthCounts", "uniqueCounts"}
        for a, name in enumerate(("counts", "CPM", "fullLengthCounts", "uniqueCounts")):
            want = np.array([[1000 * a + 10 * s + i / 100 + 0.5 for s in range(n_total)] for i in range(n_tx)])
            ok &= res[name].shape == (n_tx, n_total) and np.array_equal(res[name], want)
        q.put(bool(ok))
    else:
        q.put(res is None)
    dist.destroy_process_group()


def test_cohort_assays_assembled_on_rank0_world2():
    """samples sharded over two ranks, every rank writes its assays into the shared block, rank 0 assembles the
    n_tx x n_samples matrices in cohort order (combineCountSes)"""
    world = 2
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cohort_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(res)
