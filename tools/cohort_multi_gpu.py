"""bambu_b200.dist.cohort_assays on N ranks (torchrun): a small cohort sharded with shard_samples, every rank runs
bambu_quantDT_cohort on its samples, rank 0 assembles the assays and checks them, double for double, against ONE
bambu_quantDT_cohort call over the whole cohort on its own GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tools/cohort_multi_gpu.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    from bambu_b200 import _lib, dist as bdist, synthetic as S
    from bambu_b200.quantify import bambu_quantDT_cohort_native
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().bambu_em_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scale, n_samples = float(os.environ.get("SCALE", "0.1")), int(os.environ.get("SAMPLES", "10"))
    ann4 = S.annotation_for("config4", scale=scale)
    n_tx = int(ann4.m.sum())
    ann = np.arange(1, n_tx + 1, dtype=np.int64)
    inc = np.arange(n_samples, dtype=np.float64) * 7
    rows_guess = [2_000_000 * scale * (1 + 0.01 * s) for s in range(n_samples)]
    shards = bdist.shard_samples(rows_guess, world)
    mine = [int(s) for s in shards[rank]]
    tables = [S.sample_arena(ann4, s, d_rate=None, resample_frac=0.2, long_table=True) for s in mine]
    dist.barrier()
    t0 = time.perf_counter()
    res = bdist.cohort_assays(tables, mine, n_samples, ann, inc[mine])
    dt = time.perf_counter() - t0
    if rank == 0:
        all_tables = [S.sample_arena(ann4, s, d_rate=None, resample_frac=0.2, long_table=True) for s in range(n_samples)]
        one = bambu_quantDT_cohort_native(all_tables, ann, inc)
        same = all(np.array_equal(res[n], one[n], equal_nan=True) for n in ("counts", "CPM", "fullLengthCounts", "uniqueCounts"))
        print(f"cohort_assays on {world} ranks: {n_samples} samples x {n_tx} transcripts in {dt * 1e3:.1f} ms, shards "
              f"{[len(x) for x in shards]}, identical to one call on one GPU: {same}", flush=True)
        if not same:
            raise SystemExit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
