"""What the box can deliver: N ranks copy pinned host memory to their GPUs at the same time, nothing else running.

    python tools/h2d_ceiling.py                                             # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/h2d_ceiling.py [--gb 3.66]                                    # every GPU of the box at once

Prints one JSON line (rank 0): aggregate and slowest-rank GB/s for H2D alone, D2H alone and both directions together.
This is the ceiling bench.py's end-to-end numbers are held against (the arena has to cross PCIe once per step)."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=3.66, help="bytes per rank and direction (one 100-sample slab: 3.66 GB)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bind", action="store_true", help="bind each rank to its GPU's CPUs first (NVML affinity)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(args.gb * 1e9) // 8
    h_in = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h_in.fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.float64, device=dev)
    d_out = torch.ones(n, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        best = None
        for _ in range(args.reps + 1):
            barrier()
            t = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            barrier()
            x = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(x, op=dist.ReduceOp.MAX)
            dt = float(x[0])
            best = dt if best is None else min(best, dt)       # first pass is the warm-up, min over the rest
        return best

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    gb = n * 8 / 1e9
    res = {}
    for name, fn, dirs in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both", both, 2)):
        t = timed(fn)
        res[name] = {"seconds_slowest_rank": t, "per_rank_gbs": gb * dirs / t, "aggregate_gbs": gb * dirs * world / t}
    if rank == 0:
        print(json.dumps({"what": "pinned cudaMemcpyAsync, all ranks at once, max over ranks, best of %d" % args.reps,
                          "n_gpus": world, "gb_per_rank_per_direction": gb, "cpus": os.cpu_count(), **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
