"""Where an end-to-end step goes at N ranks: every rank runs bambu_em_batched_v2 on its slab with BAMBU_EM_TIMING=1
(the library prints the host-side breakdown of every call) under a few variants of the pipeline.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
        tools/e2e_probe.py [--samples 50]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=50)
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--chunks", action="store_true", help="also sweep the chunk size")
    args = ap.parse_args()
    ctx = bench.Ctx(args)
    arena = bench.make_cohort(range(ctx.rank * args.samples, (ctx.rank + 1) * args.samples), args.scale)
    slab = bench.Slab(ctx, arena)
    dev_ms = ctx.reduce([slab.run_device(3, 2) / 3.0])[0]
    if ctx.rank == 0:
        print(f"ranks {ctx.world}, {args.samples} samples per rank, device {dev_ms:.1f} ms per step, "
              f"H2D {slab.h2d / 1e6:.0f} MB, D2H {slab.d2h / 1e6:.0f} MB per rank and step", flush=True)

    def run(label, env, steps=3, solo=False):
        for k, v in env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        slab.e2e_call()
        ctx.barrier()
        os.environ["BAMBU_EM_TIMING"] = "1" if ctx.rank in (0, ctx.world - 1) else ""
        if not os.environ["BAMBU_EM_TIMING"]:
            os.environ.pop("BAMBU_EM_TIMING")
        if ctx.rank == 0:
            print(f"--- {label}", flush=True)
        if solo and ctx.rank != 0:
            ctx.barrier()
            return
        t = time.perf_counter()
        for _ in range(steps):
            slab.e2e_call()
        dt = (time.perf_counter() - t) / steps * 1e3
        os.environ.pop("BAMBU_EM_TIMING", None)
        if solo:
            print(f"    rank 0 alone: {dt:.1f} ms per step", flush=True)
            ctx.barrier()
            return
        mx = ctx.reduce([dt])[0]
        mn = -ctx.reduce([-dt])[0]
        if ctx.rank == 0:
            print(f"    e2e per step: slowest rank {mx:.1f} ms, fastest {mn:.1f} ms", flush=True)

    run("rank 0 alone, default chunks", {"BAMBU_EM_CHUNK_MB": None}, solo=True)
    if args.chunks:
        for frac in ("1", "2", "3", "4", "1000"):
            run(f"all ranks, chunks of at least 1/{frac} of the call", {"BAMBU_EM_CHUNK_MIN_FRACTION": frac})
        os.environ.pop("BAMBU_EM_CHUNK_MIN_FRACTION", None)
    run("all ranks, default chunks (96 MB), tables into torch-pinned memory", {"BAMBU_EM_CHUNK_MB": None})
    if args.chunks:
        run("all ranks, 32 MB chunks", {"BAMBU_EM_CHUNK_MB": "32"})
        run("all ranks, 256 MB chunks", {"BAMBU_EM_CHUNK_MB": "256"})
    if ctx.world > 1:
        # the bench's gather target: one shared, page-locked table in /dev/shm
        gather = bench.HostGather(ctx, slab.rows)
        keep = slab.h_est
        call = slab.e2e_call

        def with_gather():
            slab.h_est = gather.mine
            call()
            gather.gather()
        slab.e2e_call = with_gather
        run("all ranks, tables into the shared host table, published by counters (what bench.py times)", {"BAMBU_EM_CHUNK_MB": None})
        slab.e2e_call = call
        slab.h_est = keep

        def pinned_with_barrier():
            call()
            ctx.barrier()
        slab.e2e_call = pinned_with_barrier
        run("all ranks, torch-pinned tables + barrier per step", {"BAMBU_EM_CHUNK_MB": None})
        slab.e2e_call = call
        gather.close()
    # copies alone, the way the library issues them (plan upload = one cudaMemcpyAsync per array of the whole slab)
    torch = ctx.torch
    for label, fn in (("upload of the whole wire slab (plan API), all ranks", slab.upload),):
        ctx.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = ctx.reduce([(time.perf_counter() - t) * 1e3])[0]
        if ctx.rank == 0:
            print(f"--- {label}: {dt:.1f} ms -> {slab.h2d / dt / 1e6:.1f} GB/s per rank", flush=True)
    ctx.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    slab.plan.download_ptrs(slab.h_est.ctypes.data, slab.h_iters.ctypes.data, slab.h_status.ctypes.data)
    torch.cuda.synchronize()
    dt = ctx.reduce([(time.perf_counter() - t) * 1e3])[0]
    if ctx.rank == 0:
        print(f"--- download of the tables (plan API), all ranks: {dt:.1f} ms -> {slab.d2h / dt / 1e6:.1f} GB/s per rank", flush=True)
    slab.close()
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
