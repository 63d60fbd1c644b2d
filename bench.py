#!/usr/bin/env python
"""bench.py -- the quantification EM on synthetic read-class sets (BASELINE.json).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--samples S]

A *step* is one pass of the hot path (pack kernels + EM kernels, i.e. the
reference's ``abundance_quantification``) over one batch: the per-GPU slab of a
synthetic cohort -- ``S`` samples of the BASELINE.json config-2 shape (60k genes,
250k transcripts, 500k read classes, ~2M compatibility entries each) drawn as
config 4 prescribes (shared annotation, per-sample counts / d_rate / 20 % of the
multi-transcript classes re-drawn).  Samples are sharded over the ranks with no
data-path collective (weak scaling: every rank owns ``S`` samples).

metric   EM nnz-iterations / s  =  sum_groups nnz_g * iters_g / time, nnz_g counting
         the entries with aval != 0, iters_g the EM updates of the group
value    device time (CUDA events on the launch stream), arena resident in HBM
e2e      host wall clock around the C-ABI call ``bambu_em_batched`` with pinned HOST
         buffers: plan + H2D + kernels + D2H inside, plus the NCCL gather to rank 0 at N > 1
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EM_PAR = dict(maxiter=10000, minvalue=1e-8, conv=1e-2)     # R/bambu_utilityFunctions.R:48-53
ALG_BYTES_PER_NNZ_ITER = 24.0                              # 2 passes x (8 B value + 4 B index)
SEED = 20260004


# ----------------------------------------------------------------------------- data
def _gen(args):
    scale, s = args
    from bambu_b200 import synthetic as S
    ann = S.annotation_for("config4", SEED, scale)
    a = S.cohort_sample(ann, s)
    return (a.grp_row_ptr, a.grp_col_ptr, a.row_nnz_ptr, a.col_idx, a.aval, a.nz_flags, a.Y, a.K, a.col_flags, a.txid)


def make_cohort(sample_ids, scale=1.0):
    from bambu_b200.arena import Arena
    workers = max(1, min(len(sample_ids), (os.cpu_count() or 8), 32))
    jobs = [(scale, int(s)) for s in sample_ids]
    if workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_gen, jobs, chunksize=1)
    else:
        parts = [_gen(j) for j in jobs]
    return Arena.concatenate([Arena(*p) for p in parts])


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference(arena, mode, steps, warmup, threads=0):
    """The reference's CPU path restated (oracle/em_oracle.c): dense column-major
    matrices per group with the sweeps/temporaries/theta_trace of src/em.cpp, groups
    spread over all host cores (the reference itself is serial within a sample)."""
    import oracle
    nz = arena.group_nonzeros()
    best = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        est, iters, status = oracle.em_batched(arena, mode=mode, n_threads=threads, **EM_PAR)
        dt = time.perf_counter() - t
        if i >= warmup:
            best.append(dt)
    work = float((nz * iters).sum())
    cpu_reference.last = (est, iters, status)
    return work, float(np.sum(best)), oracle.num_threads() if threads <= 0 else threads


def parity_block(arena, groups, gpu, cpu):
    """Whole-table comparison (tests/testthat/test_quant.R:12-27 style) of the GPU results for ``groups`` of ``arena``
    with the oracle's on the sub-arena of exactly those groups.  ``gpu`` / ``cpu``: (est[3, rows], iters, status)."""
    M = np.diff(arena.grp_row_ptr)[groups]
    starts = arena.grp_row_ptr[groups]
    rows = np.repeat(starts - np.concatenate(([0], np.cumsum(M)[:-1])), M) + np.arange(int(M.sum()))
    g_est, g_it, g_st = gpu[0][:, rows], gpu[1][groups], gpu[2][groups]
    c_est, c_it, c_st = cpu
    nan_same = bool(np.array_equal(np.isnan(g_est), np.isnan(c_est)))
    a, b = np.nan_to_num(g_est, nan=0.0), np.nan_to_num(c_est, nan=0.0)
    ndiff = int(np.count_nonzero(a.view(np.int64) != b.view(np.int64)))
    den = np.maximum(np.abs(b), 1e-300)
    max_rel = float(np.max(np.abs(a - b) / den)) if a.size else 0.0
    return {"groups": int(len(groups)), "rows": int(len(rows)), "iters_equal": bool(np.array_equal(g_it, c_it)),
            "status_equal": bool(np.array_equal(g_st, c_st)), "doubles_equal": bool(nan_same and ndiff == 0),
            "doubles_differing": ndiff, "max_rel": max_rel,
            "against": "oracle/em_oracle.c sparse mode on the same groups of cohort sample 0 (CPU)"}


def bounded_cpu_sample(scale, frac, seed=0):
    """A random 1/frac of the gene groups of cohort sample 0: the bounded sample
    the CPU arms are timed on (dense B1 needs ~25 s per full sample on 8 cores)."""
    a = make_cohort([0], scale)
    rng = np.random.default_rng(seed)
    k = max(1, int(round(a.n_groups / frac)))
    sel = np.sort(rng.choice(a.n_groups, k, replace=False))
    return a.select_groups(sel), a.n_groups


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    sub, n_all = bounded_cpu_sample(args.scale, args.cpu_frac)
    # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    work, secs, cores = cpu_reference(sub, oracle.MODE_DENSE_TRACE, args.steps, args.warmup, threads=os.cpu_count() or 1)
    value = work * args.steps / secs
    sample = (f"{sub.n_groups} of the {n_all} gene groups of cohort sample 0 (random 1/{args.cpu_frac:g}), "
              f"{int(sub.group_nonzeros().sum())} non-zeros, per step")
    line = {
        "impl": "reference", "metric": "EM nnz-iterations/sec", "value": value, "unit": "nnz-iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "nnz-iterations/s", "cores": cores, "kind": "port",
                         "sample": sample,
                         "note": "dense literal restatement of src/em.cpp (oracle/em_oracle.c, mode dense+theta_trace), "
                                 "gene groups spread over all host cores with OpenMP; the reference's own em.cpp needs "
                                 "RcppArmadillo + R, which this image does not have (no oracle/_ref)"},
        "e2e": {"value": value, "unit": "nnz-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n):
    return {"workload": f"config4 cohort: {args.samples} samples per GPU x (60k genes, 250k transcripts, 500k read "
                        f"classes, ~2M nnz) [BASELINE.json configs[3]; each sample has the configs[1] shape]"
                        + (f" scaled x{args.scale:g}" if args.scale != 1.0 else ""),
            "samples_per_gpu": args.samples, "samples_total": args.samples * n, "parallelism": f"samples sharded over {n} GPU(s)",
            "em": EM_PAR, "l2": "inputs larger than L2 (arena %.0f MB per sample)" % (36.6 * args.scale)}


# ----------------------------------------------------------------------------- B200 arm
def quantdt_stage(args):
    """Latency of the whole per-sample stage (bambu.quantDT, R/bambu-quantify.R:53-74) on ONE config-2 sample given as
    the long table: bambu_quantDT (table H2D, device pack, EM on the arena in HBM, merge by txid) next to the same
    call with the host packer.  An extra object in the JSON line; not part of `value` / `e2e`."""
    from bambu_b200 import quantify as Q, synthetic
    ann = synthetic.annotation_for("config2", scale=args.scale)
    tab = synthetic.sample_arena(ann, 0, long_table=True)

    def best(reps, **kw):
        ts = []
        for _ in range(reps):
            t = time.perf_counter()
            out = Q.bambu_quantDT_native(tab, **kw)
            ts.append(time.perf_counter() - t)
        return min(ts) * 1e3, float(np.median(ts)) * 1e3, out

    Q.bambu_quantDT_native(tab)
    dev_best, dev_med, out = best(5)
    host_best, host_med, out_h = best(2, host_pack=True)
    same = bool(np.array_equal(out["txid"], out_h["txid"]) and out["counts"].tobytes() == out_h["counts"].tobytes())
    return {"rows": int(len(tab["txid"])), "transcripts": int(len(out["txid"])), "ms": dev_med, "ms_best": dev_best,
            "ms_host_packer": host_med, "identical_to_host_packer": same,
            "what": "python call of bambu_quantDT on one config-2 sample (long table in pageable host memory -> "
                    "counts by txid), median of 5; ms_host_packer = the same with csrc/quant_pack.cpp in front"}


def bind_to_gpu_numa_node(local):
    """Keep this rank's host threads -- and with them the first-touched pinned buffers -- on the CPUs next to its GPU
    (NVML's ideal affinity), so that eight ranks do not pull their arenas across the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:          # no NVML / not permitted: run unbound
        return 0


def run_b200(args):
    import torch
    import torch.distributed as dist
    from bambu_b200 import _lib
    from bambu_b200.em import EmPlan, abundance_quantification_into

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    bound_cpus = bind_to_gpu_numa_node(local) if world > 1 else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.lib().bambu_em_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- this rank's slab of the cohort, in pinned host memory ----------------------
    t0 = time.time()
    arena = make_cohort(range(rank * args.samples, (rank + 1) * args.samples), args.scale)
    gen_s = time.time() - t0
    payload = ("col_idx", "aval", "nz_flags", "Y", "K", "col_flags")
    pinned = {}
    for name in payload:
        src = torch.from_numpy(np.ascontiguousarray(getattr(arena, name)))
        pinned[name] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        pinned[name].copy_(src)
        setattr(arena, name, pinned[name].numpy())
    rows, n_groups = arena.n_rows, arena.n_groups
    h_est = torch.empty((3, rows), dtype=torch.float64, pin_memory=True)
    h_iters = torch.empty(max(n_groups, 1), dtype=torch.int32, pin_memory=True)
    h_status = torch.empty(max(n_groups, 1), dtype=torch.int32, pin_memory=True)
    h2d = sum(pinned[n].numel() * pinned[n].element_size() for n in payload) + 8 * (2 * (n_groups + 1) + rows + 1)
    d2h = h_est.numel() * 8 + 4 * 2 * n_groups

    # ---- resident plan -----------------------------------------------------------------
    stream = torch.cuda.Stream()      # a real stream: the library treats the NULL stream as 'use your own'
    torch.cuda.set_stream(stream)
    plan = EmPlan(arena)
    plan.set_stream(stream.cuda_stream)
    plan.upload_ptrs(*[pinned[n].data_ptr() for n in payload])
    plan.run(**EM_PAR)
    est, iters, status, nz = plan.download(want_nnz=True)
    _, nz_live = plan.download_live()
    work = float((nz * iters).sum())
    work_live = float((nz_live.astype(np.int64) * iters).sum())
    launches_per_step = plan.last_run_launches()

    for _ in range(args.warmup):
        plan.run(**EM_PAR)
    plan.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pack_ms = em_ms = 0.0
    e0.record(stream)
    for _ in range(args.steps):
        plan.run(**EM_PAR)
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    last = plan.last_run_ms()
    clocks = sampler.stop() if rank == 0 else None
    tot = torch.tensor([ms, work, work_live], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        ms = float(mx[0])
    work_all, live_all = float(tot[1]), float(tot[2])
    value = work_all * args.steps / (ms * 1e-3)

    # ---- the stages on their own: H2D of the arena / pack + EM kernels / D2H of the outputs, each between
    #      CUDA events on the plan's stream, median of 5 -------------------------------------------------------------
    def staged():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        plan.upload_ptrs(*[pinned[n].data_ptr() for n in payload])
        ev[1].record(stream)
        plan.run(**EM_PAR)
        plan.sync()
        r = plan.last_run_ms()
        ev[2].record(stream)
        plan.download_ptrs(h_est.data_ptr(), h_iters.data_ptr(), h_status.data_ptr())
        ev[3].record(stream)
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]), r["pack_ms"], r["em_ms"], ev[2].elapsed_time(ev[3])
    st = np.median(np.array([staged() for _ in range(5)]), axis=0)
    stages = {"h2d_ms": float(st[0]), "pack_ms": float(st[1]), "em_ms": float(st[2]), "d2h_ms": float(st[3]),
              "sum_ms": float(st.sum()), "note": "one GPU's slab, stages serialised (the pipelined e2e overlaps them)"}

    # ---- end to end through the C-ABI with host buffers ------------------------------------
    gatherer = None
    if world > 1:     # outputs to rank 0 (the reference's combineCountSes cbind, R/bambu_utilityFunctions.R:215-241)
        from bambu_b200.dist import EstimateGatherer
        # overlap: step k's tables travel to rank 0 while step k+1 computes (two host tables alternate)
        gatherer = EstimateGatherer(rows, dev, overlap=True)
    h_est2 = torch.empty((3, rows), dtype=torch.float64, pin_memory=True) if world > 1 else h_est
    flip = [0]

    def e2e_step():
        tab = (h_est, h_est2)[flip[0] & 1]
        flip[0] += 1
        abundance_quantification_into(arena, tab.numpy(), h_iters.numpy(), h_status.numpy(), **EM_PAR)
        if gatherer is not None:
            gatherer(tab)
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    if gatherer is not None:
        gatherer.wait()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    if gatherer is not None:
        gatherer.wait()               # every step's tables are on rank 0's host when the clock stops
    torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = work_all * args.steps / e2e_s
    assert np.array_equal(h_iters.numpy()[:n_groups], iters), "e2e and resident runs disagree"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        step_s = ms * 1e-3 / args.steps
        achieved = ALG_BYTES_PER_NNZ_ITER * (work_all / world) / step_s / 1e9
        line = {
            "metric": "EM nnz-iterations/sec", "value": value, "unit": "nnz-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "samples_per_sec": args.samples * world * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "nnz-iterations/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "samples_per_sec": args.samples * world * args.steps / e2e_s,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "timer": "host wall clock around bambu_em_batched (plan + H2D + kernels + D2H"
                             + (" + NCCL gather to rank 0 and its D2H, overlapped with the next step's EM)" if world > 1 else ")")},
            "stages_ms": stages,
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # DRAM bytes of the dominant kernel (em_sell_kernel<64>, 13 KB class) per launch, from the ncu
                         # --set full capture profiles/r1_v4_em_sell_metrics.txt: 147.9 MB for a 4-sample slab, scaled
                         "traffic": 147.92e6 / 4.0 * args.samples * args.scale,
                         "traffic_note": "dram__bytes_read+write of em_sell_kernel<64> (13 KB class): 37.0 MB per sample "
                                         "(profiles/r1_v4_em_sell_metrics.txt), i.e. the group images are read once; "
                                         "the 24 B x nnz x iters of 'achieved' never leave shared memory",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "kernel": "em_sell_kernel (all size classes; pack + SELL kernels included in the time)",
                         "algorithmic_bytes": "24 B x nnz_g x iters_g (SURVEY 8d streaming-equivalent); the resident "
                                              "tier keeps the slice in shared memory, so HBM traffic is far below this",
                         "achieved_live": ALG_BYTES_PER_NNZ_ITER * (live_all / world) / step_s / 1e9,
                         "live_note": "achieved_live counts only the entries the kernel iterates on "
                                      "(columns with Y != 0, aval != 0)"},
            "work": {"nnz_iterations_per_step": work_all, "live_nnz_iterations_per_step": live_all,
                     "groups_per_gpu": n_groups, "max_iters": int(iters.max()), "median_iters": float(np.median(iters)),
                     "last_step_ms": last, "gen_seconds": gen_s, "cpus_bound_to_gpu_numa_node": bound_cpus},
        }
        if world == 1 and not args.no_cpu:
            import oracle
            oracle.build()
            sel = np.sort(np.random.default_rng(0).choice(
                int(arena.sample_grp_ptr[1]), max(1, int(arena.sample_grp_ptr[1] / args.cpu_frac)), replace=False))
            sub = arena.select_groups(sel)
            w1, s1, cores = cpu_reference(sub, oracle.MODE_DENSE_TRACE, 1, 0, threads=os.cpu_count() or 1)
            w2, s2, _ = cpu_reference(sub, oracle.MODE_SPARSE, 3, 1, threads=os.cpu_count() or 1)
            # the oracle as the checker: the e2e run's tables for these groups must be the oracle's, double for double
            line["parity"] = parity_block(arena, sel, (h_est.numpy(), h_iters.numpy()[:n_groups], h_status.numpy()[:n_groups]),
                                          cpu_reference.last)
            parity_ok = line["parity"]["iters_equal"] and line["parity"]["status_equal"] and line["parity"]["doubles_equal"]
            line["cpu_baseline"] = {
                "value": w1 / s1, "unit": "nnz-iterations/s", "cores": cores, "kind": "port",
                "sample": f"{sub.n_groups} of the {int(arena.sample_grp_ptr[1])} gene groups of cohort sample 0 "
                          f"(random 1/{args.cpu_frac:g}), {int(sub.group_nonzeros().sum())} non-zeros, {s1:.1f} s",
                "note": "B1: dense literal restatement of src/em.cpp incl. theta_trace, groups over all cores (OpenMP)",
                "sparse_value": w2 * 3 / s2,
                "sparse_note": "B2: same arithmetic on the non-zeros only (CSR/CSC), all cores"}
        if world == 1 and not args.no_stage:
            line["quantDT_stage"] = quantdt_stage(args)
        print(json.dumps(line))
        if world == 1 and not args.no_cpu and not parity_ok:
            raise SystemExit("bench: GPU results differ from the oracle on cohort sample 0 (see 'parity')")
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=int(os.environ.get("BAMBU_BENCH_SAMPLES", "100")),
                    help="cohort samples per GPU")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every sample (debugging)")
    ap.add_argument("--cpu-frac", type=float, default=1.0, help="CPU arms run on 1/cpu-frac of one sample's groups")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-stage", action="store_true", help="skip the quantDT_stage latency measurement")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
