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
value    device time (CUDA events on the launch stream), wire form of the arena resident in HBM
e2e      host wall clock around the C-ABI call ``bambu_em_batched_v2`` with pinned HOST
         buffers: plan + H2D + kernels + D2H inside; at N > 1 every rank's tables are on rank 0's
         host (one shared, page-locked table) when the step ends
also     parity (cohort sample 0 of the e2e run vs the sparse oracle; the run fails if it differs),
         roofline.binding / traffic (from profiles/r2_roofline.json, an ncu capture), cpu_baseline,
         strong (the 100-sample cohort over N GPUs via dist.shard_samples), config3 / config3_sharded
         (giant loci, groups over N GPUs via dist.shard_groups_by_weight), config2_single,
         quantDT_stage, quant_stage_cohort (bambu_quantDT_cohort vs the stage on all host cores)
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
        # forkserver: the workers descend from a clean helper process, not from this one (CUDA context, NCCL, pinned memory)
        with mp.get_context("forkserver").Pool(workers) as pool:
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


def _gen_table(args):
    scale, s = args
    from bambu_b200 import synthetic as S
    ann = S.annotation_for("config4", SEED, scale)
    return S.sample_arena(ann, s, d_rate=None, resample_frac=0.2, long_table=True)


def _cpu_stage_one(tab):
    """the per-sample quantification stage on one host core: host packer (csrc/quant_pack.cpp), the sparse oracle EM
    (single thread), merge by txid -- what a BiocParallel worker does per sample, without CPM / rounding (which favours
    the CPU arm)"""
    import oracle
    from bambu_b200 import quantify as Q
    pk = Q.pack_quantDT_native(tab)
    est = oracle.em_batched(pk.arena, mode=oracle.MODE_SPARSE, n_threads=1, **EM_PAR)[0]
    q = Q.modifyQuantOut_removeDuplicates(pk, est)
    out = np.full(int(q["txid"].max()), np.nan)
    out[q["txid"] - 1] = q["uniqueCounts"]          # the unrounded assay: compared with the GPU's, double for double
    return out


def quant_stage_cohort(ctx, args):
    """The whole quantification stage of a cohort from the long tables (what bambu() runs under bplapply, R/bambu.R:187-192,
    after genEquiRCs): ONE call of bambu_quantDT_cohort -> the four n_tx x n_samples assays, tables in page-locked host
    memory.  Beside it: the same stage per sample on the host cores (host packer + sparse oracle EM + host epilogue, one
    sample per core, all cores), on a bounded number of samples."""
    import torch
    from bambu_b200 import quantify as Q
    S_gpu = max(1, min(args.stage_samples, 100))
    with mp.get_context("forkserver").Pool(max(1, min(S_gpu, os.cpu_count() or 8, 32))) as pool:
        tables = pool.map(_gen_table, [(args.scale, s) for s in range(S_gpu)], chunksize=1)
    n_tx = int(max(int(np.max(t["txid"])) for t in tables))
    ann = np.arange(1, n_tx + 1, dtype=np.int64)
    rows = int(sum(len(t["txid"]) for t in tables))
    pinned = []
    for t in tables:                      # page-locked copies of the table columns (what a caller would register once)
        p = {}
        for k, dt in (("GENEID", np.int64), ("txid", np.int64), ("eqClassId", np.int64), ("nobs", np.float64),
                      ("equal", np.uint8), ("rcWidth", np.float64), ("txlen", np.float64)):
            src = torch.from_numpy(np.ascontiguousarray(np.asarray(t[k]).astype(dt)))
            dst = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            dst.copy_(src)
            p[k] = dst.numpy()
        p["equal"] = p["equal"].view(np.bool_)
        pinned.append(p)
    inc = np.zeros(S_gpu)
    Q.bambu_quantDT_cohort_native(pinned[:min(4, S_gpu)], ann, inc[:min(4, S_gpu)])           # warm-up: workspaces, plans
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = Q.bambu_quantDT_cohort_native(pinned, ann, inc)
        ts.append(time.perf_counter() - t0)
    gpu_s = min(ts)
    res = {"samples": S_gpu, "table_rows": rows, "transcripts": n_tx, "ms": gpu_s * 1e3, "ms_per_sample": gpu_s * 1e3 / S_gpu,
           "samples_per_sec": S_gpu / gpu_s, "h2d_bytes": rows * 49, "used_device": int(out["used_device"].sum()),
           "what": "bambu_quantDT_cohort: long tables (page-locked host memory) -> GPU pack -> EM -> device epilogue -> "
                   "4 assays n_tx x n_samples in host memory; best of 3"}
    if not args.no_cpu:
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        n_cpu = max(1, min(cores, S_gpu))
        t0 = time.perf_counter()
        with mp.get_context("forkserver").Pool(n_cpu) as pool:
            cpu_counts = pool.map(_cpu_stage_one, tables[:n_cpu], chunksize=1)
        cpu_s = time.perf_counter() - t0
        same = all(np.array_equal(c, out["uniqueCounts"][:len(c), k], equal_nan=True) for k, c in enumerate(cpu_counts))
        res["cpu"] = {"samples": n_cpu, "cores": cores, "seconds": cpu_s, "samples_per_sec": n_cpu / cpu_s,
                      "uniqueCounts_identical_to_gpu": bool(same),
                      "what": "per sample on one core: host packer + sparse oracle EM + host merge (no CPM / rounding); "
                              "one sample per core, all cores at once (pool start-up included)"}
        res["speedup_vs_cpu"] = res["samples_per_sec"] / res["cpu"]["samples_per_sec"]
    return res


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


class Ctx:
    """torch / torch.distributed plumbing of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        self.bound_cpus = bind_to_gpu_numa_node(self.local) if self.world > 1 else 0
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        from bambu_b200 import _lib
        _lib.check(_lib.lib().bambu_em_set_device(self.local))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream()      # a real stream: the library treats the NULL stream as 'use your own'
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, values, op="max"):
        """max / sum of a few doubles over the ranks (device tensors: NCCL)"""
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def bcast(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]


def make_table(ctx, rows):
    """the host-side gather target (dist.SharedHostTable); None when the box has no room in /dev/shm"""
    from bambu_b200 import dist as bdist
    try:
        return bdist.SharedHostTable(rows)
    except RuntimeError:
        return None


class HostGather:
    """One interface over the two gathers: the shared host table (a barrier) or, without shared memory, NCCL."""

    def __init__(self, ctx, rows):
        self.ctx, self.table, self.nccl = ctx, make_table(ctx, rows), None
        if self.table is None:
            from bambu_b200 import dist as bdist
            self.nccl = bdist.EstimateGatherer(rows, ctx.dev)
            self.buf = ctx.torch.empty((3, rows), dtype=ctx.torch.float64, pin_memory=True)
        self.kind = ("host side: every rank's D2H lands in one shared, page-locked table (dist.SharedHostTable)"
                     if self.table is not None else "NCCL gather to GPU 0 + D2H (no shared memory available)")

    @property
    def mine(self):
        """host buffer for this rank's tables of its NEXT step (the shared table is double-buffered)"""
        return self.table.mine if self.table is not None else self.buf.numpy()

    def gather(self):
        if self.table is not None:
            self.table.gather()
        else:
            self.nccl(self.buf)

    def close(self):
        if self.table is not None:
            self.table.close()


class Slab:
    """One rank's share of a workload: the arena's wire form v2 in pinned host memory, a resident plan, host output
    buffers.  `run_device` times plan runs with CUDA events on the launch stream; `e2e_call` is the C-ABI call a user
    makes (bambu_em_batched_v2: H2D of the wire form, pack + EM kernels, D2H of the tables) with HOST buffers."""

    def __init__(self, ctx, arena, est_out=None):
        import torch
        from bambu_b200.em import EmPlan
        from bambu_b200.wire import Wire, PAYLOAD
        self.ctx, self.arena = ctx, arena
        t0 = time.time()
        self.wire = w = Wire.from_arena(arena)
        self.wire_s = time.time() - t0
        self.pinned = {}
        for name in PAYLOAD:                                   # the bytes that cross PCIe every e2e step
            src = torch.from_numpy(np.ascontiguousarray(getattr(w, name)))
            self.pinned[name] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            self.pinned[name].copy_(src)
            setattr(w, name, self.pinned[name].numpy())
        self.rows, self.n_groups = w.n_rows, w.n_groups
        self.h_est = est_out if est_out is not None else torch.empty((3, self.rows), dtype=torch.float64, pin_memory=True).numpy()
        self.h_iters = torch.empty(max(self.n_groups, 1), dtype=torch.int32, pin_memory=True).numpy()
        self.h_status = torch.empty(max(self.n_groups, 1), dtype=torch.int32, pin_memory=True).numpy()
        self.h2d = int(w.nbytes())
        self.d2h = int(3 * self.rows * 8 + 4 * 2 * self.n_groups)
        self.plan = EmPlan(w)
        self.plan.set_stream(ctx.stream.cuda_stream)
        self.upload()
        self.plan.run(**EM_PAR)
        self.est, self.iters, self.status = self.plan.download()
        _, nz_live = self.plan.download_live()
        _, slots = self.plan.download_image_stats()
        # SELL fill of the resident images: live entries over the 2 x 32 places of every slot, both sides together
        self.sell_fill = float(2.0 * nz_live.sum() / max(64.0 * slots.sum(), 1.0))
        self.work = float((w.nnz_nonzero * self.iters).sum())              # nnz_g (aval != 0) x iters_g
        self.work_live = float((nz_live.astype(np.int64) * self.iters).sum())
        self.launches = self.plan.last_run_launches()

    def upload(self):
        p, w = self.pinned, self.wire
        self.plan.upload_ptrs_v2(p["row_end"].data_ptr(), p["rs"].data_ptr(), p["aval"].data_ptr(), p["cidx"].data_ptr(),
                                 p["eq_bits"].data_ptr(), w.col_mode, p["colY"].data_ptr(), p["colK"].data_ptr(),
                                 p["multi_bits"].data_ptr())

    def run_device(self, steps, warmup):
        """ms for `steps` resident runs (CUDA events on the launch stream), this rank only"""
        torch, st = self.ctx.torch, self.ctx.stream
        for _ in range(warmup):
            self.plan.run(**EM_PAR)
        self.plan.sync()
        self.ctx.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            self.plan.run(**EM_PAR)
        e1.record(st)
        torch.cuda.synchronize()
        self.ctx.barrier()
        return e0.elapsed_time(e1)

    def e2e_call(self):
        from bambu_b200.em import abundance_quantification_wire
        abundance_quantification_wire(self.wire, out=(self.h_est, self.h_iters, self.h_status), **EM_PAR)

    def close(self):
        self.plan.close()


def timed_e2e(ctx, step, steps, warmup, finish=None, solo=False):
    """host wall clock around `steps` calls of `step` (barrier + synchronize on both sides, max over ranks)"""
    torch = ctx.torch
    for _ in range(warmup):
        step()
    if finish:
        finish()
    if not solo:
        ctx.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    if finish:
        finish()
    torch.cuda.synchronize()
    if not solo:
        ctx.barrier()
    dt = time.perf_counter() - t0
    return dt if solo else ctx.reduce([dt])[0]


def cohort_strong(ctx, args, slab0, t1_dev_ms, t1_e2e_ms):
    """BASELINE.json configs[3] as it is worded: ONE 100-sample cohort sharded over the N GPUs (strong scaling).
    Samples -> ranks with the product's partitioner (dist.shard_samples: LPT on nnz x iterations per sample, the
    iterations taken from the one-GPU pass rank 0 has just made over the whole cohort)."""
    from bambu_b200 import dist as bdist
    S = args.samples
    w = None
    if ctx.rank == 0:
        sgp = slab0.arena.sample_grp_ptr
        per_group = slab0.wire.nnz_nonzero * slab0.iters
        w = [float(per_group[sgp[s]:sgp[s + 1]].sum()) for s in range(S)]
    w = ctx.bcast(w)
    shards = bdist.shard_samples(w, ctx.world)
    mine = [int(s) for s in shards[ctx.rank]]
    if ctx.rank == 0:
        sgp = slab0.arena.sample_grp_ptr
        arena = slab0.arena.select_groups(np.concatenate([np.arange(sgp[s], sgp[s + 1]) for s in mine]))
    else:
        arena = make_cohort(mine, args.scale)
    table = HostGather(ctx, arena.n_rows)
    slab = Slab(ctx, arena, est_out=table.mine)
    work = ctx.reduce([slab.work], "sum")[0]
    dev_ms = ctx.reduce([slab.run_device(3, 1) / 3.0])[0]

    def step():
        slab.h_est = table.mine
        slab.e2e_call()
        table.gather()
    e2e_ms = timed_e2e(ctx, step, 3, 2) / 3.0 * 1e3
    loads = ctx.reduce([sum(w[s] for s in mine)])[0] / (sum(w) / ctx.world)
    out = {"samples_total": S, "samples_per_rank": [len(x) for x in shards], "partitioner": "dist.shard_samples (LPT on nnz x iters per sample)",
           "max_load_over_mean": loads, "ms": dev_ms, "e2e_ms": e2e_ms,
           "one_gpu_ms": t1_dev_ms, "one_gpu_e2e_ms": t1_e2e_ms,
           "efficiency": t1_dev_ms / (ctx.world * dev_ms), "e2e_efficiency": t1_e2e_ms / (ctx.world * e2e_ms),
           "value": work / (dev_ms * 1e-3), "e2e_value": work / (e2e_ms * 1e-3),
           "gather": table.kind,
           "note": "one_gpu_*: the same 100 samples on rank 0's GPU alone, measured in this run"}
    slab.close()
    table.close()
    return out


def config3_block(ctx, args):
    """BASELINE.json configs[2] (5k genes incl. 50 giant loci), full size.  N = 1: the whole sample on one GPU.
    N > 1: its gene groups sharded by dist.shard_groups_by_weight (LPT on the a-priori entry counts, giant loci
    first); every rank draws only its own loci; rank 0 also times the whole sample alone."""
    from bambu_b200 import dist as bdist, synthetic as S
    sc = args.config3_scale
    plan3 = S.config3_plan(sc)
    small = S.sample_arena(plan3.small, 0)
    _, _, e_small = small.group_shapes()
    weights = np.concatenate((e_small.astype(np.float64), plan3.giant_expected_entries()))
    n_small = small.n_groups
    workers = max(1, min(16, (os.cpu_count() or 8) // max(1, ctx.world)))

    def measure(arena, table=None, solo=False):
        slab = Slab(ctx, arena, est_out=None if table is None else table.mine)
        if solo:
            ms = []
            for _ in range(3):
                slab.plan.run(**EM_PAR)
                slab.plan.sync()
                ms.append(slab.plan.last_run_ms())
            dev, pack = min(m["total_ms"] for m in ms), min(m["pack_ms"] for m in ms)
            e2e = timed_e2e(ctx, slab.e2e_call, 2, 1, solo=True) / 2.0 * 1e3
        else:
            dev = ctx.reduce([slab.run_device(2, 1) / 2.0])[0]
            pack = ctx.reduce([slab.plan.last_run_ms()["pack_ms"]])[0]

            def step():
                slab.h_est = table.mine
                slab.e2e_call()
                table.gather()
            e2e = timed_e2e(ctx, step, 2, 1) / 2.0 * 1e3
        res = (slab.work, dev, pack, e2e, int(slab.wire.nbytes()), int(arena.nnz), int(slab.iters.max()) if slab.n_groups else 0)
        slab.close()
        return res

    out = {"scale": sc, "giant_loci": int(plan3.n_giant), "groups": int(n_small + plan3.n_giant)}
    arena = None
    if ctx.world > 1:      # every rank draws its own shard (in parallel), then rank 0 draws and times the whole sample alone
        from bambu_b200.arena import Arena
        shards = bdist.shard_groups_by_weight(weights, ctx.world)
        mine = shards[ctx.rank]
        loci = [int(g - n_small) for g in mine if g >= n_small]
        parts = [small.select_groups(np.array([g for g in mine if g < n_small], dtype=np.int64))]
        if loci:
            parts.append(S.config3_arena(sc, loci=loci, small=False, workers=workers, mp_context="forkserver"))
        arena = Arena.concatenate(parts)
        ctx.barrier()
    one = None
    if ctx.rank == 0:
        t0 = time.time()
        full = S.config3_arena(sc, workers=min(16, os.cpu_count() or 8), mp_context="forkserver")
        gen_s = time.time() - t0
        work, dev, pack, e2e, wire_b, entries, mx = measure(full, solo=True)
        one = {"entries": entries, "nnz_iterations": work, "max_iters": mx, "ms": dev, "pack_ms": pack, "e2e_ms": e2e,
               "value": work / (dev * 1e-3), "e2e_value": work / (e2e * 1e-3), "h2d_bytes": wire_b, "gen_seconds": gen_s}
        del full
    if ctx.world == 1:
        out.update(one)
        return out
    ctx.barrier()
    table = HostGather(ctx, arena.n_rows)
    work, dev, pack, e2e, wire_b, entries, mx = measure(arena, table=table)
    work_all, ent_all = ctx.reduce([work, entries], "sum")
    work_max = ctx.reduce([work])[0]
    table.close()
    one = ctx.bcast(one)
    out.update({"partitioner": "dist.shard_groups_by_weight (LPT on a-priori entries per group, giant loci first)",
                "loci_per_rank": [int(sum(1 for g in x if g >= n_small)) for x in shards],
                "entries": ent_all, "nnz_iterations": work_all, "max_work_over_mean": work_max / (work_all / ctx.world),
                "ms": dev, "pack_ms": pack, "e2e_ms": e2e, "value": work_all / (dev * 1e-3), "e2e_value": work_all / (e2e * 1e-3),
                "one_gpu": one, "efficiency": one["ms"] / (ctx.world * dev), "e2e_efficiency": one["e2e_ms"] / (ctx.world * e2e),
                "note": "the sharded sample is drawn locus by locus from per-locus seeds, so it is the same sample rank 0 "
                        "quantifies alone; work differs only through the iteration counts"})
    return out


def config2_single(ctx, args):
    """BASELINE.json configs[1]: one sample on one GPU -- the latency case (the longest group bounds it)."""
    from bambu_b200 import synthetic as S
    a = S.config2(scale=args.scale)
    slab = Slab(ctx, a)
    ms = []
    for _ in range(5):
        slab.plan.run(**EM_PAR)
        slab.plan.sync()
        ms.append(slab.plan.last_run_ms()["total_ms"])
    e2e = timed_e2e(ctx, slab.e2e_call, 5, 2, solo=True) / 5.0 * 1e3
    out = {"ms": float(np.median(ms)), "ms_best": min(ms), "e2e_ms": e2e, "nnz_iterations": slab.work,
           "value": slab.work / (min(ms) * 1e-3), "max_iters": int(slab.iters.max()), "groups": slab.n_groups}
    slab.close()
    return out


def roofline_capture():
    """profiles/r2_roofline.json: ncu counters of the dominant kernel on the bench slab, written by
    tools/roofline_capture.py (the capture itself runs under ncu, never inside a timed bench run)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_roofline.json")))
    except (OSError, ValueError):
        return None


def run_b200(args):
    ctx = Ctx(args)
    torch, world, rank = ctx.torch, ctx.world, ctx.rank
    from bambu_b200 import dist as bdist

    # ---- this rank's slab of the cohort: wire form v2 in pinned host memory, resident plan -----------------
    t0 = time.time()
    arena = make_cohort(range(rank * args.samples, (rank + 1) * args.samples), args.scale)
    gen_s = time.time() - t0
    table = HostGather(ctx, arena.n_rows) if world > 1 else None
    slab = Slab(ctx, arena, est_out=None if table is None else table.mine)
    v1_bytes = int(arena.nbytes())

    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    my_ms = slab.run_device(args.steps, args.warmup)
    last = slab.plan.last_run_ms()
    clocks = sampler.stop() if rank == 0 else None
    ms = ctx.reduce([my_ms])[0]
    work_all, live_all = ctx.reduce([slab.work, slab.work_live], "sum")
    value = work_all * args.steps / (ms * 1e-3)

    # ---- the stages on their own: H2D of the wire form / pack + EM kernels / D2H of the outputs, each between
    #      CUDA events on the plan's stream, median of 5 -------------------------------------------------------------
    def staged():
        st = ctx.stream
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(st)
        slab.upload()
        ev[1].record(st)
        slab.plan.run(**EM_PAR)
        slab.plan.sync()
        r = slab.plan.last_run_ms()
        ev[2].record(st)
        slab.plan.download_ptrs(slab.h_est.ctypes.data, slab.h_iters.ctypes.data, slab.h_status.ctypes.data)
        ev[3].record(st)
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]), r["pack_ms"], r["em_ms"], ev[2].elapsed_time(ev[3])
    st = np.median(np.array([staged() for _ in range(5)]), axis=0)
    stages = {"h2d_ms": float(st[0]), "pack_ms": float(st[1]), "em_ms": float(st[2]), "d2h_ms": float(st[3]),
              "sum_ms": float(st.sum()), "note": "one GPU's slab, stages serialised (the pipelined e2e overlaps them)"}

    # ---- end to end through the C-ABI with host buffers ------------------------------------
    # every step: bambu_em_batched_v2 (H2D of the wire form, kernels, D2H of the tables); at N > 1 the tables of all
    # ranks are on rank 0's host when the step ends (the reference's combineCountSes cbind): each rank's D2H lands in
    # a shared page-locked table, the gather itself is a barrier
    def e2e_step():
        if table is not None:
            slab.h_est = table.mine
        slab.e2e_call()
        if table is not None:
            table.gather()
    e2e_s = timed_e2e(ctx, e2e_step, args.steps, min(args.warmup, 2))
    e2e_value = work_all * args.steps / e2e_s
    assert np.array_equal(slab.h_iters[:slab.n_groups], slab.iters), "e2e and resident runs disagree"

    # the NCCL alternative of the gather (tables -> GPU 0 -> rank 0's host), for comparison
    nccl_e2e_ms = None
    if world > 1:
        h_a = torch.empty((3, slab.rows), dtype=torch.float64, pin_memory=True)
        h_b = torch.empty((3, slab.rows), dtype=torch.float64, pin_memory=True)
        gatherer = bdist.EstimateGatherer(slab.rows, ctx.dev, overlap=True)
        keep, flip = slab.h_est, [0]

        def nccl_step():
            tab = (h_a, h_b)[flip[0] & 1]
            flip[0] += 1
            slab.h_est = tab.numpy()
            slab.e2e_call()
            gatherer(tab)
        nccl_e2e_ms = timed_e2e(ctx, nccl_step, 3, 1, finish=gatherer.wait) / 3.0 * 1e3
        slab.h_est = keep
        del gatherer, h_a, h_b

    # rank 0's slab alone (the other GPUs idle): the one-GPU end-to-end time of the same 100 samples
    solo_e2e_ms = e2e_s / args.steps * 1e3
    if world > 1:
        ctx.barrier()
        if rank == 0:
            solo_e2e_ms = timed_e2e(ctx, slab.e2e_call, 3, 1, solo=True) / 3.0 * 1e3
        ctx.barrier()
        solo_e2e_ms = ctx.bcast(solo_e2e_ms)
    t1_dev_ms = ctx.bcast(my_ms / args.steps)

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        step_s = ms * 1e-3 / args.steps
        achieved = ALG_BYTES_PER_NNZ_ITER * (work_all / world) / step_s / 1e9
        cap = roofline_capture()
        h2d, d2h = slab.h2d, slab.d2h
        line = {
            "metric": "EM nnz-iterations/sec", "value": value, "unit": "nnz-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "samples_per_sec": args.samples * world * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "nnz-iterations/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "samples_per_sec": args.samples * world * args.steps / e2e_s,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "api": "bambu_em_batched_v2 (wire form v2 of the arena: live entries + rs, %.0f %% of the v1 arena's "
                           "%d bytes per rank)" % (100.0 * h2d / v1_bytes, v1_bytes),
                    "timer": "host wall clock around the C-ABI call (plan + H2D + kernels + D2H"
                             + ("; every rank's D2H lands in its slice of one shared, double-buffered host table and is "
                                "published by a counter; rank 0 waits for all counters = the gather to rank 0)" if world > 1 else ")"),
                    "one_gpu_ms_per_step": solo_e2e_ms, "nccl_gather_ms_per_step": nccl_e2e_ms},
            "stages_ms": stages,
            "gpu_launches": int(slab.launches * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ((cap or {}).get("dram_bytes_per_sample") or 0) * args.samples * args.scale or None,
                         "binding": (cap or {}).get("binding"),
                         "capture": (cap or {}).get("capture"),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "kernel": "em_sell_kernel (all size classes; pack + SELL kernels included in the time)",
                         "algorithmic_bytes": "24 B x nnz_g x iters_g (SURVEY 8d streaming-equivalent); the resident "
                                              "tier keeps the slice in shared memory, so HBM traffic is far below this and "
                                              "`frac` can exceed 1: `binding` holds the measured fractions of the resources "
                                              "that do bind (issue slots, shared-memory pipe, FP64 pipe, DRAM)",
                         "achieved_live": ALG_BYTES_PER_NNZ_ITER * (live_all / world) / step_s / 1e9,
                         "live_note": "achieved_live counts only the entries the kernel iterates on "
                                      "(columns with Y != 0, aval != 0)"},
            "work": {"nnz_iterations_per_step": work_all, "live_nnz_iterations_per_step": live_all,
                     "groups_per_gpu": slab.n_groups, "max_iters": int(slab.iters.max()),
                     "median_iters": float(np.median(slab.iters)), "last_step_ms": last, "gen_seconds": gen_s,
                     "wire_seconds": slab.wire_s, "cpus_bound_to_gpu_numa_node": ctx.bound_cpus,
                     "sell_fill": slab.sell_fill,
                     "sell_fill_note": "live entries / places of the SELL slots (row + column side): what the padding to the "
                                       "longest line of a 32-line slice costs (DESIGN.md 9)"},
        }
    parity_ok = True
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        oracle.build()
        sel = np.sort(np.random.default_rng(0).choice(
            int(arena.sample_grp_ptr[1]), max(1, int(arena.sample_grp_ptr[1] / args.cpu_frac)), replace=False))
        sub = arena.select_groups(sel)
        w1, s1, cores = cpu_reference(sub, oracle.MODE_DENSE_TRACE, 1, 0, threads=os.cpu_count() or 1)
        w2, s2, _ = cpu_reference(sub, oracle.MODE_SPARSE, 3, 1, threads=os.cpu_count() or 1)
        line["cpu_baseline"] = {
            "value": w1 / s1, "unit": "nnz-iterations/s", "cores": cores, "kind": "port",
            "sample": f"{sub.n_groups} of the {int(arena.sample_grp_ptr[1])} gene groups of cohort sample 0 "
                      f"(random 1/{args.cpu_frac:g}), {int(sub.group_nonzeros().sum())} non-zeros, {s1:.1f} s",
            "note": "B1: dense literal restatement of src/em.cpp incl. theta_trace, groups over all cores (OpenMP)",
            "sparse_value": w2 * 3 / s2,
            "sparse_note": "B2: same arithmetic on the non-zeros only (CSR/CSC), all cores"}
        # the oracle as the checker: the e2e run's tables for these groups must be the oracle's, double for double
        line["parity"] = parity_block(arena, sel, (slab.h_est, slab.h_iters[:slab.n_groups], slab.h_status[:slab.n_groups]),
                                      cpu_reference.last)
        parity_ok = line["parity"]["iters_equal"] and line["parity"]["status_equal"] and line["parity"]["doubles_equal"]

    # ---- BASELINE configs[3] as worded: the 100-sample cohort over N GPUs; configs[2] sharded by groups ----------
    strong = None
    if world > 1 and not args.no_strong:
        strong = cohort_strong(ctx, args, slab, t1_dev_ms, solo_e2e_ms)
    slab.close()
    if table is not None:
        table.close()
    del slab, arena
    c3 = None if args.no_config3 else config3_block(ctx, args)
    if rank == 0:
        if world == 1:
            line["strong"] = {"samples_total": args.samples, "ms": line["ms_per_step"], "e2e_ms": line["e2e"]["ms_per_step"],
                              "efficiency": 1.0, "e2e_efficiency": 1.0, "note": "N = 1: the cohort on one GPU is the main line"}
        elif strong is not None:
            line["strong"] = strong
        if c3 is not None:
            line["config3_sharded" if world > 1 else "config3"] = c3
        if world == 1 and not args.no_stage:
            line["config2_single"] = config2_single(ctx, args)
            line["quantDT_stage"] = quantdt_stage(args)
            line["quant_stage_cohort"] = quant_stage_cohort(ctx, args)
        print(json.dumps(line))
        if not parity_ok:
            raise SystemExit("bench: GPU results differ from the oracle on cohort sample 0 (see 'parity')")
    if world > 1:
        ctx.dist.destroy_process_group()


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
    ap.add_argument("--no-stage", action="store_true", help="skip the quantDT_stage / config2_single latency measurements")
    ap.add_argument("--stage-samples", type=int, default=32, help="samples of the quant_stage_cohort measurement")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling pass (N > 1)")
    ap.add_argument("--no-config3", action="store_true", help="skip the config-3 (giant loci) block")
    ap.add_argument("--config3-scale", type=float, default=float(os.environ.get("BAMBU_BENCH_CONFIG3_SCALE", "1.0")),
                    help="1.0 = BASELINE configs[2] at full size (50 giant loci)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
