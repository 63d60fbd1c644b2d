"""Throughput on the synthetic read-class sets BASELINE.json names (configs[1..4]), one GPU:
device-resident (plan API, CUDA events) and end to end (bambu_em_batched, host buffers), next to
the CPU sparse oracle (all cores) on the same arena.  Prints one line per config.
usage: measure_configs.py [config3_scale] [cohort_samples]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from bambu_b200 import synthetic as S
from bambu_b200.arena import Arena
from bambu_b200.em import EmPlan, abundance_quantification

c3_scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
n_cohort = int(sys.argv[2]) if len(sys.argv) > 2 else 16
PAR = dict(maxiter=10000, minvalue=1e-8, conv=1e-2)


def measure(name, a, cpu=True, reps=3):
    with EmPlan(a) as plan:
        plan.upload()
        ms = []
        for _ in range(reps + 1):
            plan.run(**PAR)
            plan.sync()
            ms.append(plan.last_run_ms())
        est, iters, status, nz = plan.download(want_nnz=True)
    dev = min(m["total_ms"] for m in ms[1:])
    pack = min(m["pack_ms"] for m in ms[1:])
    work = float((nz * iters).sum())
    t = []
    for _ in range(2):
        t0 = time.perf_counter()
        e2, i2, s2 = abundance_quantification(a, **PAR)
        t.append(time.perf_counter() - t0)
    assert np.array_equal(i2, iters)
    line = (f"{name:<34} groups {a.n_groups:>7} entries {a.nnz/1e6:8.2f} M  nnz-iters {work/1e9:8.3f} G  max iters {int(iters.max()):>5}  "
            f"device {dev:9.2f} ms (pack {pack:7.2f}) -> {work/dev/1e6:8.1f} G nnz-it/s   e2e {min(t)*1e3:9.2f} ms -> {work/min(t)/1e9:7.1f} G nnz-it/s")
    if cpu:
        t0 = time.perf_counter()
        o_est, o_it, o_st = oracle.em_batched(a, mode=oracle.MODE_SPARSE, n_threads=os.cpu_count(), **PAR)
        dt = time.perf_counter() - t0
        same = np.array_equal(o_it, iters) and np.array_equal(est, o_est, equal_nan=True)
        line += f"   CPU sparse x{os.cpu_count()} {dt:7.2f} s -> {work/dt/1e9:6.2f} G nnz-it/s   identical: {same}"
    print(line, flush=True)


def main():
    t0 = time.time()
    measure("config2 single sample", S.config2())
    a3 = S.config3_arena(c3_scale, workers=min(16, os.cpu_count() or 8), mp_context="forkserver")   # locus by locus
    measure(f"config3 skewed loci (scale {c3_scale:g}: {S.config3_plan(c3_scale).n_giant} giant)", a3)
    ann4 = S.annotation_for("config4")
    measure(f"config4 cohort slab ({n_cohort} samples)", Arena.concatenate([S.cohort_sample(ann4, s) for s in range(n_cohort)]), cpu=False)
    ann5 = S.annotation_for("config5")
    measure(f"config5 cohort slab ({n_cohort // 2} samples)", Arena.concatenate([S.cohort_sample(ann5, s) for s in range(n_cohort // 2)]), cpu=False)
    print("total %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
