"""Device throughput on a config-5 slab (BASELINE configs[4] shape: 75k genes / 400k transcripts / 800k read classes /
~3.2 M entries per sample), wire form v2, resident plan.  usage: config5_slab.py [samples]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import bench
    from bambu_b200 import synthetic as S
    from bambu_b200.arena import Arena
    from bambu_b200.em import EmPlan, abundance_quantification_wire
    from bambu_b200.wire import Wire
    import time
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ann = S.annotation_for("config5")
    a = Arena.concatenate([S.cohort_sample(ann, s) for s in range(n)])
    w = Wire.from_arena(a)
    with EmPlan(w) as plan:
        plan.upload()
        ms = []
        for _ in range(4):
            plan.run(**bench.EM_PAR)
            plan.sync()
            ms.append(plan.last_run_ms()["total_ms"])
        est, iters, status = plan.download()
    work = float((w.nnz_nonzero * iters).sum())
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        abundance_quantification_wire(w, **bench.EM_PAR)
        t.append(time.perf_counter() - t0)
    print(f"config5 slab, {n} samples: {a.n_groups} groups, {a.nnz / 1e6:.1f} M entries ({a.nbytes() / 1e6:.0f} MB v1, {w.nbytes() / 1e6:.0f} MB wire), "
          f"{work / 1e9:.2f} G nnz-iterations, device {min(ms[1:]):.2f} ms -> {work / min(ms[1:]) / 1e6:.1f} G nnz-it/s, "
          f"e2e (pageable host buffers) {min(t) * 1e3:.1f} ms")


if __name__ == "__main__":
    main()
