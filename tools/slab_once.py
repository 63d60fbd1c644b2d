"""One resident pass of the bench slab (what ncu wraps): S cohort samples -> wire form v2 -> plan -> warm-up run + one run.
usage: slab_once.py [samples] [runs]      (kernels of the LAST run are the ones to capture: skip 0 launches with
`-k regex:em_sell_kernel -s <6 * (runs - 1)>`)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bambu_b200.em import EmPlan  # noqa: E402
from bambu_b200.wire import Wire  # noqa: E402

samples = int(sys.argv[1]) if len(sys.argv) > 1 else 100
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
arena = bench.make_cohort(range(samples))
w = Wire.from_arena(arena)
with EmPlan(w) as plan:
    plan.upload()
    for _ in range(runs):
        plan.run(**bench.EM_PAR)
        plan.sync()
    ms = plan.last_run_ms()
    est, iters, status = plan.download()
print(f"slab of {samples} samples: {w.n_groups} groups, last run {ms}, nnz-iterations {float((w.nnz_nonzero * iters).sum()):.4g}")
