"""One resident pass of a bench slab (what ncu wraps): S cohort samples -> wire form v2 -> plan -> `runs` runs.
The slab is cached in /tmp (the plain run that has to precede an ncu run builds it, the run under ncu only loads it, so
no worker processes are started under the profiler).
usage: slab_once.py [samples] [runs]     capture the LAST run: `-k regex:em_sell_kernel -s <6 * (runs - 1)> -c 6`
(per run the six size classes launch in the order 1024, 512, 256, 128, 64 (13 KB), 64 (6 KB) threads)"""
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bambu_b200.em import EmPlan  # noqa: E402
from bambu_b200.wire import Wire  # noqa: E402

def main():
    samples = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cache = f"/tmp/bambu_slab_{samples}.pkl"
    if os.path.exists(cache):
        w = pickle.load(open(cache, "rb"))
    else:
        w = Wire.from_arena(bench.make_cohort(range(samples)))
        pickle.dump(w, open(cache, "wb"), protocol=4)
    with EmPlan(w) as plan:
        plan.upload()
        for _ in range(runs):
            plan.run(**bench.EM_PAR)
            plan.sync()
        ms = plan.last_run_ms()
        est, iters, status = plan.download()
    print(f"slab of {samples} samples: {w.n_groups} groups, last run {ms}, nnz-iterations {float((w.nnz_nonzero * iters).sum()):.4g}")



if __name__ == "__main__":
    main()
