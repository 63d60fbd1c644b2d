"""Quick device timing of one config-2 sample through the plan API (scratch tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bambu_b200 import synthetic as S
from bambu_b200.em import EmPlan, abundance_quantification

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = sys.argv[3] if len(sys.argv) > 3 else "config2"
maxiter = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
ann = S.annotation_for(cfg, scale=scale)
t = time.time()
arenas = [S.cohort_sample(ann, s) if cfg != "config3" else S.sample_arena(ann, s) for s in range(nsamp)]
from bambu_b200.arena import Arena
a = Arena.concatenate(arenas) if nsamp > 1 else arenas[0]
print("gen %.1fs groups %d nnz %d bytes %.1f MB" % (time.time() - t, a.n_groups, a.nnz, a.nbytes() / 1e6))
with EmPlan(a) as plan:
    plan.upload()
    for rep in range(4):
        plan.run(maxiter=maxiter)
        plan.sync()
        ms = plan.last_run_ms()
        print("run", rep, ms)
    est, iters, status, nz = plan.download(want_nnz=True)
    nl, nzl = plan.download_live(); sm, sl = plan.download_image_stats(); print("smem pct", np.percentile(sm,[0,10,50,90,99,100]), "slots*640 total MB", sl.sum()*640/1e6, "smem class hist", np.bincount(np.searchsorted([7168,14336,28672,57344,115712,232192], sm)))
    work = float((nz * iters).sum())
    live = float((nzl.astype(np.int64) * iters).sum())
    print("nnz-iters %.3f G (live %.3f G)  -> %.1f G nnz-it/s (live %.1f)  iters max %d  device MB %.0f launches %d" % (
        work / 1e9, live / 1e9, work / ms["total_ms"] / 1e6, live / ms["total_ms"] / 1e6, iters.max(),
        plan.device_bytes() / 1e6, plan.last_run_launches()))
t = time.time()
for rep in range(3):
    t = time.time()
    abundance_quantification(a, maxiter=maxiter)
    print("one-shot host call %.1f ms" % ((time.time() - t) * 1e3))
