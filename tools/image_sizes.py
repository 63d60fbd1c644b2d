"""Shared-memory footprint of the resident groups of one config-4 sample, weighted by EM updates:
where the class boundaries (7/14/28 KB ...) cut the distribution.  usage: image_sizes.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bambu_b200 import synthetic as S
from bambu_b200.em import EmPlan
ann = S.annotation_for("config4")
a = S.cohort_sample(ann, 0)
with EmPlan(a) as plan:
    plan.upload()
    plan.run(maxiter=10000, minvalue=1e-8, conv=1e-2)
    plan.sync()
    est, iters, status = plan.download()
    smem, slots = plan.download_image_stats()
w = iters.astype(np.float64)
edges = [0, 4096, 6144, 7168, 8192, 10240, 11264, 12288, 13312, 13568, 14336, 16384, 18432, 20480, 24576, 28672, 57344, 1 << 30]
tot = w.sum()
print("groups %d, total updates %.0f" % (len(smem), tot))
for lo, hi in zip(edges[:-1], edges[1:]):
    m = (smem > lo) & (smem <= hi)
    print("smem in (%6d, %9d]: groups %5d  share of updates %5.1f %%" % (lo, hi, m.sum(), 100 * w[m].sum() / tot))
