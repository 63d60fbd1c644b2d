"""bambu_quantDT_cohort on S config-4 tables (page-locked) for several worker counts.  usage: cohort_workers.py [S]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from bambu_b200 import quantify as Q  # noqa: E402
import multiprocessing as mp  # noqa: E402

def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    with mp.get_context("forkserver").Pool(min(S, os.cpu_count() or 8)) as pool:
        tables = pool.map(bench._gen_table, [(1.0, s) for s in range(S)], chunksize=1)
    pinned = []
    for t in tables:
        p = {}
        for k, dt in (("GENEID", np.int64), ("txid", np.int64), ("eqClassId", np.int64), ("nobs", np.float64),
                      ("equal", np.uint8), ("rcWidth", np.float64), ("txlen", np.float64)):
            src = torch.from_numpy(np.ascontiguousarray(np.asarray(t[k]).astype(dt)))
            dst = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            dst.copy_(src)
            p[k] = dst.numpy()
        p["equal"] = p["equal"].view(np.bool_)
        pinned.append(p)
    ann = np.arange(1, 250_001, dtype=np.int64)
    inc = np.zeros(S)
    for workers in (1, 2, 3, 4, 6, 8):
        os.environ["BAMBU_COHORT_WORKERS"] = str(workers)
        Q.bambu_quantDT_cohort_native(pinned[:workers], ann, inc[:workers])
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            Q.bambu_quantDT_cohort_native(pinned, ann, inc)
            ts.append(time.perf_counter() - t0)
        print(f"workers {workers}: {min(ts) * 1e3 / S:.2f} ms per sample ({S / min(ts):.1f} samples/s)", flush=True)
    os.environ["BAMBU_COHORT_WORKERS"] = "3"
    os.environ["BAMBU_PACK_TIMING"] = "1"
    Q.bambu_quantDT_cohort_native(pinned[:1], ann, inc[:1])



if __name__ == "__main__":
    main()
