"""One resident pass of a config-3 shard (giant loci only) -- what ncu wraps for the streamed tier.
usage: giant_once.py [n_loci] [runs]      (loci 3, 11, 19, ... of the full config 3; cached in /tmp)
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,... \
        -k regex:"gp_|em_giant" -s <13 * (runs - 1)> -c 13 python tools/giant_once.py 6 2
(per run: the 12 gp_* kernels of giant_pack.cuh, then em_giant_kernel)"""
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from bambu_b200 import synthetic as S
    from bambu_b200.em import EmPlan
    from bambu_b200.wire import Wire
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cache = f"/tmp/bambu_giant_{n}.pkl"
    if os.path.exists(cache):
        w = pickle.load(open(cache, "rb"))
    else:
        loci = [(3 + 8 * k) % 50 for k in range(n)]
        w = Wire.from_arena(S.config3_arena(1.0, loci=loci, small=False, workers=1))
        pickle.dump(w, open(cache, "wb"), protocol=4)
    with EmPlan(w) as plan:
        plan.upload()
        for _ in range(runs):
            plan.run(**bench.EM_PAR)
            plan.sync()
        ms = plan.last_run_ms()
        est, iters, status = plan.download()
    print(f"{n} giant loci: {w.nnz} live entries, last run {ms}, updates {iters.tolist()}, "
          f"nnz-iterations {float((w.nnz_nonzero * iters).sum()):.4g}")


if __name__ == "__main__":
    main()
