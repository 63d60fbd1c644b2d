"""Per-update latency of ONE group alone on the GPU (no contention): the longest-running groups of a config-2 sample,
each quantified on its own, in the latency and in the throughput instance of the EM kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from bambu_b200 import synthetic as S  # noqa: E402
from bambu_b200.em import EmPlan  # noqa: E402

def main():
    a = S.config2()
    with EmPlan(a) as plan:
        plan.upload()
        plan.run()
        est, iters, status = plan.download()
        print("whole sample:", plan.last_run_ms())
    M, N, e = a.group_shapes()
    for g in list(np.argsort(iters)[-3:]) + [int(np.argsort(iters)[len(iters) // 2])]:
        sub = a.select_groups([g])
        for mode, env in (("latency", None), ("throughput", "0")):
            if env is None:
                os.environ.pop("BAMBU_EM_LATENCY_GROUPS", None)
            else:
                os.environ["BAMBU_EM_LATENCY_GROUPS"] = env
            with EmPlan(sub) as plan:
                plan.upload()
                ms = []
                for _ in range(4):
                    plan.run()
                    plan.sync()
                    ms.append(plan.last_run_ms())
                _, it, _ = plan.download()
                nl, nzl = plan.download_live()
                sm, sl = plan.download_image_stats()
            best = min(m["em_ms"] for m in ms)
            print(f"group {g}: M {M[g]} N {N[g]} live cols {nl[0]} live entries {nzl[0]} slots {sl[0]} smem {sm[0]} iters {it[0]}  {mode:10s} "
                  f"EM {best:.3f} ms (pack {min(m['pack_ms'] for m in ms):.3f}) -> {best * 1e3 / max(it[0], 1):.3f} us per update")



if __name__ == "__main__":
    main()
