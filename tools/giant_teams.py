"""Streamed tier: device time of config 3 (all 50 giant loci) and of a 6-locus shard (what one of 8 GPUs gets) for several
team counts of em_giant_kernel (BAMBU_EM_GIANT_TEAMS).  usage: giant_teams.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bambu_b200 import synthetic as S
    from bambu_b200.em import EmPlan
    from bambu_b200.wire import Wire
    import bench
    full = S.config3_arena(1.0, workers=min(16, os.cpu_count() or 8), mp_context="forkserver")
    shard = S.config3_arena(1.0, loci=[3, 11, 19, 27, 35, 43], small=False, workers=6, mp_context="forkserver")
    for name, arena in (("50 loci", full), ("6 loci", shard)):
        w = Wire.from_arena(arena)
        for teams, srt in (("12", "rb2"), ("12", "rb1"), ("10", "rb2")):
            os.environ["BAMBU_EM_GIANT_TEAMS"] = teams
            os.environ["BAMBU_EM_GIANT_RB"] = srt[2:]
            teams = teams + "/" + srt
            with EmPlan(w) as plan:
                plan.upload()
                ms = []
                for _ in range(3):
                    plan.run(**bench.EM_PAR)
                    plan.sync()
                    ms.append(plan.last_run_ms())
            print(f"{name}: teams {teams:8s} total {min(m['total_ms'] for m in ms):8.2f} ms  (pack {min(m['pack_ms'] for m in ms):6.2f})", flush=True)


if __name__ == "__main__":
    main()
