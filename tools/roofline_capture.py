"""profiles/r2_roofline.json from an ncu report of the bench slab (tools/slab_once.py under `ncu --set full`).

    # on the GPU box (one call; the plain run first, as the profiling recipe asks):
    python tools/slab_once.py 100 2 > gpurun_out/slab_plain.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:em_sell_kernel -s 9 -c 3 \
        -o gpurun_out/r2_em_sell_slab100 python tools/slab_once.py 100 2
    # here (no GPU needed):
    python tools/roofline_capture.py gpurun_out/r2_em_sell_slab100.ncu-rep 100

Per EM kernel launch of ONE step (the six size classes): duration, DRAM bytes, and the fractions of the resources that
can bind a shared-memory resident kernel -- issue slots, shared-memory data pipe, FP64 pipe, DRAM.  bench.py reads the
file and reports `roofline.traffic` (sum of the DRAM bytes, scaled to its slab) and `roofline.binding` (the dominant
kernel's fractions) from it, with the git commit the capture was made at."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {
    "time_ns": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "warps_active": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "block": "launch__block_size",
    "grid": "launch__grid_size",
    "smem_dyn": "launch__shared_mem_per_block_dynamic",
    "inst": "smsp__inst_executed.sum",
    "lts_bytes": "lts__t_bytes.sum",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    k = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return None if v is None else v * k.get(unit.split("/")[0], 1)


def to_ns(v, unit):
    k = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
    return None if v is None else v * k.get(unit, 1)


def main():
    rep = sys.argv[1]
    samples = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in body:
        k = {"kernel": r[col["Kernel Name"]].split("(")[0]}
        for key, name in M.items():
            if name not in col:
                k[key] = None
                continue
            v, u = num(r[col[name]]), units[col[name]]
            if key in ("dram_read", "dram_write", "lts_bytes", "smem_dyn"):
                v = to_bytes(v, u)
            elif key == "time_ns":
                v = to_ns(v, u)
            k[key] = v
        kernels.append(k)
    kernels = [k for k in kernels if k["time_ns"]]
    total_ns = sum(k["time_ns"] for k in kernels)
    dom = max(kernels, key=lambda k: k["time_ns"])
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    res = {
        "capture": {"report": os.path.basename(rep), "samples": samples, "commit": commit,
                    "command": "ncu --set full --clock-control none --import-source on -k regex:em_sell_kernel -s 9 -c 3 "
                               f"python tools/slab_once.py {samples} 2",
                    "note": "durations under ncu are cold-cache and serialised: use the SHARES and the byte counts"},
        "kernels": [{"kernel": k["kernel"], "block": k["block"], "grid": k["grid"], "smem_dyn": k["smem_dyn"], "regs": k["regs"],
                     "share_of_em_time": k["time_ns"] / total_ns, "ms": k["time_ns"] / 1e6,
                     "dram_bytes": (k["dram_read"] or 0) + (k["dram_write"] or 0), "l2_bytes": k["lts_bytes"],
                     "issue_active": k["issue"], "smem_pipe": k["smem"], "fp64_pipe": k["fp64"], "dram": k["dram"],
                     "warps_active": k["warps_active"], "smem_wavefronts": k["smem_wavefronts"], "warp_instructions": k["inst"]}
                    for k in kernels],
        "dram_bytes_per_step": sum((k["dram_read"] or 0) + (k["dram_write"] or 0) for k in kernels),
        "dram_bytes_per_sample": sum((k["dram_read"] or 0) + (k["dram_write"] or 0) for k in kernels) / samples,
        "binding": {"kernel": f'{dom["kernel"]} <{int(dom["block"])} threads, {int(dom["smem_dyn"])} B>', "share_of_em_time": dom["time_ns"] / total_ns,
                    "issue": dom["issue"] / 100.0, "smem": dom["smem"] / 100.0, "fp64": dom["fp64"] / 100.0, "dram": dom["dram"] / 100.0,
                    "warps_active": dom["warps_active"] / 100.0,
                    "what": "measured fractions of peak for the dominant EM kernel: issue slots (smsp__issue_active), shared-memory "
                            "data pipe (l1tex lsu wavefronts), FP64 pipe, DRAM throughput"},
    }
    dst = os.path.join(ROOT, "profiles", "r2_roofline.json")
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res["binding"], indent=1))
    for k in res["kernels"]:
        print("%-28s block %4d smem %6d  share %.3f  issue %5.1f%%  smem %5.1f%%  fp64 %5.1f%%  dram %5.1f%%  warps %5.1f%%  dram MB %8.1f" % (
            k["kernel"][-28:], k["block"], k["smem_dyn"], k["share_of_em_time"], k["issue_active"], k["smem_pipe"], k["fp64_pipe"], k["dram"],
            k["warps_active"], k["dram_bytes"] / 1e6))
    print("written", dst)


if __name__ == "__main__":
    main()
