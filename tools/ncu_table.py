"""Pivot an `ncu --metrics ... --csv` log into one row per kernel launch.  usage: ncu_table.py file.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid, ib, ig = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Block Size"), hdr.index("Grid Size")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iid], {"k": r[ik].split("(")[0].replace("void ", ""), "b": r[ib], "g": r[ig]})[r[im]] = r[iv]
short = [("gpu__time_duration.sum", "time_us", 1e-3), ("dram__bytes_read.sum", "dram_rd_MB", 1e-6), ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
         ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%", 1),
         ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1),
         ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1), ("launch__registers_per_thread", "regs", 1), ("launch__shared_mem_per_block_dynamic", "smem_KB", 1e-3)]
print("%-34s %-10s %-8s " % ("kernel", "block", "grid") + " ".join("%10s" % s[1] for s in short))
seen = set()
for k, v in d.items():
    key = (v["k"], v["b"], v.get("launch__shared_mem_per_block_dynamic"))
    if key in seen:
        continue
    seen.add(key)
    def f(name, sc):
        try:
            return "%10.2f" % (float(v.get(name, "nan").replace(",", "")) * sc)
        except ValueError:
            return "%10s" % v.get(name, "-")
    print("%-34s %-10s %-8s " % (v["k"][:34], v["b"].replace(" ", ""), v["g"].replace(" ", "").strip("()").split(",")[0]) + " ".join(f(n, sc) for n, _, sc in short))
