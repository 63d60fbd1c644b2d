"""Latency of bambu.quantDT on one BASELINE-config-2 sample (2M-row long table): the one-call native entry with the
device packer vs the host packer, next to the numpy mirror.  Usage: python tools/quantdt_time.py [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bambu_b200 import quantify as Q, synthetic  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ann = synthetic.annotation_for("config2")
tab = synthetic.sample_arena(ann, 0, long_table=True)
print(f"long table: {len(tab['txid'])} rows")
os.environ["BAMBU_PACK_TIMING"] = "1"
for name, kw in (("device pack", {}), ("host pack", {"host_pack": True})):
    ts = []
    for r in range(reps):
        t = time.perf_counter()
        out = Q.bambu_quantDT_native(tab, **kw)
        ts.append(time.perf_counter() - t)
    print(f"bambu_quantDT [{name}]: best {min(ts) * 1e3:.1f} ms, median {np.median(ts) * 1e3:.1f} ms "
          f"({len(out['txid'])} transcripts, sum counts {out['counts'].sum():.6f}, d_rate {out['d_rate']:.6g})")
