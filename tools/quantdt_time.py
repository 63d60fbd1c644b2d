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

# the .C()-style entry: R's own vector types (int ids, logical-as-int), widened on the device
import ctypes  # noqa: E402
from bambu_b200 import _lib  # noqa: E402
L = _lib.lib()
ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)      # noqa: E731
i32 = lambda v: np.array([v], dtype=np.int32)          # noqa: E731
f64 = lambda v: np.array([v], dtype=np.float64)        # noqa: E731
n = len(tab["txid"])
cols = [np.asarray(tab["GENEID"]).astype(np.int32), np.asarray(tab["txid"], np.int32), np.asarray(tab["eqClassId"], np.int32),
        np.asarray(tab["nobs"], np.float64), np.asarray(tab["equal"], bool).astype(np.int32),
        np.asarray(tab["rcWidth"], np.float64), np.asarray(tab["txlen"], np.float64)]
t_out, c0, c1, c2 = np.zeros(n, np.int32), np.zeros(n), np.zeros(n), np.zeros(n)
k, d, rc = i32(0), f64(0), i32(-99)
os.environ.pop("BAMBU_PACK_TIMING", None)
ts = []
for r in range(reps + 1):
    t = time.perf_counter()
    L.bambu_quantDT_dotC(ptr(i32(n)), *[ptr(c) for c in cols], ptr(i32(1)), ptr(i32(10000)), ptr(f64(1e-8)), ptr(f64(1e-2)),
                         ptr(t_out), ptr(c0), ptr(c1), ptr(c2), ptr(k), ptr(d), ptr(rc))
    ts.append(time.perf_counter() - t)
assert rc[0] == 0
m = int(k[0])
same = np.array_equal(t_out[:m], out["txid"]) and c0[:m].tobytes() == out["counts"].tobytes()
print(f"bambu_quantDT_dotC [int32 ids, C call only]: best {min(ts[1:]) * 1e3:.1f} ms, median {np.median(ts[1:]) * 1e3:.1f} ms, "
      f"same result as bambu_quantDT: {same}")
