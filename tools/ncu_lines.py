"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu_lines.py both.csv <kernel substring> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
agg = {}
fn = fpath = hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fpath = r[1]; continue
    if r[0] == 'Function Name':
        fn = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == '' or pat not in (fn or ''):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    iex, ism, ith = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
    key = (fpath.split('/')[-1], line, r[1].strip()[:100])
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += int(r[iex]); a[1] += int(r[ism]); a[2] += int(r[ith])
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print('total warp-instructions', tot, 'samples', tots)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-18s %4d ex%5.1f%% smp%5.1f%% thr/inst %4.1f | %s' % (k[0], k[1], 100 * a[0] / tot, 100 * a[1] / max(tots, 1), a[2] / max(a[0], 1), k[2]))
