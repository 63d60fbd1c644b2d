"""Hot SASS instructions in address order from a `--print-source cuda,sass` csv.
usage: ncu_sass.py both.csv <kernel substring> [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.35
fn = hdr = cur_line = fpath = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == 'Function Name':
        fn = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if r[0] == 'File Path':
        fpath = r[1]; continue
    if hdr is None or len(r) != len(hdr) or pat not in (fn or ''):
        continue
    if r[0] != '':
        try:
            cur_line = (fpath.split('/')[-1], int(r[0]))
        except ValueError:
            cur_line = None
        continue
    if cur_line is None:
        continue
    iex, ith = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    try:
        out.append((r[2], cur_line, r[3].strip(), int(r[iex]), int(r[ith])))
    except ValueError:
        pass
out.sort()
tot = sum(o[3] for o in out)
for addr, line, sass, ex, th in out:
    if ex > minpct / 100 * tot:
        print(addr[-5:], '%-12s%4d' % line, '%5.2f%%' % (100 * ex / tot), '%4.1f' % (th / max(ex, 1)), sass[:100])
