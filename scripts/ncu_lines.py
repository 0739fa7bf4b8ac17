"""Aggregate an ncu source-page CSV (`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`) by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = []; tot = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; ie = hdr.index('Instructions Executed'); ti = hdr.index('Thread Instructions Executed'); ws = hdr.index('Warp Stall Sampling (All Samples)'); continue
    if r[0] == 'Function Name' or not hdr or r[0] == '': continue
    try:
        v = float(r[ie]); t = float(r[ti]); w = float(r[ws] or 0)
    except (ValueError, IndexError):
        continue
    agg.append((v, cur, r[0], r[1][:105], t, w)); tot += v
agg.sort(reverse=True)
ts = sum(a[5] for a in agg) or 1
print("total warp inst", tot)
for v, f, l, src, t, w in agg[:top]:
    print("%5.1f%% inst %5.1f%% stall  thr %4.1f  %s:%s | %s" % (100 * v / tot, 100 * w / ts, t / max(v, 1), f[:14], l, src))
