"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
executed warp-instructions and stall samples (top N lines)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.defaultdict(lambda: [0, 0])
cur_file = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
        continue
    if hdr is None or r[0] in ("Function Name",) or len(r) <= iI:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    try:
        a = agg[(cur_file, ln, r[1].strip()[:110])]
        a[0] += int(float(r[iS] or 0)); a[1] += int(float(r[iI] or 0))
    except ValueError:
        pass
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print("total samples %d warp-inst %d" % (ts, ti))
for (f, ln, src), v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100 * v[1] / max(ti, 1), 100 * v[0] / max(ts, 1), f, ln, src))
