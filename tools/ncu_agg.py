"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(io.StringIO("".join(lines))):
    if "gpu__time_duration" not in row.get("Metric Name", ""):
        continue
    ms = float(row["Metric Value"].replace(",", "")) / 1e6
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += ms
    tot += ms
print("total %.2f ms" % tot)
for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("%-50s %6d %10.3f ms %5.1f%%" % (k[:50], n, ms, 100 * ms / tot))
