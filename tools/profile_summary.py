"""Summarise an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch,
--csv) into a per-kernel table + the DRAM traffic of the counting kernel family.
    python tools/profile_summary.py gpurun_out/r02_launches.csv profiles/r02_launches_summary.md profiles/count_kernel_traffic.json
"""
import collections, csv, io, json, sys

src, out_md, out_json = sys.argv[1:4]
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.defaultdict(lambda: dict(n=0, ms=0.0, rd=0.0, wr=0.0))
ids = collections.defaultdict(set)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
for row in csv.DictReader(io.StringIO("".join(lines))):
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    m, v, u = row["Metric Name"], float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
    a = agg[name]
    if m == "gpu__time_duration.sum":
        a["ms"] += v * UNIT[u]
        ids[name].add(row["ID"])
    elif m == "dram__bytes_read.sum":
        a["rd"] += v * UNIT[u]
    elif m == "dram__bytes_write.sum":
        a["wr"] += v * UNIT[u]
for name in agg:
    agg[name]["n"] = len(ids[name])
tot = sum(a["ms"] for n, a in agg.items() if not n.startswith("k_synth"))
fam = ["k_v3_l1", "k_v3_plan", "k_v3_chunks", "k_v3_l2", "k_part_count32<1>", "k_part_count32<0>", "k_part_count32",
       "k_hist1", "k_scatter_l1", "k_scatter_l2", "k_scan_seg_totals", "k_scan_segs", "k_scan_apply", "k_scatter_prepare",
       "k_bucket_scan"]
with open(out_md, "w") as f:
    f.write("| kernel | launches | total ms | share of the step (input synthesis excluded) | DRAM read GB | DRAM write GB |\n|---|---|---|---|---|---|\n")
    for name, a in sorted(agg.items(), key=lambda x: -x[1]["ms"]):
        if a["ms"] < 0.05:
            continue
        share = "-" if name.startswith("k_synth") else "%.1f%%" % (100 * a["ms"] / tot)
        f.write("| `%s` | %d | %.2f | %s | %.2f | %.2f |\n" % (name[:60], a["n"], a["ms"], share, a["rd"] / 1e9, a["wr"] / 1e9))
    f.write("\nlisted kernel time without input synthesis: %.1f ms\n" % tot)
fam = [k for k in agg if k in fam or k.startswith("k_part_count32")]       # (any instantiation of the counter)
calls = max(agg[k]["n"] for k in agg if k.startswith("k_part_count32"))
dram = sum(agg[k]["rd"] + agg[k]["wr"] for k in fam if k in agg)
ms = sum(agg[k]["ms"] for k in fam if k in agg)
json.dump({"kernel_family": "spk_pcount_canonical_ex (" + ", ".join(k for k in fam if k in agg) + ")",
           "source": src + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, full C3 genome)",
           "calls": calls, "dram_bytes_total": dram, "dram_bytes_per_launch": dram / max(calls, 1),
           "ms_total_under_ncu": ms, "share_of_listed_step": ms / tot}, open(out_json, "w"), indent=1)
print(open(out_md).read())
print(open(out_json).read())
