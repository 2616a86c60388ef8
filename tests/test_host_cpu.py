"""CPU: host-only pieces of the drop-in surface against reference-generated fixtures (no GPU needed)."""
import io
import json
import os

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_stat_enrich_matches_reference(tmp_path):
    """stat_enrich.main (stat_enrich.py:4-37): byte-identical summary for the reference's own 4-column input, and the
    same numbers when the table carries the two extra columns Stats.enrich_ltr writes today (Stats.py:67)."""
    from subphaser_b200 import stat_enrich
    for i, case in enumerate(load("stat_enrich.json")):
        p = tmp_path / ("enrich%d.tsv" % i)
        p.write_text(case["input"])
        out = io.StringIO()
        stat_enrich.main(inTsv=str(p), outStat=out)
        assert out.getvalue() == case["output"]
        six = "\n".join(l + ("\tpotential_exchange\tp_corrected" if l.startswith("#") else "\tnone\t0.5")
                        for l in case["input"].splitlines()) + "\n"
        p6 = tmp_path / ("enrich%d_6col.tsv" % i)
        p6.write_text(six)
        out6 = io.StringIO()
        stat_enrich.main(inTsv=str(p6), outStat=out6)
        assert out6.getvalue() == case["output"]


def test_stat_enrich_missing_pair_and_import_without_argv(tmp_path):
    """An (annotation, subgenome) pair that never occurs counts as zero rows; importing the module must not read
    sys.argv (the reference does at definition time, stat_enrich.py:4)."""
    import importlib
    import sys
    argv, sys.argv = sys.argv, ["prog"]
    try:
        mod = importlib.reload(importlib.import_module("subphaser_b200.stat_enrich"))
    finally:
        sys.argv = argv
    p = tmp_path / "e.tsv"
    p.write_text("#id\tsubgenome\tp_value\tcounts\nCopia-1\tSG1\t0.1\t3,4\nGypsy-2\tSG2\t0.2\t10,20\nCopia-3\tSG1\t0.3\t1,1\n")
    out = io.StringIO()
    mod.main(inTsv=str(p), outStat=out)
    assert out.getvalue() == "Copia\t2\t0\t4\t5\nGypsy\t0\t1\t10\t20\n"


def test_group_exchanges_matches_reference_groups():
    """Stats.group_exchanges (Stats.py:119-132) on the rows of the pipeline fixture: the `.bin.group` the reference wrote."""
    from subphaser_b200 import Stats
    Gp = os.path.join(G, "pipeline_small")
    meta = json.load(open(os.path.join(Gp, "meta.json")))
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(Gp, "ref.bin.enrich"))][1:]
    lines = [[r[0], int(r[1]), int(r[2]), None if r[3] == "None" else r[3]] + r[4:] for r in rows]
    got = ["\t".join(map(str, g)) for g in Stats.group_exchanges(lines, meta["d_sg"])]
    ref = open(os.path.join(Gp, "ref.bin.group")).read().splitlines()[1:]
    assert got == ref
    # unsorted starts inside a chromosome block and a block without any enriched row
    lines2 = [["c1", 20, 30, "SG1"], ["c1", 0, 10, "SG1"], ["c1", 10, 20, "SG2"], ["c2", 0, 10, None], ["c1", 40, 50, "SG1"]]
    got2 = list(Stats.group_exchanges(lines2, {"c1": "SG1"}))
    assert got2 == [["c1", 0, 10, "SG1", "SG1", 1, "no"], ["c1", 10, 20, "SG2", "SG1", 1, "yes"],
                    ["c1", 20, 30, "SG1", "SG1", 1, "no"], ["c1", 40, 50, "SG1", "SG1", 1, "no"]]


def test_clock_sampler_summary_counts_only_samples_after_mark():
    """bench.ClockSampler: the `clocks` object of the bench line is built from the samples of the timed region only
    (the sampler is started before the warm-up), medians the SM clock and reports every active throttle reason."""
    import time
    import bench
    s = bench.ClockSampler(0)
    s.nvml = (None, None, 1965)            # (pretend the NVML poller is the source; nothing is started)
    t0 = time.perf_counter()
    s.lines = [(t0 - 5.0, "0,1200,1965,0,0x4,Not Active,Not Active,Not Active,Active"),        # warm-up: ignored
               (t0 + 0.1, "0,1965,1965,0,0x0,Not Active,Not Active,Not Active,Not Active"),
               (t0 + 0.3, "0,1950,1965,0,0x4,Not Active,Not Active,Not Active,Active"),
               (t0 + 0.5, "0,1965,1965,0,0x0,Not Active,Not Active,Not Active,Not Active"),
               (t0 + 0.7, "garbage")]
    s.t0 = t0
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["source"] == "nvml"
    s2 = bench.ClockSampler(0)
    assert s2.stop()["reasons"] == ["unavailable"]
