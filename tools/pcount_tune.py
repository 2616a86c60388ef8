"""Scratch: time the partitioned counter (v2) vs the global-table counter (v1) on one synthetic
wheat-like chromosome; run under `ncu --metrics gpu__time_duration.sum` for the per-phase split."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from subphaser_b200 import engine, synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 17
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["partitioned", "global"]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
plan = synth.GenomePlan(303, "ABD", [n] * 3)
d, nb = synth.synth_chromosome(plan, plan.chroms[0])
seq = engine.pack_fasta(d, nb)
del d
for mode in modes:
    tab = engine.CountTable(seq.n_bases, k, 3, mode=mode)
    for rep in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        dump = engine.count_packed(seq, k, 3, table=tab)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s split=%s: %.1f ms  %.2f G kmers/s  (valid %d distinct %d dumped %d)" % (
            mode, os.environ.get("SPK_PCOUNT_SPLIT", "6"), ms,
            dump.n_valid_kmers / ms / 1e6, dump.n_valid_kmers, dump.n_distinct, len(dump)))
    del tab
