"""Scratch: per-kernel CUDA-event timing of the count family on one synthetic wheat-like chromosome.
usage: python tools/pcount_kernels.py [bases] [k] [genome_max_bases]   (env SPK_PCOUNT_PIPE=v2 for the old pipeline)
Prints the count+dump time; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from subphaser_b200 import engine, synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 676_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 17
gmax = int(float(sys.argv[3])) if len(sys.argv) > 3 else 851_000_000
reps = 3
plan = synth.GenomePlan(303, "ABD", [n] * 3)
d, nb = synth.synth_chromosome(plan, plan.chroms[0])
seq = engine.pack_fasta(d, nb)
del d
res = {}
for pipe in ("v3", "v2"):
    os.environ["SPK_PCOUNT_PIPE"] = pipe
    tab = engine.CountTable(seq.n_bases, k, 3, mode="partitioned", genome_max_bases=gmax)
    best = 1e9
    for rep in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        dump = engine.count_packed(seq, k, 3, table=tab)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    keys, counts = dump.to_host()
    o = np.argsort(keys, kind="stable")
    res[pipe] = (keys[o], counts[o], dump.length, dump.n_valid_kmers, dump.n_distinct)
    print("%s pbits=%d: %.2f ms  %.2f G kmers/s  (valid %d distinct %d dumped %d)" % (
        pipe, tab.pbits, best, dump.n_valid_kmers / best / 1e6, dump.n_valid_kmers, dump.n_distinct, len(dump)), flush=True)
    del tab, dump
    torch.cuda.empty_cache()
a, b = res["v3"], res["v2"]
print("v3 == v2:", bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:] == b[2:]))
