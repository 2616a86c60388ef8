"""Scratch micro-benchmark of K1+K2+K3 on iid random sequence (not the judged bench)."""
import sys, time
import torch
sys.path.insert(0, ".")
from subphaser_b200 import engine, _lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 17
torch.cuda.init()
g = torch.Generator(device="cuda"); g.manual_seed(1)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
codes = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.uint8)
ascii_ = lut[codes.long()] if n <= 200_000_000 else torch.cat([lut[c.long()] for c in codes.split(100_000_000)])
del codes
d = torch.empty(n + 16, dtype=torch.uint8, device="cuda"); d[:n] = ascii_; del ascii_
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); seq = engine.pack_fasta(d, n, trim=False); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pack: %.1f ms  %.2f Gbases/s" % ((t1 - t0) * 1e3, n / (t1 - t0) / 1e9))
for mode in ("partitioned", "global"):
    tab = engine.CountTable(n, k, 3, mode=mode)
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dump = engine.count_packed(seq, k, 3, table=tab); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s count+dump: %.1f ms  %.2f G kmers/s (distinct %d)" % (mode, ms, n / ms / 1e6, dump.n_distinct))
    del tab
