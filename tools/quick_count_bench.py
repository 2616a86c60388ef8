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
tab = engine.CountTable(n, k)
print("layout", tab.layout, "table GB", tab.table_bytes / 1e9)
for rep in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    st = engine._stream()
    _lib.call("spk_count_table_init", engine._p(tab.table), tab.table_bytes, k, tab.layout, st)
    tab.stats.zero_()
    e0.record()
    _lib.call("spk_count_canonical", engine._p(seq.packed), engine._p(seq.valid), seq.n_bases, k, engine._p(tab.table), tab.table_bytes, tab.layout, engine._p(tab.stats), st)
    e1.record()
    _lib.call("spk_table_stats", engine._p(tab.table), tab.table_bytes, k, tab.layout, 3, engine._p(tab.stats[4:]), engine._p(tab.block_counts), None, 0, st)
    e2.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("count: %.1f ms  %.2f G kmers/s   scan %.1f ms  stats %s" % (ms, n / ms / 1e6, e1.elapsed_time(e2), tab.stats.tolist()))
