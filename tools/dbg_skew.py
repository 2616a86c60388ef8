import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import spk_testutil as util
from subphaser_b200 import engine
from oracle import kmers
rng = np.random.default_rng(77)
which = sys.argv[1] if len(sys.argv) > 1 else "full"
if which == "full":
    seq = "A" * 4_600_000 + util.random_seq(rng, 40_000) + "AC" * 300_000 + "N" * 100 + "ACGGT" * 100_000
else:
    seq = "A" * 300_000 + util.random_seq(rng, 40_000) + "AC" * 30_000
fa = util.fasta([("skew", seq)])
for k, lower, gmax in ((17, 1, 700_000_000), (15, 2, 100_000_000)):
    d, n = engine.to_device_bytes(fa)
    s = engine.pack_fasta(d, n)
    table = engine.CountTable(max(s.n_bases, 1), k, lower, mode="partitioned", genome_max_bases=gmax)
    dump = engine.count_packed(s, k, lower, table=table)
    keys, counts = dump.to_host()
    okeys, ocounts, st = kmers.count_fasta(fa, k, lower)
    o = np.argsort(keys, kind="stable")
    print(k, lower, table.pbits, "equal:", np.array_equal(keys[o], okeys) and np.array_equal(counts[o], ocounts), flush=True)
