import sys, torch
sys.path.insert(0, ".")
from subphaser_b200 import engine, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
plan = synth.GenomePlan(303, "ABD", [n] * 3)
d, nb = synth.synth_chromosome(plan, plan.chroms[0])
for rep in range(3):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); seq = engine.pack_fasta(d, nb, trim=False); e1.record(); torch.cuda.synchronize()
    print("pack %.2f ms  %.1f Gbases/s  (bases %d valid %d)" % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6, seq.n_bases, seq.n_valid))
