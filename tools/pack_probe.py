"""Pack one wheat-sized synthetic chromosome a few times (for `ncu --kernel-name regex:k_pack` launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from subphaser_b200 import engine, synth

plan, cfg = synth.plan_for("C3", scale=1.0)
d_lib = torch.from_numpy(plan.library).cuda()
d, nb = synth.synth_chromosome(plan, plan.chroms[4], d_library=d_lib)
for _ in range(3):
    seq = engine.pack_fasta(d, nb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    seq = engine.pack_fasta(d, nb)
e1.record()
torch.cuda.synchronize()
print("bytes", nb, "bases", seq.n_bases, "path", seq.pack_path, "ms/pack", e0.elapsed_time(e1) / 5)
