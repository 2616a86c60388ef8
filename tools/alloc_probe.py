"""cudaMalloc / cudaFree calls per hotpath.run step (torch allocator statistics): a steady-state step should make none."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from subphaser_b200 import engine, hotpath, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
plan, cfg = synth.plan_for("C3", scale=scale)
d_lib = torch.from_numpy(plan.library).cuda()
dev_inputs = [synth.synth_chromosome(plan, c, d_library=d_lib) for c in plan.chroms]
torch.cuda.synchronize()
kw = dict(labels=plan.labels, sgs=plan.sgs, k=cfg["k"], lower_count=3, min_fold=2, baseline=1, ratio=1, min_freq=200,
          max_freq=10000, nsg=len(plan.sg_letters), replicates=1000, max_pval=0.05, bin_size=10000,
          chunk_size=10_000_000, window_size=cfg["window"], seed=0)
prev = torch.cuda.memory_stats()
for step in range(8):
    t = hotpath.StageTimer(True)
    res = hotpath.run(dev_inputs, timer=t, **kw)
    ms = t.totals_ms()
    st = torch.cuda.memory_stats()
    print(step, "device_alloc", st["num_device_alloc"] - prev["num_device_alloc"], "device_free",
          st["num_device_free"] - prev["num_device_free"], "reserved_GB", round(st["reserved_bytes.all.current"] / 1e9, 2),
          "matrix_ms", round(ms.get("matrix", 0), 1), "run_ms", round(ms.get("_run", 0), 1), flush=True)
    prev = st
