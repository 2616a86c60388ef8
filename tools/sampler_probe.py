"""Does polling the clocks perturb the measured step?  Per-step matrix / run times of hotpath.run with no sampler, the
in-process NVML sampler and the nvidia-smi child (bench.ClockSampler variants)."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from subphaser_b200 import engine, hotpath, synth

plan, cfg = synth.plan_for("C3", scale=1.0)
d_lib = torch.from_numpy(plan.library).cuda()
dev_inputs = [synth.synth_chromosome(plan, c, d_library=d_lib) for c in plan.chroms]
torch.cuda.synchronize()
kw = dict(labels=plan.labels, sgs=plan.sgs, k=cfg["k"], lower_count=3, min_fold=2, baseline=1, ratio=1, min_freq=200,
          max_freq=10000, nsg=len(plan.sg_letters), replicates=1000, max_pval=0.05, bin_size=10000,
          chunk_size=10_000_000, window_size=cfg["window"], seed=0)
for _ in range(3):
    hotpath.run(dev_inputs, **kw)
for mode in ("none", "nvml", "smi", "none"):
    s = None
    if mode != "none":
        s = bench.ClockSampler(0)
        if mode == "smi":
            import pynvml
            real = pynvml.nvmlInit
            pynvml.nvmlInit = lambda: (_ for _ in ()).throw(RuntimeError("forced"))
            s.start()
            pynvml.nvmlInit = real
        else:
            s.start()
        time.sleep(1.5)
    out = []
    for step in range(8):
        t = hotpath.StageTimer(True)
        hotpath.run(dev_inputs, timer=t, **kw)
        ms = t.totals_ms()
        out.append((round(ms["matrix"], 1), round(ms["_run"], 1)))
    if s is not None:
        s.mark()
        print(mode, s.stop())
    print(mode, out, flush=True)
