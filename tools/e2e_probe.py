"""Stage times of hotpath.run in its four input/output modes (device or pinned-host inputs, results kept on the device
or copied to the host) on one GPU: which stages pay for the host copies.  Diagnostic, prints one line per mode."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from subphaser_b200 import engine, hotpath, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
plan, cfg = synth.plan_for("C3", scale=scale)
lengths = [c["length"] for c in plan.chroms]
d_lib = torch.from_numpy(plan.library).cuda()
dev_inputs = [synth.synth_chromosome(plan, c, d_library=d_lib) for c in plan.chroms]
host_inputs = []
for d, nb in dev_inputs:
    h = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    h.copy_(d[:nb])
    host_inputs.append((h, nb))
torch.cuda.synchronize()
kw = dict(labels=plan.labels, sgs=plan.sgs, k=cfg["k"], lower_count=3, min_fold=2, baseline=1, ratio=1, min_freq=200,
          max_freq=10000, nsg=len(plan.sg_letters), replicates=1000, max_pval=0.05, bin_size=10000,
          chunk_size=10_000_000, window_size=cfg["window"], seed=0)
for host_in in (False, True):
    for ret in (False, True):
        inp = host_inputs if host_in else dev_inputs
        for _ in range(3):
            hotpath.run(inp, host_inputs=host_in, return_host=ret, **kw)
        t = hotpath.StageTimer(True)
        for _ in range(3):
            hotpath.run(inp, host_inputs=host_in, return_host=ret, timer=t, **kw)
        print(json.dumps({"host_inputs": host_in, "return_host": ret,
                          **{k: round(v / 3, 2) for k, v in sorted(t.totals_ms().items())}}), flush=True)
