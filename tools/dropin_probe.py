import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for env in ({}, {"SPK_PACK_MODE": "single"}, {"SPK_PCOUNT_TABLE": "sweep"}, {"SPK_PMATRIX_KERNEL": "general"}, {"SPK_MAP_KERNEL": "tile"}):
    os.environ.update(env)
    d = bench.e2e_dropin()
    print(env, {k: d[k] for k in ("seconds", "split_genomes_s", "pipeline_s")}, flush=True)
    for k in env: os.environ.pop(k)
