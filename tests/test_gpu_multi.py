"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the sharded hot path must give bit-identical results to one
rank — differential matrix (keys, normalised values), window counts, enrichment, subgenome labels."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _digest(res):
    from subphaser_b200 import engine
    dm = res["dm"]
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    return dict(keys=h(engine.u64_numpy(dm.keys)), norm=h(dm.norm.cpu().numpy()), tot=h(dm.tot.cpu().numpy()),
                windows=h(res["window_counts"].cpu().numpy()), pvals=h(res["enrich"]["pvals"]),
                qvals=h(res["enrich"]["qvals"]), labels=res["labels_full"], d_bs=res["d_bs"], n_union=res["n_union"],
                n_diff=res["n_diff"], n_sig=res["n_sig"], n_windows=res["n_windows"], lengths=res["lengths"],
                n_kmers=res["n_kmers"])


def _genome():
    from subphaser_b200 import synth
    plan = synth.GenomePlan(77, "AB", [1_300_000, 900_000, 1_100_000, 1_250_000, 700_000, 1_000_000], n_fam=12,
                            n_shared=6, fam_len=(300, 900))
    kw = dict(labels=plan.labels, sgs=plan.sgs, k=15, lower_count=3, min_freq=50, nsg=2, replicates=64,
              window_size=100_000, bin_size=10_000, chunk_size=400_000, seed=3)
    return plan, kw


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from subphaser_b200 import hotpath, synth
    plan, kw = _genome()
    lengths = [c["length"] for c in plan.chroms]
    owner = hotpath.lpt_assign(lengths, world)
    inputs = []
    for i, c in enumerate(plan.chroms):
        if owner[i] == rank:
            inputs.append(synth.synth_chromosome(plan, c))
        else:
            inputs.append((None, plan.fasta_nbytes(c)[1]))
    for mode in ("peer", "class", "gather"):
        os.environ["SPK_EXCHANGE"] = mode
        res = hotpath.run(inputs, dist=dist, owner=owner, **kw)
        res = hotpath.run(inputs, dist=dist, owner=owner, **kw)      # a second pass re-uses the peer buffers
        with open(os.path.join(out_dir, "rank%d_%s.json" % (rank, mode)), "w") as f:
            json.dump(_digest(res), f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_bit_identical_to_one(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from subphaser_b200 import hotpath, synth
    plan, kw = _genome()
    inputs = [synth.synth_chromosome(plan, c) for c in plan.chroms]
    want = _digest(hotpath.run(inputs, **kw))
    assert want["n_diff"] > 100 and want["n_windows"] > 20
    port = 29600 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        for mode in ("peer", "class", "gather"):
            got = json.load(open(os.path.join(tmp_path, "rank%d_%s.json" % (rank, mode))))
            assert got == want, (rank, mode)
