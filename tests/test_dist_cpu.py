"""CPU, world_size 2, gloo: the host logic of the multi-GPU path — chromosome sharding (LPT) and the one
exchange step (dump merge + window gather) — without a GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_assignment_balances_wheat():
    from subphaser_b200 import hotpath, synth
    lengths = [m * 1_000_000 for m in synth.WHEAT_MB]
    for world in (1, 2, 4, 8):
        owner = hotpath.lpt_assign(lengths, world)
        assert len(owner) == 21 and set(owner) == set(range(world))
        load = [sum(l for l, o in zip(lengths, owner) if o == r) for r in range(world)]
        # greedy LPT gives 1.135 at 8 ranks; the local search brings it to 1.08 (and ~1.00 at 2 / 4)
        assert max(load) / (sum(load) / world) < {1: 1.0001, 2: 1.005, 4: 1.01, 8: 1.09}[world]
    assert hotpath.lpt_assign(lengths, 8) == hotpath.lpt_assign(lengths, 8)   # deterministic on every rank


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from subphaser_b200 import hotpath
    n = 5
    lengths = [50, 40, 30, 20, 10]
    owner = hotpath.lpt_assign(lengths, world)
    dev = torch.device("cpu")
    g = torch.Generator().manual_seed(7)
    full = {}
    for i in range(n):
        m = 3 + 2 * i if i != 3 else 0                          # one chromosome with an empty dump
        full[i] = (torch.randint(0, 2**40, (m,), generator=g, dtype=torch.int64),
                   torch.randint(3, 100, (m,), generator=g, dtype=torch.int32), 1000 + i)
    local = {i: full[i] for i in range(n) if owner[i] == rank}
    out, total = hotpath.exchange_dumps(local, n, owner, dist, dev, n_kmers_local=100 * (rank + 1))
    ok = total == sum(100 * (r + 1) for r in range(world))
    for i in range(n):
        ok &= bool(torch.equal(out[i][0], full[i][0]) and torch.equal(out[i][1], full[i][1]) and out[i][2] == full[i][2])
    # partition indices of the dumps travel with them (same partition bits on every rank)
    pbits = 3
    pidx_full = {i: torch.arange(2 << pbits, dtype=torch.int32) + 100 * i for i in range(n)}
    pidx, pb = hotpath.exchange_pindex({i: pidx_full[i] for i in range(n) if owner[i] == rank}, n, owner, dist, dev, pbits)
    ok &= pb == pbits and all(bool(torch.equal(pidx[i], pidx_full[i])) for i in range(n))
    # ranks that disagree on the partition bits (or lack an index) fall back to the plain union on every rank
    pidx2, pb2 = hotpath.exchange_pindex({i: pidx_full[i] for i in range(n) if owner[i] == rank}, n, owner, dist, dev,
                                         pbits + rank)
    ok &= pb2 == 0 and pidx2 == {}
    # class-wise exchange: a rank receives only the hash partitions p = rank (mod world) of the foreign dumps
    pb = 4
    P = 1 << pb
    dumps_full, pidx_all = {}, {}
    for i in range(n):
        gi = torch.Generator().manual_seed(100 + i)
        cnt = torch.randint(0, 6, (P,), generator=gi)
        if i == 3:
            cnt[:] = 0
        order = torch.randperm(P, generator=gi)                     # partitions lie in arbitrary order in the dump
        start = torch.zeros(P, dtype=torch.int64)
        start[order] = torch.cumsum(cnt[order], 0) - cnt[order]
        tot = int(cnt.sum())
        kk = torch.randint(0, 2**40, (tot,), generator=gi, dtype=torch.int64)
        cc = torch.randint(3, 100, (tot,), generator=gi, dtype=torch.int32)
        pi = torch.zeros(2 * P, dtype=torch.int32)
        pi[0::2] = start.to(torch.int32)
        pi[1::2] = cnt.to(torch.int32)
        dumps_full[i] = (kk, cc, 2000 + i)
        pidx_all[i] = pi
    got, total2 = hotpath.exchange_dumps_by_class({i: dumps_full[i] for i in range(n) if owner[i] == rank},
                                                  {i: pidx_all[i] for i in range(n) if owner[i] == rank}, pb, n, owner,
                                                  dist, dev, n_kmers_local=7 * (rank + 1))
    ok &= total2 == sum(7 * (r + 1) for r in range(world))
    ok &= sorted(got) == [i for i in range(n) if owner[i] != rank]
    for i, (kk, cc, length, pi) in got.items():
        fk, fc, fl = dumps_full[i]
        ok &= length == fl
        for p_ in range(P):
            a, m = int(pi[2 * p_]), int(pi[2 * p_ + 1])
            if p_ % world != rank:
                ok &= m == 0
                continue
            fa, fm = int(pidx_all[i][2 * p_]), int(pidx_all[i][2 * p_ + 1])
            ok &= m == fm and bool(torch.equal(kk[a:a + m], fk[fa:fa + fm]) and torch.equal(cc[a:a + m], fc[fa:fa + fm]))
    wins = {i: torch.full((i + 1, 3), i, dtype=torch.int64) for i in range(n) if owner[i] == rank}
    allw = hotpath.exchange_windows(wins, n, 3, owner, dist, dev)
    for i in range(n):
        ok &= bool(torch.equal(allw[i], torch.full((i + 1, 3), i, dtype=torch.int64)))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _worker_class(rank, world, port, q):
    """class-wise exchange only, for world sizes that do not divide the partition count evenly"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from subphaser_b200 import hotpath
    n, pb = 7, 5
    P = 1 << pb
    owner = hotpath.lpt_assign([70, 60, 50, 40, 30, 20, 10], world)
    dev = torch.device("cpu")
    full, pidx_all = {}, {}
    for i in range(n):
        gi = torch.Generator().manual_seed(500 + i)
        cnt = torch.randint(0, 9, (P,), generator=gi)
        order = torch.randperm(P, generator=gi)
        start = torch.zeros(P, dtype=torch.int64)
        start[order] = torch.cumsum(cnt[order], 0) - cnt[order]
        tot = int(cnt.sum())
        full[i] = (torch.randint(0, 2**40, (tot,), generator=gi, dtype=torch.int64),
                   torch.randint(3, 100, (tot,), generator=gi, dtype=torch.int32), 3000 + i)
        pi = torch.zeros(2 * P, dtype=torch.int32)
        pi[0::2] = start.to(torch.int32)
        pi[1::2] = cnt.to(torch.int32)
        pidx_all[i] = pi
    mine = [i for i in range(n) if owner[i] == rank]
    got, total = hotpath.exchange_dumps_by_class({i: full[i] for i in mine}, {i: pidx_all[i] for i in mine}, pb, n,
                                                 owner, dist, dev, n_kmers_local=rank + 1)
    ok = total == world * (world + 1) // 2 and sorted(got) == [i for i in range(n) if owner[i] != rank]
    for i, (kk, cc, length, pi) in got.items():
        fk, fc, fl = full[i]
        ok &= length == fl and int(pi[1::2].sum()) == int(kk.numel()) == int(cc.numel())
        for p_ in range(P):
            a, m = int(pi[2 * p_]), int(pi[2 * p_ + 1])
            fa, fm = int(pidx_all[i][2 * p_]), int(pidx_all[i][2 * p_ + 1])
            if p_ % world != rank:
                ok &= m == 0
            else:
                ok &= m == fm and bool(torch.equal(kk[a:a + m], fk[fa:fa + fm]) and torch.equal(cc[a:a + m], fc[fa:fa + fm]))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [3, 4])
def test_class_exchange_more_ranks_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_class, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def _worker_sig(rank, world, port, q):
    """gather_sig (both forms) and exchange_rows_meta: every rank ends with every rank's significant k-mers"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from subphaser_b200 import hotpath
    dev = torch.device("cpu")

    def part(r):
        g = torch.Generator().manual_seed(100 + r)
        n = [0, 5, 1000, 37][r % 4]                     # (rank 0 has no significant k-mer at all)
        return (torch.randint(0, 1 << 34, (n,), generator=g, dtype=torch.int64),
                torch.randint(0, 3, (n,), generator=g, dtype=torch.int64).to(torch.uint8))

    keys, vals = part(rank)
    want_k = torch.cat([part(r)[0] for r in range(world)])
    want_v = torch.cat([part(r)[1] for r in range(world)])
    ok = True
    for kw in (dict(max_rows=1000), dict(), dict(max_rows=1000, key_bits=60)):
        kb = kw.pop("key_bits", 34)
        gk, gv = hotpath.gather_sig(keys, vals, kb, dist, dev, **kw)
        ok &= bool(torch.equal(gk, want_k) and torch.equal(gv, want_v))

    class Shard:
        n_fold_pass = 10 * (rank + 1)

        def __len__(self):
            return 3 + rank

    m, n_union, n_fold = hotpath.exchange_rows_meta(Shard(), 100 + rank, dist, dev)
    ok &= m == [3 + r for r in range(world)] and n_union == sum(100 + r for r in range(world))
    ok &= n_fold == sum(10 * (r + 1) for r in range(world))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_sig_and_row_meta_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_sig, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]
