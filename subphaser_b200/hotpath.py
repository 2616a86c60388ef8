"""Whole hot path on device buffers: count -> matrix -> filter -> cluster/bootstrap -> specific k-mers ->
map -> windows -> Fisher/BH, without touching the file system.  This is the same sequence the drop-in
modules run for `__main__.py` (pipeline.py); bench.py and smoke() time it here so that the timed
region contains only the copies and kernels of the path.

Multi-GPU (one process per GPU, torch.distributed/NCCL): chromosomes are sharded over ranks by
longest-processing-time assignment; each rank packs, counts and later maps only its chromosomes.  The
one exchange step is the merge of the per-chromosome dumps into the global k-mer table: dumps are
broadcast from their owners over NVLink (exact, variable-size) and `lengths` are all-reduced; the
per-window counts are all-gathered before the genome-wide Fisher step.
"""
import os
import time

import numpy as np
import torch

from . import _lib, engine


def lpt_assign(lengths, world):
    """Chromosomes -> owner rank.  Longest-processing-time greedy, then deterministic local search (moves and
    pairwise swaps that lower the larger of two ranks' loads): with 21 chromosomes on 8 ranks the greedy
    alone leaves the heaviest rank 13.6 % above the mean, and every rank waits for it at the exchange."""
    n = len(lengths)
    load = [0] * world
    owner = [0] * n
    for i in sorted(range(n), key=lambda i: (-lengths[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += lengths[i]
    improved = True
    while improved:
        improved = False
        a = max(range(world), key=lambda r: (load[r], -r))          # heaviest rank
        best = None                                                   # (new pair maximum, move/swap)
        for i in range(n):
            if owner[i] != a:
                continue
            for b in range(world):
                if b == a:
                    continue
                cand = max(load[a] - lengths[i], load[b] + lengths[i])          # move i: a -> b
                if cand < load[a] and (best is None or cand < best[0]):
                    best = (cand, i, None, b)
                for j in range(n):
                    if owner[j] != b:
                        continue
                    d = lengths[i] - lengths[j]
                    cand = max(load[a] - d, load[b] + d)                        # swap i <-> j
                    if cand < load[a] and (best is None or cand < best[0]):
                        best = (cand, i, j, b)
        if best is not None:
            _, i, j, b = best
            owner[i] = b
            load[a] -= lengths[i]
            load[b] += lengths[i]
            if j is not None:
                owner[j] = a
                load[b] -= lengths[j]
                load[a] += lengths[j]
            improved = True
    return owner


def _gather_concat(parts, sizes, owner, n, dist, dev, dtype, group=None):
    """parts: {chromosome index: 1-D tensor} owned by this rank; sizes[i]: its length on every rank.
    One all_gather of the per-rank concatenation (padded to the longest) -> {i: view} for every chromosome."""
    rank, world = dist.get_rank(), dist.get_world_size()
    per_rank = [[i for i in range(n) if owner[i] == r] for r in range(world)]
    tot = [sum(sizes[i] for i in per_rank[r]) for r in range(world)]
    width = max(max(tot), 1)
    send = torch.empty(width, dtype=dtype, device=dev)       # the padding is never read
    off = 0
    for i in per_rank[rank]:
        if sizes[i]:
            send[off:off + sizes[i]] = parts[i]
        off += sizes[i]
    recv = torch.empty(world, width, dtype=dtype, device=dev)
    try:
        dist.all_gather_into_tensor(recv.view(-1), send, group=group)   # one flat collective, no per-rank output copies
    except (RuntimeError, NotImplementedError, AttributeError):
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    out = {}
    for r in range(world):
        off = 0
        for i in per_rank[r]:
            out[i] = recv[r, off:off + sizes[i]]
            off += sizes[i]
    return out


def exchange_dumps(local, n, owner, dist, dev, n_kmers_local=0):
    """The one exchange step of the path: after it every rank holds every chromosome's dump.
    local: {chromosome index: (keys int64 tensor, counts int32 tensor, length)} for the chromosomes this
    rank owns.  Sizes / lengths / k-mer totals travel in one all_reduce, the dumps in two all_gathers of the
    per-rank concatenations (keys, counts) — 3 collectives instead of 2 per chromosome.  Works on any
    backend (NCCL on device tensors; gloo on CPU tensors in the tests)."""
    meta = torch.zeros(n + 1, 2, dtype=torch.int64, device=dev)
    for i, (kk, cc, length) in local.items():
        meta[i, 0], meta[i, 1] = int(kk.numel()), int(length)
    meta[n, 0] = int(n_kmers_local)
    dist.all_reduce(meta)
    meta_h = meta.cpu().tolist()
    sizes = [int(meta_h[i][0]) for i in range(n)]
    keys = _gather_concat({i: v[0] for i, v in local.items()}, sizes, owner, n, dist, dev, torch.int64)
    counts = _gather_concat({i: v[1] for i, v in local.items()}, sizes, owner, n, dist, dev, torch.int32)
    out = {i: (keys[i], counts[i], int(meta_h[i][1])) for i in range(n)}
    return out, int(meta_h[n][0])


def exchange_pindex(local, n, owner, dist, dev, pbits_local):
    """Partition indices of the dumps (int32 [2 << pbits] each) -> present on every rank.  All ranks must use
    the same partition bits (the table is sized for the largest chromosome of the genome); returns ({}, 0)
    when any rank has none."""
    ok = all(v is not None for v in local.values())
    meta = torch.tensor([int(pbits_local) if ok else -1, -(int(pbits_local) if ok else -1)], dtype=torch.int64, device=dev)
    dist.all_reduce(meta, op=dist.ReduceOp.MAX)
    pmax, pmin = int(meta[0].item()), -int(meta[1].item())
    if pmax != pmin or pmax <= 0:
        return {}, 0
    out = _gather_concat(local, [2 << pmax] * n, owner, n, dist, dev, torch.int32)
    return out, pmax


def _all_to_all(send, in_splits, out_splits, dist, dev):
    """Ragged all-to-all of a 1-D tensor; NCCL does it in one collective, other backends (gloo in the CPU
    tests) get it as one broadcast per source rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    recv = torch.empty(int(sum(out_splits)), dtype=send.dtype, device=dev)
    try:
        dist.all_to_all_single(recv, send, [int(x) for x in out_splits], [int(x) for x in in_splits])
        return recv
    except (RuntimeError, NotImplementedError, AttributeError):
        pass
    sizes = torch.zeros(world, world, dtype=torch.int64, device=dev)
    sizes[rank] = torch.tensor([int(x) for x in in_splits], dtype=torch.int64, device=dev)
    dist.all_reduce(sizes)
    sizes_h = sizes.cpu().tolist()
    off_out = 0
    for src in range(world):
        tot = int(sum(sizes_h[src]))
        buf = send if src == rank else torch.empty(tot, dtype=send.dtype, device=dev)
        if tot:
            dist.broadcast(buf, src=src)
        a = int(sum(sizes_h[src][:rank]))
        m = int(sizes_h[src][rank])
        recv[off_out:off_out + m] = buf[a:a + m]
        off_out += m
    return recv


def exchange_dumps_by_class(local, pindex_local, pbits, n, owner, dist, dev, n_kmers_local=0):
    """The exchange step with 1/world of the volume: rank r builds the union rows of the hash partitions
    p = r (mod world) only, so it needs just that class of every other chromosome's dump.  Every owned dump is
    regrouped class-major (its partitions are contiguous runs located by pindex), the classes travel in three
    ragged all-to-alls (keys, counts, per-partition sizes), and the receiver rebuilds a full-size partition
    index for the slice it got.  -> ({i: (keys, counts, length, pindex)} for the chromosomes of OTHER ranks,
    total k-mers)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    P = 1 << int(pbits)
    mine = [i for i in range(n) if owner[i] == rank]
    part = torch.arange(P, device=dev)
    perm = torch.argsort(part % world, stable=True)                 # class-major, partitions ascending inside
    ncls = [len(range(c, P, world)) for c in range(world)]          # partitions per class
    cls_poff = [0]
    for c in range(world):
        cls_poff.append(cls_poff[-1] + ncls[c])
    meta = torch.zeros(n + 1, world + 1, dtype=torch.int64, device=dev)
    regrouped = {}
    for i in mine:
        kk, cc, length = local[i]
        pidx = pindex_local[i].to(torch.int64)
        start_p, cnt_p = pidx[0::2][perm], pidx[1::2][perm]
        new_start = torch.cumsum(cnt_p, 0) - cnt_p
        total = int(kk.numel())
        csum = torch.cumsum(cnt_p, 0)
        ends = torch.stack([csum[cls_poff[c + 1] - 1] if ncls[c] else csum.new_zeros(()) for c in range(world)])
        tot_c = ends - torch.cat([ends.new_zeros(1), ends[:-1]])
        meta[i, :world] = tot_c
        meta[i, world] = int(length)
        if kk.is_cuda:      # one warp per partition run (spk_dump_regroup); the index math above is 2^pbits elements
            ns = torch.empty(P, dtype=torch.int32, device=dev)
            ns[perm] = new_start.to(torch.int32)
            rk_, rc_ = torch.empty_like(kk), torch.empty_like(cc)
            _lib.call("spk_dump_regroup", engine._p(kk), engine._p(cc), engine._p(pindex_local[i]), engine._p(ns),
                      int(pbits), engine._p(rk_), engine._p(rc_), engine._stream())
        else:               # CPU tensors (gloo tests): the same permutation with torch indexing
            src = (torch.repeat_interleave(start_p - new_start, cnt_p, output_size=total) +
                   torch.arange(total, device=dev))
            rk_, rc_ = kk[src], cc[src]
        regrouped[i] = (rk_, rc_, cnt_p.to(torch.int32))
    meta[n, 0] = int(n_kmers_local)
    dist.all_reduce(meta)
    meta_h = meta.cpu().tolist()
    tot = [[int(x) for x in meta_h[i][:world]] for i in range(n)]
    # send order: for every destination class, my chromosomes in index order
    cls_off = {i: [0] for i in mine}
    for i in mine:
        for c in range(world):
            cls_off[i].append(cls_off[i][-1] + tot[i][c])

    def pack(which, per_class_len=None):
        pieces, splits = [], []
        for d in range(world):
            m = 0
            for i in mine:
                if per_class_len is None:
                    a, b = cls_off[i][d], cls_off[i][d + 1]
                else:
                    a, b = cls_poff[d], cls_poff[d + 1]
                pieces.append(regrouped[i][which][a:b])
                m += b - a
            splits.append(m)
        dtype = {0: torch.int64, 1: torch.int32, 2: torch.int32}[which]
        return (torch.cat(pieces) if pieces else torch.empty(0, dtype=dtype, device=dev)), splits

    owned_by = [[i for i in range(n) if owner[i] == s] for s in range(world)]
    out_e = [sum(tot[i][rank] for i in owned_by[s]) for s in range(world)]
    out_p = [ncls[rank] * len(owned_by[s]) for s in range(world)]
    send_k, in_e = pack(0)
    send_c, _ = pack(1)
    send_p, in_p = pack(2, per_class_len=True)
    rk = _all_to_all(send_k, in_e, out_e, dist, dev)
    rc = _all_to_all(send_c, in_e, out_e, dist, dev)
    rp = _all_to_all(send_p, in_p, out_p, dist, dev)
    out = {}
    oe = op = 0
    my_parts = torch.arange(rank, P, world, device=dev)
    for s in range(world):
        for i in owned_by[s]:
            m = tot[i][rank]
            if s != rank:
                pc = rp[op:op + ncls[rank]].to(torch.int64)
                pidx = torch.zeros(2 * P, dtype=torch.int32, device=dev)
                pidx[2 * my_parts] = (torch.cumsum(pc, 0) - pc).to(torch.int32)
                pidx[2 * my_parts + 1] = pc.to(torch.int32)
                out[i] = (rk[oe:oe + m], rc[oe:oe + m], int(meta_h[i][world]), pidx)
            oe += m
            op += ncls[rank]
    return out, int(meta_h[n][0])


def exchange_rows_meta(dm, n_union_local, dist, dev):
    """Sizes of every rank's differential-matrix shard, union rows and fold-test passes of the whole genome (one
    all_reduce) -> ([rows per rank], n_union, n_fold_pass)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = torch.zeros(world + 2, dtype=torch.int64, device=dev)
    meta[rank] = len(dm)
    meta[world] = int(n_union_local)
    meta[world + 1] = int(dm.n_fold_pass)
    dist.all_reduce(meta)
    meta_h = [int(x) for x in meta.cpu().tolist()]
    return meta_h[:world], meta_h[world], meta_h[world + 1]


def gather_rows(dm, m, n_fold_pass, dist, dev, group=None):
    """Differential-matrix shards (rows of this rank's partitions) -> the full matrix, sorted by k-mer, on every
    rank: three all_gathers (keys, normalised rows, totals) and one key sort.  `group`: the process group the
    collectives run on (the caller issues this on a side stream with its own communicator)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ncol = dm.norm.shape[1]
    ident = list(range(world))
    kk = _gather_concat({rank: dm.keys.contiguous()}, m, ident, world, dist, dev, torch.int64, group)
    nn = _gather_concat({rank: dm.norm.contiguous().view(-1)}, [x * ncol for x in m], ident, world, dist, dev,
                        torch.float64, group)
    tt = _gather_concat({rank: dm.tot.contiguous()}, m, ident, world, dist, dev, torch.int64, group)
    keys = torch.cat([kk[r] for r in range(world)])
    norm = torch.cat([nn[r].view(-1, ncol) for r in range(world)])
    tot = torch.cat([tt[r] for r in range(world)])
    # global row order = ascending k-mer, as on one GPU (keys are < 2^63 except k = 32: use the sort kernel)
    order = engine.argsort_keys(keys, 2 * dm.k)
    return engine.DiffMatrix(keys[order].contiguous(), norm[order].contiguous(), tot[order].contiguous(), dm.k,
                             dm.labels, int(n_fold_pass))


def exchange_rows(dm, n_union_local, dist, dev):
    """exchange_rows_meta + gather_rows on the default group -> (full matrix, n_union)."""
    m, n_union, n_fold = exchange_rows_meta(dm, n_union_local, dist, dev)
    return gather_rows(dm, m, n_fold, dist, dev), n_union


def gather_sig(keys, vals, key_bits, dist, dev, max_rows=None):
    """Significant k-mers of every rank's row shard -> all of them on every rank, in rank order (the probe table the
    map stage builds from the list does not depend on the order of its keys).
    With `max_rows` (an upper bound on any rank's list, known from the row counts already exchanged) and keys of at
    most 56 bits: ONE all_gather of fixed-width rows — the subgenome id rides in the top byte of the key, unused slots
    hold -1 — and no size exchange.  Otherwise: one all_reduce of the sizes and two all_gathers."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if max_rows is not None and key_bits <= 56:
        width = max(int(max_rows), 1)
        send = torch.full((width,), -1, dtype=torch.int64, device=dev)
        n_mine = int(keys.numel())
        if n_mine:
            send[:n_mine] = keys | (vals.to(torch.int64) << 56)
        recv = torch.empty(world * width, dtype=torch.int64, device=dev)
        try:
            dist.all_gather_into_tensor(recv, send)
        except (RuntimeError, NotImplementedError, AttributeError):
            dist.all_gather(list(recv.view(world, width).unbind(0)), send)
        got = recv[recv != -1]
        return (got & ((1 << 56) - 1)).contiguous(), (got >> 56).to(torch.uint8).contiguous()
    sz = torch.zeros(world, dtype=torch.int64, device=dev)
    sz[rank] = int(keys.numel())
    dist.all_reduce(sz)
    m = [int(x) for x in sz.cpu().tolist()]
    ident = list(range(world))
    kk = _gather_concat({rank: keys.contiguous()}, m, ident, world, dist, dev, torch.int64)
    vv = _gather_concat({rank: vals.contiguous()}, m, ident, world, dist, dev, torch.uint8)
    return torch.cat([kk[r] for r in range(world)]), torch.cat([vv[r] for r in range(world)])


def exchange_windows(win_counts, n, nsg, owner, dist, dev, nw_known=None):
    """Per-chromosome window count matrices (int64 [W_i, S]) -> present on every rank (one all_gather of the per-rank
    concatenations; when the row counts are not known on every rank, one all_reduce of them first)."""
    if nw_known is not None:
        nw_h = [int(x) for x in nw_known]
    else:
        nw = torch.zeros(n, dtype=torch.int64, device=dev)
        for i, w in win_counts.items():
            nw[i] = w.shape[0]
        dist.all_reduce(nw)
        nw_h = [int(x) for x in nw.cpu().tolist()]
    flat = _gather_concat({i: w.contiguous().view(-1) for i, w in win_counts.items()}, [x * nsg for x in nw_h],
                          owner, n, dist, dev, torch.int64)
    return {i: flat[i].view(nw_h[i], nsg) for i in range(n)}


_PINNED = {}
_SCRATCH = {}

class PeerExchange:
    """The exchange step over peer memory (no collective, no host round trip per chromosome): every rank owns
    symmetric receive buffers (torch symmetric memory: peer-mapped over NVLink / NVSwitch) with one region per
    chromosome; the owner of a chromosome writes each hash-partition run of its dump straight into the region of
    the rank that merges that partition class (`spk_dump_scatter_peers`).  The stores ride the counting stream, so
    the transfer of chromosome j overlaps the packing / counting of chromosome j + 1; one all_reduce of the sizes
    at the end is both the barrier and the only host synchronisation.
    A region holds at most SPK_PEER_CAP_FRAC (default 0.12) of the chromosome's bases per world; a dump that does
    not fit is reported by `finish()` and the caller uses the collective exchange instead."""

    def __init__(self, dist, nbytes_list, pbits, lower_count, dev):
        import torch.distributed._symmetric_memory as symm
        self.dist, self.dev = dist, dev
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.n, self.pbits = len(nbytes_list), int(pbits)
        self.P = 1 << self.pbits
        self.ncls = (self.P + self.world - 1) // self.world
        frac = float(os.environ.get("SPK_PEER_CAP_FRAC", "0.12"))
        self.caps = []
        for nb in nbytes_list:
            hard = nb // max(int(lower_count), 1) + 1                 # a dumped k-mer occurs >= lower_count times
            c = int(min(hard, nb * frac) / self.world * 1.15) + 4096
            self.caps.append((c + 63) // 64 * 64)
        self.offs = [0]
        for c in self.caps:
            self.offs.append(self.offs[-1] + c)
        total = self.offs[-1]
        self.rk = symm.empty(total, dtype=torch.int64, device=dev)
        self.rc = symm.empty(total, dtype=torch.int32, device=dev)
        self.rp = symm.empty(self.n * self.ncls, dtype=torch.int32, device=dev)
        group = dist.group.WORLD
        self._handles = [symm.rendezvous(t, group) for t in (self.rk, self.rc, self.rp)]
        ptrs = [[int(p) for p in h.buffer_ptrs] for h in self._handles]
        self.peer = torch.tensor(ptrs, dtype=torch.int64).to(dev)         # [3, world] device pointer tables
        self.dst_off = torch.empty(self.P, dtype=torch.int32, device=dev)
        self.class_tot = torch.zeros(self.n, self.world, dtype=torch.int64, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int64, device=dev)

    def key(self):
        return (self.n, self.world, self.pbits, tuple(self.caps))

    def begin(self):
        self.class_tot.zero_()
        self.overflow.zero_()

    def scatter(self, i, dump):
        """dump of chromosome i (owned by this rank) -> the receive regions of all ranks (asynchronous)."""
        _lib.call("spk_dump_scatter_peers", engine._p(dump.keys), engine._p(dump.counts), engine._p(dump.pindex),
                  self.pbits, self.world, self.offs[i], self.caps[i], i * self.ncls, engine._p(self.peer[0]),
                  engine._p(self.peer[1]), engine._p(self.peer[2]), engine._p(self.dst_off),
                  engine._p(self.class_tot[i]), engine._p(self.overflow), engine._stream())

    def finish(self, owner, lengths_local, bases_local, n_kmers_local):
        """Barrier + sizes.  -> ({i: (keys, counts, length, pindex)} for the chromosomes of other ranks, total k-mers,
        bases of every chromosome), or (None, total, bases) when some dump did not fit its region."""
        n, world, rank, dev = self.n, self.world, self.rank, self.dev
        meta = torch.zeros(n + 1, world + 2, dtype=torch.int64, device=dev)
        meta[:n, :world] = self.class_tot
        for i, length in lengths_local.items():
            meta[i, world] = int(length)
            meta[i, world + 1] = int(bases_local[i])
        meta[n, 0] = int(n_kmers_local)
        meta[n, 1:2] = self.overflow
        self.dist.all_reduce(meta)        # every rank's stores precede its contribution on its stream
        meta_h = meta.cpu().tolist()
        total_kmers = int(meta_h[n][0])
        bases = [int(meta_h[i][world + 1]) for i in range(n)]
        if int(meta_h[n][1]):
            return None, total_kmers, bases
        my_parts = torch.arange(rank, self.P, world, device=dev)
        npc = int(my_parts.numel())
        out = {}
        for i in range(n):
            if owner[i] == rank:
                continue
            m = int(meta_h[i][rank])
            pc = self.rp[i * self.ncls:i * self.ncls + npc].to(torch.int64)
            pidx = torch.zeros(2 * self.P, dtype=torch.int32, device=dev)
            pidx[2 * my_parts] = (torch.cumsum(pc, 0) - pc).to(torch.int32)
            pidx[2 * my_parts + 1] = pc.to(torch.int32)
            a = self.offs[i]
            out[i] = (self.rk[a:a + m], self.rc[a:a + m], int(meta_h[i][world]), pidx)
        return out, total_kmers, bases


def _peer_exchange(dist, nbytes_list, pbits, lower_count, dev):
    """Cached PeerExchange (allocation + rendezvous are collective and cost milliseconds); None when symmetric
    memory is unavailable (then the NCCL all-to-all path is used)."""
    if os.environ.get("SPK_EXCHANGE", "peer") != "peer" or not str(dev).startswith("cuda"):
        return None
    px = _SCRATCH.get("peer")
    want = (len(nbytes_list), dist.get_world_size(), int(pbits), tuple(int(x) for x in nbytes_list), int(lower_count))
    if px is not None and px[0] == want:
        return px[1]
    _SCRATCH.pop("peer", None)
    try:
        obj = PeerExchange(dist, nbytes_list, pbits, lower_count, dev)
        ok = 1
    except Exception as exc:      # symmetric memory not available on this build / topology
        import logging
        logging.getLogger("subphaser_b200").warning("peer-memory exchange unavailable (%s): using NCCL all-to-all", exc)
        obj, ok = None, 0
    flag = torch.tensor([ok], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # all ranks take the same path
    if int(flag.item()) == 0:
        obj = None
    _SCRATCH["peer"] = (want, obj)
    return obj



def _scratch_table(max_bytes, k, lower_count, genome_max):
    """The multi-GB count scratch (partition streams, dump buffers) is kept between calls: re-allocating it
    every pass makes the caching allocator split and re-grow its largest blocks (cudaMalloc/cudaFree of GBs,
    each a device-wide synchronisation — up to 50 ms of a 500-ms pass, varying from run to run)."""
    key = (int(k), int(lower_count), int(genome_max))
    tab = _SCRATCH.get("table")
    if tab is None or tab[0] != key or tab[1].max_bases < max_bytes:
        _SCRATCH.pop("table", None)
        tab = (key, engine.CountTable(max_bytes, k, lower_count, genome_max_bases=genome_max))
        _SCRATCH["table"] = tab
    return tab[1]


def release_scratch():
    _SCRATCH.clear()
    _PINNED.clear()


_COPY_CHUNK = 32 << 20


def _copy_chunked(dst, src):
    """Asynchronous host<->device copy of a flat tensor in 32-MiB pieces.  One cudaMemcpyAsync of a whole chromosome
    (0.7 GB) or of the differential matrix (0.6 GB) occupies the copy engine of its direction for tens of
    milliseconds, and every small copy of the main stream in the same direction (kernel parameters, labels, sizes)
    queues behind it; in pieces, the small copies slip in between."""
    n = dst.numel()
    step = max(_COPY_CHUNK // max(dst.element_size(), 1), 1)
    for off in range(0, n, step):
        dst[off:off + step].copy_(src[off:off + step], non_blocking=True)


def _to_host_pinned(name, t):
    """Device tensor -> numpy through a cached pinned staging buffer (async copy; caller synchronises)."""
    n = t.numel()
    buf = _PINNED.get(name)
    if buf is None or buf.numel() < n or buf.dtype != t.dtype:
        buf = torch.empty(max(n, 1), dtype=t.dtype, pin_memory=True)
        _PINNED[name] = buf
    _copy_chunked(buf[:n], t.contiguous().view(-1))
    return buf[:n].view(t.shape)


def _host_copy_plan(name, t):
    """Like _to_host_pinned, but nothing is copied yet: -> (host view, [(dst piece, src piece), ...])."""
    n = t.numel()
    buf = _PINNED.get(name)
    if buf is None or buf.numel() < n or buf.dtype != t.dtype:
        buf = torch.empty(max(n, 1), dtype=t.dtype, pin_memory=True)
        _PINNED[name] = buf
    src = t.contiguous().view(-1)
    step = max(_COPY_CHUNK // max(src.element_size(), 1), 1)
    return buf[:n].view(t.shape), [(buf[off:min(off + step, n)], src[off:min(off + step, n)]) for off in range(0, n, step)]


class StageTimer:
    """CUDA-event timers on the launching stream, accumulated per stage name."""

    def __init__(self, enabled=True):
        self.enabled = enabled
        self.pairs = {}

    def start(self, name):
        if not self.enabled:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        self.pairs.setdefault(name, []).append((e0, e1))
        return e1

    @staticmethod
    def stop(e1):
        if e1 is not None:
            e1.record()

    def totals_ms(self):
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.pairs.items()}

    def counts(self):
        return {k: len(v) for k, v in self.pairs.items()}


def run(chrom_inputs, labels, sgs, k, lower_count=3, min_fold=2, baseline=1, ratio=1, min_freq=200,
        max_freq=10000, nsg=None, replicates=1000, max_pval=0.05, bin_size=10000, chunk_size=10_000_000,
        window_size=1_000_000, seed=0, host_inputs=False, timer=None, dist=None, owner=None,
        keep_seqs=None, return_host=False):
    """chrom_inputs: per chromosome either (device uint8 tensor, nbytes) or, with host_inputs=True,
    (pinned host uint8 tensor, nbytes) — the H2D copy then happens inside this call.
    Returns a dict of host-side results (small) and device handles."""
    engine.require_cuda()
    t = timer or StageTimer(False)
    dev = engine._dev()
    n = len(chrom_inputs)
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    if owner is None:
        owner = [0] * n
    mine = [i for i in range(n) if owner[i] == rank]
    h2d_bytes = 0
    e_run = t.start("_run")
    e_head = t.start("_head")

    # ---- K1-K3 per chromosome -----------------------------------------------------------------------
    max_bytes = max([chrom_inputs[i][1] for i in mine] + [1])
    table = _scratch_table(max_bytes, k, lower_count, max([c[1] for c in chrom_inputs] + [1]))
    seqs, dumps = {}, {}
    n_kmers = 0
    # host inputs: the H2D copy of chromosome j+1 runs on a side stream while chromosome j is counted
    copy_stream = torch.cuda.Stream() if host_inputs else None
    main_stream = torch.cuda.current_stream()

    def start_copy(i):
        buf, nbytes = chrom_inputs[i]
        d = torch.empty(nbytes + 16, dtype=torch.uint8, device=dev)
        copy_stream.wait_stream(main_stream)       # the block may have been used by earlier main-stream work
        with torch.cuda.stream(copy_stream):
            _copy_chunked(d[:nbytes], buf[:nbytes])
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d, ev

    # host inputs: the first copy cannot hide behind anything, so the smallest chromosome goes first; two copies are
    # kept in flight ahead of the chromosome being counted
    order_in = sorted(mine, key=lambda i: (chrom_inputs[i][1], i)) if host_inputs else list(mine)
    PREFETCH = 2
    pending = [start_copy(i) for i in order_in[:PREFETCH]] if host_inputs else []
    # multi-GPU: dumps travel over peer memory right after they are counted (PeerExchange)
    px = None
    if world > 1 and table.pbits > 0 and n <= 128 and os.environ.get("SPK_MATRIX_MODE", "partitioned") != "plain":
        px = _peer_exchange(dist, [c[1] for c in chrom_inputs], table.pbits, lower_count, dev)
        if px is not None:
            px.begin()
    t.stop(e_head)
    e_loop = t.start("_loop_pack_count")      # coarse brackets ("_..."): stage sums vs. whole-loop time = host gaps
    for pos, i in enumerate(order_in):
        buf, nbytes = chrom_inputs[i]
        if host_inputs:
            d, ev = pending.pop(0)
            if pos + PREFETCH < len(order_in):
                pending.append(start_copy(order_in[pos + PREFETCH]))
            e = t.start("h2d_wait")
            main_stream.wait_event(ev)
            t.stop(e)
            h2d_bytes += nbytes
        else:
            d = buf
        e = t.start("pack")
        seq = engine.pack_fasta(d, nbytes, name=labels[i], trim=True)
        t.stop(e)
        if host_inputs:
            del d
        dump = engine.count_packed(seq, k, lower_count, table=table, timer=t)
        seqs[i], dumps[i] = seq, dump
        n_kmers += dump.n_valid_kmers
        if px is not None:
            # the peer stores of this dump run on their own stream, underneath the next chromosome's pack + count
            px_stream = _SCRATCH.setdefault("px_stream", torch.cuda.Stream())
            px_stream.wait_stream(main_stream)
            with torch.cuda.stream(px_stream):
                e = t.start("scatter_side")
                if dump.pindex is not None and dump.pbits == px.pbits:
                    px.scatter(i, dump)
                else:                               # counted by the global-table fallback: no partition index
                    px.overflow += 1
                t.stop(e)
    t.stop(e_loop)
    del table          # stays alive in _SCRATCH for the next call

    # ---- exchange: every rank needs every dump (exact merge of the global k-mer table) -------------
    bases_all = None
    got = None
    if world > 1 and px is not None:
        e = t.start("exchange")
        main_stream.wait_stream(_SCRATCH["px_stream"]) if "px_stream" in _SCRATCH else None
        e2 = t.start("_x_wait")                      # the sizes' all_reduce doubles as the barrier with the slowest rank
        got, n_kmers_total, bases_all = px.finish(owner, {i: dumps[i].length for i in mine},
                                                  {i: seqs[i].n_bases for i in mine}, n_kmers)
        t.stop(e2)
        if got is not None:
            for i, (kk, cc, length, pidx) in got.items():
                dumps[i] = engine.KmerDump(kk, cc, k, length, 0, 0, labels[i], None, pidx, px.pbits)
            by_class = True
        t.stop(e)
    if world > 1 and got is None:
        e = t.start("exchange")
        e2 = t.start("_x_wait")                      # time until the slowest rank has finished counting
        dist.barrier()
        t.stop(e2)
        local = {i: (dumps[i].keys, dumps[i].counts, dumps[i].length) for i in mine}
        # same partition bits everywhere (the table is sized for the largest chromosome of the genome)?
        pb = torch.tensor([dumps[i].pbits if dumps[i].pindex is not None else -1 for i in mine] or [-2],
                          dtype=torch.int64, device=dev)
        pmeta = torch.stack([pb.max(), -pb.min()])
        dist.all_reduce(pmeta, op=dist.ReduceOp.MAX)
        pb_max, pb_min = int(pmeta[0].item()), -int(pmeta[1].item())
        by_class = (pb_max == pb_min and pb_max > 0 and n <= 128 and
                    os.environ.get("SPK_EXCHANGE", "peer") in ("class", "peer") and
                    os.environ.get("SPK_MATRIX_MODE", "partitioned") != "plain")
        e2 = t.start("_x_dumps")
        if by_class:
            # each rank only needs the partition class it will merge: 1/world of every foreign dump
            got, n_kmers_total = exchange_dumps_by_class(local, {i: dumps[i].pindex for i in mine}, pb_max, n, owner,
                                                         dist, dev, n_kmers)
            for i, (kk, cc, length, pidx) in got.items():
                dumps[i] = engine.KmerDump(kk, cc, k, length, 0, 0, labels[i], None, pidx, pb_max)
            t.stop(e2)
        else:
            everything, n_kmers_total = exchange_dumps(local, n, owner, dist, dev, n_kmers)
            t.stop(e2)
            e2 = t.start("_x_pindex")
            pidx, pbits = exchange_pindex({i: dumps[i].pindex for i in mine}, n, owner, dist, dev,
                                          dumps[mine[0]].pbits if mine else 0)
            t.stop(e2)
            for i in range(n):
                if owner[i] != rank:
                    kk, cc, length = everything[i]
                    dumps[i] = engine.KmerDump(kk, cc, k, length, 0, 0, labels[i], None, pidx.get(i), pbits)
        t.stop(e)
    elif world == 1:
        n_kmers_total = n_kmers
        by_class = False
    dump_list = [dumps[i] for i in range(n)]

    # ---- K3b/K4 matrix + filter: rows are independent -> each rank builds the rows of its share ----------
    dm = None
    use_p = engine.can_pmatrix(dump_list)
    if use_p:
        # partitioned union + filter in shared memory (rank r: partitions p % world == r)
        e = t.start("matrix")
        try:
            dm, n_union = engine.pmatrix_filter(dump_list, sgs, labels, min_fold=min_fold, baseline=baseline,
                                                ratio=ratio, min_freq=min_freq, max_freq=max_freq, nparts=world,
                                                part=rank,
                                                full_dumps=[owner[i] == rank or not by_class for i in range(n)])
        except OverflowError:          # a partition did not fit the shared-memory table (adversarial skew)
            dm = None
        t.stop(e)
    if world > 1 and use_p:            # every rank must take the same path: the row shards differ between them
        ok = torch.tensor([0 if dm is None else 1], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            dm = None
    if dm is None:
        if world > 1 and by_class:
            raise OverflowError("a hash partition overflowed the shared-memory union table and the ranks hold only "
                                "their partition class of the dumps: rerun with SPK_EXCHANGE=gather")
        e = t.start("matrix")
        cm = engine.build_matrix(dump_list, labels, nparts=world, part=rank)
        t.stop(e)
        e = t.start("filter")
        dm = engine.filter_matrix(cm, sgs, labels, min_fold=min_fold, baseline=baseline, ratio=ratio,
                                  min_freq=min_freq, max_freq=max_freq)
        t.stop(e)
        n_union = len(cm)
        del cm
    shard = dm          # world > 1: this rank's rows (its partition class), ascending k-mer
    side = _SCRATCH.setdefault("side_stream", torch.cuda.Stream())
    if world > 1:
        e = t.start("exchange")
        e2 = t.start("_x_rows")
        m_rows, n_union, n_fold_all = exchange_rows_meta(shard, n_union, dist, dev)
        M = sum(m_rows)
        t.stop(e2)
        t.stop(e)
    else:
        M = len(dm)
    if M == 0:
        raise ValueError("0 kmer remained after filtering. Please reset the filter options.")
    # Side work is ENQUEUED where the host would otherwise wait for the main stream (just before a stage's result is
    # read back): enqueuing it first kept the main stream idle for the ~1 ms of host time its launches take.
    ev_rows = torch.cuda.Event()
    ev_rows.record()                                   # the row shard (world 1: the matrix) is final here
    host_copy = None
    copy_pieces = []
    return_host = bool(return_host) and rank == 0     # one copy of the results leaves the node, not one per rank

    def enqueue_gather():
        nonlocal dm
        if world > 1:
            # Only the report (bootstrap, PCA input, the matrix handed back to the host) needs every row on every
            # rank: the 0.6-GB all-gather + sort runs on the side stream over its own communicator, underneath the
            # Gram pass, the t-test and the map stage, which work on the row shards.
            side_pg = _SCRATCH.get("side_pg")
            if side_pg is None or side_pg[0] != world:
                side_pg = (world, dist.new_group(ranks=list(range(world))))
                _SCRATCH["side_pg"] = side_pg
            side.wait_event(ev_rows)
            with torch.cuda.stream(side):
                e_g = t.start("gather_rows_side")
                dm = gather_rows(shard, m_rows, n_fold_all, dist, dev, group=side_pg[1])
                t.stop(e_g)
                ev_rows.record()                       # (now: the FULL matrix is final)

    def enqueue_host_copy():
        # The device->host copy of the differential matrix (0.6 GB for wheat, ~13 ms of PCIe) runs on its own stream
        # underneath the map stage — and only there: while it is in flight every small result read of the main
        # stream (labels, sizes, masks: each a host synchronisation) arrives ~10 ms late, whichever engine moves the
        # bulk, so it is started after the last of those reads.
        # The pieces are handed to the copy engine a few at a time from inside the map loop (pump_host_copy): issuing
        # all of them at once kept the host in the driver for the duration of the transfer, with nothing queued on the
        # main stream.
        nonlocal host_copy, copy_pieces
        if return_host:
            d2h_stream = _SCRATCH.setdefault("d2h_stream", torch.cuda.Stream())
            d2h_stream.wait_event(ev_rows)
            hk, pk = _host_copy_plan("dm_keys", dm.keys)
            hn, pn = _host_copy_plan("dm_norm", dm.norm)
            host_copy = (hk, hn)
            copy_pieces = pk + pn

    def pump_host_copy(k):
        if copy_pieces:
            with torch.cuda.stream(_SCRATCH["d2h_stream"]):
                for _ in range(min(k, len(copy_pieces))):
                    dst, src = copy_pieces.pop(0)
                    dst.copy_(src, non_blocking=True)

    # ---- K5-K8 cluster ---------------------------------------------------------------------------------
    if nsg is None:
        nsg = max(len(sg) for sg in sgs)
    e = t.start("cluster")
    if world > 1:
        # the Gram pass (one stream over all rows) runs on the row shards and the n x n partial sums are all-reduced
        G = (engine.gram(engine.zscore_rows(shard.norm)) if len(shard) else
             torch.zeros(n, n, dtype=torch.float64, device=dev))
        dist.all_reduce(G)
        Z = None
    else:
        Z = engine.zscore_rows(dm.norm)
        G = engine.gram(Z)
    # (device copies of small host lists are made here, on the main stream: a pageable host->device copy issued on
    # the side stream would block the host until the side stream reaches it, and the main stream would idle)
    order = torch.tensor([i for _, i in sorted(zip(labels, range(n)))], dtype=torch.int32, device=dev)
    lab_full, inertia = engine.kmeans_gram(G, nsg, order=order, seed=seed)
    ev_lab = torch.cuda.Event()
    ev_lab.record()
    enqueue_gather()
    lab_full_h = lab_full[0].cpu().numpy()
    t.stop(e)
    # The bootstrap (replicates x tiny K-Means problems: a few dozen thread blocks for milliseconds) and the PCA (one
    # thread block) only feed the report; nothing downstream waits for them.  They run on a side stream underneath the
    # t-test, the table build and the map kernels and are joined at the end.
    R = int(replicates)
    lab_b = ari = vm = None
    eig = scores = pratio = None

    def enqueue_bootstrap():
        nonlocal Z, lab_b, ari, vm, eig, scores, pratio
        side.wait_event(ev_lab)
        with torch.cuda.stream(side):
            e_side = t.start("bootstrap_pca_side")
            if Z is None and R > 0:
                Z = engine.zscore_rows(dm.norm)
            if R > 0:
                d_idx = engine.resample_indices(M, R, seed)        # same generator state on every rank: same plan
                if world > 1:
                    # the replicates are independent: rank r clusters replicates [lo, hi), the labels are all-gathered
                    per = (R + world - 1) // world
                    lo, hi = min(rank * per, R), min((rank + 1) * per, R)
                    lab_pad = torch.zeros(per, n, dtype=torch.int32, device=dev)
                    if hi > lo:
                        Gb = engine.gram_batched(Z, d_idx[lo:hi].contiguous())
                        lab_loc, _ = engine.kmeans_gram(Gb, nsg, order=order, seed=seed + 1, r0=lo)
                        lab_pad[:hi - lo] = lab_loc
                    lab_all = torch.empty(world * per, n, dtype=torch.int32, device=dev)
                    dist.all_gather_into_tensor(lab_all.view(-1), lab_pad.view(-1), group=_SCRATCH["side_pg"][1])
                    lab_b = lab_all[:R].contiguous()
                else:
                    Gb = engine.gram_batched(Z, d_idx)
                    lab_b, _ = engine.kmeans_gram(Gb, nsg, order=order, seed=seed + 1)
                ari, vm = engine.cluster_scores(lab_full[0], lab_b)
            eig, scores, pratio = engine.pca_gram(G, min(nsg, n))
            t.stop(e_side)

    e = t.start("ttest")
    if len(shard):
        best, pval, means = engine.ttest_groups(shard.norm, lab_full_h.tolist(), nsg)
        enqueue_bootstrap()
        keep = ~(pval > max_pval)
        sig_keys = shard.keys[keep].contiguous()
        sig_vals = best[keep].to(torch.uint8).contiguous()
    else:
        enqueue_bootstrap()
        sig_keys = torch.empty(0, dtype=torch.int64, device=dev)
        sig_vals = torch.empty(0, dtype=torch.uint8, device=dev)
    if world > 1:
        sig_keys, sig_vals = gather_sig(sig_keys, sig_vals, 2 * k, dist, dev, max_rows=max(m_rows))
    t.stop(e)

    # ---- K9 map ----------------------------------------------------------------------------------------
    e = t.start("sigtable")
    sig = engine.SigTable(sig_keys, sig_vals, k, track_hits=False, S=nsg)
    t.stop(e)
    win_counts = {}
    enqueue_host_copy()
    per_iter = (len(copy_pieces) + max(len(mine), 1) - 1) // max(len(mine), 1)
    e_loop = t.start("_loop_map_stack")
    for i in mine:
        e = t.start("map")
        lines, _ = engine.map_bins(seqs[i], sig, nsg, bin_size, chunk_size, sync=False)   # (no host read in this loop)
        t.stop(e)
        pump_host_copy(per_iter)
        # ---- stack lines into windows (Circos.stack_matrix) on the device ----
        e = t.start("stack")
        L = seqs[i].n_bases
        nl = lines.shape[0]
        nwin = int(((max(L - 1, 0) // bin_size) * bin_size) // window_size) + 1 if L else 0
        out = torch.zeros(max(nwin, 1), nsg, dtype=torch.int64, device=dev)
        if nl:
            _lib.call("spk_stack_lines", engine._p(lines), nl, nsg, k, int(bin_size), int(chunk_size),
                      int(window_size), L, engine._p(out), nwin, engine._stream())
        win_counts[i] = out[:nwin]
        t.stop(e)
    t.stop(e_loop)
    pump_host_copy(len(copy_pieces))
    if keep_seqs is not None:
        keep_seqs.update(seqs)

    # ---- gather windows, K10 ---------------------------------------------------------------------------
    if world > 1:
        e = t.start("exchange")
        e2 = t.start("_x_windows")
        nw_known = None
        if bases_all is not None:      # window counts per chromosome follow from its length: no size exchange
            nw_known = [int(((max(L - 1, 0) // bin_size) * bin_size) // window_size) + 1 if L else 0 for L in bases_all]
        win_counts = exchange_windows(win_counts, n, nsg, owner, dist, dev, nw_known)
        t.stop(e2)
        t.stop(e)
    e = t.start("enrich")
    allw = torch.cat([win_counts[i] for i in range(n)], dim=0)
    nz = allw.any(dim=1)                      # zero-hit windows are absent in the reference (Circos.py:737)
    allw = allw[nz].contiguous()
    enr = engine.fisher_enrich(allw, max_pval=max_pval)
    t.stop(e)
    e_tail = t.start("_tail")
    side.synchronize()
    d_bs = None
    if lab_b is not None:
        lab_b_h = lab_b.cpu().numpy()
        d_bs = [int(100 * int(np.sum(lab_b_h[:, i] == lab_full_h[i])) / R) for i in range(n)]
    d2h_bytes = int(dm.norm.numel() * 8 + dm.keys.numel() * 8 + allw.numel() * 8 * 4)
    pca_host = (scores.cpu().numpy(), pratio.cpu().numpy())
    t.stop(e_tail)
    t.stop(e_run)
    return dict(n_kmers=n_kmers_total, n_kmers_local=n_kmers, n_union=n_union, n_diff=M, n_sig=int(sig_keys.numel()),
                n_windows=int(allw.shape[0]), labels_full=lab_full_h.tolist(), d_bs=d_bs,
                lengths=[d.length for d in dump_list], enrich=enr, dm=dm, window_counts=allw,
                pca=pca_host, bootstrap_scores=(ari, vm),      # per-replicate ARI / V-measure against the full run (device)
                h2d_bytes=h2d_bytes, d2h_bytes=d2h_bytes if return_host else int(allw.numel() * 8 * 4),
                matrix_host=_matrix_host(host_copy) if return_host else None)


def _matrix_host(host_copy):
    _SCRATCH["d2h_stream"].synchronize()
    k, v = host_copy
    return k.numpy().view(np.uint64), v.numpy()
