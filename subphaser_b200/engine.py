"""Device-side engine: thin Python orchestration of the libspk kernels.

torch is used for three things only: allocating device/pinned buffers, naming the current CUDA stream
and (in hotpath.py / bench.py) bootstrapping NCCL through torch.distributed.  All arithmetic happens in libspk.so.  Nothing here falls back
to the CPU: without a CUDA device every entry point raises.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import call


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.SpkError("no CUDA device visible: subphaser_b200 has no CPU fallback")
    _lib.load()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _empty(n, dtype):
    return torch.empty(int(n), dtype=dtype, device=_dev())


def _zeros(n, dtype):
    return torch.zeros(int(n), dtype=dtype, device=_dev())


def u64_numpy(t):
    """int64 device tensor holding uint64 payloads -> numpy uint64 on the host."""
    return t.detach().cpu().numpy().view(np.uint64)


# ----------------------------------------------------------------------------------------------------
# K1: FASTA -> packed sequence
# ----------------------------------------------------------------------------------------------------
class PackedSeq:
    """A chromosome resident in HBM: 2-bit codes + validity bits (include/spk.h K1 layout)."""

    def __init__(self, packed, valid, n_bases, n_valid, n_records, name=None):
        self.packed, self.valid = packed, valid
        self.n_bases, self.n_valid, self.n_records = int(n_bases), int(n_valid), int(n_records)
        self.name = name

    def nbytes(self):
        return self.packed.numel() * 4 + self.valid.numel() * 4


def read_fasta_bytes(path):
    """File -> uint8 numpy array (gzip transparently, as the reference's `zcat`, Jellyfish.py:696)."""
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        import gzip
        with gzip.open(path, "rb") as f:
            return np.frombuffer(f.read(), dtype=np.uint8)
    return np.fromfile(path, dtype=np.uint8)


def to_device_bytes(buf):
    """host uint8 array / bytes -> device uint8 tensor (padded to 16 B)."""
    require_cuda()
    if isinstance(buf, (bytes, bytearray, memoryview)):
        buf = np.frombuffer(buf, dtype=np.uint8)
    n = int(buf.size)
    d = _empty(n + 16, torch.uint8)
    if n:
        d[:n].copy_(torch.from_numpy(np.ascontiguousarray(buf)), non_blocking=False)
    return d, n


def pack_fasta(d_ascii, nbytes, name=None, trim=True):
    """K1 on device bytes.  `trim` re-allocates the outputs to the exact base count."""
    require_cuda()
    lib = _lib.load()
    cap = max(int(nbytes), 1)
    packed = _empty(lib.spk_packed_words(cap), torch.int32)
    valid = _empty(lib.spk_valid_words(cap), torch.int32)
    ws_bytes = lib.spk_pack_workspace_bytes(nbytes)
    ws = _empty(ws_bytes, torch.uint8)
    info = _zeros(4, torch.int64)
    call("spk_pack_fasta", _p(d_ascii), nbytes, _p(packed), _p(valid), cap, _p(info), _p(ws), ws_bytes,
         _stream())
    n_bases, n_valid, n_rec, path = (int(x) for x in info.cpu().tolist())
    if trim:
        pw, vw = lib.spk_packed_words(max(n_bases, 1)), lib.spk_valid_words(max(n_bases, 1))
        if pw < 0.9 * packed.numel():          # (a chromosome file is ~98 % bases: not worth a 0.26-GB copy)
            packed = packed[:pw].clone()
            valid = valid[:vw].clone()
        else:
            packed, valid = packed[:pw], valid[:vw]
    del ws
    seq = PackedSeq(packed, valid, n_bases, n_valid, n_rec, name)
    seq.pack_path = ("regular", "3pass", "single")[path]       # which K1 kernel produced it (diagnostic)
    return seq


# ----------------------------------------------------------------------------------------------------
# K2/K3: count -> dump
# ----------------------------------------------------------------------------------------------------
class KmerDump:
    """`jellyfish dump -c -L` content of one chromosome, resident on the device."""

    def __init__(self, keys, counts, k, length, n_valid_kmers, n_distinct, name=None, histo=None,
                 pindex=None, pbits=0):
        self.keys, self.counts, self.k = keys, counts, int(k)
        self.pindex, self.pbits = pindex, int(pbits)   # partition index of the dump (spk_pcount_canonical_ex)
        self.length = int(length)              # sum of dumped counts == lengths[i] (Jellyfish.py:97)
        self.n_valid_kmers = int(n_valid_kmers)
        self.n_distinct = int(n_distinct)
        self.name = name
        self.histo = histo

    def __len__(self):
        return int(self.keys.numel())

    def to_host(self):
        return u64_numpy(self.keys), self.counts.cpu().numpy().view(np.uint32)


class CountTable:
    """Reusable device scratch for counting chromosomes of up to `max_bases` bases.

    mode "partitioned" (default): spk_pcount_canonical, tables stay in L2 (spk_pcount.cu).
    mode "global": the v1 single open-addressed table in HBM (spk_count.cu); used for chromosomes of
    2^32 bases or more, or when SPK_COUNT_MODE=global."""

    def __init__(self, max_bases, k, lower_count=1, mode=None, common_pbits=True, genome_max_bases=None):
        require_cuda()
        lib = _lib.load()
        self.k = int(k)
        self.max_bases = int(max_bases)
        self.lower_count = max(int(lower_count), 1)
        if mode is None:
            mode = os.environ.get("SPK_COUNT_MODE", "partitioned")
        if self.max_bases >= 2**32 - 1:
            mode = "global"
        self.mode = mode
        self.stats = _zeros(8, torch.int64)
        self.pbits = 0
        if mode == "partitioned":
            # common_pbits: every chromosome counted with this table uses the partition bits of the
            # largest one, and its dump carries a partition index (-> engine.pmatrix_filter)
            # genome_max_bases: the largest chromosome of the whole genome (other ranks included)
            gmax = max(self.max_bases, int(genome_max_bases or 0))
            self.pbits = lib.spk_pcount_pbits(gmax, self.k) if (common_pbits and gmax < 2**32 - 1) else 0
            self.ws_bytes = lib.spk_pcount_workspace_bytes_ex(self.max_bases, self.k, self.pbits)
            self.ws = _empty(self.ws_bytes, torch.uint8)
            self.cap = self.max_bases // self.lower_count + 1024
            self.out_keys = _empty(self.cap, torch.int64)
            self.out_counts = _empty(self.cap, torch.int32)
        else:
            self.layout = lib.spk_count_layout(self.max_bases, self.k)
            self.table_bytes = lib.spk_count_table_bytes(self.max_bases, self.k)
            self.table = _empty(self.table_bytes, torch.uint8)
            nb = lib.spk_table_scan_blocks()
            self.block_counts = _empty(3 * nb + 2, torch.int32)

    def slots(self):
        return _lib.load().spk_count_table_slots(self.table_bytes, self.layout)


def count_packed(seq, k, lower_count, table=None, histo_len=0, timer=None):
    """K2 + K3 on one packed chromosome -> KmerDump.  `timer` (hotpath.StageTimer) brackets the
    stages with CUDA events on the launching stream."""
    require_cuda()
    if (table is None or table.max_bases < seq.n_bases or table.k != k
            or (table.mode == "partitioned" and table.lower_count > max(int(lower_count), 1))):
        table = CountTable(max(seq.n_bases, 1), k, lower_count)
    st = _stream()
    tick = (lambda name: timer.start(name)) if timer is not None else (lambda name: None)
    tock = (lambda e: timer.stop(e)) if timer is not None else (lambda e: None)
    histo = _zeros(histo_len, torch.int64) if histo_len else None
    if table.mode == "partitioned":
        e = tick("count")
        pindex = _empty(2 << table.pbits, torch.int32) if table.pbits else None
        call("spk_pcount_canonical_ex", _p(seq.packed), _p(seq.valid), seq.n_bases, k, lower_count, _p(table.ws),
             table.ws_bytes, _p(table.out_keys), _p(table.out_counts), table.cap, _p(table.stats), _p(histo),
             histo_len, table.pbits, _p(pindex), st)
        tock(e)
        n_valid, n_fail, _, _, distinct, n_ge, sum_ge, _ = (int(x) for x in table.stats.cpu().tolist())
        if n_fail or os.environ.get("SPK_PCOUNT_FORCE_FAIL") == "1":      # (the variable is a test hook)
            # a hash partition did not fit its shared-memory table (low-complexity / adversarial input): count this
            # chromosome with the global open-addressed table instead.  Its dump has no partition index, so the
            # genome's union goes through the plain path (spk_union_insert) — slower, never wrong.
            import logging
            logging.getLogger("subphaser_b200").warning(
                "partitioned counter: %d inserts did not fit a shared-memory partition table; "
                "recounting %s with the global table", n_fail, seq.name or "chromosome")
            gtab = CountTable(max(seq.n_bases, 1), k, lower_count, mode="global")
            return count_packed(seq, k, lower_count, table=gtab, histo_len=histo_len, timer=timer)
        if n_ge > table.cap:
            raise OverflowError("dump capacity exceeded: %d > %d" % (n_ge, table.cap))
        e = tick("scan")
        keys = table.out_keys[:n_ge].clone()
        counts = table.out_counts[:n_ge].clone()
        tock(e)
        return KmerDump(keys, counts, k, sum_ge, n_valid, distinct, seq.name, histo, pindex, table.pbits)
    e = tick("table_init")
    call("spk_count_table_init", _p(table.table), table.table_bytes, k, table.layout, st)
    table.stats.zero_()
    tock(e)
    e = tick("count")
    call("spk_count_canonical", _p(seq.packed), _p(seq.valid), seq.n_bases, k, _p(table.table),
         table.table_bytes, table.layout, _p(table.stats), st)
    tock(e)
    e = tick("scan")
    call("spk_table_stats", _p(table.table), table.table_bytes, k, table.layout, lower_count,
         _p(table.stats[4:]), _p(table.block_counts), _p(histo), histo_len, st)
    n_valid, n_fail, _, _, distinct, n_ge, sum_ge, _ = (int(x) for x in table.stats.cpu().tolist())
    if n_fail:
        raise OverflowError("k-mer table full: %d inserts failed" % n_fail)
    keys = _empty(n_ge, torch.int64)
    counts = _empty(n_ge, torch.int32)
    if n_ge:
        call("spk_table_extract", _p(table.table), table.table_bytes, k, table.layout, lower_count,
             _p(table.block_counts), _p(keys), _p(counts), n_ge, st)
    tock(e)
    return KmerDump(keys, counts, k, sum_ge, n_valid, distinct, seq.name, histo)


# ----------------------------------------------------------------------------------------------------
# K3b/K4: union -> count matrix -> differential filter
# ----------------------------------------------------------------------------------------------------
class CountMatrix:
    """kmer -> [count per chromosome] (the reference's d_mat, Jellyfish.py:439-460), on the device."""

    def __init__(self, matrix, row_keys, lengths, k, labels=None):
        self.matrix = matrix          # int32 [U, n] (uint32 payload)
        self.row_keys = row_keys      # int64 [U]   (uint64 payload)
        self.lengths = list(lengths)
        self.k = int(k)
        self.labels = labels

    def __len__(self):
        return int(self.row_keys.numel())

    @property
    def ncol(self):
        return len(self.lengths)


def build_matrix(dumps, labels=None, nparts=1, part=0):
    """JellyfishDumps.to_matrix on the device: union of the dumped k-mers, one column per chromosome.
    nparts/part: build only the rows whose k-mer hashes into partition `part` (multi-GPU row sharding)."""
    require_cuda()
    st = _stream()
    n = len(dumps)
    total = sum(len(d) for d in dumps)
    uslots = max(int(total / 0.5 / nparts * (1.1 if nparts > 1 else 1.0)) + 1024, 2048)   # >= 2 x #distinct
    ukeys = torch.full((uslots,), -1, dtype=torch.int64, device=_dev())
    urows = _empty(uslots, torch.int32)
    nrows = _zeros(1, torch.int32)
    fail = _zeros(1, torch.int64)
    for d in dumps:
        call("spk_union_insert", _p(d.keys), len(d), _p(ukeys), _p(urows), uslots, _p(nrows), _p(fail),
             nparts, part, st)
    U = int(nrows.item())
    if int(fail.item()):
        raise OverflowError("union table full")
    matrix = _zeros(U * n, torch.int32).view(U, n) if U else _zeros(0, torch.int32).view(0, n)
    row_keys = _empty(U, torch.int64)
    for i, d in enumerate(dumps):
        call("spk_matrix_fill", _p(d.keys), _p(d.counts), len(d), _p(ukeys), _p(urows), uslots, _p(matrix),
             _p(row_keys), n, i, nparts, part, st)
    k = dumps[0].k if dumps else 0
    return CountMatrix(matrix, row_keys, [d.length for d in dumps], k, labels)


class DiffMatrix:
    """The differential-k-mer matrix (rows sorted by k-mer): what `.kmer.mat` holds."""

    def __init__(self, keys, norm, tot, k, labels, n_fold_pass, fold_tots=None):
        self.keys, self.norm, self.tot = keys, norm, tot   # int64 [M], float64 [M, n], int64 [M]
        self.k, self.labels = int(k), labels
        self.n_fold_pass = int(n_fold_pass)
        self.fold_tots = fold_tots                         # totals of all fold-pass k-mers (histogram)

    def __len__(self):
        return int(self.keys.numel())


class FoldHistogram:
    """What `plot_histogram(tot_freqs, ...)` (Jellyfish.py:511,650-666) needs of the totals of all fold-passing k-mers —
    matplotlib's histogram with bins of `step` occurrences and the 99th percentile that limits the x axis — computed on
    the device from the filter's outputs; the list of totals (10^8 entries for wheat) is never copied to the host."""

    def __init__(self, tot, flags, U, step=25, xlim=99):
        lib = _lib.load()
        st = _stream()
        self._tot, self._flags, self._U = tot, flags, int(U)
        mm = _zeros(3, torch.int64)
        call("spk_tot_minmax", _p(tot), _p(flags), self._U, _p(mm), st)
        self.n, mn, mx = (int(x) for x in mm.cpu().numpy().view(np.uint64).tolist())
        self.step, self.xlim_pct = step, xlim
        if self.n == 0:
            self.min = self.max = 0
            self.nbins, self.hist, self.edges, self.xlim = 0, np.zeros(0, np.int64), np.zeros(1), 0.0
            return
        self.min, self.max = mn, mx
        self.nbins = max(int((mx - 0) / step), 1)               # `nbins = int((_max-_min)/step)` with _min = 0 (:653-654)
        hist = _zeros(self.nbins, torch.int64)
        call("spk_tot_histogram", _p(tot), _p(flags), self._U, float(mn), float(mx), self.nbins, _p(hist), st)
        self.hist = hist.cpu().numpy()
        self.edges = np.linspace(float(mn), float(mx), self.nbins + 1)
        self.xlim = self.percentile(xlim)

    def _order_stat(self, rank):
        """rank-th smallest total (0-based): radix select, 16 bits per device pass"""
        prefix, hist = 0, _zeros(65536, torch.int64)
        for shift in (48, 32, 16, 0):
            if (self.max >> shift) == 0 and shift:
                continue                                         # every value is zero in these bits
            call("spk_tot_select_pass", _p(self._tot), _p(self._flags), self._U, shift, prefix, _p(hist), _stream())
            h = hist.cpu().numpy()
            c = np.cumsum(h)
            b = int(np.searchsorted(c, rank, side="right"))
            rank -= int(c[b - 1]) if b else 0
            prefix = (prefix << 16) | b
        return prefix

    def percentile(self, q):
        """np.percentile(data, q) ('linear'): interpolation between two order statistics, numpy's _lerp arithmetic."""
        pos = (q / 100.0) * (self.n - 1)
        lo = int(np.floor(pos))
        t = pos - lo
        a = float(self._order_stat(lo))
        b = float(self._order_stat(min(lo + 1, self.n - 1)))
        d = b - a
        return float(b - d * (1 - t)) if t >= 0.5 else float(a + d * t)


def flatten_sgs(sgs, labels):
    """sgs (list of sets -> list of groups -> list of labels) -> CSR int32 arrays of column indices."""
    col = {lab: i for i, lab in enumerate(labels)}
    set_off, grp_off, members = [0], [0], []
    for sg in sgs:
        for chrs in sg:
            for c in chrs:
                if c not in col:
                    raise KeyError(c)
                members.append(col[c])
            grp_off.append(len(members))
        set_off.append(len(grp_off) - 1)
    return (np.array(set_off, np.int32), np.array(grp_off, np.int32), np.array(members, np.int32))


def filter_matrix(cm, sgs, labels, min_fold=2, baseline=1, ratio=1, min_freq=200, max_freq=10000,
                  by_count=False, want_fold_tots=False):
    """JellyfishDumps.filter arithmetic (Jellyfish.py:462-512) on the device -> DiffMatrix."""
    require_cuda()
    st = _stream()
    U, n = len(cm), cm.ncol
    set_off, grp_off, members = flatten_sgs(sgs, labels)
    d_set, d_grp, d_mem = (torch.from_numpy(a).to(_dev()) for a in (set_off, grp_off, members))
    d_len = torch.tensor(cm.lengths, dtype=torch.int64, device=_dev())
    flags = _empty(max(U, 1), torch.uint8)
    tot = _empty(max(U, 1), torch.int64)
    counters = _zeros(4, torch.int64)
    call("spk_filter_differential", _p(cm.matrix), U, n, _p(d_len), _p(d_set), len(set_off) - 1, _p(d_grp),
         len(grp_off) - 1, _p(d_mem), len(members), float(min_fold), int(baseline), int(bool(by_count)), float(ratio),
         float(min_freq), float(max_freq), _p(flags), _p(tot), _p(counters), st)
    n_fold, n_keep = (int(x) for x in counters[:2].cpu().tolist())
    scan = _empty(U + 2 + U // 8192 + 1, torch.int32)
    keys = _empty(n_keep, torch.int64)
    rows = _empty(n_keep, torch.int32)
    call("spk_filter_select", _p(cm.row_keys), _p(flags), U, _p(scan), _p(keys), _p(rows), n_keep, st)
    if n_keep > 1:   # deterministic row order: ascending k-mer
        lib = _lib.load()
        ws_bytes = lib.spk_sort_workspace_bytes(n_keep)
        ws = _empty(ws_bytes, torch.uint8)
        kt, rt = _empty(n_keep, torch.int64), _empty(n_keep, torch.int32)
        call("spk_sort_pairs_u64", _p(keys), _p(rows), _p(kt), _p(rt), n_keep, 2 * cm.k, _p(ws), ws_bytes, st)
    norm = _empty(n_keep * n, torch.float64).view(n_keep, n)
    otot = _empty(n_keep, torch.int64)
    call("spk_filter_emit", _p(cm.matrix), _p(tot), _p(rows), n_keep, n, _p(d_len), _p(norm), _p(otot), st)
    fold_tots = None
    if want_fold_tots and U:
        fold_tots = FoldHistogram(tot, flags, U)          # histogram on the device; the totals stay there
    return DiffMatrix(keys, norm, otot, cm.k, labels, n_fold, fold_tots)


_PM_CAP = {}


def can_pmatrix(dumps):
    """True when every dump carries a partition index with the same partition bits."""
    # (<= 128 columns: a shared-memory table slot is an 8-byte key + 4 bytes per chromosome)
    return (len(dumps) > 0 and len(dumps) <= 128 and all(d.pindex is not None and d.pbits > 0 for d in dumps)
            and len({d.pbits for d in dumps}) == 1 and os.environ.get("SPK_MATRIX_MODE", "partitioned") != "plain")


def pmatrix_union_size(dumps):
    """len(d_mat) of JellyfishDumps.to_matrix without building it: union rows counted per hash partition."""
    require_cuda()
    ptrs = torch.tensor([[d.keys.data_ptr() for d in dumps], [d.counts.data_ptr() for d in dumps],
                         [d.pindex.data_ptr() for d in dumps]], dtype=torch.int64).to(_dev())
    counters = _zeros(8, torch.int64)
    call("spk_pmatrix_filter", _p(ptrs[0]), _p(ptrs[1]), _p(ptrs[2]), len(dumps), dumps[0].pbits, 1, 0, None,
         None, 0, None, 0, None, 0, 0.0, 0, 0, 0.0, 0.0, 0.0, None, None, 0, sum(len(d) for d in dumps),
         _p(counters), _stream())
    n_union, _, _, n_over = (int(x) for x in counters[:4].cpu().tolist())
    if n_over or os.environ.get("SPK_PMATRIX_FORCE_OVERFLOW") == "1":     # (the variable is a test hook)
        raise OverflowError("partition table overflow in spk_pmatrix_filter")
    return n_union


class LazyUnion:
    """d_mat of the partitioned path: the dumps themselves; the union is merged partition by partition inside
    spk_pmatrix_filter and never materialised.  len() = number of distinct dumped k-mers (Jellyfish.py:417)."""

    def __init__(self, dumps, labels=None):
        self.dumps, self.labels = list(dumps), labels
        self.lengths = [d.length for d in dumps]
        self.k = dumps[0].k
        self._len = None

    def __len__(self):
        if self._len is None:
            self._len = pmatrix_union_size(self.dumps)
        return self._len


def pmatrix_filter(dumps, sgs, labels, min_fold=2, baseline=1, ratio=1, min_freq=200, max_freq=10000,
                   by_count=False, want_fold_tots=False, nparts=1, part=0, lengths=None, full_dumps=None):
    """to_matrix + filter in one partition-by-partition pass (spk_pmatrix_filter) -> (DiffMatrix, n_union).
    Rows come back sorted by k-mer, normalised by `lengths` exactly like filter_matrix."""
    require_cuda()
    st = _stream()
    n = len(dumps)
    k, pbits = dumps[0].k, dumps[0].pbits
    lengths = [d.length for d in dumps] if lengths is None else list(lengths)
    set_off, grp_off, members = flatten_sgs(sgs, labels)
    d_set, d_grp, d_mem = (torch.from_numpy(a).to(_dev()) for a in (set_off, grp_off, members))
    d_len = torch.tensor(lengths, dtype=torch.int64, device=_dev())
    ptrs = torch.tensor([[d.keys.data_ptr() for d in dumps], [d.counts.data_ptr() for d in dumps],
                         [d.pindex.data_ptr() for d in dumps]], dtype=torch.int64).to(_dev())
    counters = _zeros(8, torch.int64)
    total = sum(len(d) for d in dumps)
    # entries that fall into THIS rank's partitions (full_dumps: which dumps hold all partitions — the rank's own
    # chromosomes — as opposed to the class slices received from other ranks); the kernel sizes its shared-memory
    # table for the mean entries per partition, total_entries / 2^pbits
    if nparts > 1 and full_dumps is not None:
        mine = sum((len(d) // nparts) if f else len(d) for d, f in zip(dumps, full_dumps))
        table_total = mine * nparts
    else:
        table_total = total
    # candidates are a few % of the union; grown on demand.  (On several ranks `total` counts the foreign dumps by
    # their partition class only, so the share per rank is estimated more generously.)
    cap = max(total // (16 * nparts) if nparts == 1 else total // (4 * nparts), 1 << 16)
    cap = max(cap, _PM_CAP.get((n, total // 1024), 0))   # what an earlier pass over the same dumps needed
    while True:
        okeys = _empty(cap, torch.int64)
        ocnt = _empty(cap * n, torch.int32).view(cap, n)
        call("spk_pmatrix_filter", _p(ptrs[0]), _p(ptrs[1]), _p(ptrs[2]), n, pbits, nparts, part, _p(d_len),
             _p(d_set), len(set_off) - 1, _p(d_grp), len(grp_off) - 1, _p(d_mem), len(members), float(min_fold),
             int(baseline), int(bool(by_count)), float(ratio), float(min_freq), float(max_freq), _p(okeys), _p(ocnt),
             cap, table_total, _p(counters), st)
        n_union, _, n_cand, n_over = (int(x) for x in counters[:4].cpu().tolist())
        if n_over or os.environ.get("SPK_PMATRIX_FORCE_OVERFLOW") == "1":
            raise OverflowError("partition table overflow in spk_pmatrix_filter")
        if n_cand <= cap:
            break
        del okeys, ocnt
        cap = n_cand + n_cand // 16
        _PM_CAP[(n, total // 1024)] = cap
    # the compact candidate matrix goes through the ordinary filter kernels (full occupancy); rows rejected
    # by the pre-screen all fail the fold test, so its counters are those of the whole union
    cm = CountMatrix(ocnt[:n_cand], okeys[:n_cand], lengths, k, labels)
    dm = filter_matrix(cm, sgs, labels, min_fold=min_fold, baseline=baseline, ratio=ratio, min_freq=min_freq,
                       max_freq=max_freq, by_count=by_count, want_fold_tots=want_fold_tots)
    return dm, n_union


def argsort_keys(keys, key_bits):
    """Stable ascending argsort of uint64 keys held in an int64 tensor (spk_sort_pairs_u64)."""
    require_cuda()
    lib = _lib.load()
    n = int(keys.numel())
    idx = torch.arange(n, dtype=torch.int32, device=_dev())
    if n > 1:
        k2 = keys.clone()
        kt, it = _empty(n, torch.int64), _empty(n, torch.int32)
        ws_bytes = lib.spk_sort_workspace_bytes(n)
        ws = _empty(ws_bytes, torch.uint8)
        call("spk_sort_pairs_u64", _p(k2), _p(idx), _p(kt), _p(it), n, int(key_bits), _p(ws), ws_bytes, _stream())
    return idx.long()


# ----------------------------------------------------------------------------------------------------
# K5-K8: cluster statistics
# ----------------------------------------------------------------------------------------------------
def resample_indices(M, R, seed):
    """The bootstrap's resampling plan (sklearn.utils.resample(replace=True, n_samples=R), R times): int32 [R, R]
    indices into the M k-mers, drawn on the device (the reference is unseeded, Cluster.py:90; drawing 10^6
    numbers on the host and copying them cost more than the whole clustering)."""
    require_cuda()
    g = torch.Generator(device=_dev())
    g.manual_seed(int(seed))
    return torch.randint(0, int(M), (int(R), int(R)), generator=g, device=_dev(), dtype=torch.int32)


def zscore_rows(X):
    """X float64 [M, n] device -> Z (Cluster.normalize_data on data.T, Cluster.py:76-80)."""
    require_cuda()
    M, n = X.shape
    Z = torch.empty_like(X)
    call("spk_zscore_rows", _p(X), M, n, _p(Z), _stream())
    return Z


def gram(Z, idx=None):
    require_cuda()
    lib = _lib.load()
    M, n = Z.shape
    G = _empty(n * n, torch.float64).view(n, n)
    ws_bytes = lib.spk_gram_workspace_bytes(n)
    ws = _empty(ws_bytes, torch.uint8)
    call("spk_gram", _p(Z), M, n, _p(idx), 0 if idx is None else idx.numel(), _p(G), _p(ws), ws_bytes,
         _stream())
    return G


def kmeans_gram(G, S, order=None, n_init=10, max_iter=300, seed=0, r0=0):
    """K-Means of the n points behind Gram matrices G [R, n, n] -> (labels int32 [R, n], inertia [R]).
    r0: global number of the first replicate (ranks sharing a bootstrap keep the one-rank random streams)."""
    require_cuda()
    lib = _lib.load()
    if G.dim() == 2:
        G = G.unsqueeze(0)
    R, n, _ = G.shape
    G = G.contiguous()
    labels = _empty(R * n, torch.int32).view(R, n)
    inertia = _empty(R, torch.float64)
    ws_bytes = lib.spk_kmeans_workspace_bytes(R)
    ws = _empty(ws_bytes, torch.uint8)
    d_order = None if order is None else torch.as_tensor(order, dtype=torch.int32, device=_dev())
    call("spk_kmeans_gram_at", _p(G), R, int(r0), n, S, n_init, max_iter, seed, _p(d_order), _p(labels), _p(inertia),
         _p(ws), ws_bytes, _stream())
    return labels, inertia


def gram_batched(Z, idx):
    """idx int32 [R, B] device -> G [R, n, n]."""
    require_cuda()
    M, n = Z.shape
    R, B = idx.shape
    G = _empty(R * n * n, torch.float64).view(R, n, n)
    call("spk_gram_batched", _p(Z), M, n, _p(idx.contiguous()), R, B, _p(G), _stream())
    return G


def cluster_scores(ref_labels, labels):
    require_cuda()
    R, n = labels.shape
    ari = _empty(R, torch.float64)
    vm = _empty(R, torch.float64)
    ref = torch.as_tensor(ref_labels, dtype=torch.int32, device=_dev()).contiguous()
    call("spk_cluster_scores", _p(ref), _p(labels.contiguous()), R, n, _p(ari), _p(vm), _stream())
    return ari, vm


def centroids(Z, labels, S):
    require_cuda()
    M, n = Z.shape
    C = _empty(S * M, torch.float64).view(S, M)
    lab = torch.as_tensor(labels, dtype=torch.int32, device=_dev()).contiguous()
    call("spk_centroids", _p(Z), M, n, _p(lab), S, _p(C), _stream())
    return C


def ttest_groups(X, col_group, S):
    """Cluster._output_kmers arithmetic -> (best int32 [M], pval [M], means [M, S])."""
    require_cuda()
    M, n = X.shape
    best = _empty(M, torch.int32)
    pval = _empty(M, torch.float64)
    means = _empty(M * S, torch.float64).view(M, S)
    cg = torch.as_tensor(col_group, dtype=torch.int32, device=_dev()).contiguous()
    call("spk_ttest_groups", _p(X), M, n, _p(cg), S, _p(best), _p(pval), _p(means), _stream())
    return best, pval, means


_RANK_METHODS = {"kruskal": 1, "mannwhitneyu": 2, "wilcoxon": 3}


def _mwu_cdf(n1, n2):
    """P(U <= u), u = 0 .. n1*n2//2, of the Mann-Whitney statistic without ties: exact integer counts (number of ways to
    pick the n1 ranks of the first sample) divided by C(n1+n2, n1)."""
    import math
    from functools import lru_cache
    umax = n1 * n2 // 2

    @lru_cache(maxsize=None)
    def ways(u, m, n):          # arrangements of m x-values and n y-values with U = u
        if u < 0:
            return 0
        if m == 0 or n == 0:
            return 1 if u == 0 else 0
        return ways(u - n, m - 1, n) + ways(u, m, n - 1)

    total = math.comb(n1 + n2, n1)
    acc, out = 0, []
    for u in range(umax + 1):
        acc += ways(u, n1, n2)
        out.append(acc / total)
    return out


def _wilcoxon_tables(n):
    """cdf[k] = P(W <= k), then sf[k] = P(W >= k), k = 0 .. n(n+1)/2, of the signed-rank statistic of n untied
    differences (number of subsets of {1..n} with sum k, over 2^n)."""
    K = n * (n + 1) // 2
    cnt = [0] * (K + 1)
    cnt[0] = 1
    for r in range(1, n + 1):
        for k in range(K, r - 1, -1):
            cnt[k] += cnt[k - r]
    tot = 2 ** n
    cdf, acc = [], 0
    for c in cnt:
        acc += c
        cdf.append(acc / tot)
    sf, acc = [0.0] * (K + 1), 0
    for k in range(K, -1, -1):
        acc += cnt[k]
        sf[k] = acc / tot
    return cdf + sf


def ranktest_groups(X, col_group, S, method):
    """Cluster._output_kmers with test_method kruskal / mannwhitneyu / wilcoxon -> (best, pval, means, flags)."""
    require_cuda()
    if method not in _RANK_METHODS:
        raise ValueError("unknown test_method {!r}".format(method))
    M, n = X.shape
    sizes = [int(sum(1 for g in col_group if g == s)) for s in range(S)]
    pair_off = np.full(S * S, -1, dtype=np.int32)
    tables, cache = [], {}
    for a in range(S):
        for b in range(S):
            if a == b or not sizes[a] or not sizes[b]:
                continue
            if method == "mannwhitneyu" and min(sizes[a], sizes[b]) <= 8:
                key = ("m", min(sizes[a], sizes[b]), max(sizes[a], sizes[b]))
                if key not in cache:
                    cache[key] = len(tables)
                    tables.extend(_mwu_cdf(key[1], key[2]))
                pair_off[a * S + b] = cache[key]
            elif method == "wilcoxon" and sizes[a] == sizes[b] and sizes[a] <= 25:
                key = ("w", sizes[a])
                if key not in cache:
                    cache[key] = len(tables)
                    tables.extend(_wilcoxon_tables(sizes[a]))
                pair_off[a * S + b] = cache[key]
    d_tab = torch.tensor(tables or [0.0], dtype=torch.float64, device=_dev())
    d_off = torch.from_numpy(pair_off).to(_dev())
    best = _empty(M, torch.int32)
    pval = _empty(M, torch.float64)
    means = _empty(M * S, torch.float64).view(M, S)
    flags = _zeros(1, torch.int32)
    cg = torch.as_tensor(col_group, dtype=torch.int32, device=_dev()).contiguous()
    call("spk_ranktest_groups", _p(X), M, n, _p(cg), S, _RANK_METHODS[method], _p(d_off), _p(d_tab), _p(best), _p(pval),
         _p(means), _p(flags), _stream())
    return best, pval, means, int(flags.item())


def pca_gram(G, ncomp):
    require_cuda()
    lib = _lib.load()
    n = G.shape[0]
    eig = _empty(n, torch.float64)
    scores = _empty(n * ncomp, torch.float64).view(n, ncomp)
    ratio = _empty(ncomp, torch.float64)
    ws_bytes = lib.spk_pca_workspace_bytes(n)
    ws = _empty(ws_bytes, torch.uint8)
    call("spk_pca_gram", _p(G.contiguous()), n, ncomp, _p(eig), _p(scores), _p(ratio), _p(ws), ws_bytes,
         _stream())
    return eig, scores, ratio


# ----------------------------------------------------------------------------------------------------
# K9: map specific k-mers to bins
# ----------------------------------------------------------------------------------------------------
class SigTable:
    """canonical specific k-mer -> subgenome index on the device.

    Shipped layout: bucketed quotient table (spk_qtable_*, one 16/32-byte load per lookup, L2-resident)
    with a small open-addressed stash; `layout="open"` (or no bucketed layout for this k / S / size)
    uses the plain open-addressed table + one-hash bitmap of spk_sig_table_build."""

    def __init__(self, keys, vals, k, track_hits=True, S=None, layout=None):
        require_cuda()
        lib = _lib.load()
        n = int(keys.numel())
        self.k = int(k)
        self.pack_vals = 1 if self.k <= 28 else 0
        self.n = n
        self.S = int(S) if S is not None else (int(vals.max().item()) + 1 if n else 1)
        layout = layout or os.environ.get("SPK_MAP_LAYOUT", "bucket")
        self.bucket = False
        fail = _zeros(1, torch.int64)
        if layout == "bucket":
            sb, bb = ctypes.c_int(0), ctypes.c_int(0)
            if lib.spk_qtable_plan(n, self.k, self.S, ctypes.byref(sb), ctypes.byref(bb)) == 0:
                self.bucket = True
                self.slot_bits, self.bucket_bits = sb.value, bb.value
        if self.bucket:
            nslots = 8 << self.bucket_bits
            self.buckets = torch.full((nslots,), -1, dtype=torch.int16 if self.slot_bits == 16 else torch.int32,
                                      device=_dev())
            self.slots = max(n // 4 + 64, 1024)           # stash: ~1 % of the keys land here
            self.skeys = torch.full((self.slots,), -1, dtype=torch.int64, device=_dev())
            self.svals = _zeros(self.slots, torch.uint8)
            call("spk_qtable_build", _p(keys), _p(vals), n, self.k, self.S, _p(self.buckets), self.slot_bits,
                 self.bucket_bits, _p(self.skeys), _p(self.svals), self.slots, self.pack_vals, _p(fail), _stream())
            nflags = nslots + self.slots
        else:
            self.slots = max(2 * n + 64, 1024)
            self.skeys = torch.full((self.slots,), -1, dtype=torch.int64, device=_dev())
            self.svals = _zeros(self.slots, torch.uint8)
            self.filter_bits = 1024
            while self.filter_bits < 16 * n:
                self.filter_bits *= 2
            self.filter = _zeros(self.filter_bits // 32, torch.int32)
            call("spk_sig_table_build", _p(keys), _p(vals), n, _p(self.skeys), _p(self.svals), self.slots,
                 _p(self.filter), self.filter_bits, self.pack_vals, _p(fail), _stream())
            nflags = self.slots
        if int(fail.item()):
            raise OverflowError("specific k-mer table full")
        self.hit_flags = _zeros((nflags + 3) // 4 * 4, torch.uint8) if track_hits else None

    def n_mapped(self):
        """number of distinct k-mer strings seen in the mapped sequences, both orientations counted separately (what
        `len(mapped_cat)` is in Seqs.py:113); the plain open-addressed layout only knows the canonical hits"""
        if self.hit_flags is None:
            return 0
        f = self.hit_flags
        return int(((f & 1) != 0).sum().item() + ((f & 2) != 0).sum().item())


def map_bins(seq, sig, S, bin_size, chunk_size, record_lengths=None, sync=True):
    """-> (line_counts int32 [n_lines, S] device, n_hits) — n_hits as a device tensor when sync=False.
    record_lengths (multi-record FASTA): bases of every record in file order; the rows of record r then start at
    sum(spk_map_num_lines(L_q) for q < r) and use the record's own coordinates (Seqs.py:121-153)."""
    require_cuda()
    lib = _lib.load()
    rec_start = rec_line0 = None
    n_rec = 0
    if record_lengths is not None and len(record_lengths) > 1:
        if not sig.bucket:
            raise NotImplementedError("multi-record map needs the bucketed k-mer table (k <= ~24)")
        lens = [int(x) for x in record_lengths]
        n_rec = len(lens)
        starts = np.zeros(n_rec + 1, np.int64)
        starts[1:] = np.cumsum(np.asarray(lens, np.int64) + 1)      # + the separator base before the next record
        starts[n_rec] = seq.n_bases
        nl = [lib.spk_map_num_lines(L, sig.k, int(bin_size), int(chunk_size)) for L in lens]
        line0 = np.zeros(n_rec, np.int64)
        line0[1:] = np.cumsum(nl)[:-1]
        n_lines = int(sum(nl))
        rec_start = torch.from_numpy(starts).to(_dev())
        rec_line0 = torch.from_numpy(line0).to(_dev())
    else:
        n_lines = lib.spk_map_num_lines(seq.n_bases, sig.k, int(bin_size), int(chunk_size))
    counts = _zeros(max(n_lines, 1) * S, torch.int32).view(max(n_lines, 1), S)
    nhits = _zeros(1, torch.int64)
    if seq.n_bases and sig.bucket:
        if S < sig.S:
            raise ValueError("map_bins: table holds %d subgenomes, S=%d" % (sig.S, S))
        call("spk_map_bins_q", _p(seq.packed), _p(seq.valid), seq.n_bases, sig.k, _p(sig.buckets), sig.slot_bits,
             sig.bucket_bits, _p(sig.skeys), _p(sig.svals), sig.slots, sig.pack_vals, sig.S, int(bin_size),
             int(chunk_size), _p(counts), max(n_lines, 1), _p(sig.hit_flags), _p(nhits), _p(rec_start),
             _p(rec_line0), n_rec, _stream())
    elif seq.n_bases:
        call("spk_map_bins", _p(seq.packed), _p(seq.valid), seq.n_bases, sig.k, _p(sig.skeys), _p(sig.svals),
             sig.slots, S, _p(sig.filter), sig.filter_bits, sig.pack_vals, int(bin_size), int(chunk_size), _p(counts),
             max(n_lines, 1), _p(sig.hit_flags), _p(nhits), _stream())
    return counts[:n_lines], (int(nhits.item()) if sync else nhits)


def stack_windows(line_counts, line_window, n_windows):
    """Circos.stack_matrix summation: int64 [L, S] lines -> int64 [W, S] windows (host arrays in/out)."""
    require_cuda()
    lc = torch.as_tensor(np.ascontiguousarray(line_counts, dtype=np.int64)).to(_dev())
    lw = torch.as_tensor(np.ascontiguousarray(line_window, dtype=np.int32)).to(_dev())
    L, S = lc.shape
    out = _zeros(max(n_windows, 1) * S, torch.int64).view(max(n_windows, 1), S)
    call("spk_stack_windows", _p(lc), _p(lw), L, S, _p(out), _stream())
    return out[:n_windows].cpu().numpy()


# ----------------------------------------------------------------------------------------------------
# text wire formats
# ----------------------------------------------------------------------------------------------------
def format_rows(keys, vals, k, kind=0, rows=None, label=None, label_names=None, pval=None):
    """Rows of `.kmer.mat` (kind 0) / `.sig.kmer-subgenome.tsv` (kind 1) formatted on the device (spk_format_rows)
    -> the text as a host uint8 numpy array (one contiguous buffer, rows in order).
    keys int64 [M] device, vals float64 [M, n] device; rows: optional int32 device tensor of row ids;
    kind 1: label int32 [M] device, label_names list of str (<= 16 bytes each), pval float64 [M] device."""
    require_cuda()
    M, n = vals.shape
    vals = vals.contiguous()
    n_rows = int(rows.numel()) if rows is not None else int(M)
    if n_rows == 0:
        return np.zeros(0, np.uint8)
    d_text = d_len = None
    if kind == 1:
        raw = [s.encode() for s in label_names]
        if any(len(b) > 16 for b in raw):
            raise ValueError("subgenome names longer than 16 bytes are not supported by the device writer")
        tab = np.zeros((len(raw), 16), np.uint8)
        for i, b in enumerate(raw):
            tab[i, :len(b)] = np.frombuffer(b, np.uint8)
        d_text = torch.from_numpy(tab).to(_dev())
        d_len = torch.tensor([len(b) for b in raw], dtype=torch.int32, device=_dev())
    row_len = _empty(n_rows, torch.int32)
    st = _stream()
    args = (_p(keys), _p(vals), M, n, int(k), int(kind), _p(rows), n_rows, _p(label), _p(d_text), _p(d_len), _p(pval))
    call("spk_format_rows", *args, None, _p(row_len), None, st)
    off = torch.cumsum(row_len.to(torch.int64), 0)
    total = int(off[-1].item())
    off = (off - row_len).contiguous()
    out = _empty(total, torch.uint8)
    call("spk_format_rows", *args, _p(off), None, _p(out), st)
    host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    host.copy_(out, non_blocking=False)
    return host.numpy()


def write_text(fout, data, chunk=1 << 26):
    """uint8 array -> an open text (or binary) file handle, without building one giant str."""
    buf = getattr(fout, "buffer", None)
    if buf is not None and hasattr(fout, "flush"):
        fout.flush()
        buf.write(memoryview(data))
        return
    for a in range(0, len(data), chunk):
        piece = data[a:a + chunk].tobytes()
        try:
            fout.write(piece.decode("ascii"))
        except TypeError:
            fout.write(piece)


# ----------------------------------------------------------------------------------------------------
# K10: Fisher / enrichment / BH
# ----------------------------------------------------------------------------------------------------
def fisher_enrich(counts, max_pval=0.05, cutoff=1.0, min_ratio=0.5):
    """counts: int64 [W, S] (host or device) -> dict of host numpy arrays
    (pvals [W,S], idx [W], sig [W], ratios [W,S], qvals [W], totals [S])."""
    require_cuda()
    lib = _lib.load()
    c = torch.as_tensor(np.asarray(counts, dtype=np.int64) if not torch.is_tensor(counts) else counts)
    c = c.to(device=_dev(), dtype=torch.int64).contiguous()
    W, S = c.shape
    st = _stream()
    totals = _empty(S, torch.int64)
    call("spk_colsum_i64", _p(c), W, S, _p(totals), st)   # column sums (Stats.py:145)
    pvals = _empty(W * S, torch.float64).view(W, S)
    call("spk_fisher_right_tail", _p(c), _p(totals), W, S, _p(pvals), st)
    idx = _empty(W, torch.int32)
    sig = _empty(W, torch.uint8)
    ratios = _empty(W * S, torch.float64).view(W, S)
    pmin = _empty(W, torch.float64)
    call("spk_enrich_rows", _p(c), _p(totals), _p(pvals), W, S, float(max_pval), float(cutoff),
         float(min_ratio), _p(idx), _p(sig), _p(ratios), _p(pmin), st)
    q = _empty(W, torch.float64)
    if W:
        ws_bytes = lib.spk_bh_workspace_bytes(W)
        ws = _empty(ws_bytes, torch.uint8)
        call("spk_bh_adjust", _p(pmin), _p(q), W, _p(ws), ws_bytes, st)
    return dict(pvals=pvals.cpu().numpy(), idx=idx.cpu().numpy(), sig=sig.cpu().numpy().astype(bool),
                ratios=ratios.cpu().numpy(), pmin=pmin.cpu().numpy(), qvals=q.cpu().numpy(),
                totals=totals.cpu().numpy())
