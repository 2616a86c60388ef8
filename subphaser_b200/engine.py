"""Device-side engine: thin Python orchestration of the libspk kernels.

torch is used for three things only: allocating device/pinned buffers, naming the current CUDA stream
and (in parallel.py) bootstrapping NCCL.  All arithmetic happens in libspk.so.  Nothing here falls back
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
    n_bases, n_valid, n_rec, _ = (int(x) for x in info.cpu().tolist())
    if trim:
        pw, vw = lib.spk_packed_words(max(n_bases, 1)), lib.spk_valid_words(max(n_bases, 1))
        packed = packed[:pw].clone()
        valid = valid[:vw].clone()
    del ws
    return PackedSeq(packed, valid, n_bases, n_valid, n_rec, name)


# ----------------------------------------------------------------------------------------------------
# K2/K3: count -> dump
# ----------------------------------------------------------------------------------------------------
class KmerDump:
    """`jellyfish dump -c -L` content of one chromosome, resident on the device."""

    def __init__(self, keys, counts, k, length, n_valid_kmers, n_distinct, name=None, histo=None):
        self.keys, self.counts, self.k = keys, counts, int(k)
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
    """Reusable device scratch for counting chromosomes of up to `max_bases` bases."""

    def __init__(self, max_bases, k):
        require_cuda()
        lib = _lib.load()
        self.k = int(k)
        self.max_bases = int(max_bases)
        self.layout = lib.spk_count_layout(self.max_bases, self.k)
        self.table_bytes = lib.spk_count_table_bytes(self.max_bases, self.k)
        self.table = _empty(self.table_bytes, torch.uint8)
        nb = lib.spk_table_scan_blocks()
        self.block_counts = _empty(3 * nb + 2, torch.int32)
        self.stats = _zeros(8, torch.int64)

    def slots(self):
        return _lib.load().spk_count_table_slots(self.table_bytes, self.layout)


def count_packed(seq, k, lower_count, table=None, histo_len=0):
    """K2 + K3 on one packed chromosome -> KmerDump."""
    require_cuda()
    if table is None or table.max_bases < seq.n_bases or table.k != k:
        table = CountTable(max(seq.n_bases, 1), k)
    st = _stream()
    call("spk_count_table_init", _p(table.table), table.table_bytes, k, table.layout, st)
    table.stats.zero_()
    call("spk_count_canonical", _p(seq.packed), _p(seq.valid), seq.n_bases, k, _p(table.table),
         table.table_bytes, table.layout, _p(table.stats), st)
    histo = _zeros(histo_len, torch.int64) if histo_len else None
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
    return KmerDump(keys, counts, k, sum_ge, n_valid, distinct, seq.name, histo)
