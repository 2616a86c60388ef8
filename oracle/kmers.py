"""ORACLE — TEST INFRASTRUCTURE ONLY.  Canonical k-mer counting on the CPU.

`count_fasta` drives oracle/kmer_count.c (the restatement of `jellyfish count -m K --canonical` +
`jellyfish dump -c -L`, reference call site subphaser/Jellyfish.py:697-700).  `brute_count` is an
independent string-level counter (slices, str reverse complement, min of the two strings) used to pin
the C code on small inputs.  PARITY UNPINNED against jellyfish 2.2.10 itself (binary absent).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("n_bases", "n_records", "n_valid_kmers", "n_distinct", "n_dumped", "sum_dumped")]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(path)
        lib.orc_count_fasta.restype = ctypes.c_int
        lib.orc_count_fasta.argtypes = [
            ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_int,
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
            ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(_Stats)]
        lib.orc_free.argtypes = [ctypes.c_void_p]
        lib.orc_fasta_to_codes.restype = ctypes.c_uint64
        lib.orc_fasta_to_codes.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                                           ctypes.POINTER(ctypes.c_uint64)]
        lib.orc_map_bins.restype = ctypes.c_uint64
        lib.orc_map_bins.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64,
                                     ctypes.c_uint64, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def count_fasta(fasta_bytes, k, lower_count=1, nthreads=1):
    """-> (keys uint64 sorted ascending, counts uint32 aligned, stats dict)."""
    buf = np.frombuffer(fasta_bytes, dtype=np.uint8) if not isinstance(fasta_bytes, np.ndarray) else fasta_bytes
    buf = np.ascontiguousarray(buf)
    keys_p, counts_p = ctypes.c_void_p(), ctypes.c_void_p()
    n = ctypes.c_uint64()
    st = _Stats()
    rc = _lib().orc_count_fasta(buf.ctypes.data, buf.size, k, lower_count, nthreads,
                                ctypes.byref(keys_p), ctypes.byref(counts_p), ctypes.byref(n),
                                ctypes.byref(st))
    if rc != 0:
        raise RuntimeError("orc_count_fasta failed: %d" % rc)
    m = n.value
    keys = np.ctypeslib.as_array(ctypes.cast(keys_p, ctypes.POINTER(ctypes.c_uint64)), shape=(m + 1,))[:m].copy()
    counts = np.ctypeslib.as_array(ctypes.cast(counts_p, ctypes.POINTER(ctypes.c_uint32)), shape=(m + 1,))[:m].copy()
    _lib().orc_free(keys_p)
    _lib().orc_free(counts_p)
    order = np.argsort(keys, kind="stable")
    stats = {f: getattr(st, f) for f, _ in _Stats._fields_}
    return keys[order], counts[order], stats


def fasta_to_codes(fasta_bytes):
    buf = np.ascontiguousarray(np.frombuffer(fasta_bytes, dtype=np.uint8))
    codes = np.empty(buf.size + 1, dtype=np.uint8)
    nrec = ctypes.c_uint64()
    n = _lib().orc_fasta_to_codes(buf.ctypes.data, buf.size, codes.ctypes.data, ctypes.byref(nrec))
    return codes[:n].copy(), nrec.value


def map_bins(codes, k, keys_sorted, sgs, S, bin_size, chunk, n_lines):
    """Restates Seqs.map_kmer_each4 (Seqs.py:209-237) on a code array -> (counts [n_lines,S], hits)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    keys_sorted = np.ascontiguousarray(keys_sorted, dtype=np.uint64)
    sgs = np.ascontiguousarray(sgs, dtype=np.uint8)
    counts = np.zeros((n_lines, S), dtype=np.uint32)
    hits = _lib().orc_map_bins(codes.ctypes.data, codes.size, k, keys_sorted.ctypes.data,
                               sgs.ctypes.data, keys_sorted.size, S, int(bin_size), int(chunk),
                               counts.ctypes.data)
    return counts, hits


_COMP = str.maketrans("ACGT", "TGCA")


def revcomp(s):
    return s.translate(_COMP)[::-1]


def brute_count(records, k):
    """String-level canonical counter: records = list of sequence strings (one per FASTA record)."""
    d = {}
    for seq in records:
        seq = seq.upper()
        for i in range(len(seq) - k + 1):
            kmer = seq[i:i + k]
            if any(c not in "ACGT" for c in kmer):
                continue
            rc = revcomp(kmer)
            canon = kmer if kmer < rc else rc
            d[canon] = d.get(canon, 0) + 1
    return d


def key_to_str(key, k):
    key = int(key)
    return "".join("ACGT"[(key >> (2 * (k - 1 - i))) & 3] for i in range(k))


def str_to_key(s):
    v = 0
    for c in s:
        v = (v << 2) | "ACGT".index(c)
    return v
