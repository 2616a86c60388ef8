"""In-process registry of device-resident artefacts keyed by the file path the reference passes
between pipeline steps (`__main__.py` hands paths from step to step; the GPU objects ride along so a
file is not re-parsed when the producing step ran in this process)."""
import os
from collections import OrderedDict

_SEQS = OrderedDict()     # chromosome FASTA path -> engine.PackedSeq
_DUMPS = {}               # dump path (<chromfile>_<k>.fa) -> engine.KmerDump
_MATS = {}                # .kmer.mat path -> (signature, engine.DiffMatrix)
_BINS = {}                # .bin.count path -> (signature, parsed arrays)


def _key(path):
    return os.path.realpath(path)


def _sig(path):
    try:
        st = os.stat(path)
        return (st.st_size, st.st_mtime_ns)
    except OSError:
        return None


def seq_budget_bytes():
    env = os.environ.get("SPK_SEQ_CACHE_GB")
    if env is not None:
        return int(float(env) * 1e9)
    try:
        import torch
        free, total = torch.cuda.mem_get_info()
        return int(total * 0.25)
    except Exception:
        return 8 << 30


def put_seq(path, seq):
    k = _key(path)
    _SEQS.pop(k, None)
    _SEQS[k] = (_sig(path), seq)
    budget = seq_budget_bytes()
    while len(_SEQS) > 1 and sum(s.nbytes() for _, s in _SEQS.values()) > budget:
        _SEQS.popitem(last=False)


def get_seq(path):
    v = _SEQS.get(_key(path))
    if v is None or v[0] != _sig(path):
        return None
    return v[1]


def put_dump(path, dump):
    _DUMPS[_key(path)] = dump


def get_dump(path):
    return _DUMPS.get(_key(path))


def put_matrix(path, dm):
    _MATS[_key(path)] = (_sig(path), dm)


def get_matrix(path):
    v = _MATS.get(_key(path))
    if v is None or v[0] != _sig(path):
        return None
    return v[1]


def put_bins(path, obj):
    _BINS[_key(path)] = (_sig(path), obj)


def get_bins(path):
    v = _BINS.get(_key(path))
    if v is None or v[0] != _sig(path):
        return None
    return v[1]


def clear():
    _SEQS.clear()
    _DUMPS.clear()
    _MATS.clear()
    _BINS.clear()
