"""In-process registry of device-resident artefacts keyed by the file path the reference passes
between pipeline steps (`__main__.py` hands paths from step to step; the GPU objects ride along so a
file is not re-parsed when the producing step ran in this process)."""
import os
from collections import OrderedDict

_SEQS = OrderedDict()     # chromosome FASTA path -> engine.PackedSeq
_DUMPS = OrderedDict()    # dump path (<chromfile>_<k>.fa) -> (engine.KmerDump, k, lower_count, source signature)
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


DUMP_BUDGET = 256          # dumps kept on the device at most (a 21-chromosome genome uses 21; oldest go first)


def put_dump(path, dump, k=None, lower_count=None, src_sig=None):
    """dump of `path`, remembered together with the parameters and the source-file signature it was made from"""
    key = _key(path)
    _DUMPS.pop(key, None)
    _DUMPS[key] = (dump, k, lower_count, src_sig)
    while len(_DUMPS) > DUMP_BUDGET:
        _DUMPS.pop(next(iter(_DUMPS)))


def get_dump(path, k=None, lower_count=None, src_sig=None):
    """the registered dump, or None when it was made with other parameters / from a file that changed since"""
    v = _DUMPS.get(_key(path))
    if v is None:
        return None
    dump, k0, lc0, sig0 = v
    if k is not None and k0 is not None and int(k) != int(k0):
        return None
    if lower_count is not None and lc0 is not None and int(lower_count) != int(lc0):
        return None
    if src_sig is not None and sig0 is not None and tuple(src_sig) != tuple(sig0):
        return None
    return dump


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
