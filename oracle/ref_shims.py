"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by subphaser_b200/).

Imports the UNMODIFIED reference modules from /root/reference under import shims for the third-party
packages that are absent in this image (xopen, Bio, fisher, statsmodels, matplotlib, ...), so that the
reference's own Python functions can serve as the oracle for everything except the jellyfish shell-out
(SURVEY.md §8c, Appendix A).  /root/reference exists only in the build container: callers must check
`available()`; the committed fixtures under tests/golden/ (made by tests/golden/make_golden.py with
this module) are what travels to the GPU box.

Shimmed numerics (documented stand-ins, per BASELINE.json "scipy.stats CPU path"):
  fisher.pvalue(a,b,c,d).right_tail  -> scipy.stats.hypergeom.sf(a-1, a+b+c+d, a+b, a+c)
  statsmodels multipletests(fdr_bh)  -> the statsmodels formula restated in numpy
"""
import gzip
import importlib
import importlib.abc
import importlib.machinery
import io
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "subphaser"))


class _Seq(str):
    _comp = str.maketrans("ACGTacgtNn", "TGCAtgcaNn")

    def reverse_complement(self):
        return _Seq(str(self).translate(self._comp)[::-1])

    def upper(self):
        return _Seq(str.upper(self))


class _Record:
    def __init__(self, rid, seq, desc=""):
        self.id = rid
        self.seq = _Seq(seq)
        self.description = desc

    def __len__(self):
        return len(self.seq)


def _fasta_parse(handle, fmt="fasta"):
    if isinstance(handle, str):
        handle = open(handle)
    rid, desc, chunks = None, "", []
    for line in handle:
        if isinstance(line, bytes):
            line = line.decode()
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            if rid is not None:
                yield _Record(rid, "".join(chunks).replace(" ", "").replace("\r", ""), desc)
            desc = line[1:]
            rid = desc.split()[0] if desc.split() else ""
            chunks = []
        elif rid is not None:
            chunks.append(line)
    if rid is not None:
        yield _Record(rid, "".join(chunks), desc)


def _fasta_write(records, handle, fmt="fasta"):
    if isinstance(records, _Record):
        records = [records]
    n = 0
    for rc in records:
        # Bio.SeqIO.FastaIO.FastaWriter.write_record (BioPython 1.79): the parsed title is kept; a changed id goes in front
        rid, desc = str(rc.id), str(rc.description or "")
        if desc and desc.split(None, 1)[0] == rid:
            title = desc
        elif desc:
            title = "{} {}".format(rid, desc)
        else:
            title = rid
        handle.write(">{}\n".format(title))
        s = str(rc.seq)
        for i in range(0, len(s), 60):
            handle.write(s[i:i + 60] + "\n")
        n += 1
    return n


def _xopen(path, mode="r", **kw):
    if isinstance(path, str) and path.endswith(".gz"):
        return gzip.open(path, mode if "b" in mode else mode + "t")
    if "b" in mode:
        return open(path, mode)
    return open(path, mode)


class _FisherP:
    def __init__(self, a, b, c, d):
        from scipy.stats import hypergeom
        self.right_tail = float(hypergeom.sf(a - 1, a + b + c + d, a + b, a + c))
        self.left_tail = float(hypergeom.cdf(a, a + b + c + d, a + b, a + c))


def bh_statsmodels(pvals):
    """statsmodels.stats.multitest.multipletests(method='fdr_bh')[1] restated (statsmodels 0.13.1,
    multitest.py: pvals_sorted / ecdf, reverse cumulative minimum, clip to 1, unsort)."""
    import numpy as np
    p = np.asarray(pvals, dtype=float)
    n = len(p)
    order = np.argsort(p)
    ps = np.take(p, order)
    ecdf = np.arange(1, n + 1) / float(n)
    raw = ps / ecdf
    corr = np.minimum.accumulate(raw[::-1])[::-1]
    corr[corr > 1] = 1
    out = np.empty_like(corr)
    out[order] = corr
    return out


class _Permissive(types.ModuleType):
    """A module that fabricates any attribute (used for packages only imported, never called)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Permissive(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return _Permissive(self.__name__ + "()")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("Bio", "matplotlib", "svglib", "reportlab", "pp", "drmaa", "statsmodels", "xopen", "fisher")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.PREFIXES and name not in sys.modules:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Permissive(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Register the shims (idempotent) and make `import subphaser.X` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference not present at " + REFERENCE_ROOT)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    mod("xopen", xopen=_xopen)
    bio = mod("Bio")
    bio.Seq = mod("Bio.Seq", Seq=_Seq)
    bio.SeqIO = mod("Bio.SeqIO", parse=_fasta_parse, write=_fasta_write)
    mod("fisher", pvalue=lambda a, b, c, d: _FisherP(a, b, c, d))
    sm = mod("statsmodels")
    sm.stats = mod("statsmodels.stats")
    sm.stats.multitest = mod(
        "statsmodels.stats.multitest",
        multipletests=lambda p, method="fdr_bh", **k: (None, bh_statsmodels(p)),
    )

    class _Plt(_Permissive):
        def switch_backend(self, *a, **k):
            pass

    mpl = _Permissive("matplotlib")
    mpl.__path__ = []
    sys.modules["matplotlib"] = mpl
    plt = _Plt("matplotlib.pyplot")
    sys.modules["matplotlib.pyplot"] = plt
    mpl.pyplot = plt
    sys.meta_path.append(_Finder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load(name):
    """Return reference module subphaser.<name> (e.g. 'Jellyfish', 'Stats', 'Seqs', 'Cluster')."""
    install()
    return importlib.import_module("subphaser." + name)
