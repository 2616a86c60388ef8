"""Drop-in for subphaser/Stats.py (reference v1.2.7): the functions `__main__.py` calls (`enrich_bin`
Stats.py:75, `enrich_ltr` :33) and the helpers other code imports (`enrich` :140, `fisher_test` :14,
`correct_pvals` :11, `group_exchanges` :119, `is_exchange` :133), with the same file formats and return shapes.

The reference tests one window per Pool task and builds every output line inside that loop.  Here all windows
go through the device at once (K10: column sums, Fisher right tails, enrichment decision, BH — `engine.
fisher_enrich`) and the text tables are produced afterwards from whole columns: `_Table` holds the result arrays,
the formatters below turn a column into strings in one pass, and the exchange groups are found by run-length
encoding the (chromosome, subgenome) columns instead of nested `groupby` loops.
"""
import logging
import re

import numpy as np

from . import engine

logger = logging.getLogger("subphaser_b200")

MAX_INT = 2147483647 // 10          # the clamp of Stats.py:9,24-25 (applied inside spk_fisher_right_tail)
_LTR_ID = re.compile(r"(\S+?):\d+\-\d+")


def correct_pvals(pvals, method="fdr_bh"):
    """Benjamini-Hochberg as statsmodels `multipletests(method='fdr_bh')[1]` (Stats.py:11-12), on the device."""
    import torch
    if method != "fdr_bh":
        raise NotImplementedError("only fdr_bh is implemented")
    engine.require_cuda()
    lib = engine._lib.load()
    p = torch.as_tensor(np.asarray(pvals, dtype=np.float64)).to(engine._dev())
    n = p.numel()
    q = torch.empty_like(p)
    if n:
        ws_bytes = lib.spk_bh_workspace_bytes(n)
        ws = engine._empty(ws_bytes, torch.uint8)
        engine.call("spk_bh_adjust", engine._p(p), engine._p(q), n, engine._p(ws), ws_bytes, engine._stream())
    return q.cpu().numpy()


def fisher_test(each, total):
    """Right-tail Fisher exact p-value of every column of `each` against the totals (Stats.py:14-31)."""
    import torch
    assert len(each) == len(total)
    engine.require_cuda()
    c = torch.as_tensor(np.asarray([list(each)], dtype=np.int64)).to(engine._dev())
    t = torch.as_tensor(np.asarray(list(total), dtype=np.int64)).to(engine._dev())
    S = len(each)
    p = engine._empty(S, torch.float64)
    engine.call("spk_fisher_right_tail", engine._p(c), engine._p(t), 1, S, engine._p(p), engine._stream())
    return p.cpu().numpy().tolist()


def is_exchange(obs_sg, exp_sg):
    if not exp_sg or not obs_sg:
        return "none"
    return "no" if obs_sg == exp_sg else "yes"


class Pvalue:
    """One row's result as the reference's `_min` object (Stats.py:150-168,194-198)."""

    def __init__(self, pval, key, idx):
        self.pval = pval
        self.key = key
        self.idx = idx


class _Table:
    """All rows of one enrichment call: device results as host columns."""

    def __init__(self, matrix, colnames, rownames, min_ratio=0.5, max_pval=0.05, cutoff=1):
        self.matrix = matrix
        self.colnames = list(colnames)
        self.rownames = list(rownames)
        arr = np.array(matrix)
        assert arr.shape == (len(self.rownames), len(self.colnames)), "{} != {}".format(
            arr.shape, (len(self.rownames), len(self.colnames)))
        assert len(self.colnames) > 1
        self.counts = arr.astype(np.int64)
        self.W, self.S = self.counts.shape
        res = engine.fisher_enrich(self.counts, max_pval=max_pval, cutoff=cutoff, min_ratio=min_ratio)
        self.pvals = res["pvals"]                         # [W, S]
        self.idx = res["idx"].astype(np.int64)            # most enriched subgenome of every row
        self.sig = res["sig"]
        self.ratios = res["ratios"]
        self.qvals = res["qvals"]
        self.pmin = self.pvals[np.arange(self.W), self.idx] if self.W else np.zeros(0)
        names = np.array(self.colnames, dtype=object)
        self.key = names[self.idx] if self.W else names[:0]                 # Pvalue.key
        self.called = np.where(self.sig, self.key, None)                    # enriched subgenome or None

    def exchange_column(self, obs):
        """is_exchange() of every row; obs: observed subgenome of the row's chromosome (object array, None = unknown)."""
        known = np.array([bool(o) for o in obs], dtype=bool) & self.sig
        same = np.array([o == c for o, c in zip(obs, self.called)], dtype=bool)
        out = np.full(self.W, "none", dtype=object)
        out[known & same] = "no"
        out[known & ~same] = "yes"
        return out

    def onehot(self):
        """enrich vector per row: 1 at the enriched subgenome, or in the extra last column (Stats.py:162-166)."""
        oh = np.zeros((self.W, self.S + 1), dtype=np.int64)
        oh[np.arange(self.W), np.where(self.sig, self.idx, self.S)] = 1
        return oh

    def row(self, r):
        m = Pvalue(float(self.pmin[r]), self.colnames[self.idx[r]], int(self.idx[r]))
        m.sig = bool(self.sig[r])
        m.counts = self.matrix[r]
        m.pvals = self.pvals[r].tolist()
        m.ratios = self.ratios[r]
        m.ratio = m.ratios[m.idx]
        m.enrich = [0] * (self.S + 1)
        m.enrich[m.idx if m.sig else -1] = 1
        m.rowname = self.rownames[r]
        m.qval = self.qvals[r]
        return m


def _table(matrix, colnames=None, rownames=None, ncpu=4, min_ratio=0.5, max_pval=0.05, cutoff=1, **kargs):
    if len(matrix) == 0:
        return None
    return _Table(matrix, colnames, rownames, min_ratio=min_ratio, max_pval=max_pval, cutoff=cutoff)


def enrich(matrix, colnames=None, rownames=None, ncpu=4, min_ratio=0.5, **kargs):
    """Stats.py:140-168: one result object per row (.rowname .key .idx .sig .pval .pvals .counts .ratios .ratio
    .enrich), all rows evaluated together on the device."""
    t = _table(matrix, colnames=colnames, rownames=rownames, min_ratio=min_ratio, **kargs)
    for r in range(t.W if t is not None else 0):
        yield t.row(r)


# ---- column formatters (Python's str() of what the reference holds in each field) -----------------------------
def _f(x):
    return repr(float(x))       # str(float) / str(np.float64): shortest round-trip repr


def _ints_joined(mat):
    return [",".join(map(str, r)) for r in mat.tolist()]


def _floats_joined(mat):
    return [",".join(map(_f, r)) for r in mat.tolist()]


def _log_consistency(xchg, always):
    total = len(xchg)
    consistent = int(np.sum(xchg == "no"))
    exchange = int(np.sum(xchg == "yes"))
    if always or (exchange > 0 and consistent > 0):
        logger.info("Consistent with subgenome assignment: {} ({:.2%}); potential exchange: {} ({:.2%})".format(
            consistent, consistent / total, exchange, exchange / total))


def enrich_ltr(fout, d_sg, *args, **kargs):
    """Per-sequence enrichment table (Stats.py:33-73): 6 columns, returns (d_enriched, d_exchange)."""
    t = _table(*args, **kargs)
    fout.write("\t".join(["#id", "subgenome", "p_value", "counts", "potential_exchange", "p_corrected"]) + "\n")
    if t is None:
        return {}, {}
    ids = [name[0] for name in t.rownames]
    hits = [_LTR_ID.match(i) if isinstance(i, str) else None for i in ids]
    obs = np.array([d_sg.get(h.group(1)) if h else d_sg.get(None) for h in hits], dtype=object)
    xchg = t.exchange_column(obs)
    _log_consistency(xchg, always=False)
    cols = (ids, ["None" if c is None else c for c in t.called], list(map(_f, t.pmin)), _ints_joined(t.counts),
            xchg.tolist(), list(map(_f, t.qvals)))
    fout.write("".join("\t".join(fields) + "\n" for fields in zip(*cols)))
    d_enriched = {i: c for i, c in zip(ids, t.called) if c}
    d_exchange = dict(zip(ids, xchg.tolist()))
    return d_enriched, d_exchange


def enrich_bin(fout, fout2, d_sg, *args, **kargs):
    """Per-window enrichment (Stats.py:75-118): `.bin.enrich` (11 columns) to fout, the exchange groups to fout2;
    returns the rows as lists (what `Circos.out_sg_lines` consumes)."""
    t = _table(*args, **kargs)
    header = ["#chrom", "start", "end", "subgenome", "p_value", "counts", "ratios", "enrich", "pvals",
              "potential_exchange", "p_corrected"]
    header2 = ["#chrom", "start", "end", "exchange_from", "exchange_to", "N_bins", "potential_exchange"]
    if t is None:
        raise ZeroDivisionError("division by zero")           # the reference's log line divides by the row count
    chroms = [name[0] for name in t.rownames]
    starts = [name[1] for name in t.rownames]
    ends = [name[2] for name in t.rownames]
    obs = np.array([d_sg.get(c) for c in chroms], dtype=object)
    xchg = t.exchange_column(obs)
    _log_consistency(xchg, always=True)
    counts_s, ratios_s = _ints_joined(t.counts), _floats_joined(t.ratios)
    enrich_s, pvals_s = _ints_joined(t.onehot()), _floats_joined(t.pvals)
    lines = [list(fields) for fields in zip(chroms, starts, ends, t.called.tolist(), t.pmin.tolist(), counts_s, ratios_s,
                                            enrich_s, pvals_s, xchg.tolist(), t.qvals.tolist())]
    fout.write("\t".join(header) + "\n")
    text_cols = (chroms, map(str, starts), map(str, ends), ["None" if c is None else c for c in t.called],
                 map(_f, t.pmin), counts_s, ratios_s, enrich_s, pvals_s, xchg.tolist(), map(_f, t.qvals))
    fout.write("".join("\t".join(fields) + "\n" for fields in zip(*text_cols)))
    fout2.write("\t".join(header2) + "\n")
    fout2.write("".join("\t".join(map(str, g)) + "\n" for g in group_exchanges(lines, d_sg)))
    return lines


def group_exchanges(lines, d_sg):
    """Runs of consecutive bins enriched for the same subgenome (Stats.py:119-132): per block of consecutive rows
    of one chromosome, the enriched rows in order of start coordinate, run-length encoded by subgenome."""
    n = len(lines)
    if n == 0:
        return
    chrom = np.array([ln[0] for ln in lines], dtype=object)
    block_start = np.flatnonzero(np.concatenate(([True], chrom[1:] != chrom[:-1])))
    block_end = np.concatenate((block_start[1:], [n]))
    for a, b in zip(block_start.tolist(), block_end.tolist()):
        rows = [ln for ln in lines[a:b] if ln[3] is not None]
        if not rows:
            continue
        rows.sort(key=lambda ln: ln[1])                                    # stable, as sorted() in the reference
        sg = np.array([ln[3] for ln in rows], dtype=object)
        run_start = np.flatnonzero(np.concatenate(([True], sg[1:] != sg[:-1])))
        run_end = np.concatenate((run_start[1:], [len(rows)]))
        obs_sg = d_sg.get(chrom[a])
        for i, j in zip(run_start.tolist(), run_end.tolist()):
            yield [chrom[a], rows[i][1], rows[j - 1][2], sg[i], obs_sg, j - i, is_exchange(obs_sg, sg[i])]
