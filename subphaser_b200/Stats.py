"""Drop-in for subphaser/Stats.py (reference v1.2.7): same functions and file formats
(`enrich_bin` :75, `enrich_ltr` :33, `enrich` :140, `fisher_test` :14, `correct_pvals` :11,
`group_exchanges` :119, `is_exchange` :133), with the per-window Fisher exact tests, the enrichment
decision and the BH correction batched on the GPU (K10) instead of one Pool task per row."""
import logging
import re
from itertools import groupby

import numpy as np

from . import engine

logger = logging.getLogger("subphaser_b200")

MAX_INT = 2147483647 // 10


def correct_pvals(pvals, method="fdr_bh"):
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


class Pvalue:
    """The `_min` object the reference passes around (Stats.py:194-198 + attributes set in _enrich)."""

    def __init__(self, pval, key, idx):
        self.pval = pval
        self.key = key
        self.idx = idx


def enrich(matrix, colnames=None, rownames=None, ncpu=4, min_ratio=0.5, max_pval=0.05, cutoff=1, **kargs):
    """Stats.py:140-168: yields one result per row with .rowname .key .idx .sig .pval .pvals .counts
    .ratios .ratio .enrich — computed for all rows at once on the device."""
    arr = np.array(matrix)
    if colnames is not None and rownames is not None:
        assert arr.shape == (len(rownames), len(colnames)), "{} != {}".format(
            arr.shape, (len(rownames), len(colnames)))
    if arr.size == 0:
        return
    assert len(colnames) > 1
    res = engine.fisher_enrich(arr.astype(np.int64), max_pval=max_pval, cutoff=cutoff, min_ratio=min_ratio)
    S = len(colnames)
    pvals_all = res["pvals"].tolist()
    for r, (row, rowname) in enumerate(zip(matrix, rownames)):
        idx = int(res["idx"][r])
        _min = Pvalue(pvals_all[r][idx], colnames[idx], idx)
        _min.sig = bool(res["sig"][r])
        _min.counts = row
        _min.pvals = pvals_all[r]
        _min.ratios = res["ratios"][r]
        _min.ratio = _min.ratios[idx]
        _min.enrich = [0] * (S + 1)
        if _min.sig:
            _min.enrich[idx] = 1
        else:
            _min.enrich[-1] = 1
        _min.rowname = rowname
        _min.qval = res["qvals"][r]
        yield _min


def _qvals(results, pvalues):
    # BH over the minimum p-values was already computed on the device with the batch
    if results and all(hasattr(r, "qval") for r in results):
        return [r.qval for r in results]
    return correct_pvals(pvalues)


def enrich_ltr(fout, d_sg, *args, **kargs):
    """Output LTR enrichments (Stats.py:33-73)"""
    total, consistent, exchange = 0, 0, 0
    d_enriched = {}
    d_exchange = {}
    lines = []
    pvalues = []
    results = []
    for res in enrich(*args, **kargs):
        ltr, *_ = res.rowname
        try:
            chrom = re.compile(r"(\S+?):\d+\-\d+").match(ltr).groups()[0]
        except (TypeError, AttributeError):
            chrom = None
        obs_sg = d_sg.get(chrom)
        sg = res.key if res.sig else None
        potential_exchange = is_exchange(obs_sg, sg)
        counts = ",".join(map(str, res.counts))
        line = [ltr, sg, res.pval, counts, potential_exchange]
        lines += [line]
        pvalues += [res.pval]
        results += [res]
        if sg:
            d_enriched[ltr] = sg
        d_exchange[ltr] = potential_exchange
        total += 1
        if potential_exchange == "yes":
            exchange += 1
        elif potential_exchange == "no":
            consistent += 1
    if exchange > 0 and consistent > 0:
        logger.info("Consistent with subgenome assignment: {} ({:.2%}); potential exchange: {} ({:.2%})".format(
            consistent, consistent / total, exchange, exchange / total))
    qvals = _qvals(results, pvalues)
    line = ["#id", "subgenome", "p_value", "counts", "potential_exchange", "p_corrected"]
    fout.write("\t".join(line) + "\n")
    for line, qval in zip(lines, qvals):
        line += [qval]
        fout.write("\t".join(map(_s, line)) + "\n")
    return d_enriched, d_exchange


def enrich_bin(fout, fout2, d_sg, *args, **kargs):
    """Enrich by chromosome bins (Stats.py:75-118)"""
    total, consistent, exchange = 0, 0, 0
    lines = []
    pvalues = []
    results = []
    for res in enrich(*args, **kargs):
        chrom, start, end = res.rowname
        key = res.key if res.sig else None
        obs_sg = d_sg.get(chrom)
        potential_exchange = is_exchange(obs_sg, key)
        counts = ",".join(map(str, res.counts))
        enrichs = ",".join(map(str, res.enrich))
        ratios = ",".join(map(_s, res.ratios))
        pvals = ",".join(map(_s, res.pvals))
        line = [chrom, start, end, key, res.pval, counts, ratios, enrichs, pvals, potential_exchange]
        lines += [line]
        pvalues += [res.pval]
        results += [res]
        total += 1
        if potential_exchange == "yes":
            exchange += 1
        elif potential_exchange == "no":
            consistent += 1
    logger.info("Consistent with subgenome assignment: {} ({:.2%}); potential exchange: {} ({:.2%})".format(
        consistent, consistent / total, exchange, exchange / total))
    qvals = _qvals(results, pvalues)
    line = ["#chrom", "start", "end", "subgenome", "p_value", "counts", "ratios", "enrich", "pvals",
            "potential_exchange", "p_corrected"]
    fout.write("\t".join(line) + "\n")
    for line, qval in zip(lines, qvals):
        line += [qval]
        fout.write("\t".join(map(_s, line)) + "\n")
    line = ["#chrom", "start", "end", "exchange_from", "exchange_to", "N_bins", "potential_exchange"]
    fout2.write("\t".join(line) + "\n")
    for line in group_exchanges(lines, d_sg):
        fout2.write("\t".join(map(str, line)) + "\n")
    return lines


def _s(x):
    """str() as the reference prints values: floats by shortest repr (Python float and numpy float64
    print identically), everything else by str()."""
    if isinstance(x, (float, np.floating)):
        return repr(float(x))
    return str(x)


def group_exchanges(lines, d_sg):
    for chrom, items in groupby(lines, key=lambda x: x[0]):
        obs_sg = d_sg.get(chrom)
        items = [line for line in items if line[3] is not None]
        items = sorted(items, key=lambda x: x[1])
        for sg, xlines in groupby(items, key=lambda x: x[3]):
            potential_exchange = is_exchange(obs_sg, sg)
            xlines = list(xlines)
            start = xlines[0][1]
            end = xlines[-1][2]
            yield [chrom, start, end, sg, obs_sg, len(xlines), potential_exchange]


def is_exchange(obs_sg, exp_sg):
    if not exp_sg or not obs_sg:
        return "none"
    if obs_sg == exp_sg:
        return "no"
    return "yes"
