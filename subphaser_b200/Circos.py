"""Drop-in for the hot-path surface of subphaser/Circos.py: `stack_matrix` (:831-842) on top of
`_bed_density(stack=True)` (:709-742) — re-bin the 10-kb count lines into `window_size` windows,
first-seen order, zero-hit windows absent, window end not clipped.  The summation is the
spk_stack_windows kernel; parsing / ordering is host glue.  Everything else in the reference module is
visualisation (circos files) and out of scope."""
import numpy as np

from . import _registry, engine


def _parse_bin_file(path):
    """-> list of (chrom, starts int64[], ends int64[], counts int64[,S]) preserving file order."""
    cached = _registry.get_bins(path) if isinstance(path, str) else None
    if cached is not None:
        return cached
    import pandas as pd
    df = pd.read_csv(path, sep=r"\s+", header=None, comment="#", dtype={0: str}, engine="c")
    if df.shape[0] == 0:
        return []
    chroms = df[0].to_numpy()
    starts = df[1].to_numpy(dtype=np.int64)
    ends = df[2].to_numpy(dtype=np.int64)
    vals = df.iloc[:, 3:].to_numpy(dtype=np.int64)
    out = []
    change = np.nonzero(chroms[1:] != chroms[:-1])[0] + 1
    bounds = [0] + change.tolist() + [len(chroms)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        out.append((chroms[a], starts[a:b], ends[a:b], vals[a:b]))
    return out


def stack_matrix(inBedCount, window_size=100000):
    """stack short bins (Circos.py:831-842) -> (coords [(chrom, start, end)], counts [[int]*S])"""
    groups = _parse_bin_file(inBedCount)
    # window ids in first-seen order per (chrom, BIN) with chromosomes in first-seen order
    order = {}
    line_counts, line_window = [], []
    for chrom, starts, ends, vals in groups:
        bins = (starts // window_size).astype(np.int64) if float(window_size).is_integer() else \
            np.array([int(s // window_size) for s in starts.tolist()], dtype=np.int64)
        d = order.setdefault(chrom, {})
        for b in bins.tolist():
            if b not in d:
                d[b] = None
        line_counts.append(vals)
        line_window.append((chrom, bins))
    coords, index = [], {}
    for chrom, d in order.items():
        for b in d:
            index[(chrom, b)] = len(coords)
            start = int(b * window_size)
            coords.append((chrom, start, int(start + window_size)))
    if not coords:
        return [], []
    lw = np.concatenate([np.array([index[(chrom, b)] for b in bins.tolist()], dtype=np.int32)
                         for chrom, bins in line_window])
    lc = np.concatenate(line_counts, axis=0)
    out = engine.stack_windows(lc, lw, len(coords))
    return coords, out.tolist()


def abnormal(data, k=1.5, high_tile=99, low_tile=1):
    """Circos.py:973-980: the clipping bounds of a density track (99th / 1st percentile, numpy 'linear')."""
    return np.percentile(data, high_tile), np.percentile(data, low_tile)


def stack_bed_density(inBedCount, outpre, colnames, window_size=100000, trim=True):
    """Circos.py:777-806: one circos density track per subgenome from the 10-kb bin table — windows stacked on the
    device (`stack_matrix`), every column clipped at its 99th percentile, written as `chrom start end count`.
    Returns {subgenome: file}."""
    import sys
    coords, counts = stack_matrix(inBedCount, window_size=window_size)
    colnames = list(colnames)
    d_outfiles = {key: "{}.{}.txt".format(outpre, key) for key in colnames}
    arr = np.array(counts, dtype=np.int64).reshape(len(coords), -1) if coords else np.zeros((0, len(colnames)), np.int64)
    assert arr.shape[0] == 0 or arr.shape[1] == len(colnames), "{} != {}".format(len(colnames), arr.shape[1])
    uppers = {}
    if trim:
        for j, key in enumerate(colnames):
            upper, _ = abnormal(arr[:, j])          # np.percentile of an empty column raises, like the reference
            uppers[key] = upper
            print("using cutoff: upper {} for {}".format(upper, key), file=sys.stderr)
    prefix = ["{} {} {} ".format(*c) for c in coords]
    for j, key in enumerate(colnames):
        col = arr[:, j].tolist()
        if trim:
            up = uppers[key]
            col = [c if c <= up else up for c in col]        # min(count, upper): the float bound prints as a float
        with open(d_outfiles[key], "w") as f:
            f.write("".join(p + str(c) + "\n" for p, c in zip(prefix, col)))
    return d_outfiles
