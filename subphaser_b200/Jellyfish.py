"""Drop-in for the hot-path surface of subphaser/Jellyfish.py (reference v1.2.7).

Same names, argument meaning and error behaviour as the reference functions `__main__.py` calls
(`run_jellyfish_dumps` :671, `JellyfishDumps` :430 with `.to_matrix` :439, `.filter` :462,
`.write_matrix` :515, `.heatmap` :521), but the arithmetic runs in libspk.so on the GPU:

  run_jellyfish_dump  : FASTA -> K1 pack -> K2 canonical count -> K3 dump (>= lower_count)
                        instead of `jellyfish count | histo | dump` (Jellyfish.py:697-700)
  to_matrix           : union hash table + count matrix on the device (d_mat never becomes a dict)
  filter              : _filter_kmer (Jellyfish.py:611-648) as one kernel over all rows

File contract kept: `<prefix>_<k>.fa` is returned per chromosome and `<prefix>_<k>.fa.ok` marks it done;
the dump content itself is kept on the device and in a binary side-car `<prefix>_<k>.fa.spk.npz`
(10^8-line text dumps are what the GPU path exists to avoid; set SPK_TEXT_DUMPS=1 to also write the
byte-compatible `KMER COUNT` text).  The unused reads/KMC helpers of the reference module
(Jellyfish.py:100-428, 705-824) are out of scope.
"""
import logging
import os
import sys
from collections import OrderedDict

import numpy as np

from . import _registry, engine, kmer_codec

logger = logging.getLogger("subphaser_b200")

SIDE_SUFFIX = ".spk.npz"


def _touch(path):
    with open(path, "a"):
        os.utime(path, None)


def _write_text_dump(path, keys, counts, k):
    strs = kmer_codec.keys_to_strs(keys, k)
    with open(path, "w") as f:
        for s, c in zip(strs, counts.tolist()):
            f.write("%s %d\n" % (s, c))


def _parse_text_dump(path, k=None):
    """A text dump written by real jellyfish (`KMER COUNT` per line, Jellyfish.py:19-25)."""
    seqs, freqs = [], []
    opener = open
    if path.endswith(".gz"):
        import gzip
        opener = gzip.open
    with opener(path, "rt") as f:
        for line in f:
            t = line.split()
            if len(t) < 2:
                continue
            seqs.append(t[0])
            freqs.append(int(t[1]))
    if not seqs:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint32), (k or 0)
    k = len(seqs[0])
    keys, valid = kmer_codec.strs_to_keys(seqs, k)
    if not valid.all():
        raise ValueError("non-ACGT k-mer in dump file {}".format(path))
    return keys, np.array(freqs, dtype=np.uint32), k


def run_jellyfish_dumps(seqfiles, ncpu=4, **kargs):
    """Jellyfish.py:671-676.  Chromosomes are counted one after another on the current GPU (each count
    already fills the device); `ncpu` is accepted for compatibility."""
    dumpfiles = []
    # one scratch table sized for the largest chromosome: every dump then uses the same hash partitions and
    # carries a partition index, which is what JellyfishDumps.to_matrix/filter merge on chip
    table = None
    try:
        sizes = [os.path.getsize(f) for f in seqfiles if isinstance(f, str) and not f.endswith(".gz")]
        if len(sizes) == len(seqfiles) and sizes and "k" in kargs:
            table = engine.CountTable(max(sizes), int(kargs["k"]), int(kargs.get("lower_count", 2)))
    except (OSError, engine._lib.SpkError):
        table = None
    for seqfile in seqfiles:
        dumpfiles += [run_jellyfish_dump(seqfile, _table=table, **kargs)]
    return dumpfiles


def run_jellyfish_dump(seqfile, threads=4, k=17, prefix=None, lower_count=2, method="jellyfish",
                       overwrite=False, _table=None):
    """Jellyfish.py:681-704 without the shell-out.  Returns the dump path `<prefix>_<k>.fa`."""
    if isinstance(seqfile, (list, tuple, set)):
        files = list(seqfile)
        _seqfile = files[0]
    else:
        files = [seqfile]
        _seqfile = seqfile
        if prefix is None:
            prefix = seqfile
    output = "{prefix}_{KMER}.fa".format(KMER=k, prefix=prefix)
    # Checkpoints.  `<output>.ok` keeps the reference's meaning — a jellyfish-style text dump is complete at `output`
    # (Jellyfish.py:691-694) — and is only touched when that text is written (SPK_TEXT_DUMPS=1).  The binary side-car
    # (SPK_DUMP_SIDECAR=1; 12 bytes per dumped k-mer) has its own marker `<output>.spk.ok`, so a later stock SubPhaser
    # run in the same directory never mistakes an empty placeholder for a finished dump.  Without either, a re-run
    # simply recounts (the GPU counts a chromosome faster than a multi-GB dump loads).
    ckp_file = output + ".ok"
    side = output + SIDE_SUFFIX
    side_ckp = output + ".spk.ok"
    src_sig = tuple(_registry._sig(f) for f in files)
    if not overwrite:
        if _registry.get_dump(output, k=int(k), lower_count=int(lower_count), src_sig=src_sig) is not None:
            return output
        if os.path.exists(side_ckp) and os.path.exists(side):
            with np.load(side) as z:
                if int(z["k"]) == int(k) and int(z["lower_count"]) == int(lower_count):
                    return output       # loaded lazily by JellyfishDumps
        if os.path.exists(ckp_file) and os.path.exists(output) and os.path.getsize(output) > 0:
            return output               # a text dump left by real jellyfish (or SPK_TEXT_DUMPS=1)
    engine.require_cuda()
    # a chromosome that Seqs.split_genomes produced in this process is already packed on the device
    seq = _registry.get_seq(files[0]) if len(files) == 1 else None
    if seq is None:
        bufs = [engine.read_fasta_bytes(f) for f in files]
        if len(bufs) > 1:
            nl = np.frombuffer(b"\n", dtype=np.uint8)
            parts = []
            for b in bufs:
                parts += [b, nl]
            buf = np.concatenate(parts)
        else:
            buf = bufs[0]
        d_ascii, nbytes = engine.to_device_bytes(buf)
        seq = engine.pack_fasta(d_ascii, nbytes, name=os.path.basename(_seqfile))
        del d_ascii
    dump = engine.count_packed(seq, int(k), int(lower_count), table=_table, histo_len=100002)
    if len(files) == 1:
        _registry.put_seq(files[0], seq)
    _registry.put_dump(output, dump, k=int(k), lower_count=int(lower_count), src_sig=src_sig)
    want_text = os.environ.get("SPK_TEXT_DUMPS") == "1"
    want_side = os.environ.get("SPK_DUMP_SIDECAR") == "1"
    for stale in (ckp_file, side_ckp) + (() if want_side else (side,)):     # an earlier run's files describe other content now
        if os.path.exists(stale):
            os.remove(stale)
    if want_text or want_side:
        keys, counts = dump.to_host()
    if want_side:
        extra = {}
        if dump.pindex is not None:
            extra = dict(pindex=dump.pindex.cpu().numpy(), pbits=np.int64(dump.pbits))
        np.savez(side, keys=keys, counts=counts, k=np.int64(k), lower_count=np.int64(lower_count),
                 length=np.int64(dump.length), n_valid_kmers=np.int64(dump.n_valid_kmers),
                 n_distinct=np.int64(dump.n_distinct), **extra)
        _touch(side_ckp)
    histo = dump.histo.cpu().numpy()
    with open("{}_{}.histo".format(prefix, k), "w") as f:
        for c in np.nonzero(histo)[0]:
            if c > 0:
                f.write("%d %d\n" % (c, histo[c]))
    if want_text:
        _write_text_dump(output, keys, counts, int(k))
        _touch(ckp_file)
    elif not os.path.exists(output):
        _touch(output)                  # the path `__main__.py` carries around; its content lives on the device
    logger.info("Counted {}: {:,} bases, {:,} k-mers, {:,} distinct, {:,} with count >= {}".format(
        _seqfile, seq.n_bases, dump.n_valid_kmers, dump.n_distinct, len(dump), lower_count))
    return output


def load_dump(dumpfile):
    """dump path -> engine.KmerDump on the device (registry, binary side-car, or jellyfish text)."""
    import torch
    d = _registry.get_dump(dumpfile)
    if d is not None:
        return d
    engine.require_cuda()
    pindex, pbits = None, 0
    side = dumpfile + SIDE_SUFFIX
    if os.path.exists(side) and os.path.exists(dumpfile + ".spk.ok"):
        with np.load(side) as z:
            keys, counts, k = z["keys"], z["counts"], int(z["k"])
            length = int(z["length"])
            nv, nd = int(z["n_valid_kmers"]), int(z["n_distinct"])
            if "pindex" in z.files:
                pindex, pbits = z["pindex"], int(z["pbits"])
    else:
        keys, counts, k = _parse_text_dump(dumpfile)
        length = int(counts.sum(dtype=np.int64))
        nv, nd = length, len(keys)
    dev = engine._dev()
    dk = torch.from_numpy(keys.view(np.int64).copy()).to(dev)
    dc = torch.from_numpy(counts.view(np.int32).copy()).to(dev)
    dp = torch.from_numpy(np.ascontiguousarray(pindex).view(np.int32).copy()).to(dev) if pindex is not None else None
    d = engine.KmerDump(dk, dc, k, length, nv, nd, os.path.basename(dumpfile), None, dp, pbits)
    _registry.put_dump(dumpfile, d)
    return d


class JellyfishDumps:
    """Jellyfish.py:430-522."""

    def __init__(self, dumpfiles, labels=None, ncpu=4, method="map", chunksize=None, **kargs):
        self.dumpfiles = dumpfiles
        self.labels = labels
        self.ncpu = ncpu
        self.method = method
        self.chunksize = chunksize

    def __len__(self):
        return len(self.dumpfiles)

    def to_matrix(self, array=False):
        """-> engine.CountMatrix (len() = number of distinct dumped k-mers); sets self.lengths."""
        dumps = []
        for dumpfile in self.dumpfiles:
            logger.info("Loading " + dumpfile)
            dumps.append(load_dump(dumpfile))
        cm = None
        if engine.can_pmatrix(dumps):      # dumps counted with common partition bits: no union table at all
            cm = engine.LazyUnion(dumps, self.labels)
            try:
                len(cm)                    # `__main__.py:417` asks for it right away: one union-only pass
            except OverflowError:          # a hash partition did not fit the shared-memory table
                cm = None
        if cm is None:
            cm = engine.build_matrix(dumps, self.labels)
        self.lengths = list(cm.lengths)
        return cm

    def filter(self, d_mat, lengths, sgs, outfig=None, by_count=False,
               min_freq=200, max_freq=10000, min_fold=2, baseline=1,
               min_prop=None, max_prop=None, ratio=1):
        """Jellyfish.py:462-512 -> engine.DiffMatrix (len() = number of differential k-mers)."""
        tot_lens = sum(self.lengths)
        if min_prop is not None:
            min_freq = min_prop * tot_lens
            logger.info("Adjust `min_freq` to {} according to `min_prop`".format(min_freq))
        if max_prop is not None:
            max_freq = max_prop * tot_lens
            logger.info("Adjust `max_freq` to {} according to `max_prop`".format(max_freq))
        if min_freq > max_freq:
            raise ValueError("`min_freq` ({}) should be lower than `max_freq` ({})".format(min_freq, max_freq))
        i = 0
        for sg in sgs:
            if len(sg) == 1:
                logger.warning("Singleton `{}` is ignored".format(sg))
                i += 1
        if i == len(sgs):
            raise ValueError("All singletons are not allowed")
        d_lens = OrderedDict(zip(self.labels, self.lengths))
        lens0 = [lab for lab, _len in d_lens.items() if _len == 0]
        if lens0:
            raise ValueError("Chromosomes `{}` have only 0 kmers".format(lens0))
        for sg in sgs:   # freqs[baseline] must exist (the reference would raise IndexError per k-mer)
            if len(sg) > 1 and not (-len(sg) <= baseline < len(sg)):
                raise IndexError("list index out of range")
            if len(sg) > 32:
                raise ValueError("more than 32 groups in one homoeologous set is not supported")
        if isinstance(d_mat, engine.LazyUnion):
            try:
                dm, n_all = engine.pmatrix_filter(d_mat.dumps, sgs, list(self.labels), min_fold=min_fold,
                                                  baseline=baseline, ratio=ratio, min_freq=min_freq,
                                                  max_freq=max_freq, by_count=by_count,
                                                  want_fold_tots=outfig is not None, lengths=list(lengths))
                d_mat._len = n_all
            except OverflowError:      # a hash partition did not fit the shared-memory table: plain union
                d_mat = engine.build_matrix(d_mat.dumps, self.labels)
        if isinstance(d_mat, engine.LazyUnion):
            pass
        elif isinstance(d_mat, engine.CountMatrix):
            if list(lengths) != list(d_mat.lengths):
                d_mat = engine.CountMatrix(d_mat.matrix, d_mat.row_keys, lengths, d_mat.k, d_mat.labels)
            n_all = len(d_mat)
            dm = engine.filter_matrix(d_mat, sgs, list(self.labels), min_fold=min_fold, baseline=baseline,
                                      ratio=ratio, min_freq=min_freq, max_freq=max_freq, by_count=by_count,
                                      want_fold_tots=outfig is not None)
        else:
            raise TypeError("d_mat must come from JellyfishDumps.to_matrix()")
        remain = len(dm)
        # without outfig the reference gates on frequency first and only counts kept k-mers (:617,502)
        total = dm.n_fold_pass if outfig is not None else remain
        logger.info("After filtering, remained {} ({:.2%}) differential (freq >= {}) and {} ({:.2%}) "
                    "candidate (freq > 0) kmers".format(remain, remain / max(n_all, 1), min_freq, total,
                                                        total / max(n_all, 1)))
        if outfig is not None:
            if total == 0:
                raise ValueError("0 kmer with fold > {}. Please reset the filter options.".format(min_fold))
            logger.info("Plot " + outfig)
            plot_histogram(dm.fold_tots, outfig, vline=None)
        return dm

    def write_matrix(self, d_mat, fout):
        """Jellyfish.py:515-520: header `kmer<TAB>labels`, rows `KMER<TAB>str(float)...`."""
        fout.write("\t".join(["kmer"] + list(self.labels)) + "\n")
        if isinstance(d_mat, engine.DiffMatrix):
            # rows are formatted on the device (k-mer strings, shortest-repr floats: spk_format_rows)
            engine.write_text(fout, engine.format_rows(d_mat.keys, d_mat.norm, d_mat.k, kind=0))
            fout.flush()
            name = getattr(fout, "name", None)
            if isinstance(name, str) and os.path.exists(name):
                _registry.put_matrix(name, d_mat)
        else:   # a plain mapping kmer -> list, as the reference builds
            for kmer, counts in d_mat.items():
                fout.write("\t".join(map(str, [kmer] + list(counts))) + "\n")

    def heatmap(self, matfile, **kargs):
        return _heatmap(matfile, **kargs)


def _heatmap(matfile, **kargs):
    """Jellyfish.py:524-609 is an R script (visualisation, outside the accelerated path).  If the
    reference package is installed its own implementation is used; otherwise the plot is skipped."""
    try:
        from subphaser.Jellyfish import _heatmap as ref_heatmap   # noqa: the real reference, if present
    except Exception:
        logger.warning("heatmap skipped: needs the reference package and Rscript (visualisation only)")
        return None
    return ref_heatmap(matfile, **kargs)


def plot_histogram(data, outfig, step=25, xlim=99, xlabel="Kmer occurrence", ylabel="Count", vline=None):
    """Jellyfish.py:650-666.  `data`: the totals themselves (any sequence) or an engine.FoldHistogram (binned on the
    device).  The binned counts always go to `<outfig>.tsv`; the figure is drawn when matplotlib is available."""
    if isinstance(data, engine.FoldHistogram):
        hist, edges, limit = data.hist, data.edges, data.xlim
    else:
        data = np.asarray(data)
        _max = int(data.max()) if data.size else 0
        nbins = max(int((_max - 0) / step), 1)
        hist, edges = np.histogram(data, bins=nbins)
        limit = float(np.percentile(data, xlim)) if data.size else 0.0
    with open(outfig + ".tsv", "w") as f:
        f.write("#bin_start\tbin_end\tcount\n")
        for a, b, c in zip(edges[:-1], edges[1:], hist):
            f.write("%g\t%g\t%d\n" % (a, b, c))
        f.write("#xlim (percentile %g)\t%r\n" % (xlim, limit))
    try:
        from matplotlib import pyplot as plt
    except Exception:
        logger.warning("matplotlib not available: wrote {}.tsv instead of the figure".format(outfig))
        return
    plt.switch_backend("agg")
    plt.figure(figsize=(7, 5), dpi=300, tight_layout=True)
    plt.bar(edges[:-1], hist, width=np.diff(edges), align="edge")       # the bars plt.hist(data, bins=nbins) draws
    plt.xlim(0, limit)
    plt.xlabel(xlabel)
    plt.ylabel(ylabel, ha="center", va="center")
    plt.ticklabel_format(style="plain")
    if vline is not None:
        plt.axvline(vline, ls="--", c="grey")
    plt.savefig(outfig, bbox_inches="tight", dpi=300)
