"""The hot-path step sequence of the reference's `Pipeline.run()` (subphaser/__main__.py:403-498),
driven through the drop-in modules of this package: count -> matrix -> filter -> `.kmer.mat` ->
Cluster (+bootstrap) -> specific k-mers -> map to bins -> stack windows -> Fisher enrichment.
It writes the same files with the same names the reference writes.  Used by the tests, smoke() and
bench.py; a SubPhaser installation gets the same effect by swapping the modules."""
import logging
import os
from collections import Counter

from . import Circos, Jellyfish, Seqs, Stats
from .Cluster import Cluster

logger = logging.getLogger("subphaser_b200")


def run_hot_path(chromfiles, labels, sgs, outdir, prefix="", k=15, lower_count=3, min_fold=2, baseline=1,
                 min_freq=200, max_freq=10000, min_prop=None, max_prop=None, ratio=1, nsg=None,
                 sg_assigned=None, replicates=1000, jackknife=80, max_pval=0.05, test_method="ttest_ind",
                 bin_size=10000, map_window=10e6, window_size=1000000, overwrite=True, seed=None,
                 resample_idx=None, ncpu=1):
    os.makedirs(outdir, exist_ok=True)
    res = {}
    logger.info("###Step: Kmer Count")
    dumpfiles = Jellyfish.run_jellyfish_dumps(chromfiles, k=k, ncpu=ncpu, lower_count=lower_count, threads=1,
                                              overwrite=overwrite)
    dumps = Jellyfish.JellyfishDumps(dumpfiles, labels, ncpu=ncpu)
    basename = "k{}_q{}_f{}".format(k, min_freq, min_fold)
    para_prefix = os.path.join(outdir, prefix + basename)
    matfile = para_prefix + ".kmer.mat"
    d_mat = dumps.to_matrix()
    res["kmer_count"] = len(d_mat)
    logger.info("{} kmers in total".format(len(d_mat)))
    lengths = dumps.lengths
    res["lengths"] = list(lengths)
    d_mat = dumps.filter(d_mat, lengths, sgs, outfig=para_prefix + ".kmer_freq.pdf", min_fold=min_fold,
                         baseline=baseline, min_freq=min_freq, max_freq=max_freq, min_prop=min_prop,
                         max_prop=max_prop, ratio=ratio)
    if len(d_mat) == 0:
        raise ValueError("0 kmer remained after filtering. Please reset the filter options.")
    res["n_diff"] = len(d_mat)
    with open(matfile, "w") as fout:
        dumps.write_matrix(d_mat, fout)
    logger.info("###Step: Cluster")
    if nsg is None:
        nsg = max(len(sg) for sg in sgs)
    cluster = Cluster(matfile, n_clusters=nsg, sg_prefix="SG", sg_assigned=sg_assigned or {},
                      replicates=replicates, jackknife=jackknife, seed=seed, resample_idx=resample_idx)
    d_sg = cluster.d_sg
    sg_names = cluster.sg_names
    with open(para_prefix + ".chrom-subgenome.tsv", "w") as fout:
        cluster.output_subgenomes(fout)
    with open(para_prefix + ".sig.kmer-subgenome.tsv", "w") as fout:
        d_kmers = cluster.output_kmers(fout, max_pval=max_pval, ncpu=ncpu, test_method=test_method)
    logger.info("{} significant subgenome-specific kmers".format(len(d_kmers) // 2))
    for sg, count in sorted(Counter(d_kmers.values()).items()):
        logger.info("\t{} {}-specific kmers".format(count // 2, sg))
    cluster.pca(para_prefix + ".kmer_pca.pdf", n_components=nsg, sg_color=None)
    sg_map = para_prefix + ".subgenome.bin.count"
    with open(sg_map, "w") as fout:
        Seqs.map_kmer3(chromfiles, d_kmers, fout=fout, k=k, bin_size=bin_size, sg_names=sg_names, ncpu=ncpu,
                       method="map", window_size=map_window)
    bins, counts = Circos.stack_matrix(sg_map, window_size=window_size)
    with open(para_prefix + ".bin.enrich", "w") as fout, open(para_prefix + ".bin.group", "w") as fout2:
        sg_lines = Stats.enrich_bin(fout, fout2, d_sg, counts, colnames=sg_names, rownames=bins,
                                    max_pval=max_pval, ncpu=ncpu)
    res.update(para_prefix=para_prefix, cluster=cluster, d_sg=d_sg, sg_names=sg_names, d_kmers=d_kmers,
               bins=bins, counts=counts, sg_lines=sg_lines, matfile=matfile)
    return res
