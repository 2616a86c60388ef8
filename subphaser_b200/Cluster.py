"""Drop-in for subphaser/Cluster.py (reference v1.2.7): same constructor and methods `__main__.py`
uses (`Cluster(...)` :18, `.d_sg`, `.sg_names`, `.output_subgenomes` :144, `.output_kmers` :151,
`.pca` :48), with the numerics in libspk.so:

  normalize_data  -> spk_zscore_rows            (population z-score per k-mer, Cluster.py:76-80)
  fit (KMeans)    -> spk_gram + spk_kmeans_gram (k-means++ / Lloyd, n_init=10 as sklearn 0.24.2)
  bootstrap       -> spk_gram_batched + spk_kmeans_gram + spk_cluster_scores (Cluster.py:82-112)
  _output_kmers   -> spk_ttest_groups           (scipy ttest_ind per k-mer, Cluster.py:178-194)
  pca             -> spk_pca_gram               (exact PCA from the n x n Gram matrix)

The reference is unseeded (Cluster.py:49,90,116); here `seed` (env SPK_SEED, default 0) makes every
run reproducible, and `resample_idx` lets a test inject the bootstrap indices.
"""
import itertools
import logging
import os
import sys
from collections import OrderedDict
from collections.abc import Mapping

import numpy as np

from . import _registry, engine, kmer_codec
from .Data import LoadData

logger = logging.getLogger("subphaser_b200")


class _KMeansResult:
    """What the reference keeps of sklearn's fitted KMeans."""

    def __init__(self, labels, inertia, cluster):
        self.labels_ = np.asarray(labels)
        self.inertia_ = float(inertia)
        self._cluster = cluster
        self._centers = None

    @property
    def cluster_centers_(self):
        if self._centers is None:
            c = self._cluster
            self._centers = engine.centroids(c._Z, self.labels_, c.n_clusters).cpu().numpy()
        return self._centers


class KmerSGMap(Mapping):
    """The reference's `d_kmers` dict (kmer -> SG, revcomp(kmer) -> SG; Cluster.py:174-175) backed by
    sorted canonical keys; also carries the device-ready arrays for Seqs.map_kmer3."""

    def __init__(self, keys, sg_idx, sg_names, k):
        order = np.argsort(keys, kind="stable")
        self.keys_arr = np.asarray(keys, dtype=np.uint64)[order]
        self.sg_idx = np.asarray(sg_idx, dtype=np.uint8)[order]
        self.sg_names = list(sg_names)
        self.k = int(k)
        rc = kmer_codec.revcomp_keys(self.keys_arr, self.k) if len(self.keys_arr) else self.keys_arr
        self._n_pal = int(np.sum(rc == self.keys_arr))

    def __len__(self):
        return 2 * len(self.keys_arr) - self._n_pal

    def _lookup(self, kmer):
        if not isinstance(kmer, str) or len(kmer) != self.k:
            raise KeyError(kmer)
        key, valid = kmer_codec.strs_to_keys([kmer], self.k)
        if not valid[0] or kmer != kmer.upper():
            raise KeyError(kmer)
        canon = kmer_codec.canonical_keys(key, self.k)[0]
        i = int(np.searchsorted(self.keys_arr, canon))
        if i < len(self.keys_arr) and self.keys_arr[i] == canon:
            return i
        raise KeyError(kmer)

    def __getitem__(self, kmer):
        return self.sg_names[self.sg_idx[self._lookup(kmer)]]

    def __iter__(self):
        fwd = kmer_codec.keys_to_strs(self.keys_arr, self.k)
        for s in fwd:
            yield s
            r = kmer_codec.revcomp_str(s)
            if r != s:
                yield r

    def values(self):
        rc = kmer_codec.revcomp_keys(self.keys_arr, self.k) if len(self.keys_arr) else self.keys_arr
        mult = np.where(rc == self.keys_arr, 1, 2)
        return list(itertools.chain.from_iterable(
            itertools.repeat(self.sg_names[i], int(m)) for i, m in zip(self.sg_idx, mult)))


class Cluster:
    def __init__(self, datafile, n_clusters, sg_prefix="SG", sg_color=None,
                 sg_assigned={}, re_assign=True,  # use priors
                 bootstrap=True, replicates=1000, jackknife=80, seed=None, resample_idx=None, **kargs):
        import torch
        engine.require_cuda()
        self.seed = int(os.environ.get("SPK_SEED", "0")) if seed is None else int(seed)
        self._resample_idx = resample_idx
        dm = _registry.get_matrix(datafile) if isinstance(datafile, str) else None
        if dm is not None:           # produced by JellyfishDumps.write_matrix in this process
            self.chrs = list(dm.labels)
            self._X = dm.norm
            self._keys = engine.u64_numpy(dm.keys)
            self._d_keys = dm.keys
            self._k = dm.k
            self._kmers = None
        else:
            data = LoadData(datafile)
            data.load_matrix()
            self.chrs = data.colnames
            self._kmers = data.rownames
            self._X = torch.from_numpy(data.data).to(engine._dev())
            self._k = len(self._kmers[0]) if self._kmers else 0
            self._keys = None
        self._Z = engine.zscore_rows(self._X)              # [M, n]; reference self.data is Z.T
        self.n_clusters = n_clusters
        self.sg_prefix = sg_prefix
        self.kargs = kargs
        if sg_assigned:
            logger.info("Skip k-means clustering")
            labels = [sg_assigned[chr] for chr in self.chrs]
            self.n_clusters = len(set(sg_assigned.values()))
            if re_assign:
                self.d_sg = self.assign_subgenomes(labels=labels)
            else:
                self.d_sg = sg_assigned
        else:
            self.kmean = self.fit(None, n_clusters, **kargs)
            self.d_sg = self.assign_subgenomes()
        if bootstrap:
            self.d_bs = self.bootstrap(replicates, jackknife)

    # -- lazily materialised host views (the reference exposes them as attributes) --
    @property
    def kmers(self):
        if self._kmers is None:
            self._kmers = kmer_codec.keys_to_strs(self._keys, self._k)
        return self._kmers

    @property
    def raw_data(self):
        return self._X.cpu().numpy()

    @property
    def data(self):
        return self._Z.cpu().numpy().transpose()

    @property
    def d_kmers(self):
        return dict(zip(self.kmers, self.raw_data.tolist()))

    def normalize_data(self, data, axis=0):
        """Z normalization: only axis=0 work (Cluster.py:76-80); data is [rows, cols] host array."""
        import torch
        X = torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64).T)).to(engine._dev())
        return engine.zscore_rows(X).cpu().numpy().T

    def _name_order(self):
        # chromosome indices in the order sort_subgenomes visits them (Cluster.py:122)
        return [i for _, i in sorted(zip(self.chrs, range(len(self.chrs))), key=lambda x: x[0])]

    def fit(self, data, n_clusters, **kargs):
        """fit KMeans cluster (Cluster.py:114-118): sklearn 0.24.2 defaults (k-means++, n_init=10,
        max_iter=300).  data=None means the full z-scored matrix."""
        import torch
        if data is None:
            Z = self._Z
        else:   # [n, B] host array as the reference passes
            Z = torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64).T)).to(engine._dev())
        self._G = G = engine.gram(Z)
        labels, inertia = engine.kmeans_gram(G, n_clusters, order=None, n_init=10, max_iter=300,
                                             seed=self.seed)
        return _KMeansResult(labels[0].cpu().numpy(), inertia[0].item(), self)

    def bootstrap(self, replicates=1000, jackknife=80):
        """Cluster.py:82-112.  NOTE (reference quirk kept): each replicate resamples `replicates`
        k-mers with replacement; `jackknife` is computed but never used (:85,:90)."""
        import torch
        logger.info("Performing bootstrap of {} replicates, with each replicate resampling {}% data"
                    " with replacement".format(replicates, jackknife))
        M, n = self._Z.shape
        jackknife = max(int(jackknife / 100 * M), 100)     # unused, as in the reference
        if replicates <= 0 or M == 0:
            return {chr: 0 for chr in self.chrs}
        if self._resample_idx is not None:
            idx = np.asarray(self._resample_idx, dtype=np.int32)
            d_idx = torch.from_numpy(np.ascontiguousarray(idx)).to(engine._dev())
        else:   # sklearn.utils.resample(replace=True): randint(0, M, size=n_samples) per replicate
            d_idx = engine.resample_indices(M, replicates, self.seed)
        G = engine.gram_batched(self._Z, d_idx)
        labels, _ = engine.kmeans_gram(G, self.n_clusters, order=self._name_order(), n_init=10,
                                       max_iter=300, seed=self.seed + 1)
        ari, vm = engine.cluster_scores(self.labels, labels)
        xlabels = labels.cpu().numpy()
        self.boot_labels = xlabels
        d_bs = {}
        R = xlabels.shape[0]
        for i, (label, chr) in enumerate(zip(self.labels, self.chrs)):
            bs = int(np.sum(xlabels[:, i] == label))
            d_bs[chr] = int(100 * bs / R)
        self.mean_adjusted_rand_score = float(np.mean(ari.cpu().numpy()))
        self.mean_v_measure_score = float(np.mean(vm.cpu().numpy()))
        logger.info("Bootstrap: mean Adjusted Rand-Index: {:.4f}; mean V-measure score: {:.4f}".format(
            self.mean_adjusted_rand_score, self.mean_v_measure_score))
        return d_bs

    def sort_subgenomes(self, labels):
        assert len(self.chrs) == len(labels)
        d_map = {}
        for label, chr in sorted(zip(labels, self.chrs), key=lambda x: x[1]):
            if label not in d_map:
                try:
                    d_map[label] = max(d_map.values()) + 1
                except ValueError:
                    d_map[label] = 0
        return [d_map[label] for label in labels]

    def assign_subgenomes(self, base=1, labels=None):
        """labels is the same order as self.chrs (Cluster.py:128-143)"""
        if labels is None:
            labels = list(self.kmean.labels_)
        max_len = len(str(self.n_clusters))
        fmtstr = "{{}}{{:0>{}d}}".format(max_len)
        d_sg = OrderedDict()
        sg_names = set([])
        self.labels = labels = self.sort_subgenomes(labels)
        assert len(self.chrs) == len(labels)
        for label, chr in zip(labels, self.chrs):
            sg = fmtstr.format(self.sg_prefix, label + base)
            d_sg[chr] = sg
            sg_names.add(sg)
        self.sg_names = sorted(sg_names)
        return d_sg

    def output_subgenomes(self, fout=sys.stdout):
        line = ["#chrom", "subgenome", "bootstrap"]
        print("\t".join(line), file=fout)
        for chr, sg in sorted(self.d_sg.items(), key=lambda x: x[1]):
            line = [chr, sg, self.d_bs[chr]]
            print("\t".join(map(str, line)), file=fout)

    def output_kmers(self, fout=sys.stdout, max_pval=0.05, ncpu=4, method="map", test_method="ttest_ind"):
        """Cluster.py:151-176 -> KmerSGMap (the d_kmers mapping)."""
        if test_method not in ("ttest_ind", "kruskal", "mannwhitneyu", "wilcoxon"):     # the CLI's choices (__main__.py)
            raise AttributeError("module 'scipy.stats' has no attribute {!r}".format(test_method))   # eval() at :160
        sgs = sorted(set(self.d_sg.values()))                 # groups in sorted-SG order (:180)
        col_group = [sgs.index(self.d_sg[chr]) for chr in self.chrs]
        S = len(sgs)
        if S < 2:
            raise IndexError("list index out of range")      # grouped[1] in the reference (:187)
        if test_method == "ttest_ind":
            if min(col_group.count(g) for g in range(S)) == 1:
                logger.warning("a subgenome holds a single chromosome: scipy 1.7.1 (pinned by SubPhaser) gives NaN p-values "
                               "for tests against it and such k-mers are kept (Cluster.py:167); set "
                               "SPK_TTEST_SINGLETON=zero for the behaviour of scipy >= 1.9")
            best, pval, means = engine.ttest_groups(self._X, col_group, S)
        else:
            best, pval, means, flags = engine.ranktest_groups(self._X, col_group, S, test_method)
            if flags & 1:       # what scipy raises inside the reference's Pool workers
                raise ValueError("The samples x and y must have the same length.")
            if flags & 2:
                raise ValueError("All numbers are identical in kruskal")
        import torch
        d_keep = ~(pval > max_pval)                           # NaN p-values are kept (:167)
        d_idx = torch.nonzero(d_keep).flatten().to(torch.int32)
        print("\t".join(["#kmer", "subgenome", "p_value", "ratios"]), file=fout)
        if self._keys is None:                                # matrix came from a text file: encode its k-mers once
            keys_np, valid = kmer_codec.strs_to_keys(self.kmers, self._k)
            if not valid.all():
                raise ValueError("non-ACGT k-mer in {}".format("the matrix file"))
            self._keys = keys_np
        d_keys = getattr(self, "_d_keys", None)
        if d_keys is None:
            d_keys = self._d_keys = torch.from_numpy(self._keys.view(np.int64).copy()).to(engine._dev())
        # rows are formatted on the device (spk_format_rows kind 1): kmer, subgenome, p-value, comma-joined group means
        engine.write_text(fout, engine.format_rows(d_keys, means, self._k, kind=1, rows=d_idx, label=best,
                                                   label_names=sgs, pval=pval))
        idxs = d_idx.cpu().numpy().astype(np.int64)
        return KmerSGMap(self._keys[idxs], best.cpu().numpy()[idxs], sgs, self._k)

    def pca(self, outfig, n_components=2, sg_color=None):
        """Cluster.py:48-75.  Scores/explained variance come from the exact Gram-matrix PCA; they are
        written to `<outfig>.tsv`, and plotted when matplotlib is available."""
        G = getattr(self, "_G", None)
        if G is None:
            G = engine.gram(self._Z)
        eig, scores, ratio = engine.pca_gram(G, n_components)
        X_pca = scores.cpu().numpy()
        percent = ratio.cpu().numpy() * 100
        X_pca = (X_pca - X_pca.mean(axis=0)) / X_pca.std(axis=0)   # Cluster.py:52
        self.pca_scores, self.pca_percent = X_pca, percent
        with open(outfig + ".tsv", "w") as f:
            f.write("#chrom\tsubgenome\t" + "\t".join(
                "PC{} ({:.1f}%)".format(i + 1, p) for i, p in enumerate(percent)) + "\n")
            for chr, row in zip(self.chrs, X_pca):
                f.write("\t".join([chr, str(self.d_sg[chr])] + [repr(float(v)) for v in row]) + "\n")
        try:
            from matplotlib import pyplot as plt
        except Exception:
            logger.warning("matplotlib not available: wrote {}.tsv instead of the figure".format(outfig))
            return
        plt.switch_backend("agg")
        plt.figure(figsize=(7, 7), dpi=300, tight_layout=True)
        d_coord = {}
        for _x, _y, _c, _l in zip(X_pca[:, 0], X_pca[:, 1], self.chrs, self.labels):
            sg = self.d_sg[_c]
            if sg not in d_coord:
                d_coord[sg] = [[], [], sg_color.colors_hex[_l] if sg_color is not None else None]
            d_coord[sg][0] += [_x]
            d_coord[sg][1] += [_y]
        for sg, (_x, _y, _c) in sorted(d_coord.items()):
            plt.scatter(_x, _y, c=_c, marker="o", label=sg)
        plt.axhline(0, ls="--", c="grey")
        plt.axvline(0, ls="--", c="grey")
        plt.xlabel("PC1 ({:.1f}%)".format(percent[0]))
        plt.ylabel("PC2 ({:.1f}%)".format(percent[1]), ha="center", va="center")
        plt.legend()
        plt.savefig(outfig, bbox_inches="tight", dpi=300)


def _fmt(x):
    """str() of a Python float / numpy float64 as the reference prints it (shortest repr; nan)."""
    return repr(float(x))
