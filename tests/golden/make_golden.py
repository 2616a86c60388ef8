"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own Python code
(/root/reference/subphaser, imported unmodified under oracle/ref_shims.py).  Runs only in the build
container; the fixtures it writes are committed and are what the tests (CPU and GPU) read.

    python tests/golden/make_golden.py

Stand-ins, stated once: the jellyfish shell-out is replaced by oracle/kmer_count.c (jellyfish is not
installed; parity with it is unpinned); fisher -> scipy hypergeom.sf; statsmodels fdr_bh -> numpy
formula; sklearn KMeans is called with n_init=10 (the pinned 0.24.2 default) and random_state=0;
sklearn.utils.resample is replaced by a recording equivalent so the bootstrap indices are part of the
fixture.  Plotting functions are stubbed.
"""
import functools
import io
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import spk_testutil as util  # noqa: E402
from oracle import kmers, ref_shims  # noqa: E402


def jdump(obj, name):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f)
    print("wrote", name)


def gen_filter(J):
    rng = np.random.default_rng(1)
    cases = []
    labels13 = [str(i) for i in range(1, 14)]
    arab = [[["1"], ["6", "7"]], [["2", "3"], ["9", "8", "10"]], [["4", "5"], ["13", "11", "12"]]]
    wheat_labels = ["%d%s" % (c, s) for c in range(1, 4) for s in "ABD"]
    wheat = [[["%d%s" % (c, s)] for s in "ABD"] for c in range(1, 4)]
    mixed_labels = ["a", "b", "c", "d", "e"]
    mixed = [[["a"], ["b"]], [["c"]], [["d"], ["e"]]]          # one singleton set
    for labels, sgs, name in ((labels13, arab, "arab"), (wheat_labels, wheat, "wheat"), (mixed_labels, mixed, "mixed")):
        n = len(labels)
        lengths = [int(x) for x in rng.integers(50_000, 5_000_000, n)]
        from collections import OrderedDict
        d_lens = OrderedDict(zip(labels, lengths))
        for params in (dict(min_freq=200, max_freq=10000, min_fold=2, baseline=1, ratio=1),
                       dict(min_freq=50, max_freq=100000, min_fold=1.5, baseline=-1, ratio=0.5),
                       dict(min_freq=0, max_freq=1e9, min_fold=3, baseline=1, ratio=0.34)):
            if name == "mixed" and params["baseline"] == 1:
                pass
            rows = []
            for _ in range(300):
                style = rng.integers(0, 4)
                if style == 0:
                    counts = rng.integers(0, 400, n)
                elif style == 1:
                    counts = rng.integers(0, 5, n) * rng.integers(0, 2, n)
                elif style == 2:
                    counts = np.where(rng.random(n) < 0.4, rng.integers(100, 3000, n), rng.integers(0, 30, n))
                else:
                    base = int(rng.integers(3, 600))
                    counts = np.full(n, base) + rng.integers(0, 3, n)   # near the fold threshold / ties
                counts = [int(c) for c in counts]
                arg = ("K", counts, d_lens, sgs, "fig", False, params["min_freq"], params["max_freq"],
                       params["min_fold"], params["baseline"], params["ratio"])
                kmer, freqs, tot = J._filter_kmer(arg)
                rows.append(dict(counts=counts, freqs=freqs, tot=tot))
            cases.append(dict(name=name, labels=labels, sgs=sgs, lengths=lengths, params=params, rows=rows))
    jdump(cases, "filter_kmer.json")


def gen_fisher_enrich(S_mod):
    rng = np.random.default_rng(2)
    cases = []
    for S in (2, 3, 4):
        for scale in (10, 1000, 100000, 30_000_000):
            W = 40
            mat = rng.integers(0, scale, (W, S))
            mat[0] = 0
            mat[1] = [scale] + [0] * (S - 1)
            mat[2] = mat[3]                       # equal rows -> ties
            mat[4] = [5] * S                      # tied p-values inside a row
            if scale >= 30_000_000:
                mat[:, 0] += 200_000_000          # exercises the MAX_INT clamp (Stats.py:24-25)
            total = [int(x) for x in mat.sum(axis=0)]
            rows = []
            colnames = ["SG%d" % (i + 1) for i in range(S)]
            for r in range(W):
                row = [int(x) for x in mat[r]]
                pvals = S_mod.fisher_test(row, total)
                res = S_mod._enrich((row, ("c", r * 10, r * 10 + 10), total, colnames, 0.5, {"max_pval": 0.05}))
                rows.append(dict(row=row, pvals=[float(p) for p in pvals], idx=int(res.idx), sig=bool(res.sig),
                                 key=res.key, ratios=[float(x) for x in res.ratios], enrich=list(res.enrich)))
            pmin = [r["pvals"][r["idx"]] for r in rows]
            q = [float(x) for x in S_mod.correct_pvals(pmin)]
            cases.append(dict(S=S, scale=scale, total=total, rows=rows, qvals=q))
    jdump(cases, "fisher_enrich.json")


def gen_enrich_ltr(S_mod):
    """Stats.enrich_ltr (Stats.py:33-73) on per-sequence rows: ids `chrom:start-end` (LTRs) and ids of chromosomes
    that are not in d_sg; the 6-column TSV and both returned dicts are the fixture."""
    rng = np.random.default_rng(21)
    cases = []
    for S, nrow, scale in ((2, 60, 40), (3, 80, 300), (3, 25, 5000)):
        colnames = ["SG%d" % (i + 1) for i in range(S)]
        chroms = ["%d%s" % (c + 1, "ABD"[s]) for c in range(3) for s in range(S)]
        d_sg = {c: colnames["ABD".index(c[-1])] for c in chroms}
        d_sg.pop(chroms[-1])                       # a chromosome without an assignment -> obs_sg None
        rownames, matrix = [], []
        for r in range(nrow):
            c = chroms[int(rng.integers(0, len(chroms)))] if r % 11 else "scaffold_%d" % r
            a = int(rng.integers(0, 10_000_000))
            rid = "%s:%d-%d" % (c, a, a + int(rng.integers(100, 9000)))
            own = "ABD".index(c[-1]) if c[-1] in "ABD"[:S] and r % 5 else int(rng.integers(0, S))
            row = rng.integers(0, max(scale // 8, 2), S)
            row[own] += int(rng.integers(0, scale))
            if r % 13 == 0:
                row[:] = 0
                row[int(rng.integers(0, S))] = 1
            rownames.append((rid, 0, 100000000))
            matrix.append([int(x) for x in row])
        matrix = [m for m in matrix]
        keep = [i for i, m in enumerate(matrix) if sum(m) > 0]     # stack_matrix never emits all-zero rows
        rownames = [rownames[i] for i in keep]
        matrix = [matrix[i] for i in keep]
        out = io.StringIO()
        d_enriched, d_exchange = S_mod.enrich_ltr(out, d_sg, matrix, colnames=colnames, rownames=rownames,
                                                  max_pval=0.05, ncpu=1)
        cases.append(dict(S=S, colnames=colnames, d_sg=d_sg, rownames=[list(r) for r in rownames], matrix=matrix,
                          text=out.getvalue(), d_enriched=d_enriched, d_exchange=d_exchange))
    jdump(cases, "enrich_ltr.json")


def gen_stat_enrich():
    """stat_enrich.main (stat_enrich.py:4-37) on 4-column enrich tables (the layout it unpacks, :11)."""
    argv = sys.argv
    sys.argv = [argv[0], os.devnull]               # the reference evaluates sys.argv[1] at definition time (:4)
    try:
        SE = ref_shims.load("stat_enrich")
    finally:
        sys.argv = argv
    rng = np.random.default_rng(22)
    cases = []
    for S, nrow in ((2, 40), (3, 90)):
        sgs = ["SG%d" % (i + 1) for i in range(S)]
        anns = ["Copia", "Gypsy", "Ty3", "unknown", "LINE"]
        lines = ["#id\tsubgenome\tp_value\tcounts"]
        # every (annotation, subgenome) pair occurs at least once: the reference adds a zero vector of length
        # len(subgenomes) for a missing pair, which only broadcasts when that equals the number of count columns
        pairs = [(a, g) for a in anns for g in sgs]
        for r in range(nrow):
            a, g = pairs[r] if r < len(pairs) else (anns[int(rng.integers(0, len(anns)))], sgs[int(rng.integers(0, S))])
            counts = ",".join(str(int(x)) for x in rng.integers(0, 500, S))
            lines.append("%s-%d_%s\t%s\t%r\t%s" % (a, r, "x" * int(rng.integers(1, 4)), g, float(rng.random()), counts))
        text = "\n".join(lines) + "\n"
        tmp = tempfile.mkdtemp()
        path = os.path.join(tmp, "enrich.tsv")
        with open(path, "w") as f:
            f.write(text)
        out = io.StringIO()
        SE.main(inTsv=path, outStat=out)
        shutil.rmtree(tmp)
        cases.append(dict(input=text, output=out.getvalue()))
    jdump(cases, "stat_enrich.json")


def gen_split_genomes(Seqs):
    """Seqs.split_genomes (Seqs.py:27-71): two genome files (one gzipped, one with 70-column CRLF lines), prefixes,
    renamed targets (`new|old`), records that are not targets.  The per-chromosome files, labels, id map and sizes."""
    rng = np.random.default_rng(51)
    tmp = tempfile.mkdtemp()
    recs_a = [("chr1 assembled molecule", util.messy_seq(rng, 1900)), ("chr2", util.messy_seq(rng, 120)),
              ("scaffold_7 unplaced", util.random_seq(rng, 300)), ("chr3", util.messy_seq(rng, 60)), ("chrM", "")]
    recs_b = [("Chr01 len=800", util.messy_seq(rng, 800)), ("Chr02", util.messy_seq(rng, 1261)), ("ctg9", "ACGT" * 10)]
    ga, gb = os.path.join(tmp, "A.fa"), os.path.join(tmp, "B.fa.gz")
    with open(ga, "wb") as f:
        f.write(util.fasta(recs_a))
    import gzip
    with gzip.open(gb, "wb") as f:
        f.write(util.fasta(recs_b, width=70, crlf=True))
    cases = []
    for prefixes, targets in ((["", ""], ["chr1", "chr3", "Chr02", "B1|Chr01", "missing9"]),
                              (["A_", "B_"], ["A_chr1", "A2|A_chr2", "B_Chr01", "Chr02"])):
        outdir = os.path.join(tmp, "chroms%d/" % len(cases))
        os.makedirs(outdir)
        # the reference opens gz files with plain open(): give it an inflated copy of the second genome, as its caller
        # would have to; the drop-in reads the .gz itself
        gb_plain = os.path.join(tmp, "B.plain.fa")
        with open(gb_plain, "wb") as f:
            f.write(gzip.open(gb, "rb").read())
        outfas, labels, d_t2, d_size = Seqs.split_genomes([ga, gb_plain], prefixes, targets, outdir)
        cases.append(dict(prefixes=prefixes, targets=targets, files=[os.path.basename(p) for p in outfas], labels=labels,
                          d_targets2=dict(d_t2), d_size={k_: int(v) for k_, v in d_size.items()},
                          contents={os.path.basename(p): open(p).read() for p in outfas}))
    genomes = dict(A=open(ga, "rb").read().decode(), B_plain=gzip.open(gb, "rb").read().decode("latin1"))
    shutil.rmtree(tmp)
    jdump(dict(genomes=genomes, cases=cases), "split_genomes.json")


def gen_map_stack(Seqs, Circos):
    rng = np.random.default_rng(3)
    cases = []
    for k, L, bin_size, window, chunk in ((5, 3000, 100, 700, True), (7, 5000, 250, 1000, True),
                                          (7, 5000, 256, 999, True), (15, 20000, 10000, 10e6, True),
                                          (6, 4000, 300, 10e6, False), (9, 2500, 10000000, 10e6, False)):
        seq = util.messy_seq(rng, L, n_frac=0.02)
        up = seq.upper()
        # choose specific k-mers from the sequence itself + revcomps (as Cluster.output_kmers builds them)
        d_kmers = {}
        sg_names = ["SG1", "SG2", "SG3"]
        starts = rng.integers(0, L - k, 120)
        for s in starts:
            km = up[s:s + k]
            if any(c not in "ACGT" for c in km):
                continue
            rc = kmers.revcomp(km)
            canon = min(km, rc)
            if canon in d_kmers:
                continue
            sg = sg_names[int(rng.integers(0, 3))]
            d_kmers[canon] = sg
            d_kmers[kmers.revcomp(canon)] = sg
        tmp = tempfile.mkdtemp()
        fa = os.path.join(tmp, "c.fasta")
        with open(fa, "wb") as f:
            f.write(util.fasta([("chrX desc", seq)]))
        out = io.StringIO()
        Seqs.map_kmer3([fa], d_kmers, fout=out, k=k, window_size=window, bin_size=bin_size, sg_names=sg_names,
                       ncpu=1, method="map", chunk=chunk)
        text = out.getvalue()
        bc = os.path.join(tmp, "bin.count")
        with open(bc, "w") as f:
            f.write(text)
        stacks = {}
        for ws in (bin_size, bin_size * 3, 1000, 1000000):
            coords, counts = Circos.stack_matrix(bc, window_size=ws)
            stacks[str(ws)] = dict(coords=[list(c) for c in coords], counts=[[int(x) for x in c] for c in counts])
        # circos density tracks (Circos.py:777-806): files per subgenome, 99th-percentile clipping
        dens = {}
        for ws in (bin_size, bin_size * 3):
            files = Circos.stack_bed_density(bc, os.path.join(tmp, "dens%d" % ws), sg_names, window_size=ws)
            dens[str(ws)] = {key: open(path).read() for key, path in files.items()}
        shutil.rmtree(tmp)
        cases.append(dict(k=k, seq=seq, bin_size=bin_size, window_size=window, chunk=chunk, sg_names=sg_names,
                          d_kmers=d_kmers, bin_count_text=text, stacks=stacks, density=dens))
    jdump(cases, "map_stack.json")


def gen_map_multi(Seqs):
    """Multi-record FASTA through Seqs.map_kmer3 (the LTR / custom-feature calls of __main__.py:509-518,574-589 use
    chunk=False with one row per sequence; chunk=True treats every record like a chromosome)."""
    rng = np.random.default_rng(17)
    cases = []
    for k, lens, bin_size, window, chunk in ((7, [900, 40, 1500, 5, 2300, 700], 10000000, 10e6, False),
                                             (9, [3000, 1200, 6, 4100], 500, 1300, True),
                                             (5, [60, 61, 59, 120, 3, 0, 300], 50, 10e6, False)):
        seqs = [util.messy_seq(rng, L, n_frac=0.03) if L else "" for L in lens]
        allup = "N".join(s.upper() for s in seqs)
        d_kmers = {}
        sg_names = ["SG1", "SG2"]
        for s in rng.integers(0, max(len(allup) - k, 1), 150):
            km = allup[s:s + k]
            if len(km) < k or any(c not in "ACGT" for c in km):
                continue
            canon = min(km, kmers.revcomp(km))
            if canon in d_kmers:
                continue
            sg = sg_names[int(rng.integers(0, 2))]
            d_kmers[canon] = sg
            d_kmers[kmers.revcomp(canon)] = sg
        tmp = tempfile.mkdtemp()
        fa = os.path.join(tmp, "m.fasta")
        records = [("rec%d some description" % i, s) for i, s in enumerate(seqs)]
        with open(fa, "wb") as f:
            f.write(util.fasta(records))
        out = io.StringIO()
        Seqs.map_kmer3([fa], d_kmers, fout=out, k=k, window_size=window, bin_size=bin_size, sg_names=sg_names,
                       ncpu=1, method="map", chunk=chunk)
        shutil.rmtree(tmp)
        cases.append(dict(k=k, records=[[n, s] for n, s in records], bin_size=bin_size, window_size=window,
                          chunk=chunk, sg_names=sg_names, d_kmers=d_kmers, bin_count_text=out.getvalue()))
    jdump(cases, "map_multi.json")


def gen_cluster_units(C):
    rng = np.random.default_rng(4)
    from collections import OrderedDict
    cases = []
    for n, groups in ((6, {"SG1": [0, 2, 4], "SG2": [1, 3, 5]}), (9, {"SG1": [0, 3, 6], "SG2": [1, 4, 7], "SG3": [2, 5, 8]}),
                      (21, {"SG1": list(range(0, 21, 3)), "SG2": list(range(1, 21, 3)), "SG3": list(range(2, 21, 3))}),
                      (20, {"SG1": list(range(0, 10)), "SG2": list(range(10, 20))}),     # groups >= 8: pairwise sums
                      (5, {"SG1": [0], "SG2": [1, 2], "SG3": [3, 4]})):                  # singleton group -> NaN
        rows = []
        for r in range(120):
            arr = rng.random(n) * 1e-5
            own = list(groups.values())[int(rng.integers(0, len(groups)))]
            arr[own] += rng.random(len(own)) * 5e-5
            if r % 17 == 0:
                arr[:] = 3e-5                     # zero variance, equal means -> NaN
            if r % 19 == 0:
                arr[:] = 1e-6
                arr[own] = 4e-5                   # zero variance, unequal means -> p = 0
            arr = [float(x) for x in arr]
            from scipy import stats
            kmer, max_sg, pvalue, rc_kmer, mean_vals = C._output_kmers(("ACGT", arr, OrderedDict(groups), stats.ttest_ind))
            rows.append(dict(array=arr, max_sg=max_sg, pvalue=float(pvalue), mean_vals=[float(m) for m in mean_vals]))
        cases.append(dict(n=n, groups=groups, rows=rows))
    jdump(cases, "ttest_rows.json")


def gen_ranktest_units(C):
    """Cluster._output_kmers (Cluster.py:178-194) with the other `-test_method` choices: scipy.stats.kruskal and
    mannwhitneyu through the reference's own function with the installed scipy; wilcoxon through the restatement of the
    pinned scipy 1.7.1 mode rules (oracle/restate.py wilcoxon_171 — the installed 1.18 treats ties / zeros differently)."""
    from collections import OrderedDict

    from scipy import stats

    from oracle import restate
    rng = np.random.default_rng(44)
    cases = []
    layouts = ((6, {"SG1": [0, 2, 4], "SG2": [1, 3, 5]}),
               (21, {"SG1": list(range(0, 21, 3)), "SG2": list(range(1, 21, 3)), "SG3": list(range(2, 21, 3))}),
               (20, {"SG1": list(range(0, 10)), "SG2": list(range(10, 20))}),       # both > 8: normal approximation
               (14, {"SG1": list(range(0, 5)), "SG2": list(range(5, 14))}),         # unequal sizes (not for wilcoxon)
               (60, {"SG1": list(range(0, 30)), "SG2": list(range(30, 60))}))       # n > 25: wilcoxon approximation
    for method in ("kruskal", "mannwhitneyu", "wilcoxon"):
        fn = dict(kruskal=stats.kruskal, mannwhitneyu=stats.mannwhitneyu, wilcoxon=restate.wilcoxon_171)[method]
        for n, groups in layouts:
            sizes = {len(v) for v in groups.values()}
            if method == "wilcoxon" and len(sizes) > 1:
                continue
            rows = []
            for r in range(100):
                arr = rng.random(n) * 1e-5
                own = list(groups.values())[int(rng.integers(0, len(groups)))]
                arr[own] += rng.random(len(own)) * (5e-5 if r % 3 else 5e-6)
                if r % 7 == 0:                                 # ties: counts / length of equal counts
                    arr = np.round(arr * 2e5) / 2e5
                if r % 11 == 0:                                # zero differences / many ties
                    arr[:] = 1e-6
                    arr[own[:2]] = 3e-6
                arr = [float(x) for x in arr]
                kmer, max_sg, pvalue, rc_kmer, mean_vals = C._output_kmers(("ACGT", arr, OrderedDict(groups), fn))
                rows.append(dict(array=arr, max_sg=max_sg, pvalue=float(pvalue), mean_vals=[float(m) for m in mean_vals]))
            cases.append(dict(method=method, n=n, groups=groups, rows=rows))
    jdump(cases, "ranktest_rows.json")


def arab_genome(seed):
    """13 chromosomes laid out like example_data/Arabidopsis_suecica_sg.config: ids 1..13, subgenome A = 1-5,
    subgenome B = 6-13, homoeologous sets whose groups hold one to three chromosomes."""
    rng = np.random.default_rng(seed)
    fam = {sg: [util.random_seq(rng, 400) for _ in range(6)] for sg in "AB"}
    shared = [util.random_seq(rng, 400) for _ in range(3)]
    records = []
    for i in range(13):
        sg = "A" if i < 5 else "B"
        L = int(26000 * (0.8 + 0.5 * rng.random()))
        seq = np.array(list(util.random_seq(rng, L)))
        covered = 0
        while covered < 0.7 * L:
            f = fam[sg][int(rng.integers(0, 6))] if rng.random() < 0.9 else shared[int(rng.integers(0, 3))]
            copy_ = np.array(list(f))
            mut = rng.random(len(copy_)) < 0.03
            copy_[mut] = np.array(list("ACGT"))[rng.integers(0, 4, int(mut.sum()))]
            pos = int(rng.integers(0, L - len(copy_)))
            seq[pos:pos + len(copy_)] = copy_
            covered += len(copy_)
        p = int(rng.integers(0, L - 200))
        seq[p:p + int(rng.integers(10, 100))] = "N"
        records.append((str(i + 1), "".join(seq)))
    sgs = [[["1"], ["6", "7"]], [["2", "3"], ["9", "8", "10"]], [["4", "5"], ["13", "11", "12"]]]
    return records, sgs


def gen_pipeline(J, C, Seqs, Circos, S_mod, fixture="pipeline_small", genome=None):
    """The reference's step sequence (__main__.py:403-498) on a tiny genome: 2x3 chromosomes (pipeline_small) or the
    Arabidopsis-like layout with multi-chromosome groups and uneven subgenomes (pipeline_arab)."""
    import sklearn.cluster
    out = os.path.join(HERE, fixture)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    records, sgs = genome if genome is not None else util.subgenome_genome(7, n_sg=2, chr_per_sg=3, chr_len=30000,
                                                                           n_fam=6, fam_len=400)
    labels = [name for name, _ in records]
    chromfiles = []
    for name, seq in records:
        p = os.path.join(out, name + ".fasta")
        with open(p, "wb") as f:
            f.write(util.fasta([(name, seq)]))
        chromfiles.append(p)
    k, lower_count, min_freq, nsg, replicates = 15, 3, 10, 2, 40
    tmp = tempfile.mkdtemp()

    def fake_jellyfish(seqfile, threads=4, k=17, prefix=None, lower_count=2, method="jellyfish", overwrite=False):
        output = os.path.join(tmp, os.path.basename(seqfile) + "_%d.fa" % k)
        keys, counts, _ = kmers.count_fasta(open(seqfile, "rb").read(), k, lower_count)
        with open(output, "w") as f:
            for a, b in zip(keys, counts):
                f.write("%s %d\n" % (kmers.key_to_str(a, k), b))
        return output

    dumpfiles = [fake_jellyfish(f, k=k, lower_count=lower_count) for f in chromfiles]
    J.plot_histogram = lambda *a, **kw: None
    dumps = J.JellyfishDumps(dumpfiles, labels, ncpu=2)
    d_mat = dumps.to_matrix()
    lengths = dumps.lengths
    n_union = len(d_mat)
    d_mat2 = dumps.filter(d_mat, lengths, sgs, outfig="x.pdf", min_fold=2, baseline=1, min_freq=min_freq,
                          max_freq=10000, min_prop=None, max_prop=None, ratio=1)
    matfile = os.path.join(out, "ref.kmer.mat")
    with open(matfile, "w") as f:
        dumps.write_matrix(d_mat2, f)
    # Cluster with deterministic KMeans / recorded resampling
    C.KMeans = functools.partial(sklearn.cluster.KMeans, n_init=10, random_state=0)
    M = len(d_mat2)
    rs = np.random.RandomState(12345)
    recorded = []

    def resample(data, replace=True, n_samples=None):
        idx = rs.randint(0, data.shape[0], size=(n_samples,))
        recorded.append(idx)
        return data[idx]

    C.resample = resample
    C.Cluster.pca = lambda *a, **kw: None
    cluster = C.Cluster(matfile, n_clusters=nsg, sg_prefix="SG", sg_assigned={}, replicates=replicates, jackknife=80)
    with open(os.path.join(out, "ref.chrom-subgenome.tsv"), "w") as f:
        cluster.output_subgenomes(f)
    with open(os.path.join(out, "ref.sig.kmer-subgenome.tsv"), "w") as f:
        d_kmers = cluster.output_kmers(f, max_pval=0.05, ncpu=2, test_method="ttest_ind")
    sg_map = os.path.join(out, "ref.subgenome.bin.count")
    with open(sg_map, "w") as f:
        Seqs.map_kmer3(chromfiles, d_kmers, fout=f, k=k, bin_size=1000, sg_names=cluster.sg_names, ncpu=2,
                       method="map", window_size=7000)
    bins, counts = Circos.stack_matrix(sg_map, window_size=5000)
    with open(os.path.join(out, "ref.bin.enrich"), "w") as f1, open(os.path.join(out, "ref.bin.group"), "w") as f2:
        S_mod.enrich_bin(f1, f2, cluster.d_sg, counts, colnames=cluster.sg_names, rownames=bins, max_pval=0.05, ncpu=2)
    # PCA numbers made deterministic (the reference's randomized solver is not): full SVD
    from sklearn.decomposition import PCA
    pca = PCA(n_components=nsg, svd_solver="full")
    X = pca.fit_transform(cluster.data)
    meta = dict(k=k, lower_count=lower_count, min_freq=min_freq, nsg=nsg, replicates=replicates, labels=labels,
                sgs=sgs, lengths=[int(x) for x in lengths], n_union=n_union, n_diff=len(d_mat2),
                d_sg=dict(cluster.d_sg), d_bs=cluster.d_bs, labels_full=[int(x) for x in cluster.labels],
                n_sig=len(d_kmers), bin_size=1000, map_window=7000, enrich_window=5000,
                mean_ari=float(cluster.mean_adjusted_rand_score), mean_vm=float(cluster.mean_v_measure_score),
                pca_scores=[[float(v) for v in r] for r in X], pca_ratio=[float(v) for v in pca.explained_variance_ratio_],
                inertia=float(cluster.kmean.inertia_),
                centers_head=[[float(v) for v in r[:50]] for r in cluster.kmean.cluster_centers_],
                centers_labels=[int(x) for x in cluster.kmean.labels_])
    with open(os.path.join(out, "meta.json"), "w") as f:
        json.dump(meta, f)
    np.savez_compressed(os.path.join(out, "resample_idx.npz"), idx=np.array(recorded, dtype=np.int32))
    shutil.rmtree(tmp)
    print("wrote", fixture, ": union", n_union, "diff", len(d_mat2), "sig", len(d_kmers) // 2, "d_sg", dict(cluster.d_sg))


def main(only=None):
    if not ref_shims.available():
        sys.exit("reference not present: fixtures can only be regenerated in the build container")
    J = ref_shims.load("Jellyfish")
    C = ref_shims.load("Cluster")
    S_mod = ref_shims.load("Stats")
    Seqs = ref_shims.load("Seqs")
    Circos = ref_shims.load("Circos")
    import logging
    logging.getLogger().setLevel(logging.WARNING)
    gen_filter(J)
    gen_fisher_enrich(S_mod)
    gen_enrich_ltr(S_mod)
    gen_stat_enrich()
    gen_split_genomes(Seqs)
    gen_map_stack(Seqs, Circos)
    gen_map_multi(Seqs)
    gen_cluster_units(C)
    gen_ranktest_units(C)
    gen_pipeline(J, C, Seqs, Circos, S_mod)
    gen_pipeline(J, C, Seqs, Circos, S_mod, fixture="pipeline_arab", genome=arab_genome(13))


if __name__ == "__main__":
    main()
