"""ORACLE — TEST / BENCH INFRASTRUCTURE ONLY (never imported by subphaser_b200/).

The reference's hot path timed on the host CPU, stage by stage, for bench.py's `cpu_baseline` leg and
`--impl reference` arm.  What runs:

  count     oracle/kmer_count.c (C port of `jellyfish count --canonical | dump -L`, Jellyfish.py:697-700), all threads
  matrix    restate.to_matrix                    (JellyfishDumps.to_matrix, Jellyfish.py:439-460)
  filter    restate.filter_matrix                (_filter_kmer per k-mer, Jellyfish.py:611-648)          1 core *
  cluster   sklearn KMeans(n_init=10) + bootstrap replicates + PCA   (Cluster.py:48-51,82-118)
  ttest     restate.output_kmer                  (scipy ttest_ind per k-mer, Cluster.py:178-194)         1 core *
  map       restate.map_kmer_lines               (per-base dict lookup, Seqs.py:209-237)                 1 core *
  stack     restate.stack_matrix                 (Circos.py:709-742)
  enrich    restate.enrich_row + restate.bh      (Stats.py:14-31,140-192; fisher -> scipy hypergeom.sf)  1 core *

`restate` is the numpy/Python restatement of the reference's functions, pinned to fixtures produced by the
reference's own code (tests/test_oracle_restate.py); /root/reference itself does not exist on the GPU box.
Stages marked * are per-element Python loops that the reference fans out over `multiprocessing.Pool(ncpu)`:
they are timed on one core over a bounded subsample and credited with PERFECT scaling over all host cores
(seconds = elements / rate / cores) — generous to the CPU.  Every stage reports what was measured and what was
extrapolated.
"""
import time
from collections import OrderedDict

import numpy as np

from . import kmers, restate

_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def _keys_to_strs(keys, k):
    shifts = np.arange(2 * (k - 1), -2, -2, dtype=np.uint64)
    codes = ((keys[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    return [bytes(r).decode() for r in letters]


def _revcomp_strs(strs):
    tr = str.maketrans("ACGT", "TGCA")
    return [s.translate(tr)[::-1] for s in strs]


def run(fastas, labels, sgs, k, window_size, threads, lower_count=3, min_freq=200, max_freq=10000, min_fold=2,
        baseline=1, ratio=1, nsg=None, replicates=1000, max_pval=0.05, bin_size=10000, chunk_size=10_000_000,
        max_filter_rows=150_000, max_ttest_rows=2500, max_boot=30, map_bases=1_200_000, max_enrich_windows=400,
        keep_dumps=False):
    """fastas: list of uint8 arrays (one chromosome each).  -> dict(stages=OrderedDict, totals...)"""
    from scipy import stats  # noqa: F401  (imported here so that its import time is not inside a timed stage)
    from sklearn.cluster import KMeans
    from sklearn.decomposition import PCA
    S = nsg or max(len(sg) for sg in sgs)
    n = len(fastas)
    st = OrderedDict()

    # ---- count ----
    t0 = time.perf_counter()
    dumps, n_kmers, n_bases = [], 0, 0
    for fa in fastas:
        kk, cc, s = kmers.count_fasta(fa, k, lower_count, nthreads=threads)
        dumps.append((kk, cc))
        n_kmers += s["n_valid_kmers"]
        n_bases += s["n_bases"]
    t_count = time.perf_counter() - t0
    st["count"] = dict(seconds=t_count, measured="all %d chromosomes, %d threads" % (n, threads), units=n_kmers,
                       unit="k-mers")

    # ---- matrix + filter ----
    t0 = time.perf_counter()
    allk, mat, lengths = restate.to_matrix(dumps)
    t_matrix = time.perf_counter() - t0
    U = len(allk)
    st["matrix"] = dict(seconds=t_matrix, measured="full union (numpy)", units=U, unit="union rows")
    sel = np.arange(U) if U <= max_filter_rows else np.linspace(0, U - 1, max_filter_rows).astype(np.int64)
    t0 = time.perf_counter()
    fkeys, fnorm, ftot, n_fold = restate.filter_matrix(allk[sel], mat[sel], lengths, labels, sgs, min_freq=min_freq,
                                                       max_freq=max_freq, min_fold=min_fold, baseline=baseline,
                                                       ratio=ratio)
    t_f = time.perf_counter() - t0
    rate_f = len(sel) / t_f
    st["filter"] = dict(seconds=U / rate_f / threads, measured="%d of %d rows on 1 core: %.2f s" % (len(sel), U, t_f),
                        units=U, unit="union rows", rate_per_core=rate_f, extrapolated="rows / rate / %d cores" % threads)
    M_sel = len(fkeys)
    M = int(round(M_sel * U / max(len(sel), 1)))
    if M_sel < 2 * S:
        raise RuntimeError("CPU sample too small: %d differential k-mers" % M_sel)

    # ---- cluster: z-score, KMeans(n_init=10), bootstrap, PCA ----
    t0 = time.perf_counter()
    Z = restate.zscore(fnorm)
    lab, km = restate.kmeans_labels(Z, S, labels, seed=0)
    t_full = (time.perf_counter() - t0) * (M / M_sel)
    R = int(replicates)
    nb = min(R, max_boot)
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    for r in range(nb):
        idx = rng.integers(0, M_sel, R)
        KMeans(n_clusters=S, n_init=10, random_state=r).fit(np.ascontiguousarray(Z[idx].T))
    t_boot = (time.perf_counter() - t0) * (R / max(nb, 1)) if nb else 0.0
    t0 = time.perf_counter()
    PCA(n_components=min(S, n)).fit_transform(np.ascontiguousarray(Z.T))
    t_pca = (time.perf_counter() - t0) * (M / M_sel)
    st["cluster"] = dict(seconds=t_full + t_boot + t_pca, seconds_fixed=t_boot,
                         measured="z-score + KMeans(n_init=10) + PCA on %d rows, %d of %d bootstrap replicates" % (M_sel, nb, R),
                         units=R + 1, unit="K-Means fits",
                         extrapolated="full fit and PCA linear in rows; the %d bootstrap fits (21 x %d each) do not grow with the genome" % (R, R))

    # ---- t-test per k-mer ----
    groups = OrderedDict()
    for i, g in enumerate(lab):
        groups.setdefault("SG%d" % (g + 1), []).append(i)
    rows = np.arange(M_sel) if M_sel <= max_ttest_rows else np.linspace(0, M_sel - 1, max_ttest_rows).astype(np.int64)
    t0 = time.perf_counter()
    for r in rows:
        restate.output_kmer(fnorm[r].tolist(), groups)
    t_t = time.perf_counter() - t0
    rate_t = len(rows) / t_t
    st["ttest"] = dict(seconds=M / rate_t / threads, measured="%d of %d rows on 1 core: %.2f s" % (len(rows), M, t_t),
                       units=M, unit="differential k-mers", rate_per_core=rate_t,
                       extrapolated="rows / rate / %d cores" % threads)

    # ---- specific k-mers (all differential k-mers of the sample, subgenome = group with the largest mean) ----
    gmeans = np.stack([fnorm[:, idx].mean(axis=1) for idx in groups.values()], axis=1)
    best = gmeans.argmax(axis=1).astype(np.uint8)
    sg_names = list(groups.keys())
    strs = _keys_to_strs(fkeys, k)
    d_kmers = {}
    for s_, rc, b in zip(strs, _revcomp_strs(strs), best.tolist()):
        d_kmers[s_] = sg_names[b]
        d_kmers[rc] = sg_names[b]

    # ---- map: the reference's per-base Python loop on a bounded piece of chromosome 0 ----
    codes0, _ = kmers.fasta_to_codes(fastas[0])
    piece = codes0[:map_bases]
    seq_piece = "".join("ACGTN"[min(c, 4)] for c in piece.tolist())
    t0 = time.perf_counter()
    restate.map_kmer_lines(labels[0], seq_piece, d_kmers, k, bin_size, sg_names, chunk=True, window_size=chunk_size)
    t_m = time.perf_counter() - t0
    rate_m = len(piece) / t_m
    st["map"] = dict(seconds=n_bases / rate_m / threads,
                     measured="first %d bases of %s on 1 core: %.2f s" % (len(piece), labels[0], t_m), units=n_bases,
                     unit="positions", rate_per_core=rate_m, extrapolated="positions / rate / %d cores" % threads)

    # ---- bin counts of the whole sample (C helper, NOT timed: it only provides realistic window counts) ----
    order = np.argsort(fkeys)
    skeys, ssg = fkeys[order], best[order]
    lines = []
    for lab_i, fa in zip(labels, fastas):
        codes, _ = kmers.fasta_to_codes(fa)
        L = len(codes)
        n_lines = L // bin_size + 1 + L // chunk_size + 2
        cnt, _ = kmers.map_bins(codes, k, skeys, ssg, len(sg_names), bin_size, chunk_size, n_lines)
        nz = np.flatnonzero(cnt.any(axis=1))
        for li in nz.tolist():
            # (line id -> bin start: lines are ordered by position; duplicates at chunk borders keep their bin's start)
            per = max(chunk_size // bin_size, 1) + 1
            b = max(li - li // per, 0)
            lines.append("%s\t%d\t%d\t%s\n" % (lab_i, b * bin_size, min((b + 1) * bin_size, L),
                                                 "\t".join(map(str, cnt[li].tolist()))))
    # ---- stack + enrich ----
    t0 = time.perf_counter()
    coords, counts = restate.stack_matrix(lines, window_size)
    t_stack = time.perf_counter() - t0
    W = len(counts)
    st["stack"] = dict(seconds=t_stack, measured="%d lines -> %d windows" % (len(lines), W), units=W, unit="windows")
    total = [int(x) for x in np.array(counts).sum(axis=0)] if W else [0] * len(sg_names)
    wsel = list(range(W)) if W <= max_enrich_windows else np.linspace(0, W - 1, max_enrich_windows).astype(int).tolist()
    t0 = time.perf_counter()
    pmin = []
    for w in wsel:
        pmin.append(restate.enrich_row([int(x) for x in counts[w]], total, max_pval=max_pval)["pval"])
    if pmin:
        restate.bh(pmin)
    t_e = time.perf_counter() - t0
    rate_e = len(wsel) / t_e if wsel and t_e > 0 else float("inf")
    st["enrich"] = dict(seconds=(W / rate_e / threads) if W else 0.0,
                        measured="%d of %d windows on 1 core: %.2f s" % (len(wsel), W, t_e), units=W, unit="windows",
                        rate_per_core=rate_e, extrapolated="windows / rate / %d cores" % threads)

    total_s = sum(v["seconds"] for v in st.values())
    fixed_s = sum(v.get("seconds_fixed", 0.0) for v in st.values())
    win_s = st["map"]["seconds"] + st["stack"]["seconds"] + st["enrich"]["seconds"]
    out = dict(stages=st, seconds=total_s, n_kmers=n_kmers, n_bases=n_bases, n_union=U, n_diff=M, n_windows=W,
               kmers_per_s=n_kmers / total_s, windows_per_s=(W / win_s) if win_s > 0 else None, cores=threads,
               labels=lab, seconds_fixed=fixed_s, window_seconds=win_s)
    if keep_dumps:
        out["dumps"] = dumps
    return out


def extrapolate(res, full_bases, full_windows=None):
    """Whole-path CPU rate at the full workload from a scaled-down replica: every per-element stage grows linearly
    with the genome, the bootstrap (a fixed number of fixed-size fits) does not.
    -> dict(seconds, kmers_per_s, windows_per_s, factor)"""
    f = float(full_bases) / float(res["n_bases"])
    secs = (res["seconds"] - res["seconds_fixed"]) * f + res["seconds_fixed"]
    n_kmers = res["n_kmers"] * f
    W = full_windows if full_windows is not None else res["n_windows"] * f
    win_secs = res["window_seconds"] * f
    return dict(seconds=secs, kmers_per_s=n_kmers / secs, windows_per_s=(W / win_secs) if win_secs > 0 else None,
                factor=f)
