"""Synthetic polyploid genomes for tests and bench.py — TEST / BENCH INFRASTRUCTURE, not product.

Real genomes cannot be downloaded here.  A genome is described by a small plan drawn on the host
(numpy): per subgenome a private library of repeat families plus a shared library; each chromosome is
an iid background with repeat copies pasted in until `te_frac` of its length (85 % from its own
subgenome's library, 15 % shared; per-copy substitution rate `div`), one run of N per Mb and 30 %
soft-masked bases.  The FASTA bytes (header, 60-column lines) are then produced directly in device
memory by the spk_synth_fasta kernel, so a 14-Gb wheat-shaped genome never exists on the host.
"""
import ctypes

import numpy as np


# IWGSC RefSeq v2.1-like chromosome lengths (Mb), order 1A 1B 1D 2A ... 7D
WHEAT_MB = [598, 700, 498, 787, 812, 656, 754, 851, 619, 754, 673, 518, 713, 714, 569, 622, 731, 495, 744,
            764, 642]

CONFIGS = {
    # name: (subgenome letters, chromosomes per subgenome, lengths in bp (None = uniform), k, window)
    "C1": dict(sg="AB", lengths=None, k=15, window=1_000_000),      # Arabidopsis-shaped, filled below
    "C2": dict(sg="AB", lengths=[127_000_000] * 20, k=15, window=1_000_000),
    "C3": dict(sg="ABD", lengths=[m * 1_000_000 for m in WHEAT_MB], k=17, window=1_000_000),
    # C4 (bootstrap K-Means + PCA on the chrom x k-mer matrix) is the clustering stage of the C3 run
    "C4": dict(sg="ABD", lengths=[m * 1_000_000 for m in WHEAT_MB], k=17, window=1_000_000),
    # synthetic hexaploid, 21 x 2.38 Gb (chromosomes longer than 2^31 bp), k=21, 100-kb windows
    "C5": dict(sg="ABD", lengths=[2_380_000_000] * 21, k=21, window=100_000),
}
# C1: the layout of example_data/Arabidopsis_suecica_sg.config — 13 chromosomes renamed 1..13, subgenome A = 1-5,
# subgenome B = 6-13, three homoeologous sets whose groups hold one to three chromosomes each
CONFIGS["C1"]["lengths"] = [int(x) for x in np.linspace(14e6, 26e6, 13)]
CONFIGS["C1"]["names"] = [str(i) for i in range(1, 14)]
CONFIGS["C1"]["sg_of"] = [0] * 5 + [1] * 8
CONFIGS["C1"]["sgs"] = [[["1"], ["6", "7"]], [["2", "3"], ["9", "8", "10"]], [["4", "5"], ["13", "11", "12"]]]


class GenomePlan:
    def __init__(self, seed, sg_letters, lengths, n_fam=200, n_shared=100, fam_len=(2000, 10000),
                 te_frac=0.7, div=0.03, soft_frac=0.3, own_frac=0.85, n_per_mb=1.0, names=None, sg_of=None, sgs=None):
        self.seed = int(seed)
        rng = np.random.default_rng(seed)
        self.sg_letters = list(sg_letters)
        S = len(self.sg_letters)
        self.te_frac, self.div, self.soft_frac, self.own_frac = te_frac, div, soft_frac, own_frac
        self.n_per_mb = n_per_mb
        # libraries: concatenated 2-bit codes; family offsets/lengths per owner (0..S-1 private, S shared)
        fams = []
        for owner in range(S + 1):
            nf = n_fam if owner < S else n_shared
            lens = rng.integers(fam_len[0], fam_len[1] + 1, nf)
            fams.append(lens)
        total = int(sum(l.sum() for l in fams))
        self.library = rng.integers(0, 4, total, dtype=np.uint8)
        self.fam_off, self.fam_len = [], []
        off = 0
        for lens in fams:
            self.fam_off.append(off + np.concatenate([[0], np.cumsum(lens)[:-1]]))
            self.fam_len.append(lens)
            off += int(lens.sum())
        # chromosomes: set j = chromosome j of every subgenome (homoeolog config rows)
        n = len(lengths)
        self.chroms = []
        per = n // S
        for i, L in enumerate(lengths):
            if names is not None:      # explicit layout (multi-chromosome groups, uneven subgenomes)
                self.chroms.append(dict(name=names[i], sg=int(sg_of[i]), length=int(L), index=i))
                continue
            c, s = divmod(i, S) if n % S == 0 else (i // S, i % S)
            self.chroms.append(dict(name="%d%s" % (c + 1, self.sg_letters[s]), sg=s, length=int(L), index=i))
        self.labels = [c["name"] for c in self.chroms]
        if sgs is not None:
            self.sgs = [[list(g) for g in row] for row in sgs]
        else:
            sets = {}
            for c in self.chroms:
                sets.setdefault(c["name"][:-1], []).append([c["name"]])
            self.sgs = [v for v in sets.values()]

    def segments(self, chrom):
        """Segment table of one chromosome: (seg_start u64, seg_src i64 (-1 = background), seg_seed u32,
        nrun_start u64, nrun_end u64)."""
        rng = np.random.default_rng([self.seed, 7919, chrom["index"]])
        L, s = chrom["length"], chrom["sg"]
        S = len(self.sg_letters)
        mean_len = float(np.mean(np.concatenate(self.fam_len)))
        gap_mean = mean_len * (1 - self.te_frac) / self.te_frac
        n_est = int(L / (mean_len + gap_mean) * 1.3) + 16
        owner = np.where(rng.random(n_est) < self.own_frac, s, S)
        fam = np.empty(n_est, dtype=np.int64)
        for o in (s, S):
            m = owner == o
            fam[m] = rng.integers(0, len(self.fam_len[o]), int(m.sum()))
        te_len = np.where(owner == s, self.fam_len[s][np.minimum(fam, len(self.fam_len[s]) - 1)],
                          self.fam_len[S][np.minimum(fam, len(self.fam_len[S]) - 1)]).astype(np.int64)
        te_off = np.where(owner == s, self.fam_off[s][np.minimum(fam, len(self.fam_off[s]) - 1)],
                          self.fam_off[S][np.minimum(fam, len(self.fam_off[S]) - 1)]).astype(np.int64)
        gaps = np.maximum(rng.exponential(gap_mean, n_est).astype(np.int64), 1)
        # interleave gap, te, gap, te ...
        lens = np.empty(2 * n_est, dtype=np.int64)
        lens[0::2], lens[1::2] = gaps, te_len
        src = np.empty(2 * n_est, dtype=np.int64)
        src[0::2], src[1::2] = -1, te_off
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
        keep = starts < L
        starts, src = starts[keep], src[keep]
        seeds = rng.integers(1, 2**32 - 1, len(starts), dtype=np.uint32)
        n_runs = max(int(L / 1e6 * self.n_per_mb), 1 if L >= 2000 else 0)
        if n_runs:
            rs = np.sort(rng.integers(0, max(L - 1, 1), n_runs)).astype(np.int64)
            rl = rng.integers(100, 10001, n_runs).astype(np.int64)
            re = np.minimum(rs + rl, L)
            re[:-1] = np.minimum(re[:-1], rs[1:])          # no overlaps
        else:
            rs = re = np.zeros(0, np.int64)
        return (starts.astype(np.uint64), src, seeds, rs.astype(np.uint64), re.astype(np.uint64))

    def fasta_nbytes(self, chrom, line_width=60):
        header = (">%s\n" % chrom["name"]).encode()
        L = chrom["length"]
        return len(header), len(header) + L + (L + line_width - 1) // line_width


def synth_chromosome(plan, chrom, line_width=60, d_library=None):
    """-> (device uint8 tensor holding the FASTA bytes (16-B padded), nbytes)."""
    import torch
    from . import _lib, engine
    engine.require_cuda()
    dev = engine._dev()
    hlen, nbytes = plan.fasta_nbytes(chrom, line_width)
    starts, src, seeds, rs, re = plan.segments(chrom)
    if d_library is None:
        d_library = torch.from_numpy(plan.library).to(dev)
    out = torch.empty(nbytes + 16, dtype=torch.uint8, device=dev)
    header = np.frombuffer((">%s\n" % chrom["name"]).encode(), dtype=np.uint8)
    out[:hlen].copy_(torch.from_numpy(header.copy()))
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt).copy()).to(dev)  # noqa: E731
    d_start, d_src, d_seed = t(starts, np.int64), t(src, np.int64), t(seeds, np.int32)
    d_rs, d_re = t(rs, np.int64), t(re, np.int64)
    _lib.call("spk_synth_fasta", engine._p(out), nbytes, hlen, chrom["length"], line_width, engine._p(d_start),
              engine._p(d_src), engine._p(d_seed), len(starts), engine._p(d_library), int(plan.library.size),
              engine._p(d_rs) if len(rs) else ctypes.c_void_p(0), engine._p(d_re) if len(rs) else ctypes.c_void_p(0),
              len(rs), float(plan.div), float(plan.soft_frac), plan.seed, engine._stream())
    torch.cuda.current_stream().synchronize()
    return out, nbytes


def _hash64(x):
    x = x ^ (x >> np.uint64(33))
    x = x * np.uint64(0xff51afd7ed558ccd)
    x = x ^ (x >> np.uint64(33))
    x = x * np.uint64(0xc4ceb9fe1a85ec53)
    return x ^ (x >> np.uint64(33))


def _mix(a, b):
    return _hash64(a * np.uint64(0x9E3779B97F4A7C15) + b + np.uint64(0x632BE59BD9B4E019))


def synth_chromosome_host(plan, chrom, line_width=60, block=1 << 22):
    """numpy twin of spk_synth_fasta (csrc/spk_synth.cu): the same FASTA bytes as synth_chromosome(), built on the
    host without libspk — for the CPU reference arm of bench.py (which must not load the product library) and for
    CPU tests.  -> uint8 array of nbytes."""
    hlen, nbytes = plan.fasta_nbytes(chrom, line_width)
    starts, src, seeds, rs, re = plan.segments(chrom)
    L = chrom["length"]
    lib = plan.library
    out = np.empty(nbytes, dtype=np.uint8)
    out[:hlen] = np.frombuffer((">%s\n" % chrom["name"]).encode(), dtype=np.uint8)
    seed = np.uint64(plan.seed)
    div_thr = np.uint64(plan.div * 18446744073709551615.0)
    soft_thr = np.uint64(plan.soft_frac * 18446744073709551615.0)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    old = np.seterr(over="ignore")
    try:
        for a in range(hlen, nbytes, block):
            o = np.arange(a, min(a + block, nbytes), dtype=np.uint64)
            q = o - np.uint64(hlen)
            line = q // np.uint64(line_width + 1)
            col = q - line * np.uint64(line_width + 1)
            i = line * np.uint64(line_width) + col
            newline = (col == np.uint64(line_width)) | (i >= np.uint64(L))
            ii = np.minimum(i, np.uint64(max(L - 1, 0)))
            isn = np.zeros(len(o), dtype=bool)
            if len(rs):
                lo = np.searchsorted(rs, ii, side="right")
                ok = lo > 0
                isn[ok] = ii[ok] < re[lo[ok] - 1]
            sidx = np.searchsorted(starts, ii, side="right")
            seg = np.maximum(sidx, 1) - 1
            ssrc = src[seg]
            bg = (sidx == 0) | (ssrc < 0)
            b = (_mix(seed, ii) & np.uint64(3)).astype(np.uint32)
            te = ~bg
            if te.any():
                within = ii[te] - starts[seg[te]]
                pos = np.minimum(ssrc[te].astype(np.uint64) + within, np.uint64(len(lib) - 1))
                tb = (lib[pos.astype(np.int64)] & 3).astype(np.uint32)
                h = _mix(seed ^ (seeds[seg[te]].astype(np.uint64) << np.uint64(32)), within)
                mut = h < div_thr
                tb[mut] = (tb[mut] + 1 + ((h[mut] >> np.uint64(7)) % np.uint64(3)).astype(np.uint32)) & 3
                b[te] = tb
            ch = acgt[b]
            soft = _mix(seed ^ np.uint64(0xABCDEF), ii >> np.uint64(9)) < soft_thr
            ch = np.where(soft, ch | 0x20, ch).astype(np.uint8)
            ch[isn] = ord("N")
            ch[newline] = ord("\n")
            out[a:a + len(o)] = ch
    finally:
        np.seterr(**old)
    return out


def plan_for(config, seed=None, scale=1.0):
    """GenomePlan of one of the BASELINE.json configs; `scale` shrinks every chromosome (tests)."""
    cfg = CONFIGS[config]
    seeds = {"C1": 101, "C2": 202, "C3": 303, "C4": 303, "C5": 505}
    lengths = [max(int(L * scale), 1000) for L in cfg["lengths"]]
    # the repeat library is sized for the genome (200 + 100 families per 14.2 Gb) so that copy numbers per family stay
    # wheat-like whatever the configuration or scale: the default thresholds (min_freq 200) then select k-mers
    g = float(sum(lengths))
    n_fam = max(int(round(200 * g / 14.2e9)), 4)
    n_shared = max(int(round(100 * g / 14.2e9)), 2)
    return GenomePlan(seeds[config] if seed is None else seed, cfg["sg"], lengths, n_fam=n_fam,
                      n_shared=n_shared, names=cfg.get("names"), sg_of=cfg.get("sg_of"), sgs=cfg.get("sgs")), cfg
