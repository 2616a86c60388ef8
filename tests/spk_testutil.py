"""Small synthetic FASTA builders shared by the tests (numpy, deterministic)."""
import numpy as np


def random_seq(rng, n, alphabet="ACGT"):
    return "".join(np.array(list(alphabet))[rng.integers(0, len(alphabet), n)])


def wrap(seq, width=60):
    return "\n".join(seq[i:i + width] for i in range(0, len(seq), width))


def fasta(records, width=60, crlf=False):
    """records: list of (name, seq) -> bytes"""
    out = []
    for name, seq in records:
        out.append(">" + name)
        if seq:
            out.append(wrap(seq, width))
    txt = "\n".join(out) + "\n"
    if crlf:
        txt = txt.replace("\n", "\r\n")
    return txt.encode()


def messy_seq(rng, n, n_frac=0.01, lower_frac=0.3, repeat_unit=None):
    """ACGT with N runs, IUPAC codes, lower-case stretches and an optional tandem repeat block."""
    s = list(random_seq(rng, n))
    i = 0
    while i < n:
        if rng.random() < n_frac:
            run = int(rng.integers(1, 40))
            for j in range(i, min(n, i + run)):
                s[j] = "N" if rng.random() < 0.9 else "RYKM"[int(rng.integers(0, 4))]
            i += run
        i += int(rng.integers(1, 200))
    if repeat_unit:
        start = n // 3
        rep = (repeat_unit * (n // (4 * len(repeat_unit)) + 1))[: n // 4]
        s[start:start + len(rep)] = list(rep)
    s = "".join(s)
    out = []
    i = 0
    while i < n:
        run = int(rng.integers(50, 500))
        seg = s[i:i + run]
        out.append(seg.lower() if rng.random() < lower_frac else seg)
        i += run
    return "".join(out)


def subgenome_genome(seed, n_sg=2, chr_per_sg=3, chr_len=60000, n_fam=6, fam_len=400, te_frac=0.6, div=0.03):
    """Tiny polyploid: each subgenome has private repeat families, plus shared ones.
    -> (list of (name, seq), sgs config [[['1A'],['1B']], ...])"""
    rng = np.random.default_rng(seed)
    sg_names = "ABCDEFGH"[:n_sg]
    private = {sg: [random_seq(rng, fam_len) for _ in range(n_fam)] for sg in sg_names}
    shared = [random_seq(rng, fam_len) for _ in range(max(1, n_fam // 2))]
    records = []
    for c in range(chr_per_sg):
        for sg in sg_names:
            L = int(chr_len * (0.8 + 0.4 * rng.random()))
            seq = np.array(list(random_seq(rng, L)))
            covered = 0
            while covered < te_frac * L:
                fam = private[sg][int(rng.integers(0, n_fam))] if rng.random() < 0.85 else shared[
                    int(rng.integers(0, len(shared)))]
                copy = np.array(list(fam))
                mut = rng.random(len(copy)) < div
                copy[mut] = np.array(list("ACGT"))[rng.integers(0, 4, int(mut.sum()))]
                pos = int(rng.integers(0, L - len(copy)))
                seq[pos:pos + len(copy)] = copy
                covered += len(copy)
            # one N run and some soft-masking
            p = int(rng.integers(0, L - 200))
            seq[p:p + int(rng.integers(10, 150))] = "N"
            s = "".join(seq)
            q = int(rng.integers(0, L - 3000))
            s = s[:q] + s[q:q + 3000].lower() + s[q + 3000:]
            records.append(("%d%s" % (c + 1, sg), s))
    sgs = [[["%d%s" % (c + 1, sg)] for sg in sg_names] for c in range(chr_per_sg)]
    return records, sgs
