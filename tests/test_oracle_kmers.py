"""CPU: pin the C k-mer oracle (oracle/kmer_count.c) to hand-computed vectors and to an independent
string-level counter.  jellyfish 2.2.10 itself is not available: PARITY UNPINNED against it."""
import numpy as np
import pytest

import spk_testutil as util
from oracle import kmers


def _as_dict(keys, counts, k):
    return {kmers.key_to_str(a, k): int(b) for a, b in zip(keys, counts)}


def test_hand_vectors():
    # ACGTACGT, k=3: ACG CGT GTA TAC ACG CGT; ACG==rc(CGT); GTA==rc(TAC)
    keys, counts, st = kmers.count_fasta(b">s\nACGTACGT\n", 3, 1)
    assert _as_dict(keys, counts, 3) == {"ACG": 4, "GTA": 2}
    assert st["n_valid_kmers"] == 6 and st["n_distinct"] == 2 and st["sum_dumped"] == 6
    # palindromic k-mers are their own reverse complement (even k): GAATTC
    keys, counts, _ = kmers.count_fasta(b">s\nGAATTCGAATTC\n", 6, 1)
    d = _as_dict(keys, counts, 6)
    assert d["GAATTC"] == 2
    # N resets the window; lower-case counts; k > len gives nothing
    keys, counts, st = kmers.count_fasta(b">s\nacgNacgt\n", 3, 1)
    assert _as_dict(keys, counts, 3) == {"ACG": 3}      # acg | acg, cgt(=ACG)
    keys, counts, st = kmers.count_fasta(b">s\nACG\n", 5, 1)
    assert len(keys) == 0 and st["n_valid_kmers"] == 0
    # k-mers never span records; -L drops low counts; lengths = sum of dumped counts only
    keys, counts, st = kmers.count_fasta(b">a\nAAAA\n>b\nAAAA\n", 3, 1)
    assert _as_dict(keys, counts, 3) == {"AAA": 4}
    keys, counts, st = kmers.count_fasta(b">a\nAAAACCCC\n", 3, 2)
    assert _as_dict(keys, counts, 3) == {"AAA": 2, "CCC": 2} and st["sum_dumped"] == 4
    assert st["n_valid_kmers"] == 6 and st["n_distinct"] == 4


@pytest.mark.parametrize("k", [1, 2, 4, 7, 15, 17, 21, 31, 32])
def test_against_bruteforce(k):
    rng = np.random.default_rng(k)
    recs = [("r%d" % i, util.messy_seq(rng, 3000, repeat_unit="AT" if i else None)) for i in range(3)]
    fa = util.fasta(recs, width=int(rng.integers(5, 90)))
    want = kmers.brute_count([s for _, s in recs], k)
    for threads in (1, 4):
        keys, counts, st = kmers.count_fasta(fa, k, 1, nthreads=threads)
        assert _as_dict(keys, counts, k) == want
        assert st["n_valid_kmers"] == sum(want.values())
    keys, counts, st = kmers.count_fasta(fa, k, 3)
    assert _as_dict(keys, counts, k) == {a: b for a, b in want.items() if b >= 3}
    assert st["sum_dumped"] == sum(b for b in want.values() if b >= 3)


def test_line_wrap_and_crlf_independent():
    rng = np.random.default_rng(5)
    seq = util.messy_seq(rng, 5000)
    base = kmers.count_fasta(util.fasta([("c", seq)], width=60), 11, 1)
    for fa in (util.fasta([("c", seq)], width=7), util.fasta([("c", seq)], width=60, crlf=True),
               util.fasta([("c", seq)], width=10**6)):
        cur = kmers.count_fasta(fa, 11, 1)
        assert np.array_equal(cur[0], base[0]) and np.array_equal(cur[1], base[1])
