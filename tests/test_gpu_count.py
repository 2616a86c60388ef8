"""GPU parity: K1 pack + K2 count + K3 dump vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

import spk_testutil as util

pytestmark = pytest.mark.gpu


MODE = ["partitioned"]


@pytest.fixture(params=["partitioned", "global"], autouse=True)
def count_mode(request):
    """every test runs against both counters: partitioned (shared-memory tables) and v1 (global table)"""
    MODE[0] = request.param
    yield


def _gpu_count(fasta_bytes, k, lower):
    from subphaser_b200 import engine
    d, n = engine.to_device_bytes(fasta_bytes)
    seq = engine.pack_fasta(d, n)
    table = engine.CountTable(max(seq.n_bases, 1), k, lower, mode=MODE[0])
    dump = engine.count_packed(seq, k, lower, table=table)
    keys, counts = dump.to_host()
    order = np.argsort(keys, kind="stable")
    return seq, dump, keys[order], counts[order]


def _check(fasta_bytes, k, lower):
    from oracle import kmers
    okeys, ocounts, st = kmers.count_fasta(fasta_bytes, k, lower)
    seq, dump, keys, counts = _gpu_count(fasta_bytes, k, lower)
    assert seq.n_bases == st["n_bases"]
    assert seq.n_records == st["n_records"]
    assert dump.n_valid_kmers == st["n_valid_kmers"]
    assert dump.n_distinct == st["n_distinct"]
    assert dump.length == st["sum_dumped"]
    assert len(keys) == st["n_dumped"]
    np.testing.assert_array_equal(keys, okeys)
    np.testing.assert_array_equal(counts, ocounts)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 11, 15, 16, 17, 21, 27, 31, 32])
def test_random_messy(k):
    rng = np.random.default_rng(100 + k)
    seq = util.messy_seq(rng, 50000, repeat_unit="AT")
    _check(util.fasta([("chr1", seq)]), k, 1)
    _check(util.fasta([("chr1", seq)]), k, 3)


def test_known_answer_small():
    from oracle import kmers
    fa = b">s\nACGTACGT\n"
    _, _, keys, counts = _gpu_count(fa, 3, 1)
    got = {kmers.key_to_str(a, 3): int(b) for a, b in zip(keys, counts)}
    # ACG,CGT,GTA,TAC,ACG,CGT ; canonical: ACG(=CGT rc) x4, GTA(rc TAC)->GTA x2
    assert got == {"ACG": 4, "GTA": 2}


def test_edge_cases():
    rng = np.random.default_rng(7)
    cases = [
        b"",                                   # empty file
        b">only_header\n",                     # no sequence
        b">s\nNNNNNNNNNNNNNNNNNNNNNNNN\n",      # all N
        b">s\nACG\n",                          # shorter than k
        b"ACGTACGTACGTAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n",  # no header
        util.fasta([("a", "ACGT" * 50), ("b", "ACGT" * 50), ("c", "")]),  # multi-record + empty record
        util.fasta([("a", util.random_seq(rng, 5000))], width=7),
        util.fasta([("a", util.random_seq(rng, 5000))], width=60, crlf=True),
        util.fasta([("p", "A" * 10000)]),       # homopolymer (run-merge + warp aggregation)
        util.fasta([("p", "AT" * 5000 + "ACGT" * 2500 + "AACCGGTT" * 1250)]),  # microsatellites
        util.fasta([("pal", "ACGT" * 3 + "N" + "GAATTC" * 100)]),  # even-k palindromes
    ]
    for fa in cases:
        for k in (4, 6, 15, 17):
            _check(fa, k, 1)
            _check(fa, k, 2)


def test_n_every_k_minus_1():
    k = 15
    unit = "ACGTTGCAAGGCTA" + "N"      # 14 valid bases then N: no k-mer at all
    _check(util.fasta([("x", unit * 500)]), k, 1)
    unit2 = "ACGTTGCAAGGCTAC" + "N"    # exactly one k-mer per unit
    _check(util.fasta([("x", unit2 * 500)]), k, 1)


def test_line_wrap_independence():
    rng = np.random.default_rng(3)
    seq = util.messy_seq(rng, 30000)
    ref = None
    for width in (1, 13, 60, 80, 100000):
        _, dump, keys, counts = _gpu_count(util.fasta([("c", seq)], width=width), 17, 1)
        cur = (keys.tobytes(), counts.tobytes())
        if ref is None:
            ref = cur
        assert cur == ref


def test_larger_than_tile_and_table_growth():
    rng = np.random.default_rng(11)
    seq = util.messy_seq(rng, 1_500_000, repeat_unit="ACGGT")
    _check(util.fasta([("big", seq)]), 17, 3)
    _check(util.fasta([("big", seq)]), 21, 1)


@pytest.mark.parametrize("k", [15, 17, 21])
def test_two_level_scatter_path(k):
    """>= 6.3 M bases switch the partitioned counter to its two-level shared-memory staged scatter."""
    if MODE[0] != "partitioned":
        pytest.skip("partitioned counter only")
    rng = np.random.default_rng(k)
    unit = util.random_seq(rng, 3000)
    parts = []
    for i in range(30):                      # repeats (counts > 1) between stretches of unique sequence
        parts.append(util.random_seq(rng, 300000))
        parts.append(unit if i % 3 else unit[:1500] + "N" * 20 + unit[1500:])
    seq = "".join(parts)
    assert len(seq) > 9_000_000
    _check(util.fasta([("big", seq)]), k, 1)
    _check(util.fasta([("big", seq)]), k, 3)


def test_retry_queue_overflow_paths(monkeypatch):
    """k_part_count32 queues the entries whose home slot is taken and probes them later; with the queue capped
    at a few entries (test hook) every partition overflows it, so the inline-probe and the straddling-
    reservation paths must give the same counts."""
    if MODE[0] != "partitioned":
        pytest.skip("partitioned counter only")
    rng = np.random.default_rng(5)
    seq = util.messy_seq(rng, 400_000, repeat_unit="ACGTTGCA")
    fa = util.fasta([("c", seq)])
    for cap in ("0", "7", "100"):
        monkeypatch.setenv("SPK_PCOUNT_RETRY_CAP", cap)
        _check(fa, 17, 1)
        _check(fa, 15, 2)
    monkeypatch.delenv("SPK_PCOUNT_RETRY_CAP")


def _gpu_count_forced(fasta_bytes, k, lower, genome_max_bases):
    """Partitioned counter with the partition bits of a genome whose largest chromosome has `genome_max_bases`
    bases (what hotpath does for every chromosome of a genome): small inputs then run the large-genome kernels."""
    from subphaser_b200 import engine
    d, n = engine.to_device_bytes(fasta_bytes)
    seq = engine.pack_fasta(d, n)
    table = engine.CountTable(max(seq.n_bases, 1), k, lower, mode="partitioned", genome_max_bases=genome_max_bases)
    dump = engine.count_packed(seq, k, lower, table=table)
    keys, counts = dump.to_host()
    order = np.argsort(keys, kind="stable")
    return seq, dump, table, keys[order], counts[order]


def _check_forced(fa, k, lower, genome_max_bases, want_pbits):
    from oracle import kmers
    okeys, ocounts, st = kmers.count_fasta(fa, k, lower)
    seq, dump, table, keys, counts = _gpu_count_forced(fa, k, lower, genome_max_bases)
    assert table.pbits == want_pbits
    assert dump.n_valid_kmers == st["n_valid_kmers"]
    assert dump.n_distinct == st["n_distinct"]
    assert dump.length == st["sum_dumped"]
    np.testing.assert_array_equal(keys, okeys)
    np.testing.assert_array_equal(counts, ocounts)
    # the dump is grouped by partition and pindex describes it exactly
    pidx = dump.pindex.cpu().numpy().astype(np.int64)
    assert int(pidx[1::2].sum()) == len(keys)
    return dump


@pytest.mark.parametrize("k,gmax,pbits", [(17, 100_000_000, 15), (17, 800_000_000, 18), (15, 300_000_000, 17),
                                          (20, 1_500_000_000, 19), (16, 600_000_000, 18)])
def test_v3_descriptor_pipeline(k, gmax, pbits, monkeypatch):
    """Partition bits 15..19 select the descriptor pipeline (k_v3_l1 / k_v3_chunks / k_v3_l2 / gathering counter):
    bit-exact against the oracle, and identical to the two-level scatter pipeline (SPK_PCOUNT_PIPE=v2)."""
    if MODE[0] != "partitioned":
        pytest.skip("partitioned counter only")
    rng = np.random.default_rng(k + pbits)
    unit = util.random_seq(rng, 3000)
    parts = []
    for i in range(8):
        parts.append(util.messy_seq(rng, 250_000, repeat_unit="ACGTTGCA" if i == 3 else None))
        parts.append(unit if i % 3 else unit[:1500] + "N" * 20 + unit[1500:])
    fa = util.fasta([("big", "".join(parts))])
    for lower in (1, 3):
        d3 = _check_forced(fa, k, lower, gmax, pbits)
        for variant in ("versioned", "sweep"):      # (default: the kept-slot-list kernel) no-clear table; full sweep
            monkeypatch.setenv("SPK_PCOUNT_TABLE", variant)
            _check_forced(fa, k, lower, gmax, pbits)
            monkeypatch.delenv("SPK_PCOUNT_TABLE")
        monkeypatch.setenv("SPK_PCOUNT_PIPE", "v2")
        d2 = _check_forced(fa, k, lower, gmax, pbits)
        monkeypatch.delenv("SPK_PCOUNT_PIPE")
        # same partition function: the per-partition dump sizes agree between the two pipelines
        np.testing.assert_array_equal(d3.pindex.cpu().numpy()[1::2], d2.pindex.cpu().numpy()[1::2])


def test_v3_skewed_buckets():
    """Low-complexity input: one bucket receives (almost) everything — runs of 16384 entries, chunks that start in
    the middle of a run, a bucket with more chunks than the counter's descriptor registers cover (> 256), empty
    buckets, and sub-partitions far above the table's first-probe capacity."""
    if MODE[0] != "partitioned":
        pytest.skip("partitioned counter only")
    rng = np.random.default_rng(77)
    seq = "A" * 4_600_000 + util.random_seq(rng, 40_000) + "AC" * 300_000 + "N" * 100 + "ACGGT" * 100_000
    fa = util.fasta([("skew", seq)])
    _check_forced(fa, 17, 1, 700_000_000, 18)
    _check_forced(fa, 17, 3, 700_000_000, 18)
    _check_forced(fa, 15, 2, 100_000_000, 15)


def test_v3_tiny_and_empty_inputs():
    if MODE[0] != "partitioned":
        pytest.skip("partitioned counter only")
    rng = np.random.default_rng(78)
    for seq in ("", "ACGTACGTAC", util.random_seq(rng, 5000), "N" * 20000, util.random_seq(rng, 16384 + 16),
                util.random_seq(rng, 4 * 16384 + 17)):
        fa = util.fasta([("t", seq)])
        _check_forced(fa, 17, 1, 700_000_000, 18)


def _unpack(seq):
    """PackedSeq -> uint8 codes (0..3, 4 = invalid) on the host"""
    n = seq.n_bases
    pk = seq.packed.cpu().numpy().view(np.uint32)
    vd = seq.valid.cpu().numpy().view(np.uint32)
    i = np.arange(n, dtype=np.int64)
    codes = ((pk[i >> 4] >> ((i & 15) * 2).astype(np.uint32)) & 3).astype(np.uint8)
    ok = ((vd[i >> 5] >> (i & 31).astype(np.uint32)) & 1).astype(bool)
    codes[~ok] = 4
    return codes


@pytest.mark.parametrize("mode", ["regular", "single", "3pass"])
def test_pack_matches_oracle_codes(mode, monkeypatch):
    """K1: the regular-layout kernel (positions by arithmetic, layout verified), the single-pass kernel (decoupled
    look-back, local line state) and the three-pass fallback (lines longer than the look-back window) give the oracle's
    code stream for wrapped, unwrapped, CRLF, multi-record and odd inputs — each mode handing over to the next on the
    device when the input is not what it handles."""
    if MODE[0] != "partitioned":
        pytest.skip("independent of the counter")
    from oracle import kmers
    from subphaser_b200 import engine
    if mode != "regular":
        monkeypatch.setenv("SPK_PACK_MODE", mode)
    rng = np.random.default_rng(31)
    big = util.messy_seq(rng, 300_000)
    reg = []
    for width in (16, 17, 31, 32, 33, 59, 60, 61, 64, 80, 127, 4096, 8191, 8192, 8193, 70000):
        for n in (0, 1, width - 1, width, width + 1, 8192, 8192 * 3 + 5, 123_457):
            body = big[:n]
            lines = "\n".join(body[i:i + width] for i in range(0, len(body), width))
            # (a first line longer than the 64-KiB probe window is left to the general kernels)
            want_reg = width < 65536 or n < 65536
            reg.append(((">c%d_%d some text\n" % (width, n) + lines + "\n").encode(), want_reg))
            reg.append(((">c\n" + lines).encode(), want_reg))           # no final newline
    reg += [
        (b">a\n" + big[:1000].encode() + b"\n\n", None),                # blank line at the end
        (b">a\n" + big[:1000].encode() + b"\n\n\n", False),
        ((">a\n" + big[:60] + "\n" + big[60:100] + "\n" + big[100:160] + "\n").encode(), False),   # a short line in the middle
        ((">a\n" + big[:60] + "\n" + big[60:120] + "\r\n").encode(), False),
        ((">a\n" + big[:60] + "\n>" + big[60:119] + "\n").encode(), False),   # '>' at the beginning of a line: a second record
        ((">a\n" + big[:30] + ">" + big[30:59] + "\n").encode(), False),      # '>' inside a line: an invalid base
        (b">only a header", None),
        (b">only a header\n", True),
        ((">" + "h" * 70000 + "\n" + big[:100] + "\n").encode(), False),       # header longer than the probe window
    ]
    cases = reg + [(c, None) for c in [
        util.fasta([("a", big)]),                                   # 60-column lines: single pass
        util.fasta([("a", big)], width=100000),                     # lines of 100 kb: look-back window exceeded
        util.fasta([("a", big)], width=511), util.fasta([("a", big)], width=512), util.fasta([("a", big)], width=513),
        util.fasta([("a", big)], width=61, crlf=True),
        util.fasta([("r%d" % i, util.messy_seq(rng, int(rng.integers(0, 9000)))) for i in range(40)], width=70),
        util.fasta([("h" * 700, big[:5000]), ("x" * 5000 + " long header", big[5000:9000])]),   # headers longer than the window
        b">a\n" + big[:4093].encode() + b"\n>b\n" + big[:100].encode(),          # header at a tile boundary, no final newline
        big[:20000].encode(),                                        # no header, no newline at all
        b"\n\n>x\n\nAC\n\nGT\n>y\n>z\nN\n",
    ]]
    for ci, (fa, want_reg) in enumerate(cases):
        want, nrec = kmers.fasta_to_codes(fa)
        d, n = engine.to_device_bytes(fa)
        seq = engine.pack_fasta(d, n)
        if mode == "regular" and want_reg is not None:
            # wrapped single records take the arithmetic path, and its layout check rejects everything else
            assert (seq.pack_path == "regular") == want_reg, (ci, fa[:40], len(fa), seq.pack_path)
        elif mode != "regular":
            assert seq.pack_path != "regular"
        assert seq.n_bases == len(want) and seq.n_records == nrec
        got = _unpack(seq)
        np.testing.assert_array_equal(got, np.minimum(want, 4))
        assert seq.n_valid == int(np.sum(want < 4))


@pytest.mark.parametrize("table", ["list", "sweep", "versioned"])
def test_v3_dense_partitions_and_histogram(table, monkeypatch):
    """70 Mb of random sequence, lower_count 1: ~3000 distinct k-mers per partition, all dumped (more than the
    kept-slot list of the versioned table holds: scan fallback), and the count histogram path (`histo_len`)."""
    if MODE[0] != "partitioned":
        pytest.skip("partitioned counter only")
    monkeypatch.setenv("SPK_PCOUNT_TABLE", table)
    from oracle import kmers
    from subphaser_b200 import engine
    rng = np.random.default_rng(99)
    codes = rng.integers(0, 4, 70_000_000, dtype=np.uint8)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    fa = b">r\n" + seq.tobytes() + b"\n"
    d, n = engine.to_device_bytes(fa)
    ps = engine.pack_fasta(d, n)
    del d
    table = engine.CountTable(ps.n_bases, 17, 1, mode="partitioned")
    assert table.pbits == 15
    dump = engine.count_packed(ps, 17, 1, table=table, histo_len=300)
    keys, counts = dump.to_host()
    okeys, ocounts, st = kmers.count_fasta(fa, 17, 1, nthreads=8)
    o = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(keys[o], okeys)
    np.testing.assert_array_equal(counts[o], ocounts)
    assert dump.n_distinct == st["n_distinct"] and dump.length == st["sum_dumped"]
    h = dump.histo.cpu().numpy()
    want = np.bincount(np.minimum(ocounts, 299).astype(np.int64), minlength=300)
    np.testing.assert_array_equal(h[1:], want[1:])
    # without a histogram ("list": every partition keeps more slots than its list holds -> swept one by one)
    dump2 = engine.count_packed(ps, 17, 1, table=table)
    keys2, counts2 = dump2.to_host()
    o2 = np.argsort(keys2, kind="stable")
    np.testing.assert_array_equal(keys2[o2], okeys)
    np.testing.assert_array_equal(counts2[o2], ocounts)
    np.testing.assert_array_equal(dump2.pindex.cpu().numpy()[1::2], dump.pindex.cpu().numpy()[1::2])
