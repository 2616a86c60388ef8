"""GPU: the drop-in module surface beyond the main pipeline fixture — Stats.enrich_ltr, the error paths of
JellyfishDumps.filter, min_prop / max_prop, Cluster(sg_assigned=...), gz input, jellyfish text dumps, the
partition-overflow fallbacks and the host-buffer C entry point — against reference-generated fixtures
(tests/golden/*.json, made by tests/golden/make_golden.py) or the CPU oracle."""
import ctypes
import gzip
import io
import json
import os

import numpy as np
import pytest

import spk_testutil as util

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


# ---- a13: Stats.enrich_ltr (Stats.py:33-73) -------------------------------------------------------------------
def test_enrich_ltr_matches_reference_text():
    from subphaser_b200 import Stats
    for case in load("enrich_ltr.json"):
        out = io.StringIO()
        rownames = [tuple(r) for r in case["rownames"]]
        d_enriched, d_exchange = Stats.enrich_ltr(out, case["d_sg"], case["matrix"], colnames=case["colnames"],
                                                  rownames=rownames, max_pval=0.05, ncpu=1)
        assert d_enriched == case["d_enriched"]
        assert d_exchange == case["d_exchange"]
        ref = [l.split("\t") for l in case["text"].splitlines()]
        got = [l.split("\t") for l in out.getvalue().splitlines()]
        assert len(ref) == len(got) and got[0] == ref[0]
        for r, g in zip(ref[1:], got[1:]):
            assert g[0] == r[0] and g[1] == r[1] and g[3] == r[3] and g[4] == r[4]     # id, subgenome, counts, exchange
            assert float(g[2]) == pytest.approx(float(r[2]), abs=1e-10, rel=1e-9)      # p_value
            assert float(g[5]) == pytest.approx(float(r[5]), abs=1e-10, rel=1e-9)      # p_corrected


def test_enrich_generator_objects_match_table():
    """Stats.enrich keeps the reference's per-row object surface (Stats.py:150-168)."""
    from subphaser_b200 import Stats
    case = load("fisher_enrich.json")[4]
    colnames = ["SG%d" % (i + 1) for i in range(case["S"])]
    rows = [r["row"] for r in case["rows"]]
    names = [("c", i * 10, i * 10 + 10) for i in range(len(rows))]
    for res, ref in zip(Stats.enrich(rows, colnames=colnames, rownames=names, max_pval=0.05), case["rows"]):
        assert res.idx == ref["idx"] and res.sig == ref["sig"] and res.key == ref["key"]
        assert list(res.enrich) == ref["enrich"] and list(res.counts) == ref["row"]
        np.testing.assert_allclose(res.pvals, ref["pvals"], rtol=1e-9, atol=1e-10)
        assert [repr(float(x)) for x in res.ratios] == [repr(x) for x in ref["ratios"]] or np.allclose(res.ratios, ref["ratios"], rtol=0, atol=0, equal_nan=True)


# ---- a3: JellyfishDumps.filter error paths and proportional thresholds (Jellyfish.py:462-512) -------------------
@pytest.fixture(scope="module")
def small_dumps(tmp_path_factory):
    from subphaser_b200 import Jellyfish
    work = tmp_path_factory.mktemp("filt")
    records, sgs = util.subgenome_genome(5, n_sg=2, chr_per_sg=2, chr_len=20000, n_fam=4, fam_len=300)
    files = []
    for name, seq in records:
        p = os.path.join(work, name + ".fasta")
        with open(p, "wb") as f:
            f.write(util.fasta([(name, seq)]))
        files.append(p)
    labels = [n for n, _ in records]
    os.environ["SPK_DUMP_SIDECAR"] = "1"        # other tests clear the in-process registry: keep the dumps loadable
    try:
        dumpfiles = Jellyfish.run_jellyfish_dumps(files, k=13, ncpu=1, lower_count=2, threads=1, overwrite=True)
    finally:
        del os.environ["SPK_DUMP_SIDECAR"]
    return files, labels, sgs, dumpfiles


def test_filter_value_errors(small_dumps):
    from subphaser_b200 import Jellyfish
    files, labels, sgs, dumpfiles = small_dumps
    dumps = Jellyfish.JellyfishDumps(dumpfiles, labels, ncpu=1)
    d_mat = dumps.to_matrix()
    lengths = dumps.lengths
    with pytest.raises(ValueError, match="should be lower than `max_freq`"):          # Jellyfish.py:474-475
        dumps.filter(d_mat, lengths, sgs, outfig="x", min_freq=500, max_freq=100)
    with pytest.raises(ValueError, match="All singletons are not allowed"):            # :482-483
        dumps.filter(d_mat, lengths, [[[labels[0]]], [[labels[1]]]], outfig="x")
    with pytest.raises(ValueError, match="have only 0 kmers"):                         # :487-489
        dumps2 = Jellyfish.JellyfishDumps(dumpfiles, labels, ncpu=1)
        dm2 = dumps2.to_matrix()
        dumps2.lengths = [0] + list(dumps2.lengths[1:])
        dumps2.filter(dm2, dumps2.lengths, sgs, outfig="x")
    with pytest.raises(ValueError, match="0 kmer with fold"):                          # :508-509
        dumps.filter(d_mat, lengths, sgs, outfig=os.path.join(os.path.dirname(files[0]), "h.pdf"), min_fold=1e30,
                     min_freq=1)
    with pytest.raises(IndexError):                                                    # freqs[baseline] of :639-642
        dumps.filter(d_mat, lengths, sgs, outfig="x", baseline=2)


def test_filter_min_prop_max_prop(small_dumps):
    """min_prop / max_prop rescale the frequency gate by the total dumped length (Jellyfish.py:467-473)."""
    from oracle import restate
    from subphaser_b200 import Jellyfish, engine
    files, labels, sgs, dumpfiles = small_dumps
    dumps = Jellyfish.JellyfishDumps(dumpfiles, labels, ncpu=1)
    d_mat = dumps.to_matrix()
    tot = sum(dumps.lengths)
    fig = os.path.join(os.path.dirname(files[0]), "p.pdf")
    dm = dumps.filter(d_mat, dumps.lengths, sgs, outfig=fig, min_prop=5.0 / tot, max_prop=900.0 / tot,
                      min_freq=10**9, max_freq=0)          # overridden by the proportions
    dm_ref = dumps.filter(dumps.to_matrix(), dumps.lengths, sgs, outfig=fig, min_freq=5.0, max_freq=900.0)
    assert len(dm) == len(dm_ref) > 0
    np.testing.assert_array_equal(engine.u64_numpy(dm.keys), engine.u64_numpy(dm_ref.keys))
    # and against the oracle restatement of _filter_kmer
    host = [Jellyfish.load_dump(d).to_host() for d in dumpfiles]
    allk, mat, lengths = restate.to_matrix(host)
    okeys, onorm, _, _ = restate.filter_matrix(allk, mat, lengths, labels, sgs, min_freq=5.0, max_freq=900.0,
                                               min_fold=2, baseline=1, ratio=1)
    np.testing.assert_array_equal(engine.u64_numpy(dm.keys), okeys)
    assert dm.norm.cpu().numpy().tobytes() == onorm.tobytes()


# ---- a5: Cluster with sg_assigned (Cluster.py:35-42) ------------------------------------------------------------
def test_cluster_sg_assigned_skips_kmeans(tmp_path):
    from subphaser_b200.Cluster import Cluster
    Gp = os.path.join(G, "pipeline_small")
    meta = json.load(open(os.path.join(Gp, "meta.json")))
    ref_sg = meta["d_sg"]
    # user-supplied labels as `-sg_assigned` gives them (any hashable label per chromosome)
    assigned = {c: ("x" if sg == "SG2" else "y") for c, sg in ref_sg.items()}
    cl = Cluster(os.path.join(Gp, "ref.kmer.mat"), n_clusters=5, sg_prefix="SG", sg_assigned=assigned, replicates=0)
    assert not hasattr(cl, "kmean")
    assert cl.n_clusters == 2
    # relabelled by first appearance over name-sorted chromosomes (Cluster.py:119-126): same partition as the reference run
    assert dict(cl.d_sg) == ref_sg
    assert cl.sg_names == sorted(set(ref_sg.values()))
    out = io.StringIO()
    d_kmers = cl.output_kmers(out, max_pval=0.05)
    ref_lines = open(os.path.join(Gp, "ref.sig.kmer-subgenome.tsv")).read().splitlines()
    got_lines = out.getvalue().splitlines()
    assert sorted(l.split("\t")[0] for l in got_lines[1:]) == sorted(l.split("\t")[0] for l in ref_lines[1:])
    assert len(d_kmers) == meta["n_sig"]
    cl2 = Cluster(os.path.join(Gp, "ref.kmer.mat"), n_clusters=2, sg_assigned=assigned, re_assign=False, replicates=0)
    assert cl2.d_sg == assigned


# ---- ingest variants -------------------------------------------------------------------------------------------
def test_gz_fasta_and_multi_file_input(tmp_path):
    """`zcat` leg of Jellyfish.py:696 and the several-files form of :682-684."""
    from oracle import kmers
    from subphaser_b200 import Jellyfish
    rng = np.random.default_rng(9)
    a = util.fasta([("a", util.messy_seq(rng, 30000))])
    b = util.fasta([("b1", util.messy_seq(rng, 9000)), ("b2", util.random_seq(rng, 500))], width=70)
    pa, pb = str(tmp_path / "a.fasta.gz"), str(tmp_path / "b.fasta")
    with gzip.open(pa, "wb") as f:
        f.write(a)
    with open(pb, "wb") as f:
        f.write(b)
    for files, data, prefix in ((pa, a, None), ([pa, pb], a + b"\n" + b, str(tmp_path / "both"))):
        out = Jellyfish.run_jellyfish_dump(files, k=15, lower_count=2, prefix=prefix, overwrite=True)
        keys, counts = Jellyfish.load_dump(out).to_host()
        okeys, ocounts, _ = kmers.count_fasta(data, 15, 2)
        o = np.argsort(keys, kind="stable")
        np.testing.assert_array_equal(keys[o], okeys)
        np.testing.assert_array_equal(counts[o], ocounts)
        # no jellyfish-style marker for a dump whose text was not written (a stock run must not trust it)
        assert not os.path.exists(out + ".ok") and os.path.getsize(out) == 0


def test_jellyfish_text_dump_roundtrip(tmp_path, monkeypatch):
    """SPK_TEXT_DUMPS=1 writes the byte-compatible `KMER COUNT` text (Jellyfish.py:699); a text dump without side-car
    (left by real jellyfish) is ingested through the plain union path and gives the same matrix."""
    from subphaser_b200 import Jellyfish, _registry, engine
    records, sgs = util.subgenome_genome(6, n_sg=2, chr_per_sg=2, chr_len=15000, n_fam=4, fam_len=300)
    labels = [n for n, _ in records]
    files = []
    for name, seq in records:
        p = str(tmp_path / (name + ".fasta"))
        with open(p, "wb") as f:
            f.write(util.fasta([(name, seq)]))
        files.append(p)
    monkeypatch.setenv("SPK_TEXT_DUMPS", "1")
    dumpfiles = Jellyfish.run_jellyfish_dumps(files, k=13, ncpu=1, lower_count=2, threads=1, overwrite=True)
    dumps = Jellyfish.JellyfishDumps(dumpfiles, labels)
    dm = dumps.filter(dumps.to_matrix(), dumps.lengths, sgs, outfig=None, min_freq=5)
    from oracle import kmers
    for f, d in zip(files, dumpfiles):
        okeys, ocounts, _ = kmers.count_fasta(open(f, "rb").read(), 13, 2)
        want = sorted("%s %d" % (kmers.key_to_str(a, 13), b) for a, b in zip(okeys, ocounts))
        assert sorted(open(d).read().splitlines()) == want
        assert os.path.exists(d + ".ok") and not os.path.exists(d + Jellyfish.SIDE_SUFFIX)   # what real jellyfish leaves
    _registry.clear()
    dumps2 = Jellyfish.JellyfishDumps(dumpfiles, labels)
    cm = dumps2.to_matrix()
    assert isinstance(cm, engine.CountMatrix)          # no partition index in a text dump: plain union
    assert dumps2.lengths == dumps.lengths
    dm2 = dumps2.filter(cm, dumps2.lengths, sgs, outfig=None, min_freq=5)
    np.testing.assert_array_equal(engine.u64_numpy(dm2.keys), engine.u64_numpy(dm.keys))
    assert dm2.norm.cpu().numpy().tobytes() == dm.norm.cpu().numpy().tobytes()


# ---- overflow fallbacks ----------------------------------------------------------------------------------------
def test_partition_overflow_falls_back_to_global_counter(monkeypatch):
    """A chromosome whose k-mers overflow a shared-memory partition table must still be counted (global table)."""
    from oracle import kmers
    from subphaser_b200 import engine
    rng = np.random.default_rng(12)
    fa = util.fasta([("c", util.random_seq(rng, 300_000))])
    d, n = engine.to_device_bytes(fa)
    seq = engine.pack_fasta(d, n)
    monkeypatch.setenv("SPK_PCOUNT_FORCE_FAIL", "1")   # test hook: the partitioned counter reports failed inserts
    dump = engine.count_packed(seq, 17, 2, table=engine.CountTable(seq.n_bases, 17, 2, mode="partitioned"))
    monkeypatch.delenv("SPK_PCOUNT_FORCE_FAIL")
    keys, counts = dump.to_host()
    okeys, ocounts, st = kmers.count_fasta(fa, 17, 2)
    o = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(keys[o], okeys)
    np.testing.assert_array_equal(counts[o], ocounts)
    assert dump.pindex is None and dump.length == st["sum_dumped"]


def test_pmatrix_overflow_falls_back_to_plain_union(small_dumps, monkeypatch):
    """spk_pmatrix_filter reports a partition that does not fit its shared-memory table -> plain union + filter."""
    from subphaser_b200 import Jellyfish, engine
    files, labels, sgs, dumpfiles = small_dumps
    dumps = Jellyfish.JellyfishDumps(dumpfiles, labels)
    want = dumps.filter(dumps.to_matrix(), dumps.lengths, sgs, outfig=None, min_freq=5)
    monkeypatch.setenv("SPK_PMATRIX_FORCE_OVERFLOW", "1")
    dumps2 = Jellyfish.JellyfishDumps(dumpfiles, labels)
    cm = dumps2.to_matrix()
    assert isinstance(cm, engine.CountMatrix)
    got = dumps2.filter(cm, dumps2.lengths, sgs, outfig=None, min_freq=5)
    monkeypatch.delenv("SPK_PMATRIX_FORCE_OVERFLOW")
    np.testing.assert_array_equal(engine.u64_numpy(got.keys), engine.u64_numpy(want.keys))
    assert got.norm.cpu().numpy().tobytes() == want.norm.cpu().numpy().tobytes()


# ---- the host-buffer C entry point -------------------------------------------------------------------------------
def test_count_fasta_host_entry_point():
    """spk_count_fasta_host: FASTA bytes in HOST memory in, dump on the device out (include/spk.h)."""
    import torch
    from oracle import kmers
    from subphaser_b200 import _lib, engine
    lib = _lib.load()
    rng = np.random.default_rng(4)
    fa = util.fasta([("h", util.messy_seq(rng, 120_000, repeat_unit="ACGT"))])
    k, lower = 17, 2
    cap = len(fa)
    dev = "cuda"
    d_ascii = torch.empty(cap + 16, dtype=torch.uint8, device=dev)
    d_packed = torch.empty(lib.spk_packed_words(cap), dtype=torch.int32, device=dev)
    d_valid = torch.empty(lib.spk_valid_words(cap), dtype=torch.int32, device=dev)
    ws_bytes = lib.spk_pack_workspace_bytes(cap)
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    table_bytes = lib.spk_count_table_bytes(cap, k)
    d_table = torch.empty(table_bytes, dtype=torch.uint8, device=dev)
    d_blocks = torch.empty(3 * lib.spk_table_scan_blocks() + 2, dtype=torch.int32, device=dev)
    keys = torch.empty(cap, dtype=torch.int64, device=dev)
    counts = torch.empty(cap, dtype=torch.int32, device=dev)
    d_info = torch.zeros(12, dtype=torch.int64, device=dev)
    h_out = (ctypes.c_uint64 * 8)()
    buf = np.frombuffer(fa, dtype=np.uint8)
    P = engine._p
    _lib.call("spk_count_fasta_host", buf.ctypes.data_as(ctypes.c_void_p), len(fa), k, lower, P(d_ascii), P(d_packed),
              P(d_valid), cap, P(d_ws), ws_bytes, P(d_table), table_bytes, P(d_blocks), P(keys), P(counts), cap,
              P(d_info), ctypes.cast(h_out, ctypes.c_void_p), engine._stream())
    torch.cuda.synchronize()
    okeys, ocounts, ost = kmers.count_fasta(fa, k, lower)
    out = [int(x) for x in h_out]
    assert out[0] == ost["n_bases"] and out[2] == ost["n_records"] and out[3] == ost["n_valid_kmers"]
    assert out[4] == ost["n_distinct"] and out[5] == ost["n_dumped"] and out[6] == ost["sum_dumped"] and out[7] == 0
    n = out[5]
    gk = engine.u64_numpy(keys[:n])
    gc = counts[:n].cpu().numpy().view(np.uint32)
    o = np.argsort(gk, kind="stable")
    np.testing.assert_array_equal(gk[o], okeys)
    np.testing.assert_array_equal(gc[o], ocounts)


# ---- f1: genome ingest -----------------------------------------------------------------------------------------
def test_split_genomes_matches_reference(tmp_path):
    """Seqs.split_genomes (Seqs.py:27-71) with the records found and packed on the device: same files (byte for byte),
    labels, id map and sizes as the reference's BioPython loop; the second genome is read from its .gz; the packed
    chromosomes are registered, so counting never re-reads the files and gives the oracle's k-mers."""
    from oracle import kmers
    from subphaser_b200 import Jellyfish, Seqs, _registry
    fx = load("split_genomes.json")
    ga, gb = tmp_path / "A.fa", tmp_path / "B.fa.gz"
    ga.write_bytes(fx["genomes"]["A"].encode())
    with gzip.open(gb, "wb") as f:
        f.write(fx["genomes"]["B_plain"].encode("latin1"))
    for ci, case in enumerate(fx["cases"]):
        outdir = str(tmp_path / ("chroms%d" % ci)) + "/"
        os.makedirs(outdir)
        _registry.clear()
        outfas, labels, d_t2, d_size = Seqs.split_genomes([str(ga), str(gb)], case["prefixes"], case["targets"], outdir)
        assert [os.path.basename(p) for p in outfas] == case["files"]
        assert labels == case["labels"] and dict(d_t2) == case["d_targets2"] and d_size == case["d_size"]
        for p in outfas:
            assert open(p).read() == case["contents"][os.path.basename(p)], p
            seq = _registry.get_seq(p)
            assert seq is not None and seq.n_bases == d_size[labels[outfas.index(p)]]
        # counting uses the registered sequence (the file is not read again) and matches the oracle on the file
        big = outfas[0]
        data = open(big, "rb").read()
        with open(big, "r+b") as f:          # same size and mtime-preserving overwrite is not possible: check via result only
            pass
        out = Jellyfish.run_jellyfish_dump(big, k=11, lower_count=1, overwrite=True)
        keys, counts = Jellyfish.load_dump(out).to_host()
        okeys, ocounts, _ = kmers.count_fasta(data, 11, 1)
        o = np.argsort(keys, kind="stable")
        np.testing.assert_array_equal(keys[o], okeys)
        np.testing.assert_array_equal(counts[o], ocounts)
    with pytest.raises(KeyError):            # no targets at all: `d_targets[rc.id]` fails in the reference too (Seqs.py:61)
        Seqs.split_genomes([str(ga)], [""], [], str(tmp_path) + "/x_")


def test_dump_checkpoints_and_registry_validation(tmp_path, monkeypatch):
    """`.ok` only with a text dump, `.spk.ok` with the binary side-car (SPK_DUMP_SIDECAR=1), and the in-process registry
    does not hand out a dump made with another lower_count or from a file that changed."""
    from subphaser_b200 import Jellyfish, _registry
    rng = np.random.default_rng(3)
    fa = tmp_path / "c.fasta"
    fa.write_bytes(util.fasta([("c", util.messy_seq(rng, 20000, repeat_unit="ACGTT"))]))
    monkeypatch.setenv("SPK_DUMP_SIDECAR", "1")
    out = Jellyfish.run_jellyfish_dump(str(fa), k=13, lower_count=3, overwrite=True)
    assert os.path.exists(out + ".spk.ok") and os.path.exists(out + Jellyfish.SIDE_SUFFIX) and not os.path.exists(out + ".ok")
    n3 = len(Jellyfish.load_dump(out))
    assert Jellyfish.run_jellyfish_dump(str(fa), k=13, lower_count=3) == out          # registry hit
    out1 = Jellyfish.run_jellyfish_dump(str(fa), k=13, lower_count=1)                 # other threshold: recounted
    assert out1 == out and len(Jellyfish.load_dump(out)) > n3
    _registry.clear()
    assert len(Jellyfish.load_dump(out)) > n3                                         # from the side-car of the last run
    fa.write_bytes(util.fasta([("c", util.random_seq(rng, 5000))]))                   # the chromosome file changes
    monkeypatch.delenv("SPK_DUMP_SIDECAR")
    out2 = Jellyfish.run_jellyfish_dump(str(fa), k=13, lower_count=1)
    # (the stale side-car is trusted only through its own marker and parameters; a changed source with the registry
    #  cleared is the caller's `overwrite` business, as with the reference's .ok files)
    assert out2 == out
