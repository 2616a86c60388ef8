"""GPU: each libspk kernel family against the reference-generated golden vectors (tests/golden/*.json)
and the CPU oracle (oracle/restate.py), through the C ABI (ctypes) wrappers of subphaser_b200.engine."""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def _dev(a, dtype):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).cuda()


# ---- K4 filter ------------------------------------------------------------------------------------------
def test_filter_matches_reference_vectors():
    import torch
    from subphaser_b200 import engine
    for case in load("filter_kmer.json"):
        rows = case["rows"]
        n = len(case["labels"])
        mat = np.array([r["counts"] for r in rows], dtype=np.int32)
        cm = engine.CountMatrix(_dev(mat, np.int32), _dev(np.arange(len(rows)), np.int64), case["lengths"], 15,
                                case["labels"])
        p = case["params"]
        dm = engine.filter_matrix(cm, case["sgs"], case["labels"], min_fold=p["min_fold"], baseline=p["baseline"],
                                  ratio=p["ratio"], min_freq=p["min_freq"], max_freq=p["max_freq"],
                                  want_fold_tots=True)
        kept = [i for i, r in enumerate(rows) if r["freqs"]]
        fold = [i for i, r in enumerate(rows) if r["tot"] is not None]
        assert dm.n_fold_pass == len(fold)
        assert engine.u64_numpy(dm.keys).tolist() == kept          # keys are the row ids, sorted ascending
        norm = dm.norm.cpu().numpy()
        want = np.array([rows[i]["freqs"] for i in kept], dtype=np.float64).reshape(len(kept), n)
        assert norm.tobytes() == want.tobytes()                   # bit-exact fp64
        assert dm.tot.cpu().numpy().tolist() == [rows[i]["tot"] for i in kept]
        # the totals of the fold-passing rows stay on the device; their histogram / order statistics are what is exposed
        ref_tots = np.array(sorted(rows[i]["tot"] for i in fold), dtype=np.float64)
        fh = dm.fold_tots
        assert fh.n == len(ref_tots) and fh.min == ref_tots.min() and fh.max == ref_tots.max()
        nb = max(int(int(ref_tots.max()) / 25), 1)
        np.testing.assert_array_equal(fh.hist, np.histogram(ref_tots, bins=nb)[0])
        assert fh.xlim == float(np.percentile(ref_tots, 99))
        assert [fh._order_stat(r) for r in range(0, len(ref_tots), max(len(ref_tots) // 7, 1))] == \
            [int(x) for x in ref_tots[::max(len(ref_tots) // 7, 1)]]


def test_union_matrix_matches_oracle():
    from oracle import kmers, restate
    from subphaser_b200 import engine
    import spk_testutil as util
    records, sgs = util.subgenome_genome(3, n_sg=3, chr_per_sg=2, chr_len=20000)
    dumps_o, dumps_g = [], []
    for name, seq in records:
        fa = util.fasta([(name, seq)])
        k_, c_, _ = kmers.count_fasta(fa, 13, 2)
        dumps_o.append((k_, c_))
        d, n = engine.to_device_bytes(fa)
        dumps_g.append(engine.count_packed(engine.pack_fasta(d, n), 13, 2))
    allk, mat, lengths = restate.to_matrix(dumps_o)
    cm = engine.build_matrix(dumps_g)
    assert cm.lengths == lengths and len(cm) == len(allk)
    rk = engine.u64_numpy(cm.row_keys)
    order = np.argsort(rk)
    np.testing.assert_array_equal(rk[order], allk)
    np.testing.assert_array_equal(cm.matrix.cpu().numpy()[order].astype(np.int64), mat)


@pytest.mark.parametrize("k,lower,chr_len,min_freq", [(13, 2, 20000, 8), (17, 1, 60000, 20), (21, 3, 30000, 10)])
def test_partitioned_union_filter_equals_plain_path(k, lower, chr_len, min_freq):
    """spk_pmatrix_filter (shared-memory union + filter per hash partition, dumps indexed by
    spk_pcount_canonical_ex with common partition bits) == spk_union_insert/matrix_fill/filter_* and the
    oracle restatement of to_matrix + filter, bit for bit."""
    from oracle import kmers, restate
    from subphaser_b200 import engine
    import spk_testutil as util
    records, sgs = util.subgenome_genome(k, n_sg=3, chr_per_sg=2, chr_len=chr_len)
    labels = [r[0] for r in records]
    packed = []
    for name, seq in records:
        d, n = engine.to_device_bytes(util.fasta([(name, seq)]))
        packed.append(engine.pack_fasta(d, n))
    table = engine.CountTable(max(p.n_bases for p in packed), k, lower)
    dumps = [engine.count_packed(p, k, lower, table=table) for p in packed]
    assert engine.can_pmatrix(dumps)
    # the partition index covers the dump exactly once
    for d in dumps:
        idx = d.pindex.cpu().numpy().astype(np.int64).reshape(-1, 2)
        assert idx[:, 1].sum() == len(d)
        nz = idx[idx[:, 1] > 0]
        order = np.argsort(nz[:, 0])
        assert np.array_equal(nz[order, 0], np.concatenate([[0], np.cumsum(nz[order, 1])[:-1]]))
    kw = dict(min_fold=2, baseline=1, ratio=1, min_freq=min_freq, max_freq=10000)
    cm = engine.build_matrix(dumps, labels)
    want = engine.filter_matrix(cm, sgs, labels, want_fold_tots=True, **kw)
    got, n_union = engine.pmatrix_filter(dumps, sgs, labels, want_fold_tots=True, **kw)
    assert n_union == len(cm) and len(got) == len(want) > 0
    assert got.n_fold_pass == want.n_fold_pass
    np.testing.assert_array_equal(engine.u64_numpy(got.keys), engine.u64_numpy(want.keys))
    assert got.norm.cpu().numpy().tobytes() == want.norm.cpu().numpy().tobytes()
    np.testing.assert_array_equal(got.tot.cpu().numpy(), want.tot.cpu().numpy())
    assert got.fold_tots.n == want.fold_tots.n and got.fold_tots.xlim == want.fold_tots.xlim
    np.testing.assert_array_equal(got.fold_tots.hist, want.fold_tots.hist)
    # two "ranks": the row shards partition the result
    parts = [engine.pmatrix_filter(dumps, sgs, labels, nparts=2, part=r, **kw) for r in range(2)]
    assert parts[0][1] + parts[1][1] == n_union
    both = np.sort(np.concatenate([engine.u64_numpy(p[0].keys) for p in parts]))
    np.testing.assert_array_equal(both, engine.u64_numpy(want.keys))
    # and the oracle
    dumps_o = [kmers.count_fasta(util.fasta([r]), k, lower)[:2] for r in records]
    allk, mat, lengths = restate.to_matrix(dumps_o)
    okeys, onorm, otot, _ = restate.filter_matrix(allk, mat, lengths, labels, sgs, **kw)
    np.testing.assert_array_equal(engine.u64_numpy(got.keys), okeys)
    assert got.norm.cpu().numpy().tobytes() == onorm.tobytes()


@pytest.mark.parametrize("ratio,min_fold", [(1, 2), (0.5, 2), (0.3, 1.5), (1, 0)])
def test_pmatrix_presence_mask_kernel_equals_general_kernel(ratio, min_fold, monkeypatch):
    """k_pmatrix_filter2 (presence masks, candidate rows only) and k_pmatrix_filter (full count rows) emit the same
    candidate rows; ratios that make most rows candidates overflow the candidate buffer of the former, which then
    hands the call to the latter on the device (counters[5]); min_fold 0 never uses the former."""
    from subphaser_b200 import engine
    import spk_testutil as util
    k, lower = 15, 1
    records, sgs = util.subgenome_genome(k, n_sg=4, chr_per_sg=2, chr_len=150000)
    labels = [r[0] for r in records]
    packed = []
    for name, seq in records:
        d, n = engine.to_device_bytes(util.fasta([(name, seq)]))
        packed.append(engine.pack_fasta(d, n))
    table = engine.CountTable(max(p.n_bases for p in packed), k, lower)
    dumps = [engine.count_packed(p, k, lower, table=table) for p in packed]
    kw = dict(min_fold=min_fold, baseline=1, ratio=ratio, min_freq=5, max_freq=10000)
    a, ua = engine.pmatrix_filter(dumps, sgs, labels, **kw)
    monkeypatch.setenv("SPK_PMATRIX_KERNEL", "general")
    b, ub = engine.pmatrix_filter(dumps, sgs, labels, **kw)
    monkeypatch.delenv("SPK_PMATRIX_KERNEL")
    cm = engine.build_matrix(dumps, labels)
    want = engine.filter_matrix(cm, sgs, labels, **kw)
    assert ua == ub == len(cm) and len(a) == len(b) == len(want) > 0
    for got in (a, b):
        np.testing.assert_array_equal(engine.u64_numpy(got.keys), engine.u64_numpy(want.keys))
        assert got.norm.cpu().numpy().tobytes() == want.norm.cpu().numpy().tobytes()


# ---- sort -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,bits", [(1, 8), (2, 64), (1000, 34), (2049, 64), (300000, 42), (1 << 20, 30)])
def test_radix_sort(n, bits):
    import torch
    from subphaser_b200 import engine, _lib
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2**bits if bits < 64 else 2**63, n, dtype=np.uint64)
    if bits == 64:
        keys = keys * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    keys[: n // 3] = keys[n // 3: 2 * (n // 3)][: n // 3]           # duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    dk, dv = _dev(keys.view(np.int64), np.int64), _dev(vals.view(np.int32), np.int32)
    kt, vt = torch.empty_like(dk), torch.empty_like(dv)
    lib = _lib.load()
    wsb = lib.spk_sort_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.call("spk_sort_pairs_u64", engine._p(dk), engine._p(dv), engine._p(kt), engine._p(vt), n, bits,
              engine._p(ws), wsb, engine._stream())
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(engine.u64_numpy(dk), keys[order])
    np.testing.assert_array_equal(dv.cpu().numpy().view(np.uint32), vals[order])


# ---- K9 map + stack -------------------------------------------------------------------------------------
def test_map_and_stack_match_reference_text(tmp_path):
    import io
    from subphaser_b200 import Circos, Seqs, _registry
    for ci, case in enumerate(load("map_stack.json")):
        fa = tmp_path / ("c%d.fasta" % ci)
        seq = case["seq"]
        fa.write_text(">chrX desc\n" + "\n".join(seq[i:i + 60] for i in range(0, len(seq), 60)) + "\n")
        out = tmp_path / ("c%d.bin.count" % ci)
        with open(out, "w") as f:
            Seqs.map_kmer3([str(fa)], case["d_kmers"], fout=f, k=case["k"], window_size=case["window_size"],
                           bin_size=case["bin_size"], sg_names=case["sg_names"], ncpu=1, chunk=case["chunk"])
        assert out.read_text() == case["bin_count_text"]
        for reg in (True, False):                       # registry shortcut and text re-parse
            if not reg:
                _registry.clear()
            for ws, st in case["stacks"].items():
                coords, counts = Circos.stack_matrix(str(out), window_size=float(ws) if "." in ws else int(ws))
                assert [list(c) for c in coords] == st["coords"]
                assert counts == st["counts"]


def test_map_multi_record_fasta_matches_reference_text(tmp_path):
    """Seqs.map_kmer3 on multi-record FASTA (the LTR / custom-feature calls use chunk=False, one row per
    sequence): byte-identical to the text the reference's own code wrote (tests/golden/map_multi.json)."""
    from subphaser_b200 import Seqs, _registry
    import spk_testutil as util
    for ci, case in enumerate(load("map_multi.json")):
        fa = tmp_path / ("m%d.fasta" % ci)
        fa.write_bytes(util.fasta([(n, s) for n, s in case["records"]]))
        out = tmp_path / ("m%d.bin.count" % ci)
        _registry.clear()
        with open(out, "w") as f:
            Seqs.map_kmer3([str(fa)], case["d_kmers"], fout=f, k=case["k"], window_size=case["window_size"],
                           bin_size=case["bin_size"], sg_names=case["sg_names"], ncpu=1, chunk=case["chunk"])
        assert out.read_text() == case["bin_count_text"]


# ---- K10 Fisher / enrich / BH ------------------------------------------------------------------------------
def test_fisher_enrich_bh_match_reference_vectors():
    from subphaser_b200 import engine
    for case in load("fisher_enrich.json"):
        mat = np.array([r["row"] for r in case["rows"]], dtype=np.int64)
        res = engine.fisher_enrich(mat)
        assert res["totals"].tolist() == case["total"]
        want_p = np.array([r["pvals"] for r in case["rows"]])
        # tolerance of the north star: 1e-10 absolute vs the scipy path; we also demand 1e-8 relative
        # down to 1e-290 (scipy/Boost itself is only ~1e-9 accurate for p < 1e-50 at N ~ 1e8, see
        # test_fisher_against_exact_rational for the ground truth)
        # For margins >~ 1e6 scipy/Boost's Lanczos path is itself ~8e-10 off the exact value (e.g. the
        # table (5,10,1683053,3820424): exact 0.5044287592884893 = ours, scipy 0.5044287588910722);
        # test_fisher_against_exact_rational / test_hypergeom_mass_is_one pin OUR accuracy to 1e-12.
        atol = 1e-10 if case["scale"] <= 1000 else 1e-9
        np.testing.assert_allclose(res["pvals"], want_p, rtol=0, atol=atol)
        np.testing.assert_allclose(res["pvals"], want_p, rtol=1e-8, atol=1e-300)
        assert res["idx"].tolist() == [r["idx"] for r in case["rows"]]
        assert res["sig"].tolist() == [r["sig"] for r in case["rows"]]
        want_r = np.array([r["ratios"] for r in case["rows"]])
        assert np.array_equal(np.isnan(res["ratios"]), np.isnan(want_r))
        assert np.nan_to_num(res["ratios"]).tobytes() == np.nan_to_num(want_r).tobytes()   # bit-exact
        np.testing.assert_allclose(res["qvals"], case["qvals"], rtol=1e-9, atol=1e-300)


def test_fisher_against_scipy_wide_range():
    from scipy.stats import hypergeom
    from subphaser_b200 import Stats
    rng = np.random.default_rng(5)
    for _ in range(60):
        S = int(rng.integers(2, 6))
        scale = int(10 ** rng.uniform(0.5, 8.2))
        total = [int(x) for x in rng.integers(scale // 2 + 1, scale + 2, S)]
        each = [int(rng.integers(0, min(t, max(2, scale // int(rng.integers(1, 50)))) + 1)) for t in total]
        got = Stats.fisher_test(each, total)
        se, st = sum(each), sum(total)
        for i in range(S):
            x11, x12 = each[i], se - each[i]
            x21 = total[i] - x11
            x22 = st - x21 - x12
            x21, x22 = min(x21, Stats.MAX_INT), min(x22, Stats.MAX_INT)
            want = float(hypergeom.sf(x11 - 1, x11 + x12 + x21 + x22, x11 + x12, x11 + x21))
            assert got[i] == pytest.approx(want, abs=1e-10, rel=1e-8)


def _exact_right_tail(x11, x12, x21, x22):
    """P(X >= x11) by exact integer arithmetic (term recurrence on Python ints, one final division)."""
    from fractions import Fraction
    from math import comb
    N, K, n = x11 + x12 + x21 + x22, x11 + x21, x11 + x12
    hi = min(n, K)
    if x11 > hi:
        return 0.0
    t = comb(K, x11) * comb(N - K, n - x11)
    num = t
    for x in range(x11, hi):
        t = t * ((K - x) * (n - x)) // ((x + 1) * (N - K - n + x + 1))
        num += t
    return float(Fraction(num, comb(N, n)))


def test_fisher_against_exact_rational():
    """Ground truth by exact integer arithmetic: the kernel is accurate to 1e-12 relative."""
    from subphaser_b200 import Stats
    rng = np.random.default_rng(17)
    worst = 0.0
    for _ in range(25):
        S = int(rng.integers(2, 4))
        scale = int(10 ** rng.uniform(1, 4.0))
        total = [int(x) for x in rng.integers(scale // 2 + 1, scale + 2, S)]
        each = [int(rng.integers(0, min(t, max(2, scale // int(rng.integers(1, 30)))) + 1)) for t in total]
        got = Stats.fisher_test(each, total)
        se, st = sum(each), sum(total)
        for i in range(S):
            x11, x12 = each[i], se - each[i]
            x21 = total[i] - x11
            x22 = st - x21 - x12
            exact = _exact_right_tail(x11, x12, x21, x22)
            if exact > 1e-300:
                worst = max(worst, abs(got[i] - exact) / exact)
                assert got[i] == pytest.approx(exact, rel=1e-12, abs=1e-300)
            else:
                assert got[i] <= 1e-299
    assert worst < 1e-12
    # the golden vectors where scipy (the reference-side stand-in for `fisher`) and the kernel differ by
    # ~4e-10: exact rational arithmetic sides with the kernel
    for tab in ((5, 10, 1683053, 3820424), (5, 15, 1746320, 6433200)):
        x11, x12, x21, x22 = tab
        # rebuild `each`/`total` such that fisher_test() forms exactly this table (Stats.py:20-23)
        each, total = [x11, x12], [x11 + x21, x12 + x22 - x11]
        got = Stats.fisher_test(each, total)[0]
        assert got == pytest.approx(_exact_right_tail(*tab), rel=1e-12)


def test_hypergeom_mass_is_one():
    """Point masses + recurrences of the Fisher kernel sum to 1 within 1e-12 for margins up to 2^31."""
    import torch
    from subphaser_b200 import _lib, engine
    rng = np.random.default_rng(23)
    trip = []
    for _ in range(200):
        N = int(10 ** rng.uniform(1, 9.3))
        K = int(rng.integers(0, N + 1))
        n = int(rng.integers(0, N + 1))
        trip.append((N, K, n))
    trip += [(660_000_000, 230_000_000, 235_000_000), (2_000_000_000, 214_748_364, 900_000_000), (10, 0, 5), (7, 7, 7)]
    d = _dev(np.array(trip, dtype=np.int64), np.int64)
    out = torch.empty(len(trip), dtype=torch.float64, device="cuda")
    _lib.call("spk_debug_hypergeom_mass", engine._p(d), len(trip), engine._p(out), engine._stream())
    np.testing.assert_allclose(out.cpu().numpy(), 1.0, rtol=0, atol=1e-12)


def test_bh_edge_cases():
    from oracle import restate
    from subphaser_b200 import Stats
    for p in ([0.5], [0.0, 0.0, 1.0], [1e-300, 0.04, 0.04, 0.9, 0.2], list(np.linspace(0, 1, 1000)),
              list(np.random.default_rng(0).random(5000) ** 3)):
        np.testing.assert_allclose(Stats.correct_pvals(p), restate.bh(p), rtol=1e-15, atol=0)
    assert len(Stats.correct_pvals([])) == 0


# ---- K5-K8 cluster statistics ------------------------------------------------------------------------------
def test_zscore_bit_exact_vs_numpy():
    from oracle import restate
    from subphaser_b200 import engine
    rng = np.random.default_rng(0)
    for n in (2, 6, 7, 8, 13, 21, 38, 64, 128):
        raw = rng.random((3000, n)) * 1e-4
        Z = engine.zscore_rows(_dev(raw, np.float64)).cpu().numpy()
        assert Z.tobytes() == np.ascontiguousarray(restate.zscore(raw)).tobytes()


def test_ttest_rows_match_reference_vectors(monkeypatch):
    from subphaser_b200 import engine
    for case in load("ttest_rows.json"):
        sgs = sorted(case["groups"])
        n = case["n"]
        singleton = any(len(v) == 1 for v in case["groups"].values())
        if singleton:
            # the fixture comes from the installed scipy (>= 1.9: a one-observation group has variance 0); the default of
            # the kernel is the pinned scipy 1.7.1 behaviour (NaN, which the reference keeps) — checked below
            monkeypatch.setenv("SPK_TTEST_SINGLETON", "zero")
        else:
            monkeypatch.delenv("SPK_TTEST_SINGLETON", raising=False)
        col_group = [0] * n
        for gi, sg in enumerate(sgs):
            for c in case["groups"][sg]:
                col_group[c] = gi
        X = np.array([r["array"] for r in case["rows"]], dtype=np.float64)
        best, pval, means = (t.cpu().numpy() for t in engine.ttest_groups(_dev(X, np.float64), col_group, len(sgs)))
        for i, r in enumerate(case["rows"]):
            assert sgs[best[i]] == r["max_sg"]
            assert means[i].tolist() == r["mean_vals"]            # bit-exact group means
            if math.isnan(r["pvalue"]):
                assert math.isnan(pval[i])
            else:
                assert pval[i] == pytest.approx(r["pvalue"], rel=1e-9, abs=1e-300)
        if singleton:
            monkeypatch.delenv("SPK_TTEST_SINGLETON")
            best2, pval2, _ = (t.cpu().numpy() for t in engine.ttest_groups(_dev(X, np.float64), col_group, len(sgs)))
            sizes = [len(case["groups"][sg]) for sg in sgs]
            np.testing.assert_array_equal(best2, best)
            for i in range(len(X)):
                order = np.argsort([-np.mean(X[i][case["groups"][sg]]) for sg in sgs], kind="stable")
                if 1 in (sizes[order[0]], sizes[order[1]]):
                    assert math.isnan(pval2[i])
                else:
                    assert pval2[i] == pval[i] or (math.isnan(pval2[i]) and math.isnan(pval[i]))


def test_ranktest_rows_match_reference_vectors(tmp_path):
    """a7: `-test_method kruskal | mannwhitneyu | wilcoxon` (Cluster.py:160,191) — best group, exact group means and the
    p-value of the reference's _output_kmers with scipy's test (wilcoxon: the pinned scipy 1.7.1 mode rules)."""
    from subphaser_b200 import engine
    for case in load("ranktest_rows.json"):
        sgs = sorted(case["groups"])
        n = case["n"]
        col_group = [0] * n
        for gi, sg in enumerate(sgs):
            for c in case["groups"][sg]:
                col_group[c] = gi
        X = np.array([r["array"] for r in case["rows"]], dtype=np.float64)
        best, pval, means, flags = engine.ranktest_groups(_dev(X, np.float64), col_group, len(sgs), case["method"])
        best, pval, means = best.cpu().numpy(), pval.cpu().numpy(), means.cpu().numpy()
        assert flags == 0
        for i, r in enumerate(case["rows"]):
            assert sgs[best[i]] == r["max_sg"]
            assert means[i].tolist() == r["mean_vals"]
            if math.isnan(r["pvalue"]):
                assert math.isnan(pval[i]), (case["method"], n, i)
            else:
                assert pval[i] == pytest.approx(r["pvalue"], rel=1e-9, abs=1e-300), (case["method"], n, i)


def test_ranktest_against_installed_scipy_and_errors(tmp_path):
    """The branches both scipy versions share, checked against the installed scipy directly, and the two conditions under
    which scipy raises (the drop-in raises the same ValueError from Cluster.output_kmers)."""
    from scipy import stats
    from subphaser_b200 import engine
    from subphaser_b200.Cluster import Cluster
    rng = np.random.default_rng(8)
    n = 14
    col_group = [0] * 7 + [1] * 7
    X = rng.random((400, n)) * 1e-5
    X[:, :7] += rng.random((400, 7)) * 3e-5
    dX = _dev(X, np.float64)
    for method, fn in (("kruskal", stats.kruskal), ("mannwhitneyu", stats.mannwhitneyu), ("wilcoxon", stats.wilcoxon)):
        best, pval, means, flags = engine.ranktest_groups(dX, col_group, 2, method)
        pval, best = pval.cpu().numpy(), best.cpu().numpy()
        for i in range(len(X)):
            a, b = (X[i, :7], X[i, 7:]) if best[i] == 0 else (X[i, 7:], X[i, :7])
            assert pval[i] == pytest.approx(float(fn(a, b).pvalue), rel=1e-9, abs=1e-300), (method, i)
    # wilcoxon on groups of different size, kruskal on identical values
    _, _, _, flags = engine.ranktest_groups(dX, [0] * 6 + [1] * 8, 2, "wilcoxon")
    assert flags & 1
    same = np.full((3, n), 2e-5)
    _, _, _, flags = engine.ranktest_groups(_dev(same, np.float64), col_group, 2, "kruskal")
    assert flags & 2
    # through the drop-in class
    mat = tmp_path / "m.kmer.mat"
    chrs = ["c%02d" % i for i in range(n)]
    with open(mat, "w") as f:
        f.write("kmer\t" + "\t".join(chrs) + "\n")
        for i in range(50):
            kmer = "".join("ACGT"[(i >> (2 * b)) & 3] for b in range(15))
            f.write(kmer + "\t" + "\t".join(repr(float(v)) for v in X[i]) + "\n")
    assigned = {c: ("SG1" if i < 6 else "SG2") for i, c in enumerate(chrs)}
    cl = Cluster(str(mat), n_clusters=2, sg_assigned=assigned, replicates=0)
    import io
    with pytest.raises(ValueError, match="same length"):
        cl.output_kmers(io.StringIO(), test_method="wilcoxon")
    d = cl.output_kmers(io.StringIO(), test_method="mannwhitneyu")
    assert len(d) > 0
    with pytest.raises(AttributeError):
        cl.output_kmers(io.StringIO(), test_method="no_such_test")


def _blobs(rng, n_per, S, M, sep=6.0):
    centers = rng.normal(0, sep, (S, M))
    X = np.concatenate([centers[s] + rng.normal(0, 1.0, (n_per, M)) for s in range(S)])
    perm = rng.permutation(len(X))
    return X[perm]


@pytest.mark.parametrize("n_per,S,M", [(3, 2, 500), (7, 3, 2000), (5, 4, 300), (12, 3, 64)])
def test_gram_kmeans_centroids_vs_sklearn(n_per, S, M):
    from oracle import restate
    from subphaser_b200 import engine
    rng = np.random.default_rng(n_per * 100 + S)
    pts = _blobs(rng, n_per, S, M)                     # [n, M]: n chromosomes in M dimensions
    n = len(pts)
    raw = np.ascontiguousarray(pts.T)                  # matrix layout [M, n]
    Z = engine.zscore_rows(_dev(raw, np.float64))
    Zh = Z.cpu().numpy()
    Gm = engine.gram(Z).cpu().numpy()
    np.testing.assert_allclose(Gm, Zh.T @ Zh, rtol=1e-12, atol=1e-9)
    chrs = ["c%02d" % i for i in range(n)]
    order = list(range(n))
    labels, inertia = engine.kmeans_gram(engine.gram(Z), S, order=order, seed=3)
    want, km = restate.kmeans_labels(Zh, S, chrs, seed=0)
    assert labels[0].cpu().numpy().tolist() == want
    assert inertia[0].item() == pytest.approx(km.inertia_, rel=1e-10)
    C = engine.centroids(Z, labels[0].cpu().numpy(), S).cpu().numpy()
    mine_to_sk = {}
    for i in range(n):
        mine_to_sk[want[i]] = km.labels_[i]
    for s in range(S):
        np.testing.assert_allclose(C[s], km.cluster_centers_[mine_to_sk[s]], rtol=0, atol=1e-10)


def test_bootstrap_scores_vs_sklearn():
    from oracle import restate
    from subphaser_b200 import engine
    rng = np.random.default_rng(9)
    pts = _blobs(rng, 7, 3, 1500, sep=1.2)              # modest separation: some replicates disagree
    n = len(pts)
    Z = engine.zscore_rows(_dev(np.ascontiguousarray(pts.T), np.float64))
    Zh = Z.cpu().numpy()
    chrs = ["c%02d" % i for i in range(n)]
    ref_labels, _ = restate.kmeans_labels(Zh, 3, chrs)
    R, B = 64, 200
    idx = rng.integers(0, Zh.shape[0], (R, B)).astype(np.int32)
    G = engine.gram_batched(Z, _dev(idx, np.int32))
    np.testing.assert_allclose(G[5].cpu().numpy(), Zh[idx[5]].T @ Zh[idx[5]], rtol=1e-12, atol=1e-9)
    labels, _ = engine.kmeans_gram(G, 3, order=list(range(n)), seed=1)
    want = restate.bootstrap_labels(Zh, 3, chrs, idx)
    got = labels.cpu().numpy()
    agree = np.mean([(g == w).all() for g, w in zip(got, want)])
    assert agree >= 0.95          # both sides are local-search heuristics with different RNG streams
    ari, vm = engine.cluster_scores(ref_labels, labels)
    for r in range(R):
        a, v = restate.cluster_scores(ref_labels, got[r])
        assert ari[r].item() == pytest.approx(a, abs=1e-12)
        assert vm[r].item() == pytest.approx(v, abs=1e-12)


def test_pca_vs_sklearn_full_svd():
    from oracle import restate
    from subphaser_b200 import engine
    rng = np.random.default_rng(2)
    pts = _blobs(rng, 7, 3, 4000)
    Z = engine.zscore_rows(_dev(np.ascontiguousarray(pts.T), np.float64))
    eig, scores, ratio = (t.cpu().numpy() for t in engine.pca_gram(engine.gram(Z), 3))
    X, want_ratio = restate.pca_scores(Z.cpu().numpy(), 3)
    np.testing.assert_allclose(ratio, want_ratio, rtol=0, atol=1e-10)
    for j in range(3):
        sgn = 1.0 if np.dot(scores[:, j], X[:, j]) >= 0 else -1.0
        np.testing.assert_allclose(sgn * scores[:, j], X[:, j], rtol=0, atol=1e-8)


def test_stack_lines_kernel_matches_reference_stack(tmp_path):
    """hotpath's device-side stacking (spk_stack_lines) == Circos.stack_matrix of the reference text."""
    import torch
    from subphaser_b200 import Seqs, _lib, engine
    for ci, case in enumerate(load("map_stack.json")):
        seq = case["seq"]
        fa = (">chrX\n" + seq + "\n").encode()
        d, n = engine.to_device_bytes(fa)
        ps = engine.pack_fasta(d, n)
        sig, k = Seqs._sig_table(case["d_kmers"], case["k"], case["sg_names"])
        S = len(case["sg_names"])
        chunk = int(case["window_size"]) if case["chunk"] else 0
        lines, nh = engine.map_bins(ps, sig, S, case["bin_size"], chunk)
        for ws, st in case["stacks"].items():
            ws_i = int(float(ws))
            L = ps.n_bases
            nwin = ((max(L - 1, 0) // case["bin_size"]) * case["bin_size"]) // ws_i + 1
            out = torch.zeros(nwin, S, dtype=torch.int64, device="cuda")
            _lib.call("spk_stack_lines", engine._p(lines), lines.shape[0], S, k, case["bin_size"], chunk, ws_i, L,
                      engine._p(out), nwin, engine._stream())
            got = out.cpu().numpy()
            nz = got.any(axis=1)
            assert [[int(x) for x in r] for r in got[nz]] == st["counts"]
            assert [int(w) * ws_i for w in np.nonzero(nz)[0]] == [c[1] for c in st["coords"]]


@pytest.mark.parametrize("k", [9, 17, 21, 29, 31])
def test_map_bins_vs_oracle_mapper(k):
    """K9 against oracle/kmer_count.c:orc_map_bins: 16-bit (k=9,17) and 32-bit (k=21) bucketed quotient tables,
    the open-addressed table for wide k, including k > 28 (separate value array)."""
    import torch
    import spk_testutil as util
    from oracle import kmers
    from subphaser_b200 import engine
    rng = np.random.default_rng(k)
    seq = util.messy_seq(rng, 60000, n_frac=0.01)
    fa = util.fasta([("c", seq)])
    codes, _ = kmers.fasta_to_codes(fa)
    keys_all, _, _ = kmers.count_fasta(fa, k, 1)
    pick = rng.choice(len(keys_all), size=min(4000, len(keys_all)), replace=False)
    keys = np.sort(keys_all[pick])
    sgs = rng.integers(0, 3, len(keys)).astype(np.uint8)
    for bin_size, chunk, qt_mean in ((1000, 7000, None), (10000, 0, "8"), (333, 50000, None)):
        L = len(codes)
        n_lines = (L - 1) // bin_size + ((L - 1 + k - 1) // chunk if chunk else 0) + 1
        want, hits = kmers.map_bins(codes, k, keys, sgs, 3, bin_size, chunk, n_lines)
        d, n = engine.to_device_bytes(fa)
        ps = engine.pack_fasta(d, n)
        if qt_mean:      # ~8 keys per 7-entry bucket: a third of the keys overflow into the stash
            os.environ["SPK_QT_MEAN"] = qt_mean
        try:
            sig = engine.SigTable(torch.from_numpy(keys.view(np.int64).copy()).cuda(), torch.from_numpy(sgs).cuda(), k)
            # without per-entry hit flags the genome-scale kernel runs (k_map_bins_w: warp-private, 16-bit slots)
            sig_w = engine.SigTable(torch.from_numpy(keys.view(np.int64).copy()).cuda(), torch.from_numpy(sgs).cuda(), k,
                                    track_hits=False)
        finally:
            os.environ.pop("SPK_QT_MEAN", None)
        got_w, nh_w = engine.map_bins(ps, sig_w, 3, bin_size, chunk)
        assert nh_w == hits
        np.testing.assert_array_equal(got_w.cpu().numpy().view(np.uint32), want)
        if qt_mean and k == 9:   # (wider k: the remainder width, not the load, fixes the bucket count)
            assert sig.bucket and int((sig.skeys != -1).sum().item()) > 0          # the stash is really in use
        got, nh = engine.map_bins(ps, sig, 3, bin_size, chunk)
        assert nh == hits
        np.testing.assert_array_equal(got.cpu().numpy().view(np.uint32), want)
        if sig.bucket:
            # distinct mapped k-mer STRINGS, the two orientations counted separately (len(mapped_cat), Seqs.py:109-113)
            keyset = set(keys.tolist())
            up = seq.upper()
            seen = set()
            for i in range(len(up) - k + 1):
                km = up[i:i + k]
                if km in seen or any(c not in "ACGT" for c in km):
                    continue
                if kmers.str_to_key(min(km, kmers.revcomp(km))) in keyset:
                    seen.add(km)
            assert sig.n_mapped() == len(seen)
        else:
            assert sig.n_mapped() <= len(keys)


def test_stack_bed_density_matches_reference_files(tmp_path):
    """f4: Circos.stack_bed_density (Circos.py:777-806) — per-subgenome circos density tracks, byte-identical files."""
    from subphaser_b200 import Circos
    for ci, case in enumerate(load("map_stack.json")):
        bc = tmp_path / ("bin%d.count" % ci)
        bc.write_text(case["bin_count_text"])
        for ws, want in case["density"].items():
            files = Circos.stack_bed_density(str(bc), str(tmp_path / ("d%d_%s" % (ci, ws))), case["sg_names"],
                                             window_size=int(ws))
            assert sorted(files) == sorted(want)
            for key, path in files.items():
                assert open(path).read() == want[key], (ci, ws, key)


def test_device_text_writers_match_python_repr():
    """f2: spk_format_rows — `.kmer.mat` and `.sig.kmer-subgenome.tsv` rows formatted on the device are byte-identical to
    Python's str() of the same values (shortest round-trip repr, exponent / fixed notation rules, nan, inf, -0.0)."""
    import torch
    from subphaser_b200 import engine, kmer_codec
    rng = np.random.default_rng(5)
    M, n, k = 5000, 7, 17
    vals = rng.integers(0, 2**63, (M, n), dtype=np.uint64).view(np.float64).copy()
    vals[:1000] = rng.integers(0, 5000, (1000, n)) / rng.integers(1, 800_000_000, (1000, n))      # count / length
    vals[1000:1200] = rng.integers(0, 10**9, (200, n)).astype(np.float64)
    vals[1200, :] = [0.0, -0.0, 1e16, 9999999999999998.0, 1e-4, 9.999e-5, 5e-324]
    vals[1201, :] = [np.inf, -np.inf, np.nan, 1e22, 1e23, 0.1, 2.0 ** -44]
    keys = rng.integers(0, 4 ** k, M, dtype=np.uint64)
    d_keys = torch.from_numpy(keys.view(np.int64).copy()).cuda()
    d_vals = torch.from_numpy(vals).cuda()
    strs = kmer_codec.keys_to_strs(keys, k)
    got = engine.format_rows(d_keys, d_vals, k, kind=0).tobytes().decode()
    want = "".join(s + "\t" + "\t".join(map(repr, r)) + "\n" for s, r in zip(strs, vals.tolist()))
    assert got == want
    # kind 1 on a row subset
    rows = np.sort(rng.choice(M, 700, replace=False)).astype(np.int32)
    label = rng.integers(0, 3, M).astype(np.int32)
    pval = rng.random(M)
    pval[rows[:5]] = [np.nan, 0.0, 1.0, 1e-300, 3e-5]
    names = ["SG1", "SG2", "SG03_long_name"]
    got = engine.format_rows(d_keys, d_vals, k, kind=1, rows=torch.from_numpy(rows).cuda(),
                             label=torch.from_numpy(label).cuda(), label_names=names,
                             pval=torch.from_numpy(pval).cuda()).tobytes().decode()
    want = "".join("{}\t{}\t{}\t{}\n".format(strs[i], names[label[i]], repr(float(pval[i])),
                                            ",".join(map(repr, vals[i].tolist()))) for i in rows.tolist())
    assert got == want


def test_fold_histogram_matches_numpy():
    """f3: the histogram of the fold-passing totals (Jellyfish.py:499-511,650-666) binned on the device equals
    np.histogram / np.percentile of the materialised list."""
    import torch
    from subphaser_b200 import engine
    rng = np.random.default_rng(12)
    for n, hi in ((200_000, 5000), (50_000, 3_000_000_000), (1000, 60), (3, 10)):
        tot = rng.integers(1, hi, n).astype(np.uint64)
        tot[: n // 3] = rng.integers(200, 260, n // 3)              # a dense peak: ties across bin edges
        flags = (rng.random(n) < 0.7).astype(np.uint8) | (rng.integers(0, 2, n).astype(np.uint8) << 1)
        flags[0] |= 1
        data = tot[(flags & 1) == 1].astype(np.float64)
        fh = engine.FoldHistogram(torch.from_numpy(tot.view(np.int64)).cuda(), torch.from_numpy(flags).cuda(), n)
        nbins = max(int(int(data.max()) / 25), 1)
        want, edges = np.histogram(data, bins=nbins)
        assert fh.n == len(data) and fh.nbins == nbins
        np.testing.assert_array_equal(fh.hist, want)
        np.testing.assert_array_equal(fh.edges, edges)
        assert fh.xlim == float(np.percentile(data, 99))
        assert fh.percentile(50) == float(np.percentile(data, 50)) and fh.percentile(0) == data.min()
