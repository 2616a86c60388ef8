"""CPU: pin oracle/restate.py to the fixtures produced by the reference's own code
(tests/golden/make_golden.py).  Bit-exact for integers and for the fp64 values the reference computes
with plain Python/numpy arithmetic; scipy-version tolerance where a library call is involved."""
import json
import math
import os

import numpy as np
import pytest

from oracle import kmers, restate

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_filter_kmer_matches_reference():
    for case in load("filter_kmer.json"):
        p = case["params"]
        for row in case["rows"]:
            freqs, tot = restate.filter_kmer(row["counts"], case["lengths"], case["labels"], case["sgs"], **p)
            assert tot == row["tot"]
            assert freqs == row["freqs"]          # exact float equality (IEEE division)


def test_fisher_enrich_matches_reference():
    for case in load("fisher_enrich.json"):
        pmin = []
        for row in case["rows"]:
            res = restate.enrich_row(row["row"], case["total"])
            np.testing.assert_allclose(res["pvals"], row["pvals"], rtol=1e-12, atol=0)
            assert res["idx"] == row["idx"] and res["sig"] == row["sig"]
            assert res["enrich"] == row["enrich"]
            np.testing.assert_array_equal(np.nan_to_num(res["ratios"], nan=-1), np.nan_to_num(row["ratios"], nan=-1))
            pmin.append(row["pvals"][row["idx"]])
        np.testing.assert_array_equal(restate.bh(pmin), case["qvals"])
        from scipy.stats import false_discovery_control
        np.testing.assert_allclose(restate.bh(pmin), false_discovery_control(pmin, method="bh"), rtol=1e-12)


def test_map_and_stack_match_reference():
    for case in load("map_stack.json"):
        lines = restate.map_kmer_lines("chrX", case["seq"], case["d_kmers"], case["k"], case["bin_size"],
                                       case["sg_names"], chunk=case["chunk"], window_size=case["window_size"])
        ref_lines = case["bin_count_text"].splitlines(keepends=True)
        assert lines == ref_lines[1:]
        for ws, st in case["stacks"].items():
            coords, counts = restate.stack_matrix(ref_lines, float(ws) if "." in ws else int(ws))
            assert [list(c) for c in coords] == st["coords"]
            assert counts == st["counts"]


def test_c_map_bins_matches_reference_lines():
    """oracle/kmer_count.c:orc_map_bins (the CPU-baseline mapper) against the same fixture."""
    from subphaser_b200 import kmer_codec
    for case in load("map_stack.json"):
        k, S = case["k"], len(case["sg_names"])
        fa = (">chrX\n" + case["seq"] + "\n").encode()
        codes, _ = kmers.fasta_to_codes(fa)
        strs = sorted(s for s in case["d_kmers"] if s <= kmers.revcomp(s))
        keys = np.array([kmers.str_to_key(s) for s in strs], dtype=np.uint64)
        sg = np.array([case["sg_names"].index(case["d_kmers"][s]) for s in strs], dtype=np.uint8)
        order = np.argsort(keys)
        chunk = int(case["window_size"]) if case["chunk"] else 0
        L = len(codes)
        n_lines = (L - 1) // case["bin_size"] + ((L - 1 + k - 1) // chunk if chunk else 0) + 1
        counts, hits = kmers.map_bins(codes, k, keys[order], sg[order], S, case["bin_size"], chunk, n_lines)
        ref = [l.split("\t") for l in case["bin_count_text"].splitlines()[1:]]
        got = counts[counts.any(axis=1)]
        want = np.array([[int(x) for x in l[3:]] for l in ref], dtype=np.uint32).reshape(-1, S)
        np.testing.assert_array_equal(got, want)
        assert hits == int(want.sum())


def test_ttest_rows_match_reference():
    from collections import OrderedDict
    for case in load("ttest_rows.json"):
        groups = OrderedDict(case["groups"])
        for row in case["rows"]:
            with np.errstate(all="ignore"):
                sg, p, means = restate.output_kmer(row["array"], groups)
            assert sg == row["max_sg"]
            assert [float(m) for m in means] == row["mean_vals"]
            if math.isnan(row["pvalue"]):
                assert math.isnan(p)
            else:
                assert p == pytest.approx(row["pvalue"], rel=1e-12, abs=0)


def test_pipeline_fixture_consistent_with_oracle_counter():
    """lengths / union size of the reference run are reproduced by the oracle counter + restate.to_matrix."""
    d = os.path.join(G, "pipeline_small")
    meta = json.load(open(os.path.join(d, "meta.json")))
    dumps = []
    for lab in meta["labels"]:
        keys, counts, _ = kmers.count_fasta(open(os.path.join(d, lab + ".fasta"), "rb").read(), meta["k"], meta["lower_count"])
        dumps.append((keys, counts))
    allk, mat, lengths = restate.to_matrix(dumps)
    assert lengths == meta["lengths"] and len(allk) == meta["n_union"]
    keys, norm, tot, n_fold = restate.filter_matrix(allk, mat, lengths, meta["labels"], meta["sgs"],
                                                    min_freq=meta["min_freq"], max_freq=10000, min_fold=2,
                                                    baseline=1, ratio=1)
    assert len(keys) == meta["n_diff"]
    ref = {}
    for line in open(os.path.join(d, "ref.kmer.mat")).read().splitlines()[1:]:
        t = line.split("\t")
        ref[t[0]] = [float(x) for x in t[1:]]
    got = {kmers.key_to_str(a, meta["k"]): list(r) for a, r in zip(keys, norm)}
    assert got == ref


def test_zscore_and_numpy_pairwise_model():
    """restate.zscore is the reference expression; also check the pairwise-summation model the CUDA
    kernels implement (spk_cluster.cu:np_pairwise) against numpy itself, bit for bit."""
    def np_pairwise(a):
        n = len(a)
        if n < 8:
            s = 0.0
            for x in a:
                s += x
            return s
        if n <= 128:
            r = [a[j] for j in range(8)]
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] += a[i + j]
                i += 8
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
            while i < n:
                res += a[i]
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return np_pairwise(a[:n2]) + np_pairwise(a[n2:])

    rng = np.random.default_rng(0)
    for n in (2, 3, 6, 7, 8, 9, 13, 20, 21, 38, 64, 127, 128):
        raw = rng.random((50, n)) * 1e-4
        Z = restate.zscore(raw)
        for m in range(50):
            x = [float(v) for v in raw[m]]
            mean = np_pairwise(x) / n
            sd = math.sqrt(np_pairwise([(v - mean) * (v - mean) for v in x]) / n)
            mine = [(v - mean) / sd for v in x]
            assert mine == [float(v) for v in Z[m]], n
