"""GPU: the drop-in modules driven in the reference's step order reproduce, byte for byte where the
format allows, the files the REFERENCE's own code produced for the same genome
(tests/golden/pipeline_small, made by tests/golden/make_golden.py)."""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# pipeline_small: 2 x 3 chromosomes, single-chromosome groups; pipeline_arab: the layout of
# example_data/Arabidopsis_suecica_sg.config (13 chromosomes, 5 + 8 per subgenome, groups of one to three chromosomes)
@pytest.fixture(scope="module", params=["pipeline_small", "pipeline_arab"])
def run(request, tmp_path_factory):
    from subphaser_b200 import _registry, pipeline
    G = os.path.join(GOLDEN, request.param)
    _registry.clear()
    meta = json.load(open(os.path.join(G, "meta.json")))
    work = tmp_path_factory.mktemp("pipe")
    chromfiles = []
    for lab in meta["labels"]:
        dst = os.path.join(work, lab + ".fasta")
        with open(os.path.join(G, lab + ".fasta"), "rb") as fi, open(dst, "wb") as fo:
            fo.write(fi.read())
        chromfiles.append(dst)
    idx = np.load(os.path.join(G, "resample_idx.npz"))["idx"]
    res = pipeline.run_hot_path(chromfiles, meta["labels"], meta["sgs"], os.path.join(work, "out"), k=meta["k"],
                                lower_count=meta["lower_count"], min_freq=meta["min_freq"], nsg=meta["nsg"],
                                replicates=meta["replicates"], bin_size=meta["bin_size"],
                                map_window=meta["map_window"], window_size=meta["enrich_window"],
                                resample_idx=idx, seed=0)
    return G, meta, res


def _rows(path):
    return [l.rstrip("\n").split("\t") for l in open(path)]


def test_counts_and_matrix(run):
    G, meta, res = run
    assert res["lengths"] == meta["lengths"]
    assert res["kmer_count"] == meta["n_union"]
    assert res["n_diff"] == meta["n_diff"]
    ref = _rows(os.path.join(G, "ref.kmer.mat"))
    got = _rows(res["matfile"])
    assert got[0] == ref[0]
    # row order of the reference is jellyfish hash order (arbitrary): compare as sorted sets, exact text
    assert sorted(map(tuple, got[1:])) == sorted(map(tuple, ref[1:]))


def test_cluster_assignment_and_bootstrap(run):
    G, meta, res = run
    assert dict(res["d_sg"]) == meta["d_sg"]
    assert [int(x) for x in res["cluster"].labels] == meta["labels_full"]
    assert res["cluster"].d_bs == meta["d_bs"]
    assert res["cluster"].mean_adjusted_rand_score == pytest.approx(meta["mean_ari"], abs=1e-12)
    assert res["cluster"].mean_v_measure_score == pytest.approx(meta["mean_vm"], abs=1e-12)
    ref = open(os.path.join(G, "ref.chrom-subgenome.tsv")).read()
    assert open(res["para_prefix"] + ".chrom-subgenome.tsv").read() == ref
    assert res["cluster"].kmean.inertia_ == pytest.approx(meta["inertia"], rel=1e-10)


def test_centroids_and_pca(run):
    G, meta, res = run
    c = res["cluster"]
    # rows of the fixture follow sklearn's cluster numbering; match clusters through the labels
    ref_lab, my_lab = meta["centers_labels"], list(c.kmean.labels_)
    centers = c.kmean.cluster_centers_
    # the fixture's matrix rows are in reference (hash) order; ours are sorted by k-mer: align by k-mer
    ref_kmers = [r[0] for r in _rows(os.path.join(G, "ref.kmer.mat"))[1:]]
    pos = {km: i for i, km in enumerate(c.kmers)}
    cols = [pos[km] for km in ref_kmers[:50]]
    for chrom_i in range(len(ref_lab)):
        np.testing.assert_allclose(centers[my_lab[chrom_i]][cols], meta["centers_head"][ref_lab[chrom_i]],
                                   rtol=0, atol=1e-10)
    ref_scores = np.array(meta["pca_scores"])
    eig, scores, ratio = (t.cpu().numpy() for t in __import__("subphaser_b200").engine.pca_gram(c._G, meta["nsg"]))
    np.testing.assert_allclose(ratio, meta["pca_ratio"], rtol=0, atol=1e-10)
    for j in range(meta["nsg"]):
        sgn = 1.0 if np.dot(scores[:, j], ref_scores[:, j]) >= 0 else -1.0
        np.testing.assert_allclose(sgn * scores[:, j], ref_scores[:, j], rtol=0, atol=1e-8)


def test_specific_kmers(run):
    G, meta, res = run
    assert len(res["d_kmers"]) == meta["n_sig"]
    ref = {r[0]: r for r in _rows(os.path.join(G, "ref.sig.kmer-subgenome.tsv"))[1:]}
    got = {r[0]: r for r in _rows(res["para_prefix"] + ".sig.kmer-subgenome.tsv")[1:]}
    assert set(ref) == set(got)
    for km, r in ref.items():
        g = got[km]
        assert g[1] == r[1]
        pr, pg = float(r[2]), float(g[2])
        if math.isnan(pr):
            assert math.isnan(pg)
        else:
            assert pg == pytest.approx(pr, rel=1e-9, abs=1e-300)
        assert g[3] == r[3]                      # group means: exact text
    # mapping behaviour of the returned d_kmers
    from subphaser_b200 import kmer_codec
    km = next(iter(ref))
    assert res["d_kmers"][km] == ref[km][1]
    assert res["d_kmers"][kmer_codec.revcomp_str(km)] == ref[km][1]
    with pytest.raises(KeyError):
        res["d_kmers"]["A" * meta["k"] if "A" * meta["k"] not in ref else "C" * meta["k"]]


def test_bin_count_file_identical(run):
    G, meta, res = run
    ref = open(os.path.join(G, "ref.subgenome.bin.count")).read()
    assert open(res["para_prefix"] + ".subgenome.bin.count").read() == ref


def test_enrichment_files(run):
    G, meta, res = run
    ref = _rows(os.path.join(G, "ref.bin.enrich"))
    got = _rows(res["para_prefix"] + ".bin.enrich")
    assert len(ref) == len(got) and got[0] == ref[0]
    for r, g in zip(ref[1:], got[1:]):
        assert g[:4] == r[:4]                                   # chrom start end subgenome
        assert g[5] == r[5] and g[7] == r[7] and g[9] == r[9]   # counts, enrich one-hot, exchange
        assert g[6] == r[6]                                     # ratios: exact text
        for a, b in zip([g[4]] + g[8].split(",") + [g[10]], [r[4]] + r[8].split(",") + [r[10]]):
            assert float(a) == pytest.approx(float(b), abs=1e-10, rel=1e-9)
    assert open(res["para_prefix"] + ".bin.group").read() == open(os.path.join(G, "ref.bin.group")).read()
