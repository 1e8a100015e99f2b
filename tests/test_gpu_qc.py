"""GPU parity, SURVEY 8(f) rank 4: scuttle::perCellQCMetrics as doQC calls it (R/doQC.R:106-108) and the
neighbour distances that let uwot::umap re-use the exact kNN search (R/doNormAndReduce.R:102-106)."""
import numpy as np
import pytest

import cellula_b200
from cellula_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_subsets", [0, 1, 3, 6])
def test_qc_metrics_bit_exact(gpu_ctx, n_subsets):
    d = synth.make_counts(3000, 5000, nnz_per_cell=250, seed=21)
    rng = np.random.default_rng(n_subsets)
    subsets = [np.sort(rng.choice(5000, size, replace=False)) for size in (13, 1, 400, 0, 77, 5000)[:n_subsets]]
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    got = gpu_ctx.qc_metrics(subsets)
    want = O.per_cell_qc_metrics(d["n_genes"], d["colptr"], d["rowidx"], d["x"], subsets)
    assert np.array_equal(got["sum"], want["sum"]) and np.array_equal(got["detected"], want["detected"])
    assert len(got["subsets"]) == n_subsets
    for g, w in zip(got["subsets"], want["subsets"]):   # integer counts: sums are exact, so is the percentage
        assert np.array_equal(g["sum"], w["sum"]) and np.array_equal(g["detected"], w["detected"])
        assert np.array_equal(g["percent"], w["percent"])


def test_qc_metrics_threshold_and_bad_subset(gpu_ctx):
    d = synth.make_counts(500, 800, nnz_per_cell=100, seed=22)
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    got = gpu_ctx.qc_metrics([np.arange(50)], threshold=2.0)
    want = O.per_cell_qc_metrics(d["n_genes"], d["colptr"], d["rowidx"], d["x"], [np.arange(50)], threshold=2.0)
    assert np.array_equal(got["detected"], want["detected"])
    assert np.array_equal(got["subsets"][0]["detected"], want["subsets"][0]["detected"])
    with pytest.raises(cellula_b200.Cb200Error, match="outside"):
        gpu_ctx.qc_metrics([np.array([800])])


@pytest.mark.parametrize("n,d,k", [(4000, 20, 15), (1500, 50, 200), (700, 3, 26)])
def test_knn_distances_match_the_search_arithmetic(gpu_ctx, n, d, k):
    emb = synth.make_embedding(n, d, seed=n)
    idx = gpu_ctx.knn(emb, k)
    dist = gpu_ctx.knn_distances()
    o_idx, o_dist = O.find_knn(emb, k, want_dist=True)
    assert np.array_equal(idx, o_idx)
    assert np.array_equal(dist, O.knn_distances(emb, idx))
    assert np.allclose(dist, o_dist, rtol=1e-14, atol=0)
    assert np.all(np.diff(dist, axis=1) >= 0)


def test_knn_distances_of_a_query_range(gpu_ctx):
    emb = synth.make_embedding(3000, 10, seed=9)
    idx = gpu_ctx.knn(emb, 12, q_begin=1000, q_end=1700)
    dist = gpu_ctx.knn_distances()
    full = O.knn_distances(emb, O.find_knn(emb, 12))
    assert dist.shape == (700, 12) and np.array_equal(dist, full[1000:1700])
