"""GPU parity, SURVEY 8(f) rank 2: the per-resolution cluster diagnostics of makeGraphsAndClusters
(bluster::approxSilhouette, bluster::pairwiseModularity) against the oracle restatement."""
import numpy as np
import pytest

import cellula_b200
from cellula_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def labels_for(n, k, seed, emb=None):
    rng = np.random.default_rng(seed)
    if emb is None:
        return rng.integers(0, k, n)
    cent = emb[rng.choice(n, k, replace=False)]
    return ((emb[:, None, :] - cent[None]) ** 2).sum(-1).argmin(1)     # nearest of k random cells


@pytest.mark.parametrize("n,d,k", [(3000, 20, 7), (5000, 50, 40), (400, 3, 2), (1000, 10, 1)])
def test_approx_silhouette_matches_oracle(gpu_ctx, n, d, k):
    emb = synth.make_embedding(n, d, seed=n + k)
    cl = labels_for(n, k, 1, emb) * 3 + 10                 # arbitrary label values, as the mirror maps them
    got = cellula_b200.approxSilhouette(emb, cl)
    other, width, _ = O.approx_silhouette(emb, cl)
    if k == 1:
        assert np.all(np.isnan(got["width"])) and np.all(np.isnan(width))       # (Inf - s) / Inf in R
        return
    assert np.array_equal(got["other"], other)
    assert np.allclose(got["width"], width, rtol=1e-10, atol=1e-12)
    assert np.all(got["width"] <= 1) and np.all(got["width"] >= -1)


def test_approx_silhouette_resident_scores_and_string_labels(gpu_ctx):
    emb = synth.make_embedding(2000, 12, seed=5)
    cl = np.array(["b", "a", "c", "a"])[labels_for(2000, 4, 2)]
    got = cellula_b200.approxSilhouette(emb, cl)
    other, width, _ = O.approx_silhouette(emb, cl)
    assert np.array_equal(got["other"], other) and np.allclose(got["width"], width, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("scheme", ["rank", "jaccard"])
@pytest.mark.parametrize("k", [3, 25, 80])
def test_pairwise_modularity_matches_oracle(gpu_ctx, scheme, k):
    emb = synth.make_embedding(6000, 15, seed=9)
    g = cellula_b200.makeSNNGraph(emb, k=12, type=scheme)
    cl = labels_for(6000, k, 3, emb)
    want = O.pairwise_modularity(g.frm, g.to, g.weight, cl, get_weights=True)
    got = cellula_b200.pairwiseModularity(g, cl, get_weights=True)
    assert np.allclose(got["observed"], want["observed"], rtol=1e-11, atol=1e-9)
    assert np.allclose(got["expected"], want["expected"], rtol=1e-11, atol=1e-9)
    assert np.all(np.tril(got["observed"], -1) == 0) and np.all(np.tril(got["expected"], -1) == 0)
    ratio = cellula_b200.pairwiseModularity(g, cl, as_ratio=True)
    wr = O.pairwise_modularity(g.frm, g.to, g.weight, cl, as_ratio=True)
    iu = np.triu_indices(ratio.shape[0])
    assert np.allclose(ratio[iu], wr[iu], rtol=1e-10, equal_nan=True)
    # the whole weight is accounted for, and the total modularity is the sum of the entries
    assert np.isclose(got["observed"].sum(), g.weight.sum(), rtol=1e-12)
    assert np.isclose(got["expected"].sum(), g.weight.sum(), rtol=1e-12)


def test_pairwise_modularity_resident_edges(gpu_ctx):
    emb = synth.make_embedding(4000, 10, seed=2)
    g = cellula_b200.makeSNNGraph(emb, k=10, type="rank")          # leaves the edges in HBM
    cl = labels_for(4000, 5, 7, emb)
    uclust = np.unique(cl)
    obs, expd, tot = cellula_b200.get_context().pairwise_modularity(None, np.searchsorted(uclust, cl).astype(np.int32),
                                                                    uclust.size)
    want = O.pairwise_modularity(g.frm, g.to, g.weight, cl, get_weights=True)
    assert np.allclose(obs, want["observed"], rtol=1e-11) and np.allclose(expd, want["expected"], rtol=1e-11)
    assert np.isclose(tot, g.weight.sum(), rtol=1e-12)


def test_diagnostics_argument_errors(gpu_ctx):
    emb = synth.make_embedding(100, 4, seed=1)
    with pytest.raises(cellula_b200.Cb200Error, match="out of"):
        gpu_ctx.approx_silhouette(emb, np.full(100, 5, dtype=np.int32), 3)
    with pytest.raises(cellula_b200.Cb200Error, match="out of range"):
        gpu_ctx.pairwise_modularity((np.array([1, 200]), np.array([2, 3]), np.array([1.0, 1.0])),
                                    np.zeros(100, dtype=np.int32), 1)
