"""Parity at the BASELINE.json config shapes:

  cfg1  2 700 x 32 738, 30 PCs, k = 10, jaccard  -- exact shape, every stage against the full oracle
  cfg2  68 579 x 32 738, 50 PCs, k = 20, rank    -- normalisation / gene statistics against the C oracle, PCA
        against an fp64 eigen-decomposition, ALL kNN rows and the COMPLETE SNN edge list bit-exact
  cfg4  fixed 1 000 000 x 50 embedding, k = 10 / 30 / 100 x rank / jaccard -- 2 048 sampled query rows against
        brute force, and the edges of a j-range against the oracle SNN run on the GPU's kNN result

Tolerances are BASELINE.json's north_star: size factors / logcounts 1e-6 relative (asserted much tighter),
identical HVG ranks, PCA subspace angle <= 1e-4 and sign-aligned scores <= 1e-4 relative, kNN / SNN bit-exact.
Everything goes through the C ABI (ctypes); the oracle is the checker only."""
import numpy as np
import pytest

import cellula_b200 as cb
from cellula_b200 import synth, trendvar
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def subspace_angle(a, b):
    """Largest principal angle between the column spaces of a and b (orthonormalised first)."""
    qa, _ = np.linalg.qr(a)
    qb, _ = np.linalg.qr(b)
    s = np.linalg.svd(qa.T @ qb, compute_uv=False)
    return float(np.arccos(np.clip(s.min(), -1.0, 1.0)))


def aligned_score_error(got, want):
    sign = np.sign(np.sum(got * want, axis=0))
    return float((np.abs(got * sign - want).max(axis=0) / np.abs(want).max(axis=0)).max())


@pytest.mark.timeout(900)
def test_cfg1_exact_shape_end_to_end():
    n, g, nd, k = 2700, 32738, 30, 10
    d = synth.make_counts(n, g, nnz_per_cell=850.0, seed=synth.SEED)
    ref = O.pipeline(g, d["colptr"], d["rowidx"], d["x"], ndims=nd, hvg_ntop=2000, neighbors=k, snn_type="jaccard")
    cb.options["skip_umap"] = True
    cb.options["size_factors"] = "library"
    names = [f"g{i}" for i in range(g)]
    sce = cb.SingleCellExperiment.from_counts(d["colptr"], d["rowidx"], d["x"], g, rownames=names)
    sce = cb.doNormAndReduce(sce, ndims=nd, hvg_ntop=2000, verbose=False)
    assert np.allclose(sce.colData["sizeFactor"], ref["sf"], rtol=1e-13, atol=0)
    assert np.allclose(sce.assays["logcounts"].x, ref["logx"], rtol=1e-12, atol=0)
    assert np.allclose(sce.rowData["mean"], ref["mean"], rtol=1e-11, atol=1e-300)
    assert np.allclose(sce.rowData["total"], ref["var"], rtol=1e-9, atol=1e-300)
    assert sce.metadata["hvgs"] == [names[i] for i in ref["hvgs"]]            # identical HVGs, identical order
    pca = sce.reducedDims["PCA"]
    assert subspace_angle(pca.attrs["rotation"], ref["pca"]["rotation"]) <= 1e-4
    assert aligned_score_error(np.asarray(pca), ref["pca"]["scores"]) <= 1e-4
    assert np.allclose(pca.attrs["percentVar"], ref["pca"]["percent_var"], rtol=1e-6)

    # the reference's own PCA matrix fed in: kNN and SNN bit-exact (north_star)
    ctx = cb.get_context()
    knn = ctx.knn(ref["pca"]["scores"], k)
    assert np.array_equal(knn, ref["knn"])
    for scheme in ("jaccard", "rank", "number"):
        f, t, w = ctx.snn(knn, scheme)
        wf, wt, ww = ref["edges"] if scheme == "jaccard" else O.snn_edges(ref["knn"], scheme)
        assert np.array_equal(f, wf) and np.array_equal(t, wt) and np.array_equal(w, ww), scheme

    # and through the reference-facing call on the drop-in's own PCA
    sce = cb.makeGraphsAndClusters(sce, neighbors=k, weighting_scheme="jaccard", ndims=nd, save_graphs=True)
    graph = sce.metadata["SNN_10_jaccard"]
    own_knn = O.find_knn(np.asarray(pca)[:, :nd], k)
    wf, wt, ww = O.snn_edges(own_knn, "jaccard")
    assert graph.n == n and np.array_equal(graph.frm, wf) and np.array_equal(graph.to, wt)
    assert np.array_equal(graph.weight, ww)


@pytest.mark.timeout(2400)
def test_cfg2_full_knn_snn_and_statistics(gpu_ctx):
    n, g, nd, k = 68579, 32738, 50, 20
    gpu_ctx.synth_counts(g, n, 525.0, seed=synth.SEED)
    colptr, rowidx, x = gpu_ctx.export_csc()
    o_sf, o_logx, o_mean, o_var = O.c_lognorm_genestats(g, colptr, rowidx, x)
    sf = gpu_ctx.size_factors()
    logx, mean, var = gpu_ctx.lognorm_genestats()
    assert np.allclose(sf, o_sf, rtol=1e-13, atol=0)
    assert np.allclose(logx, o_logx, rtol=1e-12, atol=0)
    assert np.allclose(mean, o_mean, rtol=1e-11, atol=1e-300)
    assert np.allclose(var, o_var, rtol=1e-9, atol=1e-300)
    hv = trendvar.get_top_hvgs(trendvar.model_gene_var(mean, var), 2000)
    assert np.array_equal(hv, O.get_top_hvgs(O.model_gene_var(o_mean, o_var)["bio"], 2000))

    pca = gpu_ctx.pca(hv, nd)
    import scipy.sparse as sp
    m = sp.csc_matrix((o_logx, rowidx.astype(np.int64), colptr), shape=(g, n)).tocsr()[hv.astype(np.int64)].T
    m = np.asarray(m.todense())
    mu = m.mean(axis=0)
    m -= mu
    evals, evecs = np.linalg.eigh(m.T @ m)
    order = np.argsort(-evals)[:nd]
    rot = evecs[:, order]
    assert subspace_angle(pca["rotation"], rot) <= 1e-4
    assert aligned_score_error(pca["scores"], m @ rot) <= 1e-4
    assert np.allclose(pca["var_explained"], evals[order] / (n - 1), rtol=1e-6)   # measured 1e-8: the fixed-point digits of the cross-product
    del m

    scores = pca["scores"]
    knn = gpu_ctx.knn(scores, k)
    want = O.find_knn(scores, k)                                  # all 68 579 rows, brute force, OpenMP
    bad = np.flatnonzero((knn != want).any(axis=1))
    assert bad.size == 0, (bad.size, bad[:10])
    assert np.array_equal(gpu_ctx.knn(None, k, n_total=n, d=nd), want)   # the embedding left in HBM: same answer
    for scheme in ("rank", "jaccard"):
        f, t, w = gpu_ctx.snn(knn, scheme)
        wf, wt, ww = O.snn_edges(want, scheme)                     # complete edge list, emission order
        assert f.size == wf.size
        assert np.array_equal(f, wf) and np.array_equal(t, wt) and np.array_equal(w, ww), scheme


@pytest.fixture(scope="module")
def cfg4():
    n, d = 1000000, 50
    emb = synth.make_embedding(n, d, seed=synth.EMB_SEED)
    queries = np.sort(np.random.default_rng(4).choice(n, 2048, replace=False))
    brute = O.find_knn_queries(emb, queries, 100)               # prefixes answer every smaller k
    return dict(n=n, d=d, emb=emb, queries=queries, brute=brute)


@pytest.mark.timeout(2400)
@pytest.mark.parametrize("k", [10, 30, 100])
def test_cfg4_knn_snn_sweep(gpu_ctx, cfg4, k):
    n, emb, q = cfg4["n"], cfg4["emb"], cfg4["queries"]
    knn = gpu_ctx.knn(emb, k)
    assert knn.shape == (n, k)
    bad = np.flatnonzero((knn[q] != cfg4["brute"][:, :k]).any(axis=1))
    assert bad.size == 0, (k, bad.size, q[bad[:5]])
    j0, j1 = 600000, 600000 + (4000 if k == 100 else 20000)
    for scheme in ("rank", "jaccard"):
        f, t, w = gpu_ctx.snn(knn, scheme, j_begin=j0, j_end=j1)
        wf, wt, ww = O.snn_edges(knn, scheme, j_begin=j0, j_end=j1)
        assert f.size == wf.size and f.size > (j1 - j0) * k
        assert np.array_equal(f, wf) and np.array_equal(t, wt) and np.array_equal(w, ww), (k, scheme)
