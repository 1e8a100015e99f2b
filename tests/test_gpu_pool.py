"""GPU parity, SURVEY 8(f) rank 1: deconvolution size factors (scran::computeSumFactors, R/doNormAndReduce.R:66-71)
against the oracle restatement (dense per-cluster arrays, numpy medians, lstsq on the explicit design matrix --
structurally independent of the GPU's sliding fixed-point windows, bucket selection and circulant DFT solve).
Contract (north_star): size factors within 1e-6 relative."""
import numpy as np
import pytest
import scipy.sparse as sp

from cellula_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def planted(n_genes, n_cells, seed, depth=2.0, groups=1):
    """Poisson counts with planted per-cell scale factors and `groups` expression programmes."""
    rng = np.random.default_rng(seed)
    lam = rng.gamma(2.0, depth, size=(groups, n_genes))
    grp = rng.integers(0, groups, n_cells)
    s = np.exp(rng.normal(0, 0.4, size=n_cells))
    x = rng.poisson(lam[grp].T * s[None, :]).astype(np.float64)
    x[0, x.sum(axis=0) == 0] = 1.0
    m = sp.csc_matrix(x)
    return dict(n_genes=n_genes, n_cells=n_cells, colptr=m.indptr.astype(np.int64), rowidx=m.indices.astype(np.int32),
                x=m.data.astype(np.float64)), s, grp


def run_both(gpu_ctx, d, clusters, **kw):
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    got, nneg = gpu_ctx.pooled_size_factors(clusters, **kw)
    want = O.pooled_size_factors(d["n_genes"], d["colptr"], d["rowidx"], d["x"], clusters, **kw)
    return got, nneg, want


def test_single_cluster_recovers_planted_factors(gpu_ctx):
    d, s, _ = planted(700, 320, seed=1)
    got, nneg, want = run_both(gpu_ctx, d, None)
    assert nneg == int((want <= 0).sum()) == 0
    assert np.allclose(got, want, rtol=RTOL, atol=0)
    assert np.corrcoef(got, s)[0, 1] > 0.99
    assert abs(got.mean() - 1.0) < 1e-12


def test_clusters_of_unequal_size_and_truncated_pool_sizes(gpu_ctx):
    # 3 expression programmes as clusters: 60 cells (only the pool sizes <= 60 apply), ~150 and ~290 cells
    d, s, grp = planted(901, 500, seed=2, groups=3)
    order = np.argsort(np.bincount(grp, minlength=3))
    cl = np.empty(500, dtype=np.int32)
    small = np.flatnonzero(grp == order[0])[:60]
    cl[:] = 1 + (grp == order[2])
    cl[small] = 0
    got, nneg, want = run_both(gpu_ctx, d, cl)
    assert np.allclose(got, want, rtol=RTOL, atol=0)
    assert nneg == int((want <= 0).sum())


def test_cluster_size_limit_splits_round_robin(gpu_ctx):
    d, _, _ = planted(600, 700, seed=3)
    cl = (np.arange(700) >= 100).astype(np.int32) * 7          # codes 0 and 7: 100 and 600 cells
    got, _, want = run_both(gpu_ctx, d, cl, max_cluster_size=250)   # 600 cells -> 3 sub-clusters of 200
    assert np.allclose(got, want, rtol=RTOL, atol=0)
    lim, ncl = O.limit_cluster_size(cl, 250)
    assert ncl == 4 and np.bincount(lim).tolist() == [100, 200, 200, 200]


def test_sparse_umi_like_counts_with_zero_medians(gpu_ctx):
    # the generator of the benchmark (heavy-tailed, ~8 % dense): many windows have a zero or tied median
    d = synth.make_counts(900, 2500, nnz_per_cell=200, seed=5)
    cl = (np.arange(900) % 3).astype(np.int32)
    got, nneg, want = run_both(gpu_ctx, d, cl, min_mean=0.05)
    assert nneg == int((want <= 0).sum())
    assert np.allclose(got, want, rtol=RTOL, atol=1e-12)


def test_custom_sizes_and_min_mean(gpu_ctx):
    d, _, _ = planted(500, 260, seed=4, depth=0.5)
    got, _, want = run_both(gpu_ctx, d, None, sizes=[10, 15, 40, 33], min_mean=1.0)
    assert np.allclose(got, want, rtol=RTOL, atol=0)


def test_errors_follow_scuttle(gpu_ctx):
    import cellula_b200
    d, _, _ = planted(300, 200, seed=6)
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    with pytest.raises(cellula_b200.Cb200Error, match="not enough cells in at least one cluster"):
        gpu_ctx.pooled_size_factors((np.arange(200) < 10).astype(np.int32))
    with pytest.raises(cellula_b200.Cb200Error, match="not unique"):
        gpu_ctx.pooled_size_factors(None, sizes=[21, 21])
    x = d["x"].copy()
    x[d["colptr"][5]:d["colptr"][6]] = 0.0
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], x)
    with pytest.raises(cellula_b200.Cb200Error, match="non-zero library sizes"):
        gpu_ctx.pooled_size_factors(None)


def test_deconvolution_factors_feed_logcounts(gpu_ctx):
    """doNormAndReduce order (R/doNormAndReduce.R:66-78): pooled factors -> logNormCounts re-centres and uses them."""
    d, _, _ = planted(400, 300, seed=7)
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    sf_pool, _ = gpu_ctx.pooled_size_factors(None)
    sf = gpu_ctx.size_factors(sf_pool)
    logx, _, _ = gpu_ctx.lognorm_genestats()
    o_sf = O.library_size_factors(d["colptr"], d["x"], size_factors_in=O.pooled_size_factors(
        d["n_genes"], d["colptr"], d["rowidx"], d["x"]))
    assert np.allclose(sf, o_sf, rtol=RTOL)
    assert np.allclose(logx, O.log_norm_counts(d["colptr"], d["x"], o_sf), rtol=1e-6, atol=1e-9)


def test_mirror_default_order_with_deconvolution_option(gpu_ctx):
    """cellula_b200.doNormAndReduce(options size_factors = 'deconvolution'): pooled factors -> centred -> logcounts."""
    import cellula_b200 as cb
    from cellula_b200.sce import SingleCellExperiment
    d, _, grp = planted(700, 400, seed=8, groups=2)
    old = dict(cb.options)
    try:
        cb.options.update(skip_umap=True, size_factors="deconvolution")
        sce = SingleCellExperiment.from_counts(d["colptr"], d["rowidx"], d["x"], 700)
        sce.colData["quickCluster"] = np.array(["b", "a"])[grp]
        sce = cb.doNormAndReduce(sce, ndims=5, hvg_ntop=200, verbose=False)
    finally:
        cb.options.update(old)
    codes = np.unique(np.array(["b", "a"])[grp], return_inverse=True)[1]
    want = O.pooled_size_factors(700, d["colptr"], d["rowidx"], d["x"], codes)
    want = want / want.mean()
    assert np.allclose(sce.colData["sizeFactor"], want, rtol=RTOL)
