"""Host-side mirror of the reference entry points: argument validation and error text follow
/root/reference/R/doNormAndReduce.R:45-53 and R/makeGraphsAndClusters.R:65-72.  These checks
fire before any device work, so they run without a GPU."""
import numpy as np
import pytest

import cellula_b200 as cb
from cellula_b200 import trendvar
from cellula_b200.dist import shard_bounds, slice_csc


def tiny_sce():
    return cb.SingleCellExperiment.from_counts([0, 1, 2], [0, 1], [1.0, 2.0], n_genes=3)


def test_donormandreduce_validation():
    with pytest.raises(TypeError, match="must provide a SingleCellExperiment object"):
        cb.doNormAndReduce(object())
    with pytest.raises(ValueError, match="batch column not found in the colData of the object"):
        cb.doNormAndReduce(tiny_sce(), batch="nope")
    with pytest.raises(ValueError, match="hvg_ntop cannot be higher than the number of features"):
        cb.doNormAndReduce(tiny_sce(), hvg_ntop=10)
    # (the batch-aware branch itself runs on the GPU: tests/test_gpu_batch.py)


def test_makegraphsandclusters_validation():
    with pytest.raises(TypeError, match="must provide a SingleCellExperiment object"):
        cb.makeGraphsAndClusters(42)
    sce = tiny_sce()
    with pytest.raises(ValueError, match='must be one of "leiden" or "louvain"'):
        cb.makeGraphsAndClusters(sce, method="walktrap")
    with pytest.raises(ValueError, match='must be one of "jaccard", "rank", or "number"'):
        cb.makeGraphsAndClusters(sce, weighting_scheme="cosine")
    with pytest.raises(ValueError, match="there are no dimensionality reductions"):
        cb.makeGraphsAndClusters(sce)
    sce.reducedDims["PCA"] = np.zeros((2, 2))
    with pytest.raises(ValueError, match="the dr name supplied was not found"):
        cb.makeGraphsAndClusters(sce, dr="UMAP")


def test_signatures_match_reference_defaults():
    import inspect
    s = inspect.signature(cb.doNormAndReduce).parameters
    assert list(s) == ["sce", "batch", "name", "ndims", "hvg_ntop", "umap_seed", "verbose", "parallel_param"]
    assert (s["ndims"].default, s["hvg_ntop"].default, s["umap_seed"].default) == (20, 2000, 11)
    g = inspect.signature(cb.makeGraphsAndClusters).parameters
    assert list(g) == ["sce", "neighbors", "weighting_scheme", "sweep_on", "method", "k", "dr", "ndims",
                       "calculate_modularity", "calculate_silhouette", "leiden_iterations", "save_graphs",
                       "prefix", "verbose"]
    assert (g["neighbors"].default, g["weighting_scheme"].default, g["ndims"].default) == (10, "jaccard", 20)


def test_shard_bounds_and_slicing():
    assert shard_bounds(10, 3) == [0, 4, 7, 10]
    assert shard_bounds(8, 8) == list(range(9))
    colptr = np.array([0, 2, 2, 5, 6])
    rowidx = np.arange(6, dtype=np.int32)
    x = np.arange(6, dtype=float)
    p, i, v = slice_csc(colptr, rowidx, x, 1, 3)
    assert p.tolist() == [0, 0, 3] and i.tolist() == [2, 3, 4] and v.tolist() == [2.0, 3.0, 4.0]


def test_host_trend_agrees_with_oracle_restatement(small_counts):
    """The package's host-side HVG step and the oracle's restatement are written separately;
    fed the same means/variances they must pick the same HVGs in the same order."""
    from oracle import oracle as O
    d = small_counts
    sf = O.library_size_factors(d["colptr"], d["x"])
    lx = O.log_norm_counts(d["colptr"], d["x"], sf)
    mean, var = O.gene_mean_var(d["n_genes"], d["n_cells"], d["rowidx"], lx)
    a = trendvar.model_gene_var(mean, var)
    b = O.model_gene_var(mean, var)
    assert np.allclose(a["tech"], b["tech"], rtol=1e-7, atol=1e-12)
    assert np.array_equal(trendvar.get_top_hvgs(a, 800), O.get_top_hvgs(b["bio"], 800))
    # trend properties: positive where the mean is positive, zero at zero
    tr, sd = trendvar.fit_trend_var(mean, var)
    assert sd > 0 and tr(np.array([0.0]))[0] == 0.0 and np.all(tr(np.linspace(0.01, 5, 50)) > 0)


def test_snn_bounds_balance_and_cover():
    """SNN cell ranges of the sharded pipeline: contiguous, strictly increasing, covering [0, N), and with
    about equal numbers of (j, h < j) pairs per rank (work grows like j)."""
    from cellula_b200.dist import shard_bounds, snn_bounds
    for n, world in [(1306127, 8), (1306127, 3), (100, 4), (5, 4), (7, 1)]:
        b = snn_bounds(n, world)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1
        assert all(b[r] < b[r + 1] for r in range(world))
        assert shard_bounds(n, world)[-1] == n
    b = snn_bounds(1306127, 8)
    work = [b[r + 1] ** 2 - b[r] ** 2 for r in range(8)]
    assert max(work) / min(work) < 1.001


def test_pool_layout_matches_the_oracle():
    """Host half of cb200_pooled_size_factors (cluster size limit + ring order) against the oracle restatement of
    scuttle:::.limit_cluster_size / .generateSphere -- no GPU needed (cb200_host_pool_layout does no device work)."""
    import ctypes
    import numpy as np
    from cellula_b200 import _lib
    from oracle import oracle as O
    L = _lib.lib()
    rng = np.random.default_rng(5)
    for n, ncl, max_size in ((1, 1, 3000), (50, 1, 0), (997, 4, 100), (3000, 7, 250), (400, 3, 3000)):
        clusters = (rng.integers(0, ncl, n) * 3 + 2).astype(np.int32)          # arbitrary non-contiguous codes
        lib = rng.integers(5, 40, n).astype(np.float64)                         # many ties: the sort must be stable
        cl, ring = np.empty(n, np.int32), np.empty(n, np.int32)
        ptr, ml = np.empty(n + 1, np.int64), np.empty(n)
        nc = ctypes.c_int32()
        rc = L.cb200_host_pool_layout(n, clusters.ctypes.data, lib.ctypes.data, max_size, cl.ctypes.data, ring.ctypes.data,
                                      ptr.ctypes.data, ml.ctypes.data, ctypes.cast(ctypes.byref(nc), ctypes.c_void_p))
        assert rc == 0
        o_cl, o_n = O.limit_cluster_size(clusters, max_size if max_size > 0 else None)
        assert nc.value == o_n and np.array_equal(cl, o_cl)
        for c in range(o_n):
            cells = np.flatnonzero(o_cl == c)
            assert ptr[c + 1] - ptr[c] == cells.size
            assert np.array_equal(ring[ptr[c]:ptr[c + 1]], cells[O.pool_ring(lib[cells])])
            assert ml[c] == lib[cells].sum() / cells.size
    bad = np.array([0, -1, 2], dtype=np.int32)
    rc = L.cb200_host_pool_layout(3, bad.ctypes.data, np.ones(3).ctypes.data, 0, np.empty(3, np.int32).ctypes.data,
                                  np.empty(3, np.int32).ctypes.data, np.empty(4, np.int64).ctypes.data, None,
                                  ctypes.cast(ctypes.byref(ctypes.c_int32()), ctypes.c_void_p))
    assert rc == 2


def test_deconvolution_option_routes_cluster_labels_to_the_pooling(monkeypatch):
    """options['size_factors'] = 'deconvolution' (R/doNormAndReduce.R:57-71 order): the labels of colData['quickCluster']
    become 0-based codes for cb200_pooled_size_factors, its vector is what logNormCounts re-centres, and non-positive
    estimates stop the call.  Driven with a recording stand-in for the device context (no GPU here)."""
    import cellula_b200.api as api

    class Stop(Exception):
        pass

    class FakeCtx:
        def __init__(self, n_bad):
            self.calls, self.n_bad = [], n_bad

        def load_csc(self, *a, **k):
            self.calls.append("load")

        def pooled_size_factors(self, codes):
            self.calls.append(("pool", None if codes is None else codes.copy()))
            return np.array([2.0, 4.0]), self.n_bad

        def size_factors(self, sf_in):
            self.calls.append(("sf", None if sf_in is None else np.asarray(sf_in).copy()))
            raise Stop()

    old = dict(api.options)
    try:
        api.options.update(size_factors="deconvolution", skip_umap=True)
        sce = tiny_sce()
        sce.colData["quickCluster"] = np.array(["z", "a"])
        fake = FakeCtx(0)
        monkeypatch.setattr(api, "get_context", lambda: fake)
        with pytest.raises(Stop):
            api.doNormAndReduce(sce, hvg_ntop=2, verbose=False)
        assert fake.calls[0] == "load" and fake.calls[1][0] == "pool" and fake.calls[1][1].tolist() == [1, 0]
        assert fake.calls[2][0] == "sf" and fake.calls[2][1].tolist() == [2.0, 4.0]
        fake = FakeCtx(0)
        monkeypatch.setattr(api, "get_context", lambda: fake)
        with pytest.raises(Stop):
            api.doNormAndReduce(tiny_sce(), hvg_ntop=2, verbose=False)          # no labels: one cluster
        assert fake.calls[1] == ("pool", None)
        fake = FakeCtx(1)
        monkeypatch.setattr(api, "get_context", lambda: fake)
        with pytest.raises(ValueError, match="non-positive size factor"):
            api.doNormAndReduce(tiny_sce(), hvg_ntop=2, verbose=False)
        api.options.update(size_factors="library")
        fake = FakeCtx(0)
        monkeypatch.setattr(api, "get_context", lambda: fake)
        with pytest.raises(Stop):
            api.doNormAndReduce(tiny_sce(), hvg_ntop=2, verbose=False)
        assert fake.calls[1] == ("sf", None)                                      # library-size factors: no pooling
    finally:
        api.options.clear()
        api.options.update(old)
