"""End to end through the host mirror of the reference API, the device-resident path, the device
generator, and size-independent properties at a BASELINE-sized input."""
import numpy as np
import pytest

import cellula_b200 as cb
from cellula_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_api_end_to_end_against_oracle(small_counts, small_oracle):
    d, ref = small_counts, small_oracle
    cb.options["skip_umap"] = True
    sce = cb.SingleCellExperiment.from_counts(d["colptr"], d["rowidx"], d["x"], d["n_genes"],
                                              rownames=[f"g{i}" for i in range(d["n_genes"])])
    sce = cb.doNormAndReduce(sce, ndims=20, hvg_ntop=800, verbose=False)
    assert np.allclose(sce.colData["sizeFactor"], ref["sf"], rtol=1e-13)
    lc = sce.assays["logcounts"]
    assert lc.i is sce.counts.i and lc.p is sce.counts.p          # shares i/p like the R dgCMatrix
    assert np.allclose(lc.x, ref["logx"], rtol=1e-12)
    assert sce.metadata["hvgs"] == [f"g{i}" for i in ref["hvgs"]]   # identical HVGs, identical order
    for key in ("mean", "total", "tech", "bio", "p_value", "FDR"):
        assert key in sce.rowData
    pca = sce.reducedDims["PCA"]
    assert pca.shape == (d["n_cells"], 20) and pca.attrs["rotation"].shape == (800, 20)
    want = ref["pca"]["scores"]
    sign = np.sign(np.sum(np.asarray(pca) * want, axis=0))
    err = np.abs(np.asarray(pca) * sign - want).max(axis=0) / np.abs(want).max(axis=0)
    assert err.max() <= 1e-4
    assert np.allclose(pca.attrs["percentVar"], ref["pca"]["percent_var"], rtol=1e-6)

    sce = cb.makeGraphsAndClusters(sce, neighbors=10, weighting_scheme="rank", ndims=20, save_graphs=True)
    g = sce.metadata["SNN_10_rank"]
    knn = O.find_knn(np.asarray(pca), 10)                    # oracle fed the same embedding
    want_e = O.snn_edges(knn, "rank")
    assert np.array_equal(g.frm, want_e[0]) and np.array_equal(g.to, want_e[1]) and np.array_equal(g.weight, want_e[2])
    assert g.n == d["n_cells"]


def test_resident_path_equals_host_path(gpu_ctx, small_counts, small_oracle):
    d = small_counts
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    gpu_ctx.size_factors(want=False)
    gpu_ctx.lognorm_genestats(want_logx=False, want_stats=False)
    pca = gpu_ctx.pca(small_oracle["hvgs"], 20)
    a = gpu_ctx.knn(None, 10, n_total=d["n_cells"], d=20)       # embedding taken from HBM
    b = gpu_ctx.knn(pca["scores"], 10)
    assert np.array_equal(a, b)
    c = gpu_ctx.knn(None, 10, n_total=d["n_cells"], d=12)       # leading 12 PCs only (ndims truncation)
    assert np.array_equal(c, O.find_knn(pca["scores"][:, :12], 10))


def test_device_generator_roundtrip(gpu_ctx):
    nnz = gpu_ctx.synth_counts(3000, 5000, 250.0, seed=123)
    colptr, rowidx, x = gpu_ctx.export_csc()
    assert colptr[0] == 0 and colptr[-1] == nnz == len(x)
    lens = np.diff(colptr)
    assert lens.min() >= 1 and 150 < lens.mean() < 400
    assert np.all(x >= 1) and np.all(x == np.round(x))
    assert rowidx.min() >= 0 and rowidx.max() < 3000
    for c in (0, 17, 4999):                                    # rows sorted within a column
        seg = rowidx[colptr[c]:colptr[c + 1]]
        assert np.all(np.diff(seg) > 0)
    # same seed, shard of the same matrix: identical columns
    nnz2 = gpu_ctx.synth_counts(3000, 1000, 250.0, seed=123, cell_offset=2000, n_cells_total=5000)
    c2, r2, x2 = gpu_ctx.export_csc()
    lo, hi = colptr[2000], colptr[3000]
    assert nnz2 == hi - lo and np.array_equal(r2, rowidx[lo:hi]) and np.array_equal(x2, x[lo:hi])
    # and the generated matrix goes through the path
    d = dict(n_genes=3000, colptr=colptr, rowidx=rowidx, x=x)
    gpu_ctx.load_csc(3000, colptr, rowidx, x)
    sf = gpu_ctx.size_factors()
    assert np.allclose(sf, O.library_size_factors(colptr, x), rtol=1e-13)


def exact_knn_subset(emb, queries, k):
    """Oracle arithmetic (sequential fp64, no FMA, sqrt, (dist, idx)) for a few query rows."""
    acc = np.zeros((len(queries), emb.shape[0]))
    for t in range(emb.shape[1]):
        delta = emb[queries, t][:, None] - emb[None, :, t]
        acc += delta * delta
    dist = np.sqrt(acc)
    dist[np.arange(len(queries)), queries] = np.inf
    out = np.empty((len(queries), k), dtype=np.int32)
    for r in range(len(queries)):
        o = np.lexsort((np.arange(emb.shape[0]), dist[r]))
        out[r] = o[:k] + 1
    return out


@pytest.mark.timeout(1200)
def test_cfg2_sized_properties(gpu_ctx):
    """PBMC-68k-shaped input (68,579 x 32,738): properties that do not need the full oracle."""
    n, g, k, nd = 68579, 32738, 20, 50
    nnz = gpu_ctx.synth_counts(g, n, 525.0, seed=20261019)
    colptr, rowidx, x = gpu_ctx.export_csc()
    sf = gpu_ctx.size_factors()
    assert abs(sf.mean() - 1.0) < 1e-12 and sf.min() > 0
    logx, mean, var = gpu_ctx.lognorm_genestats()
    assert logx.min() > 0 and np.all(var >= 0)
    sub = np.random.default_rng(0).choice(g, 300, replace=False)      # checksum of per-gene sums
    sums = np.bincount(rowidx, weights=logx, minlength=g)
    assert np.allclose(mean[sub] * n, sums[sub], rtol=1e-10, atol=1e-9)
    from cellula_b200 import trendvar
    hv = trendvar.get_top_hvgs(trendvar.model_gene_var(mean, var), 2000)
    pca = gpu_ctx.pca(hv, nd)
    sc = pca["scores"]
    assert np.allclose(sc.mean(axis=0), 0, atol=1e-9 * np.abs(sc).max())
    assert np.allclose(sc.var(axis=0, ddof=1), pca["var_explained"], rtol=1e-6)
    gram = sc.T @ sc                                                  # scores are orthogonal (U D)
    off = gram - np.diag(np.diag(gram))
    assert np.abs(off).max() <= 1e-6 * np.diag(gram).max()
    knn = gpu_ctx.knn(sc, k)
    q = np.random.default_rng(1).choice(n, 256, replace=False)
    assert np.array_equal(knn[q], exact_knn_subset(sc, q, k))          # bit-exact on a sample
    f, t, w = gpu_ctx.snn(knn, "rank")
    assert np.all(t < f) and np.all(np.diff(f) >= 0)                   # (j, h<j), grouped by ascending j
    assert w.min() >= 1e-6 and w.max() <= k
    assert np.all((w * 2 == np.round(w * 2)) | (w == 1e-6))          # k - r/2, floored at 1e-6
    key = f.astype(np.int64) * (n + 1) + t
    assert len(np.unique(key)) == len(key)                            # simple graph
    # every kNN pair is an SNN edge (they share the neighbour itself)
    probe = np.random.default_rng(2).choice(n, 2000, replace=False)
    have = set(key[np.isin(f, probe + 1) | np.isin(t, probe + 1)].tolist())
    for j in probe[:300]:
        for h in knn[j]:
            a, b = int(max(j + 1, h)), int(min(j + 1, h))
            assert a * (n + 1) + b in have


def test_cuda_shard_refuses_a_context_on_another_stream(gpu_ctx):
    """NCCL collectives are ordered against torch's current stream only, so CudaShard accepts nothing but a
    context that runs on it (a private library stream would let an all-reduce read a buffer that is still being written)."""
    import torch
    import cellula_b200
    from cellula_b200.dist import CudaShard
    with pytest.raises(RuntimeError, match="torch's current CUDA stream"):
        CudaShard(gpu_ctx, 0)                       # session context: private stream made by the library
    s = torch.cuda.Stream(device=0)
    ctx = cellula_b200.Context(0, stream=s.cuda_stream)
    try:
        with pytest.raises(RuntimeError, match="torch's current CUDA stream"):
            CudaShard(ctx, 0)                       # right stream, but not current
        with torch.cuda.stream(s):
            CudaShard(ctx, 0)                       # accepted
    finally:
        ctx.close()
