"""cb200_multi_*: the hot path over several GPUs from ONE host process (the R-host boundary of SURVEY.md 8b).
Every stage must reproduce the single-context result -- bit for bit for the sums (fixed-point integers), the PCA
(integer cross-product), the kNN index and the SNN edge list.  Runs with every GPU count the box offers
(1 GPU: the same code with the collectives skipped)."""
import numpy as np
import pytest

import cellula_b200 as cb
from cellula_b200 import synth, trendvar
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def gpu_counts():
    try:
        import torch
        n = torch.cuda.device_count()
    except Exception:
        n = 1
    return [c for c in (1, 2, 4, 8) if c <= max(n, 1)]


@pytest.fixture(scope="module")
def counts():
    return synth.make_counts(6000, 5000, nnz_per_cell=300, seed=77)


@pytest.fixture(scope="module")
def single(counts):
    d = counts
    ctx = cb.Context(0)
    ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    sf = ctx.size_factors()
    logx, mean, var = ctx.lognorm_genestats()
    hv = trendvar.get_top_hvgs(trendvar.model_gene_var(mean, var), 1000)
    pca = ctx.pca(hv, 20)
    knn = ctx.knn(None, 15, n_total=d["n_cells"], d=20)
    edges = ctx.snn(None, "rank", n_total=d["n_cells"], k=15)
    digest = ctx.edges_checksum()
    ctx.close()
    return dict(sf=sf, logx=logx, mean=mean, var=var, hv=hv, pca=pca, knn=knn, edges=edges, digest=digest)


@pytest.mark.parametrize("n_gpus", gpu_counts())
def test_multi_equals_single_bit_for_bit(counts, single, n_gpus):
    d = counts
    m = cb.MultiContext(n_gpus)
    assert m.n_gpus == n_gpus
    m.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    assert np.array_equal(m.size_factors(), single["sf"])
    logx, mean, var = m.lognorm_genestats()
    assert np.array_equal(logx, single["logx"])
    assert np.array_equal(mean, single["mean"]) and np.array_equal(var, single["var"])      # integer sums: exact
    pca = m.pca(single["hv"], 20)
    assert np.array_equal(pca["scores"], single["pca"]["scores"])                            # integer cross-product
    assert np.array_equal(pca["rotation"], single["pca"]["rotation"])
    assert np.array_equal(pca["var_explained"], single["pca"]["var_explained"])
    knn = m.knn(None, 15, n_total=d["n_cells"], d=20)
    assert np.array_equal(knn, single["knn"])
    f, t, w = m.snn(None, "rank", n_total=d["n_cells"], k=15)
    for a, b in zip((f, t, w), single["edges"]):
        assert np.array_equal(a, b)                                                          # same edges, same ORDER
    assert m.edges_checksum() == single["digest"]
    m.close()


@pytest.mark.parametrize("n_gpus", gpu_counts())
def test_multi_knn_snn_on_a_host_matrix(n_gpus):
    emb = synth.make_embedding(9000, 12, seed=5)
    m = cb.MultiContext(n_gpus)
    knn = m.knn(emb, 10)
    assert np.array_equal(knn, O.find_knn(emb, 10))
    for scheme in ("rank", "jaccard", "number"):
        got = m.snn(knn, scheme)
        want = O.snn_edges(knn, scheme)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), scheme
    m.close()


def test_multi_errors_are_reported():
    with pytest.raises(cb.Cb200Error, match="GPUs requested"):
        cb.MultiContext(64)
    m = cb.MultiContext(1)
    with pytest.raises(cb.Cb200Error, match="no embedding given"):
        m.knn(None, 5, n_total=100, d=3)
    m.close()
