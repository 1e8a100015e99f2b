"""GPU parity, stage K6 (exact kNN): neighbour indices must be bit-exact against the oracle's
(distance, index) ordering, including ties, duplicates and every supported k -- with both tile
engines (tcgen05 tensor cores, fp32 CUDA cores)."""
import json
import os

import numpy as np
import pytest

import cellula_b200
from cellula_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
ENGINE_ID = {"simt": 1, "tc": 2}


@pytest.fixture(params=["tc", "simt"])
def engine(request, monkeypatch):
    monkeypatch.setenv("CB200_KNN_ENGINE", request.param)
    return request.param


def test_kat_knn(gpu_ctx, engine):
    x = np.asarray(KAT["snn1"]["x"], dtype=np.float64)[:, None]
    assert gpu_ctx.knn(x, 2).tolist() == KAT["snn1"]["knn"]
    assert gpu_ctx.timings()["knn_engine"] == ENGINE_ID[engine]
    x = np.asarray(KAT["knn_tie"]["x"], dtype=np.float64)[:, None]
    assert gpu_ctx.knn(x, 1).tolist() == KAT["knn_tie"]["knn"]


@pytest.mark.parametrize("n,d,k", [(3000, 50, 20), (3000, 50, 10), (2500, 30, 30), (1500, 7, 5), (777, 1, 3),
                                   (4097, 61, 24), (129, 50, 52), (128, 16, 40), (257, 3, 25)])
def test_knn_bit_exact_random(gpu_ctx, engine, n, d, k):
    emb = synth.make_embedding(n, d, seed=n + d + k)
    got = gpu_ctx.knn(emb, k)
    assert gpu_ctx.timings()["knn_engine"] == ENGINE_ID[engine]
    assert got.dtype == np.int32 and got.shape == (n, k)
    want = O.find_knn(emb, k)
    bad = np.flatnonzero((got != want).any(axis=1))
    assert bad.size == 0, (bad[:10], got[bad[:3]], want[bad[:3]])


@pytest.mark.parametrize("n,d,k", [(2000, 50, 100), (700, 20, 53), (5000, 30, 77)])
def test_knn_large_k_on_the_tensor_engine(gpu_ctx, monkeypatch, n, d, k):
    """52 < k <= 100: 112-entry heaps and a single query-image buffer (cfg4's k = 100)."""
    monkeypatch.setenv("CB200_KNN_ENGINE", "tc")
    for clusters in ("", "9"):
        if clusters:
            monkeypatch.setenv("CB200_KNN_CLUSTERS", clusters)
        emb = synth.make_embedding(n, d, seed=n + d + k)
        assert np.array_equal(gpu_ctx.knn(emb, k), O.find_knn(emb, k))
        assert gpu_ctx.timings()["knn_engine"] == 2


@pytest.mark.parametrize("n,d,k", [(4097, 64, 24), (900, 62, 10), (1500, 20, 105)])
def test_knn_shapes_only_the_simt_engine_takes(gpu_ctx, n, d, k):
    emb = synth.make_embedding(n, d, seed=n + d + k)
    assert np.array_equal(gpu_ctx.knn(emb, k), O.find_knn(emb, k))
    assert gpu_ctx.timings()["knn_engine"] == 1


def test_knn_bit_exact_20k(gpu_ctx, engine):
    emb = synth.make_embedding(20000, 50, seed=99)
    got, want = gpu_ctx.knn(emb, 20), O.find_knn(emb, 20)
    bad = np.flatnonzero((got != want).any(axis=1))
    assert bad.size == 0, (bad.size, bad[:10])
    # the margin test must prove (nearly) every query without the exact-scan fallback
    assert gpu_ctx.timings()["knn_fallback_queries"] <= 20


def test_knn_integer_lattice_ties(gpu_ctx, engine):
    """Massive exact ties: neighbours must come out in index order within equal distances.
    Also forces the margin test to fail, so the exact-scan fallback is exercised."""
    rng = np.random.default_rng(5)
    emb = rng.integers(-2, 3, size=(1500, 4)).astype(np.float64)
    got = gpu_ctx.knn(emb, 15)
    assert np.array_equal(got, O.find_knn(emb, 15))
    # 9 distinct positions, ~440 copies of each: far more exact ties than candidates are kept
    emb = rng.integers(-1, 2, size=(4000, 2)).astype(np.float64)
    got = gpu_ctx.knn(emb, 15)
    assert np.array_equal(got, O.find_knn(emb, 15))
    assert gpu_ctx.timings()["knn_fallback_queries"] > 0


def test_knn_duplicates(gpu_ctx, engine):
    emb = synth.make_embedding(600, 10, seed=1)
    emb[100:140] = emb[100]          # 40 identical cells
    emb[300] = emb[7]
    got = gpu_ctx.knn(emb, 12)
    assert np.array_equal(got, O.find_knn(emb, 12))
    assert all(i + 1 not in got[i] for i in range(600))


def test_knn_query_range(gpu_ctx, engine):
    emb = synth.make_embedding(2000, 20, seed=2)
    full = O.find_knn(emb, 10)
    part = gpu_ctx.knn(emb, 10, q_begin=513, q_end=1400)
    assert np.array_equal(part, full[513:1400])


@pytest.mark.parametrize("shift,scale", [(1000.0, 1.0), (0.0, 1e-6), (0.0, 1e5), (-3e4, 7.0)])
def test_knn_exact_under_offsets_and_scales(gpu_ctx, engine, shift, scale):
    """Large offsets make the norm expansion lossy, tiny/huge values stress the fp16 split;
    the result must stay exact (through the margin test / fallback)."""
    emb = synth.make_embedding(1500, 20, seed=4) * scale + shift
    assert np.array_equal(gpu_ctx.knn(emb, 10), O.find_knn(emb, 10))


def test_knn_engines_agree_on_candidates_quality(gpu_ctx, monkeypatch):
    """Both engines must need (almost) no fallback on well-separated data."""
    emb = synth.make_embedding(6000, 50, seed=21)
    for eng in ("tc", "simt"):
        monkeypatch.setenv("CB200_KNN_ENGINE", eng)
        gpu_ctx.knn(emb, 20, want=False)
        assert gpu_ctx.timings()["knn_fallback_queries"] <= 6, eng


def test_knn_degenerate_first_coordinate(gpu_ctx, engine):
    """Constant, two-valued and outlier-ridden coordinates (degenerate clusters, a scale set by a few
    outliers) must not change the (exact) result."""
    rng = np.random.default_rng(9)
    emb = synth.make_embedding(2000, 12, seed=8)
    for mutate in ("const", "binary", "outlier"):
        e = emb.copy()
        if mutate == "const":
            e[:, 0] = 3.25
        elif mutate == "binary":
            e[:, 0] = rng.integers(0, 2, 2000) * 50.0
        else:
            e[::400, 0] = 1e6
        assert np.array_equal(gpu_ctx.knn(e, 10), O.find_knn(e, 10)), mutate


@pytest.mark.parametrize("clusters", [1, 2, 7, 23, 256])
def test_knn_any_cluster_count_is_exact(gpu_ctx, monkeypatch, clusters):
    """The tensor-core engine orders the references by (cluster, distance to the cluster centre) and
    skips tiles by the triangle inequality; the number of clusters must never change the result --
    random data, a query sub-range, duplicates and a tie-heavy lattice."""
    monkeypatch.setenv("CB200_KNN_ENGINE", "tc")
    monkeypatch.setenv("CB200_KNN_CLUSTERS", str(clusters))
    emb = synth.make_embedding(3000, 50, seed=77)
    want = O.find_knn(emb, 20)
    assert np.array_equal(gpu_ctx.knn(emb, 20), want)
    assert gpu_ctx.timings()["knn_engine"] == 2
    assert np.array_equal(gpu_ctx.knn(emb, 20, q_begin=1000, q_end=1777), want[1000:1777])
    emb[100:140] = emb[100]
    assert np.array_equal(gpu_ctx.knn(emb, 12), O.find_knn(emb, 12))
    rng = np.random.default_rng(5)
    lat = rng.integers(-2, 3, size=(1500, 4)).astype(np.float64)
    assert np.array_equal(gpu_ctx.knn(lat, 15), O.find_knn(lat, 15))
    same = np.ones((700, 6))                      # every point identical: all distances tie at 0
    assert np.array_equal(gpu_ctx.knn(same, 9), O.find_knn(same, 9))


def test_knn_cluster_pruning_on_blobs(gpu_ctx, monkeypatch):
    """40 Gaussian blobs in 50 dimensions: most of the tile grid must be skipped, exactly."""
    monkeypatch.setenv("CB200_KNN_ENGINE", "tc")
    emb = synth.make_embedding(60000, 50, seed=31)
    got = gpu_ctx.knn(emb, 20)
    t = gpu_ctx.timings()
    assert t["knn_tiles_scanned"] < 0.35 * t["knn_tiles_total"], t
    assert t["knn_fallback_queries"] <= 20
    q = np.random.default_rng(1).choice(60000, 300, replace=False)
    sub = O.find_knn_queries(emb, q, 20) if hasattr(O, "find_knn_queries") else None
    if sub is None:
        acc = np.zeros((300, 60000))
        for c in range(50):                       # the reference's summation order
            dlt = emb[q, c][:, None] - emb[None, :, c]
            acc += dlt * dlt
        dist = np.sqrt(acc)
        dist[np.arange(300), q] = np.inf
        sub = np.array([np.lexsort((np.arange(60000), dist[r]))[:20] + 1 for r in range(300)], dtype=np.int32)
    assert np.array_equal(got[q], sub)


def test_knn_rejects_non_finite(gpu_ctx):
    emb = synth.make_embedding(500, 8, seed=8)
    emb[17, 3] = np.nan
    with pytest.raises(cellula_b200.Cb200Error, match="non-finite"):
        gpu_ctx.knn(emb, 5)


def test_knn_pruning_reports_tiles(gpu_ctx):
    """Well-separated slabs along the first coordinate: most of the tile grid must be pruned."""
    rng = np.random.default_rng(3)
    emb = rng.standard_normal((40000, 10))
    emb[:, 0] += np.repeat(np.arange(40) * 40.0, 1000)       # 40 slabs far apart in PC 1
    emb = emb[rng.permutation(40000)]
    got = gpu_ctx.knn(emb, 10)
    t = gpu_ctx.timings()
    assert t["knn_engine"] == 2 and t["knn_tiles_scanned"] < 0.2 * t["knn_tiles_total"], t
    q = rng.choice(40000, 200, replace=False)
    acc = np.zeros((200, 40000))
    for c in range(10):
        dlt = emb[q, c][:, None] - emb[None, :, c]
        acc += dlt * dlt
    dist = np.sqrt(acc)
    dist[np.arange(200), q] = np.inf
    want = np.array([np.lexsort((np.arange(40000), dist[r]))[:10] + 1 for r in range(200)], dtype=np.int32)
    assert np.array_equal(got[q], want)


def test_knn_argument_errors(gpu_ctx):
    emb = synth.make_embedding(50, 5, seed=3)
    with pytest.raises(cellula_b200.Cb200Error, match="k <= n-1"):
        gpu_ctx.knn(emb, 50)
    with pytest.raises(cellula_b200.Cb200Error, match="k <= 2048"):
        gpu_ctx.knn(synth.make_embedding(3000, 5, seed=3), 2500)


@pytest.mark.parametrize("n,d,k", [(300, 5, 120), (100, 70, 5), (5000, 20, 113), (12544, 30, 112), (9000, 16, 255),
                                   (4500, 100, 67), (2500, 8, 600)])
def test_knn_general_path_for_large_k_and_wide_embeddings(gpu_ctx, n, d, k):
    """Shapes outside the tile engines: k > 112 (rebuildModularity's default k = round(sqrt(N)), uwot's
    n_neighbors = floor(sqrt(N))) and more than 64 dimensions go through the exact general kernel."""
    emb = synth.make_embedding(n, d, seed=n + d + k)
    assert np.array_equal(gpu_ctx.knn(emb, k), O.find_knn(emb, k))
    if k > 112 or d > 64:
        assert gpu_ctx.timings()["knn_engine"] == 3


def test_knn_general_path_with_ties(gpu_ctx):
    rng = np.random.default_rng(11)
    lat = rng.integers(-3, 4, size=(4000, 3)).astype(np.float64)       # hundreds of exact ties per query
    assert np.array_equal(gpu_ctx.knn(lat, 150), O.find_knn(lat, 150))


def test_knn_two_product_mode_is_exact(gpu_ctx, monkeypatch):
    """CB200_KNN_PRODUCTS=2 (experimental): the q_l r_h product is dropped and the recorded error bound widened;
    the result must stay exact (margin proof or fallback)."""
    monkeypatch.setenv("CB200_KNN_ENGINE", "tc")
    monkeypatch.setenv("CB200_KNN_PRODUCTS", "2")
    for n, d, k, seed in [(20000, 50, 20, 99), (3000, 12, 10, 3)]:
        emb = synth.make_embedding(n, d, seed=seed)
        assert np.array_equal(gpu_ctx.knn(emb, k), O.find_knn(emb, k))
    rng = np.random.default_rng(5)
    lat = rng.integers(-2, 3, size=(1500, 4)).astype(np.float64)
    assert np.array_equal(gpu_ctx.knn(lat, 15), O.find_knn(lat, 15))
