"""GPU parity, stage K7 (SNN edge weighting): edges and weights bit-exact against the oracle,
in canonical form and in emission order, for rank / number / jaccard."""
import json
import os

import numpy as np
import pytest

import cellula_b200
from cellula_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
SCHEMES = ["rank", "number", "jaccard"]


def assert_same_edges(got, want):
    for g, w in zip(got, want):
        assert np.array_equal(g, w)                      # same order (libscran emission order)
    cg, cw = O.canonical_edges(*got), O.canonical_edges(*want)
    for g, w in zip(cg, cw):
        assert np.array_equal(g, w)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_kat_snn(gpu_ctx, scheme):
    knn = np.asarray(KAT["snn1"]["knn"], dtype=np.int32)
    f, t, w = gpu_ctx.snn(knn, scheme)
    want = sorted((min(a, b), max(a, b), float(c)) for a, b, c in KAT["snn1"]["edges"][scheme])
    got = sorted((min(a, b), max(a, b), float(c)) for a, b, c in zip(f.tolist(), t.tolist(), w.tolist()))
    assert [g[:2] for g in got] == [x[:2] for x in want]
    assert np.allclose([g[2] for g in got], [x[2] for x in want], rtol=0, atol=1e-15)
    assert [[a, b] for a, b in zip(f.tolist(), t.tolist())] == KAT["snn1"]["emission_order_8c1_7"]["pairs"]


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("n,k", [(3000, 10), (3000, 20), (2000, 30), (1200, 100), (500, 1), (64, 40)])
def test_snn_bit_exact(gpu_ctx, scheme, n, k):
    knn = O.find_knn(synth.make_embedding(n, 20, seed=n + k), k)
    assert_same_edges(gpu_ctx.snn(knn, scheme), O.snn_edges(knn, scheme))


def test_snn_hub_cells(gpu_ctx):
    """A few cells sit in thousands of neighbour lists: long reverse lists."""
    rng = np.random.default_rng(8)
    n, k = 6000, 10
    knn = np.empty((n, k), dtype=np.int32)
    for j in range(n):
        cand = np.setdiff1d(np.concatenate([np.arange(5), rng.choice(n, 30, replace=False)]), [j],
                            assume_unique=False)
        hubs = [c for c in range(5) if c != j][:4]
        rest = [c for c in cand if c not in hubs][:k - len(hubs)]
        knn[j] = np.array(hubs + rest) + 1
    for scheme in SCHEMES:
        assert_same_edges(gpu_ctx.snn(knn, scheme), O.snn_edges(knn, scheme))


def test_snn_cell_range_shards_concatenate(gpu_ctx):
    knn = O.find_knn(synth.make_embedding(2500, 20, seed=6), 15)
    want = O.snn_edges(knn, "rank")
    parts = [gpu_ctx.snn(knn, "rank", j_begin=a, j_end=b) for a, b in [(0, 700), (700, 1301), (1301, 2500)]]
    got = tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
    assert_same_edges(got, want)


def test_snn_resident_knn_equals_host_index(gpu_ctx):
    emb = synth.make_embedding(3000, 30, seed=12)
    knn = gpu_ctx.knn(emb, 20)
    a = gpu_ctx.snn(None, "jaccard", n_total=3000, k=20)     # uses the kNN left in HBM
    b = gpu_ctx.snn(knn, "jaccard")
    assert_same_edges(a, b)
    assert gpu_ctx.snn_resident(20, "jaccard") == len(b[0])


@pytest.mark.parametrize("scheme", SCHEMES)
def test_snn_two_phase_into_caller_buffers(gpu_ctx, scheme):
    """count -> caller sizes its arrays -> emit in pieces with the copies behind the emission: same
    edges, same order, for the whole range and for a sub-range of cells."""
    import torch
    emb = synth.make_embedding(20000, 20, seed=44)
    knn = gpu_ctx.knn(emb, 20)
    want = O.snn_edges(knn, scheme)
    ne = gpu_ctx.snn_count_device(0, 20000, 20, scheme, 0, 20000)       # 0: the resident kNN result
    assert ne == len(want[0])
    f = torch.empty(ne + 5, dtype=torch.int32).pin_memory().numpy()
    t = torch.empty(ne + 5, dtype=torch.int32).pin_memory().numpy()
    w = torch.empty(ne + 5, dtype=torch.float64).pin_memory().numpy()
    assert_same_edges(gpu_ctx.snn_emit(ne, f, t, w), want)
    assert_same_edges(gpu_ctx.fetch_edges(ne), want)                     # and they are left in HBM too
    with pytest.raises(cellula_b200.Cb200Error, match="no counted SNN range"):
        gpu_ctx.snn_emit(ne, f, t, w)
    full = gpu_ctx.snn(knn, scheme, j_begin=5000, j_end=12345)
    ne2 = gpu_ctx.snn_count_device(0, 20000, 20, scheme, 5000, 12345)
    assert_same_edges(gpu_ctx.snn_emit(ne2), full)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_snn_hash_and_merge_engines_agree(gpu_ctx, monkeypatch, scheme):
    """The hash-table kernel (default) and the k-way merge kernel (fallback for cells with more partners
    than the table holds; forced here with CB200_SNN_ENGINE=merge) must emit identical edge lists."""
    knn = O.find_knn(synth.make_embedding(30000, 10, seed=3), 20)
    a = gpu_ctx.snn(knn, scheme)
    monkeypatch.setenv("CB200_SNN_ENGINE", "merge")
    b = gpu_ctx.snn(knn, scheme)
    assert_same_edges(a, b)
    assert_same_edges(a, O.snn_edges(knn, scheme))


def test_snn_argument_errors(gpu_ctx):
    knn = O.find_knn(synth.make_embedding(100, 5, seed=1), 5)
    bad = knn.copy()
    bad[3, 2] = 101
    with pytest.raises(cellula_b200.Cb200Error, match="out of range"):
        gpu_ctx.snn(bad, "rank")
    with pytest.raises(KeyError):
        gpu_ctx.snn(knn, "cosine")


@pytest.mark.parametrize("n,k", [(3000, 127), (2500, 200), (1500, 255)])
def test_snn_large_k_as_rebuild_modularity_uses(gpu_ctx, n, k):
    """k up to 255 (rebuildModularity's round(sqrt(N)) for up to 65 025 cells): every cell goes to the largest table
    tiers or the merge kernel; edges, weights and emission order stay bit-exact."""
    emb = synth.make_embedding(n, 10, seed=n + k)
    knn = O.find_knn(emb, k)
    for scheme in ("rank", "jaccard"):
        got = gpu_ctx.snn(knn, scheme, j_begin=n - 300, j_end=n)
        want = O.snn_edges(knn, scheme, j_begin=n - 300, j_end=n)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), (k, scheme)


def test_untrusted_index_is_validated(gpu_ctx):
    """a malformed neighbour index (out of range, own cell, repeated cell) is refused, never indexed with."""
    import cellula_b200
    emb = synth.make_embedding(600, 5, seed=3)
    good = O.find_knn(emb, 8)
    for bad_value, pattern in ((601, "out of range"), (0, "out of range")):
        bad = good.copy()
        bad[17, 3] = bad_value
        with pytest.raises(cellula_b200.Cb200Error, match=pattern):
            gpu_ctx.snn(bad, "rank")
    bad = good.copy()
    bad[40, 2] = 41                      # 1-based: row 40 is cell 41
    with pytest.raises(cellula_b200.Cb200Error, match="own cell"):
        gpu_ctx.snn(bad, "rank")
    bad = good.copy()
    bad[5, 7] = bad[5, 0]
    with pytest.raises(cellula_b200.Cb200Error, match="twice"):
        gpu_ctx.snn(bad, "number")
    got = O.canonical_edges(*gpu_ctx.snn(good, "rank"))          # the context is still usable afterwards
    want = O.canonical_edges(*O.snn_edges(good, "rank"))
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


@pytest.mark.parametrize("scheme", SCHEMES)
def test_snn_cta_tables_equal_warp_tables(gpu_ctx, monkeypatch, scheme):
    """Cells whose flat sequence needs the 8 Ki / 16 Ki-slot tables run on the CTA-per-cell kernel (one shared table,
    atomic insertion, emission order rebuilt from the smallest flat index of every partner); CB200_SNN_CTA=0 keeps the
    warp-per-cell kernel.  Same edges, weights and order, and both equal the oracle."""
    knn = O.find_knn(synth.make_embedding(4000, 12, seed=17), 90)
    a = gpu_ctx.snn(knn, scheme)
    monkeypatch.setenv("CB200_SNN_CTA", "0")
    b = gpu_ctx.snn(knn, scheme)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    want = O.snn_edges(knn, scheme)
    assert all(np.array_equal(x, y) for x, y in zip(a, want))


def test_snn_long_sequences_and_their_hand_over_to_the_merge_kernel(gpu_ctx, monkeypatch):
    """Cells whose flat sequence exceeds the largest table's capacity run on the long CTA tier (distinct partners are
    counted while inserting); with a forced tiny capacity every such cell is handed to the merge kernel's list from
    inside the kernel.  All three routes (long tier, hand-over, CB200_SNN_CTA=0 = merge kernel directly) emit the
    oracle's edge list."""
    knn = O.find_knn(synth.make_embedding(2600, 8, seed=23), 180)      # flat sequences of up to ~30 000 entries
    want = O.snn_edges(knn, "rank", j_begin=2300, j_end=2600)
    a = gpu_ctx.snn(knn, "rank", j_begin=2300, j_end=2600)
    assert all(np.array_equal(x, y) for x, y in zip(a, want))
    monkeypatch.setenv("CB200_SNN_LONG_CAP", "200")
    b = gpu_ctx.snn(knn, "rank", j_begin=2300, j_end=2600)
    assert all(np.array_equal(x, y) for x, y in zip(b, want))
    monkeypatch.delenv("CB200_SNN_LONG_CAP")
    monkeypatch.setenv("CB200_SNN_CTA", "0")
    c = gpu_ctx.snn(knn, "number", j_begin=2300, j_end=2600)
    monkeypatch.delenv("CB200_SNN_CTA")
    d = gpu_ctx.snn(knn, "number", j_begin=2300, j_end=2600)
    assert all(np.array_equal(x, y) for x, y in zip(c, d))
    assert all(np.array_equal(x, y) for x, y in zip(d, O.snn_edges(knn, "number", j_begin=2300, j_end=2600)))
