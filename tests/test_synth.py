"""The ONE synthetic-data generator (csrc/cb200_synth_core.h), host compilation: determinism, shard
independence, the on-disk dataset form that baseline/run_reference.R reads, and golden digests that pin the
specification (a change of the generator changes every benchmark input, so it must be deliberate)."""
import hashlib
import json
import os

import numpy as np

from cellula_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "synth_digests.json")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_host_generator_is_deterministic_and_thread_independent():
    a = synth.make_counts(600, 3000, nnz_per_cell=250, seed=123, n_threads=1)
    b = synth.make_counts(600, 3000, nnz_per_cell=250, seed=123, n_threads=5)
    for k in ("colptr", "rowidx", "x"):
        assert np.array_equal(a[k], b[k])
    lens = np.diff(a["colptr"])
    assert lens.min() >= 1 and 200 < lens.mean() < 300
    assert np.all(a["x"] >= 1) and np.all(a["x"] == np.round(a["x"]))
    assert a["rowidx"].min() >= 0 and a["rowidx"].max() < 3000
    for c in (0, 17, 599):                                    # rows sorted within a column (dgCMatrix invariant)
        assert np.all(np.diff(a["rowidx"][a["colptr"][c]:a["colptr"][c + 1]]) > 0)


def test_any_block_of_cells_tiles_the_same_matrix():
    full = synth.make_counts(900, 2000, nnz_per_cell=150, seed=9)
    part = synth.make_counts(300, 2000, nnz_per_cell=150, seed=9, cell_offset=400, n_cells_total=900)
    lo, hi = full["colptr"][400], full["colptr"][700]
    assert np.array_equal(part["rowidx"], full["rowidx"][lo:hi]) and np.array_equal(part["x"], full["x"][lo:hi])
    assert np.array_equal(part["colptr"], full["colptr"][400:701] - lo)


def test_golden_digests_pin_the_generator_specification():
    want = json.load(open(GOLDEN))
    d = synth.make_counts(500, 2000, nnz_per_cell=200, seed=synth.SEED)
    assert digest(d["colptr"], d["rowidx"], d["x"]) == want["counts_500x2000_nnz200"]
    e = synth.make_embedding(1000, 8, seed=synth.EMB_SEED)
    assert digest(e) == want["embedding_1000x8"]


def test_dataset_roundtrip(tmp_path):
    d = synth.make_counts(200, 1000, nnz_per_cell=100, seed=4)
    man = synth.write_dataset(str(tmp_path / "ds"), d, seed=4, nnz_per_cell=100)
    assert man["nnz"] == d["x"].size and set(man["sha256"]) == {"p.i64", "i.i32", "x.f64"}
    back, man2 = synth.read_dataset(str(tmp_path / "ds"))
    for k in ("colptr", "rowidx", "x"):
        assert np.array_equal(back[k], d[k])
    assert man2["n_genes"] == 1000 and man2["n_cells"] == 200
    with open(tmp_path / "ds" / "x.f64", "r+b") as fh:        # a corrupted file is detected
        fh.seek(8)
        fh.write(b"\x01")
    try:
        synth.read_dataset(str(tmp_path / "ds"))
    except ValueError as e:
        assert "sha256" in str(e)
    else:
        raise AssertionError("corruption not detected")


def test_embedding_shape_and_scales():
    e = synth.make_embedding(20000, 50)
    assert e.shape == (20000, 50) and e.flags["F_CONTIGUOUS"]
    sd = e.std(axis=0)
    assert sd[0] > 10 * sd[-1] and np.all(np.diff(sd[::7]) < 0)      # per-PC scale decays 20 -> 1
    assert len(np.unique(e[:, 0])) == 20000                          # continuous: no duplicate cells
