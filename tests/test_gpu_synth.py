"""The device compilation of the generator (cb200_synth_counts) produces bit for bit the matrix of the host
compilation (cb200_synth_host): one specification, csrc/cb200_synth_core.h.  This is what lets the bench
(device-generated, 1.3 M cells), the parity tests, the CPU baseline legs and the on-disk dataset for
baseline/run_reference.R all see identical synthetic counts."""
import numpy as np
import pytest

from cellula_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,g,nnz,seed", [(3000, 5000, 250.0, 123), (2700, 32738, 850.0, synth.SEED),
                                          (5000, 27998, 1500.0, synth.SEED), (64, 64, 10.0, 1)])
def test_device_generator_equals_host_generator(gpu_ctx, n, g, nnz, seed):
    got_nnz = gpu_ctx.synth_counts(g, n, nnz, seed=seed)
    colptr, rowidx, x = gpu_ctx.export_csc()
    want = synth.make_counts(n, g, nnz_per_cell=nnz, seed=seed)
    assert got_nnz == want["x"].size
    assert np.array_equal(colptr, want["colptr"])
    assert np.array_equal(rowidx, want["rowidx"])
    assert np.array_equal(x, want["x"])


def test_device_shard_equals_host_shard_of_a_large_matrix(gpu_ctx):
    """Cells [1 000 000, 1 000 400) of the cfg3 matrix (1 306 127 x 27 998): the calibration probes the first
    cells of the GLOBAL matrix on both sides, so shards generated anywhere tile one matrix."""
    n_total, g = 1306127, 27998
    gpu_ctx.synth_counts(g, 400, 1500.0, seed=synth.SEED, cell_offset=1000000, n_cells_total=n_total)
    colptr, rowidx, x = gpu_ctx.export_csc()
    want = synth.make_counts(400, g, nnz_per_cell=1500.0, seed=synth.SEED, cell_offset=1000000, n_cells_total=n_total)
    assert np.array_equal(colptr, want["colptr"]) and np.array_equal(rowidx, want["rowidx"])
    assert np.array_equal(x, want["x"])
    assert 1300 < np.diff(colptr).mean() < 1700
