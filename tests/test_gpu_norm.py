"""GPU parity, stages K1-K3 (size factors, logcounts, per-gene mean/variance) through the C ABI
against the oracle.  north_star tolerance: 1e-6 relative; asserted much tighter."""
import numpy as np
import pytest

import cellula_b200
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def run_norm(ctx, d, colptr_dtype=np.int64):
    ctx.load_csc(d["n_genes"], d["colptr"].astype(colptr_dtype), d["rowidx"], d["x"])
    sf = ctx.size_factors()
    logx, mean, var = ctx.lognorm_genestats()
    return sf, logx, mean, var


def check_against_oracle(d, sf, logx, mean, var):
    o_sf = O.library_size_factors(d["colptr"], d["x"])
    o_logx = O.log_norm_counts(d["colptr"], d["x"], o_sf)
    o_mean, o_var = O.gene_mean_var(d["n_genes"], len(d["colptr"]) - 1, d["rowidx"], o_logx)
    assert rel(sf, o_sf) <= 1e-13
    assert rel(logx, o_logx) <= 1e-12          # contract: 1e-6
    assert np.allclose(mean, o_mean, rtol=1e-11, atol=1e-300)   # contract: 1e-6
    assert np.allclose(var, o_var, rtol=1e-9, atol=1e-18)
    assert np.all(var >= 0)


def test_norm_small(gpu_ctx, small_counts):
    check_against_oracle(small_counts, *run_norm(gpu_ctx, small_counts))


def test_norm_int32_colptr_like_dgcmatrix(gpu_ctx, small_counts):
    check_against_oracle(small_counts, *run_norm(gpu_ctx, small_counts, np.int32))


def test_norm_kat(gpu_ctx):
    # KAT-NORM-1 / KAT-VAR-1 of SURVEY.md 8c.3: column sums 2, 4, 6; count 3 in the first column
    colptr = np.array([0, 2, 3, 5], dtype=np.int64)
    rowidx = np.array([0, 1, 0, 0, 1], dtype=np.int32)
    x = np.array([3.0, -1.0 + 0.0, 4.0, 5.0, 1.0])
    x[1] = -1.0  # (3 + -1 = 2): values need not be counts for the arithmetic
    gpu_ctx.load_csc(2, colptr, rowidx, x)
    sf = gpu_ctx.size_factors()
    assert np.allclose(sf, [0.5, 1.0, 1.5], rtol=0, atol=1e-16)
    gpu_ctx.lognorm_values()                    # K2 alone: the -1 makes one logcount NaN, which the gene sums reject
    logx = gpu_ctx.fetch_logcounts()
    assert abs(logx[0] - 2.807354922057604) < 1e-14
    with pytest.raises(cellula_b200.Cb200Error, match="fixed-point range"):
        gpu_ctx.lognorm_genestats()


def test_ragged_columns_and_alignment(gpu_ctx):
    """Column lengths 1..70 at every start parity: exercises the 128-bit head/tail handling."""
    rng = np.random.default_rng(3)
    lens = np.concatenate([np.arange(1, 71), rng.integers(1, 9, 300), [257, 1, 2, 3, 511]])
    colptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    G = 600
    rowidx = np.concatenate([np.sort(rng.choice(G, n, replace=False)) for n in lens]).astype(np.int32)
    x = rng.integers(1, 50, colptr[-1]).astype(np.float64)
    d = dict(n_genes=G, colptr=colptr, rowidx=rowidx, x=x)
    check_against_oracle(d, *run_norm(gpu_ctx, d))


def test_genes_without_counts_have_zero_stats(gpu_ctx, small_counts):
    d = dict(small_counts)
    d["n_genes"] = small_counts["n_genes"] + 37  # trailing all-zero genes
    sf, logx, mean, var = run_norm(gpu_ctx, d)
    assert np.all(mean[-37:] == 0) and np.all(var[-37:] == 0)
    check_against_oracle(d, sf, logx, mean, var)


def test_user_size_factors_are_recentred(gpu_ctx, small_counts):
    d = small_counts
    rng = np.random.default_rng(0)
    user = rng.uniform(0.2, 3.0, d["n_cells"])
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    sf = gpu_ctx.size_factors(user)
    assert np.allclose(sf, user / user.mean(), rtol=1e-14)
    logx, _, _ = gpu_ctx.lognorm_genestats()
    assert rel(logx, O.log_norm_counts(d["colptr"], d["x"], user / user.mean())) <= 1e-12


def test_empty_cell_fails_like_scuttle(gpu_ctx):
    colptr = np.array([0, 2, 2, 3], dtype=np.int64)  # second cell has no counts
    gpu_ctx.load_csc(4, colptr, np.array([0, 1, 2], dtype=np.int32), np.array([1.0, 2.0, 3.0]))
    with pytest.raises(cellula_b200.Cb200Error, match="size factors should be positive"):
        gpu_ctx.size_factors()


def test_stage_order_is_enforced(gpu_ctx, small_counts):
    d = small_counts
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    with pytest.raises(cellula_b200.Cb200Error, match="size factors have not been computed"):
        gpu_ctx.lognorm_genestats()


def test_values_first_split_equals_fused(gpu_ctx, small_counts):
    """cb200_lognorm_values (logcounts from values + column pointers, while the row indices still upload)
    followed by the gene sums must give bit-identical logcounts and statistics to the fused kernel's."""
    d = small_counts
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    gpu_ctx.size_factors()
    logx_f, mean_f, var_f = gpu_ctx.lognorm_genestats()
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    gpu_ctx.size_factors()
    gpu_ctx.lognorm_values()
    early = gpu_ctx.fetch_logcounts()                      # available before the gene sums
    logx_s, mean_s, var_s = gpu_ctx.lognorm_genestats()
    assert np.array_equal(early, logx_f) and np.array_equal(logx_s, logx_f)
    assert np.array_equal(mean_s, mean_f) and np.array_equal(var_s, var_f)   # integer sums: order-independent


def test_gene_sums_are_bit_reproducible_and_shard_independent(gpu_ctx, small_counts):
    """The per-gene sums are 64-bit fixed-point integers: two runs give identical bits, and the sums of two
    half-matrices (what two GPUs would all-reduce) add up to the sums of the whole matrix exactly."""
    import ctypes
    d = small_counts
    n = d["n_cells"]

    def acc_of(c0, c1):
        lo, hi = d["colptr"][c0], d["colptr"][c1]
        gpu_ctx.load_csc(d["n_genes"], d["colptr"][c0:c1 + 1] - lo, d["rowidx"][lo:hi], d["x"][lo:hi],
                         cell_offset=c0, n_cells_total=n)
        gpu_ctx.size_factors_local()
        gpu_ctx.size_factors_finish(float(d["x"].sum()), n, want=False)
        gpu_ctx.lognorm_genestats_local()
        gpu_ctx.synchronize()
        ptr, nbytes = gpu_ctx.device_ptr("gene_acc")
        import torch
        from cellula_b200.dist import device_tensor
        return device_tensor(ptr, (2 * d["n_genes"],), torch.int64, torch.device("cuda", 0)).cpu().numpy().copy()
    whole = acc_of(0, n)
    again = acc_of(0, n)
    assert np.array_equal(whole, again)
    assert np.array_equal(acc_of(0, 777) + acc_of(777, n), whole)
    assert whole[:d["n_genes"]].max() > 0


def test_malformed_row_indices_are_reported_not_written_out_of_bounds(gpu_ctx, small_counts):
    d = small_counts
    bad = d["rowidx"].copy()
    bad[12345] = d["n_genes"] + 5                        # outside [0, n_genes)
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], bad, d["x"])
    gpu_ctx.size_factors()
    with pytest.raises(cellula_b200.Cb200Error, match="row index is outside"):
        gpu_ctx.lognorm_genestats()
    bad = d["rowidx"].copy()
    bad[7] = -3
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], bad, d["x"])
    gpu_ctx.size_factors()
    with pytest.raises(cellula_b200.Cb200Error, match="row index is outside"):
        gpu_ctx.lognorm_genestats()


@pytest.mark.parametrize("n_genes", [9600, 9601, 20011, 32738])
def test_gene_slab_boundaries(gpu_ctx, n_genes):
    """1, 2, 3 and 4 gene slabs; columns that miss whole slabs; a segment longer than one stage (1024 entries)."""
    rng = np.random.default_rng(n_genes)
    lens = np.concatenate([rng.integers(1, 400, 500), [3000, 1, 2500]])
    cols = []
    for i, n in enumerate(lens):
        if i % 7 == 0:                                   # only low genes: the upper slabs get an empty segment
            cols.append(np.sort(rng.choice(min(n_genes, 5000), min(n, 5000), replace=False)))
        else:
            cols.append(np.sort(rng.choice(n_genes, n, replace=False)))
    lens = np.array([len(c) for c in cols])
    colptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rowidx = np.concatenate(cols).astype(np.int32)
    x = rng.integers(1, 60, colptr[-1]).astype(np.float64)
    dd = dict(n_genes=n_genes, colptr=colptr, rowidx=rowidx, x=x)
    check_against_oracle(dd, *run_norm(gpu_ctx, dd))


def test_table_and_direct_logcount_engines_agree_bit_for_bit(gpu_ctx, small_counts, monkeypatch):
    """The per-cell logcount tables (counts 1..32) + overflow list and the evaluate-everything-in-place kernel call
    the same device function: identical logcounts and identical (integer) gene sums."""
    a = run_norm(gpu_ctx, small_counts)
    monkeypatch.setenv("CB200_LOGNORM_ENGINE", "direct")
    b = run_norm(gpu_ctx, small_counts)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


def test_matrix_of_large_counts_falls_back_to_direct_evaluation(gpu_ctx):
    rng = np.random.default_rng(8)
    lens = rng.integers(50, 400, 800)
    colptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rowidx = np.concatenate([np.sort(rng.choice(3000, n, replace=False)) for n in lens]).astype(np.int32)
    x = rng.integers(100, 5000, colptr[-1]).astype(np.float64)           # nothing the 1..32 table covers
    d = dict(n_genes=3000, colptr=colptr, rowidx=rowidx, x=x)
    check_against_oracle(d, *run_norm(gpu_ctx, d))
