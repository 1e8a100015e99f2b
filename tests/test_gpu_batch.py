"""GPU parity of the batch-aware branch (R/doNormAndReduce.R:74-76, :84-86): per-batch averages for
batchelor::multiBatchNorm, factors taken as they are, per-block mean / variance for scran::modelGeneVar(block=),
and the Python mirror of doNormAndReduce(batch = ...) end to end -- all against the oracle restatement."""
import numpy as np
import pytest

import cellula_b200 as cb
from cellula_b200 import synth, trendvar
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def batched():
    d = synth.make_counts(3000, 4000, nnz_per_cell=300, seed=21)
    rng = np.random.default_rng(2)
    block = rng.choice(3, d["n_cells"], p=[0.5, 0.3, 0.2]).astype(np.int32)
    # batch 1 is sequenced 3x deeper: multiBatchNorm has something to rescale
    x = d["x"].copy()
    col_of = np.repeat(np.arange(d["n_cells"]), np.diff(d["colptr"]))
    x[block[col_of] == 1] *= 3.0
    d = dict(d, x=x)
    return d, block


def test_block_averages_and_rescaled_factors(gpu_ctx, batched):
    d, block = batched
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    sf = gpu_ctx.size_factors()
    ave, sfm, nb = gpu_ctx.block_averages(block, 3)
    want = O.batch_averages(d["n_genes"], d["colptr"], d["rowidx"], d["x"], sf, block, 3)
    assert np.array_equal(nb, np.bincount(block, minlength=3))
    assert np.allclose(sfm, [sf[block == b].mean() for b in range(3)], rtol=1e-13)
    assert np.allclose(ave, want, rtol=1e-9, atol=1e-12)      # fixed-point sums: 2^-39 absolute per term at this size
    got = trendvar.rescale_size_factors(ave, sf, block)
    ref = O.rescale_size_factors(want, sf, block)
    assert np.allclose(got, ref, rtol=1e-10)
    assert got[block == 1].mean() > 2.0 * got[block == 0].mean()      # the deeper batch is scaled down


def test_blocked_gene_statistics(gpu_ctx, batched):
    d, block = batched
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    sf = gpu_ctx.size_factors()
    gpu_ctx.set_blocks(block, 3)
    logx, mean, var = gpu_ctx.lognorm_genestats_blocked()
    o_logx = O.log_norm_counts(d["colptr"], d["x"], sf)
    assert np.allclose(logx, o_logx, rtol=1e-12, atol=0)
    om, ov = O.gene_mean_var_blocked(d["n_genes"], d["colptr"], d["rowidx"], o_logx, block, 3)
    assert np.allclose(mean, om, rtol=1e-11, atol=1e-300)
    assert np.allclose(var, ov, rtol=1e-9, atol=1e-18)
    # a block without cells / with one cell: NaN like scran
    block2 = block.copy()
    block2[block2 == 2] = 0
    block2[0] = 2
    gpu_ctx.set_blocks(block2, 4)
    _, mean2, var2 = gpu_ctx.lognorm_genestats_blocked(want_logx=False)
    assert np.all(np.isnan(mean2[3])) and np.all(np.isnan(var2[3])) and np.all(np.isnan(var2[2]))
    assert not np.any(np.isnan(mean2[2]))


def test_given_size_factors_are_not_recentred(gpu_ctx, batched):
    d, _ = batched
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    user = np.random.default_rng(0).uniform(0.3, 4.0, d["n_cells"])
    gpu_ctx.size_factors_given(user)
    logx, _, _ = gpu_ctx.lognorm_genestats()
    assert np.allclose(logx, O.log_norm_counts(d["colptr"], d["x"], user), rtol=1e-12, atol=0)
    with pytest.raises(cb.Cb200Error, match="size factors should be positive"):
        gpu_ctx.size_factors_given(np.where(np.arange(d["n_cells"]) == 5, 0.0, user))


def test_do_norm_and_reduce_with_batch_against_oracle(batched):
    d, block = batched
    cb.options["skip_umap"] = True
    sce = cb.SingleCellExperiment.from_counts(d["colptr"], d["rowidx"], d["x"], d["n_genes"])
    sce.colData["batch"] = np.array(["a", "b", "c"])[block]
    sce = cb.doNormAndReduce(sce, batch="batch", ndims=15, hvg_ntop=600, verbose=False)
    # oracle: library-size factors -> multiBatchNorm -> blocked modelGeneVar -> PCA
    sf0 = O.library_size_factors(d["colptr"], d["x"])
    ave = O.batch_averages(d["n_genes"], d["colptr"], d["rowidx"], d["x"], sf0, block, 3)
    sf = O.rescale_size_factors(ave, sf0, block)
    logx = O.log_norm_counts(d["colptr"], d["x"], sf)
    mb, vb = O.gene_mean_var_blocked(d["n_genes"], d["colptr"], d["rowidx"], logx, block, 3)
    stats = O.model_gene_var_blocked(mb, vb, np.bincount(block))
    hv = O.get_top_hvgs(stats["bio"], 600)
    assert np.allclose(sce.colData["sizeFactor"], sf, rtol=1e-10)
    assert np.allclose(sce.assays["logcounts"].x, logx, rtol=1e-9)
    assert np.array_equal(np.asarray(sce.metadata["hvgs"]), hv)
    ref = O.run_pca(d["n_genes"], d["colptr"], d["rowidx"], logx, hv, 15)
    got = np.asarray(sce.reducedDims["PCA"])
    sign = np.sign(np.sum(got * ref["scores"], axis=0))
    assert (np.abs(got * sign - ref["scores"]).max(axis=0) / np.abs(ref["scores"]).max(axis=0)).max() <= 1e-4
