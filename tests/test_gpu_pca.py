"""GPU parity, stage K5 (PCA on the HVG subset) against the oracle's exact SVD.
north_star tolerance: largest principal angle <= 1e-4 rad, sign-aligned scores within 1e-4
(relative to the column's max-abs value)."""
import numpy as np
import pytest

import cellula_b200
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def principal_angle(a, b):
    qa, _ = np.linalg.qr(a)
    qb, _ = np.linalg.qr(b)
    s = np.linalg.svd(qa.T @ qb, compute_uv=False)
    # sin of the largest angle, computed without cancellation
    return float(np.linalg.norm(qb - qa @ (qa.T @ qb), 2)), float(s.min())


def run_to_pca(ctx, d, hvg, ndims):
    ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    ctx.size_factors()
    ctx.lognorm_genestats(want_logx=False, want_stats=False)
    return ctx.pca(hvg, ndims)


@pytest.mark.parametrize("ndims", [5, 20])
def test_pca_matches_oracle_svd(gpu_ctx, small_counts, small_oracle, ndims):
    d, ref = small_counts, small_oracle
    hvg = ref["hvgs"]
    got = run_to_pca(gpu_ctx, d, hvg, ndims)
    want = O.run_pca(d["n_genes"], d["colptr"], d["rowidx"], ref["logx"], hvg, ndims)
    # the default cross-product engine rounds its inputs to 2^-20 (tensor-core int8 digits): ~1e-8
    assert np.allclose(got["var_explained"], want["var_explained"], rtol=1e-6)
    assert abs(got["total_var"] - want["total_var"]) <= 1e-9 * want["total_var"]
    assert np.allclose(got["percent_var"], want["percent_var"], rtol=1e-6)
    sin_max, _ = principal_angle(got["rotation"], want["rotation"])
    assert sin_max <= 1e-4, sin_max
    assert np.allclose(got["rotation"].T @ got["rotation"], np.eye(ndims), atol=1e-9)
    sign = np.sign(np.sum(got["scores"] * want["scores"], axis=0))
    err = np.abs(got["scores"] * sign - want["scores"]).max(axis=0) / np.abs(want["scores"]).max(axis=0)
    assert err.max() <= 1e-4, err
    # scores are exactly the projection of the centred data on the returned rotation
    assert np.allclose(got["scores"].var(axis=0, ddof=1), got["var_explained"], rtol=1e-6)
    gpu_ctx.timings()["eig_iterations"] > 0


def test_gram_engines_agree(gpu_ctx, small_counts, small_oracle, monkeypatch):
    """Three cross-product engines -- tcgen05 int8 digit products (default), fp64 shared-memory
    tiles, fp64 L2 reductions -- must give the same PCA: the fp64 ones to rounding, the tensor-core
    one to the 2^-20 input quantisation (far inside the 1e-4 contract)."""
    res, eng = {}, {}
    for name in ("tc", "tiled", "atomic"):
        monkeypatch.setenv("CB200_GRAM_ENGINE", name)
        res[name] = run_to_pca(gpu_ctx, small_counts, small_oracle["hvgs"], 8)
        eng[name] = gpu_ctx.timings()["gram_engine"]
    assert eng == {"tc": 3, "tiled": 2, "atomic": 1}
    monkeypatch.delenv("CB200_GRAM_ENGINE")
    run_to_pca(gpu_ctx, small_counts, small_oracle["hvgs"], 8)
    assert gpu_ctx.timings()["gram_engine"] == 3            # the default takes the tensor cores
    a, b, c = res["tiled"], res["atomic"], res["tc"]
    assert np.allclose(a["var_explained"], b["var_explained"], rtol=1e-10)
    assert np.allclose(a["rotation"], b["rotation"], atol=1e-8)
    assert np.allclose(a["scores"], b["scores"], atol=1e-7 * np.abs(a["scores"]).max())
    assert np.allclose(a["var_explained"], c["var_explained"], rtol=1e-6)
    assert np.allclose(a["rotation"], c["rotation"], atol=2e-5)
    assert np.allclose(a["scores"], c["scores"], atol=2e-5 * np.abs(a["scores"]).max())


def test_pca_sign_convention_is_deterministic(gpu_ctx, small_counts, small_oracle):
    a = run_to_pca(gpu_ctx, small_counts, small_oracle["hvgs"], 6)
    b = run_to_pca(gpu_ctx, small_counts, small_oracle["hvgs"], 6)
    # the cross-product is accumulated with fp64 reductions in L2 (order varies run to run), so
    # two runs agree to rounding, and the sign rule makes them agree in sign as well
    assert np.allclose(a["rotation"], b["rotation"], rtol=0, atol=1e-9)
    j = np.argmax(np.abs(a["rotation"]), axis=0)
    assert np.all(a["rotation"][j, np.arange(6)] > 0)


def test_pca_hvg_order_only_permutes_rotation_rows(gpu_ctx, small_counts, small_oracle):
    hvg = small_oracle["hvgs"][:300]
    perm = np.random.default_rng(1).permutation(len(hvg))
    a = run_to_pca(gpu_ctx, small_counts, hvg, 5)
    b = run_to_pca(gpu_ctx, small_counts, hvg[perm], 5)
    assert np.allclose(np.abs(a["scores"]), np.abs(b["scores"]), atol=1e-7 * np.abs(a["scores"]).max())
    assert np.allclose(np.abs(a["rotation"][perm]), np.abs(b["rotation"]), atol=1e-7)


def test_pca_argument_errors(gpu_ctx, small_counts, small_oracle):
    d = small_counts
    gpu_ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    gpu_ctx.size_factors()
    with pytest.raises(cellula_b200.Cb200Error, match="must be computed first"):
        gpu_ctx.pca(small_oracle["hvgs"], 5)
    gpu_ctx.lognorm_genestats(want_logx=False, want_stats=False)
    with pytest.raises(cellula_b200.Cb200Error, match="twice"):
        gpu_ctx.pca(np.array([1, 2, 2, 3], dtype=np.int32), 2)
    with pytest.raises(cellula_b200.Cb200Error, match="out of range"):
        gpu_ctx.pca(np.array([1, d["n_genes"]], dtype=np.int32), 1)
    with pytest.raises(cellula_b200.Cb200Error, match="exceeds"):
        gpu_ctx.pca(np.arange(10, dtype=np.int32), 11)
