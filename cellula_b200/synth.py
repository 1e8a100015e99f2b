"""Synthetic scRNA-seq count matrices and embeddings (host side, numpy).

Deterministic negative-binomial counts with planted low-rank structure, in the
dgCMatrix/CSC layout the reference's ``counts(sce)`` assay uses (SURVEY.md row a1,
generator spec SURVEY.md 8d): log lambda_gc = log m_g + sum_f W_gf Z_fc + log s_c,
F latent factors of geometrically decaying strength (clear spectral gap so the PCA
comparison is well-posed), 40 cluster centroids plus continuous noise (no duplicate
cells), Gamma-Poisson counts with dispersion 0.1.

The bench generates its (much larger) matrices on the device with
``cb200_synth_counts`` (csrc/cb200_synth.cu); this numpy version feeds the parity tests,
where oracle and CUDA path must see the same host arrays.
"""
from __future__ import annotations

import numpy as np

SEED = 20261019


def make_counts(n_cells, n_genes, nnz_per_cell=500, seed=SEED, n_factors=64, n_clusters=40,
                chunk=4096, dispersion=0.1):
    """Returns dict(colptr int64[N+1], rowidx int32[nnz], x float64[nnz], n_genes, n_cells)."""
    rng = np.random.default_rng(seed)
    F = n_factors
    sigma = 1.0 * 0.93 ** np.arange(F)
    W = np.where(rng.random((n_genes, F)) < 0.05, rng.standard_normal((n_genes, F)), 0.0) * sigma
    cent = rng.standard_normal((n_clusters, F))
    base = rng.standard_normal(n_genes) * 2.0
    # calibrate the base-mean offset so that the neutral cell has ~nnz_per_cell non-zeros
    r = 1.0 / dispersion
    lo, hi = -20.0, 10.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        lam = np.exp(base + mid)
        nnz = np.sum(1.0 - (r / (r + lam)) ** r)
        lo, hi = (mid, hi) if nnz < nnz_per_cell else (lo, mid)
    base = base + 0.5 * (lo + hi)

    cols, rows, vals = [0], [], []
    total = 0
    for c0 in range(0, n_cells, chunk):
        c1 = min(n_cells, c0 + chunk)
        crng = np.random.default_rng([seed, c0])
        nc = c1 - c0
        cl = crng.integers(0, n_clusters, nc)
        Z = cent[cl] + 0.3 * crng.standard_normal((nc, F))
        depth = np.exp(0.35 * crng.standard_normal(nc))
        loglam = Z @ W.T + base[None, :] + np.log(depth)[:, None]
        lam = np.exp(loglam)
        gam = crng.gamma(shape=r, scale=1.0 / r, size=lam.shape)
        cnt = crng.poisson(lam * gam)
        for q in range(nc):  # re-draw empty cells (logNormCounts would error on them)
            while cnt[q].sum() == 0:
                cnt[q] = crng.poisson(lam[q] * crng.gamma(shape=r, scale=1.0 / r, size=n_genes))
        cc, gg = np.nonzero(cnt)
        rows.append(gg.astype(np.int32))
        vals.append(cnt[cc, gg].astype(np.float64))
        per = np.bincount(cc, minlength=nc)
        cols.extend((total + np.cumsum(per)).tolist())
        total += per.sum()
    return dict(colptr=np.asarray(cols, dtype=np.int64), rowidx=np.concatenate(rows),
                x=np.concatenate(vals), n_genes=int(n_genes), n_cells=int(n_cells))


def make_embedding(n_cells, d=50, seed=SEED + 1, n_clusters=40):
    """cfg-4-style PC matrix: Gaussian clusters, per-PC scale decaying 20 -> 1 (N x d fp64)."""
    rng = np.random.default_rng(seed)
    scale = np.geomspace(20.0, 1.0, d)
    cent = rng.standard_normal((n_clusters, d)) * scale
    cl = rng.integers(0, n_clusters, n_cells)
    return cent[cl] + rng.standard_normal((n_cells, d)) * (0.35 * scale)
