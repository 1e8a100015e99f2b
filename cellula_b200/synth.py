"""Synthetic scRNA-seq count matrices and embeddings -- host side of THE generator of this repository.

One specification (csrc/cb200_synth_core.h: integer hashing + individually rounded IEEE operations), compiled
for the device (``Context.synth_counts`` -> ``cb200_synth_counts``, what the bench uses at 1.3 M - 4 M cells)
and for the host (``make_counts`` -> ``cb200_synth_host``, no GPU needed).  Both give the same matrix bit for
bit for the same (seed, shape, nnz_per_cell), and any contiguous block of cells can be generated on its own.
``write_dataset`` stores it in the language-neutral on-disk form that
``baseline/run_reference.R`` reads on an R host.

Model: negative-binomial counts with planted low-rank structure in the dgCMatrix/CSC layout the reference's
``counts(sce)`` assay uses: log lambda_gc = base_g + off + log depth_c + sum_z W_gz Z_fc,
64 latent factors of geometrically decaying strength (clear spectral gap so the PCA comparison is
well-posed), 40 cluster centroids plus continuous noise (no duplicate cells), dispersion 0.1.
"""
from __future__ import annotations

import ctypes
import hashlib
import json
import os

import numpy as np

from . import _lib

SEED = 20261019
EMB_SEED = 20261020


def make_counts(n_cells, n_genes, nnz_per_cell=500, seed=SEED, cell_offset=0, n_cells_total=None, n_threads=0):
    """Returns dict(colptr int64[N+1], rowidx int32[nnz], x float64[nnz], n_genes, n_cells) for the global
    cells [cell_offset, cell_offset + n_cells) of the (n_genes x n_cells_total) matrix."""
    L = _lib.lib()
    if n_cells_total is None:
        n_cells_total = cell_offset + n_cells
    nnz = _lib.c_i64()
    pc, pr, px = _lib.c_i64p(), _lib.c_i32p(), _lib.c_f64p()
    rc = L.cb200_synth_host(int(n_genes), int(n_cells), float(nnz_per_cell), int(seed), int(cell_offset),
                            int(n_cells_total), int(n_threads), ctypes.byref(nnz), ctypes.byref(pc), ctypes.byref(pr),
                            ctypes.byref(px))
    if rc != 0:
        raise ValueError(f"cb200_synth_host failed ({rc}): need n_genes >= 64, 1 <= nnz_per_cell < n_genes")
    m = nnz.value
    try:
        colptr = np.ctypeslib.as_array(pc, shape=(n_cells + 1,)).copy()
        rowidx = np.ctypeslib.as_array(pr, shape=(max(m, 1),))[:m].copy()
        x = np.ctypeslib.as_array(px, shape=(max(m, 1),))[:m].copy()
    finally:
        L.cb200_synth_host_free(pc, pr, px)
    return dict(colptr=colptr, rowidx=rowidx, x=x, n_genes=int(n_genes), n_cells=int(n_cells))


def make_embedding(n_cells, d=50, seed=EMB_SEED, n_threads=0):
    """cfg-4-style PC matrix: 40 Gaussian clusters, per-PC scale decaying 20 -> 1 (N x d fp64, column-major)."""
    emb = np.empty((n_cells, d), dtype=np.float64, order="F")
    rc = _lib.lib().cb200_synth_embedding_host(int(n_cells), int(d), int(seed), int(n_threads),
                                               ctypes.c_void_p(emb.ctypes.data))
    if rc != 0:
        raise ValueError("cb200_synth_embedding_host failed")
    return emb


def _sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def write_dataset(path, d, **meta):
    """On-disk form of a count matrix (little-endian raw binaries + manifest), readable from base R with
    readBin -> new("dgCMatrix", i=, p=as.integer(p), x=, Dim=): p.i64, i.i32, x.f64, manifest.json."""
    os.makedirs(path, exist_ok=True)
    arrays = {"p.i64": np.asarray(d["colptr"], dtype="<i8"), "i.i32": np.asarray(d["rowidx"], dtype="<i4"),
              "x.f64": np.asarray(d["x"], dtype="<f8")}
    for name, a in arrays.items():
        a.tofile(os.path.join(path, name))
    man = dict(format="cellula_b200 dataset v1", n_genes=int(d["n_genes"]), n_cells=int(d["n_cells"]),
               nnz=int(arrays["x.f64"].size), sha256={k: _sha256(v) for k, v in arrays.items()},
               generator="cellula_b200/csrc/cb200_synth_core.h", **meta)
    with open(os.path.join(path, "manifest.json"), "w") as fh:
        json.dump(man, fh, indent=1)
    return man


def read_dataset(path, verify=True):
    man = json.load(open(os.path.join(path, "manifest.json")))
    d = dict(colptr=np.fromfile(os.path.join(path, "p.i64"), dtype="<i8"),
             rowidx=np.fromfile(os.path.join(path, "i.i32"), dtype="<i4"),
             x=np.fromfile(os.path.join(path, "x.f64"), dtype="<f8"), n_genes=man["n_genes"], n_cells=man["n_cells"])
    if verify:
        for k, a in (("p.i64", d["colptr"]), ("i.i32", d["rowidx"]), ("x.f64", d["x"])):
            if _sha256(a) != man["sha256"][k]:
                raise ValueError(f"{path}/{k}: sha256 does not match the manifest")
    return d, man
