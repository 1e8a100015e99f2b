#!/usr/bin/env python
"""Compares a dump written by baseline/run_reference.R (the real reference, on an R host) with the CPU oracle
and -- when a B200 is present -- with libcellula_b200.so, stage by stage, at the tolerances of BASELINE.json's
north_star.  This is the step that turns "parity unpinned" into pinned parity: run it once where R exists and
commit the (small) dump of cfg1 under tests/golden/ref_dump_cfg1/ -- tests/test_reference_dump.py picks it up.

    python scripts/compare_reference_dump.py --data data/cfg1 --dump dump/cfg1 [--variant B] [--gpu]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_dump(dump, variant, n, ndims_hint=None):
    def vec(name, dtype):
        return np.fromfile(os.path.join(dump, f"{variant}_{name}"), dtype=dtype)
    rep = json.load(open(os.path.join(dump, "report.json")))
    nd, k = rep["ndims"], rep["k"]
    out = dict(report=rep, sf=vec("sf.f64", "<f8"), logx=vec("logcounts_x.f64", "<f8"), mean=vec("gene_mean.f64", "<f8"),
               var=vec("gene_total.f64", "<f8"), tech=vec("gene_tech.f64", "<f8"), bio=vec("gene_bio.f64", "<f8"),
               hvgs=vec("hvg_idx.i32", "<i4"), var_explained=vec("pca_var_explained.f64", "<f8"),
               percent_var=vec("pca_percent_var.f64", "<f8"))
    out["scores"] = vec("pca_scores.f64", "<f8").reshape((n, nd), order="F")
    out["rotation"] = vec("pca_rotation.f64", "<f8").reshape((-1, nd), order="F")
    out["knn"] = vec("knn_index.i32", "<i4").reshape((n, k), order="F")
    out["edges"] = (vec("snn_from.i32", "<i4"), vec("snn_to.i32", "<i4"), vec("snn_weight.f64", "<f8"))
    return out


def subspace_angle(a, b):
    qa, _ = np.linalg.qr(a)
    qb, _ = np.linalg.qr(b)
    return float(np.arccos(np.clip(np.linalg.svd(qa.T @ qb, compute_uv=False).min(), -1.0, 1.0)))


def score_error(got, want):
    sign = np.sign(np.sum(got * want, axis=0))
    return float((np.abs(got * sign - want).max(axis=0) / np.abs(want).max(axis=0)).max())


def compare(name, ours, ref, d, library_size_factors):
    """ours: dict with the same keys as the dump; returns a list of (check, value, bound, ok)."""
    from oracle import oracle as O
    rows = []

    def add(check, value, bound, ok=None):
        rows.append((f"{name}: {check}", value, bound, bool(value <= bound) if ok is None else bool(ok)))
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))  # noqa: E731
    if library_size_factors:
        add("size factors, max rel", rel(ours["sf"], ref["sf"]), 1e-6)
    add("logcounts @x, max rel", rel(ours["logx"], ref["logx"]), 1e-6)
    nz = ref["mean"] > 0
    add("gene means, max rel", rel(ours["mean"][nz], ref["mean"][nz]), 1e-6)
    add("gene variances, max rel", rel(ours["var"][nz], ref["var"][nz]), 1e-6)
    same = np.array_equal(ours["hvgs"], ref["hvgs"])
    add("HVG ranks identical", 0 if same else int((ours["hvgs"] != ref["hvgs"]).sum()), 0, same)
    add("PCA subspace angle (rad)", subspace_angle(ours["rotation"], ref["rotation"]), 1e-4)
    add("PCA scores, sign-aligned max rel (by column max)", score_error(ours["scores"], ref["scores"]), 1e-4)
    # kNN / SNN: fed the REFERENCE's PCA matrix (north_star), so PCA rounding cannot leak in
    knn_same = np.array_equal(ours["knn_on_ref_pca"], ref["knn"])
    add("kNN on the reference PCA: rows differing", int((ours["knn_on_ref_pca"] != ref["knn"]).any(axis=1).sum()), 0, knn_same)
    a = O.canonical_edges(*ours["edges_on_ref_knn"])
    b = O.canonical_edges(*ref["edges"])
    e_same = all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b))
    add("SNN edges + weights (canonical order) identical", 0 if e_same else 1, 0, e_same)
    o_same = all(np.array_equal(x, y) for x, y in zip(ours["edges_on_ref_knn"], ref["edges"]))
    add("SNN emission order identical to as_edgelist()", 0 if o_same else 1, 0, o_same)
    return rows


def oracle_results(d, ref, ndims, k, scheme, sf_in):
    from oracle import oracle as O
    g, n = d["n_genes"], d["n_cells"]
    sf = O.library_size_factors(d["colptr"], d["x"], sf_in)
    logx = O.log_norm_counts(d["colptr"], d["x"], sf)
    mean, var = O.gene_mean_var(g, n, d["rowidx"], logx)
    hv = O.get_top_hvgs(O.model_gene_var(mean, var)["bio"], len(ref["hvgs"]))
    pca = O.run_pca(g, d["colptr"], d["rowidx"], logx, ref["hvgs"], ndims)      # the reference's HVGs: isolates the PCA
    knn = O.find_knn(ref["scores"][:, :ndims], k)
    return dict(sf=sf, logx=logx, mean=mean, var=var, hvgs=hv, rotation=pca["rotation"], scores=pca["scores"],
                knn_on_ref_pca=knn, edges_on_ref_knn=O.snn_edges(ref["knn"], scheme))


def gpu_results(d, ref, ndims, k, scheme, sf_in):
    import cellula_b200 as cb
    from cellula_b200 import trendvar
    ctx = cb.Context(0)
    ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
    sf = ctx.size_factors(sf_in)
    logx, mean, var = ctx.lognorm_genestats()
    hv = trendvar.get_top_hvgs(trendvar.model_gene_var(mean, var), len(ref["hvgs"]))
    pca = ctx.pca(ref["hvgs"], ndims)
    knn = ctx.knn(ref["scores"][:, :ndims], k)
    edges = ctx.snn(ref["knn"], scheme)
    ctx.close()
    return dict(sf=sf, logx=logx, mean=mean, var=var, hvgs=hv, rotation=pca["rotation"], scores=pca["scores"],
                knn_on_ref_pca=knn, edges_on_ref_knn=edges)


def compare_side_calls(dump, d, use_gpu):
    """SURVEY 8(f)-1 / 8(f)-4 dumps (F1_*, F4_*, check_lattice_dist): pooled size factors on the reference's own
    quickCluster labels, per-cell QC metrics, neighbour distances.  Skipped when the dump predates them."""
    from oracle import oracle as O
    rows = []
    path = lambda name: os.path.join(dump, name)  # noqa: E731
    if not os.path.exists(path("F1_pooled_sf.f64")):
        return rows
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))  # noqa: E731
    codes = np.fromfile(path("F1_quickcluster_codes.i32"), dtype="<i4")
    ref_sf = np.fromfile(path("F1_pooled_sf.f64"), dtype="<f8")
    ref_sf = ref_sf / ref_sf[ref_sf > 0].mean()            # both sides on the unit-mean scale logNormCounts would use
    subsets = [np.fromfile(path(f"F4_subset_{nm}_genes.i32"), dtype="<i4") for nm in ("first", "second", "third")]
    arms = [("oracle", lambda: O.pooled_size_factors(d["n_genes"], d["colptr"], d["rowidx"], d["x"], codes),
             lambda: O.per_cell_qc_metrics(d["n_genes"], d["colptr"], d["rowidx"], d["x"], subsets))]
    if use_gpu:
        import cellula_b200 as cb
        ctx = cb.Context(0)
        ctx.load_csc(d["n_genes"], d["colptr"], d["rowidx"], d["x"])
        arms.append(("cellula_b200", lambda: ctx.pooled_size_factors(codes)[0], lambda: ctx.qc_metrics(subsets)))
    for name, pooled, qc in arms:
        sf = pooled()
        v = rel(sf / sf[sf > 0].mean(), ref_sf)
        rows.append((f"{name}: pooled size factors on the reference's quickCluster labels, max rel", v, 1e-6, v <= 1e-6))
        q = qc()
        ok = (np.array_equal(q["sum"], np.fromfile(path("F4_qc_sum.f64"), dtype="<f8")) and
              np.array_equal(q["detected"], np.fromfile(path("F4_qc_detected.i32"), dtype="<i4")))
        for s_i, nm in enumerate(("first", "second", "third")):
            ok = ok and np.array_equal(q["subsets"][s_i]["sum"], np.fromfile(path(f"F4_qc_{nm}_sum.f64"), dtype="<f8"))
            ok = ok and np.array_equal(q["subsets"][s_i]["detected"], np.fromfile(path(f"F4_qc_{nm}_detected.i32"), dtype="<i4"))
            pr = np.fromfile(path(f"F4_qc_{nm}_percent.f64"), dtype="<f8")
            ok = ok and np.allclose(q["subsets"][s_i]["percent"], pr, rtol=1e-14, atol=0, equal_nan=True)
        rows.append((f"{name}: perCellQCMetrics (sum, detected, 3 subsets) identical", 0 if ok else 1, 0, ok))
    if os.path.exists(path("check_lattice_dist.f64")):
        lat = np.fromfile(path("check_lattice_x.f64"), dtype="<f8").reshape((-1, 3), order="F")
        idx = np.fromfile(path("check_lattice_knn.i32"), dtype="<i4").reshape((lat.shape[0], -1), order="F")
        dist = np.fromfile(path("check_lattice_dist.f64"), dtype="<f8").reshape(idx.shape, order="F")
        same = np.array_equal(O.knn_distances(lat, idx), dist)
        rows.append(("oracle: findKNN(get.distance = TRUE) distances on the tie lattice identical", 0 if same else 1, 0, same))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True)
    ap.add_argument("--dump", required=True)
    ap.add_argument("--variant", default="B", choices=["A", "B"],
                    help="B: library-size factors (like for like); A: the reference's deconvolution factors are fed in")
    ap.add_argument("--gpu", action="store_true")
    a = ap.parse_args()
    from cellula_b200 import synth
    d, _ = synth.read_dataset(a.data)
    ref = load_dump(a.dump, a.variant, d["n_cells"])
    rep = ref["report"]
    sf_in = None if a.variant == "B" else ref["sf"]
    rows = compare("oracle", oracle_results(d, ref, rep["ndims"], rep["k"], rep["scheme"], sf_in), ref, d, a.variant == "B")
    if a.gpu:
        rows += compare("cellula_b200", gpu_results(d, ref, rep["ndims"], rep["k"], rep["scheme"], sf_in), ref, d,
                        a.variant == "B")
    rows += compare_side_calls(a.dump, d, a.gpu)
    width = max(len(r[0]) for r in rows)
    for check, value, bound, ok in rows:
        print(f"{'ok  ' if ok else 'FAIL'} {check:<{width}}  {value:.3e}  (bound {bound:g})")
    sys.exit(0 if all(r[3] for r in rows) else 1)


if __name__ == "__main__":
    main()
