"""The parity-pinning tool chain (scripts/write_dataset.py form -> baseline/run_reference.R dump format ->
scripts/compare_reference_dump.py) exercised end to end on the CPU with a STAND-IN dump written by the oracle in the
exact file formats run_reference.R writes.  This checks the plumbing only (file names, layouts, 0- vs 1-based indices,
column-major matrices, tolerances wired to the right quantities); it pins nothing -- a real dump needs an R host
(tests/test_reference_dump.py picks one up when committed)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def _write(path, name, arr, dtype):
    np.asarray(arr).astype(dtype).tofile(os.path.join(path, name))


def test_compare_script_reads_what_run_reference_writes(tmp_path):
    import compare_reference_dump as C
    from cellula_b200 import synth
    from oracle import oracle as O
    data, dump = str(tmp_path / "data"), str(tmp_path / "dump")
    os.makedirs(dump)
    d0 = synth.make_counts(450, 1500, nnz_per_cell=200, seed=9)
    synth.write_dataset(data, d0, seed=9, workload="test")
    d, _ = synth.read_dataset(data)
    g, n, ndims, k, scheme, hvg = d["n_genes"], d["n_cells"], 8, 6, "jaccard", 300
    # --- variant B in run_reference.R's formats (dump_path(): *_sf.f64, *_logcounts_x.f64, *_gene_*.f64, *_hvg_idx.i32,
    #     *_pca_*.f64 column-major, *_knn_index.i32 column-major 1-based, *_snn_*.{i32,f64}), values from the oracle
    r = O.pipeline(g, d["colptr"], d["rowidx"], d["x"], ndims=ndims, hvg_ntop=hvg, neighbors=k, snn_type=scheme)
    stats = O.model_gene_var(r["mean"], r["var"])
    _write(dump, "B_sf.f64", r["sf"], "<f8")
    _write(dump, "B_logcounts_x.f64", r["logx"], "<f8")
    for col, key in (("mean", "mean"), ("total", "total"), ("tech", "tech"), ("bio", "bio")):
        _write(dump, f"B_gene_{col}.f64", stats[key], "<f8")
    _write(dump, "B_hvg_idx.i32", r["hvgs"], "<i4")
    pca = r["pca"]
    _write(dump, "B_pca_scores.f64", np.asfortranarray(pca["scores"]).ravel(order="F"), "<f8")
    _write(dump, "B_pca_rotation.f64", np.asfortranarray(pca["rotation"]).ravel(order="F"), "<f8")
    _write(dump, "B_pca_var_explained.f64", pca["var_explained"], "<f8")
    _write(dump, "B_pca_percent_var.f64", pca["percent_var"], "<f8")
    _write(dump, "B_knn_index.i32", np.asfortranarray(r["knn"]).ravel(order="F"), "<i4")
    f, t, w = r["edges"]
    _write(dump, "B_snn_from.i32", f, "<i4")
    _write(dump, "B_snn_to.i32", t, "<i4")
    _write(dump, "B_snn_weight.f64", w, "<f8")
    json.dump(dict(n_genes=g, n_cells=n, nnz=int(d["x"].size), ndims=ndims, k=k, scheme=scheme, hvg=hvg, workers=1,
                   skip_umap=True, irlba=False, timings_s={}), open(os.path.join(dump, "report.json"), "w"))
    ref = C.load_dump(dump, "B", n)
    assert ref["scores"].shape == (n, ndims) and ref["knn"].shape == (n, k) and ref["rotation"].shape == (hvg, ndims)
    rows = C.compare("oracle", C.oracle_results(d, ref, ndims, k, scheme, None), ref, d, True)
    assert rows and all(ok for _, _, _, ok in rows), [r_ for r_ in rows if not r_[3]]
    # a corrupted dump must be caught: shift one neighbour, scale one size factor
    ref_bad = dict(ref)
    ref_bad["knn"] = ref["knn"].copy()
    ref_bad["knn"][3, 0] = ref["knn"][3, 1]
    ref_bad["sf"] = ref["sf"].copy()
    ref_bad["sf"][7] *= 1.001
    rows = C.compare("oracle", C.oracle_results(d, ref_bad, ndims, k, scheme, None), ref_bad, d, True)
    failed = {name.split(": ")[1] for name, _, _, ok in rows if not ok}
    assert any("size factors" in x for x in failed) and any("kNN" in x for x in failed)

    # --- the F1 / F4 side dumps
    codes = (np.arange(n) % 2).astype(np.int32)
    _write(dump, "F1_quickcluster_codes.i32", codes, "<i4")
    _write(dump, "F1_pooled_sf.f64", 2.5 * O.pooled_size_factors(g, d["colptr"], d["rowidx"], d["x"], codes), "<f8")
    subsets = dict(first=np.arange(0, g, 97), second=np.arange(4, g, 1013), third=np.arange(min(g, 50)))
    q = O.per_cell_qc_metrics(g, d["colptr"], d["rowidx"], d["x"], list(subsets.values()))
    _write(dump, "F4_qc_sum.f64", q["sum"], "<f8")
    _write(dump, "F4_qc_detected.i32", q["detected"], "<i4")
    for i, (nm, genes) in enumerate(subsets.items()):
        _write(dump, f"F4_subset_{nm}_genes.i32", genes, "<i4")
        _write(dump, f"F4_qc_{nm}_sum.f64", q["subsets"][i]["sum"], "<f8")
        _write(dump, f"F4_qc_{nm}_detected.i32", q["subsets"][i]["detected"], "<i4")
        _write(dump, f"F4_qc_{nm}_percent.f64", q["subsets"][i]["percent"], "<f8")
    rows = C.compare_side_calls(dump, d, False)
    assert len(rows) == 2 and all(ok for _, _, _, ok in rows)
