"""Deconvolution size factors at scale: device-generated matrix, synthetic cluster labels (contiguous
blocks of cells, as quickCluster's min.size = sqrt(N) would give at least), stage times of cb200_pooled_size_factors.
    python scripts/pool_bench.py --cells 1306127 --genes 27998 --nnz 1500 [--cpu-sample 600]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cellula_b200
ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=68579)
ap.add_argument("--genes", type=int, default=32738)
ap.add_argument("--nnz", type=float, default=525.0)
ap.add_argument("--clusters", type=int, default=0, help="0 = cells / 30000 (min 1)")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--cpu-sample", type=int, default=0, help="time the numpy oracle on the first cells (one cluster)")
a = ap.parse_args()
ctx = cellula_b200.Context(0)
nnz = ctx.synth_counts(a.genes, a.cells, a.nnz, 20261019)
ncl = a.clusters or max(1, a.cells // 30000)
cl = (np.arange(a.cells) * ncl // a.cells).astype(np.int32)
out = dict(cells=a.cells, genes=a.genes, nnz=int(nnz), clusters_in=ncl)
for r in range(a.reps):
    t0 = time.perf_counter()
    sf, nbad = ctx.pooled_size_factors(cl)
    out[f"rep{r}"] = dict(wall_ms=1e3 * (time.perf_counter() - t0), nonpositive=nbad, **ctx.pool_ms)
out["sf_quantiles"] = np.quantile(sf, [0, 0.01, 0.5, 0.99, 1]).tolist()
lib = ctx.size_factors()
out["corr_with_library_factors"] = float(np.corrcoef(sf, lib)[0, 1])
if a.cpu_sample:
    from oracle import oracle as O
    n = a.cpu_sample
    colptr, rowidx, x = ctx.export_csc()
    e = int(colptr[n])
    t0 = time.perf_counter()
    want = O.pooled_size_factors(a.genes, colptr[:n + 1], rowidx[:e], x[:e], None)
    out["cpu_oracle"] = dict(cells=n, seconds=time.perf_counter() - t0, note="numpy restatement (dense, python loops), one cluster")
    ctx.load_csc(a.genes, colptr[:n + 1].copy(), rowidx[:e].copy(), x[:e].copy())
    got, _ = ctx.pooled_size_factors(None)
    out["cpu_oracle"]["max_rel_diff_vs_gpu"] = float(np.max(np.abs(got - want) / np.abs(want)))
print(json.dumps(out))
