"""Stage timings on a device-generated matrix of a given shape (no e2e, no CPU baseline)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cellula_b200
from cellula_b200 import trendvar
ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=300000)
ap.add_argument("--genes", type=int, default=27998)
ap.add_argument("--nnz", type=float, default=1500.0)
ap.add_argument("--ndims", type=int, default=50)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--upto", default="snn")
a = ap.parse_args()
ctx = cellula_b200.Context(0)
nnz = ctx.synth_counts(a.genes, a.cells, a.nnz, 20261019)
print("nnz", nnz, nnz / a.cells)
for r in range(a.reps):
    ctx.size_factors(want=False)
    _, mean, var = ctx.lognorm_genestats(want_logx=False)
    hv = trendvar.get_top_hvgs(trendvar.model_gene_var(mean, var), 2000)
    ctx.pca(hv, a.ndims, want_scores=False)
    if a.upto == "snn":
        ctx.knn(None, a.k, n_total=a.cells, d=a.ndims, want=False)
        ctx.snn_resident(a.k, "rank")
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in ctx.timings().items()}))
