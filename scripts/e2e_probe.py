"""Where does the host wait while the logcounts copy is in flight?  Wall-clock per C-ABI call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cellula_b200
from cellula_b200 import trendvar

N, G = int(os.environ.get("PROBE_N", 400000)), 27998
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = cellula_b200.Context(0, stream=stream.cuda_stream)
nnz = ctx.synth_counts(G, N, 1500.0, 20261019)
logx_h = torch.empty(nnz, dtype=torch.float64).pin_memory()
for mode in ["nocopy", "copy", "nocopy", "copy"]:
    marks = []
    def mark(name):
        marks.append((name, time.perf_counter()))
    torch.cuda.synchronize(); mark("start")
    ctx.size_factors(want=False); mark("sf")
    ctx.lognorm_genestats_local(); mean, var = ctx.genestats_finish(want=True); mark("lognorm")
    if mode == "copy":
        ctx.fetch_logcounts_async(logx_h.numpy()); mark("copy_submit")
    stats = trendvar.model_gene_var(mean, var); hv = trendvar.get_top_hvgs(stats, 2000); mark("trend")
    ctx.pca_gram_local(hv); mark("gram")
    ctx.synchronize(); mark("gram_sync")
    ctx.pca_finish(50, len(hv), want_scores=False); mark("pca_finish")
    ctx.knn(None, 20, n_total=N, d=50, want=False); mark("knn")
    ne = ctx.snn_resident(20, "rank"); mark("snn")
    ctx.wait_copies(); mark("wait")
    t0 = marks[0][1]
    print(mode, " ".join("%s=%.1f" % (k, 1e3 * (t - t0)) for k, t in marks[1:]), flush=True)
