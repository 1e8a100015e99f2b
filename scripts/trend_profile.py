"""Host-side trend fit (scran::fitTrendVar / modelGeneVar restatement) timed on real gene statistics."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cellula_b200
from cellula_b200 import trendvar
ctx = cellula_b200.Context(0)
n = int(os.environ.get("PROBE_N", 200000))
ctx.synth_counts(27998, n, 1500.0, 20261019)
ctx.size_factors(want=False)
ctx.lognorm_genestats_local()
mean, var = ctx.genestats_finish(want=True)
print("genes", mean.size, "fit points", int(((mean >= 0.1) & (var > 1e-8)).sum()))
for _ in range(3):
    t = time.perf_counter(); st = trendvar.model_gene_var(mean, var); hv = trendvar.get_top_hvgs(st, 2000)
    print("trend + hvg: %.1f ms" % (1e3 * (time.perf_counter() - t)))
cProfile.run("trendvar.model_gene_var(mean, var)", "/tmp/prof")
pstats.Stats("/tmp/prof").sort_stats("tottime").print_stats(14)
