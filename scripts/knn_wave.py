"""One wave of query tiles (148 x 128 queries) against n references: per-CTA time vs n."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cellula_b200
from cellula_b200 import synth
ctx = cellula_b200.Context(0)
for n in (18944, 37888, 68579, 150000, 300000):
    emb = synth.make_embedding(n, 50, seed=1)
    for waves in (1, 2):
        q = min(n, 148 * 128 * waves)
        for rep in range(2):
            ctx.knn(emb, 20, q_begin=0, q_end=q, want=False)
        t = ctx.timings()
        print(json.dumps(dict(n=n, tiles=(n + 127) // 128, queries=q, tile_ms=round(t["knn_tile_ms"], 3),
                              us_per_tile=round(1e3 * t["knn_tile_ms"] / waves / ((n + 127) // 128), 3))))
