"""kNN + SNN stages alone on a fixed synthetic embedding (BASELINE cfg-4 style); prints stage timings."""
import argparse
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cellula_b200
from cellula_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=68579)
ap.add_argument("--d", type=int, default=50)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--scheme", default="rank")
a = ap.parse_args()
emb = synth.make_embedding(a.n, a.d, seed=20261020)
ctx = cellula_b200.Context(0)
for r in range(a.reps):
    ctx.knn(emb, a.k, want=False)
    ne = ctx.snn_resident(a.k, a.scheme)
    t = ctx.timings()
    print(json.dumps({k: round(v, 3) if isinstance(v, float) else v for k, v in t.items()
                      if k.startswith(("knn", "snn"))} | {"edges": ne}))
