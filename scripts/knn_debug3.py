import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CB200_KNN_ENGINE"] = "tc"
import cellula_b200
from cellula_b200 import synth
ctx = cellula_b200.Context(0)
emb = synth.make_embedding(int(os.environ.get("N", 60000)), 50, seed=31)
try:
    got = ctx.knn(emb, 20)
    print("ok", ctx.timings()["knn_tiles_scanned"], ctx.timings()["knn_fallback_queries"])
except Exception as e:
    print("ERR", e)
