"""Debug aid: runs chosen kNN cases under several environment settings and reports parity with the C oracle."""
import os, sys, itertools, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cellula_b200
from cellula_b200 import synth
from oracle import oracle

cases = [(2500, 30, 30), (3000, 50, 20), (129, 50, 52), (128, 16, 40), (4097, 61, 24), (257, 3, 25), (20000, 50, 30), (20000, 50, 20)]
envs = [dict(CB200_KNN_SEED=s) for s in os.environ.get("SEEDS", "0 64 -32").split()]
for env in envs:
    os.environ.update({k: str(v) for k, v in env.items()})
    ctx = cellula_b200.Context(0)
    for n, d, k in cases:
        emb = synth.make_embedding(n, d, seed=n + d + k)
        try:
            got = ctx.knn(emb, k)
            want = oracle.find_knn(emb, k)
            bad = int((got != want).any(axis=1).sum())
            tm = ctx.timings()
            print(env, (n, d, k), "rows wrong", bad, "fallback", tm.get("knn_fallback_queries"), flush=True)
        except Exception as e:
            print(env, (n, d, k), "ERROR", str(e)[:120], flush=True)
    ctx.close()
