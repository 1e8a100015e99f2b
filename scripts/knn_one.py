import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cellula_b200
from cellula_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 68579
ctx = cellula_b200.Context(0)
emb = synth.make_embedding(n, 50, seed=1)
for rep in range(3):
    try:
        ctx.knn(emb, 20, q_begin=0, q_end=min(n, 148 * 128), want=False)
    except Exception as e:  # debug modes produce garbage candidates; the tile timing is still valid
        pass
print(ctx.timings()["knn_tile_ms"])
