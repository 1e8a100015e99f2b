import os, sys
sys.path.insert(0, "/root/repo")
os.environ["CB200_KNN_TIMELINE"] = "/tmp/tl.bin"
import cellula_b200
from cellula_b200 import synth
ctx = cellula_b200.Context(0)
emb = synth.make_embedding(2500, 30, seed=2500 + 30 + 30)
try:
    ctx.knn(emb, 30)
except Exception as e:
    print("ERR", e)
