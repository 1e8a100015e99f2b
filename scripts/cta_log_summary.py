"""Per-CTA log of k_knn_tiles_tc (CB200_KNN_TIMELINE dump, after the 7 role timelines): start/end (globaltimer ns), tiles, SM."""
import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.int64)
c = a[7 * 4096 * 2:].reshape(-1, 4)
c = c[c[:, 1] > 0]
t0 = c[:, 0].min(); start = (c[:, 0] - t0) / 1e3; end = (c[:, 1] - t0) / 1e3; dur = end - start; tiles = c[:, 2]
print(f"{len(c)} CTAs, kernel span {end.max() / 1e3:.2f} ms, tiles/CTA mean {tiles.mean():.0f} min {tiles.min()} p50 {np.median(tiles):.0f} p90 {np.percentile(tiles, 90):.0f} max {tiles.max()}")
print(f"CTA duration us: mean {dur.mean():.0f} p50 {np.median(dur):.0f} p90 {np.percentile(dur, 90):.0f} max {dur.max():.0f}; ns/tile overall {1e3 * dur.sum() / tiles.sum():.0f}")
# linear model dur = a + b * tiles
A = np.vstack([np.ones_like(tiles, dtype=float), tiles.astype(float)]).T
coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
print(f"fit: duration = {coef[0]:.0f} us + {1e3 * coef[1]:.0f} ns per tile")
sm_busy = {}
for s, d in zip(c[:, 3], dur): sm_busy[s] = sm_busy.get(s, 0) + d
b = np.array(list(sm_busy.values()))
print(f"per-SM busy ms: mean {b.mean() / 1e3:.2f} min {b.min() / 1e3:.2f} max {b.max() / 1e3:.2f} over {len(b)} SMs; last CTA start at {start.max() / 1e3:.2f} ms")
last = np.argsort(end)[-5:]
print("last CTAs to finish (start ms, dur us, tiles):", [(round(start[i] / 1e3, 2), int(dur[i]), int(tiles[i])) for i in last])
q = np.argsort(c[:, 0])
n = len(q); 
for f in range(0, 10):
    sl = q[f * n // 10:(f + 1) * n // 10]
    print(f"  launch decile {f}: tiles/CTA {tiles[sl].mean():6.0f}  us/CTA {dur[sl].mean():7.0f}  ns/tile {1e3 * dur[sl].sum() / max(1, tiles[sl].sum()):6.0f}")
