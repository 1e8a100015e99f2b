"""Summarises gpurun_out/timeline.bin (CB200_KNN_TIMELINE: per-role (tag, clock) events of one CTA of k_knn_tiles_tc)."""
import sys
import numpy as np
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.bin"
a = np.fromfile(path, dtype=np.int64)[:7 * 4096 * 2].reshape(-1, 4096, 2)
ev = {r: a[r][:int((a[r][:, 1] != 0).sum())] for r in range(a.shape[0])}
t0 = ev[0][0, 1]
d = {}
for x, c in ev[1]:
    d.setdefault(x >> 8, {})[x & 255] = c - t0
te = fu = q = iss = 0; n = 0; end = 0
for sq in sorted(d):
    e = d[sq]
    if 5 not in e:
        continue
    n += 1; te += e[2] - e[1]; fu += e[3] - e[2]
    if 4 in e:
        q += e[4] - e[3]; iss += e[5] - e[4]
    else:
        iss += e[5] - e[3]
    end = e[5]
print(f"mma: {n} tiles, {end} cycles ({end / max(n, 1):.0f}/tile): wait tempty {te}, wait full {fu}, wait query image {q}, issue {iss}")
for r in (2, 3, 4):
    dd = {}
    for x, c in ev[r]:
        dd.setdefault(x >> 8, {})[x & 255] = c - t0
    w = b = k = 0; hist = []
    for sq in sorted(dd):
        e = dd[sq]
        if 3 in e:
            w += e[2] - e[1]; b += e[3] - e[2]; k += 1; hist.append(e[3] - e[2])
    hist = np.array(hist)
    print(f"scanner set {r - 2}: {k} tiles, wait {w}, busy {b} ({b / max(k, 1):.0f}/tile; median {np.median(hist):.0f}, p90 {np.percentile(hist, 90):.0f})")
dd = {}
for x, c in ev[a.shape[0] - 1]:
    dd.setdefault(x >> 8, {})[x & 255] = c - t0
bd = sum(e[3] - e[2] for e in dd.values() if 3 in e)
print(f"builder: {len(dd)} visits, {bd / max(len(dd), 1):.0f} cycles per build")
p = ev[0]
iss = [(x >> 8, c - t0) for x, c in p if (x & 15) == 2]
print("producer: first issue at", iss[0][1], "last at", iss[-1][1])

# where the MMA thread's time goes, by decile of the tile sequence (warm-up vs steady state)
sq_sorted = [sq for sq in sorted(d) if 5 in d[sq]]
if sq_sorted:
    import math
    nb = 10; per = max(1, math.ceil(len(sq_sorted) / nb))
    print("tiles   cycles/tile  wait_tempty  wait_full  wait_qimg  issue")
    for b in range(0, len(sq_sorted), per):
        grp = sq_sorted[b:b + per]
        t_first = d[grp[0]][1]; t_last = d[grp[-1]][5]
        te = sum(d[s][2] - d[s][1] for s in grp); fu = sum(d[s][3] - d[s][2] for s in grp)
        q = sum(d[s][4] - d[s][3] for s in grp if 4 in d[s])
        iss = sum((d[s][5] - d[s][4]) if 4 in d[s] else (d[s][5] - d[s][3]) for s in grp)
        n = len(grp)
        print(f"{grp[0]:4d}-{grp[-1]:4d}  {(t_last - t_first) / n:8.0f}  {te / n:10.0f}  {fu / n:9.0f}  {q / n:9.0f}  {iss / n:6.0f}")
    print("CTA span (first producer event -> last MMA issue):", d[sq_sorted[-1]][5], "cycles")
