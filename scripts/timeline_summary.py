"""Summarises gpurun_out/timeline.bin (CB200_KNN_TIMELINE: per-role (tag, clock) events of one CTA of k_knn_tiles_tc)."""
import sys
import numpy as np
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.bin"
a = np.fromfile(path, dtype=np.int64).reshape(5, 4096, 2)
ev = {r: a[r][:int((a[r][:, 1] != 0).sum())] for r in range(5)}
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
for r in (2, 3):
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
for x, c in ev[4]:
    dd.setdefault(x >> 8, {})[x & 255] = c - t0
bd = sum(e[3] - e[2] for e in dd.values() if 3 in e)
print(f"builder: {len(dd)} visits, {bd / max(len(dd), 1):.0f} cycles per build")
p = ev[0]
iss = [(x >> 8, c - t0) for x, c in p if (x & 15) == 2]
print("producer: first issue at", iss[0][1], "last at", iss[-1][1])
