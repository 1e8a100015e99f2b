"""Per-tile scanner timing and cumulative ring pushes of the traced row (tag 6 of the scanner timelines)."""
import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.int64)[:7 * 4096 * 2].reshape(-1, 4096, 2)
t0 = a[0][0, 1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rows = {}
for r in (2, 3, 4):
    e = a[r]; e = e[:int((e[:, 0] != 0).sum() if (e[:, 1] != 0).sum() == 0 else max((e[:, 1] != 0).sum(), (e[:, 0] != 0).sum()))]
    for x, c in e:
        x = int(x); rows.setdefault(x >> 8, {})[x & 255] = int(c) - (0 if (x & 255) >= 6 else int(t0))
seqs = sorted(rows)
print("seq   wait  busy  pushes(cum)  at")
prev = 0
for i in range(0, len(seqs), step):
    grp = [rows[s] for s in seqs[i:i + step] if 3 in rows[s]]
    if not grp:
        continue
    w = sum(g[2] - g[1] for g in grp) / len(grp); b = sum(g[3] - g[2] for g in grp) / len(grp)
    cum = max([g.get(6, 0) for g in grp] + [prev]); 
    ld = sum(g.get(7, 0) for g in grp) / len(grp); hit = sum(g.get(8, 0) for g in grp) / len(grp); nh = sum(g.get(9, 0) for g in grp) / len(grp)
    print(f"{seqs[i]:4d} {w:7.0f} {b:6.0f}  ld {ld:6.0f} hitpath {hit:6.0f} hitchunks {nh:.1f} (+{cum - prev if cum >= prev else 0})  {grp[0][1]}")
    prev = cum
pr = a[6][:128]
print("per-row pushes: mean %.1f median %.0f p90 %.0f max %d; inserts mean %.1f max %d" % (pr[:, 0].mean(), np.median(pr[:, 0]), np.percentile(pr[:, 0], 90), pr[:, 0].max(), pr[:, 1].mean(), pr[:, 1].max()))
print("rows by pushes (desc):", sorted(pr[:, 0].tolist(), reverse=True)[:24])
