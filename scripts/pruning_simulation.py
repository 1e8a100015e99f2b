"""CPU simulation (numpy) of the kNN tile-pruning strategies on a Gaussian-mixture proxy of the cfg3 embedding:
fraction of the N x N tile grid that survives (a) the first-coordinate ordering and (b) the cluster ordering with the
per-row triangle-inequality test of cb200_knn_tc.cu.  Needs scratch/emb_30000.npy (scratch/sim_prune.py makes it)."""
import sys, numpy as np, time
E0 = np.load('scratch/emb_30000.npy').astype(np.float64)
N = int(sys.argv[1]); KP = 32; T = 128
rng = np.random.default_rng(1)
# 40-blob GMM proxy fitted to the real embedding
def kmeans(X, C, iters, seed=0):
    r = np.random.default_rng(seed)
    cen = X[r.choice(len(X), C, replace=False)].copy()
    n2 = (X * X).sum(1)
    for it in range(iters + 1):
        a = np.empty(len(X), np.int64)
        for i in range(0, len(X), 65536):
            D = n2[i:i+65536, None] + (cen * cen).sum(1)[None, :] - 2 * X[i:i+65536] @ cen.T
            a[i:i+65536] = D.argmin(1)
        if it == iters: break
        for c in range(C):
            m = a == c
            if m.any(): cen[c] = X[m].mean(0)
    return cen, a
cen0, a0 = kmeans(E0, 40, 10)
parts = []
for c in range(40):
    X = E0[a0 == c]
    if len(X) < 60: continue
    mu = X.mean(0); L = np.linalg.cholesky(np.cov(X.T) + 1e-9 * np.eye(50))
    n = int(round(N * len(X) / len(E0)))
    parts.append(mu + rng.standard_normal((n, 50)) @ L.T)
E = np.concatenate(parts).astype(np.float32); E = E[rng.permutation(len(E))]
N, d = E.shape
print('N', N)
nrm = (E * E).sum(1)
t0 = time.time()
reach = np.empty(N, np.float32)
for i in range(0, N, 1024):
    D = nrm[i:i+1024, None] + nrm[None, :] - 2 * E[i:i+1024] @ E.T
    D[np.arange(D.shape[0]), np.arange(i, i + D.shape[0])] = np.inf
    reach[i:i+1024] = np.sqrt(np.maximum(np.partition(D, KP - 1, axis=1)[:, KP - 1], 0))
print('reach median %.3f max %.3f  (%.0fs)' % (np.median(reach), reach.max(), time.time() - t0))
nt0 = (N + T - 1) // T
o = np.argsort(E[:, 0]); x = E[o, 0]; r = reach[o]
lo = np.minimum.reduceat(x, np.arange(0, N, T)); hi = np.maximum.reduceat(x, np.arange(0, N, T))
cnt = 0
for t in range(nt0):
    q = x[t*T:(t+1)*T]; rr = (q + r[t*T:(t+1)*T]).max(); rl = (q - r[t*T:(t+1)*T]).min()
    cnt += np.sum((lo <= rr) & (hi >= rl))
print('PC1 ordering: %.4f' % (cnt / nt0 / nt0))

def frac_cluster(C, iters):
    sub = E[::max(1, N // 65536)]
    cen, _ = kmeans(sub.astype(np.float64), C, iters)
    cen = cen.astype(np.float32)
    Dqc = np.sqrt(np.maximum(nrm[:, None] + (cen * cen).sum(1)[None, :] - 2 * E @ cen.T, 0))
    a = Dqc.argmin(1); rho = Dqc[np.arange(N), a]
    tiles_c = []; tiles_idx = []
    for c in range(C):
        idx = np.flatnonzero(a == c); idx = idx[np.argsort(rho[idx])]
        for s in range(0, len(idx), T):
            tiles_c.append(c); tiles_idx.append(idx[s:s+T])
    nt = len(tiles_c); tc = np.array(tiles_c)
    rlo = np.array([rho[i].min() for i in tiles_idx]); rhi = np.array([rho[i].max() for i in tiles_idx])
    cnt_q = 0
    for idx in tiles_idx:
        dq = Dqc[idx][:, tc]
        need = ((dq - reach[idx][:, None] <= rhi[None, :]) & (dq + reach[idx][:, None] >= rlo[None, :])).any(0)
        cnt_q += need.sum()
    print('C=%d: tiles %d (pad %.1f%%) per-query annulus frac %.4f' % (C, nt, 100 * (nt / nt0 - 1), cnt_q / nt0 / nt0), flush=True)
for C in [40, 128, 256, 512]:
    frac_cluster(C, 4)
