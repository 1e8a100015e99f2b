"""Host-side mean-variance trend and HVG selection (rows a5/a6 of SURVEY.md section 8).

In the R product these steps stay *unchanged R calls* on the GPU-computed per-gene means and
variances (`scran::fitTrendVar`, `scran::getTopHVGs`; /root/reference/R/doNormAndReduce.R:88-90,
r-pkg/R/doNormAndReduce.R).  The Python mirror has no R, so this module supplies the same
decomposition on the host: G-sized work (~30k points), never on the device.

model_gene_var(mean, var) -> dict(mean, total, tech, bio, p_value, FDR)   (modelGeneVar columns)
get_top_hvgs(stats, n)    -> 0-based row indices ordered by decreasing bio, bio > 0 only
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import ndtr


def _native(name):
    """A host helper of the shared library (csrc/cb200_host.cu), or None when it has not been built."""
    try:
        from . import _lib
        return getattr(_lib.lib(), name)
    except Exception:
        return None


def _native_lowess():
    return _native("cb200_host_weighted_lowess")


def _density_weights(m):
    """1 / kernel density of the means (Gaussian, bw.nrd0, 512-point grid), mean-scaled."""
    n = m.size
    sd = m.std(ddof=1)
    iqr = np.subtract(*np.percentile(m, [75, 25]))
    lo = min(sd, iqr / 1.34) or sd or abs(m[0]) or 1.0
    bw = 0.9 * lo * n ** -0.2
    lo_g, hi_g = m.min() - 3 * bw, m.max() + 3 * bw
    grid = np.linspace(lo_g, hi_g, 512)
    dens = None
    native = _native("cb200_host_gauss_kde")
    if native is not None:
        mc = np.ascontiguousarray(m)
        out = np.empty(512)
        if native(n, mc.ctypes.data, float(bw), float(lo_g), float(hi_g), 512, out.ctypes.data) == 0:
            dens = out
    if dens is None:
        dens = np.zeros(512)
        step = max(1, 4_000_000 // 512)
        for a in range(0, n, step):
            z = (grid[:, None] - m[None, a:a + step]) / bw
            dens += np.exp(-0.5 * z * z).sum(axis=1)
        dens /= n * bw * math.sqrt(2 * math.pi)
    w = 1.0 / np.interp(m, grid, dens)
    return w / w.mean()


def _start_values(v, m):
    order = np.argsort(m, kind="stable")
    keep = order[:max(1, int(min(100, 0.1 * v.size)))]
    grad = float(m[keep] @ v[keep]) / float(m[keep] @ m[keep])
    bgrid = 2.0 ** np.linspace(-5, 5, 10)
    ngrid = 2.0 ** np.linspace(0, 10, 10)
    native = _native("cb200_host_nls_grid")
    if native is not None:   # same sums in C++ (csrc/cb200_host.cu): no 10 x 10 x G temporaries
        import ctypes
        vc, mc = np.ascontiguousarray(v), np.ascontiguousarray(m)
        ib, inn = ctypes.c_int(), ctypes.c_int()
        if native(v.size, vc.ctypes.data, mc.ctypes.data, grad, 10, bgrid.ctypes.data, 10, ngrid.ctypes.data,
                  ctypes.addressof(ib), ctypes.addressof(inn)) == 0:
            return bgrid[ib.value] * grad, bgrid[ib.value], ngrid[inn.value]
    with np.errstate(over="ignore"):
        mp = m[None, :] ** ngrid[:, None]                                   # [n-grid, genes]
        pred = grad * bgrid[None, :, None] * m[None, None, :] / (mp[:, None, :] + bgrid[None, :, None])
        rss = ((v[None, None, :] - pred) ** 2).sum(axis=2)                   # [n-grid, b-grid]
    i_n, i_b = np.unravel_index(np.argmin(rss), rss.shape)  # first minimum, b fastest: expand.grid order
    return bgrid[i_b] * grad, bgrid[i_b], ngrid[i_n]


def _fit_parametric(v, m, w, tol=1e-9, maxiter=500):
    """Weighted least squares of v ~ exp(A) m / (m^(1+exp(N)) + exp(B)), damped Gauss-Newton."""
    a0, b0, n0 = _start_values(v, m)
    th = np.array([math.log(a0), math.log(b0), math.log(max(n0 - 1.0, 1e-8))])
    native = _native("cb200_host_nls_trend")
    if native is not None:   # same iteration in C++ (csrc/cb200_host.cu)
        vc, mc, wc = np.ascontiguousarray(v), np.ascontiguousarray(m), np.ascontiguousarray(w)
        tc = th.copy()
        rc = native(v.size, vc.ctypes.data, mc.ctypes.data, wc.ctypes.data, tc.ctypes.data, float(tol), int(maxiter))
        if rc in (0, 2):
            return np.exp(tc)
    logm = np.log(m)

    def evaluate(t):
        ea, eb, en = np.exp(t)
        p = np.exp((1.0 + en) * logm)
        den = p + eb
        f = ea * m / den
        return f, np.column_stack((f, -f * eb / den, -f * p * logm * en / den))

    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        f, jac = evaluate(th)
    if not (np.all(np.isfinite(f)) and np.all(np.isfinite(jac))):
        return np.exp(th)  # scran keeps the grid-search start values when nls() cannot run
    rss = float(w @ (v - f) ** 2)
    damp = 1e-3
    sw = np.sqrt(w)
    for _ in range(maxiter):
        r = sw * (v - f)
        jw = jac * sw[:, None]
        q, _ = np.linalg.qr(jw)
        proj = q.T @ r
        num = float(proj @ proj)
        if math.sqrt(num / max(float(r @ r) - num, 1e-300)) < tol:
            break
        a = jw.T @ jw
        g = jw.T @ r
        for _try in range(40):
            try:
                step = np.linalg.solve(a + damp * np.diag(np.diag(a) + 1e-300), g)
            except np.linalg.LinAlgError:
                damp *= 10.0
                continue
            with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
                f2, jac2 = evaluate(th + step)
                rss2 = float(w @ (v - f2) ** 2)
            if np.isfinite(rss2) and rss2 < rss and np.all(np.isfinite(jac2)):
                th, f, jac, rss = th + step, f2, jac2, rss2
                damp = max(damp / 10.0, 1e-12)
                break
            damp *= 10.0
        else:
            break
    return np.exp(th)


def _wmedian(x, w):
    o = np.argsort(x, kind="stable")
    cw = np.cumsum(w[o])
    half = 0.5 * cw[-1]
    i = int(np.searchsorted(cw, half, side="left"))
    if i + 1 < x.size and cw[i] == half:
        return 0.5 * (x[o[i]] + x[o[i + 1]])
    return x[o[i]]


def _weighted_lowess(x, y, w, span=0.3, iterations=4, npts=200):
    """limma::weightedLowess: local linear fits at ~npts seeds over the sorted x, window =
    nearest points holding span * total weight, tricube kernel, bisquare robustness."""
    o = np.argsort(x, kind="stable")
    xs, ys, ws = np.ascontiguousarray(x[o]), np.ascontiguousarray(y[o]), np.ascontiguousarray(w[o])
    n = xs.size
    native = _native_lowess()
    if native is not None:   # same algorithm in C++ (csrc/cb200_host.cu); the loops below are its definition
        fitted = np.empty(n)
        rc = native(n, xs.ctypes.data, ys.ctypes.data, ws.ctypes.data, float(span), int(iterations), int(npts),
                    fitted.ctypes.data)
        if rc == 0:
            out = np.empty(n)
            out[o] = fitted
            return out
    delta = (xs[-1] - xs[0]) / npts
    seeds = [0]
    for i in range(1, n - 1):
        if xs[i] - xs[seeds[-1]] > delta:
            seeds.append(i)
    if n > 1:
        seeds.append(n - 1)
    target = span * ws.sum()
    windows = []
    for s in seeds:
        lo = hi = s
        acc = ws[s]
        while acc < target and (lo > 0 or hi < n - 1):
            go_left = hi == n - 1 or (lo > 0 and xs[s] - xs[lo - 1] < xs[hi + 1] - xs[s])
            if go_left:
                lo -= 1
                acc += ws[lo]
            else:
                hi += 1
                acc += ws[hi]
        dist = max(xs[s] - xs[lo], xs[hi] - xs[s])
        while lo > 0 and xs[s] - xs[lo - 1] <= dist:
            lo -= 1
        while hi < n - 1 and xs[hi + 1] - xs[s] <= dist:
            hi += 1
        windows.append((lo, hi + 1, dist))
    rob = np.ones(n)
    fitted = ys.copy()
    sx = xs[seeds]
    for it in range(iterations):
        sv = np.empty(len(seeds))
        for q, (s, (lo, hi, dist)) in enumerate(zip(seeds, windows)):
            xx, yy = xs[lo:hi], ys[lo:hi]
            kw = ws[lo:hi] * rob[lo:hi]
            if dist > 0:
                u = np.abs(xx - xs[s]) / dist
                kw = kw * np.where(u < 1, (1 - u ** 3) ** 3, 0.0)
            tw = kw.sum()
            if tw <= 0:
                sv[q] = ys[s]
                continue
            xm, ym = (kw @ xx) / tw, (kw @ yy) / tw
            sxx = kw @ (xx - xm) ** 2
            if sxx < 1e-12 * max(dist * dist, 1e-300) * tw:
                sv[q] = ym
            else:
                sv[q] = ym + (kw @ ((xx - xm) * (yy - ym))) / sxx * (xs[s] - xm)
        fitted = np.interp(xs, sx, sv)
        if it == iterations - 1:
            break
        res = np.abs(ys - fitted)
        cmad = 6.0 * _wmedian(res, ws)
        if cmad <= 1e-12 * max(res.mean(), 1e-300):
            break
        rob = np.where(res < cmad, (1 - (res / cmad) ** 2) ** 2, 0.0)
    out = np.empty(n)
    out[o] = fitted
    return out


def fit_trend_var(means, variances, min_mean=0.1):
    """Returns (trend_function, std_dev) like scran::fitTrendVar(means, vars)."""
    means = np.asarray(means, dtype=np.float64)
    variances = np.asarray(variances, dtype=np.float64)
    ok = np.isfinite(variances) & (variances > 1e-8) & (means >= min_mean)
    v, m = variances[ok], means[ok]
    if v.size <= 3:
        raise ValueError("need at least 4 points for non-linear curve fitting")
    w = _density_weights(m)
    ea, eb, en = _fit_parametric(v, m, w)

    def parametric(x):
        return ea * x / (x ** (1.0 + en) + eb)
    resid = np.log(v) - np.log(parametric(m))
    smooth = _weighted_lowess(m, resid, w)
    ux, inv = np.unique(m, return_inverse=True)
    uy = np.bincount(inv, weights=smooth) / np.bincount(inv)

    def unscaled(x):
        return np.exp(np.interp(x, ux, uy)) * parametric(x)
    left = v / unscaled(m)
    med = _wmedian(left, w)
    std_dev = float(_wmedian(np.abs(left / med - 1.0), w) * 1.4826)

    def trend(x):
        x = np.asarray(x, dtype=np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            t = unscaled(np.maximum(x, 0.0)) * med
        return np.where(x > 0, t, 0.0)
    return trend, std_dev


def model_gene_var(means, variances, want_pvalues=True):
    """scran::modelGeneVar's DataFrame columns from per-gene mean / total variance.
    want_pvalues=False skips p.value / FDR: cellula only feeds the table to getTopHVGs(), which ranks by
    `bio` (R/doNormAndReduce.R:88-91), so the sharded pipeline does not spend host time on them."""
    means = np.asarray(means, dtype=np.float64)
    variances = np.asarray(variances, dtype=np.float64)
    trend, std_dev = fit_trend_var(means, variances)
    tech = trend(means)
    bio = variances - tech
    if not want_pvalues:
        return dict(mean=means, total=variances, tech=tech, bio=bio)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = variances / tech
    p = np.where(np.isfinite(ratio), ndtr(-(ratio - 1.0) / std_dev), np.nan)
    fdr = np.full(p.shape, np.nan)
    valid = np.isfinite(p)
    if valid.any():
        pv = p[valid]
        o = np.argsort(pv, kind="stable")[::-1]
        adj = np.minimum.accumulate(pv[o] * pv.size / np.arange(pv.size, 0, -1))
        tmp = np.empty(pv.size)
        tmp[o] = np.minimum(adj, 1.0)
        fdr[valid] = tmp
    return dict(mean=means, total=variances, tech=tech, bio=bio, p_value=p, FDR=fdr)


def model_gene_var_blocked(means, variances, n_per_block):
    """scran::modelGeneVar(block=) from per-block statistics (rows = blocks): one trend per block with >= 2 cells,
    then the unweighted average of mean / total / tech / bio over those blocks (equiweight = TRUE)."""
    means = np.asarray(means, dtype=np.float64)
    variances = np.asarray(variances, dtype=np.float64)
    use = [b for b in range(means.shape[0]) if n_per_block[b] >= 2]
    if not use:
        raise ValueError("no block with at least 2 cells")
    per = [model_gene_var(means[b], variances[b], want_pvalues=False) for b in use]
    return {k: np.mean([p[k] for p in per], axis=0) for k in ("mean", "total", "tech", "bio")}


def rescale_size_factors(ave, sf, block, min_mean=1.0):
    """batchelor:::.rescale_size_factors on the GPU-computed batch averages (rows = batches): pairwise median ratios
    over the genes with grand mean >= min_mean; every batch is scaled down to the batch of lowest coverage."""
    ave = np.asarray(ave, dtype=np.float64)
    sf = np.asarray(sf, dtype=np.float64)
    block = np.asarray(block)
    nb = ave.shape[0]
    ratios = np.ones((nb, nb))
    sums = ave.sum(axis=1)
    for a in range(nb - 1):
        for b in range(a + 1, nb):
            grand = (ave[a] / sums[a] + ave[b] / sums[b]) / 2 * (sums[a] + sums[b]) / 2
            keep = grand >= min_mean
            with np.errstate(divide="ignore", invalid="ignore"):
                cur = np.median(ave[b][keep] / ave[a][keep]) if keep.any() else np.nan
            if not np.isfinite(cur) or cur == 0:
                raise ValueError("median ratio of averages between batches is not finite")
            ratios[a, b] = cur
            ratios[b, a] = 1.0 / cur
    smallest = int(np.argmin(ratios.min(axis=0)))
    out = np.empty_like(sf)
    for b in range(nb):
        cells = block == b
        out[cells] = sf[cells] / sf[cells].mean() / ratios[b, smallest]
    return out


def get_top_hvgs(stats, n):
    bio = np.asarray(stats["bio"] if isinstance(stats, dict) else stats)
    keep = np.flatnonzero(np.isfinite(bio) & (bio > 0))
    if keep.size > 4 * n:   # only the top of the list is needed: cut at the n-th value (ties kept), then the stable sort
        cut = np.partition(bio[keep], keep.size - n)[keep.size - n]
        keep = keep[bio[keep] >= cut]
    return keep[np.argsort(-bio[keep], kind="stable")][:n].astype(np.int32)
