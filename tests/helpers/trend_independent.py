"""TEST INFRASTRUCTURE.  A second, independently structured restatement of scran::fitTrendVar's definition
(SURVEY.md row a5), used to check cellula_b200/trendvar.py (product mirror) and oracle/oracle.py (oracle), which
were written by the same hand with the same loop structure.  Nothing here shares code or algorithmic shortcuts
with them:

  * density weights -- stats::density's definition evaluated with scipy.stats.norm on the 512-point grid
  * parametric fit   -- scipy.optimize.least_squares (trust-region reflective) on sqrt(w) (v - f(theta)),
                        start values from an explicit double loop over the 10 x 10 grid
  * weighted lowess  -- dense: for every seed the window is found by sorting ALL points by distance to the seed
                        and cutting the cumulative weight, the local line is a weighted np.polyfit
  * weighted median  -- searchsorted on the cumulative weights of the sorted values

Agreement is reported as a residual, not assumed (tests/test_trend_independent.py).
"""
import numpy as np
from scipy.optimize import least_squares
from scipy.stats import norm


def density_weights(m):
    n = m.size
    q75, q25 = np.percentile(m, [75, 25])
    spread = min(np.std(m, ddof=1), (q75 - q25) / 1.34)
    bw = 0.9 * spread * n ** (-1.0 / 5.0)
    grid = np.linspace(m.min() - 3 * bw, m.max() + 3 * bw, 512)
    dens = np.array([norm.pdf(g, loc=m, scale=bw).mean() for g in grid])
    w = 1.0 / np.interp(m, grid, dens)
    return w / np.mean(w)


def start_values(v, m):
    order = np.argsort(m, kind="stable")
    left = order[: max(1, int(min(100, v.size * 0.1)))]
    grad = np.sum(m[left] * v[left]) / np.sum(m[left] ** 2)
    best = None
    for n_try in 2.0 ** np.linspace(0, 10, 10):       # expand.grid(B = , n = ): B varies fastest
        for b_try in 2.0 ** np.linspace(-5, 5, 10):
            with np.errstate(over="ignore"):
                rss = np.sum((v - (grad * b_try) * m / (m ** n_try + b_try)) ** 2)
            if best is None or rss < best[0]:
                best = (rss, grad * b_try, b_try, n_try)
    return best[1:]


def fit_parametric(v, m, w):
    a0, b0, n0 = start_values(v, m)
    sw = np.sqrt(w)

    def resid(t):
        with np.errstate(over="ignore"):
            return sw * (v - np.exp(t[0]) * m / (m ** (1.0 + np.exp(t[2])) + np.exp(t[1])))
    t0 = np.array([np.log(a0), np.log(b0), np.log(max(n0 - 1.0, 1e-8))])
    if not np.all(np.isfinite(resid(t0))):
        return np.exp(t0)
    sol = least_squares(resid, t0, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=5000, x_scale="jac")
    return np.exp(sol.x)


def weighted_median(x, w):
    o = np.argsort(x, kind="stable")
    c = np.cumsum(w[o])
    i = int(np.searchsorted(c, 0.5 * c[-1], side="left"))
    if i + 1 < x.size and c[i] == 0.5 * c[-1]:
        return 0.5 * (x[o[i]] + x[o[i + 1]])
    return x[o[i]]


def weighted_lowess_dense(x, y, w, span=0.3, iterations=4, npts=200):
    o = np.argsort(x, kind="stable")
    xs, ys, ws = x[o], y[o], w[o]
    n = xs.size
    delta = (xs[-1] - xs[0]) / npts
    seeds = [0]
    for i in range(1, n - 1):
        if xs[i] - xs[seeds[-1]] > delta:
            seeds.append(i)
    if n > 1:
        seeds.append(n - 1)
    target = span * ws.sum()
    idx = np.arange(n)
    radius = np.empty(len(seeds))
    for q, s in enumerate(seeds):
        dist = np.abs(xs - xs[s])
        # nearest points first; equidistant points: the right one first (limma extends right on a tie), then by
        # position away from the seed
        near = np.lexsort((np.abs(idx - s), idx < s, dist))
        cum = np.cumsum(ws[near])
        cut = int(np.searchsorted(cum, target, side="left"))
        radius[q] = dist[near[min(cut, n - 1)]]
    rob = np.ones(n)
    fit = ys.copy()
    for it in range(iterations):
        sv = np.empty(len(seeds))
        for q, s in enumerate(seeds):
            dist = np.abs(xs - xs[s])
            inside = dist <= radius[q]
            kw = ws[inside] * rob[inside]
            if radius[q] > 0:
                u = dist[inside] / radius[q]
                kw = kw * np.where(u < 1, (1 - u ** 3) ** 3, 0.0)
            if kw.sum() <= 0:
                sv[q] = ys[s]
                continue
            xx, yy = xs[inside], ys[inside]
            xm = np.average(xx, weights=kw)
            if np.sum(kw * (xx - xm) ** 2) < 1e-12 * max(radius[q] ** 2, 1e-300) * kw.sum():
                sv[q] = np.average(yy, weights=kw)
            else:
                good = kw > 0
                slope, icpt = np.polyfit(xx[good] - xs[s], yy[good], 1, w=np.sqrt(kw[good]))
                sv[q] = icpt
        fit = np.interp(xs, xs[seeds], sv)
        if it == iterations - 1:
            break
        res = np.abs(ys - fit)
        cmad = 6.0 * weighted_median(res, ws)
        if cmad <= 1e-12 * max(res.mean(), 1e-300):
            break
        rob = np.where(res < cmad, (1 - (res / cmad) ** 2) ** 2, 0.0)
    out = np.empty(n)
    out[o] = fit
    return out


def model_gene_var(means, variances, min_mean=0.1):
    means = np.asarray(means, dtype=np.float64)
    variances = np.asarray(variances, dtype=np.float64)
    ok = np.isfinite(variances) & (variances > 1e-8) & (means >= min_mean)
    v, m = variances[ok], means[ok]
    w = density_weights(m)
    ea, eb, en = fit_parametric(v, m, w)

    def parametric(x):
        return ea * x / (x ** (1.0 + en) + eb)
    smooth = weighted_lowess_dense(m, np.log(v) - np.log(parametric(m)), w)
    ux, inv = np.unique(m, return_inverse=True)
    uy = np.bincount(inv, weights=smooth) / np.bincount(inv)

    def unscaled(x):
        return np.exp(np.interp(x, ux, uy)) * parametric(x)
    med = weighted_median(v / unscaled(m), w)
    with np.errstate(invalid="ignore", divide="ignore"):
        tech = np.where(means > 0, unscaled(np.maximum(means, 0.0)) * med, 0.0)
    return dict(mean=means, total=variances, tech=tech, bio=variances - tech, params=(ea, eb, en))
