"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU (numpy fp64 + plain C) restatement of the upstream routines that cellula's
``doNormAndReduce()`` (/root/reference/R/doNormAndReduce.R:57-97) and
``makeGraphsAndClusters()`` (/root/reference/R/makeGraphsAndClusters.R:80-85) call.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  Nothing under ``cellula_b200/``
imports it; the product path fails loudly without its CUDA library instead.

PARITY UNPINNED.  All arithmetic of the path lives in third-party R/Bioconductor
packages that are neither vendored in /root/reference nor version-pinned
(DESCRIPTION:14-50): scuttle (logNormCounts), scran (modelGeneVar, fitTrendVar,
getTopHVGs), scater/BiocSingular (runPCA), BiocNeighbors/knncolle (findKNN),
bluster/libscran (makeSNNGraph).  Estimated era Bioconductor 3.20.  The reference holds
no tests, golden vectors or fixtures and R cannot run in the build
image, so every function below follows the packages' *published* definitions as
collected in SURVEY.md 8c.1 and is pinned only by the hand-derived known-answer vectors
of SURVEY.md 8c.3 (tests/test_oracle_kat.py, tests/golden/kat_*.json).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/knn_snn.c -> oracle/liboracle.so (gcc, no FMA contraction)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "knn_snn.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        base = ["/usr/bin/gcc", "-O2", "-fPIC", "-ffp-contract=off", "-std=c11", "-shared",
                "-o", so, src, "-lm"]
        try:
            subprocess.run(base[:5] + ["-fopenmp"] + base[5:], check=True, capture_output=True)
        except (subprocess.CalledProcessError, FileNotFoundError):
            subprocess.run(["gcc"] + base[1:], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        i64, i32p, f64p = ctypes.c_int64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
        lib.oracle_knn.argtypes = [f64p, i64, ctypes.c_int, ctypes.c_int, i32p, f64p]
        lib.oracle_knn.restype = ctypes.c_int
        lib.oracle_snn.argtypes = [i32p, i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(i64),
                                   ctypes.POINTER(i32p), ctypes.POINTER(i32p), ctypes.POINTER(f64p)]
        lib.oracle_snn.restype = ctypes.c_int
        lib.oracle_snn_range.argtypes = [i32p, i64, ctypes.c_int, ctypes.c_int, i64, i64, ctypes.POINTER(i64),
                                         ctypes.POINTER(i32p), ctypes.POINTER(i32p), ctypes.POINTER(f64p)]
        lib.oracle_snn_range.restype = ctypes.c_int
        lib.oracle_knn_queries.argtypes = [f64p, i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(i64), i64, i32p, f64p]
        lib.oracle_knn_queries.restype = ctypes.c_int
        lib.oracle_free.argtypes = [ctypes.c_void_p]
        lib.oracle_lognorm_genestats.argtypes = [i64, i64, ctypes.POINTER(i64), i32p, f64p, f64p, f64p,
                                                 f64p, f64p]
        lib.oracle_lognorm_genestats.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# ---------------------------------------------------------------------------------------
# K1/K2: scuttle::logNormCounts  (R/doNormAndReduce.R:78)
# ---------------------------------------------------------------------------------------
def library_size_factors(colptr, x, size_factors_in=None):
    """sf = colSums(counts) / mean(colSums(counts)); a user vector is only re-centred.

    Errors like scuttle does when any factor is non-positive or NA."""
    colptr = np.asarray(colptr, dtype=np.int64)
    n = colptr.size - 1
    if size_factors_in is None:
        x = np.asarray(x, dtype=np.float64)
        ls = np.zeros(n)
        nz = np.diff(colptr) > 0
        ls[nz] = np.add.reduceat(x, colptr[:-1][nz]) if x.size else 0.0
        # reduceat on empty trailing segments is handled by the nz mask above
    else:
        ls = np.asarray(size_factors_in, dtype=np.float64).copy()
    sf = ls / ls.mean()
    if not np.all(np.isfinite(sf)) or np.any(sf <= 0):
        raise ValueError("size factors should be positive")
    return sf


def log_norm_counts(colptr, x, sf):
    """logcounts@x = log1p(x / sf[col]) / log(2) on the stored non-zeros (same i/p)."""
    colptr = np.asarray(colptr, dtype=np.int64)
    col_of = np.repeat(np.arange(colptr.size - 1), np.diff(colptr))
    return np.log1p(np.asarray(x, dtype=np.float64) / sf[col_of]) / np.log(2.0)


# ---------------------------------------------------------------------------------------
# K3: scran:::.compute_mean_var  (R/doNormAndReduce.R:88)
# ---------------------------------------------------------------------------------------
def gene_mean_var(n_genes, n_cells, rowidx, y):
    """Per-gene mean and (n-1)-variance over ALL cells, implicit zeros included.

    Two-pass form (exact mean first), mathematically what upstream's Welford computes."""
    rowidx = np.asarray(rowidx, dtype=np.int64)
    s = np.bincount(rowidx, weights=y, minlength=n_genes)
    nnz = np.bincount(rowidx, minlength=n_genes).astype(np.float64)
    mean = s / n_cells
    dev = y - mean[rowidx]
    ss = np.bincount(rowidx, weights=dev * dev, minlength=n_genes)
    ss += (n_cells - nnz) * mean * mean
    var = ss / (n_cells - 1) if n_cells > 1 else np.full(n_genes, np.nan)
    return mean, var


def c_lognorm_genestats(n_genes, colptr, rowidx, x, want_logx=True):
    """Same three steps through oracle/knn_snn.c (single thread, Welford in cell order)."""
    colptr = np.ascontiguousarray(colptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = colptr.size - 1
    sf = np.empty(n)
    logx = np.empty(x.size) if want_logx else None
    mean = np.empty(n_genes)
    var = np.empty(n_genes)
    rc = _lib().oracle_lognorm_genestats(
        n_genes, n, _p(colptr, ctypes.c_int64), _p(rowidx, ctypes.c_int32), _p(x, ctypes.c_double),
        _p(sf, ctypes.c_double), _p(logx, ctypes.c_double) if want_logx else None,
        _p(mean, ctypes.c_double), _p(var, ctypes.c_double))
    if rc == -4:
        raise ValueError("size factors should be positive")
    if rc != 0:
        raise RuntimeError(f"oracle_lognorm_genestats failed: {rc}")
    return sf, logx, mean, var


# ---------------------------------------------------------------------------------------
# K4: scran::fitTrendVar + modelGeneVar decomposition + getTopHVGs
#     (R/doNormAndReduce.R:88-90).  Approximate by
#     construction: stats::nls / limma::weightedLowess / stats::density are restated from
#     their documentation.  In the R product these stay unchanged R calls.
# ---------------------------------------------------------------------------------------
def _bw_nrd0(x):
    hi = np.std(x, ddof=1)
    q75, q25 = np.percentile(x, [75, 25])
    lo = min(hi, (q75 - q25) / 1.34)
    if lo == 0:
        lo = hi or abs(x[0]) or 1.0
    return 0.9 * lo * len(x) ** (-0.2)


def _inverse_density_weights(m, adjust=1.0):
    """scran:::.inverse_density_weights: w = 1/approx(density(m))(m), scaled to mean 1."""
    bw = _bw_nrd0(m) * adjust
    lo, hi = m.min() - 3 * bw, m.max() + 3 * bw
    grid = np.linspace(lo, hi, 512)
    # Gaussian KDE evaluated on R's default 512-point grid (R bins + FFT; direct sum here)
    dens = np.empty_like(grid)
    for a in range(0, 512, 64):
        z = (grid[a:a + 64, None] - m[None, :]) / bw
        dens[a:a + 64] = np.exp(-0.5 * z * z).sum(axis=1)
    dens /= len(m) * bw * np.sqrt(2 * np.pi)
    w = 1.0 / np.interp(m, grid, dens)
    return w / w.mean()


def _nls_starts(v, m, left_n=100, left_prop=0.1, grid_length=10, b_range=5, n_max=10):
    o = np.argsort(m, kind="stable")
    n = len(v)
    keep = o[:max(1, int(min(left_n, n * left_prop)))]
    grad = float(np.dot(m[keep], v[keep]) / np.dot(m[keep], m[keep]))  # lm(y ~ 0 + x)
    b_pts = 2.0 ** np.linspace(-b_range, b_range, grid_length)
    n_pts = 2.0 ** np.linspace(0, n_max, grid_length)
    best = (np.inf, None, None)
    for cur_n in n_pts:          # expand.grid(B=, n=): B varies fastest
        for cur_b in b_pts:
            with np.errstate(over="ignore"):
                rss = float(np.sum((v - grad * cur_b * m / (m ** cur_n + cur_b)) ** 2))
            if rss < best[0]:
                best = (rss, cur_b, cur_n)
    _, b, nn = best
    return dict(a=b * grad, b=b, n=nn)


def _nls_fit(v, m, w, start, maxiter=500, tol=1e-9):
    """Weighted least squares on v ~ exp(A) m / (m^(1+exp(N)) + exp(B)).

    stats::nls is Gauss-Newton with step halving and the relative-offset stopping rule
    (tol 1e-5); here the same normal equations are damped (Levenberg-Marquardt) and
    iterated far past R's tolerance so that two independent implementations land on the
    same optimum to ~1e-9."""
    th = np.array([np.log(start["a"]), np.log(start["b"]), np.log(max(start["n"] - 1.0, 1e-8))])
    lm = np.log(m)
    sw = np.sqrt(w)

    def model(t):
        ea, eb, en = np.exp(t[0]), np.exp(t[1]), np.exp(t[2])
        p = m ** (1.0 + en)
        den = p + eb
        f = ea * m / den
        jac = np.stack([f, -f * eb / den, -f * p * lm * en / den], axis=1)
        return f, jac

    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        f, jac = model(th)
    if not (np.all(np.isfinite(f)) and np.all(np.isfinite(jac))):
        return th  # scran keeps the grid-search start values when nls() cannot run
    rss = float(np.sum(w * (v - f) ** 2))
    damp = 1e-3
    for _ in range(maxiter):
        r = sw * (v - f)
        jw = sw[:, None] * jac
        q, _ = np.linalg.qr(jw)
        qtr = q.T @ r
        crit = np.sqrt(float(qtr @ qtr) / max(float(r @ r) - float(qtr @ qtr), 1e-300))
        if crit < tol:
            break
        a = jw.T @ jw
        g = jw.T @ r
        improved = False
        for _try in range(40):
            try:
                step = np.linalg.solve(a + damp * np.diag(np.diag(a) + 1e-300), g)
            except np.linalg.LinAlgError:
                damp *= 10.0
                continue
            with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
                f2, jac2 = model(th + step)
                rss2 = float(np.sum(w * (v - f2) ** 2))
            if np.isfinite(rss2) and rss2 < rss and np.all(np.isfinite(jac2)):
                th, f, jac, rss = th + step, f2, jac2, rss2
                damp = max(damp / 10.0, 1e-12)
                improved = True
                break
            damp *= 10.0
        if not improved:
            break  # no representable decrease left: converged to rounding level
    return th


def _weighted_median(x, w):
    o = np.argsort(x, kind="stable")
    cw = np.cumsum(w[o])
    half = cw[-1] / 2
    idx = int(np.searchsorted(cw, half, side="left"))
    if idx + 1 < len(x) and cw[idx] == half:
        return 0.5 * (x[o[idx]] + x[o[idx + 1]])
    return x[o[idx]]


def weighted_lowess(x, y, w, span=0.3, iterations=4, npts=200):
    """limma::weightedLowess restated (sorted-x local linear fits at ~npts seed points,
    tricube kernel, window sized by span * total weight, bisquare robustness)."""
    o = np.argsort(x, kind="stable")
    xs, ys, ws = x[o], y[o], w[o]
    n = len(xs)
    delta = (xs[-1] - xs[0]) / npts
    seeds = [0]
    last = 0
    for i in range(1, n - 1):
        if xs[i] - xs[last] > delta:
            seeds.append(i)
            last = i
    if n > 1:
        seeds.append(n - 1)
    seeds = np.array(seeds)
    span_w = span * ws.sum()
    cw = np.concatenate([[0.0], np.cumsum(ws)])
    # window limits per seed: grow towards the nearer neighbour until weight >= span_w
    lims = []
    for s in seeds:
        lo = hi = s
        cur = ws[s]
        while cur < span_w and (lo > 0 or hi < n - 1):
            if lo == 0:
                hi += 1
                cur += ws[hi]
            elif hi == n - 1:
                lo -= 1
                cur += ws[lo]
            elif xs[s] - xs[lo - 1] < xs[hi + 1] - xs[s]:
                lo -= 1
                cur += ws[lo]
            else:
                hi += 1
                cur += ws[hi]
        dist = max(xs[s] - xs[lo], xs[hi] - xs[s])
        while lo > 0 and xs[s] - xs[lo - 1] <= dist:
            lo -= 1
        while hi < n - 1 and xs[hi + 1] - xs[s] <= dist:
            hi += 1
        lims.append((lo, hi, dist))
    del cw
    rob = np.ones(n)
    fitted = np.empty(n)
    for it in range(iterations):
        sv = np.empty(len(seeds))
        for q, s in enumerate(seeds):
            lo, hi, dist = lims[q]
            xx, yy = xs[lo:hi + 1], ys[lo:hi + 1]
            if dist <= 0:
                kw = ws[lo:hi + 1] * rob[lo:hi + 1]
            else:
                u = np.abs(xx - xs[s]) / dist
                kw = np.where(u < 1, (1 - u ** 3) ** 3, 0.0) * ws[lo:hi + 1] * rob[lo:hi + 1]
            tw = kw.sum()
            if tw <= 0:
                sv[q] = ys[s]
                continue
            xm = (kw * xx).sum() / tw
            ym = (kw * yy).sum() / tw
            var = (kw * (xx - xm) ** 2).sum()
            if var < 1e-12 * max(dist * dist, 1e-300) * tw:
                sv[q] = ym
            else:
                slope = (kw * (xx - xm) * (yy - ym)).sum() / var
                sv[q] = ym + slope * (xs[s] - xm)
        fitted = np.interp(xs, xs[seeds], sv)
        if it == iterations - 1:
            break
        res = np.abs(ys - fitted)
        cmad = 6.0 * _weighted_median(res, ws)
        if cmad <= 1e-12 * max(res.mean(), 1e-300):
            break
        rob = np.where(res < cmad, (1 - (res / cmad) ** 2) ** 2, 0.0)
    out = np.empty(n)
    out[o] = fitted
    return out


@dataclass
class TrendFit:
    trend: object
    std_dev: float


def fit_trend_var(means, vars_, min_mean=0.1, parametric=True, lowess=True, density_weights=True):
    """scran::fitTrendVar(means, vars) with the defaults modelGeneVar uses (row a5)."""
    means = np.asarray(means, dtype=np.float64)
    vars_ = np.asarray(vars_, dtype=np.float64)
    ok = np.isfinite(vars_) & (vars_ > 1e-8) & (means >= min_mean)
    v, m = vars_[ok], means[ok]
    if len(v) < 2:
        raise ValueError("need at least 2 points for non-parametric curve fitting")
    w = _inverse_density_weights(m) if density_weights else np.ones_like(m)
    to_fit = np.log(v)
    left_edge = m.min()

    def paramfun(x):
        return np.minimum(1.0, x / left_edge)

    if parametric:
        if len(v) <= 3:
            raise ValueError("need at least 4 points for non-linear curve fitting")
        th = _nls_fit(v, m, w, _nls_starts(v, m))
        ea, eb, en = np.exp(th)

        def paramfun(x):  # noqa: F811
            return ea * x / (x ** (1.0 + en) + eb)
        to_fit = to_fit - np.log(paramfun(m))
    if lowess:
        lfit = weighted_lowess(m, to_fit, w, span=0.3)
        o = np.argsort(m, kind="stable")
        mx, ly = m[o], lfit[o]
        # approxfun(m, fitted, rule=2); ties averaged like approxfun's default ties=mean
        ux, inv = np.unique(mx, return_inverse=True)
        uy = np.bincount(inv, weights=ly) / np.bincount(inv)

        def unscaled(x):
            return np.exp(np.interp(x, ux, uy)) * paramfun(x)
    else:
        unscaled = paramfun
    leftovers = v / unscaled(m)
    med = _weighted_median(leftovers, w)
    std_dev = float(_weighted_median(np.abs(leftovers / med - 1.0), w) * 1.4826)

    def trend(x):
        x = np.asarray(x, dtype=np.float64)
        out = unscaled(np.maximum(x, 0.0)) * med
        return np.where(x > 0, out, 0.0) if np.ndim(out) else out
    return TrendFit(trend=trend, std_dev=std_dev)


def _pnorm_upper(z):
    from scipy.special import ndtr
    return ndtr(-z)


def model_gene_var(means, vars_):
    """scran::modelGeneVar decomposition: tech = trend(mean), bio = total - tech,
    p.value = pnorm(total/tech, mean=1, sd=std.dev, lower=FALSE), FDR = BH."""
    fit = fit_trend_var(means, vars_)
    tech = fit.trend(means)
    bio = vars_ - tech
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = vars_ / tech
    p = _pnorm_upper((ratio - 1.0) / fit.std_dev)
    p = np.where(np.isfinite(ratio), p, np.nan)
    fdr = np.full_like(p, np.nan)
    okp = np.isfinite(p)
    if okp.any():
        pv = p[okp]
        o = np.argsort(pv, kind="stable")[::-1]
        n = len(pv)
        adj = np.minimum.accumulate(pv[o] * n / np.arange(n, 0, -1))
        tmp = np.empty(n)
        tmp[o] = np.minimum(adj, 1.0)
        fdr[okp] = tmp
    return dict(mean=means, total=vars_, tech=tech, bio=bio, p_value=p, FDR=fdr, std_dev=fit.std_dev)


# ---------------------------------------------------------------------------------------
# batch-aware branch (R/doNormAndReduce.R:74-76, :84-86) -- batchelor::multiBatchNorm and
# scran::modelGeneVar(block=), restated from the packages' definitions (un-vendored: parity unpinned)
# ---------------------------------------------------------------------------------------
def batch_averages(n_genes, colptr, rowidx, x, sf, block, n_blocks):
    """Per batch: scuttle::calculateAverage(counts_b, size.factors = sf_b) = rowMeans(t(t(counts_b) / (sf_b / mean(sf_b))))."""
    colptr = np.asarray(colptr, dtype=np.int64)
    col_of = np.repeat(np.arange(colptr.size - 1), np.diff(colptr))
    ave = np.zeros((n_blocks, n_genes))
    for b in range(n_blocks):
        cells = np.flatnonzero(block == b)
        if cells.size == 0:
            continue
        sfc = sf / sf[cells].mean()
        sel = block[col_of] == b
        ave[b] = np.bincount(rowidx[sel], weights=x[sel] / sfc[col_of[sel]], minlength=n_genes) / cells.size
    return ave


def rescale_size_factors(ave, sf, block, min_mean=1.0):
    """batchelor:::.rescale_size_factors: pairwise median ratio of the batch averages over the genes whose grand mean
    is >= min_mean; every batch is scaled down to the batch of lowest coverage.  Returns the per-cell factors
    multiBatchNorm hands to logNormCounts(center.size.factors = FALSE)."""
    nb = ave.shape[0]
    ratios = np.ones((nb, nb))
    for first in range(nb - 1):
        fa = ave[first]
        fs = fa.sum()
        for second in range(first + 1, nb):
            sa = ave[second]
            ss = sa.sum()
            grand = (fa / fs + sa / ss) / 2 * (fs + ss) / 2
            keep = grand >= min_mean
            with np.errstate(divide="ignore", invalid="ignore"):
                cur = np.median(sa[keep] / fa[keep])
            if not np.isfinite(cur) or cur == 0:
                raise ValueError("median ratio of averages between batches is not finite")
            ratios[first, second] = cur
            ratios[second, first] = 1.0 / cur
    smallest = int(np.argmin(ratios.min(axis=0)))
    out = np.empty_like(sf)
    for b in range(nb):
        cells = block == b
        out[cells] = sf[cells] / sf[cells].mean() / ratios[b, smallest]
    return out


def gene_mean_var_blocked(n_genes, colptr, rowidx, y, block, n_blocks):
    colptr = np.asarray(colptr, dtype=np.int64)
    col_of = np.repeat(np.arange(colptr.size - 1), np.diff(colptr))
    mean = np.full((n_blocks, n_genes), np.nan)
    var = np.full((n_blocks, n_genes), np.nan)
    for b in range(n_blocks):
        n = int((block == b).sum())
        if n == 0:
            continue
        sel = block[col_of] == b
        m, v = gene_mean_var(n_genes, n, rowidx[sel], y[sel])
        mean[b] = m
        if n >= 2:
            var[b] = v
    return mean, var


def model_gene_var_blocked(means, vars_, n_per_block):
    """scran::modelGeneVar(block=): a trend per block, then the unweighted average of the per-block mean / total /
    tech / bio over the blocks with at least 2 cells (equiweight = TRUE); getTopHVGs ranks by that `bio`."""
    use = [b for b in range(means.shape[0]) if n_per_block[b] >= 2]
    per = [model_gene_var(means[b], vars_[b]) for b in use]
    out = {k: np.mean([p[k] for p in per], axis=0) for k in ("mean", "total", "tech", "bio")}
    out["per_block"] = per
    return out


# ---------------------------------------------------------------------------------------
# deconvolution size factors (R/doNormAndReduce.R:66-71 -> scran::computeSumFactors =
# scuttle::pooledSizeFactors), restated from the published algorithm (Lun, Bach & Marioni 2016, Genome Biol. 17:75,
# and the package sources as recalled: un-vendored, parity unpinned).  quickCluster's walktrap stays igraph;
# the clusters are an input here.
# ---------------------------------------------------------------------------------------
POOL_SIZES = tuple(range(21, 102, 5))
POOL_LOWWEIGHT = 1e-6


def limit_cluster_size(clusters, max_size=3000):
    """scuttle:::.limit_cluster_size: a cluster with more than max_size cells is dealt round-robin into
    ceiling(n / max_size) sub-clusters; new ids are numbered in order of first appearance."""
    clusters = np.asarray(clusters)
    out = np.zeros(clusters.size, dtype=np.int32)
    counter = 0
    _, first = np.unique(clusters, return_index=True)
    for cid in clusters[np.sort(first)]:
        cur = np.flatnonzero(clusters == cid)
        if max_size is None or cur.size <= max_size:
            out[cur] = counter
            counter += 1
            continue
        mult = -(-cur.size // max_size)
        out[cur] = counter + (np.arange(cur.size) % mult)
        counter += mult
    return out, counter


def guess_min_mean(libsizes):
    """scuttle:::.guessMinMean: UMI-like data (median library size <= 50000) -> 0.1, else 1."""
    return 0.1 if float(np.median(libsizes)) <= 50000 else 1.0


def pool_ring(libsizes):
    """scuttle:::.generateSphere: cells by ascending library size (stable), odd ranks ascending then even ranks
    descending, so that neighbours on the ring have similar library sizes."""
    o = np.argsort(libsizes, kind="stable")
    return np.concatenate([o[0::2], o[1::2][::-1]])


def pooled_factors_one_cluster(dense_counts, sizes=POOL_SIZES, min_mean=0.1):
    """scuttle:::.per_cluster_normalize on a dense genes x cells block.  Returns (final.nf, ave.cell)."""
    x = np.asarray(dense_counts, dtype=np.float64)
    n = x.shape[1]
    lib = x.sum(axis=0)
    if np.any(lib == 0):
        raise ValueError("cells should have non-zero library sizes")
    y = x / lib
    ave = y.mean(axis=1) * lib.mean()
    keep = ave >= min_mean
    yk, avek = y[keep], ave[keep]
    ring = pool_ring(lib)
    sizes = [int(s) for s in sorted(sizes) if s <= n]
    if not sizes:
        raise ValueError("not enough cells in at least one cluster for some 'sizes'")
    rows, cols, vals, rhs = [], [], [], []
    r = 0
    for s in sizes:
        for i in range(n):
            members = ring[(i + np.arange(s)) % n]
            pooled = np.zeros(yk.shape[0])
            for c in members:                     # sequential fp64 sum in ring order (pool_size_factors C++)
                pooled = pooled + yk[:, c]
            rhs.append(np.median(pooled / avek))
            rows += [r] * s
            cols += members.tolist()
            vals += [1.0] * s
            r += 1
    lw = np.sqrt(POOL_LOWWEIGHT)
    for i in range(n):
        c = ring[i]
        rhs.append(lw * np.median(yk[:, c] / avek))
        rows.append(r)
        cols.append(int(c))
        vals.append(lw)
        r += 1
    design = np.zeros((r, n))
    design[rows, cols] = vals
    nf, *_ = np.linalg.lstsq(design, np.asarray(rhs), rcond=None)   # Matrix::qr + qr.coef in the package
    return nf * lib, ave


def rescale_clusters(profiles, ref, min_mean):
    """scuttle:::.rescale_clusters: median ratio of every cluster's pseudo-cell to the reference cluster's over the
    genes whose grand mean passes min_mean."""
    out = np.zeros(len(profiles))
    refp = profiles[ref]
    for c, cur in enumerate(profiles):
        cl, rl = cur.sum(), refp.sum()
        use = (cur / cl + refp / rl) / 2 * (cl + rl) / 2 >= min_mean
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = cur[use] / refp[use]
        val = float(np.median(ratio[~np.isnan(ratio)])) if np.any(~np.isnan(ratio)) else np.nan
        if not np.isfinite(val) or val <= 0:
            raise ValueError("inter-cluster rescaling factor is not strictly positive")
        out[c] = val
    return out


def pooled_size_factors(n_genes, colptr, rowidx, x, clusters=None, sizes=POOL_SIZES, min_mean=None,
                        max_cluster_size=3000):
    """scuttle::pooledSizeFactors(counts, clusters = ...) for a CSC count matrix: per-cluster pooled factors, then
    rescaling between clusters (reference = the cluster whose pseudo-cell has most non-zero genes), scaled to unit
    mean over the positive factors.  No cleanSizeFactors step: non-positive factors are returned as they are."""
    colptr = np.asarray(colptr, dtype=np.int64)
    n = colptr.size - 1
    clusters = np.zeros(n, dtype=np.int32) if clusters is None else np.asarray(clusters)
    cl, ncl = limit_cluster_size(clusters, max_cluster_size)
    lib_all = np.add.reduceat(np.asarray(x, dtype=np.float64), colptr[:-1]) if n else np.zeros(0)
    lib_all[np.diff(colptr) == 0] = 0.0
    if min_mean is None:
        min_mean = guess_min_mean(lib_all)
    min_mean = max(min_mean, 1e-8)
    nfs, profiles, members = [], [], []
    for c in range(ncl):
        cells = np.flatnonzero(cl == c)
        dense = np.zeros((n_genes, cells.size))
        for j, cell in enumerate(cells):
            sl = slice(colptr[cell], colptr[cell + 1])
            dense[rowidx[sl], j] = x[sl]
        nf, ave = pooled_factors_one_cluster(dense, sizes, min_mean)
        nfs.append(nf)
        profiles.append(ave)
        members.append(cells)
    ref = int(np.argmax([int((p > 0).sum()) for p in profiles]))
    resc = rescale_clusters(profiles, ref, min_mean)
    sf = np.full(n, np.nan)
    for c in range(ncl):
        sf[members[c]] = nfs[c] * resc[c]
    pos = sf > 0
    return sf / sf[pos].mean()


# ---------------------------------------------------------------------------------------
# scuttle::perCellQCMetrics as doQC calls it (R/doQC.R:106-108); findKNN(get.distance = TRUE)
# ---------------------------------------------------------------------------------------
def per_cell_qc_metrics(n_genes, colptr, rowidx, x, subsets=(), threshold=0.0):
    """sum, detected (count > threshold) and per subset: sum, detected, percent = 100 * subset sum / sum."""
    colptr = np.asarray(colptr, dtype=np.int64)
    n = colptr.size - 1
    col_of = np.repeat(np.arange(n), np.diff(colptr))
    x = np.asarray(x, dtype=np.float64)
    out = dict(sum=np.bincount(col_of, weights=x, minlength=n),
               detected=np.bincount(col_of, weights=(x > threshold), minlength=n).astype(np.int32), subsets=[])
    for genes in subsets:
        m = np.zeros(n_genes, dtype=bool)
        m[np.asarray(genes, dtype=np.int64)] = True
        sel = m[rowidx]
        ssum = np.bincount(col_of[sel], weights=x[sel], minlength=n)
        sdet = np.bincount(col_of[sel], weights=(x[sel] > threshold), minlength=n).astype(np.int32)
        with np.errstate(divide="ignore", invalid="ignore"):
            out["subsets"].append(dict(sum=ssum, detected=sdet, percent=100.0 * ssum / out["sum"]))
    return out


def knn_distances(emb, index1):
    """Distances of the neighbours in a 1-based index matrix, in the search's arithmetic (sequential fp64)."""
    emb = np.asarray(emb, dtype=np.float64)
    n, k = index1.shape
    out = np.empty((n, k))
    for c in range(k):
        diff = emb - emb[index1[:, c] - 1]
        s = np.zeros(n)
        for t in range(emb.shape[1]):
            s = s + diff[:, t] * diff[:, t]
        out[:, c] = np.sqrt(s)
    return out


def get_top_hvgs(bio, n):
    """scran::getTopHVGs(stats, n=n): keep bio > 0, order(bio, decreasing=TRUE) (ties by
    row index), first n.  Returns 0-based row indices (row a6)."""
    bio = np.asarray(bio)
    keep = np.flatnonzero(np.isfinite(bio) & (bio > 0))
    o = keep[np.argsort(-bio[keep], kind="stable")]
    return o[:n].astype(np.int32)


# ---------------------------------------------------------------------------------------
# K5: scater::runPCA(subset_row = hvgs, exprs_values = "logcounts", ncomponents = d)
#     (R/doNormAndReduce.R:94-97)
# ---------------------------------------------------------------------------------------
def run_pca(n_genes, colptr, rowidx, logx, hvg_idx, d):
    """Exact truncated SVD of the column-centred (N x H) HVG log-count matrix.

    Returns scores (N x d) = U D, rotation (H x d) = V, var_explained = D^2/(N-1),
    percent_var = 100 var_explained / sum(column variances)."""
    import scipy.sparse as sp
    colptr = np.asarray(colptr, dtype=np.int64)
    n = colptr.size - 1
    mat = sp.csc_matrix((logx, np.asarray(rowidx, dtype=np.int64), colptr), shape=(n_genes, n))
    m = np.asarray(mat.tocsr()[np.asarray(hvg_idx, dtype=np.int64), :].T.todense(), dtype=np.float64)
    mu = m.mean(axis=0)
    xc = m - mu
    total_var = float((xc * xc).sum() / (n - 1))
    if n >= xc.shape[1]:
        # eigen-decomposition of the H x H cross-product in fp64 would square the
        # condition number; use LAPACK SVD (what BiocSingular::ExactParam does).
        u, s, vt = np.linalg.svd(xc, full_matrices=False)
    else:
        u, s, vt = np.linalg.svd(xc, full_matrices=False)
    scores = u[:, :d] * s[:d]
    rotation = vt[:d].T
    var_expl = s[:d] ** 2 / (n - 1)
    return dict(scores=scores, rotation=rotation, var_explained=var_expl,
                percent_var=100.0 * var_expl / total_var, total_var=total_var, center=mu)


# ---------------------------------------------------------------------------------------
# K6: BiocNeighbors::findKNN  
# ---------------------------------------------------------------------------------------
def find_knn(emb, k, want_dist=False):
    """Exact kNN through oracle/knn_snn.c.  emb: N x d; returns N x k 1-based int32."""
    emb_f = np.asfortranarray(emb, dtype=np.float64)
    n, d = emb_f.shape
    idx = np.empty((n, k), dtype=np.int32, order="F")
    dist = np.empty((n, k), dtype=np.float64, order="F") if want_dist else None
    rc = _lib().oracle_knn(_p(emb_f, ctypes.c_double), n, d, k, _p(idx, ctypes.c_int32),
                           _p(dist, ctypes.c_double) if want_dist else None)
    if rc != 0:
        raise ValueError(f"oracle_knn failed ({rc}): need 1 <= k <= N-1")
    return (idx, dist) if want_dist else idx


def find_knn_queries(emb, queries, k):
    """Exact kNN (same arithmetic / tie rule) of the 0-based rows `queries` only: len(queries) x k, 1-based."""
    emb_f = np.asfortranarray(emb, dtype=np.float64)
    n, d = emb_f.shape
    q = np.ascontiguousarray(queries, dtype=np.int64)
    idx = np.empty((q.size, k), dtype=np.int32)
    rc = _lib().oracle_knn_queries(_p(emb_f, ctypes.c_double), n, d, k, _p(q, ctypes.c_int64), q.size,
                                   _p(idx, ctypes.c_int32), None)
    if rc != 0:
        raise ValueError(f"oracle_knn_queries failed ({rc})")
    return idx


def find_knn_numpy(emb, k):
    """Same contract in numpy, for cross-checking the C oracle on small inputs:
    sequential fp64 accumulation (numpy never fuses multiply-add), sqrt, (dist, idx) order."""
    emb = np.asarray(emb, dtype=np.float64)
    n, d = emb.shape
    acc = np.zeros((n, n))
    for t in range(d):
        delta = emb[:, t][:, None] - emb[:, t][None, :]
        acc += delta * delta
    dist = np.sqrt(acc)
    out = np.empty((n, k), dtype=np.int32)
    for i in range(n):
        cand = np.array([j for j in range(n) if j != i])
        o = np.lexsort((cand, dist[i, cand]))
        out[i] = cand[o[:k]] + 1
    return out


# ---------------------------------------------------------------------------------------
# K7: bluster::neighborsToSNNGraph  
# ---------------------------------------------------------------------------------------
SNN_TYPES = {"rank": 0, "number": 1, "jaccard": 2}


def snn_edges(index, snn_type="rank", j_begin=0, j_end=None):
    """index: N x k 1-based.  Returns (from, to, weight) in emission order (1-based); j_begin / j_end restrict
    the emitting cells to the 0-based range [j_begin, j_end) (the edges of a cell depend on the lists only)."""
    idx = np.asfortranarray(index, dtype=np.int32)
    n, k = idx.shape
    j_end = n if j_end is None else j_end
    ne = ctypes.c_int64(0)
    pf, pt = ctypes.POINTER(ctypes.c_int32)(), ctypes.POINTER(ctypes.c_int32)()
    pw = ctypes.POINTER(ctypes.c_double)()
    rc = _lib().oracle_snn_range(_p(idx, ctypes.c_int32), n, k, SNN_TYPES[snn_type], j_begin, j_end, ctypes.byref(ne),
                                 ctypes.byref(pf), ctypes.byref(pt), ctypes.byref(pw))
    if rc != 0:
        raise ValueError(f"oracle_snn failed ({rc})")
    m = ne.value
    f = np.ctypeslib.as_array(pf, shape=(max(m, 1),))[:m].copy()
    t = np.ctypeslib.as_array(pt, shape=(max(m, 1),))[:m].copy()
    w = np.ctypeslib.as_array(pw, shape=(max(m, 1),))[:m].copy()
    for ptr in (pf, pt, pw):
        _lib().oracle_free(ctypes.cast(ptr, ctypes.c_void_p))
    return f, t, w


def snn_edges_python(index, snn_type="rank"):
    """Brute-force definition (all pairs), for cross-checking on tiny inputs.
    Returns a dict {(max, min) 1-based: weight}."""
    idx = np.asarray(index)
    n, k = idx.shape
    lists = [[j] + [int(v) - 1 for v in idx[j]] for j in range(n)]
    out = {}
    for j in range(n):
        for h in range(j):
            shared = set(lists[j]) & set(lists[h])
            if not shared:
                continue
            if snn_type == "rank":
                r = min(lists[j].index(s) + lists[h].index(s) for s in shared)
                wgt = k - r / 2.0
            elif snn_type == "number":
                wgt = float(len(shared))
            else:
                wgt = len(shared) / (2.0 * (k + 1) - len(shared))
            out[(j + 1, h + 1)] = max(wgt, 1e-6)
    return out


def canonical_edges(frm, to, w):
    """(min, max)-sorted edge list, the parity comparison form."""
    a = np.minimum(frm, to).astype(np.int64)
    b = np.maximum(frm, to).astype(np.int64)
    o = np.lexsort((b, a))
    return a[o], b[o], np.asarray(w)[o]


# ---------------------------------------------------------------------------------------
# whole path (for bench cpu_baseline and the end-to-end parity tests)
# ---------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------
# per-resolution cluster diagnostics (R/makeGraphsAndClusters.R:103-114)
#   bluster::approxSilhouette, bluster::pairwiseModularity -- restated from the package's
#   definitions (bluster 1.16); bluster is not vendored in the reference: parity unpinned.
# ---------------------------------------------------------------------------------------
def approx_silhouette(x, clusters):
    """bluster::approxSilhouette(x, clusters).  Returns (other, width): for every cell the closest
    other cluster (a label) and (other.dist - self.dist) / pmax(other.dist, self.dist), where the
    distance from cell i to cluster u is sqrt(|x_i - centroid_u|^2 + sum_dims mean sq. deviation of u)."""
    x = np.asarray(x, dtype=np.float64)
    clusters = np.asarray(clusters)
    uclust = np.unique(clusters)                       # sort(unique(clusters))
    cent, cvar = [], []
    for u in uclust:
        cur = x[clusters == u]
        c = cur.mean(axis=0)
        cent.append(c)
        cvar.append(float(((cur - c) ** 2).mean(axis=0).sum()))
    n = x.shape[0]
    self_d = np.full(n, np.inf)
    other_d = np.full(n, np.inf)
    other_c = np.full(n, -1, dtype=np.int64)
    for i, u in enumerate(uclust):
        D = np.sqrt(((x - cent[i]) ** 2).sum(axis=1) + cvar[i])
        is_self = clusters == u
        self_d[is_self] = D[is_self]
        better = (~is_self) & (D < other_d)
        other_d[better] = D[better]
        other_c[better] = i
    with np.errstate(invalid="ignore"):
        width = (other_d - self_d) / np.maximum(other_d, self_d)
    other = np.where(other_c >= 0, uclust[np.maximum(other_c, 0)], uclust[0])
    return other, width, other_c


def pairwise_modularity(frm, to, weight, clusters, get_weights=False, as_ratio=False):
    """bluster::pairwiseModularity(graph, clusters, get.weights, as.ratio) on an undirected weighted edge
    list (1-based).  observed: upper-triangular K x K, [x, y] = weight of the edges between clusters x and y
    (diagonal: within); expected: W p_x^2 on the diagonal, 2 W p_x p_y above it, p = share of the weighted
    degree; the lower triangles are 0 (so as_ratio gives NaN there, as in R)."""
    clusters = np.asarray(clusters)
    uclust = np.unique(clusters)
    K = uclust.size
    lab = np.searchsorted(uclust, clusters)
    a = lab[np.asarray(frm, dtype=np.int64) - 1]
    b = lab[np.asarray(to, dtype=np.int64) - 1]
    w = np.asarray(weight, dtype=np.float64)
    obs = np.zeros((K, K))
    np.add.at(obs, (np.minimum(a, b), np.maximum(a, b)), w)
    deg = np.zeros(K)
    np.add.at(deg, a, w)
    np.add.at(deg, b, w)
    W = float(w.sum())
    prop = deg / deg.sum() if deg.sum() > 0 else deg
    expd = np.outer(prop, prop) * W
    expd[np.tril_indices(K, -1)] = 0.0
    expd[np.triu_indices(K, 1)] *= 2.0
    if get_weights:
        return dict(observed=obs, expected=expd)
    with np.errstate(invalid="ignore", divide="ignore"):
        return obs / expd if as_ratio else (obs - expd) / W


def pipeline(n_genes, colptr, rowidx, x, ndims=50, hvg_ntop=2000, neighbors=20, snn_type="rank",
             fast_pca=False):
    """counts -> SNN graph through the restated steps, in the reference's call order."""
    sf = library_size_factors(colptr, x)
    logx = log_norm_counts(colptr, x, sf)
    n = len(colptr) - 1
    mean, var = gene_mean_var(n_genes, n, rowidx, logx)
    stats = model_gene_var(mean, var)
    hvgs = get_top_hvgs(stats["bio"], hvg_ntop)
    if fast_pca:
        pca = run_pca_irlba_like(n_genes, colptr, rowidx, logx, hvgs, ndims)
    else:
        pca = run_pca(n_genes, colptr, rowidx, logx, hvgs, ndims)
    knn = find_knn(pca["scores"], neighbors)
    frm, to, w = snn_edges(knn, snn_type)
    return dict(sf=sf, logx=logx, mean=mean, var=var, stats=stats, hvgs=hvgs, pca=pca, knn=knn,
                edges=(frm, to, w))


def run_pca_irlba_like(n_genes, colptr, rowidx, logx, hvg_idx, d):
    """Truncated SVD by ARPACK on the implicitly centred sparse matrix -- the CPU-baseline
    stand-in for BiocSingular::IrlbaParam (Lanczos bidiagonalisation)."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import LinearOperator, svds
    colptr = np.asarray(colptr, dtype=np.int64)
    n = colptr.size - 1
    mat = sp.csc_matrix((logx, np.asarray(rowidx, dtype=np.int64), colptr), shape=(n_genes, n))
    m = mat.tocsr()[np.asarray(hvg_idx, dtype=np.int64), :].T.tocsr()  # N x H
    mu = np.asarray(m.mean(axis=0)).ravel()
    ones = np.ones(n)
    op = LinearOperator(m.shape, dtype=np.float64,
                        matvec=lambda v: m @ v - ones * float(mu @ np.ravel(v)),
                        rmatvec=lambda u: m.T @ u - mu * float(np.ravel(u).sum()))
    rng = np.random.default_rng(0)
    u, s, vt = svds(op, k=d, v0=rng.standard_normal(min(m.shape)), tol=1e-10)
    o = np.argsort(-s)
    u, s, vt = u[:, o], s[o], vt[o]
    sq = np.asarray(m.multiply(m).sum(axis=0)).ravel()
    total_var = float((sq - n * mu * mu).sum() / (n - 1))
    var_expl = s ** 2 / (n - 1)
    return dict(scores=u * s, rotation=vt.T, var_explained=var_expl,
                percent_var=100.0 * var_expl / total_var, total_var=total_var, center=mu)
