// cb200_host.cu -- host-side helper of the Python mirror (no device code): the weighted lowess
// that scran::fitTrendVar runs on the per-gene (mean, log-variance residual) points.  In the R
// product this stays limma::weightedLowess inside the unchanged scran call; the Python mirror
// (cellula_b200/trendvar.py) has no R and would otherwise spend ~0.2 s per call in interpreted loops.
// Same algorithm as trendvar.py's reference loops: local linear fits at ~npts seed points over the
// sorted x, window = nearest points holding span * total weight, tricube kernel, bisquare
// robustness iterations, linear interpolation between the seeds.
#include <algorithm>
#include <cmath>
#include <thread>
#include <functional>
#include <atomic>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../include/cellula_b200.h"

// static split of [0, n) over a few host threads (G-sized loops of the trend fit; every index is handled by
// exactly one thread with the same arithmetic as the serial loop, so results do not depend on the thread count)
template <typename Fn>
static void host_parallel_for(int n, Fn fn)
{
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw > 8 ? 8 : (hw < 1 ? 1 : hw));
    if (nt > n)
        nt = n;
    if (n < 8 || nt <= 1) {
        for (int i = 0; i < n; ++i)
            fn(i);
        return;
    }
    std::vector<std::thread> th;
    const int per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const int a = t * per, b = std::min(n, a + per);
        if (a >= b)
            break;
        th.emplace_back([=] {
            for (int i = a; i < b; ++i)
                fn(i);
        });
    }
    for (auto &t : th)
        t.join();
}

static double weighted_median(const std::vector<double> &x, const double *w, std::vector<int> &order)
{
    const int n = (int)x.size();
    order.resize(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return x[a] < x[b]; });
    double total = 0.0;
    for (int i = 0; i < n; ++i)
        total += w[order[i]];
    const double half = 0.5 * total;
    double cw = 0.0;
    for (int i = 0; i < n; ++i) {
        cw += w[order[i]];
        if (cw >= half) {
            if (cw == half && i + 1 < n)
                return 0.5 * (x[order[i]] + x[order[i + 1]]);
            return x[order[i]];
        }
    }
    return x[order[n - 1]];
}

// xs must be sorted ascending.  fitted_out[n].  Returns 0.
extern "C" int cb200_host_weighted_lowess(int64_t n64, const double *xs, const double *ys, const double *ws, double span,
                                          int iterations, int npts, double *fitted_out)
{
    const int n = (int)n64;
    if (n < 1 || xs == nullptr || ys == nullptr || ws == nullptr || fitted_out == nullptr)
        return 1;
    const double delta = (xs[n - 1] - xs[0]) / npts;
    std::vector<int> seeds(1, 0);
    for (int i = 1; i < n - 1; ++i)
        if (xs[i] - xs[seeds.back()] > delta)
            seeds.push_back(i);
    if (n > 1)
        seeds.push_back(n - 1);
    const int ns = (int)seeds.size();
    double total = 0.0;
    for (int i = 0; i < n; ++i)
        total += ws[i];
    const double target = span * total;
    std::vector<int> lo(ns), hi(ns);
    std::vector<double> dist(ns);
    for (int q = 0; q < ns; ++q) {
        const int s = seeds[q];
        int l = s, h = s;
        double acc = ws[s];
        while (acc < target && (l > 0 || h < n - 1)) {
            const bool go_left = h == n - 1 || (l > 0 && xs[s] - xs[l - 1] < xs[h + 1] - xs[s]);
            if (go_left)
                acc += ws[--l];
            else
                acc += ws[++h];
        }
        const double d = std::max(xs[s] - xs[l], xs[h] - xs[s]);
        while (l > 0 && xs[s] - xs[l - 1] <= d)
            --l;
        while (h < n - 1 && xs[h + 1] - xs[s] <= d)
            ++h;
        lo[q] = l;
        hi[q] = h + 1;
        dist[q] = d;
    }
    std::vector<double> rob(n, 1.0), sv(ns), res(n);
    std::vector<int> order;
    for (int i = 0; i < n; ++i)
        fitted_out[i] = ys[i];
    for (int it = 0; it < iterations; ++it) {
        host_parallel_for(ns, [&](int q) {
            const int s = seeds[q];
            const double d = dist[q];
            double tw = 0.0, sx = 0.0, sy = 0.0;
            for (int i = lo[q]; i < hi[q]; ++i) {
                double k = ws[i] * rob[i];
                if (d > 0.0) {
                    const double u = std::fabs(xs[i] - xs[s]) / d;
                    const double t = 1.0 - u * u * u;
                    k *= (u < 1.0) ? t * t * t : 0.0;
                }
                tw += k;
                sx += k * xs[i];
                sy += k * ys[i];
            }
            if (!(tw > 0.0)) {
                sv[q] = ys[s];
                return;
            }
            const double xm = sx / tw, ym = sy / tw;
            double sxx = 0.0, sxy = 0.0;
            for (int i = lo[q]; i < hi[q]; ++i) {
                double k = ws[i] * rob[i];
                if (d > 0.0) {
                    const double u = std::fabs(xs[i] - xs[s]) / d;
                    const double t = 1.0 - u * u * u;
                    k *= (u < 1.0) ? t * t * t : 0.0;
                }
                const double dx = xs[i] - xm;
                sxx += k * dx * dx;
                sxy += k * dx * (ys[i] - ym);
            }
            if (sxx < 1e-12 * std::max(d * d, 1e-300) * tw)
                sv[q] = ym;
            else
                sv[q] = ym + sxy / sxx * (xs[s] - xm);
        });
        // linear interpolation between the seeds (xs[seeds] is strictly increasing except for ties)
        int q = 0;
        for (int i = 0; i < n; ++i) {
            while (q + 1 < ns - 1 && xs[i] > xs[seeds[q + 1]])
                ++q;
            const double x0 = xs[seeds[q]], x1 = xs[seeds[std::min(q + 1, ns - 1)]];
            if (ns == 1 || x1 <= x0 || xs[i] <= x0)
                fitted_out[i] = (xs[i] <= x0 || ns == 1) ? sv[q] : sv[std::min(q + 1, ns - 1)];
            else if (xs[i] >= x1)
                fitted_out[i] = sv[q + 1];
            else
                fitted_out[i] = sv[q] + (sv[q + 1] - sv[q]) * (xs[i] - x0) / (x1 - x0);
        }
        if (it == iterations - 1)
            break;
        double rmean = 0.0;
        for (int i = 0; i < n; ++i) {
            res[i] = std::fabs(ys[i] - fitted_out[i]);
            rmean += res[i];
        }
        rmean /= n;
        const double cmad = 6.0 * weighted_median(res, ws, order);
        if (cmad <= 1e-12 * std::max(rmean, 1e-300))
            break;
        for (int i = 0; i < n; ++i) {
            const double r = res[i] / cmad;
            rob[i] = res[i] < cmad ? (1.0 - r * r) * (1.0 - r * r) : 0.0;
        }
    }
    return 0;
}

// Gaussian kernel density of m[0..n) on a uniform grid of `ng` points from lo to hi with bandwidth bw
// (direct sum, what trendvar._density_weights evaluates with numpy): dens_out[ng].
extern "C" int cb200_host_gauss_kde(int64_t n, const double *m, double bw, double lo, double hi, int ng,
                                    double *dens_out)
{
    if (n < 1 || ng < 2 || m == nullptr || dens_out == nullptr || !(bw > 0.0))
        return 1;
    const double step = (hi - lo) / (ng - 1);
    std::vector<double> acc((size_t)ng, 0.0);
    const double inv = 1.0 / bw;
    // grid ranges of every data point (exp(-0.5 z^2) < 2e-16 beyond |z| = 8.5: those terms cannot change a
    // double-precision sum), then one thread per slice of the grid, data points in ascending index order
    std::vector<int> r0((size_t)n), r1((size_t)n);
    for (int64_t j = 0; j < n; ++j) {
        int i0 = (int)std::floor((m[j] - 8.5 * bw - lo) / step), i1 = (int)std::ceil((m[j] + 8.5 * bw - lo) / step);
        r0[(size_t)j] = i0 < 0 ? 0 : i0;
        r1[(size_t)j] = i1 > ng - 1 ? ng - 1 : i1;
    }
    host_parallel_for(ng, [&](int i) {
        double a = 0.0;
        const double gx = lo + step * i;
        for (int64_t j = 0; j < n; ++j) {
            if (i < r0[(size_t)j] || i > r1[(size_t)j])
                continue;
            const double z = (gx - m[j]) * inv;
            a += std::exp(-0.5 * z * z);
        }
        acc[(size_t)i] = a;
    });
    const double norm = 1.0 / ((double)n * bw * 2.5066282746310002);  // sqrt(2 pi)
    for (int i = 0; i < ng; ++i)
        dens_out[i] = acc[i] * norm;
    return 0;
}

// Weighted least squares of v ~ exp(A) m / (m^(1+exp(N)) + exp(B)) (the parametric part of
// scran::fitTrendVar), damped Gauss-Newton from theta = (A, B, N); same iteration as
// trendvar._fit_parametric.  theta is updated in place.  Returns 0, or 2 when the start point is not
// finite (the caller keeps the start values, as scran does when nls() fails).
static bool solve3(const double a[3][3], const double g[3], double x[3])
{
    double m[3][4];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            m[i][j] = a[i][j];
        m[i][3] = g[i];
    }
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        for (int r = c + 1; r < 3; ++r)
            if (std::fabs(m[r][c]) > std::fabs(m[piv][c]))
                piv = r;
        if (!(std::fabs(m[piv][c]) > 0.0) || !std::isfinite(m[piv][c]))
            return false;
        for (int j = 0; j < 4; ++j)
            std::swap(m[c][j], m[piv][j]);
        for (int r = c + 1; r < 3; ++r) {
            const double f = m[r][c] / m[c][c];
            for (int j = c; j < 4; ++j)
                m[r][j] -= f * m[c][j];
        }
    }
    for (int c = 2; c >= 0; --c) {
        double t = m[c][3];
        for (int j = c + 1; j < 3; ++j)
            t -= m[c][j] * x[j];
        x[c] = t / m[c][c];
    }
    return std::isfinite(x[0]) && std::isfinite(x[1]) && std::isfinite(x[2]);
}

// Grid search of scran's .get_nls_starts (fitTrendVar): rss[n][b] = sum_i (v_i - grad b m_i / (m_i^n + b))^2 over
// the 10 x 10 grid of (n, b) values; the first minimum in expand.grid order (b fastest) is returned.
extern "C" int cb200_host_nls_grid(int64_t n64, const double *v, const double *m, double grad, int nb, const double *bgrid,
                                   int nn, const double *ngrid, int *ib_out, int *in_out)
{
    const int n = (int)n64;
    if (n < 1 || v == nullptr || m == nullptr || bgrid == nullptr || ngrid == nullptr || nb < 1 || nn < 1 || nb > 64)
        return 1;
    // one thread per exponent of the grid; the winner is picked afterwards in grid order (same tie rule)
    std::vector<double> rss_all((size_t)nn * 64, 0.0);
    host_parallel_for(nn, [&](int a) {
        double rss[64];
        for (int b = 0; b < nb; ++b)
            rss[b] = 0.0;
        for (int i = 0; i < n; ++i) {
            const double gm = m[i], vi = v[i], p = std::pow(m[i], ngrid[a]);
            for (int b = 0; b < nb; ++b) {
                const double r = vi - grad * bgrid[b] * gm / (p + bgrid[b]);
                rss[b] += r * r;
            }
        }
        for (int b = 0; b < nb; ++b)
            rss_all[(size_t)a * 64 + b] = rss[b];
    });
    double best = 0.0;
    int bi = -1, bn = -1;
    for (int a = 0; a < nn; ++a) {
        for (int b = 0; b < nb; ++b) {
            // NaN never wins; ties keep the earlier grid point (argmin over the flattened [n][b] array)
            const double r = rss_all[(size_t)a * 64 + b];
            if (r == r && (bi < 0 || r < best)) {
                best = r;
                bi = b;
                bn = a;
            }
        }
    }
    if (bi < 0)
        return 2;
    *ib_out = bi;
    *in_out = bn;
    return 0;
}

// A few worker threads for the passes over the G genes (each pass is ~G exp() calls; the fit makes dozens).
// Static partition, partial sums combined in chunk order: the result does not depend on timing.
namespace {
struct NlsPool {
    int T;
    std::vector<std::thread> th;
    std::atomic<int> gen{0}, done{0};
    std::atomic<bool> stop{false};
    std::function<void(int)> job;
    explicit NlsPool(int t) : T(t)
    {
        for (int k = 1; k < T; ++k)
            th.emplace_back([this, k] {
                int seen = 0;
                for (;;) {
                    while (gen.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_acquire))
                        std::this_thread::yield();
                    if (stop.load(std::memory_order_acquire))
                        return;
                    seen = gen.load(std::memory_order_acquire);
                    job(k);
                    done.fetch_add(1, std::memory_order_release);
                }
            });
    }
    void run(const std::function<void(int)> &f)
    {
        job = f;
        done.store(0, std::memory_order_relaxed);
        gen.fetch_add(1, std::memory_order_release);
        f(0);
        while (done.load(std::memory_order_acquire) != T - 1)
            std::this_thread::yield();
    }
    ~NlsPool()
    {
        stop.store(true, std::memory_order_release);
        for (auto &t : th)
            t.join();
    }
};
}  // namespace

extern "C" int cb200_host_nls_trend(int64_t n64, const double *v, const double *m, const double *w, double *theta,
                                    double tol, int maxiter)
{
    const int n = (int)n64;
    if (n < 4 || v == nullptr || m == nullptr || w == nullptr || theta == nullptr)
        return 1;
    int T = 1;   // measured on the 16-core GPU host: 4 threads 22 ms vs 1 thread 17 ms for the whole trend fit
    if (const char *e = getenv("CB200_HOST_THREADS"))
        T = atoi(e) > 0 ? atoi(e) : T;
    if (T > 16)
        T = 16;
    NlsPool pool(T);
    std::vector<double> logm(n), f(n), f2(n), j0(n), j1(n), j2(n), k0(n), k1(n), k2(n);
    struct Part {
        double rss, a[3][3], g[3];
        bool ok;
        char pad[64];
    };
    std::vector<Part> part((size_t)T);
    auto range = [&](int k, int &lo, int &hi) {
        const int per = (n + T - 1) / T;
        lo = k * per;
        hi = lo + per < n ? lo + per : n;
    };
    pool.run([&](int k) {
        int lo, hi;
        range(k, lo, hi);
        for (int i = lo; i < hi; ++i)
            logm[i] = std::log(m[i]);
    });
    // f, Jacobian, rss and the normal equations J^T W J, J^T W r of the point t, in one pass
    auto evaluate = [&](const double *t, std::vector<double> &ff, std::vector<double> &a0, std::vector<double> &a1,
                        std::vector<double> &a2, double &rss, double (&A)[3][3], double (&G)[3]) {
        const double ea = std::exp(t[0]), eb = std::exp(t[1]), en = std::exp(t[2]);
        pool.run([&](int k) {
            int lo, hi;
            range(k, lo, hi);
            Part &P = part[(size_t)k];
            P.rss = 0.0;
            P.ok = true;
            for (int p = 0; p < 3; ++p) {
                P.g[p] = 0.0;
                for (int q = 0; q < 3; ++q)
                    P.a[p][q] = 0.0;
            }
            for (int i = lo; i < hi; ++i) {
                const double p = std::exp((1.0 + en) * logm[i]);
                const double den = p + eb;
                const double fi = ea * m[i] / den;
                ff[i] = fi;
                a0[i] = fi;
                a1[i] = -fi * eb / den;
                a2[i] = -fi * p * logm[i] * en / den;
                const double r = v[i] - fi;
                P.rss += w[i] * r * r;
                P.ok = P.ok && std::isfinite(fi) && std::isfinite(a1[i]) && std::isfinite(a2[i]);
                const double jr[3] = {a0[i], a1[i], a2[i]};
                for (int pp = 0; pp < 3; ++pp) {
                    P.g[pp] += w[i] * jr[pp] * r;
                    for (int q = 0; q < 3; ++q)
                        P.a[pp][q] += w[i] * jr[pp] * jr[q];
                }
            }
        });
        rss = 0.0;
        bool ok = true;
        for (int p = 0; p < 3; ++p) {
            G[p] = 0.0;
            for (int q = 0; q < 3; ++q)
                A[p][q] = 0.0;
        }
        for (int k = 0; k < T; ++k) {
            rss += part[(size_t)k].rss;
            ok = ok && part[(size_t)k].ok;
            for (int p = 0; p < 3; ++p) {
                G[p] += part[(size_t)k].g[p];
                for (int q = 0; q < 3; ++q)
                    A[p][q] += part[(size_t)k].a[p][q];
            }
        }
        return ok && std::isfinite(rss);
    };
    double rss, a[3][3], g[3];
    if (!evaluate(theta, f, j0, j1, j2, rss, a, g))
        return 2;
    double damp = 1e-3;
    for (int it = 0; it < maxiter; ++it) {
        const double rr = rss;
        // relative-offset criterion: |Q^T r|^2 = g^T (J^T W J)^-1 g
        double x[3];
        if (solve3(a, g, x)) {
            const double num = g[0] * x[0] + g[1] * x[1] + g[2] * x[2];
            if (std::sqrt(num / std::max(rr - num, 1e-300)) < tol)
                break;
        }
        bool improved = false;
        for (int tr = 0; tr < 40; ++tr) {
            double ad[3][3];
            for (int p = 0; p < 3; ++p)
                for (int q = 0; q < 3; ++q)
                    ad[p][q] = a[p][q] + (p == q ? damp * (a[p][p] + 1e-300) : 0.0);
            double step[3];
            if (!solve3(ad, g, step)) {
                damp *= 10.0;
                continue;
            }
            const double t2[3] = {theta[0] + step[0], theta[1] + step[1], theta[2] + step[2]};
            double rss2, a2m[3][3], g2[3];
            if (evaluate(t2, f2, k0, k1, k2, rss2, a2m, g2) && rss2 < rss) {
                theta[0] = t2[0];
                theta[1] = t2[1];
                theta[2] = t2[2];
                f.swap(f2);
                j0.swap(k0);
                j1.swap(k1);
                j2.swap(k2);
                rss = rss2;
                for (int p = 0; p < 3; ++p) {
                    g[p] = g2[p];
                    for (int q = 0; q < 3; ++q)
                        a[p][q] = a2m[p][q];
                }
                damp = std::max(damp / 10.0, 1e-12);
                improved = true;
                break;
            }
            damp *= 10.0;
        }
        if (!improved)
            break;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Host compilation of the synthetic-data generator (cb200_synth_core.h): bit-identical to the device
// generator of cb200_synth.cu.  Test / bench infrastructure; no device work.
// ---------------------------------------------------------------------------------------------------------
#include <cstdlib>
#include <cstring>

#include "cb200_synth_core.h"
#include "cb200_synth_host.h"

void cb200_synth_params_host(int64_t G, uint64_t seed, std::vector<float> &base, std::vector<uint8_t> &wf,
                             std::vector<float> &wv, std::vector<float> &cent)
{
    base.resize((size_t)G);
    wf.resize((size_t)G * SY_NZ);
    wv.resize((size_t)G * SY_NZ);
    cent.resize((size_t)SY_CLUSTERS * SY_F);
    for (int64_t g = 0; g < G; ++g) {
        base[(size_t)g] = sy_gene_base(seed, g);
        for (int z = 0; z < SY_NZ; ++z) {
            wf[(size_t)g * SY_NZ + z] = (uint8_t)sy_gene_factor(seed, g, z);
            wv[(size_t)g * SY_NZ + z] = sy_gene_loading(seed, g, z);
        }
    }
    for (int i = 0; i < SY_CLUSTERS * SY_F; ++i)
        cent[(size_t)i] = sy_centroid(seed, i);
}

// offset that gives the neutral cell (all factors 0, depth 1) ~nnz_per_cell non-zero genes: bisection
float cb200_synth_neutral_offset(const std::vector<float> &base, double nnz_per_cell)
{
    float lo = -30.0f, hi = 10.0f;
    for (int it = 0; it < 40; ++it) {
        const float mid = sy_fmul(0.5f, sy_fadd(lo, hi));
        double nz = 0.0;
        for (size_t g = 0; g < base.size(); ++g) {
            const float ll = sy_fadd(base[g], mid);
            nz = sy_dadd(nz, (double)sy_fadd(1.0f, -sy_p0(sy_expf(ll < SY_LL_MAX ? ll : SY_LL_MAX))));
        }
        if (nz < nnz_per_cell)
            lo = mid;
        else
            hi = mid;
    }
    return sy_fmul(0.5f, sy_fadd(lo, hi));
}

// one step of the realised-density correction (the neutral cell ignores factor and depth variance)
float cb200_synth_refine_offset(float off, double nnz_per_cell, int64_t nnz_probe, int64_t n_probe)
{
    const double actual = sy_ddiv((double)nnz_probe, (double)n_probe);
    return sy_fadd(off, (float)sy_dmul(1.3, sy_log(sy_ddiv(nnz_per_cell, actual))));
}

bool cb200_synth_offset_ok(double nnz_per_cell, int64_t nnz_probe, int64_t n_probe)
{
    const double ratio = sy_ddiv(sy_ddiv((double)nnz_probe, (double)n_probe), nnz_per_cell);
    return ratio > 0.99 && ratio < 1.01;
}

namespace {
struct SyBlock {
    std::vector<int32_t> cnt, row;
    std::vector<double> val;
};

// cells [c0, c1) of the global matrix; FILL = false only counts
void sy_host_cells(int64_t G, int64_t c0, int64_t c1, uint64_t seed, float off, const std::vector<float> &base,
                   const std::vector<uint8_t> &wf, const std::vector<float> &wv, const std::vector<float> &cent,
                   bool fill, SyBlock &out)
{
    float zs[SY_F];
    out.cnt.assign((size_t)(c1 - c0), 0);
    for (int64_t c = c0; c < c1; ++c) {
        const uint64_t cg = (uint64_t)c;
        const int cl = sy_cell_cluster(seed, cg);
        for (int f = 0; f < SY_F; ++f)
            zs[f] = sy_cell_factor(seed, cg, f, cent.data(), cl);
        const float ldepth = sy_cell_logdepth(seed, cg);
        int total = 0;
        for (int64_t g = 0; g < G; ++g) {
            const int k = sy_count(seed, cg, g, base[(size_t)g], &wf[(size_t)g * SY_NZ], &wv[(size_t)g * SY_NZ], zs, off,
                                   ldepth);
            if (k > 0) {
                ++total;
                if (fill) {
                    out.row.push_back((int32_t)g);
                    out.val.push_back((double)k);
                }
            }
        }
        if (total == 0) {
            total = 1;
            if (fill) {
                out.row.push_back((int32_t)(cg % (uint64_t)G));
                out.val.push_back(1.0);
            }
        }
        out.cnt[(size_t)(c - c0)] = total;
    }
}

template <typename Fn>
void sy_parallel_blocks(int64_t n_blocks, int n_threads, Fn fn)
{
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= n_blocks)
                return;
            fn(b);
        }
    };
    if (n_threads <= 1) {
        work();
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back(work);
    for (auto &t : th)
        t.join();
}
}  // namespace

int cb200_synth_calibrate_host(int64_t n_genes, int64_t n_cells_total, double nnz_per_cell, uint64_t seed, int n_threads,
                               const std::vector<float> &base, const std::vector<uint8_t> &wf,
                               const std::vector<float> &wv, const std::vector<float> &cent, float *off_out)
{
    float off = cb200_synth_neutral_offset(base, nnz_per_cell);
    const int64_t n_probe = n_cells_total < CB200_SYNTH_PROBE ? n_cells_total : CB200_SYNTH_PROBE;
    const int64_t BL = 16, nb = (n_probe + BL - 1) / BL;
    for (int pass = 0; pass < 6; ++pass) {
        std::vector<int64_t> part((size_t)nb, 0);
        sy_parallel_blocks(nb, n_threads, [&](int64_t b) {
            SyBlock blk;
            const int64_t c0 = b * BL, c1 = std::min(n_probe, c0 + BL);
            sy_host_cells(n_genes, c0, c1, seed, off, base, wf, wv, cent, false, blk);
            int64_t s = 0;
            for (int32_t v : blk.cnt)
                s += v;
            part[(size_t)b] = s;
        });
        int64_t nnz = 0;
        for (int64_t v : part)
            nnz += v;
        if (cb200_synth_offset_ok(nnz_per_cell, nnz, n_probe))
            break;
        off = cb200_synth_refine_offset(off, nnz_per_cell, nnz, n_probe);
    }
    *off_out = off;
    return 0;
}

extern "C" int cb200_synth_host(int64_t n_genes, int64_t n_cells, double nnz_per_cell, uint64_t seed, int64_t cell_offset,
                                int64_t n_cells_total, int n_threads, int64_t *nnz_out, int64_t **colptr_out,
                                int32_t **rowidx_out, double **values_out)
{
    if (n_genes < 64 || n_genes >= 2147483647LL || n_cells <= 0 || cell_offset < 0 ||
        cell_offset + n_cells > n_cells_total || !(nnz_per_cell >= 1.0) || !(nnz_per_cell < (double)n_genes) ||
        nnz_out == nullptr || colptr_out == nullptr || rowidx_out == nullptr || values_out == nullptr)
        return 1;
    if (n_threads <= 0)
        n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<float> base, wv, cent;
    std::vector<uint8_t> wf;
    cb200_synth_params_host(n_genes, seed, base, wf, wv, cent);
    float off = 0.f;
    cb200_synth_calibrate_host(n_genes, n_cells_total, nnz_per_cell, seed, n_threads, base, wf, wv, cent, &off);
    const int64_t BL = 32, nb = (n_cells + BL - 1) / BL;
    std::vector<SyBlock> blocks((size_t)nb);
    sy_parallel_blocks(nb, n_threads, [&](int64_t b) {
        const int64_t c0 = cell_offset + b * BL, c1 = std::min(cell_offset + n_cells, c0 + BL);
        sy_host_cells(n_genes, c0, c1, seed, off, base, wf, wv, cent, true, blocks[(size_t)b]);
    });
    int64_t nnz = 0;
    for (auto &b : blocks)
        nnz += (int64_t)b.row.size();
    int64_t *cp = (int64_t *)malloc((size_t)(n_cells + 1) * sizeof(int64_t));
    int32_t *ri = (int32_t *)malloc((size_t)std::max<int64_t>(nnz, 1) * sizeof(int32_t));
    double *xv = (double *)malloc((size_t)std::max<int64_t>(nnz, 1) * sizeof(double));
    if (cp == nullptr || ri == nullptr || xv == nullptr) {
        free(cp);
        free(ri);
        free(xv);
        return 2;
    }
    int64_t pos = 0, c = 0;
    cp[0] = 0;
    for (auto &b : blocks) {
        memcpy(ri + pos, b.row.data(), b.row.size() * sizeof(int32_t));
        memcpy(xv + pos, b.val.data(), b.val.size() * sizeof(double));
        int64_t p = pos;
        for (int32_t v : b.cnt) {
            p += v;
            cp[++c] = p;
        }
        pos += (int64_t)b.row.size();
    }
    *nnz_out = nnz;
    *colptr_out = cp;
    *rowidx_out = ri;
    *values_out = xv;
    return 0;
}

extern "C" void cb200_synth_host_free(int64_t *colptr, int32_t *rowidx, double *values)
{
    free(colptr);
    free(rowidx);
    free(values);
}

extern "C" int cb200_synth_embedding_host(int64_t n, int d, uint64_t seed, int n_threads, double *emb_colmajor)
{
    if (n <= 0 || d <= 0 || emb_colmajor == nullptr)
        return 1;
    if (n_threads <= 0)
        n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int64_t BL = 4096, nb = (n + BL - 1) / BL;
    sy_parallel_blocks(nb * d, n_threads, [&](int64_t job) {
        const int t = (int)(job / nb);
        const int64_t i0 = (job % nb) * BL, i1 = std::min(n, i0 + BL);
        const double sc = sy_emb_scale(t, d);
        for (int64_t i = i0; i < i1; ++i)
            emb_colmajor[(size_t)t * (size_t)n + (size_t)i] = sy_emb_value(seed, i, t, sc, SY_EMB_CLUSTERS);
    });
    return 0;
}
