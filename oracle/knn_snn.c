/*
 * oracle/knn_snn.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never on the product path).
 *
 * Plain-C restatement of the two upstream routines that cellula's
 * makeGraphsAndClusters() reaches through bluster::makeSNNGraph()
 * (/root/reference/R/makeGraphsAndClusters.R:83-85, :125-127, :284):
 *
 *   1. BiocNeighbors::findKNN(x, k, BNPARAM = KmknnParam(), get.distance = FALSE)
 *      -> exact Euclidean kNN, self excluded, row i ordered by (distance, index).
 *      The algorithm lives in the un-vendored dependency BiocNeighbors/knncolle
 *      (no version pin in /root/reference/DESCRIPTION:14-50; era Bioconductor 3.20,
 *      BiocNeighbors 2.0).  Its published contract: distance = sqrt(sum_t (a_t-b_t)^2),
 *      fp64, accumulated sequentially, candidates kept in a (distance, index) queue.
 *      Kmknn's k-means pruning does not change the result, so a brute-force scan with
 *      the same arithmetic is the restatement.
 *
 *   2. bluster::neighborsToSNNGraph(index, type) -> libscran build_snn_graph:
 *      shared-nearest-neighbour edges with rank / number / jaccard weights
 *      (SURVEY.md 8c.1-7, row a10).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
 * path and R/Bioconductor cannot run in the build container, so
 * this oracle is pinned only by the hand-derived known-answer vectors of SURVEY.md 8c.3
 * (tests/test_oracle_kat.py).
 *
 * Build: see oracle/Makefile ( -O2 -ffp-contract=off -fopenmp ).  -ffp-contract=off
 * reproduces R's default x86-64 build, which has no FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ kNN */

/* emb: column-major N x d (R matrix layout).  index_out: column-major N x k, 1-based.
 * dist_out (nullable): column-major N x k distances.  Returns 0, or -1 on bad args. */
int oracle_knn(const double *emb, int64_t N, int d, int k, int32_t *index_out,
               double *dist_out)
{
    if (N < 2 || d < 1 || k < 1 || k > N - 1)
        return -1;
    double *rows = (double *)malloc((size_t)N * d * sizeof(double));
    if (!rows)
        return -2;
    for (int64_t i = 0; i < N; ++i)
        for (int t = 0; t < d; ++t)
            rows[i * d + t] = emb[i + N * (int64_t)t];

#pragma omp parallel
    {
        double *bd = (double *)malloc((size_t)k * sizeof(double));  /* sqrt'd distances */
        double *br = (double *)malloc((size_t)k * sizeof(double));  /* raw sums, same order */
        int32_t *bi = (int32_t *)malloc((size_t)k * sizeof(int32_t));
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < N; ++i) {
            const double *a = rows + i * d;
            int cnt = 0;
            for (int64_t j = 0; j < N; ++j) {
                if (j == i)
                    continue;
                const double *b = rows + j * d;
                double acc = 0.0;
                for (int t = 0; t < d; ++t) {
                    double delta = a[t] - b[t];
                    acc += delta * delta;
                }
                /* j ascends, so on an exact distance tie the incumbent (smaller index)
                 * wins: insertion needs dist < worst strictly.  raw > worst_raw implies
                 * sqrt(raw) >= worst, so the sqrt can be skipped then. */
                if (cnt == k && acc > br[k - 1])
                    continue;
                double dist = sqrt(acc);
                if (cnt == k && !(dist < bd[k - 1]))
                    continue;
                int pos = (cnt < k) ? cnt : k - 1;
                while (pos > 0 && dist < bd[pos - 1]) { /* strict: ties keep smaller idx first */
                    bd[pos] = bd[pos - 1];
                    br[pos] = br[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = dist;
                br[pos] = acc;
                bi[pos] = (int32_t)j;
                if (cnt < k)
                    ++cnt;
            }
            for (int c = 0; c < k; ++c) {
                index_out[i + N * (int64_t)c] = bi[c] + 1;
                if (dist_out)
                    dist_out[i + N * (int64_t)c] = bd[c];
            }
        }
        free(bd);
        free(br);
        free(bi);
    }
    free(rows);
    return 0;
}

/* ------------------------------------------------------------------ SNN */

/* index: column-major N x k, 1-based (findKNN()$index).  type: 0 rank, 1 number, 2 jaccard.
 * Edges are emitted as (from = j+1, to = h+1 with h < j, weight) in order of first
 * encounter, exactly as SURVEY.md 8c.1-7 states.  Arrays are malloc'ed here; release
 * with oracle_free(). Returns 0 / negative on error. */
int oracle_snn_range(const int32_t *index, int64_t N, int k, int type, int64_t j_begin, int64_t j_end,
                     int64_t *n_edges, int32_t **from_out, int32_t **to_out, double **w_out)
{
    if (N < 1 || k < 1 || type < 0 || type > 2 || j_begin < 0 || j_end > N || j_begin > j_end)
        return -1;
    const int L = k + 1;
    /* hosts[s] = (cell, rank) of every cell whose extended list holds s; CSR layout. */
    int64_t *hptr = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t));
    if (!hptr)
        return -2;
    for (int64_t j = 0; j < N; ++j) {
        hptr[j + 1] += 1; /* (j, rank 0) */
        for (int i = 0; i < k; ++i) {
            int64_t s = (int64_t)index[j + N * (int64_t)i] - 1;
            if (s < 0 || s >= N) {
                free(hptr);
                return -3;
            }
            hptr[s + 1] += 1;
        }
    }
    for (int64_t s = 0; s < N; ++s)
        hptr[s + 1] += hptr[s];
    int64_t total = hptr[N];
    int32_t *hcell = (int32_t *)malloc((size_t)total * sizeof(int32_t));
    int32_t *hrank = (int32_t *)malloc((size_t)total * sizeof(int32_t));
    int64_t *fill = (int64_t *)malloc((size_t)N * sizeof(int64_t));
    memcpy(fill, hptr, (size_t)N * sizeof(int64_t));
    for (int64_t j = 0; j < N; ++j) { /* ascending j => every hosts list is sorted by cell */
        for (int i = 0; i < L; ++i) {
            int64_t s = (i == 0) ? j : (int64_t)index[j + N * (int64_t)(i - 1)] - 1;
            hcell[fill[s]] = (int32_t)j;
            hrank[fill[s]] = i;
            fill[s]++;
        }
    }
    free(fill);

    int64_t cap = (int64_t)N * L + 16, ne = 0;
    int32_t *from = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
    int32_t *to = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
    double *w = (double *)malloc((size_t)cap * sizeof(double));
    int64_t *slot = (int64_t *)malloc((size_t)N * sizeof(int64_t)); /* edge slot of h for current j */
    int64_t *stamp = (int64_t *)malloc((size_t)N * sizeof(int64_t));
    int32_t *score = (int32_t *)malloc((size_t)cap * sizeof(int32_t)); /* best rank sum or count */
    for (int64_t h = 0; h < N; ++h)
        stamp[h] = -1;

    for (int64_t j = j_begin; j < j_end; ++j) { /* the edges of cell j depend on the lists only: any j-range */
        for (int i = 0; i < L; ++i) {
            int64_t s = (i == 0) ? j : (int64_t)index[j + N * (int64_t)(i - 1)] - 1;
            for (int64_t e = hptr[s]; e < hptr[s + 1]; ++e) {
                int64_t h = hcell[e];
                if (h >= j)
                    break; /* list is sorted by cell */
                int32_t r = i + hrank[e];
                if (stamp[h] != j) {
                    stamp[h] = j;
                    if (ne == cap) {
                        cap = cap + cap / 2;
                        from = (int32_t *)realloc(from, (size_t)cap * sizeof(int32_t));
                        to = (int32_t *)realloc(to, (size_t)cap * sizeof(int32_t));
                        w = (double *)realloc(w, (size_t)cap * sizeof(double));
                        score = (int32_t *)realloc(score, (size_t)cap * sizeof(int32_t));
                    }
                    slot[h] = ne;
                    from[ne] = (int32_t)(j + 1);
                    to[ne] = (int32_t)(h + 1);
                    score[ne] = (type == 0) ? r : 1;
                    ++ne;
                } else {
                    int64_t q = slot[h];
                    if (type == 0) {
                        if (r < score[q])
                            score[q] = r;
                    } else {
                        score[q] += 1;
                    }
                }
            }
        }
    }
    for (int64_t e = 0; e < ne; ++e) {
        double v;
        if (type == 0) {
            v = (double)k - 0.5 * (double)score[e];
        } else if (type == 1) {
            v = (double)score[e];
        } else {
            v = (double)score[e] / (double)(2 * L - score[e]);
        }
        w[e] = v > 1e-6 ? v : 1e-6;
    }
    free(score);
    free(slot);
    free(stamp);
    free(hptr);
    free(hcell);
    free(hrank);
    *n_edges = ne;
    *from_out = from;
    *to_out = to;
    *w_out = w;
    return 0;
}

int oracle_snn(const int32_t *index, int64_t N, int k, int type, int64_t *n_edges,
               int32_t **from_out, int32_t **to_out, double **w_out)
{
    return oracle_snn_range(index, N, k, type, 0, N, n_edges, from_out, to_out, w_out);
}

void oracle_free(void *p) { free(p); }

/* Exact kNN of selected query rows only (same arithmetic and tie rule as oracle_knn): used to check
 * sampled rows at sizes where the full N x N scan is too slow for a test.
 * queries: 0-based rows; index_out: row-major nq x k, 1-based; dist_out nullable, row-major nq x k. */
int oracle_knn_queries(const double *emb, int64_t N, int d, int k, const int64_t *queries, int64_t nq,
                       int32_t *index_out, double *dist_out)
{
    if (N < 2 || d < 1 || k < 1 || k > N - 1 || nq < 0)
        return -1;
    double *rows = (double *)malloc((size_t)N * d * sizeof(double));
    if (!rows)
        return -2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        for (int t = 0; t < d; ++t)
            rows[i * d + t] = emb[i + N * (int64_t)t];
    int bad = 0;
#pragma omp parallel
    {
        double *bd = (double *)malloc((size_t)k * sizeof(double));
        double *br = (double *)malloc((size_t)k * sizeof(double));
        int32_t *bi = (int32_t *)malloc((size_t)k * sizeof(int32_t));
#pragma omp for schedule(dynamic, 4)
        for (int64_t qi = 0; qi < nq; ++qi) {
            const int64_t i = queries[qi];
            if (i < 0 || i >= N) {
                bad = 1;
                continue;
            }
            const double *a = rows + i * d;
            int cnt = 0;
            for (int64_t j = 0; j < N; ++j) {
                if (j == i)
                    continue;
                const double *b = rows + j * d;
                double acc = 0.0;
                for (int t = 0; t < d; ++t) {
                    double delta = a[t] - b[t];
                    acc += delta * delta;
                }
                if (cnt == k && acc > br[k - 1])
                    continue;
                double dist = sqrt(acc);
                if (cnt == k && !(dist < bd[k - 1]))
                    continue;
                int pos = (cnt < k) ? cnt : k - 1;
                while (pos > 0 && dist < bd[pos - 1]) {
                    bd[pos] = bd[pos - 1];
                    br[pos] = br[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = dist;
                br[pos] = acc;
                bi[pos] = (int32_t)j;
                if (cnt < k)
                    ++cnt;
            }
            for (int c = 0; c < k; ++c) {
                index_out[qi * k + c] = bi[c] + 1;
                if (dist_out)
                    dist_out[qi * k + c] = bd[c];
            }
        }
        free(bd);
        free(br);
        free(bi);
    }
    free(rows);
    return bad ? -3 : 0;
}

/* --------------------------------------------------- normalisation + gene statistics
 * scuttle::librarySizeFactors / normalizeCounts and scran's per-gene mean/variance
 * (SURVEY.md 8c.1-1..3), fp64, single pass per column, Welford per gene in cell order.
 * Used as the fast CPU baseline at sizes numpy is too slow for. */
int oracle_lognorm_genestats(int64_t G, int64_t N, const int64_t *colptr,
                             const int32_t *rowidx, const double *x, double *sf_out,
                             double *logx_out, double *mean_out, double *var_out)
{
    double total = 0.0;
    for (int64_t c = 0; c < N; ++c) {
        double s = 0.0;
        for (int64_t e = colptr[c]; e < colptr[c + 1]; ++e)
            s += x[e];
        sf_out[c] = s;
        total += s;
    }
    double mean_ls = total / (double)N;
    for (int64_t c = 0; c < N; ++c) {
        sf_out[c] /= mean_ls;
        if (!(sf_out[c] > 0.0))
            return -4; /* scuttle: "size factors should be positive" */
    }
    /* Welford over the non-zeros, then fold the implicit zeros in closed form. */
    double *m = (double *)calloc((size_t)G, sizeof(double));
    double *m2 = (double *)calloc((size_t)G, sizeof(double));
    int64_t *cnt = (int64_t *)calloc((size_t)G, sizeof(int64_t));
    const double ln2 = 0.693147180559945309417232121458;
    for (int64_t c = 0; c < N; ++c) {
        for (int64_t e = colptr[c]; e < colptr[c + 1]; ++e) {
            double y = log1p(x[e] / sf_out[c]) / ln2;
            if (logx_out)
                logx_out[e] = y;
            int32_t g = rowidx[e];
            cnt[g] += 1;
            double dlt = y - m[g];
            m[g] += dlt / (double)cnt[g];
            m2[g] += dlt * (y - m[g]);
        }
    }
    for (int64_t g = 0; g < G; ++g) {
        double nz = (double)cnt[g], n = (double)N;
        double mean_all = m[g] * nz / n;
        /* merge group (nz, m, m2) with group (n-nz, 0, 0): Chan et al. */
        double M2 = m2[g] + m[g] * m[g] * nz * (n - nz) / n;
        mean_out[g] = mean_all;
        var_out[g] = (N > 1) ? M2 / (n - 1.0) : NAN;
    }
    free(m);
    free(m2);
    free(cnt);
    return 0;
}
