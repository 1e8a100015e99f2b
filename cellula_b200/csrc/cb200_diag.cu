// cb200_diag.cu -- SURVEY.md 8(f) rank 2: the per-resolution cluster diagnostics of
// makeGraphsAndClusters(), /root/reference/R/makeGraphsAndClusters.R:103-114 (:144-153, :288):
//
//   bluster::approxSilhouette(space, clusters)                  -> cb200_approx_silhouette
//   bluster::pairwiseModularity(g, clusters, as.ratio = TRUE)   -> cb200_pairwise_modularity
//
// Both are reductions over data the path already holds in HBM (the embedding, the SNN edge list).
// bluster is not vendored in the reference; the definitions restated here (and in oracle/oracle.py)
// are those of bluster 1.16 (Bioconductor 3.20):
//
// approxSilhouette.  For every cluster u (sorted unique labels): centroid c_u = colMeans, spread
//   v_u = sum over dimensions of the mean squared deviation from c_u.  For every cell i and cluster u
//   D_iu = sqrt(|x_i - c_u|^2 + v_u) (root-mean-square distance from i to the cells of u);
//   self = D_i,own; other = the smallest D_iu over u != own (first such u on ties);
//   width = (other - self) / max(other, self).
//
// pairwiseModularity.  W = sum of the edge weights; observed[x][y] (x <= y, upper triangle) = weight of
//   the edges between clusters x and y (x = y: within); p_x = (sum of the weighted degrees of x) / 2W;
//   expected[x][x] = W p_x^2, expected[x][y] = 2 W p_x p_y (x < y); lower triangles are 0.
//   as.ratio = TRUE returns observed / expected; otherwise (observed - expected) / W.
#include <cfloat>

#include "cb200_common.cuh"

namespace {

// sums[u][t] += x_i[t], cnt[u] += 1
__global__ void k_sil_sums(const double *__restrict__ E, int64_t lde, int64_t n, int d, const int32_t *__restrict__ lab,
                           double *__restrict__ sums, unsigned long long *__restrict__ cnt)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < n * d) {
        const int64_t i = idx / d;
        const int t = (int)(idx % d);
        const int u = lab[i];
        atomicAdd(sums + (size_t)u * d + t, E[i * lde + t]);
        if (t == 0)
            atomicAdd(cnt + u, 1ULL);
    }
}

__global__ void k_sil_centroids(int K, int d, const unsigned long long *__restrict__ cnt, double *__restrict__ sums)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < K * d) {
        const unsigned long long c = cnt[idx / d];
        sums[idx] = c > 0 ? sums[idx] / (double)c : 0.0;
    }
}

// dev[u] += |x_i - c_own|^2 (warp per cell)
__global__ void k_sil_spread(const double *__restrict__ E, int64_t lde, int64_t n, int d, const int32_t *__restrict__ lab,
                             const double *__restrict__ cen, double *__restrict__ dev)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= n)
        return;
    const int u = lab[i];
    double s = 0.0;
    for (int t = lane; t < d; t += 32) {
        const double df = E[i * lde + t] - cen[(size_t)u * d + t];
        s += df * df;
    }
    s = warp_sum(s);
    if (lane == 0)
        atomicAdd(dev + u, s);
}

// width / other of every cell (thread per cell, clusters in label order)
__global__ void k_sil_width(const double *__restrict__ E, int64_t lde, int64_t n, int d, const int32_t *__restrict__ lab,
                            int K, const double *__restrict__ cen, const double *__restrict__ dev,
                            const unsigned long long *__restrict__ cnt, double *__restrict__ width,
                            int32_t *__restrict__ other)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int own = lab[i];
    double self = 1.0 / 0.0, best = 1.0 / 0.0;
    int bu = -1;
    for (int u = 0; u < K; ++u) {
        if (cnt[u] == 0)
            continue;   // label value without cells: not a cluster
        double s = 0.0;
        for (int t = 0; t < d; ++t) {
            const double df = E[i * lde + t] - cen[(size_t)u * d + t];
            s += df * df;
        }
        const double D = sqrt(s + dev[u] / (double)cnt[u]);
        if (u == own) {
            self = D;
        } else if (D < best) {
            best = D;
            bu = u;
        }
    }
    width[i] = (best - self) / fmax(best, self);
    other[i] = bu;
}

// observed[min(a,b)][max(a,b)] += w, deg[a] += w, deg[b] += w, total += w; block-private tables in
// shared memory (K <= 64), flushed once per block
#define MOD_MAXK_SMEM 64
__global__ void __launch_bounds__(256)
k_mod_accumulate(const int32_t *__restrict__ from, const int32_t *__restrict__ to, const double *__restrict__ w,
                 int64_t ne, const int32_t *__restrict__ lab, int K, double *__restrict__ obs, double *__restrict__ deg,
                 double *__restrict__ total)
{
    extern __shared__ double sm[];   // obs K*K | deg K | total 1  (only when K <= MOD_MAXK_SMEM)
    const bool priv = K <= MOD_MAXK_SMEM;
    double *sobs = sm, *sdeg = sm + K * K, *stot = sdeg + K;
    if (priv) {
        for (int i = threadIdx.x; i < K * K + K + 1; i += blockDim.x)
            sm[i] = 0.0;
        __syncthreads();
    }
    double tsum = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += stride) {
        const int a = lab[from[e] - 1], b = lab[to[e] - 1];
        const double we = w[e];
        const int x = a < b ? a : b, y = a < b ? b : a;
        if (priv) {
            atomicAdd(sobs + x * K + y, we);
            atomicAdd(sdeg + a, we);
            atomicAdd(sdeg + b, we);
        } else {
            atomicAdd(obs + (size_t)x * K + y, we);
            atomicAdd(deg + a, we);
            atomicAdd(deg + b, we);
        }
        tsum += we;
    }
    tsum = warp_sum(tsum);
    if ((threadIdx.x & 31) == 0) {
        if (priv)
            atomicAdd(stot, tsum);
        else
            atomicAdd(total, tsum);
    }
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < K * K; i += blockDim.x)
            if (sobs[i] != 0.0)
                atomicAdd(obs + i, sobs[i]);
        for (int i = threadIdx.x; i < K; i += blockDim.x)
            if (sdeg[i] != 0.0)
                atomicAdd(deg + i, sdeg[i]);
        if (threadIdx.x == 0)
            atomicAdd(total, *stot);
    }
}

// expected matrix from the degree sums
__global__ void k_mod_expected(int K, const double *__restrict__ deg, const double *__restrict__ total,
                               double *__restrict__ expected)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < K * K) {
        const int x = idx / K, y = idx % K;
        const double W = *total;
        const double px = deg[x] / (2.0 * W), py = deg[y] / (2.0 * W);   // sum(deg) = 2W
        double v = 0.0;
        if (x == y)
            v = px * py * W;
        else if (x < y)
            v = 2.0 * (px * py * W);
        expected[idx] = v;
    }
}

__global__ void k_colmajor_to_rowmajor_d(const double *__restrict__ in, int64_t rows, int cols, double *__restrict__ out)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < rows * cols) {
        const int64_t r = idx % rows;
        const int c = (int)(idx / rows);
        out[r * cols + c] = in[idx];
    }
}

}  // namespace

extern "C" int cb200_approx_silhouette(cb200_ctx *ctx, const double *emb, int64_t n, int d, const int32_t *labels,
                                       int n_clusters, double *width_out, int32_t *other_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, labels != nullptr && width_out != nullptr && other_out != nullptr, "cb200_approx_silhouette: NULL argument");
    CB_CHECK(ctx, n >= 1 && d >= 1 && n_clusters >= 1, "cb200_approx_silhouette: bad shape");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int64_t i = 0; i < n; ++i)
        CB_CHECK(ctx, labels[i] >= 0 && labels[i] < n_clusters, "cb200_approx_silhouette: label %d of cell %lld out of [0, %d)",
                 labels[i], (long long)i, n_clusters);
    const double *E = nullptr;
    int64_t lde = d;
    if (emb == nullptr) {
        CB_CHECK(ctx, ctx->have_scores && n == ctx->N && d <= ctx->d,
                 "cb200_approx_silhouette: no embedding given and the resident PCA scores do not match (%lld x %d)",
                 (long long)n, d);
        E = ctx->scores.p;
        lde = ctx->d;
    } else {
        CB_CUDA(ctx, ctx->stage_d.ensure((size_t)n * d));
        CB_CUDA(ctx, ctx->emb.ensure((size_t)n * d));
        CB_TRY(cb200_h2d(ctx, ctx->stage_d.p, emb, (size_t)n * d * sizeof(double), ctx->stream));
        k_colmajor_to_rowmajor_d<<<(unsigned)((n * d + 255) / 256), 256, 0, ctx->stream>>>(ctx->stage_d.p, n, d, ctx->emb.p);
        CB_LAUNCH(ctx);
        E = ctx->emb.p;
    }
    const int K = n_clusters;
    CB_CUDA(ctx, ctx->diag_lab.ensure((size_t)n));
    CB_CUDA(ctx, ctx->diag_d.ensure((size_t)K * d + K + (size_t)n));
    CB_CUDA(ctx, ctx->diag_cnt.ensure((size_t)K));
    CB_CUDA(ctx, ctx->diag_i.ensure((size_t)n));
    double *cen = ctx->diag_d.p, *dev = cen + (size_t)K * d, *width = dev + K;
    CB_CUDA(ctx, cudaMemcpyAsync(ctx->diag_lab.p, labels, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(cen, 0, ((size_t)K * d + K) * sizeof(double), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->diag_cnt.p, 0, (size_t)K * sizeof(unsigned long long), ctx->stream));
    k_sil_sums<<<(unsigned)((n * d + 255) / 256), 256, 0, ctx->stream>>>(E, lde, n, d, ctx->diag_lab.p, cen, ctx->diag_cnt.p);
    CB_LAUNCH(ctx);
    k_sil_centroids<<<(K * d + 255) / 256, 256, 0, ctx->stream>>>(K, d, ctx->diag_cnt.p, cen);
    CB_LAUNCH(ctx);
    k_sil_spread<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(E, lde, n, d, ctx->diag_lab.p, cen, dev);
    CB_LAUNCH(ctx);
    k_sil_width<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(E, lde, n, d, ctx->diag_lab.p, K, cen, dev,
                                                                       ctx->diag_cnt.p, width, ctx->diag_i.p);
    CB_LAUNCH(ctx);
    CB_CUDA(ctx, cudaMemcpyAsync(width_out, width, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(other_out, ctx->diag_i.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cb200_pairwise_modularity(cb200_ctx *ctx, const int32_t *from, const int32_t *to, const double *weight,
                                         int64_t n_edges, const int32_t *labels, int64_t n, int n_clusters,
                                         double *observed_out, double *expected_out, double *total_weight_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, labels != nullptr && observed_out != nullptr && expected_out != nullptr,
             "cb200_pairwise_modularity: NULL argument");
    CB_CHECK(ctx, n >= 1 && n_clusters >= 1 && n_clusters <= 4096, "cb200_pairwise_modularity: 1 <= clusters <= 4096 required");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int64_t i = 0; i < n; ++i)
        CB_CHECK(ctx, labels[i] >= 0 && labels[i] < n_clusters, "cb200_pairwise_modularity: label %d of cell %lld out of [0, %d)",
                 labels[i], (long long)i, n_clusters);
    const int K = n_clusters;
    const int32_t *dfrom = nullptr, *dto = nullptr;
    const double *dw = nullptr;
    int64_t ne = n_edges;
    if (from == nullptr) {   // the edges left in HBM by the last cb200_snn* call
        CB_CHECK(ctx, ctx->n_edges > 0 && ctx->efrom.p != nullptr, "cb200_pairwise_modularity: no edge list given and none resident");
        ne = ctx->n_edges;
        dfrom = ctx->efrom.p;
        dto = ctx->eto.p;
        dw = ctx->ew8.p;
    } else {
        CB_CHECK(ctx, to != nullptr && weight != nullptr && n_edges >= 0, "cb200_pairwise_modularity: NULL edge array");
        for (int64_t e = 0; e < n_edges; ++e)
            CB_CHECK(ctx, from[e] >= 1 && from[e] <= n && to[e] >= 1 && to[e] <= n,
                     "cb200_pairwise_modularity: edge %lld out of range [1, n]", (long long)e);
        CB_CUDA(ctx, ctx->diag_ef.ensure((size_t)(2 * n_edges + 1)));
        CB_CUDA(ctx, ctx->diag_ew.ensure((size_t)n_edges + 1));
        CB_TRY(cb200_h2d(ctx, ctx->diag_ef.p, from, (size_t)n_edges * sizeof(int32_t), ctx->stream));
        CB_TRY(cb200_h2d(ctx, ctx->diag_ef.p + n_edges, to, (size_t)n_edges * sizeof(int32_t), ctx->stream));
        CB_TRY(cb200_h2d(ctx, ctx->diag_ew.p, weight, (size_t)n_edges * sizeof(double), ctx->stream));
        dfrom = ctx->diag_ef.p;
        dto = ctx->diag_ef.p + n_edges;
        dw = ctx->diag_ew.p;
    }
    CB_CUDA(ctx, ctx->diag_lab.ensure((size_t)n));
    CB_CUDA(ctx, ctx->diag_d.ensure((size_t)2 * K * K + K + 1));
    double *obs = ctx->diag_d.p, *expd = obs + (size_t)K * K, *deg = expd + (size_t)K * K, *total = deg + K;
    CB_CUDA(ctx, cudaMemcpyAsync(ctx->diag_lab.p, labels, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(obs, 0, ((size_t)2 * K * K + K + 1) * sizeof(double), ctx->stream));
    if (ne > 0) {
        const size_t smem = K <= MOD_MAXK_SMEM ? ((size_t)K * K + K + 1) * sizeof(double) : 0;
        const int grid = cb_grid_for(ne, 256 * 8, ctx->num_sms * 8);
        k_mod_accumulate<<<grid, 256, smem, ctx->stream>>>(dfrom, dto, dw, ne, ctx->diag_lab.p, K, obs, deg, total);
        CB_LAUNCH(ctx);
    }
    k_mod_expected<<<(K * K + 255) / 256, 256, 0, ctx->stream>>>(K, deg, total, expd);
    CB_LAUNCH(ctx);
    CB_CUDA(ctx, cudaMemcpyAsync(observed_out, obs, (size_t)K * K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(expected_out, expd, (size_t)K * K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    double tw = 0.0;
    CB_CUDA(ctx, cudaMemcpyAsync(&tw, total, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (total_weight_out != nullptr)
        *total_weight_out = tw;
    return 0;
}
