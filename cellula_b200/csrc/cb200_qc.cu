// cb200_qc.cu -- SURVEY 8(f) rank 4: the two callers either side of the path that re-use its kernels.
//
//   cb200_qc_metrics      replaces scuttle::perCellQCMetrics(sce, subsets = ...) as doQC() calls it
//                         (/root/reference/R/doQC.R:106-108): per cell the total count, the number of detected
//                         genes (count > detection limit) and the same two numbers over every gene subset
//                         (mito / Malat1 / Ribo); the *_percent columns are 100 * subset sum / total.
//                         One streaming pass over the CSC matrix (K1 plus a gene -> subset-mask lookup that
//                         stays in L1/L2): 12 B per non-zero.
//   cb200_knn_distances   the distances findKNN(get.distance = TRUE) returns next to the index: what
//                         uwot::umap(nn_method = list(idx =, dist =)) needs to re-use the exact kNN result
//                         instead of running its own Annoy search (/root/reference/R/doNormAndReduce.R:102-106,
//                         n_neighbors = floor(sqrt(N))).  Same arithmetic as the search: sequential fp64
//                         sum of squared differences, separate multiply and add, then sqrt.
#include "cb200_common.cuh"

#define QC_MAX_SUBSETS 8

// One warp per cell.  Sums are integer-valued counts in fp64: exact below 2^53 whatever the order.
template <int S>
__global__ void __launch_bounds__(256)
k_qc_metrics(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, const double *__restrict__ x,
             const uint32_t *__restrict__ mask, int64_t n_cols, int64_t G, double limit, double *__restrict__ sums,
             int32_t *__restrict__ det, int *__restrict__ flag)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        double s[S + 1];
        int n[S + 1];
#pragma unroll
        for (int q = 0; q <= S; ++q) {
            s[q] = 0.0;
            n[q] = 0;
        }
        for (int64_t i = b + lane; i < e; i += 32) {
            const double v = ld_stream(x + i);
            const int g = ld_stream_i(rowidx + i);
            if (g < 0 || g >= G) {
                atomicOr(flag, 2);
                continue;
            }
            const bool hit = v > limit;
            s[0] += v;
            n[0] += hit ? 1 : 0;
            if (S > 0) {
                const uint32_t m = __ldg(mask + g);
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    const bool in = (m >> q) & 1u;
                    s[q + 1] += in ? v : 0.0;
                    n[q + 1] += (in && hit) ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int q = 0; q <= S; ++q) {
            const double t = warp_sum(s[q]);
            const int m = warp_sum_i(n[q]);
            if (lane == 0) {
                sums[(int64_t)q * n_cols + c] = t;
                det[(int64_t)q * n_cols + c] = m;
            }
        }
    }
}

__global__ void k_qc_mask(const int32_t *__restrict__ genes, int64_t n, int bit, int64_t G, uint32_t *__restrict__ mask,
                          int *__restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const int g = genes[i];
        if (g < 0 || g >= G)
            atomicOr(flag, 1);
        else
            atomicOr(mask + g, 1u << bit);
    }
}

extern "C" int cb200_qc_metrics(cb200_ctx *ctx, int n_subsets, const int32_t *subset_ptr, const int32_t *subset_genes,
                                double detection_limit, double *sum_out, int32_t *detected_out, double *subset_sum_out,
                                int32_t *subset_detected_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded, "cb200_qc_metrics: no count matrix loaded");
    CB_CHECK(ctx, n_subsets >= 0 && n_subsets <= QC_MAX_SUBSETS, "cb200_qc_metrics: 0 <= n_subsets <= %d (got %d)",
             QC_MAX_SUBSETS, n_subsets);
    CB_CHECK(ctx, n_subsets == 0 || (subset_ptr != nullptr && subset_genes != nullptr), "cb200_qc_metrics: NULL subset arrays");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int S = n_subsets;
    const int64_t N = ctx->N, G = ctx->G;
    CB_CUDA(ctx, ctx->qc_mask.ensure((size_t)G));
    CB_CUDA(ctx, ctx->qc_sum.ensure((size_t)(QC_MAX_SUBSETS + 1) * N));   // the 5..8-subset case runs the 8-subset kernel
    CB_CUDA(ctx, ctx->qc_det.ensure((size_t)(QC_MAX_SUBSETS + 1) * N));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->qc_mask.p, 0, (size_t)G * sizeof(uint32_t), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
    if (S > 0) {
        CB_CHECK(ctx, subset_ptr[0] == 0, "cb200_qc_metrics: subset_ptr[0] must be 0");
        for (int q = 0; q < S; ++q)
            CB_CHECK(ctx, subset_ptr[q + 1] >= subset_ptr[q], "cb200_qc_metrics: subset_ptr must be non-decreasing");
        const int64_t tot = subset_ptr[S];
        if (tot > 0) {
            CB_CUDA(ctx, ctx->stage_i.ensure((size_t)tot));
            CB_TRY(cb200_h2d(ctx, ctx->stage_i.p, subset_genes, (size_t)tot * sizeof(int32_t), ctx->stream));
            for (int q = 0; q < S; ++q) {
                const int64_t n = subset_ptr[q + 1] - subset_ptr[q];
                if (n == 0)
                    continue;
                k_qc_mask<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->stage_i.p + subset_ptr[q], n, q, G,
                                                                               ctx->qc_mask.p, ctx->flag.p);
                CB_LAUNCH(ctx);
            }
        }
    }
    CB_TRY(cb200_need_rowidx(ctx));
    const int grid = cb_grid_for(N, 8, ctx->num_sms * 8);
#define QC_LAUNCH(SS)                                                                                                   \
    k_qc_metrics<SS><<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->rowidx.p, ctx->x.p, ctx->qc_mask.p, N, G,       \
                                                    detection_limit, ctx->qc_sum.p, ctx->qc_det.p, ctx->flag.p)
    switch (S) {
    case 0: QC_LAUNCH(0); break;
    case 1: QC_LAUNCH(1); break;
    case 2: QC_LAUNCH(2); break;
    case 3: QC_LAUNCH(3); break;
    case 4: QC_LAUNCH(4); break;
    default: QC_LAUNCH(QC_MAX_SUBSETS); break;   // 5..8 subsets: unused bits stay empty
    }
#undef QC_LAUNCH
    CB_LAUNCH(ctx);
    int h_flag = 0;
    CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
    CB_CHECK(ctx, (h_flag & 1) == 0, "cb200_qc_metrics: a subset names a gene outside [0, n_genes)");
    CB_CHECK(ctx, (h_flag & 2) == 0, "cb200_qc_metrics: a row index of the count matrix is outside [0, n_genes)");
    if (sum_out != nullptr)
        CB_TRY(cb200_d2h(ctx, sum_out, ctx->qc_sum.p, (size_t)N * sizeof(double), ctx->stream));
    if (detected_out != nullptr)
        CB_TRY(cb200_d2h(ctx, detected_out, ctx->qc_det.p, (size_t)N * sizeof(int32_t), ctx->stream));
    if (S > 0 && subset_sum_out != nullptr)
        CB_TRY(cb200_d2h(ctx, subset_sum_out, ctx->qc_sum.p + N, (size_t)S * N * sizeof(double), ctx->stream));
    if (S > 0 && subset_detected_out != nullptr)
        CB_TRY(cb200_d2h(ctx, subset_detected_out, ctx->qc_det.p + N, (size_t)S * N * sizeof(int32_t), ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// thread per (query, neighbour); output column-major nq x k like the index
__global__ void k_knn_dist(const double *__restrict__ E, int64_t lde, int d, int64_t q_begin, int64_t nq, int k,
                           const int32_t *__restrict__ knn, double *__restrict__ out)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= nq * k)
        return;
    const int64_t q = idx % nq;
    const int c = (int)(idx / nq);
    const double *a = E + (q_begin + q) * lde;
    const double *b = E + (int64_t)knn[q * k + c] * lde;
    double s = 0.0;
    for (int t = 0; t < d; ++t) {
        const double df = __dsub_rn(a[t], b[t]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    out[idx] = __dsqrt_rn(s);
}

extern "C" int cb200_knn_distances(cb200_ctx *ctx, double *dist_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->have_knn && ctx->knn_E != nullptr, "cb200_knn_distances: no kNN result in the context");
    CB_CHECK(ctx, dist_out != nullptr, "cb200_knn_distances: NULL output");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nq = ctx->knn_nq;
    const int k = ctx->knn_k;
    // stage_d may hold the column-major upload of the embedding; the scores / emb buffers are what knn_E points to
    CB_CUDA(ctx, ctx->qc_sum.ensure((size_t)nq * k));
    k_knn_dist<<<(unsigned)((nq * k + 255) / 256), 256, 0, ctx->stream>>>(ctx->knn_E, ctx->knn_lde, ctx->knn_d, ctx->knn_qbegin,
                                                                        nq, k, ctx->knn.p, ctx->qc_sum.p);
    CB_LAUNCH(ctx);
    CB_TRY(cb200_d2h(ctx, dist_out, ctx->qc_sum.p, (size_t)nq * k * sizeof(double), ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
