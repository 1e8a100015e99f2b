/*
 * cellula_b200.h -- plain C ABI of libcellula_b200.so
 *
 * B200-native (sm_100a) implementation of cellula's data-parallel hot path:
 *     doNormAndReduce()        /root/reference/R/doNormAndReduce.R:37-109
 *     makeGraphsAndClusters()  /root/reference/R/makeGraphsAndClusters.R:46-157
 * from raw counts to the SNN edge list.  The reference is pure R;
 * the arithmetic it reaches lives in Bioconductor packages.  Each entry point below names
 * the reference call site (file:line) and the upstream routine it replaces.  The R side
 * binds these through a .Call shim that only unpacks SEXPs (r-pkg/src/rshim.c,
 * INTEGRATION.md); the Python mirror binds them through ctypes (cellula_b200/_lib.py).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is available
 *     from cb200_last_error(ctx).  Nothing throws or longjmps across this boundary.
 *   - host input buffers are owned by the caller, read-only, never retained after return.
 *   - "nullable" output pointers may be NULL when the caller does not need that result
 *     on the host (it stays resident in HBM for the following stages).
 *   - one cb200_ctx drives one GPU.  A multi-GPU job is one process (or thread) per GPU,
 *     each with its own context holding a contiguous shard of cells; the *_local /
 *     *_finish pairs expose the points where the caller reduces partial results
 *     (NCCL all-reduce / all-gather on the device pointers from cb200_device_ptr).
 *     The unsuffixed forms are the single-GPU composition local -> finish.
 *   - all matrices crossing the boundary use R's layouts: dgCMatrix slots for counts
 *     (0-based row indices, column pointers), column-major dense matrices, 1-based
 *     neighbour indices.
 */
#ifndef CELLULA_B200_H
#define CELLULA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb200_ctx cb200_ctx;

/* SNN weighting schemes: bluster::makeSNNGraph(type=) as passed by
 * R/makeGraphsAndClusters.R:83-85 (weighting_scheme). */
enum { CB200_SNN_RANK = 0, CB200_SNN_NUMBER = 1, CB200_SNN_JACCARD = 2 };

/* device buffers a caller may reduce / gather across GPUs (cb200_device_ptr) */
enum {
    CB200_BUF_LIBSIZE_PARTIAL = 0, /* double[2]: local sum of library sizes, local #cells    */
    CB200_BUF_GENE_ACC = 1,        /* int64[2*G]: per-gene sum(y), sum(y^2) over local cells as FIXED-POINT
                                      integers (exact, order-independent): all-reduce with an integer sum */
    CB200_BUF_GRAM = 2,            /* local X^T X of the HVG log-counts (upper triangle): int64[2*H*H] (fixed
                                      point, gram_engine 3: all-reduce with an integer sum) or double[H*H]   */
    CB200_BUF_SCORES = 3,          /* double[n_local*d] row-major: local PCA scores           */
    CB200_BUF_KNN = 4,             /* int32[n_query*k] row-major, 0-based global indices      */
    CB200_BUF_SIZE_FACTORS = 5,    /* double[n_local]                                         */
    CB200_BUF_LOGCOUNTS = 6,       /* double[nnz_local]                                       */
    CB200_BUF_BLOCK_PART = 7,      /* double[2B+1]: per block sum of size factors, #cells (SUM-reduce); max colsum/sf (MAX)  */
    CB200_BUF_BLOCK_ACC = 8,       /* int64[B][2][G]: per-block fixed-point gene sums (integer SUM-reduce)                 */
    CB200_BUF_COUNT_ = 9
};

/* per-stage device times of the most recent call of each stage, milliseconds (CUDA events
 * on the context's stream), plus launch counts; see cb200_stage_timings */
typedef struct cb200_timing {
    double h2d_ms;          /* cb200_load_csc copies                                  */
    double colsum_ms;       /* K1  k_colsum_csc                                       */
    double lognorm_ms;      /* K2+K3 k_lognorm_genestats                              */
    double hvg_extract_ms;  /* K5a HVG sub-matrix extraction                          */
    double gram_ms;         /* K5b Gram / cross-product                               */
    double eig_ms;          /* K5c subspace iteration on the H x H covariance         */
    double project_ms;      /* K5d scores = Xc V                                      */
    double knn_prep_ms;     /* K6a embedding conversion, norms                        */
    double knn_tile_ms;     /* K6b distance tiles + fused top-k' selection            */
    double knn_rerank_ms;   /* K6c exact fp64 re-rank + margin check (+ fallback)     */
    double snn_revlist_ms;  /* K7a reverse neighbour lists                            */
    double snn_edges_ms;    /* K7b sorted-intersection edge weighting                 */
    double d2h_ms;          /* result copies of the last stage that made any          */
    int64_t kernel_launches; /* kernels of this library launched since cb200_reset_timings */
    int64_t knn_fallback_queries; /* queries answered by the exact fp64 fallback      */
    int64_t eig_iterations;
    int64_t gram_engine;    /* cross-product engine of the last PCA: 1 = fp64 L2 reductions, 2 = fp64 shared-memory tiles, 3 = tcgen05 int8 digits */
    int64_t knn_tiles_scanned; /* 128 x 128 distance tiles computed by the last kNN call ...          */
    int64_t knn_tiles_total;   /* ... out of this many (the rest was pruned on the PC-1 bound)          */
    int64_t knn_engine;     /* tile engine of the last kNN call: 1 = fp32 CUDA cores, 2 = tcgen05 tensor cores */
} cb200_timing;

/* ---- context ------------------------------------------------------------------------ */

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t to run on (e.g.
 * the caller's current stream so that its own events bracket the work) or NULL to let the
 * library create one. */
/* STREAM CONTRACT for callers that run their own collectives on the library's buffers (cb200_device_ptr) between the
 * *_local and *_finish halves: the library's kernels are asynchronous on the context's stream, so such a caller must
 * issue its collectives on THAT stream (pass it here), or bracket them with cb200_synchronize().  A NULL `stream`
 * gives a private cudaStreamNonBlocking stream that nothing else is ordered against; handle 0 (the legacy default
 * stream) cannot be requested.  cellula_b200/dist.py:CudaShard refuses a context that is not on torch's current
 * stream; cb200_multi_* runs NCCL on each context's own stream and needs nothing from the caller. */
int cb200_init(int device, void *stream, cb200_ctx **out);
int cb200_free(cb200_ctx *ctx);
/* Message of the last failure on this context ("" if none); ctx == NULL returns the
 * message of the last failed cb200_init. */
const char *cb200_last_error(const cb200_ctx *ctx);
/* Library/ABI version and the -gencode it was built for, e.g. "cellula_b200 0.1 sm_100a". */
const char *cb200_version(void);
int cb200_synchronize(cb200_ctx *ctx);
int cb200_stage_timings(cb200_ctx *ctx, cb200_timing *out);
int cb200_reset_timings(cb200_ctx *ctx);
/* Device pointer and size in bytes of one of the CB200_BUF_* buffers (valid until the next
 * stage that reallocates it). */
int cb200_device_ptr(cb200_ctx *ctx, int which, void **ptr, int64_t *bytes);

/* ---- a1: counts(sce), a dgCMatrix (genes x cells, CSC) -------------------------------
 * Replaces: the assay every step of R/doNormAndReduce.R:64-97 reads.
 * colptr is int64[n_cells+1]; cb200_load_csc_i32 takes dgCMatrix's own int32 `p` slot.
 * rowidx int32[nnz] 0-based gene index, sorted within a column; values double[nnz].
 * cell_offset / n_cells_total describe the shard: this context holds global cells
 * [cell_offset, cell_offset + n_cells); single GPU: 0 and n_cells.
 * The copies are asynchronous when the host arrays are page-locked (values first, row indices behind them on
 * their own stream): keep the arrays valid until cb200_lognorm_genestats[_local] or cb200_synchronize has returned. */
int cb200_load_csc(cb200_ctx *ctx, int64_t n_genes, int64_t n_cells, const int64_t *colptr,
                   const int32_t *rowidx, const double *values, int64_t cell_offset,
                   int64_t n_cells_total);
int cb200_load_csc_i32(cb200_ctx *ctx, int64_t n_genes, int64_t n_cells, const int32_t *colptr,
                       const int32_t *rowidx, const double *values, int64_t cell_offset,
                       int64_t n_cells_total);

/* ---- a2: size factors ----------------------------------------------------------------
 * Replaces: scuttle::logNormCounts() -> librarySizeFactors()/sizeFactors() centring,
 * R/doNormAndReduce.R:78.  sf = ls / mean(ls); a caller-supplied vector (deconvolution
 * factors from R/doNormAndReduce.R:64-71) is only re-centred.  Fails with
 * "size factors should be positive" like scuttle. */
int cb200_size_factors_local(cb200_ctx *ctx, const double *sf_in /* nullable, local cells */);
int cb200_size_factors_finish(cb200_ctx *ctx, double total, int64_t n_cells_total,
                              double *sf_out /* nullable, local cells */);
int cb200_size_factors(cb200_ctx *ctx, const double *sf_in, double *sf_out);

/* ---- a3 + a4: logcounts and per-gene mean / variance ---------------------------------
 * Replaces: scuttle::normalizeCounts(log=TRUE, pseudo.count=1) (R/doNormAndReduce.R:78)
 * and scran:::.compute_mean_var inside modelGeneVar() (R/doNormAndReduce.R:88), fused:
 * logx = log1p(x / sf[col]) / log(2) on the stored non-zeros; mean/var over all cells
 * with implicit zeros, var with n-1. */
int cb200_lognorm_genestats_local(cb200_ctx *ctx);
int cb200_genestats_finish(cb200_ctx *ctx, double *gene_mean /* nullable [G] */,
                           double *gene_var /* nullable [G] */);
/* K2 alone: logcounts from the values and the column pointers (no gene sums yet).  The loaders send the
 * values first and the row indices behind them, so this -- and cb200_fetch_logcounts_async -- can run while
 * the indices are still arriving; cb200_lognorm_genestats_local afterwards only accumulates the gene sums. */
int cb200_lognorm_values(cb200_ctx *ctx);
int cb200_fetch_logcounts(cb200_ctx *ctx, double *logx_out /* [nnz_local] */);
/* Same copy issued on the context's copy stream so that it overlaps the following stages;
 * logx_out should be page-locked.  cb200_wait_copies blocks until such copies have landed. */
int cb200_fetch_logcounts_async(cb200_ctx *ctx, double *logx_out /* [nnz_local] */);
int cb200_wait_copies(cb200_ctx *ctx);
int cb200_lognorm_genestats(cb200_ctx *ctx, double *logx_out /* nullable */, double *gene_mean,
                            double *gene_var);

/* ---- the batch-aware branch -------------------------------------------------------------------
 * Replaces the arithmetic of batchelor::multiBatchNorm(sce, batch = ) and scran::modelGeneVar(sce, block = ) in
 * /root/reference/R/doNormAndReduce.R:74-76 and :84-86 (reached from cellula(batch = ...), R/cellula.r:145-150).
 * block: host int32[n_local], 0-based ids in [0, n_blocks) (levels of the batch factor).
 *   cb200_block_averages: per batch b and gene g the mean over the batch's cells of x / (sf / mean_b(sf)) --
 *     scuttle::calculateAverage with batch-centred size factors, what multiBatchNorm compares between batches.
 *     ave_out row-major B x G; sf_mean_out[B] = mean_b(sf); n_out[B] = cells per batch.  The B x B median-ratio
 *     rescaling (batchelor:::.rescale_size_factors) is G-sized host code (r-pkg/R, cellula_b200/api.py); its result
 *     comes back through cb200_size_factors_given, which takes the factors AS THEY ARE (no centring), like
 *     logNormCounts(center.size.factors = FALSE) inside multiBatchNorm.
 *   cb200_lognorm_genestats_blocked: logcounts as usual; mean / variance per gene WITHIN every block (row-major
 *     B x G, NaN where a block has < 1 / < 2 cells); the trend fits per block and their combination stay host code.
 * _local / _finish forms for multi-GPU callers: cb200_block_prepare_local fills CB200_BUF_BLOCK_PART (SUM-reduce the
 * first 2B doubles, MAX-reduce the last), cb200_block_averages_local / cb200_lognorm_genestats_blocked_local fill
 * CB200_BUF_BLOCK_ACC (integer SUM-reduce), the _finish calls take the reduced block sizes. */
int cb200_set_blocks(cb200_ctx *ctx, const int32_t *block, int n_blocks);
int cb200_block_prepare_local(cb200_ctx *ctx);
int cb200_block_averages_local(cb200_ctx *ctx, double max_ratio);
int cb200_block_averages_finish(cb200_ctx *ctx, const double *part_global /* [2B] */, double *ave_out /* [B*G] */);
int cb200_block_averages(cb200_ctx *ctx, const int32_t *block, int n_blocks, double *ave_out, double *sf_mean_out /* nullable */,
                         int64_t *n_out /* nullable */);
int cb200_size_factors_given(cb200_ctx *ctx, const double *sf /* [n_local], used as they are */);
int cb200_lognorm_genestats_blocked_local(cb200_ctx *ctx);
int cb200_genestats_blocked_finish(cb200_ctx *ctx, const int64_t *n_per_block /* [B], all ranks */, double *gene_mean /* nullable [B*G] */,
                                   double *gene_var /* nullable [B*G] */);
int cb200_lognorm_genestats_blocked(cb200_ctx *ctx, double *logx_out /* nullable */, double *gene_mean, double *gene_var);

/* ---- a7: PCA on the HVG subset -------------------------------------------------------
 * Replaces: scater::runPCA(sce, subset_row=hvgs, exprs_values="logcounts",
 * ncomponents=ndims), R/doNormAndReduce.R:94-97 -> BiocSingular::runPCA.
 * hvg_idx: 0-based gene rows in getTopHVGs order.  scores column-major n_local x d (U D),
 * rotation column-major H x d (V), var_explained[d] = D^2/(N-1), total_var = sum of the
 * HVG column variances (percentVar = 100 var_explained / total_var). */
int cb200_pca_gram_local(cb200_ctx *ctx, const int32_t *hvg_idx, int n_hvg);
int cb200_pca_finish(cb200_ctx *ctx, int d, double *scores /* nullable */,
                     double *rotation /* nullable */, double *var_explained /* nullable */,
                     double *total_var /* nullable */);
int cb200_pca(cb200_ctx *ctx, const int32_t *hvg_idx, int n_hvg, int d, double *scores,
              double *rotation, double *var_explained, double *total_var);

/* ---- a9: exact kNN -------------------------------------------------------------------
 * Replaces: BiocNeighbors::findKNN(x, k, BNPARAM=KmknnParam(), get.distance=FALSE) inside
 * bluster::makeSNNGraph(), R/makeGraphsAndClusters.R:83-85 (:125-127, :284).
 * emb: host, column-major n_total x d (the `space` matrix, :80-81); NULL = use the scores
 * left in HBM by cb200_pca_finish (single GPU only).  Queries are the global rows
 * [q_begin, q_end); index_out is column-major (q_end-q_begin) x k, 1-based, each row
 * ordered by (distance, index) with distance = sqrt(sum (a-b)^2) in fp64, self excluded.
 * The result is exact: tensor/fp32 distances only nominate candidates, an fp64 re-rank
 * with a proven error margin (or an fp64 full scan) decides. */
int cb200_knn(cb200_ctx *ctx, const double *emb, int64_t n_total, int d, int k, int64_t q_begin,
              int64_t q_end, int32_t *index_out /* nullable */);
/* Same with the full embedding already in HBM: row-major n_total x d doubles. */
int cb200_knn_device(cb200_ctx *ctx, const double *d_emb_rowmajor, int64_t n_total, int d, int k,
                     int64_t q_begin, int64_t q_end, int32_t *index_out /* nullable */);

/* ---- a10: SNN edges ------------------------------------------------------------------
 * Replaces: bluster::neighborsToSNNGraph(index, type) -> libscran build_snn_graph inside
 * makeSNNGraph(), R/makeGraphsAndClusters.R:83-85.
 * index: host, column-major n_total x k, 1-based (findKNN()$index); NULL = use the result
 * left in HBM by cb200_knn (single GPU, all queries).  Edges are produced for the cells
 * j in [j_begin, j_end) as (from = j+1, to = h+1, h < j), grouped by j ascending, within
 * a group ordered as libscran emits them (first encounter).  The arrays are allocated by
 * the library (pinned host memory); release with cb200_free_edges. */
int cb200_snn(cb200_ctx *ctx, const int32_t *index, int64_t n_total, int k, int type,
              int64_t j_begin, int64_t j_end, int64_t *n_edges, int32_t **from, int32_t **to,
              double **weight);
/* Same with the full 0-based row-major n_total x k index already in HBM.  from/to/weight may
 * be NULL: the edges then stay resident in HBM and only the count is returned. */
int cb200_snn_device(cb200_ctx *ctx, const int32_t *d_index_rowmajor0, int64_t n_total, int k,
                     int type, int64_t j_begin, int64_t j_end, int64_t *n_edges, int32_t **from,
                     int32_t **to, double **weight);
/* Device-only variant: edges stay in HBM (bench "value" leg); returns the count. */
int cb200_snn_resident(cb200_ctx *ctx, int k, int type, int64_t *n_edges);
/* Two-phase form for callers that own the result vectors (what the R shim does: count, allocVector,
 * emit): cb200_snn_count_device builds the reverse lists and counts the edges of cells
 * [j_begin, j_end) (d_index_rowmajor0 NULL = the resident kNN result); cb200_snn_emit then writes
 * them -- into from/to/weight (host, n_edges entries each; the device->host copies overlap the
 * emission) or, with NULL pointers, only into HBM (cb200_fetch_edges copies them later). */
int cb200_snn_count_device(cb200_ctx *ctx, const int32_t *d_index_rowmajor0, int64_t n_total, int k,
                           int type, int64_t j_begin, int64_t j_end, int64_t *n_edges);
int cb200_snn_emit(cb200_ctx *ctx, int32_t *from, int32_t *to, double *weight);
int cb200_free_edges(int32_t *from, int32_t *to, double *weight);
/* Copies the n_edges edges left in HBM by the last cb200_snn / cb200_snn_device / cb200_snn_resident
 * call into caller-owned arrays (what the R shim does with freshly allocated R vectors). */
int cb200_fetch_edges(cb200_ctx *ctx, int64_t n_edges, int32_t *from, int32_t *to, double *weight);

/* Order-independent digest of the edge list left in HBM by the last cb200_snn* call: sum over the edges of a
 * 64-bit hash of (from, to, weight bits) modulo 2^64, and the edge count.  The digests of the shards of a
 * multi-GPU run add up to the single-GPU digest (bench.py and the tests compare them). */
int cb200_edges_checksum(cb200_ctx *ctx, int64_t *n_edges, uint64_t *checksum);

/* ---- per-resolution cluster diagnostics -------------------------------------------
 * Replaces: bluster::approxSilhouette(space, clusters) and bluster::pairwiseModularity(g, clusters, ...)
 * in /root/reference/R/makeGraphsAndClusters.R:103-114 (:144-153, :288).  labels: 0-based positions in the
 * sorted unique cluster labels, host int32[n]; a value without cells is not a cluster.
 *
 * cb200_approx_silhouette: emb = host column-major n x d, or NULL for the PCA scores resident in the
 * context (first d components).  width_out[n] = (other - self) / max(other, self) with the root-mean-square
 * distances to the own / the closest other cluster; other_out[n] = that other cluster (0-based, -1 if none). */
int cb200_approx_silhouette(cb200_ctx *ctx, const double *emb, int64_t n, int d, const int32_t *labels,
                            int n_clusters, double *width_out, int32_t *other_out);
/* cb200_pairwise_modularity: edges 1-based as cb200_snn returns them (from == NULL: the edge list left in HBM
 * by the last cb200_snn* call).  observed_out / expected_out: row-major K x K, upper triangle filled
 * (bluster's get.weights = TRUE pair); total_weight_out (nullable) = sum of the edge weights.  The caller
 * forms observed / expected (as.ratio = TRUE) or (observed - expected) / total. */
int cb200_pairwise_modularity(cb200_ctx *ctx, const int32_t *from, const int32_t *to, const double *weight,
                              int64_t n_edges, const int32_t *labels, int64_t n, int n_clusters,
                              double *observed_out, double *expected_out, double *total_weight_out);

/* ---- deconvolution (pooled) size factors ---------------------------------------------
 * Replaces: scran::computeSumFactors(sce, clusters = sce_cl) = scuttle::pooledSizeFactors in
 * /root/reference/R/doNormAndReduce.R:66-71 (the quickCluster call above it, :57-65, keeps its igraph walktrap;
 * its labels are the `clusters` input).  Needs the whole count matrix resident in this context (cb200_load_csc
 * with all cells).  clusters: host int32[N] codes >= 0, or NULL for one cluster; sizes: pool sizes (NULL =
 * seq(21, 101, 5)), sizes above a cluster's cell count are dropped for that cluster; min_mean <= 0 = scuttle's
 * guess (0.1 when the median library size is <= 50000, else 1); max_cluster_size <= 0 = no limit (scuttle's
 * default is 3000).  sf_out[N]: the factors, scaled to unit mean over the positive ones; the count of
 * non-positive factors (scuttle would warn and, with positive = TRUE, call cleanSizeFactors -- that repair stays
 * with the caller) goes to n_nonpositive.  stage_ms_out (nullable, 8 doubles): pseudo-cells + gene filter +
 * compact store, pool medians, least-squares solves (device ms), host rescaling, whole call (host ms), then the
 * largest number of kept genes of a cluster, the number of clusters after the size limit, the kept entries. */
int cb200_pooled_size_factors(cb200_ctx *ctx, const int32_t *clusters /* nullable */, const int32_t *sizes /* nullable */,
                              int n_sizes, double min_mean, int64_t max_cluster_size, double *sf_out,
                              int64_t *n_nonpositive /* nullable */, double *stage_ms_out /* nullable */);

/* ---- the callers either side of the path that re-use its kernels ---------------------
 * cb200_qc_metrics replaces scuttle::perCellQCMetrics(sce, subsets = list(mito =, Malat1 =, Ribo =)) in
 * /root/reference/R/doQC.R:106-108: per cell the total count and the number of genes above detection_limit
 * (scuttle's threshold = 0), and the same pair over each of n_subsets <= 8 gene subsets given CSR-style
 * (subset_ptr[n_subsets + 1], subset_genes 0-based).  Outputs (host, nullable): sum_out[N], detected_out[N],
 * subset_sum_out / subset_detected_out [n_subsets x N], subset s at offset s * N; subsets_*_percent is
 * 100 * subset sum / sum, formed by the caller. */
int cb200_qc_metrics(cb200_ctx *ctx, int n_subsets, const int32_t *subset_ptr, const int32_t *subset_genes,
                     double detection_limit, double *sum_out, int32_t *detected_out, double *subset_sum_out,
                     int32_t *subset_detected_out);
/* cb200_knn_distances: the distance matrix findKNN(get.distance = TRUE) returns beside the index, for the kNN
 * result left in HBM by the last cb200_knn / cb200_knn_device call (whose device embedding must still be alive):
 * dist_out column-major n_queries x k, sqrt of the sequential fp64 sum of squares.  Lets
 * uwot::umap(n_neighbors = floor(sqrt(N)), nn_method = list(idx, dist)) in
 * /root/reference/R/doNormAndReduce.R:102-106 re-use the exact search instead of running its own. */
int cb200_knn_distances(cb200_ctx *ctx, double *dist_out);

/* ---- several GPUs of one node from ONE host process (what an R session needs: SURVEY.md 8b) -----------------
 * cb200_multi_init(n_gpus) opens devices 0 .. n_gpus-1 (n_gpus <= 0: all), one cb200_ctx each, and an NCCL
 * communicator over them (ncclCommInitAll; NCCL is loaded with dlopen here, the single-GPU path does not need it).
 * Every call below fans out over one host thread per GPU: GPU r holds the contiguous block of cells
 * [r N / R, (r+1) N / R) and the stages exchange what SURVEY.md 8e lists (all-reduce of the library-size total,
 * of the fixed-point gene sums and of the fixed-point cross-product -- integer sums, so the results are
 * bit-identical to the single-GPU ones --, all-gather of the scores and of the neighbour index).  Host arrays
 * have the same layouts as in the single-GPU calls and cover ALL cells; pageable memory is staged through pinned
 * rings by each GPU's thread.  Errors: non-zero return, message from cb200_multi_last_error (m == NULL: of the
 * last failed cb200_multi_init).
 *   cb200_multi_knn: emb NULL = the scores left by cb200_multi_pca.  cb200_multi_snn_count / _emit: the two-phase
 *   form of cb200_snn (index NULL = the resident kNN result); the edges of GPU r land behind those of GPU r-1,
 *   which is the single-GPU emission order; from == NULL keeps them in HBM. */
typedef struct cb200_multi cb200_multi;
int cb200_multi_init(int n_gpus, cb200_multi **out);
int cb200_multi_free(cb200_multi *m);
const char *cb200_multi_last_error(const cb200_multi *m);
int cb200_multi_n_gpus(const cb200_multi *m);
int cb200_multi_stage_timings(cb200_multi *m, int rank, cb200_timing *out);
int cb200_multi_load_csc(cb200_multi *m, int64_t n_genes, int64_t n_cells, const int64_t *colptr, const int32_t *rowidx,
                         const double *values);
int cb200_multi_load_csc_i32(cb200_multi *m, int64_t n_genes, int64_t n_cells, const int32_t *colptr,
                             const int32_t *rowidx, const double *values);
int cb200_multi_size_factors(cb200_multi *m, const double *sf_in /* nullable [N] */, double *sf_out /* nullable [N] */);
int cb200_multi_lognorm_genestats(cb200_multi *m, double *logx_out /* nullable [nnz] */, double *gene_mean /* [G] */,
                                  double *gene_var /* [G] */);
int cb200_multi_pca(cb200_multi *m, const int32_t *hvg_idx, int n_hvg, int d, double *scores /* nullable N x d */,
                    double *rotation, double *var_explained, double *total_var);
int cb200_multi_knn(cb200_multi *m, const double *emb /* nullable */, int64_t n_total, int d, int k,
                    int32_t *index_out /* nullable N x k */);
int cb200_multi_snn_count(cb200_multi *m, const int32_t *index /* nullable */, int64_t n_total, int k, int type,
                          int64_t *n_edges);
int cb200_multi_snn_emit(cb200_multi *m, int32_t *from, int32_t *to, double *weight);
int cb200_multi_edges_checksum(cb200_multi *m, int64_t *n_edges, uint64_t *checksum);

/* ---- synthetic input (bench / test infrastructure, not part of the reference path) ---
 * ONE generator specification (cellula_b200/csrc/cb200_synth_core.h) compiled for the device and for the
 * host; both produce the same matrix bit for bit (integer hashing + individually rounded IEEE operations).
 * cb200_synth_counts fills the context with the negative-binomial count matrix with planted low-rank
 * structure as if by cb200_load_csc; cb200_export_csc copies it to host buffers.
 * cb200_synth_host is the host compilation (no GPU needed): it mallocs colptr[n_cells+1], rowidx[nnz],
 * values[nnz] for the global cells [cell_offset, cell_offset + n_cells); release with cb200_synth_host_free.
 * cb200_synth_embedding_host: the cfg-4 style PC matrix (Gaussian clusters, per-PC scale 20 -> 1), column-major
 * n x d doubles into a caller-owned array.  n_threads <= 0: all host cores. */
int cb200_synth_counts(cb200_ctx *ctx, int64_t n_genes, int64_t n_cells, double nnz_per_cell,
                       uint64_t seed, int64_t cell_offset, int64_t n_cells_total,
                       int64_t *nnz_out);
int cb200_synth_host(int64_t n_genes, int64_t n_cells, double nnz_per_cell, uint64_t seed, int64_t cell_offset,
                     int64_t n_cells_total, int n_threads, int64_t *nnz_out, int64_t **colptr_out,
                     int32_t **rowidx_out, double **values_out);
void cb200_synth_host_free(int64_t *colptr, int32_t *rowidx, double *values);
int cb200_synth_embedding_host(int64_t n, int d, uint64_t seed, int n_threads, double *emb_colmajor);
int cb200_export_csc(cb200_ctx *ctx, int64_t *colptr, int32_t *rowidx, double *values);

/* ---- host-side helper of the Python mirror (no device work, not part of the reference FFI) ----
 * limma::weightedLowess as scran::fitTrendVar uses it, on x sorted ascending.  The R package keeps
 * calling scran/limma; cellula_b200/trendvar.py uses this instead of interpreted loops. */
int cb200_host_weighted_lowess(int64_t n, const double *x_sorted, const double *y, const double *w, double span,
                               int iterations, int npts, double *fitted_out);
/* Gaussian kernel density of m[0..n) on `ng` uniform grid points in [lo, hi] (stats::density stand-in). */
/* Parametric mean-variance trend of scran::fitTrendVar: weighted least squares of
 * v ~ exp(A) m / (m^(1+exp(N)) + exp(B)), theta = (A, B, N) updated in place from its start value. */
int cb200_host_nls_trend(int64_t n, const double *v, const double *m, const double *w, double *theta, double tol,
                         int maxiter);
/* grid search for the start values of the parametric trend (scran's .get_nls_starts); first minimum, b fastest */
int cb200_host_nls_grid(int64_t n, const double *v, const double *m, double grad, int nb, const double *bgrid, int nn,
                        const double *ngrid, int *ib_out, int *in_out);
/* Host layout of cb200_pooled_size_factors (no device work; exported so that the CPU test suite can check it against the
 * oracle): scuttle:::.limit_cluster_size + .generateSphere.  cl_out[n] = cluster codes after the size limit, ring_out[n] =
 * cells in ring order (clusters concatenated), cl_ptr_out[n + 1] (first *n_clusters_out + 1 entries used), meanlib_out[n]
 * (nullable; first *n_clusters_out used) = mean library size per cluster. */
int cb200_host_pool_layout(int64_t n, const int32_t *clusters /* nullable */, const double *lib, int64_t max_cluster_size,
                           int32_t *cl_out, int32_t *ring_out, int64_t *cl_ptr_out, double *meanlib_out,
                           int32_t *n_clusters_out);
int cb200_host_gauss_kde(int64_t n, const double *m, double bw, double lo, double hi, int ng, double *dens_out);

#ifdef __cplusplus
}
#endif
#endif /* CELLULA_B200_H */
