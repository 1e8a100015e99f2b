// cb200_norm.cu -- K1 library-size factors, K2+K3 fused log-normalisation and per-gene
// mean / variance accumulation.  HBM-bound streaming passes over the CSC count matrix.
//
//   K1  k_colsum_csc            replaces scuttle::librarySizeFactors (colSums / mean),
//                               reached from logNormCounts(), R/doNormAndReduce.R:78
//   K2  k_lognorm_genestats     replaces scuttle::normalizeCounts(log=TRUE): log1p(x/sf)/log(2)
//   K3  (same kernel)           replaces scran:::.compute_mean_var inside modelGeneVar(),
//                               R/doNormAndReduce.R:88 (mean and n-1 variance, zeros included)
//
// Algorithmic bytes: K1 8 B/nnz (+16 B/cell); K2+K3 20 B/nnz (read x 8, read i 4,
// write logcounts 8) + 8 B/cell + 16 B/gene.
#include "cb200_common.cuh"

// One warp per column (cell).  Interior of the column is read with 128-bit loads.
__global__ void __launch_bounds__(256) k_colsum_csc(const int64_t *__restrict__ colptr, const double *__restrict__ x,
                                                    int64_t n_cols, double *__restrict__ libsize)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        int64_t b2 = (b + 1) & ~1LL;
        if (b2 > e)
            b2 = e;
        int64_t e2 = e & ~1LL;
        if (e2 < b2)
            e2 = b2;
        double s = 0.0;
        if (lane == 0 && b2 > b)
            s += x[b];
        if (lane == 1 && e > e2)
            s += x[e2];
        const int64_t p0 = b2 >> 1, p1 = e2 >> 1;
        int64_t i = p0 + lane;
        // 4 independent 128-bit loads in flight per lane
        for (; i + 96 < p1; i += 128) {
            double2 v0 = ld_stream2(x2 + i);
            double2 v1 = ld_stream2(x2 + i + 32);
            double2 v2 = ld_stream2(x2 + i + 64);
            double2 v3 = ld_stream2(x2 + i + 96);
            s += (v0.x + v0.y) + (v1.x + v1.y) + (v2.x + v2.y) + (v3.x + v3.y);
        }
        for (; i < p1; i += 32) {
            double2 v = ld_stream2(x2 + i);
            s += v.x + v.y;
        }
        s = warp_sum(s);
        if (lane == 0)
            libsize[c] = s;
    }
}

// Deterministic sum of a vector: one block, fixed assignment, fixed tree.
__global__ void __launch_bounds__(1024) k_sum_vector(const double *__restrict__ v, int64_t n, double n_as_double,
                                                     double *__restrict__ out2)
{
    __shared__ double sm[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
        s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            out2[0] = t;
            out2[1] = n_as_double;
        }
    }
}

__global__ void k_sf_finish(const double *__restrict__ libsize, int64_t n, double mean_ls, double *__restrict__ sf,
                            int *__restrict__ flag)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        double v = libsize[i] / mean_ls;
        sf[i] = v;
        if (!(v > 0.0) || isinf(v))
            atomicOr(flag, 1);
    }
}

extern "C" int cb200_size_factors_local(cb200_ctx *ctx, const double *sf_in)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded, "cb200_size_factors: no count matrix loaded");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_CUDA(ctx, ctx->libsize.ensure((size_t)ctx->N));
    CB_CUDA(ctx, ctx->sf.ensure((size_t)ctx->N));
    cb_stage_begin(ctx, ST_COLSUM);
    if (sf_in != nullptr) {
        CB_CUDA(ctx, cudaMemcpyAsync(ctx->libsize.p, sf_in, (size_t)ctx->N * sizeof(double), cudaMemcpyHostToDevice,
                                     ctx->stream));
    } else {
        int grid = cb_grid_for(ctx->N, 8, ctx->num_sms * 8);
        k_colsum_csc<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->x.p, ctx->N, ctx->libsize.p);
        CB_LAUNCH(ctx);
    }
    k_sum_vector<<<1, 1024, 0, ctx->stream>>>(ctx->libsize.p, ctx->N, (double)ctx->N, ctx->part.p);
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_COLSUM);
    return 0;
}

extern "C" int cb200_size_factors_finish(cb200_ctx *ctx, double total, int64_t n_cells_total, double *sf_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded && ctx->libsize.p != nullptr, "cb200_size_factors_finish: call cb200_size_factors_local first");
    CB_CHECK(ctx, n_cells_total > 0, "cb200_size_factors_finish: n_cells_total must be positive");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const double mean_ls = total / (double)n_cells_total;
    // scuttle: stop("size factors should be positive") -- also covers NA / all-zero input
    CB_CHECK(ctx, mean_ls > 0.0 && mean_ls == mean_ls && mean_ls < 1.0 / 0.0, "size factors should be positive");
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
    k_sf_finish<<<(unsigned)((ctx->N + 255) / 256), 256, 0, ctx->stream>>>(ctx->libsize.p, ctx->N, mean_ls, ctx->sf.p,
                                                                           ctx->flag.p);
    CB_LAUNCH(ctx);
    int h_flag = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(&h_flag, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (sf_out != nullptr)
        CB_TRY(cb200_read_back(ctx, sf_out, ctx->sf.p, (size_t)ctx->N * sizeof(double)));
    CB_CHECK(ctx, h_flag == 0, "size factors should be positive");
    ctx->have_sf = true;
    return 0;
}

extern "C" int cb200_size_factors(cb200_ctx *ctx, const double *sf_in, double *sf_out)
{
    if (ctx == nullptr)
        return 1;
    CB_TRY(cb200_size_factors_local(ctx, sf_in));
    double part[2];
    CB_TRY(cb200_read_back(ctx, part, ctx->part.p, sizeof(part)));
    return cb200_size_factors_finish(ctx, part[0], ctx->N, sf_out);
}

// ---------------------------------------------------------------------------------------
// K2 + K3
// ---------------------------------------------------------------------------------------
// One warp per column.  y = log1p(x / sf) / ln 2 is written back coalesced; the per-gene
// sums go to L2 with fp64 reductions (RED.ADD.F64 -- no return value, so no round trip).
__global__ void __launch_bounds__(256)
k_lognorm_genestats(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                    const double *__restrict__ x, const double *__restrict__ sf, int64_t n_cols,
                    double *__restrict__ logx, double *__restrict__ acc_sum, double *__restrict__ acc_sq)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double ln2 = 0.693147180559945309417232121458;
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        const double s = sf[c];
        int64_t i = b + lane;
        for (; i + 96 < e; i += 128) {
            double v[4];
            int g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = ld_stream(x + i + 32 * u);
                g[u] = ld_stream_i(rowidx + i + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                double y = log1p(v[u] / s) / ln2;
                st_stream(logx + i + 32 * u, y);
                atomicAdd(acc_sum + g[u], y);
                atomicAdd(acc_sq + g[u], y * y);
            }
        }
        for (; i < e; i += 32) {
            double y = log1p(ld_stream(x + i) / s) / ln2;
            int g = ld_stream_i(rowidx + i);
            st_stream(logx + i, y);
            atomicAdd(acc_sum + g, y);
            atomicAdd(acc_sq + g, y * y);
        }
    }
}

// K2 alone: logcounts from the values and the column pointers only (no row indices, no gene sums)
__global__ void __launch_bounds__(256)
k_lognorm_values(const int64_t *__restrict__ colptr, const double *__restrict__ x, const double *__restrict__ sf,
                 int64_t n_cols, double *__restrict__ logx)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double ln2 = 0.693147180559945309417232121458;
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        const double s = sf[c];
        int64_t i = b + lane;
        for (; i + 96 < e; i += 128) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = ld_stream(x + i + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                st_stream(logx + i + 32 * u, log1p(v[u] / s) / ln2);
        }
        for (; i < e; i += 32)
            st_stream(logx + i, log1p(ld_stream(x + i) / s) / ln2);
    }
}

// K3 alone: per-gene sums of the logcounts already in HBM (same arithmetic as the fused kernel: the
// values are read back bit for bit)
__global__ void __launch_bounds__(256)
k_genestats_from_logx(const int32_t *__restrict__ rowidx, const double *__restrict__ logx, int64_t nnz,
                      double *__restrict__ acc_sum, double *__restrict__ acc_sq)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const double y = ld_stream(logx + i);
        const int g = ld_stream_i(rowidx + i);
        atomicAdd(acc_sum + g, y);
        atomicAdd(acc_sq + g, y * y);
    }
}

__global__ void k_genestats_finish(const double *__restrict__ acc_sum, const double *__restrict__ acc_sq, int64_t G,
                                   double n_total, double *__restrict__ mean, double *__restrict__ var)
{
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < G) {
        double s = acc_sum[g], q = acc_sq[g];
        double m = s / n_total;
        double ss = q - s * m;  // sum (y - m)^2 over all cells, implicit zeros included
        if (ss < 0.0)
            ss = 0.0;
        mean[g] = m;
        var[g] = n_total > 1.0 ? ss / (n_total - 1.0) : nan("");
    }
}

extern "C" int cb200_lognorm_genestats_local(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded && ctx->have_sf, "cb200_lognorm_genestats: size factors have not been computed");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_CUDA(ctx, ctx->logx.ensure((size_t)ctx->nnz));
    CB_CUDA(ctx, ctx->gene_acc.ensure((size_t)(2 * ctx->G)));
    CB_TRY(cb200_need_rowidx(ctx));
    if (!(ctx->have_logx && ctx->logx_values_only))
        cb_stage_begin(ctx, ST_LOGNORM);   // (after cb200_lognorm_values the stage is already open)
    CB_CUDA(ctx, cudaMemsetAsync(ctx->gene_acc.p, 0, (size_t)(2 * ctx->G) * sizeof(double), ctx->stream));
    if (ctx->have_logx && ctx->logx_values_only) {
        // the logcounts exist already (cb200_lognorm_values): only the gene sums are left
        k_genestats_from_logx<<<ctx->num_sms * 16, 256, 0, ctx->stream>>>(ctx->rowidx.p, ctx->logx.p, ctx->nnz,
                                                                           ctx->gene_acc.p, ctx->gene_acc.p + ctx->G);
        ctx->logx_values_only = false;
    } else {
        int grid = cb_grid_for(ctx->N, 8, ctx->num_sms * 8);
        k_lognorm_genestats<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->rowidx.p, ctx->x.p, ctx->sf.p, ctx->N,
                                                           ctx->logx.p, ctx->gene_acc.p, ctx->gene_acc.p + ctx->G);
    }
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_LOGNORM);
    ctx->have_logx = true;
    return 0;
}

// Logcounts only (K2), from the values and the column pointers: callable while the row indices are still
// being uploaded, so that cb200_fetch_logcounts_async can start early; a following
// cb200_lognorm_genestats_local then only accumulates the gene sums.
extern "C" int cb200_lognorm_values(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded && ctx->have_sf, "cb200_lognorm_values: size factors have not been computed");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_CUDA(ctx, ctx->logx.ensure((size_t)ctx->nnz));
    cb_stage_begin(ctx, ST_LOGNORM);
    const int grid = cb_grid_for(ctx->N, 8, ctx->num_sms * 8);
    k_lognorm_values<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->x.p, ctx->sf.p, ctx->N, ctx->logx.p);
    CB_LAUNCH(ctx);
    ctx->have_logx = true;
    ctx->logx_values_only = true;
    return 0;
}

extern "C" int cb200_genestats_finish(cb200_ctx *ctx, double *gene_mean, double *gene_var)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->have_logx, "cb200_genestats_finish: call cb200_lognorm_genestats_local first");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_CUDA(ctx, ctx->gene_mean.ensure((size_t)ctx->G));
    CB_CUDA(ctx, ctx->gene_var.ensure((size_t)ctx->G));
    k_genestats_finish<<<(unsigned)((ctx->G + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->gene_acc.p, ctx->gene_acc.p + ctx->G, ctx->G, (double)ctx->N_total, ctx->gene_mean.p, ctx->gene_var.p);
    CB_LAUNCH(ctx);
    if (gene_mean != nullptr)
        CB_TRY(cb200_read_back(ctx, gene_mean, ctx->gene_mean.p, (size_t)ctx->G * sizeof(double)));
    if (gene_var != nullptr)
        CB_TRY(cb200_read_back(ctx, gene_var, ctx->gene_var.p, (size_t)ctx->G * sizeof(double)));
    ctx->have_genestats = true;
    return 0;
}

extern "C" int cb200_fetch_logcounts(cb200_ctx *ctx, double *logx_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->have_logx, "cb200_fetch_logcounts: logcounts have not been computed");
    CB_CHECK(ctx, logx_out != nullptr, "cb200_fetch_logcounts: NULL output");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_stage_begin(ctx, ST_D2H);
    CB_CUDA(ctx, cudaMemcpyAsync(logx_out, ctx->logx.p, (size_t)ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    cb_stage_end(ctx, ST_D2H);
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Asynchronous variant: the copy runs on the context's copy stream behind the kernel that produced
// the logcounts, so it overlaps the following stages (PCA, kNN, SNN).  logx_out should be pinned.
// cb200_wait_copies() blocks until it has landed.
extern "C" int cb200_fetch_logcounts_async(cb200_ctx *ctx, double *logx_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->have_logx, "cb200_fetch_logcounts_async: logcounts have not been computed");
    CB_CHECK(ctx, logx_out != nullptr, "cb200_fetch_logcounts_async: NULL output");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    return cb200_copy_async(ctx, logx_out, ctx->logx.p, (size_t)ctx->nnz * sizeof(double));
}

extern "C" int cb200_wait_copies(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 1;
    return cb200_copies_wait(ctx);
}

extern "C" int cb200_lognorm_genestats(cb200_ctx *ctx, double *logx_out, double *gene_mean, double *gene_var)
{
    if (ctx == nullptr)
        return 1;
    CB_TRY(cb200_lognorm_genestats_local(ctx));
    CB_TRY(cb200_genestats_finish(ctx, gene_mean, gene_var));
    if (logx_out != nullptr)
        CB_TRY(cb200_fetch_logcounts(ctx, logx_out));
    return 0;
}
