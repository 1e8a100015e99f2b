// cb200_synth.cu -- device-side synthetic count matrices (bench / test infrastructure; not on
// the reference path).  Generator spec: SURVEY.md 8d -- negative-binomial counts with a planted
// low-rank structure: log lambda_gc = base_g + off + log depth_c + sum_f W_gf Z_fc, F = 64 latent
// factors of geometrically decaying strength (sigma_f = 0.93^f), each gene loading on 4 of them,
// Z_c = centroid[cluster(c)] + 0.3 N(0,1) over 40 clusters (continuous: no duplicate cells),
// depth_c ~ LogNormal(0, 0.35^2), counts ~ NB(mean lambda, size 10).  Every draw is a pure
// function of (seed, global cell index, gene), so any shard of cells can be generated
// independently and reproducibly, and the count and fill passes agree.
#include <cmath>
#include <vector>

#include "cb200_common.cuh"

#define SY_F 64
#define SY_NZ 4
#define SY_CLUSTERS 40
#define SY_SIZE 10.0f  // NB size parameter r = 1/dispersion

__host__ __device__ __forceinline__ uint64_t sy_mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t sy_hash(uint64_t seed, uint64_t a, uint64_t b)
{
    return sy_mix(sy_mix(seed + 0x9E3779B97F4A7C15ULL * (a + 1)) ^ (0xD1B54A32D192ED03ULL * (b + 1)));
}
__host__ __device__ __forceinline__ float sy_unif(uint64_t h)  // (0, 1)
{
    return ((float)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float sy_normal(uint64_t h)
{
    const float u1 = sy_unif(h), u2 = sy_unif(sy_mix(h + 0x632BE59BD9B4E019ULL));
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

struct SyParams {
    float *base;      // [G]
    uint8_t *wf;      // [G*SY_NZ] factor ids
    float *wv;        // [G*SY_NZ] loadings
    float *cent;      // [SY_CLUSTERS*SY_F]
};

__global__ void k_synth_params(int64_t G, uint64_t seed, float *base, uint8_t *wf, float *wv, float *cent)
{
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < G) {
        base[g] = 2.0f * sy_normal(sy_hash(seed, 0x1001, (uint64_t)g));
        for (int z = 0; z < SY_NZ; ++z) {
            const uint64_t h = sy_hash(seed, 0x2000 + z, (uint64_t)g);
            const int f = (int)(h % SY_F);
            wf[g * SY_NZ + z] = (uint8_t)f;
            wv[g * SY_NZ + z] = sy_normal(sy_mix(h)) * powf(0.93f, (float)f);
        }
    }
    if (g < SY_CLUSTERS * SY_F)
        cent[g] = sy_normal(sy_hash(seed, 0x3003, (uint64_t)g));
}

__device__ __forceinline__ int sy_draw(float lam, float u)
{
    // NB(mean lam, size r) by inversion; normal approximation for large means
    const float r = SY_SIZE;
    if (lam > 30.f) {
        const float sd = sqrtf(lam + lam * lam / r);
        const float v = lam + sd * normcdfinvf(u);
        return v < 0.f ? 0 : (int)(v + 0.5f);
    }
    const float q = lam / (r + lam);
    float p = expf(-r * log1pf(lam / r));
    float cum = p;
    int kk = 0;
    while (u > cum && kk < 4096) {
        p *= ((float)kk + r) / ((float)kk + 1.f) * q;
        cum += p;
        ++kk;
        if (p < 1e-12f)
            break;
    }
    return kk;
}

// warp per cell; FILL == 0 counts the non-zeros, FILL == 1 writes them in gene order
template <int FILL>
__global__ void __launch_bounds__(256)
k_synth_cells(int64_t G, int64_t n_cells, int64_t cell_offset, uint64_t seed, float off, SyParams P,
              int32_t *__restrict__ cnt, const int64_t *__restrict__ colptr, int32_t *__restrict__ rowidx,
              double *__restrict__ values)
{
    __shared__ float zs[8][SY_F];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + w;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t c = warp0; c < n_cells; c += nwarps) {
        const uint64_t cg = (uint64_t)(cell_offset + c);
        const int cl = (int)(sy_hash(seed, 0x4004, cg) % SY_CLUSTERS);
        for (int f = lane; f < SY_F; f += 32)
            zs[w][f] = P.cent[cl * SY_F + f] + 0.3f * sy_normal(sy_hash(seed, 0x5000 + f, cg));
        const float ldepth = 0.35f * sy_normal(sy_hash(seed, 0x6006, cg));
        __syncwarp();
        int64_t out = FILL ? colptr[c] : 0;
        int total = 0;
        for (int64_t g0 = 0; g0 < G; g0 += 32) {
            const int64_t g = g0 + lane;
            int k = 0;
            if (g < G) {
                float ll = P.base[g] + off + ldepth;
#pragma unroll
                for (int z = 0; z < SY_NZ; ++z)
                    ll = fmaf(P.wv[g * SY_NZ + z], zs[w][P.wf[g * SY_NZ + z]], ll);
                const float lam = expf(fminf(ll, 12.f));
                const float u = sy_unif(sy_hash(seed ^ 0x7007, cg, (uint64_t)g));
                k = sy_draw(lam, u);
            }
            const unsigned m = __ballot_sync(0xffffffffu, k > 0);
            if (FILL && k > 0) {
                const int64_t o = out + __popc(m & ((1u << lane) - 1u));
                rowidx[o] = (int32_t)g;
                values[o] = (double)k;
            }
            out += __popc(m);
            total += __popc(m);
        }
        if (total == 0) {  // an empty cell would make logNormCounts fail: give it one count
            if (FILL && lane == 0) {
                rowidx[colptr[c]] = (int32_t)(cg % (uint64_t)G);
                values[colptr[c]] = 1.0;
            }
            total = 1;
        }
        if (!FILL && lane == 0)
            cnt[c] = total;
        __syncwarp();
    }
}

extern "C" int cb200_synth_counts(cb200_ctx *ctx, int64_t n_genes, int64_t n_cells, double nnz_per_cell, uint64_t seed,
                                  int64_t cell_offset, int64_t n_cells_total, int64_t *nnz_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, n_genes >= 64 && n_genes < 2147483647LL && n_cells > 0, "cb200_synth_counts: bad shape");
    CB_CHECK(ctx, nnz_per_cell >= 1.0 && nnz_per_cell < (double)n_genes, "cb200_synth_counts: bad nnz_per_cell");
    CB_CHECK(ctx, cell_offset >= 0 && cell_offset + n_cells <= n_cells_total, "cb200_synth_counts: bad shard");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    SyParams P;
    DevBuf<float> base, wv, cent;
    DevBuf<uint8_t> wf;
    CB_CUDA(ctx, base.ensure((size_t)n_genes));
    CB_CUDA(ctx, wv.ensure((size_t)n_genes * SY_NZ));
    CB_CUDA(ctx, wf.ensure((size_t)n_genes * SY_NZ));
    CB_CUDA(ctx, cent.ensure((size_t)SY_CLUSTERS * SY_F));
    P.base = base.p;
    P.wf = wf.p;
    P.wv = wv.p;
    P.cent = cent.p;
    const int64_t np = n_genes > SY_CLUSTERS * SY_F ? n_genes : SY_CLUSTERS * SY_F;
    k_synth_params<<<(unsigned)((np + 255) / 256), 256, 0, ctx->stream>>>(n_genes, seed, base.p, wf.p, wv.p, cent.p);
    CB_LAUNCH(ctx);
    // calibrate the global offset so that a neutral cell has ~nnz_per_cell non-zero genes
    std::vector<float> hb((size_t)n_genes);
    CB_CUDA(ctx, cudaMemcpyAsync(hb.data(), base.p, (size_t)n_genes * sizeof(float), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double lo = -30.0, hi = 10.0;
    for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        double nz = 0.0;
        for (int64_t g = 0; g < n_genes; ++g) {
            const double lam = exp((double)hb[g] + mid);
            nz += 1.0 - exp(-(double)SY_SIZE * log1p(lam / (double)SY_SIZE));
        }
        if (nz < nnz_per_cell)
            lo = mid;
        else
            hi = mid;
    }
    float off = (float)(0.5 * (lo + hi));

    ctx->G = n_genes;
    ctx->N = n_cells;
    ctx->cell_offset = cell_offset;
    ctx->N_total = n_cells_total;
    // The neutral-cell calibration ignores the factor / depth variance: correct the offset with a
    // few count passes over a fixed probe -- the first min(N_total, 16384) cells of the *global*
    // matrix -- until the realised mean is within 1% of the request.  The probe does not depend on
    // the shard, so every rank derives the same offset and the shards tile one matrix.
    const int64_t n_probe = n_cells_total < 16384 ? n_cells_total : 16384;
    const int64_t n_buf = n_cells > n_probe ? n_cells : n_probe;
    CB_CUDA(ctx, ctx->colptr.ensure((size_t)n_buf + 1));
    CB_CUDA(ctx, ctx->hcnt.ensure((size_t)n_buf));
    int64_t nnz = 0;
    const int pgrid = cb_grid_for(n_probe, 8, ctx->num_sms * 8);
    for (int pass = 0; pass < 6; ++pass) {
        k_synth_cells<0><<<pgrid, 256, 0, ctx->stream>>>(n_genes, n_probe, 0, seed, off, P, ctx->hcnt.p, nullptr,
                                                         nullptr, nullptr);
        CB_LAUNCH(ctx);
        CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->hcnt.p, n_probe, ctx->colptr.p));
        CB_CUDA(ctx, cudaMemcpyAsync(&nnz, ctx->colptr.p + n_probe, sizeof(int64_t), cudaMemcpyDeviceToHost,
                                     ctx->stream));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const double actual = (double)nnz / (double)n_probe;
        if (fabs(actual / nnz_per_cell - 1.0) < 0.01)
            break;
        off += (float)(1.3 * log(nnz_per_cell / actual));
    }
    const int grid = cb_grid_for(n_cells, 8, ctx->num_sms * 8);
    k_synth_cells<0><<<grid, 256, 0, ctx->stream>>>(n_genes, n_cells, cell_offset, seed, off, P, ctx->hcnt.p, nullptr,
                                                    nullptr, nullptr);
    CB_LAUNCH(ctx);
    CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->hcnt.p, n_cells, ctx->colptr.p));
    CB_CUDA(ctx, cudaMemcpyAsync(&nnz, ctx->colptr.p + n_cells, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->nnz = nnz;
    CB_CUDA(ctx, ctx->rowidx.ensure((size_t)nnz));
    CB_CUDA(ctx, ctx->x.ensure((size_t)nnz));
    k_synth_cells<1><<<grid, 256, 0, ctx->stream>>>(n_genes, n_cells, cell_offset, seed, off, P, nullptr, ctx->colptr.p,
                                                    ctx->rowidx.p, ctx->x.p);
    CB_LAUNCH(ctx);
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    base.release();
    wv.release();
    wf.release();
    cent.release();
    ctx->loaded = true;
    ctx->have_sf = ctx->have_logx = ctx->have_genestats = ctx->have_gram = ctx->have_scores = false;
    ctx->have_knn = false;
    if (nnz_out != nullptr)
        *nnz_out = nnz;
    return 0;
}
