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
#include "cb200_synth_core.h"
#include "cb200_synth_host.h"

struct SyParams {
    float *base;      // [G]
    uint8_t *wf;      // [G*SY_NZ] factor ids
    float *wv;        // [G*SY_NZ] loadings
    float *cent;      // [SY_CLUSTERS*SY_F]
};

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
        const int cl = sy_cell_cluster(seed, cg);
        for (int f = lane; f < SY_F; f += 32)
            zs[w][f] = sy_cell_factor(seed, cg, f, P.cent, cl);
        const float ldepth = sy_cell_logdepth(seed, cg);
        __syncwarp();
        int64_t out = FILL ? colptr[c] : 0;
        int total = 0;
        for (int64_t g0 = 0; g0 < G; g0 += 32) {
            const int64_t g = g0 + lane;
            int k = 0;
            if (g < G)
                k = sy_count(seed, cg, g, P.base[g], P.wf + g * SY_NZ, P.wv + g * SY_NZ, zs[w], off, ldepth);
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
    // model parameters and the neutral-cell offset come from the same host code as the host generator
    std::vector<float> hb, hwv, hcent;
    std::vector<uint8_t> hwf;
    cb200_synth_params_host(n_genes, seed, hb, hwf, hwv, hcent);
    CB_CUDA(ctx, cudaMemcpyAsync(base.p, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(wv.p, hwv.data(), hwv.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(wf.p, hwf.data(), hwf.size(), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(cent.p, hcent.data(), hcent.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float off = cb200_synth_neutral_offset(hb, nnz_per_cell);

    ctx->G = n_genes;
    ctx->N = n_cells;
    ctx->cell_offset = cell_offset;
    ctx->N_total = n_cells_total;
    // The neutral-cell calibration ignores the factor / depth variance: correct the offset with a
    // few count passes over a fixed probe -- the first min(N_total, CB200_SYNTH_PROBE) cells of the *global*
    // matrix -- until the realised mean is within 1% of the request.  The probe does not depend on
    // the shard, so every rank derives the same offset and the shards tile one matrix.
    const int64_t n_probe = n_cells_total < CB200_SYNTH_PROBE ? n_cells_total : CB200_SYNTH_PROBE;
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
        if (cb200_synth_offset_ok(nnz_per_cell, nnz, n_probe))
            break;
        off = cb200_synth_refine_offset(off, nnz_per_cell, nnz, n_probe);
    }
    const int grid = cb_grid_for(n_cells, 8, ctx->num_sms * 8);
    k_synth_cells<0><<<grid, 256, 0, ctx->stream>>>(n_genes, n_cells, cell_offset, seed, off, P, ctx->hcnt.p, nullptr,
                                                    nullptr, nullptr);
    CB_LAUNCH(ctx);
    CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->hcnt.p, n_cells, ctx->colptr.p));
    CB_CUDA(ctx, cudaMemcpyAsync(&nnz, ctx->colptr.p + n_cells, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->nnz = nnz;
    CB_CUDA(ctx, ctx->rowidx.ensure((size_t)nnz + 16));
    CB_CUDA(ctx, ctx->x.ensure((size_t)nnz + 16));
    k_synth_cells<1><<<grid, 256, 0, ctx->stream>>>(n_genes, n_cells, cell_offset, seed, off, P, nullptr, ctx->colptr.p,
                                                    ctx->rowidx.p, ctx->x.p);
    CB_LAUNCH(ctx);
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    base.release();
    wv.release();
    wf.release();
    cent.release();
    ctx->loaded = true;
    ctx->gs_split_ready = ctx->gs_tables_ready = false;
    ctx->n_blocks = 0;
    ctx->have_sf = ctx->have_logx = ctx->have_genestats = ctx->have_gram = ctx->have_scores = false;
    ctx->have_knn = false;
    if (nnz_out != nullptr)
        *nnz_out = nnz;
    return 0;
}
