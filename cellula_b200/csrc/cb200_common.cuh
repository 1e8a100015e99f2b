// cb200_common.cuh -- context, error handling and small device helpers shared by the
// kernels of libcellula_b200.so (sm_100a).  See include/cellula_b200.h for the C ABI.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cellula_b200.h"

#define CB200_NUM_SM_FALLBACK 148
#define CB200_PIN_SCRATCH (1u << 20)

// ---------------------------------------------------------------------------------------
// device buffer that grows on demand and is released with the context
// ---------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;  // capacity in elements
    cudaError_t ensure(size_t count)
    {
        if (p != nullptr && count <= n)
            return cudaSuccess;
        if (p != nullptr) {
            cudaFree(p);
            p = nullptr;
            n = 0;
        }
        size_t want = count > 0 ? count : 1;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), want * sizeof(T));
        if (e == cudaSuccess)
            n = want;
        return e;
    }
    void release()
    {
        if (p != nullptr)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

enum Cb200Stage {
    ST_H2D = 0,
    ST_COLSUM,
    ST_LOGNORM,
    ST_HVG_EXTRACT,
    ST_GRAM,
    ST_EIG,
    ST_PROJECT,
    ST_KNN_PREP,
    ST_KNN_TILE,
    ST_KNN_RERANK,
    ST_SNN_REVLIST,
    ST_SNN_EDGES,
    ST_D2H,
    ST_COUNT_
};

struct cb200_ctx {
    int device = 0;
    int num_sms = CB200_NUM_SM_FALLBACK;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    unsigned char *pin_scratch = nullptr;  // pinned staging of cb200_read_back (CB200_PIN_SCRATCH bytes)
    struct CopyWorker *copier = nullptr;  // paced background device->host copies (cb200_ctx.cu)
    struct StagePool *stage = nullptr;    // pinned staging ring + memcpy threads for pageable host memory (cb200_stage.cu)
    std::string err;

    // ---- a1: counts shard (CSC, genes x local cells)
    int64_t G = 0, N = 0, nnz = 0, cell_offset = 0, N_total = 0;
    DevBuf<int64_t> colptr;
    DevBuf<int32_t> rowidx;
    DevBuf<double> x;
    bool loaded = false;
    cudaStream_t up_stream = nullptr;   // the row indices are uploaded behind the values, on their own stream ...
    cudaEvent_t ev_x = nullptr, ev_rowidx = nullptr;
    bool rowidx_pending = false;        // ... and the first consumer waits for them (cb200_need_rowidx)

    // ---- a2/a3/a4
    DevBuf<double> libsize;   // [N]
    DevBuf<double> sf;        // [N]
    DevBuf<double> part;      // [2]  local sum, local count
    DevBuf<int> flag;         // [4]  device-side error flags
    DevBuf<double> logx;      // [nnz]
    DevBuf<double> gene_acc;  // [2G] sum y, sum y^2 as 64-bit FIXED-POINT integers (scales 2^gs_s1, 2^gs_s2): exact,
                              // order-independent sums -- the buffer a multi-GPU caller all-reduces as int64
    DevBuf<int32_t> gs_split; // [N*(S-1)] cut points of the columns at the gene-slab boundaries
    DevBuf<unsigned int> gs_queue;   // [0] work queue of the slab kernel, [2..3] 64-bit length of the overflow list
    DevBuf<double> gs_tab_y;          // [N*32] logcount of the counts 1..32 of every cell
    DevBuf<unsigned long long> gs_tab_q;  // [N*32*2] its fixed-point images
    DevBuf<int64_t> gs_ovf_entry;     // entries the table does not cover
    DevBuf<int32_t> gs_ovf_cell;
    DevBuf<int32_t> gs_ovf_count;     // entries in every warp's region of the list
    int gs_s1 = 0, gs_s2 = 0, gs_ybits = 0;
    bool gs_flag_pending = false, sf_from_input = false;
    bool gs_split_ready = false;      // cut points valid for the loaded matrix
    bool gs_tables_ready = false;     // logcount tables valid for the current size factors
    // ---- 8(f)-3: blocks (batches) of cells
    int n_blocks = 0;
    DevBuf<int32_t> blk_id;           // [N] block of every local cell
    DevBuf<int32_t> blk_cells;        // [N] local cells grouped by block
    DevBuf<int64_t> blk_ptr;          // [B+1] their offsets (device)
    std::vector<int64_t> blk_ptr_h;   // [B+1] their offsets (host)
    DevBuf<double> blk_part;          // [2B+1] per block: sum of the size factors, #cells; then max_c colsum_c / sf_c
    DevBuf<double> blk_w;             // [N] scratch: column sums / weights
    int blk_s = 0;                    // fixed-point shift of the per-block sums of x / sf
    double mean_ls = 0.0;     // mean library size when the size factors are library-size factors, else 0
    DevBuf<double> gene_mean, gene_var;  // [G]
    bool have_sf = false, have_logx = false, have_genestats = false;
    bool logx_values_only = false;      // logcounts computed by cb200_lognorm_values: gene sums still to do

    // ---- a7: PCA
    int H = 0, d = 0, blk = 0;
    DevBuf<int32_t> hvg_idx;   // [H] genes of the slots, ascending
    DevBuf<int32_t> hvg_perm;  // [H] slot of the caller's h-th HVG
    DevBuf<unsigned char> gram_img;  // int8 digit planes of the HVG matrix (cb200_gram_tc.cu)
    DevBuf<uint16_t> hbptr;    // [N*(nb+1)] block boundaries of the per-cell HVG entries
    DevBuf<int32_t> slot;      // [G] gene -> hvg slot or -1
    DevBuf<int32_t> hcnt;      // [N]
    DevBuf<int64_t> hptr;      // [N+1]
    DevBuf<int32_t> hslot;     // [nnz_hvg]
    DevBuf<double> hval;       // [nnz_hvg]
    int64_t nnz_hvg = 0;
    DevBuf<double> gram;       // upper-triangular accumulation of X^T X: fp64 [H*H] (fp64 engines) or two planes of
                               // 64-bit integers [2][H*H] (tensor-core engine, gram_fixed): lo, hi; value = (lo + hi 2^16) 2^-38
    bool gram_fixed = false;
    DevBuf<double> cov;        // [H*H]
    DevBuf<double> mu;         // [H]
    DevBuf<double> eq, ew, ey, ex, es, et, etheta, eres, emuv, egs;  // eigen-solver work (egs: split-K partials)
    DevBuf<double> rot;        // [H*d] row-major rotation (V)
    DevBuf<double> scores;     // [N*d] row-major
    DevBuf<double> varexp;     // [d]
    double total_var = 0.0;
    bool have_gram = false, have_scores = false;

    // ---- a9: kNN
    DevBuf<double> emb;        // [n_total*d] row-major fp64 (when given from host)
    DevBuf<double> stage_d;    // staging for column-major <-> row-major conversions
    DevBuf<float> emb_f;       // [n_total*DP] fp32 padded
    DevBuf<int32_t> knn_perm, knn_bkt, knn_hist, knn_qlist, knn_qpos;  // cluster ordering of the references
    DevBuf<unsigned long long> knn_cursor;
    DevBuf<int64_t> knn_qoff;
    DevBuf<float> knn_tile_lo, knn_tile_hi;  // range of the distance to the cluster centre over every reference tile
    DevBuf<double> knn_cen, knn_km_sums;     // [C*d] cluster centres (scaled units), Lloyd accumulators
    DevBuf<unsigned long long> knn_km_cnt;   // [C]
    DevBuf<float> knn_dqc;                   // [n*C] distance of every point to every cluster centre
    DevBuf<int32_t> knn_cl;                  // [n] cluster of every point
    DevBuf<float> knn_rho;                   // [n] distance to the own cluster centre
    DevBuf<unsigned int> knn_cl_rhomax;      // [C] float bits of the largest rho of the cluster
    DevBuf<int32_t> knn_cl_tile_begin;       // [C+1] first (padded) tile of every cluster
    DevBuf<int64_t> knn_pad_before;          // [C] padding positions ahead of the cluster
    DevBuf<uint16_t> knn_cl_order;           // [C*C] clusters by ascending centre distance from cluster h
    DevBuf<double> knn_eperm;  // [npos*d] the embedding in sorted order (rows of the padding positions are 0)
    DevBuf<unsigned char> knn_img;  // tensor-core operand image of the references (cb200_knn_tc.cu)
    DevBuf<float> hnorm;       // [n_total]   0.5*|r|^2 (fp32)
    DevBuf<double> nrm2;       // [n_total]   |r|^2 (fp64)
    DevBuf<unsigned long long> maxnrm;  // [1] bit pattern of max |r|^2
    DevBuf<int32_t> cand_idx;  // [nq*KP]
    DevBuf<float> cand_tau;    // [nq]
    DevBuf<float> cand_err;    // [nq] per-query bound of |approximate key - exact key| (tensor-core engine)
    DevBuf<int32_t> knn;       // [nq*k] row-major 0-based
    DevBuf<double> knn_dk;     // [nq] exact k-th candidate distance
    DevBuf<int32_t> fb_list;   // [nq] queries needing the exact scan
    DevBuf<int> fb_cnt;        // [n_fallback] hits collected per flagged query
    DevBuf<double> fb_d;       // [n_fallback * KNN_FB_CAP] their distances
    DevBuf<int32_t> fb_i;      // ... and indices
    DevBuf<int32_t> stage_i;   // staging for index layout conversion
    int64_t knn_nq = 0, knn_ntotal = 0, knn_qbegin = 0;
    int knn_k = 0;
    bool have_knn = false;
    const double *knn_E = nullptr;   // embedding the resident kNN result was computed from (row-major, stride knn_lde) ...
    int64_t knn_lde = 0;             // ... kept for cb200_knn_distances
    int knn_d = 0;

    // ---- a10: SNN
    DevBuf<int32_t> index0;    // [n_total*k] row-major 0-based (when given from host)
    DevBuf<int32_t> rcnt;      // [n_total]
    DevBuf<int64_t> rptr;      // [n_total+1]
    DevBuf<int32_t> rcur;      // [n_total]
    DevBuf<uint32_t> hosts_u;  // unsorted (cell<<8 | rank)
    DevBuf<uint32_t> hosts;    // sorted by cell within each list
    DevBuf<int32_t> ecnt;      // [nj]
    DevBuf<int64_t> eptr;      // [nj+1]
    DevBuf<int32_t> bcnt;      // [nj*(k+1)] edges of j first met in list i
    DevBuf<int32_t> efrom, eto;  // [E]
    DevBuf<double> ew8;          // [E]
    int64_t n_edges = 0;
    DevBuf<int64_t> snn_bounds;    // chunk boundaries of the emission pass (device) ...
    int64_t *snn_bounds_h = nullptr;  // ... and their pinned host copy
    const int32_t *snn_nn = nullptr;  // the counted range (between cb200_snn_count_device and cb200_snn_emit)
    int64_t snn_n = 0, snn_jbegin = 0, snn_jend = 0;
    int snn_k = 0, snn_type = 0;
    bool snn_counted = false, snn_use_hash = true;
    DevBuf<int32_t> snn_cut;          // [n*(k+1)] position of (cell, rank) inside the sorted list of its neighbour
    DevBuf<int32_t> snn_tcnt;         // [nj] upper bound of the partners of every cell
    DevBuf<int64_t> snn_toff;         // [nj+1] its scan: slices of the partner store
    DevBuf<uint32_t> snn_tmp_h, snn_tmp_rc;  // partner store written by the counting pass, read by the emission
    DevBuf<int32_t> snn_jlist;        // [(tiers + 1) * nj] cells of every hash-table tier, then the cells left to the merge kernel
    DevBuf<int> snn_nlists;           // [tiers + 1] their counts
    int snn_n_merge = 0;
    DevBuf<unsigned char> snn_flag;   // [nj] cells left to the merge kernel (too many partners for the hash table)

    // ---- 8(f)-4: per-cell QC metrics, kNN distances (cb200_qc.cu)
    DevBuf<uint32_t> qc_mask;          // [G] bit s set = gene belongs to subset s
    DevBuf<double> qc_sum;             // [(1 + S) * N] total, then per subset
    DevBuf<int32_t> qc_det;            // [(1 + S) * N]

    // ---- cluster diagnostics (cb200_diag.cu)
    DevBuf<int32_t> diag_lab, diag_i, diag_ef;
    DevBuf<double> diag_d, diag_ew;
    DevBuf<unsigned long long> diag_cnt;

    // scan scratch
    DevBuf<int64_t> scan_blk;

    // ---- timing
    cudaEvent_t ev[ST_COUNT_][2];
    bool ev_used[ST_COUNT_];
    cb200_timing tm;
};

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
int cb200_fail(cb200_ctx *ctx, const char *fmt, ...);

#define CB_CUDA(ctx, call)                                                                  \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return cb200_fail((ctx), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,   \
                              cudaGetErrorString(e__));                                     \
    } while (0)

#define CB_CHECK(ctx, cond, ...)                                                            \
    do {                                                                                    \
        if (!(cond))                                                                        \
            return cb200_fail((ctx), __VA_ARGS__);                                          \
    } while (0)

#define CB_LAUNCH(ctx)                                                                      \
    do {                                                                                    \
        (ctx)->tm.kernel_launches += 1;                                                     \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess)                                                             \
            return cb200_fail((ctx), "kernel launch failed at %s:%d: %s", __FILE__,         \
                              __LINE__, cudaGetErrorString(e__));                           \
    } while (0)

#define CB_TRY(expr)                                                                        \
    do {                                                                                    \
        int rc__ = (expr);                                                                  \
        if (rc__ != 0)                                                                      \
            return rc__;                                                                    \
    } while (0)

// CB200_TRACE=1: host wall-clock marks on stderr (debugging aid for the overlap of copies and stages)
#include <chrono>
#include <cstdlib>
static inline void cb_trace(const char *label)
{
    static const bool on = getenv("CB200_TRACE") != nullptr;
    if (!on)
        return;
    static auto t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[cb200 %10.3f ms] %s\n", ms, label);
}
#define CB_TRACE(label) cb_trace(label)

static inline void cb_stage_begin(cb200_ctx *ctx, int st)
{
    cudaEventRecord(ctx->ev[st][0], ctx->stream);
}
static inline void cb_stage_end(cb200_ctx *ctx, int st)
{
    cudaEventRecord(ctx->ev[st][1], ctx->stream);
    ctx->ev_used[st] = true;
}

// Background device->host copy of `bytes` from src (device) to dst (host, ideally pinned), ordered
// behind everything submitted to ctx->stream so far.  A worker thread feeds the copy engine a few
// small chunks at a time, so the short synchronous result reads of the following stages are never
// queued behind one multi-gigabyte transfer.  cb200_copies_wait() blocks until all jobs have landed.
int cb200_copy_async(cb200_ctx *ctx, void *dst, const void *src, size_t bytes);
int cb200_copies_wait(cb200_ctx *ctx);
void cb200_copies_shutdown(cb200_ctx *ctx);

// Copies between caller memory and HBM (cb200_stage.cu).  Page-locked caller memory: one cudaMemcpyAsync.  Pageable
// memory (R vectors): chunks through a pinned ring filled / drained by several host threads.  cb200_h2d returns when
// the source is no longer needed by the host side (pageable) or has been queued (pinned: keep it valid until the
// stream has passed it); cb200_d2h is complete on return for pageable destinations, asynchronous for pinned ones.
int cb200_h2d(cb200_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes, cudaStream_t stream);
int cb200_d2h(cb200_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes, cudaStream_t stream);
int cb200_d2h_worker(cb200_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes, cudaStream_t stream);
int cb200_upload_background(cb200_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes, cudaStream_t stream,
                            cudaEvent_t done);
int cb200_upload_join(cb200_ctx *ctx);
void cb200_stage_shutdown(cb200_ctx *ctx);
bool cb200_host_is_pinned(const void *p);
int cb200_stage_thread_count(cb200_ctx *ctx);

// Small synchronous transfers (flags, counters, a few KB of statistics or indices) that do NOT go
// through the copy engines: a kernel moves the bytes between device memory and a pinned, device-
// mapped scratch buffer over PCIe.  A copy-engine transfer issued while a background copy is in
// flight waits until that stream's queue runs dry (measured: a 4-byte read behind a 4.8 GB copy
// took 66 ms), which would serialise the stages behind the logcounts copy-out.  Both calls leave
// the stream idle.  Above CB200_PIN_SCRATCH/2 bytes they fall back to a plain cudaMemcpyAsync.
int cb200_read_back(cb200_ctx *ctx, void *dst, const void *src, size_t bytes);
int cb200_write_small(cb200_ctx *ctx, void *dev_dst, const void *src, size_t bytes);

// makes ctx->stream wait for the upload of the row indices (no-op once done)
int cb200_need_rowidx(cb200_ctx *ctx);

// exclusive scan of int32 counts into int64 offsets (out has n+1 entries, out[n] = total)
int cb200_scan_i32_to_i64(cb200_ctx *ctx, const int32_t *d_in, int64_t n, int64_t *d_out);

static inline int cb_grid_for(int64_t work_items, int per_block, int max_blocks)
{
    int64_t b = (work_items + per_block - 1) / per_block;
    if (b < 1)
        b = 1;
    if (b > max_blocks)
        b = max_blocks;
    return (int)b;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// streaming 64-/128-bit loads that do not pollute L1 (data is touched once)
__device__ __forceinline__ double ld_stream(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream2(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream_i(const int *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// logcount of the path: log2(1 + x / sf) with rs = 1 / sf (computed once per cell, __drcp-exact 1.0 / sf).
// The reference evaluates log1p(x / sf) / log(2) (scuttle::normalizeCounts); this is the same function to ~1e-15
// relative (contract: 1e-6), restated so that it costs ~40 instructions instead of the ~190 of log1p() plus two
// fp64 divisions: at cfg3 there are 2e9 of them and the pass was instruction-bound, not HBM-bound.
//   u = 1 + t rounded, c = t - (u - 1) its rounding error (exact);  u = 2^e m, m in [sqrt(1/2), sqrt(2));
//   log(m) by fdlibm's atanh series in s = f / (2 + f), f = m - 1;  log2(1 + t) = e + (log(m) + c / u) / ln 2.
// Every kernel that produces logcounts calls this one function, so all paths give identical bits.
__device__ __forceinline__ double cb_logcount(double x, double rs)
{
    const double t = x * rs;
    const double u = 1.0 + t;
    const double c = t - (u - 1.0);
    if (!(u > 0.0))
        return nan("");                              // x / sf <= -1 (not a count) or NaN
    const long long bits = __double_as_longlong(u);
    int hi = (int)(bits >> 32);
    int e = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int carry = (hi + 0x95f64) & 0x100000;   // mantissa >= sqrt(2): halve it, bump the exponent
    e += carry >> 20;
    const double m = __hiloint2double(hi | (carry ^ 0x3ff00000), (int)(unsigned int)bits);
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
    const double t2 = z * (6.666666666666735130e-01 + w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 +
                                                                                          w * 1.479819860511658591e-01)));
    const double hfsq = 0.5 * f * f;
    const double lm = f - (hfsq - s * (hfsq + (t1 + t2)));          // log(m)
    const double corr = c * (double)__frcp_rn((float)u);            // c / u: a 2^-53-sized term, fp32 reciprocal is plenty
    return (double)e + (lm + corr) * 1.4426950408889634074;
}

__device__ __forceinline__ void st_stream(double *p, double v)
{
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
#endif
