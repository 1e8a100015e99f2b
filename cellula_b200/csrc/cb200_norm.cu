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
#include <cstdlib>
#include <cstring>

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
    ctx->sf_from_input = sf_in != nullptr;
    if (sf_in != nullptr) {
        CB_TRY(cb200_h2d(ctx, ctx->libsize.p, sf_in, (size_t)ctx->N * sizeof(double), ctx->stream));
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
    ctx->mean_ls = ctx->sf_from_input ? 0.0 : mean_ls;
    ctx->have_sf = true;
    ctx->gs_tables_ready = false;
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
// Per-gene sums WITHOUT global fp64 atomics (round 1: 3.9 G RED.ADD.F64 at cfg3 = 24 ms, and a result that
// depended on the order the reductions landed in).  Now:
//   * the sums are 64-bit FIXED-POINT integers: Y1 = rn(y 2^s1), Y2 = rn(y^2 2^s2).  Integer addition is
//     associative, so the result is bit-identical from run to run and for ANY split of the cells over CTAs or
//     GPUs (the all-reduce is an int64 sum).  s1 = 63 - ceil(log2 N) - ybits, s2 = 63 - ceil(log2 N) - 2 ybits,
//     with 2^ybits > max y (y <= log2(1 + mean library size) for library-size factors: x <= colsum);
//   * the genes are cut into slabs of <= GS_SLAB_MAX genes whose (Y1, Y2) table lives in shared memory; the rows
//     of a dgCMatrix column are sorted, so the slab's entries of a cell are ONE contiguous segment of the column
//     (k_gs_splits finds the cut points once).  A CTA owns (slab, chunk of cells) work items from a queue;
//   * one producer warp stages those segments with bulk async copies (TMA engine, mbarrier ring); 15 consumer
//     warps process one segment at a time: the genes inside a cell are distinct, so the table is updated with
//     plain 128-bit shared-memory loads / stores (no atomics; ATOMS costs 2 cycles per lane), and a named
//     barrier separates consecutive cells;
//   * when the slab changes the table is flushed with integer reductions to global memory (G / S x 2 per CTA).
// The variance is then formed WITHOUT cancellation error: N Q - S^2 in 128-bit integers (k_genestats_finish).
#include "cb200_tc.cuh"

#define GS_THREADS 512
#define GS_CONSUMERS 480
#define GS_STAGE_ENTRIES 1024   // entries per stage: several column segments are packed into one stage
#define GS_MAXSUB 8             // ... at most this many
#define GS_STAGES 5
#define GS_SLAB_MAX 9600
#define GS_CHUNK 512            // cells per work item
#define GS_ITEM_DATA 0
#define GS_ITEM_FLUSH 1
#define GS_ITEM_END 2
#define GS_TABLE 32             // counts 1 .. GS_TABLE come from the per-cell logcount table
#define GS_CAP8 (GS_STAGE_ENTRIES + 2 * GS_MAXSUB)   // staged 8-byte elements (every segment starts 16-byte aligned)
#define GS_CAP4 (GS_STAGE_ENTRIES + 8 * GS_MAXSUB)   // 4-byte elements: up to 3 before + 3 after every segment

struct GsSub {
    int64_t beg;      // first entry of the segment (element index into x / rowidx / logx)
    double rs;        // 1 / size factor of the cell
    int32_t n;        // entries
    int32_t pos8;     // position of entry `beg` inside the stage's 8-byte array
    int32_t pos4;     // ... inside the 4-byte array
    int32_t cell;     // local column of the segment
};
struct GsItem {
    int32_t kind;     // GS_ITEM_*
    int32_t nsub;
    int32_t slab_base;
    int32_t slab_genes;
    GsSub sub[GS_MAXSUB];
};

// cut points of every column at the slab boundaries: split[c * (S-1) + b] = number of entries of column c with
// row < (b+1) * slab_genes
__global__ void k_gs_splits(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, int64_t n_cols,
                            int n_slabs, int slab_genes, int32_t *__restrict__ split)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int nb = n_slabs - 1;
    if (i >= n_cols * nb)
        return;
    const int64_t c = i / nb;
    const int b = (int)(i % nb);
    const int32_t bound = (b + 1) * slab_genes;
    const int64_t base = colptr[c];
    int lo = 0, hi = (int)(colptr[c + 1] - base);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(rowidx + base + mid) < bound)
            lo = mid + 1;
        else
            hi = mid;
    }
    split[i] = lo;
}

__device__ __forceinline__ void gs_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GS_CONSUMERS) : "memory"); }

// MODE 0: staged 8-byte array = the logcounts already in HBM (cb200_lognorm_values ran first): gene sums only
// MODE 1: staged 8-byte array = counts; logcount and fixed-point images of the counts 1 .. GS_TABLE from the cell's
//         table, everything else appended to the warp's region of the overflow list (k_gs_overflow finishes them)
// MODE 2: staged 8-byte array = counts; every logcount evaluated in place (matrices that are not small counts)
// MODE 3: staged 8-byte array = counts; sums of x / sf per gene (batchelor::multiBatchNorm's per-batch averages):
//         first accumulator only, nothing written
// cell_list != NULL: the columns cell_list[0 .. n_cols) instead of 0 .. n_cols (the cells of one block / batch)
template <int MODE>
__global__ void __launch_bounds__(GS_THREADS, 1)
k_genestats_slab(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, const double *__restrict__ v8,
                 const double *__restrict__ sf, const int32_t *__restrict__ split, int64_t n_cols, int64_t G,
                 int n_slabs, int slab_genes, double scale1, double scale2, double y_limit,
                 double *__restrict__ logx, unsigned long long *__restrict__ acc1, unsigned long long *__restrict__ acc2,
                 unsigned int *__restrict__ queue, int *__restrict__ flag, const double *__restrict__ tab_y,
                 const ulonglong2 *__restrict__ tab_q, int64_t *__restrict__ ovf_entry, int32_t *__restrict__ ovf_cell,
                 int32_t *__restrict__ ovf_count, int64_t ovf_region, const int32_t *__restrict__ cell_list)
{
    extern __shared__ __align__(128) unsigned char gs_smem[];
    using namespace cb200tc;
    ulonglong2 *tab = reinterpret_cast<ulonglong2 *>(gs_smem);                       // [GS_SLAB_MAX]
    unsigned char *stage8 = gs_smem + (size_t)GS_SLAB_MAX * 16;                      // [GS_STAGES][GS_CAP8 * 8]
    unsigned char *stage4 = stage8 + (size_t)GS_STAGES * GS_CAP8 * 8;                 // [GS_STAGES][GS_CAP4 * 4]
    double *stageT = reinterpret_cast<double *>(stage4 + (size_t)GS_STAGES * GS_CAP4 * 4);   // [GS_STAGES][GS_MAXSUB][GS_TABLE]
    GsItem *items = reinterpret_cast<GsItem *>(stageT + (size_t)GS_STAGES * GS_MAXSUB * GS_TABLE);
    uint64_t *full = reinterpret_cast<uint64_t *>(items + GS_STAGES);
    uint64_t *empty = full + GS_STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int g = tid; g < GS_SLAB_MAX; g += GS_THREADS)
        tab[g] = make_ulonglong2(0ull, 0ull);
    if (tid == 0) {
        for (int s = 0; s < GS_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t n_chunks = (n_cols + GS_CHUNK - 1) / GS_CHUNK;
    const int64_t n_work = n_chunks * n_slabs;
    const int nb = n_slabs - 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        int stage = 0;
        uint32_t phase = 0;
        int cur_slab = -1, cur_genes = 0;
        bool open = false;          // a data stage is being packed
        int fill = 0, nsub = 0;     // its entries and segments so far
        int cur8 = 0, cur4 = 0;     // next free (16-byte aligned) position of the stage's 8- / 4-byte array
        auto next_stage = [&]() {
            if (++stage == GS_STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        };
        auto acquire = [&]() {            // all lanes wait (the stage's previous occupant has been consumed)
            mbar_wait(empty + stage, phase ^ 1u, flag);
        };
        auto close = [&]() {   // publishes the packed stage: lane q issues the bulk copies of segment q, one expect_tx
            if (!open)
                return;
            __syncwarp();      // the segment descriptors written by lane 0 are visible to the other lanes
            GsItem &it = items[stage];
            uint32_t mine = 0;
            int64_t a8 = 0, a4 = 0;
            uint32_t bytes8 = 0, bytes4 = 0;
            GsSub sb;
            if (lane < nsub) {
                sb = it.sub[lane];
                a8 = sb.beg & ~1LL;                                               // 16-byte aligned sources
                a4 = sb.beg & ~3LL;
                bytes8 = (uint32_t)(((sb.beg + sb.n - a8 + 1) & ~1LL) * 8);
                bytes4 = (uint32_t)(((sb.beg + sb.n - a4 + 3) & ~3LL) * 4);
                mine = bytes8 + bytes4 + (MODE == 1 ? GS_TABLE * 8 : 0);
            }
            const uint32_t total = __reduce_add_sync(0xffffffffu, mine);
            if (lane == 0) {
                it.kind = GS_ITEM_DATA;
                it.nsub = nsub;
                it.slab_base = cur_slab * slab_genes;
                it.slab_genes = cur_genes;
                mbar_expect_tx(full + stage, total);    // (a copy may land before this: the phase cannot complete without it)
            }
            if (lane < nsub) {
                bulk_g2s(stage8 + ((size_t)stage * GS_CAP8 + (size_t)(sb.pos8 - (int)(sb.beg - a8))) * 8, v8 + a8, bytes8,
                         full + stage);
                bulk_g2s(stage4 + ((size_t)stage * GS_CAP4 + (size_t)(sb.pos4 - (int)(sb.beg - a4))) * 4, rowidx + a4, bytes4,
                         full + stage);
                if (MODE == 1)   // the cell's logcount table rides along (256 bytes)
                    bulk_g2s(stageT + ((size_t)stage * GS_MAXSUB + lane) * GS_TABLE, tab_y + (int64_t)sb.cell * GS_TABLE,
                             GS_TABLE * 8, full + stage);
            }
            __syncwarp();
            next_stage();
            open = false;
        };
        auto control = [&](int kind, int slab) {
            close();
            acquire();
            if (lane == 0) {
                GsItem &it = items[stage];
                it.kind = kind;
                it.nsub = 0;
                it.slab_base = slab * slab_genes;
                it.slab_genes = (int)(((int64_t)(slab + 1) * slab_genes < G ? (int64_t)(slab + 1) * slab_genes : G) -
                                      (int64_t)slab * slab_genes);
                mbar_arrive(full + stage);
            }
            __syncwarp();
            next_stage();
        };
        for (;;) {
            unsigned int w = 0;
            if (lane == 0)
                w = atomicAdd(queue, 1u);
            w = __shfl_sync(0xffffffffu, w, 0);
            if ((int64_t)w >= n_work)
                break;
            const int slab = (int)(w / n_chunks);
            const int64_t chunk = (int64_t)w % n_chunks;
            if (slab != cur_slab) {
                if (cur_slab >= 0)
                    control(GS_ITEM_FLUSH, cur_slab);
                cur_slab = slab;
                const int64_t hi = (int64_t)(slab + 1) * slab_genes < G ? (int64_t)(slab + 1) * slab_genes : G;
                cur_genes = (int)(hi - (int64_t)slab * slab_genes);
            }
            const int64_t c0 = chunk * GS_CHUNK;
            const int64_t c1 = c0 + GS_CHUNK < n_cols ? c0 + GS_CHUNK : n_cols;
            for (int64_t cb = c0; cb < c1; cb += 32) {
                int64_t c = cb + lane;
                int64_t beg = 0, end = 0;
                double s = 1.0;
                if (c < c1) {
                    if (cell_list != nullptr)
                        c = cell_list[c];
                    const int64_t b = colptr[c], e = colptr[c + 1];
                    beg = slab == 0 ? b : b + split[c * nb + slab - 1];
                    end = slab == nb ? e : b + split[c * nb + slab];
                    s = 1.0 / sf[c];    // the consumers multiply
                }
                const int ncell = (int)(c1 - cb < 32 ? c1 - cb : 32);
                for (int j = 0; j < ncell; ++j) {
                    int64_t jb = __shfl_sync(0xffffffffu, beg, j);
                    const int64_t je = __shfl_sync(0xffffffffu, end, j);
                    const double js = __shfl_sync(0xffffffffu, s, j);
                    const int jc = __shfl_sync(0xffffffffu, (int)c, j);
                    while (jb < je) {
                        if (open && (fill >= GS_STAGE_ENTRIES || nsub >= GS_MAXSUB))
                            close();
                        if (!open) {
                            acquire();
                            open = true;
                            fill = 0;
                            nsub = 0;
                            cur8 = 0;
                            cur4 = 0;
                        }
                        const int room = GS_STAGE_ENTRIES - fill;
                        const int n = (int)(je - jb < room ? je - jb : room);
                        if (lane == 0) {
                            GsSub &sb = items[stage].sub[nsub];
                            sb.beg = jb;
                            sb.rs = js;
                            sb.n = n;
                            sb.cell = jc;
                            // the segment's 16-byte aligned image starts at an aligned position of the stage arrays
                            sb.pos8 = cur8 + (int)(jb & 1);
                            sb.pos4 = cur4 + (int)(jb & 3);
                        }
                        cur8 += (int)(((jb & 1) + n + 1) & ~1LL);
                        cur4 += (int)(((jb & 3) + n + 3) & ~3LL);
                        fill += n;
                        nsub += 1;
                        jb += n;
                    }
                }
            }
        }
        if (cur_slab >= 0)
            control(GS_ITEM_FLUSH, cur_slab);
        control(GS_ITEM_END, 0);
    } else {
        // ------------------------------------------------------------------ consumers
        const int ct = tid - 32;
        int stage = 0;
        uint32_t phase = 0;
        int bad = 0;
        // this warp's region of the overflow list (MODE 1): no atomics, the fill level is a register
        const int64_t ovf_base = ((int64_t)blockIdx.x * (GS_CONSUMERS / 32) + (warp - 1)) * ovf_region;
        int64_t ovf_n = 0;
        for (;;) {
            mbar_wait(full + stage, phase, flag);
            const GsItem &it = items[stage];
            const int kind = it.kind, nsub = it.nsub, slab_base = it.slab_base, sgenes = it.slab_genes;
            if (kind == GS_ITEM_DATA) {
                const double *base8 = reinterpret_cast<const double *>(stage8) + (size_t)stage * GS_CAP8;
                const int32_t *base4 = reinterpret_cast<const int32_t *>(stage4) + (size_t)stage * GS_CAP4;
                for (int q = 0; q < nsub; ++q) {
                    const GsSub sb = it.sub[q];
                    const double *s8 = base8 + sb.pos8;
                    const int32_t *s4 = base4 + sb.pos4;
                    const double *ty = stageT + ((size_t)stage * GS_MAXSUB + q) * GS_TABLE;
                    for (int t0 = ct - lane; t0 < sb.n; t0 += GS_CONSUMERS) {   // uniform trip count per warp (ballot below)
                        const int t = t0 + lane;
                        const bool in = t < sb.n;
                        double y = in ? s8[t] : 1.0;
                        const int g = in ? s4[t] - slab_base : 0;
                        ulonglong2 add;
                        if (MODE == 1) {
                            // counts are small integers: the logcount and its two fixed-point images come from the
                            // cell's table (k_gs_tables); anything else goes to this warp's region of the overflow list
                            const int xi = __double2int_rz(y);
                            const bool ovf = in && (xi < 1 || xi > GS_TABLE || (double)xi != y);
                            const unsigned om = __ballot_sync(0xffffffffu, ovf);
                            if (om != 0u) {
                                if (ovf) {
                                    const int64_t o = ovf_n + __popc(om & ((1u << lane) - 1u));
                                    if (o < ovf_region) {
                                        ovf_entry[ovf_base + o] = sb.beg + t;
                                        ovf_cell[ovf_base + o] = sb.cell;
                                    } else {
                                        bad |= 8;
                                    }
                                }
                                ovf_n += __popc(om);
                            }
                            if (ovf || !in)
                                continue;
                            y = ty[xi - 1];
                            st_stream(logx + sb.beg + t, y);
                            if (!(y < y_limit) || g < 0 || g >= sgenes) {
                                bad |= 2;
                                continue;
                            }
                            add.x = (unsigned long long)__double2ll_rn(y * scale1);
                            add.y = (unsigned long long)__double2ll_rn(y * y * scale2);
                        } else {
                            if (!in)
                                continue;
                            if (MODE == 2) {
                                y = cb_logcount(y, sb.rs);
                                st_stream(logx + sb.beg + t, y);
                            }
                            if (MODE == 3)
                                y = y * sb.rs;
                            if (!(y < y_limit) || g < 0 || g >= sgenes) {   // out of the fixed-point range / malformed row index
                                bad |= 2;
                                continue;
                            }
                            add.x = (unsigned long long)__double2ll_rn(y * scale1);
                            add.y = MODE == 3 ? 0ull : (unsigned long long)__double2ll_rn(y * y * scale2);
                        }
                        ulonglong2 a = tab[g];
                        a.x += add.x;
                        a.y += add.y;
                        tab[g] = a;
                    }
                    gs_consumer_sync();     // the next segment belongs to another cell and may hit the same genes
                }
            } else if (kind == GS_ITEM_FLUSH) {
                for (int g = ct; g < sgenes; g += GS_CONSUMERS) {
                    const ulonglong2 a = tab[g];
                    if (a.x | a.y) {
                        atomicAdd(acc1 + slab_base + g, a.x);
                        atomicAdd(acc2 + slab_base + g, a.y);
                        tab[g] = make_ulonglong2(0ull, 0ull);
                    }
                }
                gs_consumer_sync();
            }
            if (ct == 0)
                mbar_arrive(empty + stage);     // (the named barrier above ordered every reader of the stage before this)
            if (kind == GS_ITEM_END)
                break;
            if (++stage == GS_STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (bad)
            atomicOr(flag, bad);
        if (MODE == 1 && lane == 0)
            ovf_count[blockIdx.x * (GS_CONSUMERS / 32) + (warp - 1)] = (int32_t)(ovf_n < ovf_region ? ovf_n : ovf_region);
    }
}

// per cell: logcount of the counts 1 .. GS_TABLE (256 bytes; staged with the cell's segments by the slab kernel)
__global__ void __launch_bounds__(256)
k_gs_tables(const double *__restrict__ sf, int64_t n_cols, double *__restrict__ tab_y)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_cols * GS_TABLE)
        return;
    tab_y[i] = cb_logcount((double)((int)(i % GS_TABLE) + 1), 1.0 / sf[i / GS_TABLE]);
}

// the entries the table does not cover (counts above GS_TABLE, non-integer values), region by region of the
// overflow list: direct evaluation, integer reductions straight into the global sums
__global__ void __launch_bounds__(256)
k_gs_overflow(const int64_t *__restrict__ ovf_entry, const int32_t *__restrict__ ovf_cell, const int32_t *__restrict__ ovf_count,
              int n_regions, int64_t ovf_region, const int32_t *__restrict__ rowidx, const double *__restrict__ x,
              const double *__restrict__ sf, int64_t G, double scale1, double scale2, double y_limit,
              double *__restrict__ logx, unsigned long long *__restrict__ acc1, unsigned long long *__restrict__ acc2,
              int *__restrict__ flag)
{
    for (int r = blockIdx.x; r < n_regions; r += gridDim.x) {
        const int n = ovf_count[r];
        const int64_t base = (int64_t)r * ovf_region;
        for (int o = threadIdx.x; o < n; o += blockDim.x) {
            const int64_t e = ovf_entry[base + o];
            const double y = cb_logcount(x[e], 1.0 / sf[ovf_cell[base + o]]);
            logx[e] = y;
            const int g = rowidx[e];
            if (!(y < y_limit) || g < 0 || g >= G) {
                atomicOr(flag, 2);
                continue;
            }
            atomicAdd(acc1 + g, (unsigned long long)__double2ll_rn(y * scale1));
            atomicAdd(acc2 + g, (unsigned long long)__double2ll_rn(y * y * scale2));
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
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        const double rs = 1.0 / sf[c];
        int64_t i = b + lane;
        for (; i + 96 < e; i += 128) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = ld_stream(x + i + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                st_stream(logx + i + 32 * u, cb_logcount(v[u], rs));
        }
        for (; i < e; i += 32)
            st_stream(logx + i, cb_logcount(ld_stream(x + i), rs));
    }
}

// mean and n-1 variance from the fixed-point sums S = sum Y1 (scale 2^s1), Q = sum Y2 (scale 2^s2):
//   var = (N Q 2^(2 s1 - s2) - S^2) / (2^(2 s1) N (N - 1)) with the numerator formed exactly in 128-bit integers,
// so no cancellation error enters; implicit zeros are included by construction.
__global__ void k_genestats_finish(const unsigned long long *__restrict__ acc1, const unsigned long long *__restrict__ acc2,
                                   int64_t G, double n_total, int s1, int s2, double *__restrict__ mean,
                                   double *__restrict__ var)
{
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < G) {
        const unsigned long long S = acc1[g], Q = acc2[g];
        const unsigned __int128 nq = ((unsigned __int128)(unsigned long long)n_total * Q) << (2 * s1 - s2);
        const unsigned __int128 ss = (unsigned __int128)S * S;
        double num = 0.0;   // independent rounding of y and y^2 can leave a constant gene a hair below zero
        if (nq > ss) {
            const unsigned __int128 dlt = nq - ss;
            num = ldexp((double)(unsigned long long)(dlt >> 64), 64) + (double)(unsigned long long)dlt;
        }
        mean[g] = ldexp((double)S, -s1) / n_total;
        var[g] = n_total > 1.0 ? ldexp(num, -2 * s1) / (n_total * (n_total - 1.0)) : nan("");
    }
}

// fixed-point scales of the gene sums: 2^ybits > every logcount, sums over n_total cells stay below 2^63
static void gs_scales(cb200_ctx *ctx)
{
    int lgn = 0;
    while (((int64_t)1 << lgn) < ctx->N_total)
        ++lgn;
    int ybits = 6;   // caller-supplied size factors: no a-priori bound, the kernel checks y < 64
    if (ctx->mean_ls > 0.0) {   // library-size factors: x <= column sum, so x / sf <= mean library size
        const double ymax = log2(1.0 + ctx->mean_ls) * (1.0 + 1e-9) + 1e-9;
        ybits = 1;
        while (ldexp(1.0, ybits) <= ymax)
            ++ybits;
    }
    ctx->gs_ybits = ybits;
    ctx->gs_s1 = 63 - lgn - ybits;
    ctx->gs_s2 = 63 - lgn - 2 * ybits;
}

// One run of the slab kernel over the columns cell_list[0 .. n_items) (NULL: all N columns) into the accumulators
// acc1 / acc2.  mode as in k_genestats_slab; mode 1 is followed by the overflow kernel.
static int gs_launch(cb200_ctx *ctx, int mode, const int32_t *cell_list, int64_t n_items, unsigned long long *acc1,
                     unsigned long long *acc2, double sc1, double sc2, double ylim, const double *weights)
{
    if (n_items <= 0)
        return 0;
    const int n_slabs = (int)((ctx->G + GS_SLAB_MAX - 1) / GS_SLAB_MAX);
    const int slab_genes = (int)((ctx->G + n_slabs - 1) / n_slabs);
    if (n_slabs > 1 && !ctx->gs_split_ready) {
        CB_CUDA(ctx, ctx->gs_split.ensure((size_t)ctx->N * (n_slabs - 1)));
        const int64_t n = ctx->N * (n_slabs - 1);
        k_gs_splits<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->rowidx.p, ctx->N, n_slabs,
                                                                          slab_genes, ctx->gs_split.p);
        CB_LAUNCH(ctx);
        ctx->gs_split_ready = true;
    }
    const size_t smem = (size_t)GS_SLAB_MAX * 16 + (size_t)GS_STAGES * (GS_CAP8 * 8 + GS_CAP4 * 4 + GS_MAXSUB * GS_TABLE * 8) +
                        GS_STAGES * sizeof(GsItem) + 2 * GS_STAGES * sizeof(uint64_t);
    const int64_t n_work = ((n_items + GS_CHUNK - 1) / GS_CHUNK) * n_slabs;
    const int grid = (int)(n_work < ctx->num_sms ? n_work : ctx->num_sms);
    CB_CUDA(ctx, ctx->gs_queue.ensure(4));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->gs_queue.p, 0, 4 * sizeof(unsigned int), ctx->stream));
    const double *sfp = weights != nullptr ? weights : ctx->sf.p;
#define GS_ARGS(V8, LOGX, TABY, OE, OC, ON, REG)                                                                         \
    ctx->colptr.p, ctx->rowidx.p, V8, sfp, ctx->gs_split.p, n_items, ctx->G, n_slabs, slab_genes, sc1, sc2, ylim, LOGX,    \
        acc1, acc2, ctx->gs_queue.p, ctx->flag.p, TABY, nullptr, OE, OC, ON, REG, cell_list
    if (mode == 0) {
        CB_CUDA(ctx, cudaFuncSetAttribute(k_genestats_slab<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_genestats_slab<0><<<grid, GS_THREADS, smem, ctx->stream>>>(GS_ARGS(ctx->logx.p, nullptr, nullptr, nullptr, nullptr, nullptr, 0));
        CB_LAUNCH(ctx);
    } else if (mode == 1) {
        const int n_regions = grid * (GS_CONSUMERS / 32);
        const int64_t region = (ctx->nnz / 8 + n_regions - 1) / n_regions + 256;
        CB_CUDA(ctx, ctx->gs_ovf_entry.ensure((size_t)region * n_regions));
        CB_CUDA(ctx, ctx->gs_ovf_cell.ensure((size_t)region * n_regions));
        CB_CUDA(ctx, ctx->gs_ovf_count.ensure((size_t)n_regions));
        if (!ctx->gs_tables_ready) {   // logcounts of the counts 1 .. GS_TABLE of every cell
            CB_CUDA(ctx, ctx->gs_tab_y.ensure((size_t)ctx->N * GS_TABLE));
            const int64_t nt = ctx->N * GS_TABLE;
            k_gs_tables<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->sf.p, ctx->N, ctx->gs_tab_y.p);
            CB_LAUNCH(ctx);
            ctx->gs_tables_ready = true;
        }
        CB_CUDA(ctx, cudaFuncSetAttribute(k_genestats_slab<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_genestats_slab<1><<<grid, GS_THREADS, smem, ctx->stream>>>(
            GS_ARGS(ctx->x.p, ctx->logx.p, ctx->gs_tab_y.p, ctx->gs_ovf_entry.p, ctx->gs_ovf_cell.p, ctx->gs_ovf_count.p, region));
        CB_LAUNCH(ctx);
        k_gs_overflow<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->gs_ovf_entry.p, ctx->gs_ovf_cell.p, ctx->gs_ovf_count.p,
                                                                n_regions, region, ctx->rowidx.p, ctx->x.p, ctx->sf.p, ctx->G, sc1,
                                                                sc2, ylim, ctx->logx.p, acc1, acc2, ctx->flag.p);
        CB_LAUNCH(ctx);
    } else if (mode == 2) {
        CB_CUDA(ctx, cudaFuncSetAttribute(k_genestats_slab<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_genestats_slab<2><<<grid, GS_THREADS, smem, ctx->stream>>>(GS_ARGS(ctx->x.p, ctx->logx.p, nullptr, nullptr, nullptr, nullptr, 0));
        CB_LAUNCH(ctx);
    } else {
        CB_CUDA(ctx, cudaFuncSetAttribute(k_genestats_slab<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_genestats_slab<3><<<grid, GS_THREADS, smem, ctx->stream>>>(GS_ARGS(ctx->x.p, nullptr, nullptr, nullptr, nullptr, nullptr, 0));
        CB_LAUNCH(ctx);
    }
#undef GS_ARGS
    return 0;
}

// logcounts + fixed-point gene sums of the whole shard (n_blocks == 0) or of every block of cells separately
// (acc = [block][2][G]); table engine first, everything in place when the matrix is not made of small counts
static int gs_lognorm_sums(cb200_ctx *ctx, int n_blocks)
{
    CB_CUDA(ctx, ctx->logx.ensure((size_t)ctx->nnz + 16));
    const int nb = n_blocks > 0 ? n_blocks : 1;
    CB_CUDA(ctx, ctx->gene_acc.ensure((size_t)nb * 2 * ctx->G));
    gs_scales(ctx);
    CB_CHECK(ctx, ctx->gs_s2 >= 8, "cb200_lognorm_genestats: too many cells for the fixed-point gene sums");
    CB_TRY(cb200_need_rowidx(ctx));
    const bool from_logx = ctx->have_logx && ctx->logx_values_only;
    if (!from_logx)
        cb_stage_begin(ctx, ST_LOGNORM);   // (after cb200_lognorm_values the stage is already open)
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(ctx->gene_acc.p);
    const double sc1 = ldexp(1.0, ctx->gs_s1), sc2 = ldexp(1.0, ctx->gs_s2), ylim = ldexp(1.0, ctx->gs_ybits);
    auto run = [&](int mode) -> int {
        CB_CUDA(ctx, cudaMemsetAsync(ctx->gene_acc.p, 0, (size_t)nb * 2 * ctx->G * sizeof(unsigned long long), ctx->stream));
        if (n_blocks <= 0)
            return gs_launch(ctx, mode, nullptr, ctx->N, acc, acc + ctx->G, sc1, sc2, ylim, nullptr);
        for (int b = 0; b < n_blocks; ++b)
            CB_TRY(gs_launch(ctx, mode, ctx->blk_cells.p + ctx->blk_ptr_h[(size_t)b], ctx->blk_ptr_h[(size_t)b + 1] - ctx->blk_ptr_h[(size_t)b],
                             acc + (size_t)b * 2 * ctx->G, acc + (size_t)b * 2 * ctx->G + ctx->G, sc1, sc2, ylim, nullptr));
        return 0;
    };
    if (from_logx) {
        CB_TRY(run(0));   // the logcounts exist already (cb200_lognorm_values): only the gene sums are left
        ctx->logx_values_only = false;
    } else {
        const char *genv = getenv("CB200_LOGNORM_ENGINE");   // "direct": evaluate every logcount in place (cross-check)
        bool direct = genv != nullptr && strcmp(genv, "direct") == 0;
        if (!direct) {
            CB_TRY(run(1));
            // a matrix that is not made of small counts overflows its list: evaluate everything in place instead
            int h_flag = 0;
            CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
            if (h_flag & 8) {
                direct = true;
                CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
            }
        }
        if (direct)
            CB_TRY(run(2));
    }
    cb_stage_end(ctx, ST_LOGNORM);
    ctx->have_logx = true;
    ctx->gs_flag_pending = true;
    return 0;
}

extern "C" int cb200_lognorm_genestats_local(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded && ctx->have_sf, "cb200_lognorm_genestats: size factors have not been computed");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    return gs_lognorm_sums(ctx, 0);
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
    CB_CUDA(ctx, ctx->logx.ensure((size_t)ctx->nnz + 16));
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
    if (ctx->gs_flag_pending) {   // malformed input is reported here, where the caller synchronises anyway
        int h_flag = 0;
        CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
        ctx->gs_flag_pending = false;
        CB_CHECK(ctx, (h_flag & 4) == 0, "cb200_lognorm_genestats: internal pipeline error (mbarrier timeout)");
        CB_CHECK(ctx, (h_flag & 2) == 0,
                 "cb200_lognorm_genestats: a row index is outside [0, n_genes) / rows are not sorted within a column, or a "
                 "log-count is outside the fixed-point range (size factors far from the library sizes?)");
    }
    const unsigned long long *acc = reinterpret_cast<const unsigned long long *>(ctx->gene_acc.p);
    k_genestats_finish<<<(unsigned)((ctx->G + 255) / 256), 256, 0, ctx->stream>>>(
        acc, acc + ctx->G, ctx->G, (double)ctx->N_total, ctx->gs_s1, ctx->gs_s2, ctx->gene_mean.p, ctx->gene_var.p);
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
    CB_TRY(cb200_d2h(ctx, logx_out, ctx->logx.p, (size_t)ctx->nnz * sizeof(double), ctx->stream));
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


// ---------------------------------------------------------------------------------------
// the batch-aware branch of doNormAndReduce (R/doNormAndReduce.R:58-62, :74-76, :84-86)
//   batchelor::multiBatchNorm(sce, batch = ...)   -> per-batch averages of x / sf on the GPU (cb200_block_averages);
//                                                     the B x B median-ratio rescaling is G-sized host code
//   scran::modelGeneVar(sce, block = ...)         -> per-block mean / variance of the logcounts
//                                                     (cb200_lognorm_genestats_blocked); trend fits stay host code
// The per-block sums use the same fixed-point slab kernel, run once per block over that block's list of cells.
// ---------------------------------------------------------------------------------------
__global__ void k_blk_hist(const int32_t *__restrict__ blk, int64_t n, int B, int *__restrict__ cnt, int *__restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const int b = blk[i];
        if (b < 0 || b >= B)
            atomicOr(flag, 1);
        else
            atomicAdd(cnt + b, 1);
    }
}
// cells of every block in ascending order (one thread per block: B is small, the lists feed whole kernels)
__global__ void k_blk_fill(const int32_t *__restrict__ blk, int64_t n, int B, const int64_t *__restrict__ ptr,
                           int32_t *__restrict__ cells)
{
    const int b = blockIdx.x;
    __shared__ int base;
    if (threadIdx.x == 0)
        base = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < n; i0 += blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        const bool mine = i < n && blk[i] == b;
        // ordered compaction: ballot per warp, warp offsets through shared memory
        __shared__ int wcnt[32];
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0)
            wcnt[w] = __popc(m);
        __syncthreads();
        int off = 0, tot = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) {
            if (q < w)
                off += wcnt[q];
            tot += wcnt[q];
        }
        if (mine)
            cells[ptr[b] + base + off + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
        __syncthreads();
        if (threadIdx.x == 0)
            base += tot;
        __syncthreads();
    }
}

extern "C" int cb200_set_blocks(cb200_ctx *ctx, const int32_t *block, int n_blocks)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded, "cb200_set_blocks: no count matrix loaded");
    CB_CHECK(ctx, block != nullptr && n_blocks >= 1 && n_blocks <= 4096, "cb200_set_blocks: need 1 <= n_blocks <= 4096");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = n_blocks;
    CB_CUDA(ctx, ctx->blk_id.ensure((size_t)ctx->N));
    CB_CUDA(ctx, ctx->blk_cells.ensure((size_t)ctx->N));
    CB_CUDA(ctx, ctx->hcnt.ensure((size_t)(ctx->N > B ? ctx->N : B)));
    CB_CUDA(ctx, ctx->hptr.ensure((size_t)(ctx->N > B ? ctx->N : B) + 1));
    CB_TRY(cb200_h2d(ctx, ctx->blk_id.p, block, (size_t)ctx->N * sizeof(int32_t), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->hcnt.p, 0, (size_t)B * sizeof(int32_t), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
    k_blk_hist<<<(unsigned)((ctx->N + 255) / 256), 256, 0, ctx->stream>>>(ctx->blk_id.p, ctx->N, B, ctx->hcnt.p, ctx->flag.p);
    CB_LAUNCH(ctx);
    CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->hcnt.p, B, ctx->hptr.p));
    CB_CUDA(ctx, ctx->blk_ptr.ensure((size_t)B + 1));
    CB_CUDA(ctx, cudaMemcpyAsync(ctx->blk_ptr.p, ctx->hptr.p, (size_t)(B + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    k_blk_fill<<<B, 256, 0, ctx->stream>>>(ctx->blk_id.p, ctx->N, B, ctx->blk_ptr.p, ctx->blk_cells.p);
    CB_LAUNCH(ctx);
    ctx->blk_ptr_h.assign((size_t)B + 1, 0);
    CB_TRY(cb200_read_back(ctx, ctx->blk_ptr_h.data(), ctx->blk_ptr.p, (size_t)(B + 1) * sizeof(int64_t)));
    int h_flag = 0;
    CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
    CB_CHECK(ctx, h_flag == 0, "cb200_set_blocks: block ids must be in [0, n_blocks)");
    ctx->n_blocks = B;
    ctx->have_gram = ctx->have_scores = false;
    return 0;
}

// per block: [sum of the size factors, #cells], then max over the cells of colsum / sf (bounds x / sf)
__global__ void __launch_bounds__(256)
k_blk_part(const int32_t *__restrict__ cells, const int64_t *__restrict__ ptr, int B, const double *__restrict__ sf,
           const double *__restrict__ colsum, double *__restrict__ part)
{
    const int b = blockIdx.x;
    __shared__ double ssum[8], smax[8];
    double s = 0.0, mx = 0.0;
    for (int64_t i = ptr[b] + threadIdx.x; i < ptr[b + 1]; i += blockDim.x) {   // fixed assignment: deterministic
        const int c = cells[i];
        s += sf[c];
        mx = fmax(mx, colsum[c] / sf[c]);
    }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) {
        ssum[threadIdx.x >> 5] = s;
        smax[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, m = 0.0;
        for (int q = 0; q < 8; ++q) {
            t += ssum[q];
            m = fmax(m, smax[q]);
        }
        part[2 * b] = t;
        part[2 * b + 1] = (double)(ptr[b + 1] - ptr[b]);
        atomicMax(reinterpret_cast<unsigned long long *>(part + 2 * B), (unsigned long long)__double_as_longlong(m));
    }
}

// Phase 1 (rank-local): fills CB200_BUF_BLOCK_PART = double[2B + 1]: per block the sum of its cells' size factors and its
// cell count (a multi-GPU caller SUM-reduces these 2B numbers), then the largest colsum / sf (MAX-reduce).
extern "C" int cb200_block_prepare_local(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded && ctx->have_sf && ctx->n_blocks > 0, "cb200_block_averages: call cb200_size_factors and cb200_set_blocks first");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = ctx->n_blocks;
    CB_CUDA(ctx, ctx->blk_part.ensure((size_t)2 * B + 1));
    CB_CUDA(ctx, ctx->blk_w.ensure((size_t)ctx->N));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->blk_part.p, 0, (size_t)(2 * B + 1) * sizeof(double), ctx->stream));
    const int grid = cb_grid_for(ctx->N, 8, ctx->num_sms * 8);
    k_colsum_csc<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->x.p, ctx->N, ctx->blk_w.p);
    CB_LAUNCH(ctx);
    k_blk_part<<<B, 256, 0, ctx->stream>>>(ctx->blk_cells.p, ctx->blk_ptr.p, B, ctx->sf.p, ctx->blk_w.p, ctx->blk_part.p);
    CB_LAUNCH(ctx);
    return 0;
}

// Phase 2 (rank-local): per block and gene the fixed-point sum of x / sf over the block's local cells into
// CB200_BUF_GENE_ACC = int64[B][2][G] (second plane unused); max_ratio = the (reduced) last entry of the part buffer
extern "C" int cb200_block_averages_local(cb200_ctx *ctx, double max_ratio)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->n_blocks > 0 && ctx->blk_part.p != nullptr, "cb200_block_averages: call cb200_block_prepare_local first");
    CB_CHECK(ctx, max_ratio > 0.0 && max_ratio < 1e300, "cb200_block_averages: size factors should be positive");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = ctx->n_blocks;
    int lgn = 0, rbits = 1;
    while (((int64_t)1 << lgn) < ctx->N_total)
        ++lgn;
    while (ldexp(1.0, rbits) <= max_ratio * (1.0 + 1e-9))
        ++rbits;
    ctx->blk_s = 62 - lgn - rbits;
    CB_CHECK(ctx, ctx->blk_s >= -60, "cb200_block_averages: counts / size factors too large for the fixed-point sums");
    CB_CUDA(ctx, ctx->gene_acc.ensure((size_t)B * 2 * ctx->G));
    CB_TRY(cb200_need_rowidx(ctx));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->gene_acc.p, 0, (size_t)B * 2 * ctx->G * sizeof(unsigned long long), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(ctx->gene_acc.p);
    for (int b = 0; b < B; ++b)
        CB_TRY(gs_launch(ctx, 3, ctx->blk_cells.p + ctx->blk_ptr_h[(size_t)b], ctx->blk_ptr_h[(size_t)b + 1] - ctx->blk_ptr_h[(size_t)b],
                         acc + (size_t)b * 2 * ctx->G, acc + (size_t)b * 2 * ctx->G + ctx->G, ldexp(1.0, ctx->blk_s), 0.0,
                         ldexp(1.0, rbits), nullptr));
    ctx->gs_flag_pending = true;
    return 0;
}

__global__ void k_blk_ave_finish(const unsigned long long *__restrict__ acc, int64_t G, int B, int s, const double *__restrict__ part,
                                 double *__restrict__ ave)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)B * G) {
        const int b = (int)(i / G);
        const int64_t g = i % G;
        const double n = part[2 * b + 1], sfmean = n > 0.0 ? part[2 * b] / n : 1.0;
        // mean over the block of x / (sf / mean_block(sf)): scuttle::calculateAverage with block-centred factors
        ave[i] = n > 0.0 ? ldexp((double)acc[((size_t)b * 2) * G + g], -s) * sfmean / n : 0.0;
    }
}

// Phase 3: part_global = the reduced double[2B] (sums of the size factors, cell counts).  ave_out: row-major B x G.
extern "C" int cb200_block_averages_finish(cb200_ctx *ctx, const double *part_global, double *ave_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->n_blocks > 0 && part_global != nullptr && ave_out != nullptr, "cb200_block_averages_finish: bad arguments");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = ctx->n_blocks;
    if (ctx->gs_flag_pending) {
        int h_flag = 0;
        CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
        ctx->gs_flag_pending = false;
        CB_CHECK(ctx, (h_flag & 6) == 0, "cb200_block_averages: a row index is outside [0, n_genes) or a count / size factor is out of range");
    }
    CB_TRY(cb200_write_small(ctx, ctx->blk_part.p, part_global, (size_t)2 * B * sizeof(double)));
    CB_CUDA(ctx, ctx->stage_d.ensure((size_t)B * ctx->G));
    const int64_t n = (int64_t)B * ctx->G;
    k_blk_ave_finish<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long *>(ctx->gene_acc.p),
                                                                         ctx->G, B, ctx->blk_s, ctx->blk_part.p, ctx->stage_d.p);
    CB_LAUNCH(ctx);
    return cb200_read_back(ctx, ave_out, ctx->stage_d.p, (size_t)n * sizeof(double));
}

extern "C" int cb200_block_averages(cb200_ctx *ctx, const int32_t *block, int n_blocks, double *ave_out, double *sf_mean_out,
                                    int64_t *n_out)
{
    if (ctx == nullptr)
        return 1;
    CB_TRY(cb200_set_blocks(ctx, block, n_blocks));
    CB_TRY(cb200_block_prepare_local(ctx));
    std::vector<double> part((size_t)2 * n_blocks + 1);
    CB_TRY(cb200_read_back(ctx, part.data(), ctx->blk_part.p, part.size() * sizeof(double)));
    CB_TRY(cb200_block_averages_local(ctx, part[(size_t)2 * n_blocks]));
    CB_TRY(cb200_block_averages_finish(ctx, part.data(), ave_out));
    for (int b = 0; b < n_blocks; ++b) {
        if (sf_mean_out != nullptr)
            sf_mean_out[b] = part[(size_t)2 * b + 1] > 0.0 ? part[(size_t)2 * b] / part[(size_t)2 * b + 1] : 0.0;
        if (n_out != nullptr)
            n_out[b] = (int64_t)llround(part[(size_t)2 * b + 1]);
    }
    return 0;
}

// multiBatchNorm's rescaled size factors are used as they are (logNormCounts(center.size.factors = FALSE))
extern "C" int cb200_size_factors_given(cb200_ctx *ctx, const double *sf)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded && sf != nullptr, "cb200_size_factors_given: no count matrix loaded / NULL input");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_CUDA(ctx, ctx->sf.ensure((size_t)ctx->N));
    CB_CUDA(ctx, ctx->libsize.ensure((size_t)ctx->N));
    CB_TRY(cb200_h2d(ctx, ctx->libsize.p, sf, (size_t)ctx->N * sizeof(double), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
    k_sf_finish<<<(unsigned)((ctx->N + 255) / 256), 256, 0, ctx->stream>>>(ctx->libsize.p, ctx->N, 1.0, ctx->sf.p, ctx->flag.p);
    CB_LAUNCH(ctx);
    int h_flag = 0;
    CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
    CB_CHECK(ctx, h_flag == 0, "size factors should be positive");
    ctx->sf_from_input = true;
    ctx->mean_ls = 0.0;
    ctx->have_sf = true;
    ctx->gs_tables_ready = false;
    ctx->have_logx = ctx->have_genestats = false;
    return 0;
}

extern "C" int cb200_lognorm_genestats_blocked_local(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded && ctx->have_sf && ctx->n_blocks > 0,
             "cb200_lognorm_genestats_blocked: size factors and blocks (cb200_set_blocks) must be set first");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    return gs_lognorm_sums(ctx, ctx->n_blocks);
}

__global__ void k_genestats_blocked_finish(const unsigned long long *__restrict__ acc, int64_t G, int B, const double *__restrict__ nb,
                                           int s1, int s2, double *__restrict__ mean, double *__restrict__ var,
                                           double *__restrict__ mean_all, double *__restrict__ var_all, double n_all)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G)
        return;
    unsigned long long St = 0, Qt = 0;
    for (int b = 0; b < B; ++b) {
        const unsigned long long S = acc[((size_t)b * 2) * G + g], Q = acc[((size_t)b * 2 + 1) * G + g];
        St += S;
        Qt += Q;
        const double n = nb[b];
        double m = nan(""), v = nan("");
        if (n >= 1.0) {
            m = ldexp((double)S, -s1) / n;
            if (n >= 2.0) {
                const unsigned __int128 nq = ((unsigned __int128)(unsigned long long)n * Q) << (2 * s1 - s2);
                const unsigned __int128 ss = (unsigned __int128)S * S;
                double num = 0.0;
                if (nq > ss) {
                    const unsigned __int128 dlt = nq - ss;
                    num = ldexp((double)(unsigned long long)(dlt >> 64), 64) + (double)(unsigned long long)dlt;
                }
                v = ldexp(num, -2 * s1) / (n * (n - 1.0));
            }
        }
        mean[(size_t)b * G + g] = m;
        var[(size_t)b * G + g] = v;
    }
    // the statistics over all cells (what the PCA centres with) from the same integer sums
    const unsigned __int128 nq = ((unsigned __int128)(unsigned long long)n_all * Qt) << (2 * s1 - s2);
    const unsigned __int128 ss = (unsigned __int128)St * St;
    double num = 0.0;
    if (nq > ss) {
        const unsigned __int128 dlt = nq - ss;
        num = ldexp((double)(unsigned long long)(dlt >> 64), 64) + (double)(unsigned long long)dlt;
    }
    mean_all[g] = ldexp((double)St, -s1) / n_all;
    var_all[g] = n_all > 1.0 ? ldexp(num, -2 * s1) / (n_all * (n_all - 1.0)) : nan("");
}

// n_per_block: total cells of every block over all ranks.  gene_mean / gene_var: row-major B x G (NaN where a block has
// fewer than 1 / 2 cells, like scran).  The context also keeps the all-cell statistics for the PCA.
extern "C" int cb200_genestats_blocked_finish(cb200_ctx *ctx, const int64_t *n_per_block, double *gene_mean, double *gene_var)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->have_logx && ctx->n_blocks > 0 && n_per_block != nullptr, "cb200_genestats_blocked_finish: call cb200_lognorm_genestats_blocked_local first");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = ctx->n_blocks;
    if (ctx->gs_flag_pending) {
        int h_flag = 0;
        CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
        ctx->gs_flag_pending = false;
        CB_CHECK(ctx, (h_flag & 4) == 0, "cb200_lognorm_genestats: internal pipeline error (mbarrier timeout)");
        CB_CHECK(ctx, (h_flag & 2) == 0,
                 "cb200_lognorm_genestats: a row index is outside [0, n_genes) / rows are not sorted within a column, or a "
                 "log-count is outside the fixed-point range (size factors far from the library sizes?)");
    }
    std::vector<double> nb((size_t)B);
    double n_all = 0.0;
    for (int b = 0; b < B; ++b) {
        nb[(size_t)b] = (double)n_per_block[b];
        n_all += nb[(size_t)b];
    }
    CB_CHECK(ctx, (int64_t)n_all == ctx->N_total, "cb200_genestats_blocked_finish: the block sizes do not add up to the number of cells");
    CB_CUDA(ctx, ctx->blk_part.ensure((size_t)2 * B + 1));
    CB_TRY(cb200_write_small(ctx, ctx->blk_part.p, nb.data(), (size_t)B * sizeof(double)));
    CB_CUDA(ctx, ctx->gene_mean.ensure((size_t)ctx->G));
    CB_CUDA(ctx, ctx->gene_var.ensure((size_t)ctx->G));
    CB_CUDA(ctx, ctx->stage_d.ensure((size_t)2 * B * ctx->G));
    k_genestats_blocked_finish<<<(unsigned)((ctx->G + 255) / 256), 256, 0, ctx->stream>>>(
        reinterpret_cast<const unsigned long long *>(ctx->gene_acc.p), ctx->G, B, ctx->blk_part.p, ctx->gs_s1, ctx->gs_s2, ctx->stage_d.p,
        ctx->stage_d.p + (size_t)B * ctx->G, ctx->gene_mean.p, ctx->gene_var.p, n_all);
    CB_LAUNCH(ctx);
    if (gene_mean != nullptr)
        CB_TRY(cb200_read_back(ctx, gene_mean, ctx->stage_d.p, (size_t)B * ctx->G * sizeof(double)));
    if (gene_var != nullptr)
        CB_TRY(cb200_read_back(ctx, gene_var, ctx->stage_d.p + (size_t)B * ctx->G, (size_t)B * ctx->G * sizeof(double)));
    ctx->have_genestats = true;
    return 0;
}

extern "C" int cb200_lognorm_genestats_blocked(cb200_ctx *ctx, double *logx_out, double *gene_mean, double *gene_var)
{
    if (ctx == nullptr)
        return 1;
    CB_TRY(cb200_lognorm_genestats_blocked_local(ctx));
    std::vector<int64_t> nb((size_t)ctx->n_blocks);
    for (int b = 0; b < ctx->n_blocks; ++b)
        nb[(size_t)b] = ctx->blk_ptr_h[(size_t)b + 1] - ctx->blk_ptr_h[(size_t)b];
    CB_TRY(cb200_genestats_blocked_finish(ctx, nb.data(), gene_mean, gene_var));
    if (logx_out != nullptr)
        CB_TRY(cb200_fetch_logcounts(ctx, logx_out));
    return 0;
}
