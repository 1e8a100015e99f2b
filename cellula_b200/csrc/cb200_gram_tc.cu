// cb200_gram_tc.cu -- K5b on the tensor cores: the H x H cross-product X^T X of the HVG log-counts
// with tcgen05.mma kind::i8 and an error-free digit splitting, so that the result keeps fp64-grade
// accuracy (the eigenvectors of near-degenerate PCs do not survive an fp32-accumulated product).
//
// Arithmetic.  Every log-count y (0 <= y < 32) is rounded to a 24-bit fixed-point integer
// q = round(y 2^19) = d0 + d1 2^8 + d2 2^16 with 8-bit digits.  Then
//        sum_c q_ci q_cj = sum_{p,p'} 2^{8(p+p')} sum_c d_p(c,i) d_p'(c,j)
// and each inner sum is an unsigned 8-bit GEMM with int32 accumulation, which the tensor core does
// *exactly* as long as it stays below 2^31: the five orders o = p + p' get separate TMEM
// accumulators (at most 3 digit pairs x 255^2 per cell) which are flushed into fp64 every
// GT_FLUSH_CHUNKS x 128 cells.  The only error left is the input rounding (2^-20 absolute, zero-mean
// and uncorrelated with the data), whose effect on the covariance is ~1e-9 relative.
//
// Operands.  k_gram_img densifies the sparse HVG sub-matrix into digit planes in the shared-memory
// operand layout of tcgen05 (K-major, 128-byte swizzle; K = cells): for gene tile g (128 genes) and
// cell chunk c (128 cells) three 16 KB planes.  A CTA of k_gram_tc owns a 128 x 64 output tile (I, J)
// for a slice of the chunks: per chunk one 48 KB + three 8 KB bulk copies, 9 digit products x 4
// tcgen05.mma (M = 128, N = 64, K = 32).  Warp roles: 0 = bulk-copy producer, 1 = MMA issuer,
// 2-5 = flush.  The flush keeps the sums as INTEGERS: the five exact order sums are packed into two 64-bit
// words per entry (lo = S0 + S1 2^8, hi = S2 + S3 2^8 + S4 2^16, value = (lo + hi 2^16) 2^-38) and added with
// integer reductions, so the cross-product is bit-identical from run to run and for any split of the cells over
// CTAs or GPUs (the multi-GPU all-reduce is an int64 sum); it is converted to fp64 once, in k_cov_from_gram.
#include <cstdlib>
#include <cstring>

#include "cb200_common.cuh"
#include "cb200_tc.cuh"

using namespace cb200tc;

namespace {
constexpr int GT_M = 128;            // genes per A tile (UMMA M)
constexpr int GT_N = 64;             // genes per B tile (UMMA N)
constexpr int GT_KC = 128;           // cells per chunk = bytes per operand row
constexpr int GT_PLANES = 3;
constexpr int GT_PLANE_BYTES = GT_M * GT_KC;             // 16 KB
constexpr int GT_TILE_BYTES = GT_PLANES * GT_PLANE_BYTES; // 48 KB per (gene tile, chunk)
constexpr int GT_B_BYTES = GT_N * GT_KC;                  // 8 KB: half a plane
constexpr int GT_STAGE_BYTES = GT_TILE_BYTES + GT_PLANES * GT_B_BYTES;  // 72 KB
constexpr int GT_STAGES = 3;
constexpr int GT_ORDERS = 2 * GT_PLANES - 1;              // 5 accumulators of GT_N columns
constexpr uint32_t GT_TMEM_COLS = 512;                    // 5 x 64 = 320 -> next power of two
constexpr int GT_FLUSH_CHUNKS = 80;                       // 3 * 255^2 * 128 * 80 < 2^31
constexpr int GT_THREADS = 192;
constexpr int GT_SCALE_BITS = 19;                         // q = round(y 2^19) < 2^24 for y < 32
constexpr int GT_QUAD = 2;                                // gene tiles built per CTA of k_gram_img (96 KB: two CTAs per SM overlap their zero / scatter / write-out phases)
// instruction descriptor: c = S32 (2) at [4,6), a = b = unsigned 8 bit (0), K-major, N>>3, M>>4
constexpr uint32_t GT_IDESC = (2u << 4) | ((uint32_t)(GT_N >> 3) << 17) | ((uint32_t)(GT_M >> 4) << 24);
}  // namespace

// max of the HVG values (bit pattern of a non-negative double orders like an integer)
__global__ void k_gram_maxval(const double *__restrict__ v, int64_t n, unsigned long long *__restrict__ maxbits)
{
    double m = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, v[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0)
        atomicMax(maxbits, (unsigned long long)__double_as_longlong(m));
}

// Digit planes of GT_QUAD gene tiles for one chunk of 128 cells, built in shared memory and written
// out with coalesced 16-byte stores.  bptr: block boundaries (blocks of 64 slots) of the per-cell
// entries, as produced by k_hvg_blockptr.
__global__ void __launch_bounds__(1024)
k_gram_img(const int64_t *__restrict__ hptr, const int32_t *__restrict__ hslot, const double *__restrict__ hval,
           const uint16_t *__restrict__ bptr, int64_t n_cells, int nb, int n_gt, int64_t n_chunks,
           unsigned char *__restrict__ img)
{
    extern __shared__ __align__(16) unsigned char sm[];  // GT_QUAD x 48 KB
    const int64_t kc = blockIdx.x;
    const int gt0 = blockIdx.y * GT_QUAD;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint4 *sm4 = reinterpret_cast<uint4 *>(sm);
    for (int i = tid; i < GT_QUAD * GT_TILE_BYTES / 16; i += blockDim.x)
        sm4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const int blk0 = gt0 * (GT_M / 64);
    int blk1 = (gt0 + GT_QUAD) * (GT_M / 64);
    if (blk1 > nb)
        blk1 = nb;
    const double scale = (double)(1 << GT_SCALE_BITS);
    for (int cc = w; cc < GT_KC; cc += (int)(blockDim.x >> 5)) {   // warp per cell: 32 warps keep the gathers in flight
        const int64_t c = kc * GT_KC + cc;
        if (c >= n_cells)
            break;
        const uint16_t *bp = bptr + c * (nb + 1);
        const int64_t base = hptr[c];
        const int e0 = bp[blk0], e1 = bp[blk1];
        for (int e = e0 + lane; e < e1; e += 32) {
            const int s = hslot[base + e];
            double qd = rint(hval[base + e] * scale);
            if (qd > 16777215.0)
                qd = 16777215.0;
            const uint32_t q = (uint32_t)qd;
            const int lt = s / GT_M - gt0, row = s % GT_M;
            unsigned char *t = sm + (size_t)lt * GT_TILE_BYTES + swz_offset(row, cc);
            t[0] = (unsigned char)(q & 255u);
            t[GT_PLANE_BYTES] = (unsigned char)((q >> 8) & 255u);
            t[2 * GT_PLANE_BYTES] = (unsigned char)(q >> 16);
        }
    }
    __syncthreads();
    for (int lt = 0; lt < GT_QUAD; ++lt) {
        if (gt0 + lt >= n_gt)
            break;
        uint4 *dst = reinterpret_cast<uint4 *>(img + ((size_t)(gt0 + lt) * n_chunks + kc) * GT_TILE_BYTES);
        const uint4 *src = reinterpret_cast<const uint4 *>(sm + (size_t)lt * GT_TILE_BYTES);
        for (int i = tid; i < GT_TILE_BYTES / 16; i += blockDim.x)
            dst[i] = src[i];
    }
}

// A_TMEM: the A planes of a chunk are copied shared -> tensor memory (12 tcgen05.cp per chunk, two buffers of 96 columns
// behind the 320 accumulator columns) and the 36 MMAs read A from there: with both operands in shared memory an
// M = 128, N = 64, K = 32 int8 MMA needs 6 KB of operands per 32 tensor cycles = 192 B/clk of a 128 B/clk port.
template <bool A_TMEM>
__global__ void __launch_bounds__(GT_THREADS, 1)
k_gram_tc(const unsigned char *__restrict__ img, int64_t n_chunks, int n_gt, int H,
          unsigned long long *__restrict__ gram_lo, unsigned long long *__restrict__ gram_hi, int *__restrict__ err_flag)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)GT_STAGES * GT_STAGE_BYTES);
    uint64_t *full = bars;                 // [GT_STAGES]
    uint64_t *empty = bars + GT_STAGES;    // [GT_STAGES]
    uint64_t *tfull = bars + 2 * GT_STAGES;
    uint64_t *tempty = tfull + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 1);

    // tile (I, J): 128 rows starting at I*128, 64 columns starting at J*64, J*64 >= I*128
    const int n_gn = n_gt * (GT_M / GT_N);
    int I = 0, rem = blockIdx.x;
    while (rem >= n_gn - 2 * I) {
        rem -= n_gn - 2 * I;
        ++I;
    }
    const int J = 2 * I + rem;
    const int gtB = J / 2, halfB = J & 1;
    const int64_t per = (n_chunks + gridDim.y - 1) / gridDim.y;
    const int64_t kc0 = (int64_t)blockIdx.y * per;
    const int64_t kc1 = kc0 + per < n_chunks ? kc0 + per : n_chunks;
    const int64_t nk = kc1 > kc0 ? kc1 - kc0 : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < GT_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
        tmem_alloc(tmem_slot, GT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_flush = (nk + GT_FLUSH_CHUNKS - 1) / GT_FLUSH_CHUNKS;

    if (warp == 0) {
        if (lane == 0) {
            for (int64_t t = 0; t < nk; ++t) {
                const int s = (int)(t % GT_STAGES);
                const uint32_t ph = (uint32_t)((t / GT_STAGES) & 1);
                mbar_wait(empty + s, ph ^ 1u, err_flag);
                mbar_expect_tx(full + s, GT_STAGE_BYTES);
                unsigned char *st = smem + (size_t)s * GT_STAGE_BYTES;
                const unsigned char *a = img + ((size_t)I * n_chunks + (kc0 + t)) * GT_TILE_BYTES;
                const unsigned char *b = img + ((size_t)gtB * n_chunks + (kc0 + t)) * GT_TILE_BYTES + halfB * GT_B_BYTES;
                bulk_g2s(st, a, GT_TILE_BYTES, full + s);
#pragma unroll
                for (int p = 0; p < GT_PLANES; ++p)
                    bulk_g2s(st + GT_TILE_BYTES + p * GT_B_BYTES, b + (size_t)p * GT_PLANE_BYTES, GT_B_BYTES, full + s);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            for (int64_t f = 0; f < n_flush; ++f) {
                // the flush warps must have drained the accumulators of the previous period
                mbar_wait(tempty, (uint32_t)((f & 1) ^ 1), err_flag);
                tc_fence_after();
                const int64_t t0 = f * GT_FLUSH_CHUNKS;
                const int64_t t1 = t0 + GT_FLUSH_CHUNKS < nk ? t0 + GT_FLUSH_CHUNKS : nk;
                for (int64_t t = t0; t < t1; ++t) {
                    const int s = (int)(t % GT_STAGES);
                    const uint32_t ph = (uint32_t)((t / GT_STAGES) & 1);
                    mbar_wait(full + s, ph, err_flag);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * GT_STAGE_BYTES);
                    const uint32_t sb = sa + GT_TILE_BYTES;
                    // descriptors: only the 14-bit start-address field (bytes >> 4) differs between the 36 MMAs of a
                    // chunk, so it is ONE 64-bit add per operand (the issuing thread was the limiter: two full
                    // descriptor constructions per MMA cost more than the 32 cycles the MMA keeps the tensor pipe busy)
                    const uint64_t da0 = make_desc(sa), db0 = make_desc(sb);
                    const uint32_t not_first = t != t0 ? 1u : 0u;   // first chunk of a flush period overwrites
                    const uint32_t a_tm = tmem_base + GT_ORDERS * GT_N + (uint32_t)(t & 1) * (GT_PLANES * 32);
                    if (A_TMEM) {
#pragma unroll
                        for (int p = 0; p < GT_PLANES; ++p)
#pragma unroll
                            for (int kk = 0; kk < GT_KC / 32; ++kk)
                                tmem_cp_128x256b(a_tm + (uint32_t)(p * 32 + kk * 8),
                                                 da0 + (uint64_t)((p * GT_PLANE_BYTES + kk * 32) >> 4));
                    }
#pragma unroll
                    for (int p = 0; p < GT_PLANES; ++p) {
#pragma unroll
                        for (int q = 0; q < GT_PLANES; ++q) {
#pragma unroll
                            for (int kk = 0; kk < GT_KC / 32; ++kk) {
                                // (p, q, kk) is the first product of its order o = p + q iff kk == 0 and p == max(0, o - 2)
                                const bool first_of_order = kk == 0 && p == ((p + q) > 2 ? (p + q) - 2 : 0);
                                if (A_TMEM)
                                    umma_i8_ts(tmem_base + (uint32_t)((p + q) * GT_N), a_tm + (uint32_t)(p * 32 + kk * 8),
                                               db0 + (uint64_t)((q * GT_B_BYTES + kk * 32) >> 4), GT_IDESC,
                                               first_of_order ? not_first : 1u);
                                else
                                    umma_i8(tmem_base + (uint32_t)((p + q) * GT_N),
                                            da0 + (uint64_t)((p * GT_PLANE_BYTES + kk * 32) >> 4),
                                            db0 + (uint64_t)((q * GT_B_BYTES + kk * 32) >> 4), GT_IDESC,
                                            first_of_order ? not_first : 1u);
                            }
                        }
                    }
                    umma_commit(empty + s);
                }
                umma_commit(tfull);
            }
        }
    } else {
        // ===== flush: the five exact integer sums of an entry packed into two 64-bit words.  A CTA holds one flush period
        // (the host sizes the k-slices so), so when the accumulators are complete the operand stages are idle: the rows
        // (thread = TMEM lane = output row) go through shared memory and leave as COALESCED 64-bit reductions, a warp
        // covering 32 consecutive columns of one row -- with thread = row every lane hit its own 32-byte sector and the
        // 570 M scattered reductions of a step cost ~12 ms =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        constexpr int FL_STRIDE = GT_N * 16 + 16;   // bytes per staged row; odd multiple of 16: no bank conflicts
        for (int64_t f = 0; f < n_flush; ++f) {
            mbar_wait(tfull, (uint32_t)(f & 1), err_flag);
            tc_fence_after();
            const bool staged = n_flush == 1;
#pragma unroll 1
            for (int half = 0; half < GT_N / 32; ++half) {
                unsigned long long lo[32], hi[32];
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    lo[c] = hi[c] = 0ull;
#pragma unroll
                for (int o = 0; o < GT_ORDERS; ++o) {
                    uint32_t r[32];
                    tmem_ld32_issue(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(o * GT_N + half * 32), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (o < 2)
                            lo[c] += (unsigned long long)r[c] << (8 * o);
                        else
                            hi[c] += (unsigned long long)r[c] << (8 * (o - 2));
                    }
                }
                if (staged) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        *reinterpret_cast<ulonglong2 *>(smem + (size_t)row * FL_STRIDE + (size_t)(half * 32 + c) * 16) =
                            make_ulonglong2(lo[c], hi[c]);
                } else {
                    const int grow = I * GT_M + row;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int gcol = J * GT_N + half * 32 + c;
                        if (grow < H && gcol < H && gcol >= grow && (lo[c] | hi[c]) != 0ull) {
                            atomicAdd(gram_lo + (int64_t)grow * H + gcol, lo[c]);
                            atomicAdd(gram_hi + (int64_t)grow * H + gcol, hi[c]);
                        }
                    }
                }
            }
            if (staged) {
                asm volatile("bar.sync 1, 128;" ::: "memory");   // the four flush warps
                for (int r2 = warp - 2; r2 < GT_M; r2 += 4) {
                    const int grow = I * GT_M + r2;
                    if (grow >= H)
                        break;
#pragma unroll
                    for (int cc = lane; cc < GT_N; cc += 32) {
                        const int gcol = J * GT_N + cc;
                        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(smem + (size_t)r2 * FL_STRIDE + (size_t)cc * 16);
                        if (gcol < H && gcol >= grow && (v.x | v.y) != 0ull) {
                            atomicAdd(gram_lo + (int64_t)grow * H + gcol, v.x);
                            atomicAdd(gram_hi + (int64_t)grow * H + gcol, v.y);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(tempty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        tmem_dealloc(tmem_base, GT_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// Cluster version: 8 CTAs (2 row tiles x 4 column tiles = a 256 x 256 block of the cross-product) share their
// operand loads.  k_gram_tc above is bound by L2 -> shared-memory traffic (72 KB per CTA and chunk for 9.4 M MACs);
// here CTA (r, c) fetches only a quarter of its row's A tile and half of its column's B tile (24 KB) and the TMA
// engine MULTICASTS each piece to the CTAs of the row / column that need it, so every byte leaves L2 once per
// cluster instead of once per CTA (3x less).  A stage may be overwritten when every CTA it is multicast INTO has
// finished reading it: the MMA-completion arrival is multicast to the `empty` barriers of the row- and column-mates.
// ---------------------------------------------------------------------------------------------------------
#ifndef CB200_GRAM_CL_R
#define CB200_GRAM_CL_R 1
#endif
constexpr int GT_CL_R = CB200_GRAM_CL_R, GT_CL_C = 2 * GT_CL_R, GT_CL = GT_CL_R * GT_CL_C;   // square super tile: R x 128 = C x 64
constexpr int GT_A_SLICE = GT_PLANE_BYTES / GT_CL_C;   // the part of an A plane CTA (r, c) fetches for its row-mates
constexpr int GT_B_SLICE = GT_B_BYTES / GT_CL_R;       // the part of a B half-plane it fetches for its column-mates

__global__ void __cluster_dims__(GT_CL, 1, 1) __launch_bounds__(GT_THREADS, 1)
k_gram_tc_cl(const unsigned char *__restrict__ img, int64_t n_chunks, int n_st, int H,
             unsigned long long *__restrict__ gram_lo, unsigned long long *__restrict__ gram_hi, int *__restrict__ err_flag)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)GT_STAGES * GT_STAGE_BYTES);
    uint64_t *full = bars;                 // [GT_STAGES]
    uint64_t *empty = bars + GT_STAGES;    // [GT_STAGES]
    uint64_t *tfull = bars + 2 * GT_STAGES;
    uint64_t *tempty = tfull + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 1);

    // super-tile (SI, SJ), SI <= SJ, of 256 x 256; CTA (r, c) of the cluster owns rows I = 2 SI + r (128), columns
    // J = 4 SJ + c (64)
    const uint32_t rank = cluster_ctarank();
    const int r = (int)(rank / GT_CL_C), c = (int)(rank % GT_CL_C);
    int SI = 0, rem = (int)(blockIdx.x / GT_CL);
    while (rem >= n_st - SI) {
        rem -= n_st - SI;
        ++SI;
    }
    const int SJ = SI + rem;
    const int I = GT_CL_R * SI + r, J = GT_CL_C * SJ + c;
    const int gtB = J / 2, halfB = J & 1;
    const uint16_t row_mask = (uint16_t)(((1u << GT_CL_C) - 1u) << (r * GT_CL_C));      // CTAs that share my A tile
    uint16_t col_mask = 0;                                                              // CTAs that share my B tile
#pragma unroll
    for (int rr = 0; rr < GT_CL_R; ++rr)
        col_mask |= (uint16_t)(1u << (rr * GT_CL_C + c));
    const int64_t per = (n_chunks + gridDim.y - 1) / gridDim.y;
    const int64_t kc0 = (int64_t)blockIdx.y * per;
    const int64_t kc1 = kc0 + per < n_chunks ? kc0 + per : n_chunks;
    const int64_t nk = kc1 > kc0 ? kc1 - kc0 : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < GT_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, GT_CL_C + GT_CL_R - 1);   // row-mates and column-mates (myself once)
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
        tmem_alloc(tmem_slot, GT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync();            // every CTA's barriers exist before anything is multicast into it
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_flush = (nk + GT_FLUSH_CHUNKS - 1) / GT_FLUSH_CHUNKS;

    if (warp == 0) {
        if (lane == 0) {
            for (int64_t t = 0; t < nk; ++t) {
                const int s = (int)(t % GT_STAGES);
                const uint32_t ph = (uint32_t)((t / GT_STAGES) & 1);
                mbar_wait(empty + s, ph ^ 1u, err_flag);       // everyone I multicast into has released the stage
                mbar_expect_tx(full + s, GT_STAGE_BYTES);      // the row-mates' parts of A and the column-mates' parts of B land here
                unsigned char *st = smem + (size_t)s * GT_STAGE_BYTES;
                const unsigned char *a = img + ((size_t)I * n_chunks + (kc0 + t)) * GT_TILE_BYTES + (size_t)c * GT_A_SLICE;
                const unsigned char *b = img + ((size_t)gtB * n_chunks + (kc0 + t)) * GT_TILE_BYTES + halfB * GT_B_BYTES +
                                         (size_t)r * GT_B_SLICE;
#pragma unroll
                for (int p = 0; p < GT_PLANES; ++p) {
                    bulk_g2s_multicast(st + p * GT_PLANE_BYTES + c * GT_A_SLICE, a + (size_t)p * GT_PLANE_BYTES, GT_A_SLICE, full + s,
                                       row_mask);
                    bulk_g2s_multicast(st + GT_TILE_BYTES + p * GT_B_BYTES + r * GT_B_SLICE, b + (size_t)p * GT_PLANE_BYTES, GT_B_SLICE,
                                       full + s, col_mask);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            for (int64_t f = 0; f < n_flush; ++f) {
                mbar_wait(tempty, (uint32_t)((f & 1) ^ 1), err_flag);
                tc_fence_after();
                const int64_t t0 = f * GT_FLUSH_CHUNKS;
                const int64_t t1 = t0 + GT_FLUSH_CHUNKS < nk ? t0 + GT_FLUSH_CHUNKS : nk;
                for (int64_t t = t0; t < t1; ++t) {
                    const int s = (int)(t % GT_STAGES);
                    const uint32_t ph = (uint32_t)((t / GT_STAGES) & 1);
                    mbar_wait(full + s, ph, err_flag);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * GT_STAGE_BYTES);
                    const uint32_t sb = sa + GT_TILE_BYTES;
                    // descriptors: only the 14-bit start-address field (bytes >> 4) differs between the 36 MMAs of a
                    // chunk, so it is ONE 64-bit add per operand (the issuing thread was the limiter: two full
                    // descriptor constructions per MMA cost more than the 32 cycles the MMA keeps the tensor pipe busy)
                    const uint64_t da0 = make_desc(sa), db0 = make_desc(sb);
                    const uint32_t not_first = t != t0 ? 1u : 0u;   // first chunk of a flush period overwrites
                    // A planes: shared -> tensor memory (12 copies, in issue order with the MMAs), two buffers behind the accumulators
                    const uint32_t a_tm = tmem_base + GT_ORDERS * GT_N + (uint32_t)(t & 1) * (GT_PLANES * 32);
#pragma unroll
                    for (int p = 0; p < GT_PLANES; ++p)
#pragma unroll
                        for (int kk = 0; kk < GT_KC / 32; ++kk)
                            tmem_cp_128x256b(a_tm + (uint32_t)(p * 32 + kk * 8), da0 + (uint64_t)((p * GT_PLANE_BYTES + kk * 32) >> 4));
#pragma unroll
                    for (int p = 0; p < GT_PLANES; ++p) {
#pragma unroll
                        for (int q = 0; q < GT_PLANES; ++q) {
#pragma unroll
                            for (int kk = 0; kk < GT_KC / 32; ++kk) {
                                // (p, q, kk) is the first product of its order o = p + q iff kk == 0 and p == max(0, o - 2)
                                const bool first_of_order = kk == 0 && p == ((p + q) > 2 ? (p + q) - 2 : 0);
                                umma_i8_ts(tmem_base + (uint32_t)((p + q) * GT_N), a_tm + (uint32_t)(p * 32 + kk * 8),
                                           db0 + (uint64_t)((q * GT_B_BYTES + kk * 32) >> 4), GT_IDESC,
                                           first_of_order ? not_first : 1u);
                            }
                        }
                    }
                    // the stage may be refilled by its writers (row- and column-mates, me included) once these MMAs have read it
                    umma_commit_multicast(empty + s, (uint16_t)(row_mask | col_mask));
                }
                umma_commit(tfull);
            }
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int grow = I * GT_M + row;
        for (int64_t f = 0; f < n_flush; ++f) {
            mbar_wait(tfull, (uint32_t)(f & 1), err_flag);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < GT_N / 32; ++half) {
                unsigned long long lo[32], hi[32];
#pragma unroll
                for (int cc = 0; cc < 32; ++cc)
                    lo[cc] = hi[cc] = 0ull;
#pragma unroll
                for (int o = 0; o < GT_ORDERS; ++o) {
                    uint32_t rr[32];
                    tmem_ld32_issue(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(o * GT_N + half * 32), rr);
                    tmem_ld_wait();
#pragma unroll
                    for (int cc = 0; cc < 32; ++cc) {
                        if (o < 2)
                            lo[cc] += (unsigned long long)rr[cc] << (8 * o);
                        else
                            hi[cc] += (unsigned long long)rr[cc] << (8 * (o - 2));
                    }
                }
#pragma unroll
                for (int cc = 0; cc < 32; ++cc) {
                    const int gcol = J * GT_N + half * 32 + cc;
                    if (grow < H && gcol < H && gcol >= grow && (lo[cc] | hi[cc]) != 0ull) {
                        atomicAdd(gram_lo + (int64_t)grow * H + gcol, lo[cc]);
                        atomicAdd(gram_hi + (int64_t)grow * H + gcol, hi[cc]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(tempty);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();            // nobody leaves while a cluster-mate may still signal its barriers
    if (warp == 0)
        tmem_dealloc(tmem_base, GT_TMEM_COLS);
}

// Adds the local X^T X (upper triangle) of the HVG sub-matrix held by the context into ctx->gram,
// which the caller has zeroed.  Returns 1 in *used when the tensor-core path applied (all values
// below 32), 0 when the caller has to use the fp64 kernels instead.
int cb200_gram_tc(cb200_ctx *ctx, int nb, int *used)
{
    *used = 0;
    const int H = ctx->H;
    CB_CUDA(ctx, cudaMemsetAsync(ctx->maxnrm.p, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->nnz_hvg > 0) {
        k_gram_maxval<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->hval.p, ctx->nnz_hvg, ctx->maxnrm.p);
        CB_LAUNCH(ctx);
    }
    unsigned long long bits = 0;
    CB_TRY(cb200_read_back(ctx, &bits, ctx->maxnrm.p, sizeof(bits)));
    CB_TRACE("gram_tc: maxval read");
    double mx;
    memcpy(&mx, &bits, sizeof(double));
    if (!(mx < 31.99))
        return 0;  // outside the fixed-point range: fp64 path
    const int n_gt = (H + GT_M - 1) / GT_M;
    const int n_st = (n_gt + GT_CL_R - 1) / GT_CL_R;          // 256-row super tiles (cluster kernel)
    const int64_t n_chunks = (ctx->N + GT_KC - 1) / GT_KC;
    CB_CUDA(ctx, ctx->gram_img.ensure((size_t)n_st * GT_CL_R * n_chunks * GT_TILE_BYTES));
    if (n_st * GT_CL_R > n_gt)   // the padding gene tile of the last super tile: zeros
        CB_CUDA(ctx, cudaMemsetAsync(ctx->gram_img.p + (size_t)n_gt * n_chunks * GT_TILE_BYTES, 0,
                                     (size_t)(n_st * GT_CL_R - n_gt) * n_chunks * GT_TILE_BYTES, ctx->stream));
    CB_CUDA(ctx, cudaFuncSetAttribute(k_gram_img, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GT_QUAD * GT_TILE_BYTES));
    dim3 igrid((unsigned)n_chunks, (unsigned)((n_gt + GT_QUAD - 1) / GT_QUAD));
    k_gram_img<<<igrid, 1024, GT_QUAD * GT_TILE_BYTES, ctx->stream>>>(ctx->hptr.p, ctx->hslot.p, ctx->hval.p,
                                                                      ctx->hbptr.p, ctx->N, nb, n_gt, n_chunks,
                                                                      ctx->gram_img.p);
    CB_LAUNCH(ctx);

    const int n_gn = n_gt * (GT_M / GT_N);
    int n_pairs = 0;
    for (int i = 0; i < n_gt; ++i)
        n_pairs += n_gn - 2 * i;
    // k-slices of one flush period (80 chunks): the CTAs that run at the same time then work on the SAME slice of cells for
    // different tile pairs (grid x = pair runs fastest), and the digit planes of that slice -- 16 gene tiles x 80 chunks x
    // 48 KB = 61 MB -- stay in the 126 MB L2 while all 272 pairs pass over them.  With 4 long slices every CTA streamed
    // its own part of the 7.8 GB image and the kernel read 111 GB from DRAM (14 passes); CB200_GRAM_SLICE overrides
    int64_t slice_chunks = GT_FLUSH_CHUNKS;
    if (const char *senv = getenv("CB200_GRAM_SLICE"))
        if (atoi(senv) > 0)
            slice_chunks = atoi(senv);
    int64_t ksplit = (n_chunks + slice_chunks - 1) / slice_chunks;
    const int64_t max_split = (n_chunks + 31) / 32;
    if (ksplit > max_split)
        ksplit = max_split;
    if (ksplit > 65535)
        ksplit = 65535;
    if (ksplit < 1)
        ksplit = 1;
    const size_t smem = (size_t)GT_STAGES * GT_STAGE_BYTES + (2 * GT_STAGES + 2) * 8 + 16;
    CB_CUDA(ctx, cudaFuncSetAttribute(k_gram_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CB_CUDA(ctx, cudaFuncSetAttribute(k_gram_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 3, 0, sizeof(int), ctx->stream));
    CB_TRACE("gram_tc: img launched");
    unsigned long long *glo = reinterpret_cast<unsigned long long *>(ctx->gram.p);
    const char *genv = getenv("CB200_GRAM_CLUSTER");
    const bool use_cluster = genv != nullptr && strcmp(genv, "1") == 0 && n_chunks >= 64;   // measured slower: opt-in
    if (use_cluster) {
        // clusters with multicast operand loads.  k-slices: the split that minimises (waves of clusters) x (chunks per
        // slice + the cost of one flush, ~24 chunks), so that the last wave is full
        const int n_sp = n_st * (n_st + 1) / 2;
        const int slots = ctx->num_sms / GT_CL;
        int64_t ks = 1;
        double best = 1e300;
        for (int64_t cand = 1; cand <= max_split && cand <= 64; ++cand) {
            const int64_t waves = (n_sp * cand + slots - 1) / slots;
            const double cost = (double)waves * ((double)((n_chunks + cand - 1) / cand) + 24.0);
            if (cost < best) {
                best = cost;
                ks = cand;
            }
        }
        CB_CUDA(ctx, cudaFuncSetAttribute(k_gram_tc_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 cgrid((unsigned)(n_sp * GT_CL), (unsigned)ks);
        k_gram_tc_cl<<<cgrid, GT_THREADS, smem, ctx->stream>>>(ctx->gram_img.p, n_chunks, n_st, H, glo, glo + (size_t)H * H,
                                                              ctx->flag.p + 3);
    } else {
        dim3 grid((unsigned)n_pairs, (unsigned)ksplit);
        // CB200_GRAM_A=tmem: A planes copied to tensor memory (tcgen05.cp) and read from there -- measured 1.3 ms slower than
        // both operands from shared memory once the flush was fixed (27.0 vs 25.6 ms for the stage at cfg3)
        const char *aenv = getenv("CB200_GRAM_A");
        if (!(aenv != nullptr && strcmp(aenv, "tmem") == 0))
            k_gram_tc<false><<<grid, GT_THREADS, smem, ctx->stream>>>(ctx->gram_img.p, n_chunks, n_gt, H, glo,
                                                                      glo + (size_t)H * H, ctx->flag.p + 3);
        else
            k_gram_tc<true><<<grid, GT_THREADS, smem, ctx->stream>>>(ctx->gram_img.p, n_chunks, n_gt, H, glo,
                                                                     glo + (size_t)H * H, ctx->flag.p + 3);
    }
    CB_LAUNCH(ctx);
    int h_flag = 0;
    CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p + 3, sizeof(int)));
    CB_CHECK(ctx, (h_flag & 4) == 0, "cb200_pca: tensor-core cross-product kernel timed out on a barrier");
    *used = 1;
    return 0;
}
