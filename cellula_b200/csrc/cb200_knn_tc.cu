// cb200_knn_tc.cu -- K6b on the 5th-generation tensor cores: tcgen05.mma distance tiles with
// TMEM accumulators, bulk-async (TMA engine) staging of the reference tiles, and the per-query
// top-KP selection fused into the TMEM epilogue.  Same contract as k_knn_tiles (cb200_knn.cu):
// it only nominates KP candidates + a threshold per query; exactness is decided by
// k_knn_rerank / k_knn_exact in fp64.
//
// Arithmetic.  Coordinates are scaled by a power of two so that max|x'| is in [256, 512) and
// split into two fp16 terms x' = hi + lo (22 significant bits).  For a query q and reference r
// the tensor core accumulates in fp32
//        T(q, r) = qh.rh + qh.rl + ql.rh - 0.5|r'|^2            (larger T = closer)
// with three K = 64 products per 128 x 128 tile; the -0.5|r'|^2 term rides in three spare K
// columns of the (qh, rh) product: q-side constants [2^12, 1, 2^-12], r-side [-a1, -a2, -a3]
// with 0.5|r'|^2 = a1 2^12 + a2 + a3 2^-12.  Selection key = -T = 0.5 d^2 - 0.5|q'|^2 (+ error).
// Error bound used by the margin test (DESIGN.md "kNN error bound"): 2^-15 (|q|^2 + max|r|^2).
//
// Shared-memory operand image (what tcgen05.mma reads through its matrix descriptors): K-major,
// 128-byte swizzle.  A point's 64 fp16 values are one 128-byte row; 8 rows form a 1024-byte
// swizzle atom; the 16-byte chunk c of row r sits at chunk (c ^ (r & 7)).  The reference image
// in HBM is stored tile by tile in exactly this layout (hi part then lo part, 32 KB per 128
// points), so one cp.async.bulk moves a whole stage.
//
// Warp roles (320 threads): warp 0 = bulk-copy producer, warp 1 = MMA issuer (one elected lane),
// warps 2-5 / 6-9 = two epilogue sets, one per TMEM accumulator stage (even / odd reference tiles);
// within a set warp w owns TMEM lane quarter w % 4 and a thread is one query row.
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>

#include "cb200_common.cuh"
#include "cb200_knn_tc.cuh"
#include "cb200_tc.cuh"

using namespace cb200tc;

namespace {

constexpr int TC_M = 128;        // queries per CTA (UMMA M)
constexpr int TC_N = 128;        // references per tile (UMMA N)
constexpr int TC_K = 64;         // padded K (fp16 elements): 128-byte rows
constexpr int TC_UMMA_K = 16;
constexpr int TC_PART_BYTES = TC_N * TC_K * 2;   // 16 KB: one fp16 part of a tile
constexpr int TC_TILE_BYTES = 2 * TC_PART_BYTES; // hi + lo
// SETS epilogue warp sets of 4 warps each, one per TMEM accumulator stage (SETS x 128 columns)
__host__ __device__ constexpr int tc_threads(int sets) { return 64 + sets * 128; }

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=b=F16 (0), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// fp64 coordinate (already scaled) -> (hi, lo) fp16 pair
__device__ __forceinline__ void split_h(double x, __half &hi, __half &lo)
{
    hi = __double2half(x);
    lo = __double2half(x - (double)__half2float(hi));
}

}  // namespace

// ---------------------------------------------------------------------------------------
// reference image: one thread per point
// ---------------------------------------------------------------------------------------
__global__ void k_knn_tc_prep(const double *__restrict__ E, int64_t lde, int64_t n, int d, double scale,
                              unsigned char *__restrict__ img, double *__restrict__ nrm2,
                              unsigned long long *__restrict__ maxbits)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n_pad = ((n + TC_N - 1) / TC_N) * TC_N;
    if (i >= n_pad)
        return;
    unsigned char *tile = img + (i / TC_N) * (int64_t)TC_TILE_BYTES;
    const int row = (int)(i % TC_N);
    __half *hi_base = reinterpret_cast<__half *>(tile);
    __half *lo_base = reinterpret_cast<__half *>(tile + TC_PART_BYTES);
    double s = 0.0, sp = 0.0;
    for (int t = 0; t < TC_K; ++t) {
        __half h = __float2half(0.f), l = __float2half(0.f);
        if (i < n && t < d) {
            const double v = E[i * lde + t];
            s += v * v;
            const double vs = v * scale;
            sp += vs * vs;
            split_h(vs, h, l);
        }
        const uint32_t off = swz_offset(row, t * 2);
        *reinterpret_cast<__half *>(reinterpret_cast<unsigned char *>(hi_base) + off) = h;
        *reinterpret_cast<__half *>(reinterpret_cast<unsigned char *>(lo_base) + off) = l;
    }
    // -0.5|r'|^2 in the three spare K columns of the hi part
    __half a1, a2, a3;
    if (i < n) {
        const double hn = 0.5 * sp;
        a1 = __double2half(hn * (1.0 / 4096.0));
        const double r1 = hn - (double)__half2float(a1) * 4096.0;
        a2 = __double2half(r1);
        const double r2 = r1 - (double)__half2float(a2);
        a3 = __double2half(r2 * 4096.0);
        nrm2[i] = s;
        atomicMax(maxbits, (unsigned long long)__double_as_longlong(s));
    } else {
        a1 = __float2half(65504.f);  // padding rows: T = -2.7e8, never selected
        a2 = __float2half(0.f);
        a3 = __float2half(0.f);
    }
    unsigned char *hb = reinterpret_cast<unsigned char *>(hi_base);
    *reinterpret_cast<__half *>(hb + swz_offset(row, (d + 0) * 2)) = __hneg(a1);
    *reinterpret_cast<__half *>(hb + swz_offset(row, (d + 1) * 2)) = __hneg(a2);
    *reinterpret_cast<__half *>(hb + swz_offset(row, (d + 2) * 2)) = __hneg(a3);
}

// max |x| over the embedding (bit pattern of a non-negative double orders like an integer)
__global__ void k_knn_tc_maxabs(const double *__restrict__ E, int64_t lde, int64_t n, int d,
                                unsigned long long *__restrict__ maxbits)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double m = 0.0;
    if (i < n)
        for (int t = 0; t < d; ++t)
            m = fmax(m, fabs(E[i * lde + t]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0)
        atomicMax(maxbits, (unsigned long long)__double_as_longlong(m));
}

// ---------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------
// Warp-cooperative insertion of (nk, ni) into one row's candidate list: KP = 32*EPL entries sorted
// ascending by (key, idx) in shared memory, element p held by lane p % 32 in register p / 32.
// Returns the new KP-th key.  ~10 + 12*EPL instructions; out of line to keep the epilogue compact.
template <int EPL>
__device__ __noinline__ float tc_insert(float *lkey, int *lidx, float nk, int ni, int lane)
{
    float k[EPL];
    int ix[EPL];
    int pos = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        k[e] = lkey[e * 32 + lane];
        ix[e] = lidx[e * 32 + lane];
        pos += __popc(__ballot_sync(0xffffffffu, pair_less<float>(k[e], ix[e], nk, ni)));
    }
#pragma unroll
    for (int e = EPL - 1; e >= 0; --e) {
        const int p = e * 32 + lane;
        float uk = __shfl_up_sync(0xffffffffu, k[e], 1);
        int ui = __shfl_up_sync(0xffffffffu, ix[e], 1);
        if (e > 0) {
            const float lk = __shfl_sync(0xffffffffu, k[e - 1], 31);
            const int li = __shfl_sync(0xffffffffu, ix[e - 1], 31);
            if (lane == 0) {
                uk = lk;
                ui = li;
            }
        }
        if (p > pos) {
            k[e] = uk;
            ix[e] = ui;
        } else if (p == pos) {
            k[e] = nk;
            ix[e] = ni;
        }
        lkey[p] = k[e];
        lidx[p] = ix[e];
    }
    const float kth = __shfl_sync(0xffffffffu, k[EPL - 1], 31);
    __syncwarp();
    return kth;
}

// Warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer, then SETS epilogue sets of 4 warps;
// set s takes the reference tiles t = s (mod SETS) from accumulator stage s.  Within a set warp w
// owns TMEM lane quarter w % 4; a thread is one query row.  Each set keeps its own sorted top-KP
// list per row, so a query ends with SETS*KP candidates and SETS thresholds.
template <int KP, int SETS, int STAGES>
__global__ void __launch_bounds__(tc_threads(SETS), 1)
k_knn_tiles_tc(const double *__restrict__ E, int64_t lde, int d, double scale, const unsigned char *__restrict__ img,
               int64_t n, int64_t q_begin, int64_t q_end, float key_unscale, int32_t *__restrict__ cand_idx,
               float *__restrict__ cand_tau, int *__restrict__ err_flag, int mode)
{
    constexpr int EPL = KP / 32;
    constexpr int TC_SETS = SETS, TC_ACC_STAGES = SETS, TC_THREADS = tc_threads(SETS);
    constexpr uint32_t TC_TMEM_COLS = SETS * TC_N;  // 256 or 512 (power of two)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *q_img = smem;                                   // 32 KB: Q hi | Q lo
    unsigned char *r_img = smem + TC_TILE_BYTES;                   // STAGES x 32 KB
    float *lkey = reinterpret_cast<float *>(r_img + (size_t)STAGES * TC_TILE_BYTES);  // [2][TC_M][KP]
    int *lidx = reinterpret_cast<int *>(lkey + (size_t)TC_SETS * TC_M * KP);          // [2][TC_M][KP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(lidx + (size_t)TC_SETS * TC_M * KP);
    uint64_t *full = bars;                       // [STAGES]
    uint64_t *empty = bars + STAGES;             // [STAGES]
    uint64_t *tfull = bars + 2 * STAGES;         // [TC_ACC_STAGES]
    uint64_t *tempty = tfull + TC_ACC_STAGES;    // [TC_ACC_STAGES]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + TC_ACC_STAGES);
    float *thr_sh = reinterpret_cast<float *>(tmem_slot + 4);  // [SETS][TC_M] thresholds published by the sets

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = q_begin + (int64_t)blockIdx.x * TC_M;
    const int64_t n_tiles = (n + TC_N - 1) / TC_N;

    // ---- one-time setup ------------------------------------------------------------------
    // query operand image (generic-proxy stores, made visible to the async proxy below)
    for (int i = tid; i < TC_M * TC_K; i += TC_THREADS) {
        const int row = i / TC_K, t = i % TC_K;
        const int64_t g = q0 + row;
        __half h = __float2half(0.f), l = __float2half(0.f);
        if (g < q_end) {
            if (t < d) {
                split_h(E[g * lde + t] * scale, h, l);
            } else if (t == d) {
                h = __float2half(4096.f);
            } else if (t == d + 1) {
                h = __float2half(1.f);
            } else if (t == d + 2) {
                h = __float2half(1.0f / 4096.f);
            }
        }
        const uint32_t off = swz_offset(row, t * 2);
        *reinterpret_cast<__half *>(q_img + off) = h;
        *reinterpret_cast<__half *>(q_img + TC_PART_BYTES + off) = l;
    }
    for (int i = tid; i < TC_SETS * TC_M * KP; i += TC_THREADS) {
        lkey[i] = FLT_MAX;
        lidx[i] = 0x7fffffff;
    }
    for (int i = tid; i < TC_SETS * TC_M; i += TC_THREADS)
        thr_sh[i] = -FLT_MAX;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < TC_ACC_STAGES; ++a) {
            mbar_init(tfull + a, 1);
            mbar_init(tempty + a, 4);  // one arrival per epilogue warp of the set
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
        tmem_alloc(tmem_slot, TC_TMEM_COLS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer: one 32 KB bulk copy per reference tile =====
        if (lane == 0) {
            for (int64_t t = 0; t < n_tiles; ++t) {
                const int s = (int)(t % STAGES);
                const uint32_t ph = (uint32_t)((t / STAGES) & 1);
                mbar_wait(empty + s, ph ^ 1u, err_flag);
                mbar_expect_tx(full + s, TC_TILE_BYTES);
                bulk_g2s(r_img + (size_t)s * TC_TILE_BYTES, img + t * (int64_t)TC_TILE_BYTES, TC_TILE_BYTES, full + s);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t qh = smem_u32(q_img), ql = smem_u32(q_img + TC_PART_BYTES);
            for (int64_t t = 0; t < n_tiles; ++t) {
                const int s = (int)(t % STAGES);
                const uint32_t ph = (uint32_t)((t / STAGES) & 1);
                const int a = (int)(t % TC_ACC_STAGES);
                const uint32_t aph = (uint32_t)((t / TC_ACC_STAGES) & 1);
                mbar_wait(tempty + a, aph ^ 1u, err_flag);
                mbar_wait(full + s, ph, err_flag);
                tc_fence_after();
                const uint32_t rh = smem_u32(r_img + (size_t)s * TC_TILE_BYTES), rl = rh + TC_PART_BYTES;
                const uint32_t dcol = tmem_base + (uint32_t)(a * TC_N);
                uint32_t acc = 0;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t qa = (p == 2) ? ql : qh;
                    const uint32_t rb = (p == 1) ? rl : rh;
#pragma unroll
                    for (int kk = 0; kk < TC_K / TC_UMMA_K; ++kk) {
                        umma_f16(dcol, make_desc(qa + kk * TC_UMMA_K * 2), make_desc(rb + kk * TC_UMMA_K * 2),
                                 TC_IDESC, acc);
                        acc = 1;
                    }
                }
                umma_commit(empty + s);   // smem stage may be refilled once these MMAs have read it
                umma_commit(tfull + a);   // accumulator complete
            }
        }
    } else {
        // ===== epilogue: thread = query row (TMEM lane), fused top-KP selection =====
        const int set = (warp - 2) >> 2;           // 0: even tiles / accumulator 0, 1: odd tiles / accumulator 1
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;
        const int64_t grow = q0 + row;
        float *set_key = lkey + (size_t)set * TC_M * KP;
        int *set_idx = lidx + (size_t)set * TC_M * KP;
        float thr = (grow < q_end) ? -FLT_MAX : FLT_MAX;  // T must exceed thr; rows past the range never select
        const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * TC_N);
        for (int64_t t = set; t < n_tiles; t += TC_SETS) {
            const uint32_t aph = (uint32_t)((t / TC_ACC_STAGES) & 1);
            mbar_wait(tfull + set, aph, err_flag);
            tc_fence_after();
            const int64_t r0 = t * TC_N;
            // A reference that does not beat the 32nd best of ANY set cannot be among the 32 best
            // overall: take the tightest threshold any set has published for this row (stale reads
            // only cost extra hits; the union of the per-set lists still holds the global top 32).
            if (thr != FLT_MAX) {
#pragma unroll
                for (int s2 = 0; s2 < TC_SETS; ++s2)
                    thr = fmaxf(thr, thr_sh[s2 * TC_M + row]);
            }
            // Scan of 32 accumulator columns held in registers.  Fast path: 3-input max tree + one
            // vote.  Slow path (some lane beats its threshold): 32 independent ballots give the hit
            // lanes of every column; each hit is inserted into its row's sorted list by the whole
            // warp (tc_insert), which also yields the row's exact new threshold.
            auto scan_chunk = [&](uint32_t(&r)[32], int64_t c0) {
                // maxima of the four 8-column groups, then of the chunk
                float g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a0 = fmaxf(fmaxf(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1])),
                                     __uint_as_float(r[8 * j + 2]));
                    float a1 = fmaxf(fmaxf(__uint_as_float(r[8 * j + 3]), __uint_as_float(r[8 * j + 4])),
                                     __uint_as_float(r[8 * j + 5]));
                    float a2 = fmaxf(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
                    g[j] = fmaxf(fmaxf(a0, a1), a2);
                }
                const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
                if (!__any_sync(0xffffffffu, m > thr) || (mode & 8))
                    return;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (!__any_sync(0xffffffffu, g[j] > thr))
                        continue;
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) {
                        const int i = 8 * j + ii;
                        unsigned c = __ballot_sync(0xffffffffu, __uint_as_float(r[i]) > thr);
                        while (c) {
                            const int owner = __ffs(c) - 1;
                            c &= c - 1;
                            const float val = __shfl_sync(0xffffffffu, __uint_as_float(r[i]), owner);
                            const float othr = __shfl_sync(0xffffffffu, thr, owner);
                            const int64_t gcol = c0 + i;
                            const int64_t og = q0 + quarter * 32 + owner;
                            if (val > othr && gcol < n && gcol != og && !(mode & 16)) {
                                const float kth = tc_insert<EPL>(set_key + (size_t)(quarter * 32 + owner) * KP,
                                                                 set_idx + (size_t)(quarter * 32 + owner) * KP, -val,
                                                                 (int)gcol, lane);
                                if (lane == owner && kth != FLT_MAX) {
                                    thr = fmaxf(thr, -kth);
                                    thr_sh[set * TC_M + row] = -kth;
                                }
                            }
                        }
                    }
                }
            };
            // one register buffer: a second one (load of chunk ch+1 in flight while ch is scanned) was
            // measured 2.4x slower -- the doubled slow-path code thrashes the instruction cache
            uint32_t ra[32];
#pragma unroll 1
            for (int ch = 0; ch < ((mode & 2) ? 2 : TC_N / 32); ++ch) {
                if (!(mode & 4)) {
                    tmem_ld32_issue(tcol + (uint32_t)(ch * 32), ra);
                    tmem_ld_wait();
                }
                if (!(mode & 1))
                    scan_chunk(ra, r0 + ch * 32);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(tempty + set);
        }
        // final: the lists are sorted; write this set's KP candidates and threshold of every row
        // (key space, unscaled: key = 0.5 d^2 - 0.5 |q|^2)
        __syncwarp();
        if (grow < q_end) {
            const float *rk = set_key + (size_t)row * KP;
            const int *ri = set_idx + (size_t)row * KP;
            int32_t *out = cand_idx + (grow - q_begin) * (int64_t)(TC_SETS * KP) + set * KP;
            for (int p = 0; p < KP; ++p) {
                const int v = ri[p];
                out[p] = v == 0x7fffffff ? -1 : v;
            }
            const float kth = rk[KP - 1];
            cand_tau[(grow - q_begin) * TC_SETS + set] = (kth == FLT_MAX) ? FLT_MAX : kth * key_unscale;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int KP, int SETS, int STAGES>
static int launch_tc(cb200_ctx *ctx, const double *E, int64_t lde, int d, double scale, int64_t n, int64_t q_begin,
                     int64_t q_end)
{
    const size_t smem = (size_t)TC_TILE_BYTES * (1 + STAGES) + (size_t)SETS * TC_M * KP * 8 +
                        (size_t)(2 * STAGES + 2 * SETS) * 8 + 16 + (size_t)SETS * TC_M * 4;
    CB_CUDA(ctx, cudaFuncSetAttribute(k_knn_tiles_tc<KP, SETS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const int64_t nq = q_end - q_begin;
    const unsigned grid = (unsigned)((nq + TC_M - 1) / TC_M);
    const float key_unscale = (float)(1.0 / (scale * scale));
    k_knn_tiles_tc<KP, SETS, STAGES><<<grid, tc_threads(SETS), smem, ctx->stream>>>(
        E, lde, d, scale, ctx->knn_img.p, n, q_begin, q_end, key_unscale, ctx->cand_idx.p, ctx->cand_tau.p,
        ctx->flag.p + 3, getenv("CB200_TC_MODE") ? atoi(getenv("CB200_TC_MODE")) : 0);
    CB_LAUNCH(ctx);
    return 0;
}

bool cb200_knn_tc_supported(int d, int k) { return d + 3 <= TC_K && k <= 52; }

// Builds the reference image and runs the tensor-core tile kernel; fills ctx->cand_idx /
// ctx->cand_tau, ctx->nrm2, ctx->maxnrm.  *kp_out = candidates kept per query, *ntau_out =
// thresholds per query (every non-candidate has key >= the smallest of them),
// *err_coef_out = coefficient of the a-priori key error bound.
int cb200_knn_tiles_tc(cb200_ctx *ctx, const double *E, int64_t lde, int64_t n, int d, int k, int64_t q_begin,
                       int64_t q_end, int *kp_out, int *ntau_out, double *err_coef_out)
{
    const int64_t nq = q_end - q_begin;
    const int64_t n_tiles = (n + TC_N - 1) / TC_N;
    const int KP = k <= 24 ? 32 : 64;
    CB_CUDA(ctx, ctx->knn_img.ensure((size_t)n_tiles * TC_TILE_BYTES));
    CB_CUDA(ctx, ctx->nrm2.ensure((size_t)n));
    const int SETS = KP == 32 ? 4 : 2;
    CB_CUDA(ctx, ctx->cand_idx.ensure((size_t)nq * SETS * KP));
    CB_CUDA(ctx, ctx->cand_tau.ensure((size_t)nq * SETS));

    cb_stage_begin(ctx, ST_KNN_PREP);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->maxnrm.p, 0, sizeof(unsigned long long), ctx->stream));
    k_knn_tc_maxabs<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(E, lde, n, d, ctx->maxnrm.p);
    CB_LAUNCH(ctx);
    unsigned long long bits = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(&bits, ctx->maxnrm.p, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double maxabs;
    memcpy(&maxabs, &bits, sizeof(double));
    CB_CHECK(ctx, maxabs == maxabs && maxabs < 1e300, "cb200_knn: embedding contains non-finite values");
    // power-of-two scale: max |x'| in [256, 512)
    double scale = 1.0;
    if (maxabs > 0.0) {
        int e;
        frexp(maxabs, &e);          // maxabs = f * 2^e, f in [0.5, 1)
        scale = ldexp(1.0, 9 - e);  // maxabs * scale in [256, 512)
    }
    CB_CUDA(ctx, cudaMemsetAsync(ctx->maxnrm.p, 0, sizeof(unsigned long long), ctx->stream));
    const int64_t n_pad = n_tiles * TC_N;
    k_knn_tc_prep<<<(unsigned)((n_pad + 127) / 128), 128, 0, ctx->stream>>>(E, lde, n, d, scale, ctx->knn_img.p,
                                                                           ctx->nrm2.p, ctx->maxnrm.p);
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_KNN_PREP);

    cb_stage_begin(ctx, ST_KNN_TILE);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 3, 0, sizeof(int), ctx->stream));
    if (KP == 32)
        CB_TRY((launch_tc<32, 4, 2>(ctx, E, lde, d, scale, n, q_begin, q_end)));
    else
        CB_TRY((launch_tc<64, 2, 2>(ctx, E, lde, d, scale, n, q_begin, q_end)));
    cb_stage_end(ctx, ST_KNN_TILE);
    *kp_out = SETS * KP;  // candidates per query: one sorted list per epilogue set
    *ntau_out = SETS;
    // split (3.1 * 2^-22) + fp32 accumulation of <= 195 products (2^-23 each, truncation assumed)
    *err_coef_out = ldexp(1.0, -15);
    return 0;
}
