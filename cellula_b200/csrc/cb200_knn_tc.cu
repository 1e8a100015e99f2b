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
// Warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer (one elected lane), then SETS epilogue
// sets of 4 warps, one per TMEM accumulator stage; within a set warp w owns TMEM lane quarter w % 4 and
// a thread is one query row.
//
// Pruning (exact).  The references are partitioned into C clusters (a few Lloyd iterations on a
// sample; any partition is valid) and sorted by (cluster, rho = distance to the cluster centre);
// clusters are padded to whole tiles, so tile t belongs to one cluster j and covers rho in
// [lo_t, hi_t].  By the triangle inequality d(q, r) >= |d(q, c_j) - rho_r|, so a tile can hold a
// candidate of row q only if  d(q, c_j) - reach_q <= hi_t  and  d(q, c_j) + reach_q >= lo_t, where
// reach_q is the distance of the row's current KP-th candidate.  d(q, c_j) comes from a table
// computed in fp64 (k_knn_tc_assign).  The producer warp (lane = 4 query rows) visits the clusters
// by ascending centre distance from the query tile's own cluster -- the thresholds settle inside
// the first cluster -- and stages only the tiles some row can still reach; the MMA and epilogue work
// of the others is never issued.  All quantities are rounded outwards, so the pruned references
// satisfy key_exact >= threshold and the margin proof of k_knn_rerank still holds.
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <cstdio>
#include <vector>

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

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=b=F16 (0), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// fp64 coordinate (already scaled and centred) -> (hi, lo) fp16 pair: hi + lo carries 22 bits of x.
// One fp64 -> fp32 rounding (2^-24, far below the 2^-22 of the pair), then exact fp32 arithmetic.
__device__ __forceinline__ void split_h(double x, __half &hi, __half &lo)
{
    const float xf = (float)x;
    hi = __float2half_rn(xf);
    lo = __float2half_rn(xf - __half2float(hi));
}

}  // namespace

// ---------------------------------------------------------------------------------------
// ordering of the references: C clusters (a few Lloyd iterations on a strided sample), points
// sorted by (cluster, distance to the cluster centre) with a counting sort into 2^16 buckets,
// every cluster padded to whole tiles
// ---------------------------------------------------------------------------------------
#define TC_BUCKETS 65536
#define TC_MAXC 1024       // clusters at most
#define TC_KM_CHUNK 32     // centres staged per pass of k_knn_tc_assign
#define TC_KM_J 8          // centres per register block

// max |x| over the embedding
__global__ void k_knn_tc_ranges(const double *__restrict__ E, int64_t lde, int64_t n, int d,
                                unsigned long long *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double m = 0.0;
    if (i < n) {
        for (int t = 0; t < d; ++t) {
            const double v = E[i * lde + t];
            m = (v != v) ? 1.0 / 0.0 : fmax(m, fabs(v));  // NaN must not slip through fmax
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0)
        atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// initial centres: C points taken at equal strides (scaled units)
__global__ void k_knn_tc_km_init(const double *__restrict__ E, int64_t lde, int64_t n, int d, double scale, int C,
                                 double *__restrict__ cen)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < C * d) {
        const int j = idx / d, t = idx % d;
        const int64_t i = (int64_t)(((2 * (int64_t)j + 1) * n) / (2 * (int64_t)C));
        cen[idx] = E[i * lde + t] * scale;
    }
}

// Distances of the points i = m*stride (m < M) to all C centres in fp64 (scaled units).
// FINAL = false: one Lloyd step (accumulate the points of every cluster).
// FINAL = true : write the distance table dqc[i*C + j] (fp32, round to nearest), the cluster of
//                every point (ties -> lowest j), its distance rho to that centre and the largest
//                rho of every cluster.
template <bool FINAL>
__global__ void __launch_bounds__(128)
k_knn_tc_assign(const double *__restrict__ E, int64_t lde, int64_t n, int d, double scale, int64_t stride, int64_t M,
                int C, const double *__restrict__ cen, double *__restrict__ sums, unsigned long long *__restrict__ cnts,
                float *__restrict__ dqc, int32_t *__restrict__ cl, float *__restrict__ rho,
                unsigned int *__restrict__ cl_rhomax)
{
    __shared__ __align__(16) double cs[TC_K][TC_KM_CHUNK];  // centre chunk, [t][j]
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = m * stride;
    const bool valid = m < M && i < n;
    double x[TC_K];
#pragma unroll
    for (int t = 0; t < TC_K; ++t)
        x[t] = (valid && t < d) ? E[i * lde + t] * scale : 0.0;
    double best = 1.0 / 0.0;
    int bj = 0;
    for (int j0 = 0; j0 < C; j0 += TC_KM_CHUNK) {
        __syncthreads();
        for (int e = threadIdx.x; e < TC_K * TC_KM_CHUNK; e += blockDim.x) {
            const int jj = e % TC_KM_CHUNK, t = e / TC_KM_CHUNK;
            cs[t][jj] = (j0 + jj < C && t < d) ? cen[(size_t)(j0 + jj) * d + t] : 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int jb = 0; jb < TC_KM_CHUNK && j0 + jb < C; jb += TC_KM_J) {
            double acc[TC_KM_J];
#pragma unroll
            for (int u = 0; u < TC_KM_J; ++u)
                acc[u] = 0.0;
#pragma unroll
            for (int t = 0; t < TC_K; ++t) {
                if (t < d) {
#pragma unroll
                    for (int u = 0; u < TC_KM_J; u += 2) {
                        const double2 c2 = *reinterpret_cast<const double2 *>(&cs[t][jb + u]);
                        const double d0 = x[t] - c2.x, d1 = x[t] - c2.y;
                        acc[u] = fma(d0, d0, acc[u]);
                        acc[u + 1] = fma(d1, d1, acc[u + 1]);
                    }
                }
            }
            float f[TC_KM_J];
#pragma unroll
            for (int u = 0; u < TC_KM_J; ++u) {
                const double dist = sqrt(acc[u]);
                f[u] = (float)dist;
                if (j0 + jb + u < C && dist < best) {
                    best = dist;
                    bj = j0 + jb + u;
                }
            }
            if (FINAL && valid) {
                float *out = dqc + (size_t)i * C + j0 + jb;
                if (j0 + jb + TC_KM_J <= C && (C & 3) == 0) {
                    *reinterpret_cast<float4 *>(out) = make_float4(f[0], f[1], f[2], f[3]);
                    *reinterpret_cast<float4 *>(out + 4) = make_float4(f[4], f[5], f[6], f[7]);
                } else {
#pragma unroll
                    for (int u = 0; u < TC_KM_J; ++u)
                        if (j0 + jb + u < C)
                            out[u] = f[u];
                }
            }
        }
    }
    if (!valid)
        return;
    if (FINAL) {
        cl[i] = bj;
        const float r = (float)best;
        rho[i] = r;
        atomicMax(cl_rhomax + bj, __float_as_uint(r));  // r >= 0: the bit patterns are ordered
    } else {
        atomicAdd(cnts + bj, 1ULL);
#pragma unroll
        for (int t = 0; t < TC_K; ++t)
            if (t < d)
                atomicAdd(sums + (size_t)bj * d + t, x[t]);
    }
}

// centres <- mean of the assigned sample points (an empty cluster keeps its centre)
__global__ void k_knn_tc_km_update(int C, int d, const double *__restrict__ sums,
                                   const unsigned long long *__restrict__ cnts, double *__restrict__ cen)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < C * d) {
        const unsigned long long c = cnts[idx / d];
        if (c > 0)
            cen[idx] = sums[idx] / (double)c;
    }
}

// bucket = cluster * B + bin of rho within [0, rhomax of the cluster]
__global__ void k_knn_tc_bucket(int64_t n, int B, const int32_t *__restrict__ cl, const float *__restrict__ rho,
                                const unsigned int *__restrict__ cl_rhomax, int32_t *__restrict__ bkt,
                                int32_t *__restrict__ hist)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const int c = cl[i];
        const float rm = __uint_as_float(cl_rhomax[c]);
        int b = rm > 0.f ? (int)(rho[i] / rm * (float)B) : 0;
        b = b < 0 ? 0 : (b >= B ? B - 1 : b);
        const int bk = c * B + b;
        bkt[i] = bk;
        atomicAdd(hist + bk, 1);
    }
}

// every cluster starts on a tile boundary: first tile of every cluster and the number of padding
// positions ahead of it (one thread; C <= 256)
__global__ void k_knn_tc_layout(int C, int B, const unsigned long long *__restrict__ cursor,
                                int32_t *__restrict__ cl_tile_begin, int64_t *__restrict__ pad_before)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t tiles = 0;
        for (int c = 0; c < C; ++c) {
            const int64_t p0 = (int64_t)cursor[(size_t)c * B], p1 = (int64_t)cursor[(size_t)(c + 1) * B];
            cl_tile_begin[c] = (int32_t)tiles;
            pad_before[c] = tiles * TC_N - p0;
            tiles += (p1 - p0 + TC_N - 1) / TC_N;
        }
        cl_tile_begin[C] = (int32_t)tiles;
    }
}

// perm[sorted position] = original index, -1 at the padding positions (order inside a bucket is
// arbitrary: tile bounds are taken from the actual values, and every tie-break of the final result
// uses original indices)
__global__ void k_knn_tc_scatter(const int32_t *__restrict__ bkt, int64_t n, int B,
                                 unsigned long long *__restrict__ cursor, const int64_t *__restrict__ pad_before,
                                 int32_t *__restrict__ perm)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const int bk = bkt[i];
        const unsigned long long pos = atomicAdd(cursor + bk, 1ULL);
        perm[(int64_t)pos + pad_before[bk / B]] = (int32_t)i;
    }
}

// clusters by ascending centre distance from cluster h (block h)
__global__ void __launch_bounds__(TC_MAXC)
k_knn_tc_corder(int C, int d, const double *__restrict__ cen, uint16_t *__restrict__ order)
{
    __shared__ double dist[TC_MAXC];
    const int h = blockIdx.x, j = threadIdx.x;
    double s = 0.0;
    if (j < C) {
        for (int t = 0; t < d; ++t) {
            const double df = cen[(size_t)h * d + t] - cen[(size_t)j * d + t];
            s += df * df;
        }
        dist[j] = s;
    }
    __syncthreads();
    if (j < C) {
        int rank = 0;
        for (int o = 0; o < C; ++o)
            rank += (dist[o] < s || (dist[o] == s && o < j)) ? 1 : 0;
        order[(size_t)h * C + rank] = (uint16_t)j;
    }
}

// queries of [q_begin, q_end) in sorted order: flag, (scan), scatter
__global__ void k_knn_tc_qflag(const int32_t *__restrict__ perm, int64_t npos, int64_t q_begin, int64_t q_end,
                               int32_t *__restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < npos)
        flag[i] = (perm[i] >= q_begin && perm[i] < q_end) ? 1 : 0;
}
__global__ void k_knn_tc_qscatter(const int32_t *__restrict__ perm, const int32_t *__restrict__ flag,
                                  const int64_t *__restrict__ off, int64_t npos, int32_t *__restrict__ qlist,
                                  int32_t *__restrict__ qpos)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < npos && flag[i]) {
        qlist[off[i]] = perm[i];
        qpos[off[i]] = (int32_t)i;
    }
}

// ---------------------------------------------------------------------------------------
// reference image in sorted order: one block = one tile of 128 positions, one thread per point
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_knn_tc_prep(const double *__restrict__ E, int64_t lde, int d, double scale, const int32_t *__restrict__ perm,
              const float *__restrict__ rho, const int32_t *__restrict__ cl, const double *__restrict__ cen,
              unsigned char *__restrict__ img, double *__restrict__ eperm, double *__restrict__ nrm2,
              unsigned long long *__restrict__ maxbits, float *__restrict__ tile_lo, float *__restrict__ tile_hi)
{
    __shared__ float smin[4], smax[4];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // sorted position
    unsigned char *tile = img + (int64_t)blockIdx.x * TC_TILE_BYTES;
    const int row = threadIdx.x;
    __half *hi_base = reinterpret_cast<__half *>(tile);
    __half *lo_base = reinterpret_cast<__half *>(tile + TC_PART_BYTES);
    const int64_t src = (int64_t)perm[i];  // -1: padding position
    const double *cj = src >= 0 ? cen + (size_t)cl[src] * d : nullptr;   // centre of the point's cluster
    double s = 0.0, sp = 0.0;
    for (int t = 0; t < TC_K; ++t) {
        __half h = __float2half(0.f), l = __float2half(0.f);
        if (src >= 0 && t < d) {
            const double v = E[src * lde + t];
            s += v * v;
            const double vs = v * scale - cj[t];   // coordinates relative to the cluster centre
            sp += vs * vs;
            split_h(vs, h, l);
        }
        const uint32_t off = swz_offset(row, t * 2);
        *reinterpret_cast<__half *>(reinterpret_cast<unsigned char *>(hi_base) + off) = h;
        *reinterpret_cast<__half *>(reinterpret_cast<unsigned char *>(lo_base) + off) = l;
    }
    // the embedding in sorted order (k_knn_rerank reads it): one row per warp pass, lanes over the coordinates
    {
        const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
        for (int r = wid * 32; r < wid * 32 + 32; ++r) {
            const int64_t pos = blockIdx.x * (int64_t)blockDim.x + r;
            const int64_t sr = (int64_t)__shfl_sync(0xffffffffu, (int)src, r & 31);
            for (int t = ln; t < d; t += 32)
                eperm[pos * d + t] = sr >= 0 ? E[sr * lde + t] : 0.0;
        }
    }
    // -0.5|r~|^2 in the three spare K columns of the hi part
    __half a1, a2, a3;
    float rlo = INFINITY, rhi = -INFINITY;  // an all-padding tile is never needed
    if (src >= 0) {
        const double hn = 0.5 * sp;
        a1 = __double2half(hn * (1.0 / 4096.0));
        const double r1 = hn - (double)__half2float(a1) * 4096.0;
        a2 = __double2half(r1);
        const double r2 = r1 - (double)__half2float(a2);
        a3 = __double2half(r2 * 4096.0);
        nrm2[src] = s;
        atomicMax(maxbits, (unsigned long long)__double_as_longlong(s));
        // distance to the cluster centre, one fp32 ulp outwards (the table holds round-to-nearest values)
        const float r = rho[src];
        rlo = fmaxf(nextafterf(r, -FLT_MAX), 0.f);
        rhi = nextafterf(r, FLT_MAX);
    } else {
        a1 = __float2half(65504.f);  // padding rows: T = -2.7e8, below every real reference
        a2 = __float2half(0.f);
        a3 = __float2half(0.f);
    }
    unsigned char *hb = reinterpret_cast<unsigned char *>(hi_base);
    *reinterpret_cast<__half *>(hb + swz_offset(row, (d + 0) * 2)) = __hneg(a1);
    *reinterpret_cast<__half *>(hb + swz_offset(row, (d + 1) * 2)) = __hneg(a2);
    *reinterpret_cast<__half *>(hb + swz_offset(row, (d + 2) * 2)) = __hneg(a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rlo = fminf(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
        rhi = fmaxf(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        smin[threadIdx.x >> 5] = rlo;
        smax[threadIdx.x >> 5] = rhi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        tile_lo[blockIdx.x] = fminf(fminf(smin[0], smin[1]), fminf(smin[2], smin[3]));
        tile_hi[blockIdx.x] = fmaxf(fmaxf(smax[0], smax[1]), fmaxf(smax[2], smax[3]));
    }
}

// Tile bounds made monotone inside every cluster (thread = cluster): hi'[t] = max hi[<= t], lo'[t] = min lo[>= t].
// The counting sort orders by rho BIN, so a bin that straddles tiles can leave hi[t] > hi[t + 1]; the widened
// bounds still contain every point of the tile, and with both sequences non-decreasing the tiles a query row can
// reach form ONE interval of the cluster -- the producer of the tile kernel finds it by binary search.
__global__ void k_knn_tc_monotone(int C, const int32_t *__restrict__ cl_tile_begin, float *__restrict__ tile_lo,
                                  float *__restrict__ tile_hi)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C)
        return;
    const int t0 = cl_tile_begin[c], t1 = cl_tile_begin[c + 1];
    float m = -INFINITY;
    for (int t = t0; t < t1; ++t) {
        m = fmaxf(m, tile_hi[t]);
        tile_hi[t] = m;
    }
    m = INFINITY;
    for (int t = t1 - 1; t >= t0; --t) {
        m = fminf(m, tile_lo[t]);
        tile_lo[t] = m;
    }
}

// ---------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------
// Warp roles: warp 0 = bulk-copy producer (+ pruning tests), warp 1 = MMA issuer, then TC_SCAN_SETS (3)
// scanner sets of 4 warps (the n-th tile of the sequence goes to set n % 3 through accumulator stage
// n % 4; within a set warp w owns TMEM lane quarter w % 4 and a thread is one query row), 4 inserter
// warps, again one thread per query row, and 4 builder warps (query images, cluster prefilter).
//
// Scanners only compare: 3-input max tree over 32 accumulator columns + one vote; a column that beats
// the row's published threshold is pushed as (value, sorted position) into the ring of (set, row) --
// single producer, single consumer, no atomics: a slot carries the lap parity of its position, so
// the consumer can tell a written slot from a stale one without resetting it, and the producer may
// write position p once head > p - QCAP.  The inserter thread of a row keeps ONE binary min-heap of
// the KP best candidates in shared memory ([entry][row]: a lane always hits its own bank); the root
// is the row's threshold, published after every change.  Candidates leave unsorted (k_knn_rerank
// sorts by exact distance).
constexpr int TC_RING = 16;  // tile ids of the sequence in flight (producer -> MMA -> scanners)
constexpr int TC_SCAN_SETS = 3;
constexpr int TC_ACC_STAGES = 3;   // = scanner sets: set s reads accumulator s
constexpr uint32_t TC_Q_COL0 = TC_ACC_STAGES * TC_N;   // TMEM columns of the query operand: 2 buffers x (32 hi + 32 lo)
constexpr int TC_POS_BITS = 27;  // sorted positions in a ring entry
constexpr int TC_THREADS = 64 + TC_SCAN_SETS * 128 + 128 + 128;  // + inserters + query-image builders

template <int KP, int STAGES, int QCAP, int NQB>
struct TcSmem {
    static constexpr size_t q_img = 0;                                      // transit of the query image (Q hi | Q lo) on its way to TMEM
    static constexpr size_t r_img = q_img + TC_TILE_BYTES;
    static constexpr size_t hk = r_img + (size_t)STAGES * TC_TILE_BYTES;    // [KP][TC_M] heap keys (T values)
    static constexpr size_t hi = hk + (size_t)TC_M * KP * 4;                // [KP][TC_M] sorted positions
    static constexpr size_t ring = hi + (size_t)TC_M * KP * 4;              // [SETS][QCAP][TC_M] uint2
    static constexpr size_t bars = ring + (size_t)TC_SCAN_SETS * QCAP * TC_M * 8;  // full/empty[STAGES], tfull/tempty[ACC]
    static constexpr size_t thr_sh = bars + (size_t)(2 * STAGES + 2 * TC_ACC_STAGES + 10) * 8;  // [TC_M] float
    static constexpr size_t pf_reach = thr_sh + TC_M * 4;                   // [TC_M] reach of every row (prefilter input)
    static constexpr size_t pf_words = pf_reach + TC_M * 4;                 // [TC_MAXC / 32] clusters still in reach
    static constexpr size_t qhead = pf_words + (TC_MAXC / 32) * 4;          // [SETS][TC_M]
    static constexpr size_t misc = qhead + TC_SCAN_SETS * TC_M * 4;         // tmem slot, scan_done
    static constexpr size_t seq_tile = misc + 16;                           // [TC_RING] tile, [TC_RING] meta
    static constexpr size_t bytes = seq_tile + 2 * TC_RING * 4;
};

// NQB = query-image buffers: 2 (the image of the next cluster is built while the current one is in use) or,
// when the candidate heaps need the room (KP = 112), 1 (the build waits for the last MMA of the visit)
// NPROD = fp16 products per tile: 3 (q_h r_h + q_h r_l + q_l r_h, the default) or, experimentally
// (CB200_KNN_PRODUCTS=2), 2: without q_l r_h the key error grows by 2^-12 |q~||r~| -- the recorded bound
// covers it (err_rel) -- and the tensor work shrinks by a third.  Measured at cfg3: 66.2 vs 67.4 ms for the
// tiles (the MMA issue is not the limiter) and 9 142 queries lose their margin proof: not worth it, kept off
template <int KP, int STAGES, int QCAP, int NQB, int NPROD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_knn_tiles_tc(const double *__restrict__ E, int64_t lde, int d, double scale, const unsigned char *__restrict__ img,
               int64_t nq, const int32_t *__restrict__ qlist, const int32_t *__restrict__ qpos,
               const int32_t *__restrict__ perm, const double *__restrict__ eperm, int C, const double *__restrict__ cen,
               const float *__restrict__ dqc,
               const int32_t *__restrict__ cl, const uint16_t *__restrict__ cl_order,
               const int32_t *__restrict__ cl_tile_begin, const unsigned int *__restrict__ cl_rhomax,
               const float *__restrict__ tile_lo, const float *__restrict__ tile_hi, float key_unscale, int cand_stride,
               int32_t *__restrict__ cand_idx, float *__restrict__ cand_tau, float *__restrict__ cand_err, float err_rel,
               int *__restrict__ err_flag, unsigned long long *__restrict__ tiles_done, int seed_max,
               long long *__restrict__ dbg)
{
    using L = TcSmem<KP, STAGES, QCAP, NQB>;
    static_assert(TC_SCAN_SETS == 3, "the seed exchange below is written for three scanner sets");
    // CB200_KNN_TIMELINE (debugging aid): the CTA in the middle of the grid records (tag, clock) pairs
    // per role into dbg[role][TL_CAP][2]
    constexpr int TL_CAP = 4096;
    const bool tl_on = dbg != nullptr && blockIdx.x == gridDim.x / 2;
    long long *cta_log = dbg != nullptr ? dbg + (size_t)7 * TL_CAP * 2 + (size_t)blockIdx.x * 4 : nullptr;
    if (cta_log != nullptr && threadIdx.x == 0) {
        unsigned long long t_ns;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ns));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        cta_log[0] = (long long)t_ns;
        cta_log[3] = (long long)smid;
    }
    int tl_n = 0;
    auto TL = [&](int role, long long tag) {
        if (tl_on && tl_n < TL_CAP) {
            long long *e = dbg + ((size_t)role * TL_CAP + tl_n) * 2;
            e[0] = tag;
            e[1] = clock64();
            ++tl_n;
        }
    };
    auto TLV = [&](int role, long long tag, long long val) {   // a value instead of the clock
        if (tl_on && tl_n < TL_CAP) {
            long long *e = dbg + ((size_t)role * TL_CAP + tl_n) * 2;
            e[0] = tag;
            e[1] = val;
            ++tl_n;
        }
    };
    constexpr uint32_t TC_TMEM_COLS = 512;  // 3 accumulators of 128 columns + 2 query-operand buffers of 64
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *q_img = smem + L::q_img;                                // 32 KB transit: Q hi | Q lo of the visit being built
    unsigned char *r_img = smem + L::r_img;                                // STAGES x 32 KB
    float *hk = reinterpret_cast<float *>(smem + L::hk);
    int *hi_ = reinterpret_cast<int *>(smem + L::hi);
    uint2 *ring = reinterpret_cast<uint2 *>(smem + L::ring);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::bars);
    uint64_t *full = bars;                       // [STAGES]
    uint64_t *empty = bars + STAGES;             // [STAGES]
    uint64_t *tfull = bars + 2 * STAGES;         // [TC_ACC_STAGES]
    uint64_t *tempty = tfull + TC_ACC_STAGES;    // [TC_ACC_STAGES]
    uint64_t *qreq = tempty + TC_ACC_STAGES;     // [2] producer -> builders: build the image of cluster q_cluster[b]
    uint64_t *qfull = qreq + 2;                  // [2] builders -> MMA issuer (and producer): image b is ready
    uint64_t *qfree = qfull + 2;                 // [2] MMA issuer -> builders: image b is no longer read
    volatile int *q_cluster = reinterpret_cast<volatile int *>(qfree + 2);  // [2]
    uint64_t *pfreq = qfree + 3;                 // producer -> builders: run the cluster prefilter
    uint64_t *pfdone = qfree + 4;                // builders -> producer
    uint64_t *seeded = qfree + 5;                // scanners -> producer: the seed thresholds are published
    volatile float *pf_reach = reinterpret_cast<volatile float *>(smem + L::pf_reach);
    volatile unsigned int *pf_words = reinterpret_cast<volatile unsigned int *>(smem + L::pf_words);
    volatile float *thr_sh = reinterpret_cast<volatile float *>(smem + L::thr_sh);
    volatile unsigned int *qhead = reinterpret_cast<volatile unsigned int *>(smem + L::qhead);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::misc);
    unsigned int *scan_done = reinterpret_cast<unsigned int *>(smem + L::misc + 4);
    volatile int *seq_tile = reinterpret_cast<volatile int *>(smem + L::seq_tile);
    volatile int *seq_meta = seq_tile + TC_RING;  // cluster | first tile of a visit << 10 | visit << 11

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ql0 = (int64_t)blockIdx.x * TC_M;   // first entry of the (sorted) query list

    // ---- one-time setup ------------------------------------------------------------------
    for (int i = tid; i < TC_SCAN_SETS * QCAP * TC_M; i += TC_THREADS)
        ring[i] = make_uint2(0u, 0x80000000u);  // parity of lap -1
    for (int i = tid; i < TC_M; i += TC_THREADS)
        thr_sh[i] = (ql0 + i < nq) ? -FLT_MAX : FLT_MAX;  // rows past the list never select
    for (int i = tid; i < TC_SCAN_SETS * TC_M; i += TC_THREADS)
        qhead[i] = 0u;
    if (tid == 0)
        *scan_done = 0u;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < TC_ACC_STAGES; ++a) {
            mbar_init(tfull + a, 1);
            mbar_init(tempty + a, 4);  // one arrival per scanner warp of the set
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(qreq + b, 1);
            mbar_init(qfull + b, 4);   // one arrival per builder warp
            mbar_init(qfree + b, 1);
        }
        mbar_init(pfreq, 1);
        mbar_init(pfdone, 4);
        mbar_init(seeded, TC_SCAN_SETS * 4);   // one arrival per scanner warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
        tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Threshold seed (pass A).  Streaming a top-KP selection over n references in arbitrary order costs about
    // KP (1 + ln(n / KP)) heap insertions per row, and most of them fall on the first tiles, where every column
    // still beats the threshold (measured at cfg3: the first 58 of ~490 tiles took 45 % of a CTA's time, all of
    // it scanners waiting for the inserters).  So the tiles [sb, se) of the home cluster first run through the
    // tensor core in a cheap mode: a scanner thread only keeps 32 running maxima of T, one per column residue
    // class mod 32 -- no compare against a threshold, no ring, no heap.  The 3 x 32 class maxima of a row belong
    // to DISTINCT references, so their KP-th largest is a lower bound of the row's KP-th best T: a valid
    // threshold, near the 39th best of the seed set for KP = 32.  The regular pass then starts with ~40
    // insertions per row instead of several hundred.  Everything the seed rejects has T <= seed <= final
    // threshold, so the margin proof of k_knn_rerank is unchanged.  The range is a pure function of the tile's
    // position: producer, scanners and inserters compute it on their own.  seed_max < 0: the first |seed_max|
    // tiles of the cluster (its core) instead of a window around the query tile's own position.
    auto seed_range = [&](int &sb, int &se, int &home_j) {
        const int64_t nv = nq - ql0 < TC_M ? nq - ql0 : TC_M;
        const int64_t mid = ql0 + nv / 2;
        home_j = cl[qlist[mid]];
        const int pt = cl_tile_begin[home_j], pte = cl_tile_begin[home_j + 1];
        const int smax_t = KP > 96 ? 0 : (seed_max < 0 ? -seed_max : seed_max);
        sb = pt;
        se = smax_t > 0 ? pte : pt;
        if (se - sb > smax_t) {
            if (seed_max > 0) {   // a window of seed_max tiles around the query tile's own position
                const int own = qpos[mid] >> 7;
                sb = own - smax_t / 2;
                sb = sb < pt ? pt : (sb > pte - smax_t ? pte - smax_t : sb);
            }
            se = sb + smax_t;
        }
    };

    // need[j] (bit j of the result words) = some row can still reach cluster j.  Lane = 4 rows (g4, reach4);
    // every lane streams its rows of the distance table (128-bit loads, 4 clusters each), the minimum
    // over the warp's rows is one redux per cluster on order-preserving integer keys.  Words [w_lo, w_hi).
    auto prefilter_words = [&](const int (&g4)[4], const float (&reach4)[4], int w_lo, int w_hi) {
        constexpr float EPS = 3e-7f;
        const bool vec = (C & 3) == 0;
        for (int w = w_lo; w < w_hi; ++w) {
            const int j0 = w * 32;
            unsigned wbits = 0u;
#pragma unroll 2
            for (int jj = 0; jj < 32; jj += 4) {
                if (j0 + jj >= C)
                    break;
                float lb4[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (g4[u] < 0)
                        continue;
                    const float *drow = dqc + (size_t)g4[u] * C + j0 + jj;
                    float v[4];
                    if (vec) {
                        const float4 q4 = *reinterpret_cast<const float4 *>(drow);
                        v[0] = q4.x; v[1] = q4.y; v[2] = q4.z; v[3] = q4.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            v[e] = j0 + jj + e < C ? drow[e] : INFINITY;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        lb4[e] = fminf(lb4[e], v[e] * (1.f - EPS) - reach4[u]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // signed-int order of the float bits is monotone for the values that occur
                    // (-inf .. +inf, no NaN: reach = inf gives -inf, inf - inf cannot happen)
                    const int key = __float_as_int(lb4[e]) >= 0 ? __float_as_int(lb4[e])
                                                                 : (int)(0x80000000u - (unsigned)__float_as_int(lb4[e]));
                    const int mn = __reduce_min_sync(0xffffffffu, key);
                    const float lbmin = __int_as_float(mn >= 0 ? mn : (int)(0x80000000u - (unsigned)mn));
                    const int j = j0 + jj + e;
                    if (j < C && lbmin <= __uint_as_float(cl_rhomax[j]) * (1.f + EPS))
                        wbits |= 1u << (jj + e);
                }
            }
            if (lane == 0)
                pf_words[w] = wbits;
        }
    };

    if (warp == 0) {
        // ===== producer: visits the clusters by ascending centre distance from the query tile's own
        // cluster and stages only the tiles some row can still reach (whole warp: lane = 4 query rows).
        // The search for the next tile runs ahead of the wait for a free stage.  After the home cluster
        // (thresholds exist by then) one warp-parallel pass over the distance table drops every cluster
        // no row can reach -- a cluster out of reach stays out of reach, the thresholds only tighten --
        // so the serial walk only touches the few candidates left. =====
        constexpr float EPS = 3e-7f;
        constexpr int NW = TC_MAXC / 32;   // candidate words: bit b of word w = position w*32 + b of `order`
        int g[4];
        float dlo[4], dhi[4], lo[4], hi[4], dq[4], smax[4], omax[4], reach[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t ql = ql0 + u * 32 + lane;
            g[u] = ql < nq ? qlist[ql] : -1;
            dlo[u] = INFINITY;   // rows past the list need nothing
            dhi[u] = -INFINITY;
            lo[u] = INFINITY;
            hi[u] = -INFINITY;
            dq[u] = 0.f;
            reach[u] = INFINITY;
            smax[u] = 0.f;       // largest |q~||r~| + 0.5|r~|^2 bound over the clusters scanned (error bound)
            omax[u] = 0.f;       // largest key offset 0.5 |q~|^2 over the clusters scanned
        }
        const int64_t nvalid = nq - ql0 < TC_M ? nq - ql0 : TC_M;
        const int home = cl[qlist[ql0 + nvalid / 2]];
        const uint16_t *order = cl_order + (size_t)home * C;
        unsigned words[NW];            // bit j: cluster j may still hold candidates (all set until the prefilter)
#pragma unroll
        for (int w = 0; w < NW; ++w)
            words[w] = 0xffffffffu;
        bool prefiltered = false;
        int opos = 0;                  // next position of `order` to look at
        unsigned ochunk = 0u;          // order[32 * (opos / 32) + lane]
        uint32_t seq = 0;
        int req_visit = -1;            // cluster visits requested from the builders so far, minus one
        // current cluster: tiles [.., te), walked in blocks of 32 (lane = tile bt0 + lane)
        int cur_j = -1, te = 0, bt0 = 0;
        int cur_visit = -1;            // visit of the cluster being walked (-1: none)
        bool cur_issued = false;       // the current visit has staged a tile
        float crmax = 0.f;
        float blo = INFINITY, bhi = INFINITY;     // (monotone) bounds of tile bt0 + lane; +inf past the cluster
        float nblo = INFINITY, nbhi = INFINITY;   // the same for the next block of the cluster, loaded a block ahead
        unsigned mask = 0u;            // tiles of the block still to stage
        // the next cluster, looked up ahead of its use: 0 nothing, 1 its loads are in flight, 2 confirmed (image requested)
        int pl_state = 0, pl_j = -1, pl_t = 0, pl_te = 0;
        float pl_crmax = 0.f, pl_dq[4] = {0.f, 0.f, 0.f, 0.f}, pl_blo = INFINITY, pl_bhi = INFINITY;
        bool no_more = false;

        auto refresh = [&]() {
            // how far every row can still find a candidate: canonical T = -0.5 d^2, rounded up
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float thr = thr_sh[u * 32 + lane];
                reach[u] = thr == -FLT_MAX ? INFINITY : sqrtf(fmaxf(-2.0f * thr, 0.f)) * 1.000002f + 1e-30f;
                if (g[u] >= 0) {
                    lo[u] = dlo[u] - reach[u];
                    hi[u] = dhi[u] + reach[u];
                }
            }
        };
        auto prefilter = [&]() {
            // the four builder warps share the pass over the distance table (they are idle right now)
#pragma unroll
            for (int u = 0; u < 4; ++u)
                pf_reach[u * 32 + lane] = reach[u];
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                mbar_arrive(pfreq);
                mbar_wait(pfdone, 0u, err_flag);
            }
            __syncwarp();
            for (int w = 0; w < NW; ++w)
                words[w] = w * 32 < C ? pf_words[w] : 0u;
        };
        auto post_request = [&](int j) {
            ++req_visit;
            if (lane == 0) {
                const int b = req_visit % NQB, use = req_visit / NQB;
                if (use >= 1)   // the builders must have taken the previous request for this buffer
                    mbar_wait(qfull + b, (uint32_t)((use - 1) & 1), err_flag);
                q_cluster[b] = j;
                __threadfence_block();
                mbar_arrive(qreq + b);
            }
        };
        // The tiles of the current block some row can reach, all 32 at once.  Row u needs tile t iff
        // hi'[t] >= lo[u] and lo'[t] <= hi[u]; both bound sequences are non-decreasing (k_knn_tc_monotone), so the
        // row's tiles are [#(hi' < lo[u]), #(lo' <= hi[u])) -- two binary searches over the lanes' bounds (shuffles
        // with a per-lane source), four rows per lane, one OR-reduction over the warp.
        auto block_mask = [&]() -> unsigned {
            refresh();
            unsigned m = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int a = 0, b = 0;
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const float va = __shfl_sync(0xffffffffu, bhi, a + step - 1);
                    const float vb = __shfl_sync(0xffffffffu, blo, b + step - 1);
                    a += (va < lo[u]) ? step : 0;
                    b += (vb <= hi[u]) ? step : 0;
                }
                const float va = __shfl_sync(0xffffffffu, bhi, 31);
                const float vb = __shfl_sync(0xffffffffu, blo, 31);
                a += (a == 31 && va < lo[u]) ? 1 : 0;
                b += (b == 31 && vb <= hi[u]) ? 1 : 0;
                if (b > a)   // a < 32 here
                    m |= (b >= 32 ? 0xffffffffu : ((1u << b) - 1u)) & ~((1u << a) - 1u);
            }
            m = __reduce_or_sync(0xffffffffu, m);
            return m & __ballot_sync(0xffffffffu, bt0 + lane < te);
        };
        // Next position of `order` whose cluster is still marked: issues the loads of everything the visit will
        // need (per-row centre distances, tile range, bounds of its first block) and returns without touching
        // them -- confirm_plan() consumes them one tile later, so their latency hides behind the tiles in flight.
        auto pick_candidate = [&]() -> bool {
            while (opos < C) {
                if (opos == 1 && !prefiltered) {   // the home cluster is done: thresholds exist now
                    prefiltered = true;
                    refresh();
                    prefilter();
                }
                if ((opos & 31) == 0)
                    ochunk = opos + lane < C ? (unsigned)order[opos + lane] : 0u;
                const int j = (int)__shfl_sync(0xffffffffu, ochunk, opos & 31);
                ++opos;
                if (((words[j >> 5] >> (j & 31)) & 1u) == 0u)
                    continue;
                const int pt = cl_tile_begin[j], pte = cl_tile_begin[j + 1];
                if (pt >= pte)
                    continue;
                pl_j = j;
                pl_t = pt;
                pl_te = pte;
                pl_crmax = __uint_as_float(cl_rhomax[j]) * (1.f + EPS);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    pl_dq[u] = g[u] >= 0 ? dqc[(size_t)g[u] * C + j] : 0.f;
                pl_blo = pt + lane < pte ? tile_lo[pt + lane] : INFINITY;
                pl_bhi = pt + lane < pte ? tile_hi[pt + lane] : INFINITY;
                pl_state = 1;
                return true;
            }
            no_more = true;
            return false;
        };
        auto confirm_plan = [&]() {
            bool need = false;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                need |= g[u] >= 0 && pl_dq[u] * (1.f - EPS) - reach[u] <= pl_crmax;
            if (__any_sync(0xffffffffu, need)) {
                // with two image buffers the build is requested now, ahead of the first tile; with one buffer
                // it waits for that tile (an early request for a visit that ends up without tiles could
                // only be released after the MMA issuer has moved on -- which needs the next tile)
                if (NQB == 2)
                    post_request(pl_j);
                pl_state = 2;
            } else {
                pl_state = 0;
            }
        };
        // between two tiles: one step of the look-ahead, never waiting on a load issued in the same call
        auto advance_plan = [&]() {
            if (pl_state == 1)
                confirm_plan();
            else if (pl_state == 0 && !no_more && (opos != 1 || prefiltered))
                pick_candidate();
        };
        // moves to the next tile some row can reach (block, cluster) -> next_t; false when nothing is left
        int next_t = 0, cur_t0 = 0;
        auto find_next = [&]() -> bool {
            while (mask == 0u) {
                if (bt0 + 32 < te) {   // next block of the cluster (its bounds were loaded a block ago)
                    bt0 += 32;
                    blo = nblo;
                    bhi = nbhi;
                } else {
                    if (cur_visit >= 0 && !cur_issued) {
                        // The image of this visit was requested ahead (two buffers), but by now no row reaches any of
                        // its tiles.  One tile is staged anyway: the MMA issuer then releases the image buffers in
                        // visit order like for every other visit.  (Releasing a skipped visit from here needs its
                        // image to be built first, and the builders may be waiting for the buffer of the visit before
                        // the last, which the MMA issuer only frees when the NEXT visit starts: two skipped visits in
                        // a row would deadlock.)  An extra tile costs ~1 k cycles and changes nothing in the result.
                        next_t = cur_t0;
                        bt0 = te;
                        return true;
                    }
                    cur_visit = -1;
                    while (pl_state != 2) {
                        if (pl_state == 1)
                            confirm_plan();
                        else if (!pick_candidate())
                            return false;
                    }
                    cur_j = pl_j;
                    te = pl_te;
                    bt0 = pl_t;
                    cur_t0 = pl_t;
                    crmax = pl_crmax;
                    blo = pl_blo;
                    bhi = pl_bhi;
                    cur_visit = NQB == 2 ? req_visit : -1;   // one buffer: requested with the first tile
                    cur_issued = false;
                    pl_state = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        dq[u] = pl_dq[u];
                        if (g[u] >= 0) {
                            dlo[u] = dq[u] * (1.f - EPS);
                            dhi[u] = dq[u] * (1.f + EPS);
                        }
                    }
                }
                const int nt = bt0 + 32 + lane;
                nblo = nt < te ? tile_lo[nt] : INFINITY;
                nbhi = nt < te ? tile_hi[nt] : INFINITY;
                mask = block_mask();
            }
            next_t = bt0 + __ffs(mask) - 1;
            mask &= mask - 1u;
            return true;
        };

        TL(0, -1);
        {   // pass A: the seed tiles of the home cluster, all of them, no tests
            int sb, se, hj;
            seed_range(sb, se, hj);
            if (se > sb) {
                post_request(hj);
                for (int ts = sb; ts < se; ++ts) {
                    if (lane == 0) {
                        const int s = (int)(seq % STAGES);
                        mbar_wait(empty + s, (uint32_t)((seq / STAGES) & 1) ^ 1u, err_flag);
                        seq_tile[seq % TC_RING] = ts;
                        seq_meta[seq % TC_RING] = hj | ((ts == sb ? 1 : 0) << 10) | (req_visit << 11);
                        mbar_expect_tx(full + s, TC_TILE_BYTES);
                        bulk_g2s(r_img + (size_t)s * TC_TILE_BYTES, img + (int64_t)ts * TC_TILE_BYTES, TC_TILE_BYTES, full + s);
                        TL(0, ((long long)ts << 8) | 2 | (ts == sb ? 16 : 0));
                    }
                    __syncwarp();
                    ++seq;
                }
                if (lane == 0)
                    mbar_wait(seeded, 0u, err_flag);   // the regular pass is planned with the seeded thresholds
                __syncwarp();
            }
        }
        while (find_next()) {
            const int t = next_t;
            if (lane == 0)
                TL(0, ((long long)t << 8) | 1);   // found
            const bool first = !cur_issued;
            cur_issued = true;
            if (first) {
                if (NQB == 1) {
                    post_request(cur_j);
                    cur_visit = req_visit;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    smax[u] = fmaxf(smax[u], dq[u] * crmax + 0.5f * crmax * crmax);
                    omax[u] = fmaxf(omax[u], 0.5f * dq[u] * dq[u]);
                }
            }
            const int s = (int)(seq % STAGES);
            if (lane == 0) {
                mbar_wait(empty + s, (uint32_t)((seq / STAGES) & 1) ^ 1u, err_flag);
                seq_tile[seq % TC_RING] = t;
                seq_meta[seq % TC_RING] = cur_j | ((first ? 1 : 0) << 10) | (cur_visit << 11);
                mbar_expect_tx(full + s, TC_TILE_BYTES);
                bulk_g2s(r_img + (size_t)s * TC_TILE_BYTES, img + (int64_t)t * TC_TILE_BYTES, TC_TILE_BYTES, full + s);
                TL(0, ((long long)t << 8) | 2 | (first ? 16 : 0));   // issued
            }
            __syncwarp();
            ++seq;
            advance_plan();
        }
        // end of the sequence
        if (lane == 0) {
            mbar_wait(empty + (int)(seq % STAGES), (uint32_t)((seq / STAGES) & 1) ^ 1u, err_flag);
            seq_tile[seq % TC_RING] = -1;
            __threadfence_block();
            mbar_arrive(full + (int)(seq % STAGES));
            atomicAdd(tiles_done, (unsigned long long)seq);
            if (cta_log != nullptr)
                cta_log[2] = (long long)seq;
            // stop the builders
            const int v = req_visit + 1, b = v % NQB, use = v / NQB;
            if (use >= 1)
                mbar_wait(qfull + b, (uint32_t)((use - 1) & 1), err_flag);
            q_cluster[b] = -1;
            __threadfence_block();
            mbar_arrive(qreq + b);
        }
        // a-priori bound of |approximate key - exact key| of every reference this row was compared with
        // (scaled^2 units -> unscaled): fp16 split + fp32 accumulation relative to |q~||r~| + 0.5|r~|^2,
        // rounding of the key offset 0.5|q~|^2, fp16 underflow floor
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t ql = ql0 + u * 32 + lane;
            if (ql < nq)
                cand_err[ql] = (err_rel * smax[u] + 4.8e-7f * omax[u] + 0.02f) * key_unscale;
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            int prev_b = -1;
            const uint64_t desc_hi = make_desc(0u);                      // everything but the start address
            const uint32_t r_desc0 = smem_u32(r_img) >> 4;
            uint32_t q_tmem = 0u;
            for (uint32_t seq = 0;; ++seq) {
                const int s = (int)(seq % STAGES);
                const uint32_t ph = (uint32_t)((seq / STAGES) & 1);
                const int a = (int)(seq % TC_ACC_STAGES);
                const uint32_t aph = (uint32_t)((seq / TC_ACC_STAGES) & 1);
                TL(1, ((long long)seq << 8) | 1);
                mbar_wait(tempty + a, aph ^ 1u, err_flag);
                TL(1, ((long long)seq << 8) | 2);
                mbar_wait(full + s, ph, err_flag);
                TL(1, ((long long)seq << 8) | 3);
                tc_fence_after();
                if (seq_tile[seq % TC_RING] < 0) {
                    // wake every scanner set with the end marker at its next sequence slot
                    for (int j = 0; j < TC_SCAN_SETS; ++j) {
                        const uint32_t sq = seq + (uint32_t)j;
                        const int a2 = (int)(sq % TC_ACC_STAGES);
                        if (j > 0)
                            mbar_wait(tempty + a2, (uint32_t)((sq / TC_ACC_STAGES) & 1) ^ 1u, err_flag);
                        seq_tile[sq % TC_RING] = -1;
                        __threadfence_block();
                        mbar_arrive(tfull + a2);
                    }
                    break;
                }
                const int meta = seq_meta[seq % TC_RING];
                if (meta & (1 << 10)) {   // first tile of a cluster visit: switch to its query image
                    const int v = meta >> 11, b = v % NQB;
                    if (prev_b >= 0)
                        umma_commit(qfree + prev_b);   // every MMA reading the previous image has been issued
                    mbar_wait(qfull + b, (uint32_t)((v / NQB) & 1), err_flag);
                    TL(1, ((long long)seq << 8) | 4);
                    tc_fence_after();
                    prev_b = b;
                    q_tmem = tmem_base + TC_Q_COL0 + (uint32_t)b * 64u;
                }
                // B descriptors: only the 14-bit start-address field (bytes >> 4) changes between operands; the A operand
                // (query image) is read from TMEM: row = lane, 8 columns (16 fp16) per K step, hi part then lo part
                const uint32_t rb = r_desc0 + (uint32_t)s * (TC_TILE_BYTES >> 4);
                const uint32_t dcol = tmem_base + (uint32_t)(a * TC_N);
#pragma unroll
                for (int p = 0; p < NPROD; ++p) {
                    const uint32_t qa = q_tmem + ((p == 2) ? 32u : 0u);
                    const uint32_t rd = rb + ((p == 1) ? (TC_PART_BYTES >> 4) : 0u);
#pragma unroll
                    for (int kk = 0; kk < TC_K / TC_UMMA_K; ++kk) {
                        const uint64_t db = desc_hi | (uint64_t)(rd + kk * (TC_UMMA_K * 2 >> 4));
                        if (p == 0 && kk == 0)
                            umma_f16_ts_first(dcol, qa + (uint32_t)(kk * 8), db, TC_IDESC);
                        else
                            umma_f16_ts_acc(dcol, qa + (uint32_t)(kk * 8), db, TC_IDESC);
                    }
                }
                umma_commit(empty + s);   // smem stage may be refilled once these MMAs have read it
                umma_commit(tfull + a);   // accumulator complete
                TL(1, ((long long)seq << 8) | 5);
            }
        }
    } else if (warp < 2 + TC_SCAN_SETS * 4) {
        // ===== scanners: thread = query row (TMEM lane) =====
        const int set = (warp - 2) >> 2;           // sequence residue of this set
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;
        const int64_t myql = ql0 + row;
        const bool active = myql < nq;
        const float *mydq = active ? dqc + (size_t)qlist[myql] * C : nullptr;
        const uint32_t ring_addr = smem_u32(ring + (size_t)set * QCAP * TC_M + row);  // slot s at + s*TC_M*8
        volatile unsigned int *myhead = qhead + set * TC_M + row;
        unsigned tail = 0u;
        int cur_j = -1;
        float off = 0.f;   // key offset of the current cluster: 0.5 |q - c_j|^2 (canonical T = T_j - off)
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t seq = (uint32_t)set;
        {   // pass A (see seed_range): 32 running maxima per row, exchanged between the sets through the (still unused)
            // heap arrays, minimum of the 32 published as the row's first threshold
            int sb, se, hj;
            seed_range(sb, se, hj);
            if (se > sb) {
                float gm[32];
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    gm[i] = -INFINITY;
                const int mypos = active ? qpos[myql] : -1;   // the query itself is not a neighbour
                for (; seq < (uint32_t)(se - sb); seq += TC_SCAN_SETS) {
                    const int a = (int)(seq % TC_ACC_STAGES);
                    if (lane == 0 && quarter == 2)
                        TL(2 + set, ((long long)seq << 8) | 1);
                    mbar_wait(tfull + a, (uint32_t)((seq / TC_ACC_STAGES) & 1), err_flag);
                    if (lane == 0 && quarter == 2)
                        TL(2 + set, ((long long)seq << 8) | 2);
                    tc_fence_after();
                    const int t = sb + (int)seq;
                    const uint32_t tcol = tbase + (uint32_t)(a * TC_N);
                    const int selfcol = (mypos >> 7) == t ? (mypos & 127) : -1;
                    uint32_t ra[16];   // 16 columns at a time: 16 + 32 live registers with the maxima (no spills here)
#pragma unroll 1
                    for (int ch = 0; ch < TC_N / 16; ++ch) {
                        tmem_ld16_sync(tcol + (uint32_t)(ch * 16), ra);
                        if ((selfcol >> 4) == ch) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                ra[i] = (i == (selfcol & 15)) ? 0xff800000u : ra[i];   // -inf
                        }
                        if (ch & 1) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                gm[16 + i] = fmaxf(gm[16 + i], __uint_as_float(ra[i]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                gm[i] = fmaxf(gm[i], __uint_as_float(ra[i]));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(tempty + a);
                    if (lane == 0 && quarter == 2)
                        TL(2 + set, ((long long)seq << 8) | 3);
                }
                // The 96 class maxima of a row (32 per set) belong to 96 distinct references: their KP-th largest is a
                // valid threshold (KP references beat or equal it).  Every thread sorts its 32 in registers (bitonic
                // network, descending), the three sorted runs meet in shared memory -- the heap arrays and the rings,
                // unused until the seed is out: the inserters wait for it -- and every thread ranks its own run
                // against the other two.  With classes of m references the result sits near the
                // 96 ln(96 / (96 - KP)) -th best of the seed set (KP = 32: the 39th).
#pragma unroll
                for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int l = i ^ jj;
                            if (l > i) {
                                const float x0 = gm[i], x1 = gm[l];
                                const bool desc = (i & kk) == 0;   // descending overall
                                gm[i] = desc ? fmaxf(x0, x1) : fminf(x0, x1);
                                gm[l] = desc ? fminf(x0, x1) : fmaxf(x0, x1);
                            }
                        }
                    }
                }
                float *X = reinterpret_cast<float *>(smem + L::hk) + row;   // [3][32][TC_M]
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    X[(set * 32 + i) * TC_M] = gm[i];
                __syncwarp();   // bar.sync is the .aligned form: the warp must arrive converged (the loops above diverge per row)
                asm volatile("bar.sync 1, %0;" ::"n"(TC_SCAN_SETS * 128) : "memory");
                if (KP <= 96) {
                    // rank of every own value in the union = its index in the own run + the number of larger values in
                    // the other two runs (binary searches; ties: the run of the lower set counts as larger), so the
                    // ranks are a permutation of 0..95 and exactly one of the three threads of a row holds rank KP-1
                    const int s1 = (set + 1) % TC_SCAN_SETS, s2 = (set + 2) % TC_SCAN_SETS;
                    const float *X1 = X + s1 * 32 * TC_M, *X2 = X + s2 * 32 * TC_M;
                    float m = 0.f;
                    bool mine = false;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = gm[i];
                        int c1 = 0, c2 = 0;
#pragma unroll
                        for (int step = 16; step >= 1; step >>= 1) {
                            const float w1 = X1[(c1 + step - 1) * TC_M], w2 = X2[(c2 + step - 1) * TC_M];
                            c1 += (s1 < set ? w1 >= v : w1 > v) ? step : 0;
                            c2 += (s2 < set ? w2 >= v : w2 > v) ? step : 0;
                        }
                        const float w1 = X1[31 * TC_M], w2 = X2[31 * TC_M];
                        c1 += (c1 == 31 && (s1 < set ? w1 >= v : w1 > v)) ? 1 : 0;
                        c2 += (c2 == 31 && (s2 < set ? w2 >= v : w2 > v)) ? 1 : 0;
                        if (i + c1 + c2 == KP - 1) {
                            m = v;
                            mine = true;
                        }
                    }
                    if (mine && active && m > -1e30f) {   // (fewer than KP real references in the seed tiles: no seed)
                        const float dqh = mydq[hj];
                        const float offh = 0.5f * dqh * dqh;
                        // canonical T, a few ulps down: the KP references at or above m must pass `T > thr + off`
                        thr_sh[row] = (m - offh) - ((fabsf(m) + offh) * 4.8e-7f + 1e-30f);   // strictly below, also at m = offh = 0
                    }
                }
                __syncwarp();   // bar.sync is the .aligned form: the warp must arrive converged (the loops above diverge per row)
                asm volatile("bar.sync 1, %0;" ::"n"(TC_SCAN_SETS * 128) : "memory");
                for (int e = 0; e < QCAP; ++e)   // the exchange overwrote the rings: empty again (parity of lap -1)
                    ring[((size_t)set * QCAP + e) * TC_M + row] = make_uint2(0u, 0x80000000u);
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    mbar_arrive(seeded);
                }
            }
        }
        for (;; seq += TC_SCAN_SETS) {
            const int a = (int)(seq % TC_ACC_STAGES);
            if (lane == 0 && quarter == 2)
                TL(2 + set, ((long long)seq << 8) | 1);
            mbar_wait(tfull + a, (uint32_t)((seq / TC_ACC_STAGES) & 1), err_flag);
            if (lane == 0 && quarter == 2)
                TL(2 + set, ((long long)seq << 8) | 2);
            tc_fence_after();
            const int t = seq_tile[seq % TC_RING];
            if (t < 0)
                break;
            const int j = seq_meta[seq % TC_RING] & 1023;
            if (j != cur_j) {
                cur_j = j;
                const float dq = active ? mydq[j] : 0.f;
                off = 0.5f * dq * dq;
            }
            const int r0 = t * TC_N;
            const uint32_t tcol = tbase + (uint32_t)(a * TC_N);
            uint32_t r[32];
            long long tl_ld = 0, tl_hit = 0, tl_nhit = 0, tl_c0 = 0, tl_c1 = 0;
            const bool tl_me = tl_on && lane == 0 && quarter == 2;
#pragma unroll 1
            for (int ch = 0; ch < TC_N / 32; ++ch) {
                if (tl_me)
                    tl_c0 = clock64();
                tmem_ld32_issue(tcol + (uint32_t)(ch * 32), r);
                const float thr = thr_sh[row] + off;   // stale reads only cost extra pushes
                tmem_ld_wait();
                if (tl_me) {
                    tl_c1 = clock64();
                    tl_ld += tl_c1 - tl_c0;
                }
                // max tree: 32 -> 12 partial maxima (3, 3, 2 columns) -> 4 group maxima -> 1
                float pm[12], g[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    pm[3 * jj] = fmaxf(fmaxf(__uint_as_float(r[8 * jj]), __uint_as_float(r[8 * jj + 1])),
                                       __uint_as_float(r[8 * jj + 2]));
                    pm[3 * jj + 1] = fmaxf(fmaxf(__uint_as_float(r[8 * jj + 3]), __uint_as_float(r[8 * jj + 4])),
                                           __uint_as_float(r[8 * jj + 5]));
                    pm[3 * jj + 2] = fmaxf(__uint_as_float(r[8 * jj + 6]), __uint_as_float(r[8 * jj + 7]));
                    g[jj] = fmaxf(fmaxf(pm[3 * jj], pm[3 * jj + 1]), pm[3 * jj + 2]);
                }
                const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
                if (!__any_sync(0xffffffffu, m > thr))
                    continue;
                // hit path, branch-free per lane: a 32-bit mask of the columns that beat the threshold, then one warp-
                // uniform loop that takes the lowest set bit of every lane, fetches its value with a 5-level select
                // tree (registers cannot be indexed) and pushes it.  The earlier tree walk (descend only into the
                // groups with a hit) executed ~10 comparisons per hit but diverged per lane: measured 1 200 - 3 700
                // cycles per chunk with a hit, 70 % of a scanner warp's time.
                unsigned hm = 0u;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    hm |= (__uint_as_float(r[i]) > thr ? 1u : 0u) << i;
                while (__any_sync(0xffffffffu, hm != 0u)) {
                    const int i = hm != 0u ? __ffs(hm) - 1 : 0;
                    uint32_t s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        s16[q] = (i & 1) ? r[2 * q + 1] : r[2 * q];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        s8[q] = (i & 2) ? s16[2 * q + 1] : s16[2 * q];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        s4[q] = (i & 4) ? s8[2 * q + 1] : s8[2 * q];
                    s2[0] = (i & 8) ? s4[1] : s4[0];
                    s2[1] = (i & 8) ? s4[3] : s4[2];
                    const uint32_t vbits = (i & 16) ? s2[1] : s2[0];
                    if (hm != 0u) {
                        const int gcol = r0 + ch * 32 + i;
                        for (uint32_t spin = 0; (int)(tail - *myhead) >= QCAP; ++spin) {
                            if (spin > (1u << 24)) {
                                atomicOr(err_flag, 4);
                                __threadfence();
                                asm volatile("trap;");
                            }
                        }
                        const unsigned meta = (((tail / QCAP) & 1u) << 31) | (unsigned)gcol;
                        asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(
                                         ring_addr + (uint32_t)((tail % QCAP) * TC_M * 8)),
                                     "r"(__float_as_uint(__uint_as_float(vbits) - off)), "r"(meta)
                                     : "memory");
                        ++tail;
                        hm &= hm - 1u;
                    }
                }
                __syncwarp();
                if (tl_me) {
                    tl_hit += clock64() - tl_c1;
                    ++tl_nhit;
                }
            }
            if (tl_me) {
                TLV(2 + set, ((long long)seq << 8) | 7, tl_ld);
                TLV(2 + set, ((long long)seq << 8) | 8, tl_hit);
                TLV(2 + set, ((long long)seq << 8) | 9, tl_nhit);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(tempty + a);
            if (lane == 0 && quarter == 2) {
                TL(2 + set, ((long long)seq << 8) | 3);
                TLV(2 + set, ((long long)seq << 8) | 6, (long long)tail);   // pushes of this row so far
            }
        }
        if (tl_on)   // pushes of every row (summed over the sets)
            atomicAdd(reinterpret_cast<unsigned long long *>(dbg + ((size_t)6 * TL_CAP + row) * 2), (unsigned long long)tail);
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(scan_done, 1u);
        }
    } else if (warp < 2 + TC_SCAN_SETS * 4 + 4) {
        // ===== inserters: thread = query row; binary min-heap of the KP largest T in shared memory =====
        const int row = tid - (64 + TC_SCAN_SETS * 128);
        float *mk = hk + row;   // entry e at mk[e * TC_M]
        int *mi = hi_ + row;
        unsigned head[TC_SCAN_SETS];
        uint32_t raddr[TC_SCAN_SETS];
#pragma unroll
        for (int s2 = 0; s2 < TC_SCAN_SETS; ++s2) {
            head[s2] = 0u;
            raddr[s2] = smem_u32(ring + (size_t)s2 * QCAP * TC_M + row);
        }
        float root = -FLT_MAX;
        long long n_ins = 0;
        const int my_pos = ql0 + row < nq ? qpos[ql0 + row] : -1;   // sorted position of the query itself
        bool done_seen = false;
        {   // the rings carry the scanners' seed exchange first (pass A): nothing to read before it is over
            int sb, se, hj;
            seed_range(sb, se, hj);
            if (se > sb)
                mbar_wait(seeded, 0u, err_flag);
        }
        // The heap starts FULL of placeholders (no reference, key = the seeded threshold, or -FLT_MAX without a
        // seed): every insertion is the same sift-down from the root, and rows that reach KP real candidates at
        // different moments never run a divergent heapify (measured: that serialised the inserter warps for
        // ~100 k cycles per CTA once the seed spread the fill over many tiles).
        root = ql0 + row < nq ? thr_sh[row] : FLT_MAX;
        for (int e = 0; e < KP; ++e) {
            mk[e * TC_M] = root;
            mi[e * TC_M] = -1;
        }
        while (true) {
            bool got = false;
#pragma unroll
            for (int s2 = 0; s2 < TC_SCAN_SETS; ++s2) {
                unsigned ex, ey;
                asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];"
                             : "=r"(ex), "=r"(ey)
                             : "r"(raddr[s2] + (uint32_t)((head[s2] % QCAP) * TC_M * 8))
                             : "memory");
                if ((ey >> 31) != ((head[s2] / QCAP) & 1u))
                    continue;
                got = true;
                ++head[s2];
                qhead[s2 * TC_M + row] = head[s2];
                const float val = __uint_as_float(ex);
                const int pos = (int)(ey & ((1u << TC_POS_BITS) - 1u));
                if (pos == my_pos)
                    continue;   // the query itself
                if (val > root) {
                    ++n_ins;
                    int i = 0;
                    while (true) {  // sift the new value down from the root
                        const int l = 2 * i + 1;
                        if (l >= KP)
                            break;
                        const float kl = mk[l * TC_M];
                        const float kr = (l + 1 < KP) ? mk[(l + 1) * TC_M] : FLT_MAX;
                        const int c = kl <= kr ? l : l + 1;
                        const float kc = fminf(kl, kr);
                        if (val <= kc)
                            break;
                        mk[i * TC_M] = kc;
                        mi[i * TC_M] = mi[c * TC_M];
                        i = c;
                    }
                    mk[i * TC_M] = val;
                    mi[i * TC_M] = pos;
                    root = mk[0];
                    thr_sh[row] = root;
                }
            }
            if (!__any_sync(0xffffffffu, got)) {
                if (done_seen)
                    break;
                done_seen = *reinterpret_cast<volatile unsigned int *>(scan_done) == (unsigned)(TC_SCAN_SETS * 4);
                if (!done_seen)
                    __nanosleep(32);
            }
        }
        // final: candidates (sorted positions, unsorted order, -1 where a placeholder survived) and the row
        // threshold in key space, unscaled: key = 0.5 d^2 = -T.  Every reference that is not a candidate was
        // pruned or rejected against a threshold <= root.
        const int64_t myql = ql0 + row;
        if (tl_on)
            dbg[((size_t)6 * TL_CAP + row) * 2 + 1] = n_ins;
        if (myql < nq) {
            int32_t *out = cand_idx + myql * (int64_t)cand_stride;
            for (int p = 0; p < cand_stride; ++p)
                out[p] = p < KP ? mi[p * TC_M] : -1;   // k_knn_rerank maps the positions through perm
            cand_tau[myql] = root == -FLT_MAX ? FLT_MAX : -root * key_unscale;
        }
    } else {
        // ===== builders: the query operand image relative to the centre of the cluster being visited,
        // into the image buffer the producer names.  Warp w owns rows 32w .. 32w+31; a row is loaded by
        // the whole warp (lane = coordinate: coalesced), several rows in flight =====
        const int bw = warp - (2 + TC_SCAN_SETS * 4 + 4);   // share of the prefilter pass
        const int bq = warp & 3;                             // TMEM lane quarter this warp may write = its 32 query rows
        const int64_t rowq = ql0 + bq * 32 + lane;
        const int my_p = rowq < nq ? qpos[rowq] : -1;    // lane r: sorted position of row 32 bq + r
        bool pf_served = false;
        for (int visit = 0;; ++visit) {
            const int b = visit % NQB, use = visit / NQB;
            // wait for the next image request; meanwhile serve the (single) prefilter request
            // (polling only until the prefilter is served; afterwards the hardware-suspended wait below)
            for (uint32_t spin = 0; !pf_served && !mbar_test(qreq + b, (uint32_t)(use & 1)); ++spin) {
                if (mbar_test(pfreq, 0u)) {
                    pf_served = true;
                    int g4[4];
                    float reach4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t ql = ql0 + u * 32 + lane;
                        g4[u] = ql < nq ? qlist[ql] : -1;
                        reach4[u] = pf_reach[u * 32 + lane];
                    }
                    const int nw = (C + 31) / 32, per = (nw + 3) / 4;
                    const int w_lo = bw * per < nw ? bw * per : nw, w_hi = w_lo + per < nw ? w_lo + per : nw;
                    prefilter_words(g4, reach4, w_lo, w_hi);
                    __syncwarp();
                    if (lane == 0) {
                        __threadfence_block();
                        mbar_arrive(pfdone);
                    }
                    spin = 0;
                }
                if (spin > (1u << 24)) {
                    atomicOr(err_flag, 4);
                    __threadfence();
                    asm volatile("trap;");
                }
                __nanosleep(200);
            }
            mbar_wait(qreq + b, (uint32_t)(use & 1), err_flag);
            const int j = q_cluster[b];
            if (j < 0)
                break;
            if (bw == 0 && lane == 0)
                TL(5, ((long long)visit << 8) | 1);
            if (use >= 1)
                mbar_wait(qfree + b, (uint32_t)((use - 1) & 1), err_flag);
            if (bw == 0 && lane == 0)
                TL(5, ((long long)visit << 8) | 2);
            const double *cj = cen + (size_t)j * d;
            // lane = the column pair (2 lane, 2 lane + 1): one packed 32-bit store per part and row
            const int t0 = 2 * lane, t1 = 2 * lane + 1;
            const double c0 = t0 < d ? cj[t0] : 0.0, c1 = t1 < d ? cj[t1] : 0.0;
            unsigned char *qb = q_img;   // transit: this warp's 32 rows, then row-wise into TMEM buffer b
            // constants of the columns past the coordinates: [2^12, 1, 2^-12] (hi part), 0 elsewhere
            auto kconst = [&](int t) { return t == d ? 4096.f : (t == d + 1 ? 1.f : (t == d + 2 ? 1.0f / 4096.f : 0.f)); };
            const float k0 = kconst(t0), k1 = kconst(t1);
            const uint32_t lane_chunk = (uint32_t)(lane >> 2), lane_byte = (uint32_t)((lane * 4) & 15);
#pragma unroll 1
            for (int r0 = 0; r0 < 32; r0 += 8) {
                double x0[8], x1[8];
                int ps[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    ps[u] = __shfl_sync(0xffffffffu, my_p, r0 + u);
                    const double *qrow = eperm + (int64_t)(ps[u] < 0 ? 0 : ps[u]) * d;   // the sorted-order copy
                    x0[u] = (ps[u] >= 0 && t0 < d) ? qrow[t0] : 0.0;
                    x1[u] = (ps[u] >= 0 && t1 < d) ? qrow[t1] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int row = bq * 32 + r0 + u;
                    __half h0 = __float2half(ps[u] >= 0 ? k0 : 0.f), l0 = __float2half(0.f);
                    __half h1 = __float2half(ps[u] >= 0 ? k1 : 0.f), l1 = __float2half(0.f);
                    if (ps[u] >= 0 && t0 < d)
                        split_h(x0[u] * scale - c0, h0, l0);
                    if (ps[u] >= 0 && t1 < d)
                        split_h(x1[u] * scale - c1, h1, l1);
                    const uint32_t o = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128) +
                                       (((lane_chunk ^ (uint32_t)(row & 7)) & 7u) << 4) + lane_byte;
                    const uint32_t hp = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    const uint32_t lp = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                    *reinterpret_cast<uint32_t *>(qb + o) = hp;
                    *reinterpret_cast<uint32_t *>(qb + TC_PART_BYTES + o) = lp;
                }
            }
            __syncwarp();
            // transit -> TMEM: thread = query row (TMEM lane); its 128-byte hi row and lo row, chunks un-swizzled,
            // are 32 + 32 columns of query buffer b
            {
                const int row = bq * 32 + lane;
                const unsigned char *rp = qb + (size_t)((row >> 3) * 1024 + (row & 7) * 128);
                const uint32_t tq = tmem_base + ((uint32_t)(bq * 32) << 16) + TC_Q_COL0 + (uint32_t)b * 64u;
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    uint32_t v[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 q4 = *reinterpret_cast<const uint4 *>(rp + (size_t)part * TC_PART_BYTES +
                                                                          (size_t)(((c ^ (row & 7)) & 7) << 4));
                        v[4 * c] = q4.x;
                        v[4 * c + 1] = q4.y;
                        v[4 * c + 2] = q4.z;
                        v[4 * c + 3] = q4.w;
                    }
                    tmem_st32(tq + (uint32_t)(part * 32), v);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(qfull + b);
            if (bw == 0 && lane == 0)
                TL(5, ((long long)visit << 8) | 3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (cta_log != nullptr && threadIdx.x == 0) {
        unsigned long long t_ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ns));
        cta_log[1] = (long long)t_ns;
    }
    if (warp == 0)
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int KP, int STAGES, int QCAP, int NQB, int NPROD = 3>
static int launch_tc(cb200_ctx *ctx, const double *E, int64_t lde, int d, double scale, int64_t nq, int C,
                     int cand_stride)
{
    const size_t smem = TcSmem<KP, STAGES, QCAP, NQB>::bytes;
    static_assert(TcSmem<KP, STAGES, QCAP, NQB>::bytes <= 232448, "shared memory budget");
    CB_CUDA(ctx, cudaFuncSetAttribute(k_knn_tiles_tc<KP, STAGES, QCAP, NQB, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const unsigned grid = (unsigned)((nq + TC_M - 1) / TC_M);
    const float key_unscale = (float)(1.0 / (scale * scale));
    // tiles of the home cluster used for the threshold seed (0: off)
    const char *seed_env = getenv("CB200_KNN_SEED");
    const int seed_max = seed_env != nullptr ? atoi(seed_env) : 64;
    long long *dbg = nullptr;
    if (getenv("CB200_KNN_TIMELINE") != nullptr) {
        CB_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&dbg), ((size_t)7 * 4096 * 2 + (size_t)grid * 4) * sizeof(long long)));
        CB_CUDA(ctx, cudaMemsetAsync(dbg, 0, ((size_t)7 * 4096 * 2 + (size_t)grid * 4) * sizeof(long long), ctx->stream));
    }
    k_knn_tiles_tc<KP, STAGES, QCAP, NQB, NPROD><<<grid, TC_THREADS, smem, ctx->stream>>>(
        E, lde, d, scale, ctx->knn_img.p, nq, ctx->knn_qlist.p, ctx->knn_qpos.p, ctx->knn_perm.p, ctx->knn_eperm.p, C,
        ctx->knn_cen.p,
        ctx->knn_dqc.p, ctx->knn_cl.p, ctx->knn_cl_order.p, ctx->knn_cl_tile_begin.p, ctx->knn_cl_rhomax.p,
        ctx->knn_tile_lo.p, ctx->knn_tile_hi.p, key_unscale, cand_stride, ctx->cand_idx.p, ctx->cand_tau.p,
        ctx->cand_err.p, NPROD == 3 ? 4.6e-5f : 3.0e-4f, ctx->flag.p + 3, ctx->maxnrm.p + 3, seed_max, dbg);
    CB_LAUNCH(ctx);
    if (dbg != nullptr) {   // dump the timeline of the middle CTA
        std::vector<long long> h((size_t)7 * 4096 * 2 + (size_t)grid * 4);   // role timelines of the middle CTA, then [start ns, end ns, tiles, SM] of every CTA
        CB_CUDA(ctx, cudaMemcpyAsync(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        FILE *f = fopen(getenv("CB200_KNN_TIMELINE"), "wb");
        if (f != nullptr) {
            fwrite(h.data(), sizeof(long long), h.size(), f);
            fclose(f);
        }
        cudaFree(dbg);
    }
    return 0;
}

bool cb200_knn_tc_supported(int d, int k) { return d + 3 <= TC_K && k <= 100; }

// number of clusters of the reference ordering: about one per 512 points up to 128, then one per 16 384
// points up to 256 (measured at cfg3: 128 clusters 74 ms for ordering + tiles, 256 clusters 80 ms -- fewer
// tiles but more cluster visits per query tile); CB200_KNN_CLUSTERS overrides, up to TC_MAXC; any value
// gives the same, exact, result
static int tc_num_clusters(int64_t n)
{
    int64_t c = n / 512;
    const int64_t hi = n / 16384 > 128 ? n / 16384 : 128;
    if (c > hi)
        c = hi;
    if (c > 256)
        c = 256;
    const char *e = getenv("CB200_KNN_CLUSTERS");
    if (e != nullptr && atoi(e) > 0)
        c = atoi(e);
    if (c > TC_MAXC)
        c = TC_MAXC;
    if (c > n / 2)
        c = n / 2;
    if (c < 1)
        c = 1;
    return (int)c;
}

// Clusters and orders the references, builds the operand image and runs the tensor-core tile
// kernel; fills ctx->cand_idx / ctx->cand_tau (indexed by position in the sorted query list),
// ctx->nrm2, ctx->maxnrm.  *kp_out = candidates kept per query, *ntau_out = thresholds per query
// (every non-candidate has key >= the smallest of them), *qlist_out = original index of every entry
// of the query list (device), *err_coef_out = coefficient of the a-priori key error bound.
int cb200_knn_tiles_tc(cb200_ctx *ctx, const double *E, int64_t lde, int64_t n, int d, int k, int64_t q_begin,
                       int64_t q_end, int *kp_out, int *ntau_out, const int32_t **qlist_out, double *err_coef_out)
{
    const int64_t nq = q_end - q_begin;
    const int C = tc_num_clusters(n);
    const int B = TC_BUCKETS / C;
    const int64_t n_tiles_cap = (n + TC_N - 1) / TC_N + C;  // every cluster is padded to whole tiles
    const int64_t npos = n_tiles_cap * TC_N;
    // candidates kept per query (one list): enough beyond k for the margin proof of k_knn_rerank
    const int KP = k <= 24 ? 32 : (k <= 52 ? 64 : 112);
    const int CAND = KP == 112 ? 128 : KP;   // the re-rank sorts a power of two: padded with empty entries
    CB_CUDA(ctx, ctx->knn_img.ensure((size_t)n_tiles_cap * TC_TILE_BYTES));
    CB_CUDA(ctx, ctx->nrm2.ensure((size_t)n));
    CB_CUDA(ctx, ctx->knn_eperm.ensure((size_t)npos * d));
    CB_CUDA(ctx, ctx->cand_idx.ensure((size_t)nq * CAND));
    CB_CUDA(ctx, ctx->cand_tau.ensure((size_t)nq));
    CB_CUDA(ctx, ctx->cand_err.ensure((size_t)nq));
    CB_CUDA(ctx, ctx->knn_perm.ensure((size_t)npos));
    CB_CUDA(ctx, ctx->knn_bkt.ensure((size_t)npos));
    CB_CUDA(ctx, ctx->knn_hist.ensure((size_t)TC_BUCKETS));
    CB_CUDA(ctx, ctx->knn_cursor.ensure((size_t)TC_BUCKETS + 1));
    CB_CUDA(ctx, ctx->knn_tile_lo.ensure((size_t)n_tiles_cap));
    CB_CUDA(ctx, ctx->knn_tile_hi.ensure((size_t)n_tiles_cap));
    CB_CUDA(ctx, ctx->maxnrm.ensure(4));
    CB_CUDA(ctx, ctx->knn_cen.ensure((size_t)C * d));
    CB_CUDA(ctx, ctx->knn_km_sums.ensure((size_t)C * d));
    CB_CUDA(ctx, ctx->knn_km_cnt.ensure((size_t)C));
    CB_CUDA(ctx, ctx->knn_dqc.ensure((size_t)n * C));
    CB_CUDA(ctx, ctx->knn_cl.ensure((size_t)n));
    CB_CUDA(ctx, ctx->knn_rho.ensure((size_t)n));
    CB_CUDA(ctx, ctx->knn_cl_rhomax.ensure((size_t)C));
    CB_CUDA(ctx, ctx->knn_cl_tile_begin.ensure((size_t)C + 1));
    CB_CUDA(ctx, ctx->knn_pad_before.ensure((size_t)C));
    CB_CUDA(ctx, ctx->knn_cl_order.ensure((size_t)C * C));
    CB_CUDA(ctx, ctx->knn_qlist.ensure((size_t)nq));
    CB_CUDA(ctx, ctx->knn_qpos.ensure((size_t)nq));
    CB_CUDA(ctx, ctx->knn_qoff.ensure((size_t)npos + 1));

    cb_stage_begin(ctx, ST_KNN_PREP);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->maxnrm.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    const unsigned g256 = (unsigned)((n + 255) / 256);
    k_knn_tc_ranges<<<g256, 256, 0, ctx->stream>>>(E, lde, n, d, ctx->maxnrm.p);
    CB_LAUNCH(ctx);
    double maxabs = 0.0;
    CB_TRY(cb200_read_back(ctx, &maxabs, ctx->maxnrm.p, sizeof(double)));
    CB_CHECK(ctx, maxabs == maxabs && maxabs < 1e300, "cb200_knn: embedding contains non-finite values");
    // power-of-two scale: max |x'| in [256, 512)
    double scale = 1.0;
    if (maxabs > 0.0) {
        int e;
        frexp(maxabs, &e);          // maxabs = f * 2^e, f in [0.5, 1)
        scale = ldexp(1.0, 9 - e);  // maxabs * scale in [256, 512)
    }
    // clusters: Lloyd iterations on a strided sample of at most 65536 points, then the distance table
    const unsigned gcd = (unsigned)((C * d + 255) / 256);
    k_knn_tc_km_init<<<gcd, 256, 0, ctx->stream>>>(E, lde, n, d, scale, C, ctx->knn_cen.p);
    CB_LAUNCH(ctx);
    if (C > 1) {
        const int64_t stride = n > 65536 ? n / 65536 : 1;
        const int64_t M = (n + stride - 1) / stride;
        for (int it = 0; it < 4; ++it) {
            CB_CUDA(ctx, cudaMemsetAsync(ctx->knn_km_sums.p, 0, (size_t)C * d * sizeof(double), ctx->stream));
            CB_CUDA(ctx, cudaMemsetAsync(ctx->knn_km_cnt.p, 0, (size_t)C * sizeof(unsigned long long), ctx->stream));
            k_knn_tc_assign<false><<<(unsigned)((M + 127) / 128), 128, 0, ctx->stream>>>(
                E, lde, n, d, scale, stride, M, C, ctx->knn_cen.p, ctx->knn_km_sums.p, ctx->knn_km_cnt.p, nullptr, nullptr,
                nullptr, nullptr);
            CB_LAUNCH(ctx);
            k_knn_tc_km_update<<<gcd, 256, 0, ctx->stream>>>(C, d, ctx->knn_km_sums.p, ctx->knn_km_cnt.p, ctx->knn_cen.p);
            CB_LAUNCH(ctx);
        }
    }
    CB_CUDA(ctx, cudaMemsetAsync(ctx->knn_cl_rhomax.p, 0, (size_t)C * sizeof(unsigned int), ctx->stream));
    k_knn_tc_assign<true><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
        E, lde, n, d, scale, 1, n, C, ctx->knn_cen.p, nullptr, nullptr, ctx->knn_dqc.p, ctx->knn_cl.p, ctx->knn_rho.p,
        ctx->knn_cl_rhomax.p);
    CB_LAUNCH(ctx);
    k_knn_tc_corder<<<C, TC_MAXC, 0, ctx->stream>>>(C, d, ctx->knn_cen.p, ctx->knn_cl_order.p);
    CB_LAUNCH(ctx);
    // counting sort by (cluster, rho bin), clusters padded to whole tiles
    CB_CUDA(ctx, cudaMemsetAsync(ctx->knn_hist.p, 0, TC_BUCKETS * sizeof(int32_t), ctx->stream));
    k_knn_tc_bucket<<<g256, 256, 0, ctx->stream>>>(n, B, ctx->knn_cl.p, ctx->knn_rho.p, ctx->knn_cl_rhomax.p,
                                                   ctx->knn_bkt.p, ctx->knn_hist.p);
    CB_LAUNCH(ctx);
    CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->knn_hist.p, TC_BUCKETS, reinterpret_cast<int64_t *>(ctx->knn_cursor.p)));
    k_knn_tc_layout<<<1, 32, 0, ctx->stream>>>(C, B, ctx->knn_cursor.p, ctx->knn_cl_tile_begin.p, ctx->knn_pad_before.p);
    CB_LAUNCH(ctx);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->knn_perm.p, 0xFF, (size_t)npos * sizeof(int32_t), ctx->stream));
    k_knn_tc_scatter<<<g256, 256, 0, ctx->stream>>>(ctx->knn_bkt.p, n, B, ctx->knn_cursor.p, ctx->knn_pad_before.p,
                                                    ctx->knn_perm.p);
    CB_LAUNCH(ctx);
    // image, norms, tile bounds (maxnrm[0] becomes max |r|^2)
    CB_CUDA(ctx, cudaMemsetAsync(ctx->maxnrm.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    k_knn_tc_prep<<<(unsigned)n_tiles_cap, 128, 0, ctx->stream>>>(E, lde, d, scale, ctx->knn_perm.p, ctx->knn_rho.p,
                                                                  ctx->knn_cl.p, ctx->knn_cen.p, ctx->knn_img.p,
                                                                  ctx->knn_eperm.p, ctx->nrm2.p, ctx->maxnrm.p, ctx->knn_tile_lo.p,
                                                                  ctx->knn_tile_hi.p);
    CB_LAUNCH(ctx);
    k_knn_tc_monotone<<<(unsigned)((C + 127) / 128), 128, 0, ctx->stream>>>(C, ctx->knn_cl_tile_begin.p, ctx->knn_tile_lo.p,
                                                                         ctx->knn_tile_hi.p);
    CB_LAUNCH(ctx);
    // the query list: the cells of [q_begin, q_end) in sorted order, with their sorted positions
    const unsigned gpos = (unsigned)((npos + 255) / 256);
    k_knn_tc_qflag<<<gpos, 256, 0, ctx->stream>>>(ctx->knn_perm.p, npos, q_begin, q_end, ctx->knn_bkt.p);
    CB_LAUNCH(ctx);
    CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->knn_bkt.p, npos, ctx->knn_qoff.p));
    k_knn_tc_qscatter<<<gpos, 256, 0, ctx->stream>>>(ctx->knn_perm.p, ctx->knn_bkt.p, ctx->knn_qoff.p, npos,
                                                     ctx->knn_qlist.p, ctx->knn_qpos.p);
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_KNN_PREP);

    cb_stage_begin(ctx, ST_KNN_TILE);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 3, 0, sizeof(int), ctx->stream));
    CB_CHECK(ctx, npos < (1LL << TC_POS_BITS), "cb200_knn: at most 2^27 points on the tensor-core engine");
    const char *np_env = getenv("CB200_KNN_PRODUCTS");
    if (KP == 32 && np_env != nullptr && atoi(np_env) == 2)
        CB_TRY((launch_tc<32, 4, 8, 2, 2>(ctx, E, lde, d, scale, nq, C, CAND)));
    else if (KP == 32)
        CB_TRY((launch_tc<32, 4, 8, 2>(ctx, E, lde, d, scale, nq, C, CAND)));
    else if (KP == 64)
        CB_TRY((launch_tc<64, 3, 8, 2>(ctx, E, lde, d, scale, nq, C, CAND)));
    else
        CB_TRY((launch_tc<112, 2, 4, 2>(ctx, E, lde, d, scale, nq, C, CAND)));
    cb_stage_end(ctx, ST_KNN_TILE);
    *kp_out = CAND;
    *ntau_out = 1;
    *qlist_out = ctx->knn_qlist.p;
    // the bound of every query is in ctx->cand_err (see the producer of k_knn_tiles_tc); a negative
    // coefficient tells k_knn_rerank to use it with the canonical key 0.5 d^2
    *err_coef_out = -1.0;
    return 0;
}
