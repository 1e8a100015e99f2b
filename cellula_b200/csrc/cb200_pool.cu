// cb200_pool.cu -- SURVEY 8(f) rank 1: deconvolution ("pooled") size factors on the GPU.
//
// Replaces: scran::computeSumFactors(sce, clusters = ...) = scuttle::pooledSizeFactors, called by
// /root/reference/R/doNormAndReduce.R:66-71 between quickCluster (:57-65; its walktrap clustering stays igraph,
// the cluster labels are an INPUT here) and logNormCounts (:78).  Published algorithm: Lun, Bach & Marioni,
// Genome Biology 17:75 (2016); package internals restated from memory (un-vendored: parity unpinned), the same
// restatement as oracle/oracle.py:pooled_size_factors, which the GPU tests compare against.
//
// Per cluster (clusters above max_cluster_size are dealt round-robin into sub-clusters first):
//   1. y_gc = x_gc / lib_c; pseudo-cell ave_g = mean_c(y_gc) * mean(lib); genes with ave_g >= min_mean are kept.
//   2. cells are placed on a ring by library size (odd ranks ascending, even ranks descending).
//   3. for every ring position p and pool size s (21, 26, ... 101 <= n): b_s[p] = median over the kept genes of
//      (sum of y over the s cells starting at p) / ave_g; plus b_1[p] with weight sqrt(1e-6).
//   4. least squares for the per-cell factors theta: every pool says sum_{c in pool} theta_c = b.
//   5. nf_c = theta_c * lib_c; clusters are brought to one scale by the median ratio of their pseudo-cells to a
//      reference cluster's (host, G-sized), and the vector is scaled to unit mean over the positive factors.
//
// B200 design.
//   * Step 3 is the cost (N x 18 medians over ~5 000 genes; minutes on a CPU).  k_pool_medians: one CTA owns a run
//     of consecutive ring positions for ONE pool size and keeps the pooled profile of the current window in shared
//     memory as 64-bit FIXED-POINT integers (2^s2 * y, exact integer adds), so sliding the window -- add the entering
//     cell's column, subtract the leaving one's -- is exact and drift-free, and an empty gene is exactly zero.
//     The median is a bucket selection in shared memory: zeros are counted apart, the non-zero ratios are
//     histogrammed on fp32 keys into 1 024 buckets spanning [0, 4 x previous median), the bucket holding the
//     wanted rank is gathered (a handful of genes) and ranked exactly in fp64 on (ratio, gene).  The range adapts
//     (widen / narrow) until the bucket is small, so any distribution terminates; ties collapse to one value.
//     fp32 keys only decide bucket membership: a gene within 1e-7 relative of a bucket edge may be ranked across
//     it, which moves the median by <= 1e-7 relative (contract: size factors within 1e-6 relative).
//   * Step 4: in ring coordinates every position carries one pool of every size, so A^T A is CIRCULANT:
//     eigenvalues lambda_k = w + sum_s (sin(pi k s / n) / sin(pi k / n))^2, solved exactly by one DFT pair per
//     cluster (O(n^2) fp64 with a twiddle table; n <= 3 000).  lambda_min ~ 3, so the normal equations lose
//     nothing (checked against numpy lstsq: 3e-12).
//   * Step 1: RED.ADD.U64 of fixed-point y into a [clusters][genes] table (order-independent, exact), then an
//     ordered compaction of the kept genes per cluster (gene -> slot map, fp64 ave, fp32 reciprocal).
// Single context only: the clusters are arbitrary cell subsets, so the matrix has to be resident on one GPU.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <thread>

#include "cb200_common.cuh"

#define POOL_THREADS 256
#define POOL_NB 1024
#define POOL_CAP 256
#define POOL_BYTES_PER_GENE 14   // window int64 + reciprocal fp32 + bucket id uint16
#define POOL_LOWWEIGHT 1e-6

struct PoolItem {
    int32_t cluster, size, p0, p1;
    int64_t out_off;   // index of ring position 0 of this (cluster, size) in the output array
};

// library sizes: warp per cell (counts are integers in fp64: the sum is exact whatever the order)
__global__ void __launch_bounds__(256)
k_pool_colsum(const int64_t *__restrict__ colptr, const double *__restrict__ x, int64_t n_cols, double *__restrict__ lib)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        double s = 0.0;
        for (int64_t i = colptr[c] + lane; i < colptr[c + 1]; i += 32)
            s += ld_stream(x + i);
        s = warp_sum(s);
        if (lane == 0)
            lib[c] = s;
    }
}

// ---------------------------------------------------------------------------------------
// step 1: per-cluster gene sums of x / lib in fixed point, then the kept genes of every cluster
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pool_gene_sums(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, const double *__restrict__ x,
                 const double *__restrict__ lib, const int32_t *__restrict__ cl_of, int64_t n_cols, int64_t G, double scale,
                 unsigned long long *__restrict__ acc, int *__restrict__ flag)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        const double r = scale / lib[c];
        unsigned long long *row = acc + (int64_t)cl_of[c] * G;
        for (int64_t i = b + lane; i < e; i += 32) {
            const double v = ld_stream(x + i);
            const int g = ld_stream_i(rowidx + i);
            if (g < 0 || g >= G || !(v >= 0.0)) {
                atomicOr(flag, 1);
                continue;
            }
            atomicAdd(row + g, (unsigned long long)__double2ll_rn(v * r));
        }
    }
}

// one CTA per cluster: ave = (sum / n) * meanlib; ordered compaction of the genes with ave >= min_mean
__global__ void __launch_bounds__(256)
k_pool_filter(const unsigned long long *__restrict__ acc, const int64_t *__restrict__ cl_ptr, const double *__restrict__ meanlib,
              int64_t G, double inv_scale1, double inv_scale2, double min_mean, int32_t *__restrict__ pslot,
              double *__restrict__ pave, float *__restrict__ pinv, double *__restrict__ ave_full, int32_t *__restrict__ gf_out)
{
    const int c = blockIdx.x;
    const double nc = (double)(cl_ptr[c + 1] - cl_ptr[c]);
    const double ml = meanlib[c];
    __shared__ int wcnt[8];
    __shared__ int base;
    if (threadIdx.x == 0)
        base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t g0 = 0; g0 < G; g0 += blockDim.x) {
        const int64_t g = g0 + threadIdx.x;
        double ave = 0.0;
        bool keep = false;
        if (g < G) {
            ave = ((double)acc[(int64_t)c * G + g] * inv_scale1 / nc) * ml;
            ave_full[(int64_t)c * G + g] = ave;
            keep = ave >= min_mean;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0)
            wcnt[w] = __popc(m);
        __syncthreads();
        int off = 0, tot = 0;
        for (int q = 0; q < 8; ++q) {
            if (q < w)
                off += wcnt[q];
            tot += wcnt[q];
        }
        if (g < G) {
            int slot = -1;
            if (keep) {
                slot = base + off + __popc(m & ((1u << lane) - 1u));
                pave[(int64_t)c * G + slot] = ave;
                pinv[(int64_t)c * G + slot] = (float)(inv_scale2 / ave);
            }
            pslot[(int64_t)c * G + g] = slot;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        gf_out[c] = base;
}

// ---------------------------------------------------------------------------------------
// step 3: pool medians
// ---------------------------------------------------------------------------------------
struct PoolShared {
    unsigned int hist[POOL_NB];
    double list_v[POOL_CAP];
    int list_s[POOL_CAP];
    unsigned long long next_min;   // bit pattern of the smallest exact ratio beyond the gathered bucket
    double res[2];
    int wsum[POOL_THREADS / 32];
    int n_zero, n_below, n_list, n_in;
    int bucket, cum, cnt, tie_slot;
};
#define POOL_BKT_NONE 0xffffu    // zero entry, or below the range
#define POOL_BKT_ABOVE 0xfffeu   // beyond the range

// Compact store of the kept entries, in RING order: position r of the ring (clusters concatenated) owns the entries
// [cptr[r], cptr[r + 1]) = (slot of the gene among its cluster's kept genes, fixed-point y).  The cells of a window are
// consecutive ring positions, so a window is one contiguous range (two at the wrap-around) and a slide touches two
// short contiguous ranges: no dependent gene -> slot look-ups on the way.
__global__ void __launch_bounds__(256)
k_pool_compact(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, const double *__restrict__ x,
               const double *__restrict__ lib, const int32_t *__restrict__ ring, const int32_t *__restrict__ cl_of,
               const int32_t *__restrict__ pslot, int64_t n_cols, int64_t G, double scale, const int64_t *__restrict__ cptr,
               int32_t *__restrict__ kcnt, uint16_t *__restrict__ eslot, long long *__restrict__ eq)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp0; r < n_cols; r += nwarps) {
        const int cell = ring[r];
        const int64_t b = colptr[cell], e = colptr[cell + 1];
        const int32_t *slot_of = pslot + (int64_t)cl_of[cell] * G;
        const double sc = scale / lib[cell];
        int64_t out = cptr != nullptr ? cptr[r] : 0;
        int total = 0;
        for (int64_t i0 = b; i0 < e; i0 += 32) {
            const int64_t i = i0 + lane;
            int sl = -1;
            if (i < e)
                sl = __ldg(slot_of + ld_stream_i(rowidx + i));
            const unsigned m = __ballot_sync(0xffffffffu, sl >= 0);
            if (cptr != nullptr && sl >= 0) {
                const int64_t o = out + __popc(m & ((1u << lane) - 1u));
                eslot[o] = (uint16_t)sl;
                eq[o] = __double2ll_rn(ld_stream(x + i) * sc);
            }
            out += __popc(m);
            total += __popc(m);
        }
        if (cptr == nullptr && lane == 0)
            kcnt[r] = total;
    }
}

// all threads of the CTA add (sign = +1) or subtract (sign = -1) the entries [b, e) of the compact store
__device__ __forceinline__ void pool_apply_range(const uint16_t *__restrict__ eslot, const long long *__restrict__ eq, int64_t b,
                                                 int64_t e, int sign, unsigned long long *W, int tid)
{
    int64_t i = b + tid;
    for (; i + 3 * POOL_THREADS < e; i += 4 * POOL_THREADS) {   // four independent loads in flight per thread
        long long q[4];
        uint16_t sl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            q[u] = __ldg(eq + i + u * POOL_THREADS);
            sl[u] = __ldg(eslot + i + u * POOL_THREADS);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            atomicAdd(W + sl[u], (unsigned long long)(sign > 0 ? q[u] : -q[u]));
    }
    for (; i < e; i += POOL_THREADS) {
        const long long q = __ldg(eq + i);
        atomicAdd(W + __ldg(eslot + i), (unsigned long long)(sign > 0 ? q : -q));
    }
}

__device__ __forceinline__ float pool_key(unsigned long long w, float inv)
{
    return (float)(long long)w * inv;
}

// bucket code of one window entry for the range [lo, hi); counts it and, inside the range, adds it to the histogram
__device__ __forceinline__ unsigned int pool_classify(unsigned long long w, float iv, float lo, float hi, float scale,
                                                      unsigned int *hist, int &zero, int &below, int &in)
{
    const float key = pool_key(w, iv);
    const bool z = w == 0ULL;
    const bool bl = !z && key < lo;
    const bool inr = !z && !bl && key < hi;
    int bb = (int)((key - lo) * scale);   // only used inside the range
    bb = bb > POOL_NB - 1 ? POOL_NB - 1 : bb;
    bb = bb < 0 ? 0 : bb;
    if (inr)
        atomicAdd(hist + bb, 1u);
    zero += z ? 1 : 0;
    below += bl ? 1 : 0;
    in += inr ? 1 : 0;
    return inr ? (unsigned int)bb : ((z || bl) ? POOL_BKT_NONE : POOL_BKT_ABOVE);
}

// Median of the gf ratios W[g] * inv[g] of the window (zeros included), exact in fp64 up to the bucket-edge caveat in
// the header.  All threads of the CTA must call it; every thread returns the same value.  `guess` carries the scale
// of the previous median from one ring position to the next.
__device__ double pool_median_select(PoolShared *sh, const unsigned long long *W, const float *inv, uint16_t *bkt,
                                     const double *__restrict__ ave, int gf, float &guess, double inv_scale2, int *__restrict__ flag)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = gf / 2;
    const bool even = (gf & 1) == 0;
    float lo = 0.f, hi = 4.f * guess;
    if (!(hi > 0.f) || isinf(hi))
        hi = 1.f;
    for (int round = 0; round < 96; ++round) {
        {
            uint4 *h4 = reinterpret_cast<uint4 *>(sh->hist);
            for (int b = tid; b < POOL_NB / 4; b += POOL_THREADS)
                h4[b] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (tid == 0) {
            sh->n_zero = 0;
            sh->n_below = 0;
            sh->n_in = 0;
            sh->n_list = 0;
            sh->tie_slot = 0x7fffffff;
            sh->next_min = ~0ULL;
        }
        __syncthreads();
        // pass A: zeros, entries below the range, histogram of the entries inside it; the bucket of every gene is kept
        const float scale = (float)POOL_NB / (hi - lo);
        int zero = 0, below = 0, in = 0;
        {
            // two entries per thread and step: one 128-bit + one 64-bit shared load, one 32-bit store of both bucket ids
            const ulonglong2 *W2 = reinterpret_cast<const ulonglong2 *>(W);
            const float2 *inv2 = reinterpret_cast<const float2 *>(inv);
            unsigned int *bkt2 = reinterpret_cast<unsigned int *>(bkt);
            const int npair = gf >> 1;
#pragma unroll 2
            for (int i = tid; i < npair; i += POOL_THREADS) {
                const ulonglong2 w = W2[i];
                const float2 iv = inv2[i];
                const unsigned int b0 = pool_classify(w.x, iv.x, lo, hi, scale, sh->hist, zero, below, in);
                const unsigned int b1 = pool_classify(w.y, iv.y, lo, hi, scale, sh->hist, zero, below, in);
                bkt2[i] = b0 | (b1 << 16);
            }
            if ((gf & 1) && tid == 0)
                bkt[gf - 1] = (uint16_t)pool_classify(W[gf - 1], inv[gf - 1], lo, hi, scale, sh->hist, zero, below, in);
        }
        zero = warp_sum_i(zero);
        below = warp_sum_i(below);
        in = warp_sum_i(in);
        if (lane == 0) {
            if (zero)
                atomicAdd(&sh->n_zero, zero);
            if (below)
                atomicAdd(&sh->n_below, below);
            if (in)
                atomicAdd(&sh->n_in, in);
        }
        __syncthreads();
        const int nz = sh->n_zero, n_below = sh->n_below, n_in = sh->n_in;
        // the wanted rank(s) among the non-zero ratios
        if (half < nz) {
            __syncthreads();   // counters read by everybody before the next call resets them
            return 0.0;
        }
        const bool lower_is_zero = even && half - 1 < nz;                     // lower middle is a zero
        const int r = even ? (lower_is_zero ? 0 : half - 1 - nz) : half - nz;   // first wanted non-zero rank
        const bool want_next = even && !lower_is_zero;
        if (r < n_below) {   // the range starts too high (only after a narrowing step met a rounding edge)
            const float wd = hi - lo;
            hi = lo;
            lo = fmaxf(0.f, lo - 4.f * wd);
            __syncthreads();
            continue;
        }
        if (r >= n_below + n_in) {   // beyond the range: widen upwards
            const float wd = hi - lo;
            lo = hi;
            hi = hi + 16.f * wd;
            if (isinf(hi))
                hi = 3.0e38f;
            __syncthreads();
            continue;
        }
        // bucket of the wanted rank: block-wide scan of the histogram, POOL_NB / POOL_THREADS buckets per thread
        constexpr int BPT = POOL_NB / POOL_THREADS;
        unsigned int hv[BPT];
        {
            const uint4 *h4 = reinterpret_cast<const uint4 *>(sh->hist + tid * BPT);
#pragma unroll
            for (int q = 0; q < BPT / 4; ++q) {
                const uint4 v = h4[q];
                hv[4 * q] = v.x;
                hv[4 * q + 1] = v.y;
                hv[4 * q + 2] = v.z;
                hv[4 * q + 3] = v.w;
            }
        }
        int mine = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q)
            mine += (int)hv[q];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            sh->wsum[warp] = incl;
        __syncthreads();
        int excl = incl - mine;
        for (int q = 0; q < warp; ++q)
            excl += sh->wsum[q];
        const int want_in = r - n_below;
        if (want_in >= excl && want_in < excl + mine) {
            int cum = excl;
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                if (want_in >= cum && want_in < cum + (int)hv[q]) {
                    sh->bucket = tid * BPT + q;
                    sh->cum = cum;
                    sh->cnt = (int)hv[q];
                }
                cum += (int)hv[q];
            }
        }
        __syncthreads();
        const int bucket = sh->bucket, cnt = sh->cnt;
        const int want = want_in - sh->cum;   // rank inside the bucket
        const float blo = lo + (float)bucket / scale, bhi = lo + (float)(bucket + 1) / scale;
        const bool degenerate = !(bhi > blo);   // the bucket cannot be split further in fp32: its keys are one value
        if (cnt > POOL_CAP && !degenerate && round < 90) {   // narrow to this bucket (membership is recounted next round)
            lo = blo;
            hi = bhi;
            __syncthreads();
            continue;
        }
        // gather the bucket; beyond it, the smallest exact ratio (the "next" element when the bucket ends at the wanted rank)
        const bool need_outside = want_next && want == cnt - 1;
        {
            // eight bucket ids per thread and step (the ids beyond gf are POOL_BKT_NONE)
            const uint4 *b4 = reinterpret_cast<const uint4 *>(bkt);
            const int n8 = (gf + 7) >> 3;
            for (int i = tid; i < n8; i += POOL_THREADS) {
                const uint4 v = b4[i];
                const unsigned int word[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const unsigned int b = (word[h >> 1] >> (16 * (h & 1))) & 0xffffu;
                    const int g = 8 * i + h;
                    if (b == (unsigned int)bucket) {
                        if (cnt <= POOL_CAP) {
                            const int pos = atomicAdd(&sh->n_list, 1);
                            if (pos < POOL_CAP) {
                                sh->list_s[pos] = g;
                                sh->list_v[pos] = ((double)(long long)W[g] * inv_scale2) / ave[g];
                            }
                        } else {
                            atomicMin(&sh->tie_slot, g);   // more equal keys than the list holds: they are one value
                        }
                    } else if (need_outside && b > (unsigned int)bucket && b != POOL_BKT_NONE) {
                        const double vv = ((double)(long long)W[g] * inv_scale2) / ave[g];
                        atomicMin(&sh->next_min, (unsigned long long)__double_as_longlong(vv));
                    }
                }
            }
        }
        __syncthreads();
        double v0, v1;
        if (cnt > POOL_CAP) {
            const int g = sh->tie_slot;
            v0 = v1 = ((double)(long long)W[g] * inv_scale2) / ave[g];
        } else {
            const int nl = sh->n_list < POOL_CAP ? sh->n_list : POOL_CAP;
            if (tid < nl) {
                const double v = sh->list_v[tid];
                const int sl = sh->list_s[tid];
                int rank = 0;
                for (int j = 0; j < nl; ++j) {
                    const double u = sh->list_v[j];
                    rank += (u < v || (u == v && sh->list_s[j] < sl)) ? 1 : 0;
                }
                if (rank == want)
                    sh->res[0] = v;
                if (want_next && rank == want + 1)
                    sh->res[1] = v;
            }
            if (tid == 0 && sh->n_list != cnt)
                atomicOr(flag, 8);   // histogram and gather disagree: internal error
            __syncthreads();
            v0 = sh->res[0];
            v1 = need_outside ? __longlong_as_double((long long)sh->next_min) : sh->res[1];
        }
        guess = (float)v0;
        __syncthreads();   // res / counters consumed before the next call rewrites them
        if (!even)
            return v0;
        return lower_is_zero ? (0.0 + v0) / 2.0 : (v1 + v0) / 2.0;
    }
    if (tid == 0)
        atomicOr(flag, 4);   // no convergence of the bucket search (not expected)
    __syncthreads();
    return 0.0;
}

#define POOL_PREF 4   // entries per thread and column prefetched into registers ahead of a slide

__global__ void __launch_bounds__(POOL_THREADS, 4)
k_pool_medians(const PoolItem *__restrict__ items, int n_items, const int64_t *__restrict__ cptr,
               const uint16_t *__restrict__ eslot, const long long *__restrict__ eq, const int64_t *__restrict__ cl_ptr,
               const float *__restrict__ pinv, const double *__restrict__ pave, const int32_t *__restrict__ gf_of,
               const double *__restrict__ meanlib, int64_t G, int gf_max, double inv_scale2, double *__restrict__ bvals,
               int *__restrict__ flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int gfp = (gf_max + 7) & ~7;   // array strides: multiples of 8 genes keep every array 16-byte aligned
    unsigned long long *W = reinterpret_cast<unsigned long long *>(smem_raw);
    float *inv = reinterpret_cast<float *>(W + gfp);
    uint16_t *bkt = reinterpret_cast<uint16_t *>(inv + gfp);
    PoolShared *sh = reinterpret_cast<PoolShared *>(smem_raw + (size_t)gfp * POOL_BYTES_PER_GENE);
    const int tid = threadIdx.x;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const PoolItem item = items[it];
        const int c = item.cluster, size = item.size;
        const int64_t base = cl_ptr[c];
        const int n = (int)(cl_ptr[c + 1] - base);
        const int gf = gf_of[c];
        const double *ave = pave + (int64_t)c * G;
        const int64_t *cp = cptr + base;   // cp[r] .. cp[r + 1]: entries of ring position r of this cluster
        __syncthreads();   // previous item fully done with W / sh
        for (int g = tid; g < gf; g += POOL_THREADS) {
            W[g] = 0ULL;
            inv[g] = pinv[(int64_t)c * G + g];
        }
        if ((gf & 7) != 0 && tid < 8 && (gf & ~7) + tid >= gf)
            bkt[(gf & ~7) + tid] = (uint16_t)POOL_BKT_NONE;   // ids beyond gf in the last group of 8
        __syncthreads();
        // window of the first position: ring positions p0 .. p0 + size - 1 (mod n) = one or two contiguous ranges
        int po = item.p0 % n;             // ring position that leaves at the next slide ...
        int pi = (item.p0 + size) % n;    // ... and the one that enters
        {
            const int first = po + size <= n ? size : n - po;
            pool_apply_range(eslot, eq, cp[po], cp[po + first], +1, W, tid);
            if (first < size)
                pool_apply_range(eslot, eq, cp[0], cp[size - first], +1, W, tid);
        }
        __syncthreads();
        float guess = (float)(0.5 * (double)size / meanlib[c]);
        for (int p = item.p0; p < item.p1; ++p) {
            // the two columns of the coming slide are requested now and applied after the median
            const bool slide = p + 1 < item.p1 && size < n;
            int64_t ob = 0, oe = 0, ib = 0, ie = 0;
            uint16_t os[POOL_PREF], is[POOL_PREF];
            long long oq[POOL_PREF], iq[POOL_PREF];
            if (slide) {
                ob = cp[po];
                oe = cp[po + 1];
                ib = cp[pi];
                ie = cp[pi + 1];
#pragma unroll
                for (int u = 0; u < POOL_PREF; ++u) {
                    const int64_t i = ob + tid + u * POOL_THREADS, j = ib + tid + u * POOL_THREADS;
                    os[u] = i < oe ? __ldg(eslot + i) : (uint16_t)0;
                    oq[u] = i < oe ? __ldg(eq + i) : 0LL;
                    is[u] = j < ie ? __ldg(eslot + j) : (uint16_t)0;
                    iq[u] = j < ie ? __ldg(eq + j) : 0LL;
                }
            }
            const double med = pool_median_select(sh, W, inv, bkt, ave, gf, guess, inv_scale2, flag);
            if (tid == 0)
                bvals[item.out_off + p] = med;
            if (slide) {
#pragma unroll
                for (int u = 0; u < POOL_PREF; ++u) {
                    if (oq[u] != 0LL)
                        atomicAdd(W + os[u], (unsigned long long)(-oq[u]));
                    if (iq[u] != 0LL)
                        atomicAdd(W + is[u], (unsigned long long)iq[u]);
                }
                // columns longer than the prefetch window
                pool_apply_range(eslot, eq, ob + (int64_t)POOL_PREF * POOL_THREADS, oe, -1, W, tid);
                pool_apply_range(eslot, eq, ib + (int64_t)POOL_PREF * POOL_THREADS, ie, +1, W, tid);
                po = po + 1 == n ? 0 : po + 1;
                pi = pi + 1 == n ? 0 : pi + 1;
                __syncthreads();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// step 4: circulant least squares per cluster (one CTA each)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
k_pool_solve(const int64_t *__restrict__ cl_ptr, const int32_t *__restrict__ sizes, const int32_t *__restrict__ ns_of,
             int n_sizes_all, int64_t N, const double *__restrict__ bvals, const int32_t *__restrict__ ring,
             const double *__restrict__ lib, double *__restrict__ rhs, double *__restrict__ tw_re, double *__restrict__ tw_im,
             double *__restrict__ F_re, double *__restrict__ F_im, double *__restrict__ nf)
{
    const int c = blockIdx.x;
    const int64_t base = cl_ptr[c];
    const int n = (int)(cl_ptr[c + 1] - base);
    const int ns = ns_of[c];
    const double w = POOL_LOWWEIGHT;
    // right-hand side A^T b in ring coordinates and the twiddle table
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double acc = w * bvals[(int64_t)n_sizes_all * N + base + j];
        for (int si = 0; si < ns; ++si) {
            const int s = sizes[si];
            const double *b = bvals + (int64_t)si * N + base;
            int idx = j;
            for (int t = 0; t < s; ++t) {
                acc += b[idx];
                idx = idx == 0 ? n - 1 : idx - 1;
            }
        }
        rhs[base + j] = acc;
        double sn, cs;
        sincospi(2.0 * (double)j / (double)n, &sn, &cs);
        tw_re[base + j] = cs;
        tw_im[base + j] = -sn;
    }
    __syncthreads();
    // forward DFT, divided by the eigenvalues of A^T A
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double fr = 0.0, fi = 0.0;
        int idx = 0;
        for (int j = 0; j < n; ++j) {
            const double v = rhs[base + j];
            fr += v * tw_re[base + idx];
            fi += v * tw_im[base + idx];
            idx += k;
            if (idx >= n)
                idx -= n;
        }
        double lam = w;
        for (int si = 0; si < ns; ++si) {
            const int s = sizes[si];
            if (k == 0) {
                lam += (double)s * (double)s;
            } else {
                const double a = sinpi((double)(((int64_t)k * s) % n) / (double)n);
                const double b = sinpi((double)k / (double)n);
                lam += (a / b) * (a / b);
            }
        }
        F_re[base + k] = fr / lam;
        F_im[base + k] = fi / lam;
    }
    __syncthreads();
    // inverse DFT (real part), times the library size, scattered to the cell
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double acc = 0.0;
        int idx = 0;
        for (int k = 0; k < n; ++k) {
            acc += F_re[base + k] * tw_re[base + idx] + F_im[base + k] * tw_im[base + idx];
            idx += j;
            if (idx >= n)
                idx -= n;
        }
        const int cell = ring[base + j];
        nf[cell] = (acc / (double)n) * lib[cell];
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <typename F>
static void pool_parallel_for(int n, F f)
{
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw == 0 ? 4 : hw > 32 ? 32 : hw);
    if (nt > n)
        nt = n;
    if (nt <= 1) {
        for (int i = 0; i < n; ++i)
            f(i);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=]() {
            for (int i = t; i < n; i += nt)
                f(i);
        });
    for (auto &t : th)
        t.join();
}

static double pool_median(std::vector<double> &v)
{
    const size_t n = v.size();
    if (n == 0)
        return NAN;
    const size_t h = n / 2;
    std::nth_element(v.begin(), v.begin() + h, v.end());
    const double hi = v[h];
    if (n & 1)
        return hi;
    const double lo = *std::max_element(v.begin(), v.begin() + h);
    return (lo + hi) / 2.0;
}

// Host layout of the pooling (scuttle:::.limit_cluster_size + .generateSphere): clusters above max_cluster_size are dealt
// round-robin into ceil(n / max) parts, new ids in order of first appearance; per cluster the member cells (ascending),
// the ring order by library size (stable sort; odd ranks ascending, then even ranks descending) and the mean library
// size.  Returns false when a cluster code is negative.
static bool pool_layout(int64_t N, const int32_t *clusters, const double *lib, int64_t max_cluster_size, std::vector<int32_t> &cl,
                        int &C, std::vector<int64_t> &cl_ptr, std::vector<int32_t> &members, std::vector<int32_t> &ring,
                        std::vector<double> &meanlib)
{
    cl.assign((size_t)N, 0);
    C = 0;
    {
        int32_t maxc = 0;
        if (clusters != nullptr)
            for (int64_t i = 0; i < N; ++i) {
                if (clusters[i] < 0)
                    return false;
                maxc = clusters[i] > maxc ? clusters[i] : maxc;
            }
        std::vector<int64_t> cnt((size_t)maxc + 1, 0), seen((size_t)maxc + 1, 0);
        std::vector<int32_t> first_id((size_t)maxc + 1, -1), mult((size_t)maxc + 1, 1);
        for (int64_t i = 0; i < N; ++i)
            cnt[(size_t)(clusters ? clusters[i] : 0)] += 1;
        for (int64_t i = 0; i < N; ++i) {
            const size_t o = (size_t)(clusters ? clusters[i] : 0);
            if (first_id[o] < 0) {
                first_id[o] = C;
                mult[o] = (max_cluster_size > 0 && cnt[o] > max_cluster_size) ? (int32_t)((cnt[o] + max_cluster_size - 1) / max_cluster_size) : 1;
                C += mult[o];
            }
            cl[(size_t)i] = first_id[o] + (int32_t)(seen[o] % mult[o]);
            seen[o] += 1;
        }
    }
    cl_ptr.assign((size_t)C + 1, 0);
    for (int64_t i = 0; i < N; ++i)
        cl_ptr[(size_t)cl[(size_t)i] + 1] += 1;
    for (int c = 0; c < C; ++c)
        cl_ptr[(size_t)c + 1] += cl_ptr[(size_t)c];
    members.assign((size_t)N, 0);
    ring.assign((size_t)N, 0);
    meanlib.assign((size_t)C, 0.0);
    {
        std::vector<int64_t> cur(cl_ptr.begin(), cl_ptr.end() - 1);
        for (int64_t i = 0; i < N; ++i)
            members[(size_t)cur[(size_t)cl[(size_t)i]]++] = (int32_t)i;
    }
    pool_parallel_for(C, [&](int c) {
        const int64_t b = cl_ptr[(size_t)c], e = cl_ptr[(size_t)c + 1];
        std::vector<int32_t> o(members.begin() + b, members.begin() + e);
        std::stable_sort(o.begin(), o.end(), [&](int32_t a, int32_t bb) { return lib[(size_t)a] < lib[(size_t)bb]; });
        const int64_t n = e - b;
        int64_t w = b;
        for (int64_t i = 0; i < n; i += 2)
            ring[(size_t)w++] = o[(size_t)i];
        for (int64_t i = (n & 1) ? n - 2 : n - 1; i >= 1; i -= 2)
            ring[(size_t)w++] = o[(size_t)i];
        double s = 0.0;
        for (int64_t i = b; i < e; ++i)
            s += lib[(size_t)members[(size_t)i]];
        meanlib[(size_t)c] = s / (double)n;
    });
    return true;
}

// the same layout for host-side tests (no device work): cl_out[n], ring_out[n], cl_ptr_out[n + 1] (first n_clusters + 1 used)
extern "C" int cb200_host_pool_layout(int64_t n, const int32_t *clusters, const double *lib, int64_t max_cluster_size,
                                      int32_t *cl_out, int32_t *ring_out, int64_t *cl_ptr_out, double *meanlib_out,
                                      int32_t *n_clusters_out)
{
    if (n < 1 || lib == nullptr || cl_out == nullptr || ring_out == nullptr || cl_ptr_out == nullptr || n_clusters_out == nullptr)
        return 1;
    std::vector<int32_t> cl, members, ring;
    std::vector<int64_t> cl_ptr;
    std::vector<double> meanlib;
    int C = 0;
    if (!pool_layout(n, clusters, lib, max_cluster_size, cl, C, cl_ptr, members, ring, meanlib))
        return 2;
    std::copy(cl.begin(), cl.end(), cl_out);
    std::copy(ring.begin(), ring.end(), ring_out);
    std::copy(cl_ptr.begin(), cl_ptr.end(), cl_ptr_out);
    if (meanlib_out != nullptr)
        std::copy(meanlib.begin(), meanlib.end(), meanlib_out);
    *n_clusters_out = C;
    return 0;
}

extern "C" int cb200_pooled_size_factors(cb200_ctx *ctx, const int32_t *clusters, const int32_t *sizes, int n_sizes,
                                         double min_mean, int64_t max_cluster_size, double *sf_out, int64_t *n_nonpositive,
                                         double *stage_ms_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded, "cb200_pooled_size_factors: no count matrix loaded");
    CB_CHECK(ctx, ctx->N == ctx->N_total, "cb200_pooled_size_factors: the whole matrix has to be resident on one GPU "
                                           "(clusters are arbitrary cell subsets)");
    CB_CHECK(ctx, sf_out != nullptr, "cb200_pooled_size_factors: NULL output");
    static const int32_t default_sizes[17] = {21, 26, 31, 36, 41, 46, 51, 56, 61, 66, 71, 76, 81, 86, 91, 96, 101};
    if (sizes == nullptr) {
        sizes = default_sizes;
        n_sizes = 17;
    }
    CB_CHECK(ctx, n_sizes >= 1 && n_sizes <= 64, "cb200_pooled_size_factors: 1 <= number of pool sizes <= 64");
    std::vector<int32_t> sz(sizes, sizes + n_sizes);
    std::sort(sz.begin(), sz.end());
    for (int i = 0; i < n_sizes; ++i) {
        CB_CHECK(ctx, sz[i] >= 2, "cb200_pooled_size_factors: pool sizes must be >= 2");
        CB_CHECK(ctx, i == 0 || sz[i] != sz[i - 1], "cb200_pooled_size_factors: 'sizes' are not unique");
    }
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t N = ctx->N, G = ctx->G;
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    cudaEvent_t ev[4];
    for (auto &e : ev)
        CB_CUDA(ctx, cudaEventCreate(&e));
    struct EvGuard {
        cudaEvent_t *e;
        ~EvGuard()
        {
            for (int i = 0; i < 4; ++i)
                cudaEventDestroy(e[i]);
        }
    } ev_guard{ev};

    // ---- library sizes (K1) -> host
    CB_CUDA(ctx, ctx->blk_w.ensure((size_t)N));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
    {
        const int grid = cb_grid_for(N, 8, ctx->num_sms * 8);
        k_pool_colsum<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->x.p, N, ctx->blk_w.p);
        CB_LAUNCH(ctx);
    }
    std::vector<double> lib((size_t)N);
    CB_TRY(cb200_d2h(ctx, lib.data(), ctx->blk_w.p, (size_t)N * sizeof(double), ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t i = 0; i < N; ++i)
        CB_CHECK(ctx, lib[(size_t)i] > 0.0 && std::isfinite(lib[(size_t)i]), "cb200_pooled_size_factors: cells should have non-zero library sizes");
    if (!(min_mean > 0.0)) {   // scuttle:::.guessMinMean
        std::vector<double> tmp(lib);
        min_mean = pool_median(tmp) <= 50000.0 ? 0.1 : 1.0;
    }
    if (min_mean < 1e-8)
        min_mean = 1e-8;

    // ---- clusters: size limit (round-robin split, ids by first appearance), members, ring order (host: pool_layout)
    std::vector<int32_t> cl, members, ring;
    std::vector<int64_t> cl_ptr;
    std::vector<double> meanlib;
    int C = 0;
    CB_CHECK(ctx, pool_layout(N, clusters, lib.data(), max_cluster_size, cl, C, cl_ptr, members, ring, meanlib),
             "cb200_pooled_size_factors: cluster codes must be >= 0");
    std::vector<int32_t> ns_of((size_t)C);
    int n_max = 0;
    for (int c = 0; c < C; ++c) {
        const int64_t n = cl_ptr[(size_t)c + 1] - cl_ptr[(size_t)c];
        int ns = 0;
        while (ns < n_sizes && sz[(size_t)ns] <= n)
            ++ns;
        CB_CHECK(ctx, ns >= 1, "cb200_pooled_size_factors: not enough cells in at least one cluster for some 'sizes' "
                               "(%lld cells, smallest pool %d)", (long long)n, sz[0]);
        CB_CHECK(ctx, n < (1 << 30), "cb200_pooled_size_factors: cluster too large");
        ns_of[(size_t)c] = ns;
        n_max = n > n_max ? (int)n : n_max;
    }
    // ---- device buffers
    DevBuf<int32_t> d_cl, d_ring, d_sizes, d_ns, d_pslot, d_gf;
    DevBuf<int64_t> d_clptr;
    DevBuf<double> d_meanlib, d_pave, d_avefull, d_bvals, d_work, d_nf;
    DevBuf<float> d_pinv;
    DevBuf<unsigned long long> d_acc;
    DevBuf<PoolItem> d_items;
    DevBuf<int32_t> d_kcnt;          // compact ring-ordered store of the kept entries (step 3)
    DevBuf<int64_t> d_cptr;
    DevBuf<uint16_t> d_eslot;
    DevBuf<long long> d_eq;
    struct BufGuard {   // declared after every buffer it releases
        std::vector<std::function<void()>> f;
        ~BufGuard()
        {
            for (auto &g : f)
                g();
        }
    } guard;
#define POOL_OWN(b) guard.f.push_back([&]() { b.release(); })
    POOL_OWN(d_cl); POOL_OWN(d_ring); POOL_OWN(d_sizes); POOL_OWN(d_ns); POOL_OWN(d_pslot); POOL_OWN(d_gf); POOL_OWN(d_clptr);
    POOL_OWN(d_meanlib); POOL_OWN(d_pave); POOL_OWN(d_avefull); POOL_OWN(d_bvals); POOL_OWN(d_work); POOL_OWN(d_nf);
    POOL_OWN(d_pinv); POOL_OWN(d_acc); POOL_OWN(d_items);
    POOL_OWN(d_kcnt); POOL_OWN(d_cptr); POOL_OWN(d_eslot); POOL_OWN(d_eq);
#undef POOL_OWN
    const size_t CG = (size_t)C * (size_t)G;
    CB_CUDA(ctx, d_cl.ensure((size_t)N));
    CB_CUDA(ctx, d_ring.ensure((size_t)N));
    CB_CUDA(ctx, d_sizes.ensure((size_t)n_sizes));
    CB_CUDA(ctx, d_ns.ensure((size_t)C));
    CB_CUDA(ctx, d_clptr.ensure((size_t)C + 1));
    CB_CUDA(ctx, d_meanlib.ensure((size_t)C));
    CB_CUDA(ctx, d_acc.ensure(CG));
    CB_CUDA(ctx, d_pslot.ensure(CG));
    CB_CUDA(ctx, d_pave.ensure(CG));
    CB_CUDA(ctx, d_pinv.ensure(CG));
    CB_CUDA(ctx, d_avefull.ensure(CG));
    CB_CUDA(ctx, d_gf.ensure((size_t)C));
    CB_CUDA(ctx, d_bvals.ensure((size_t)(n_sizes + 1) * N));
    CB_CUDA(ctx, d_work.ensure((size_t)5 * N));
    CB_CUDA(ctx, d_nf.ensure((size_t)N));
    CB_TRY(cb200_h2d(ctx, d_cl.p, cl.data(), (size_t)N * sizeof(int32_t), ctx->stream));
    CB_TRY(cb200_h2d(ctx, d_ring.p, ring.data(), (size_t)N * sizeof(int32_t), ctx->stream));
    CB_TRY(cb200_h2d(ctx, d_sizes.p, sz.data(), (size_t)n_sizes * sizeof(int32_t), ctx->stream));
    CB_TRY(cb200_h2d(ctx, d_ns.p, ns_of.data(), (size_t)C * sizeof(int32_t), ctx->stream));
    CB_TRY(cb200_h2d(ctx, d_clptr.p, cl_ptr.data(), (size_t)(C + 1) * sizeof(int64_t), ctx->stream));
    CB_TRY(cb200_h2d(ctx, d_meanlib.p, meanlib.data(), (size_t)C * sizeof(double), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(d_acc.p, 0, CG * sizeof(unsigned long long), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(d_bvals.p, 0, (size_t)(n_sizes + 1) * N * sizeof(double), ctx->stream));
    CB_TRY(cb200_need_rowidx(ctx));

    // fixed-point scales: y <= 1; sums over at most n_max cells (pseudo-cell) / the largest pool (windows)
    int bits1 = 1, bits2 = 1;
    while ((1LL << bits1) <= (long long)n_max)
        ++bits1;
    while ((1LL << bits2) <= (long long)sz[(size_t)n_sizes - 1])
        ++bits2;
    const int s1 = 62 - bits1, s2 = 62 - bits2;

    // ---- step 1
    CB_CUDA(ctx, cudaEventRecord(ev[0], ctx->stream));
    {
        const int grid = cb_grid_for(N, 8, ctx->num_sms * 8);
        k_pool_gene_sums<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->rowidx.p, ctx->x.p, ctx->blk_w.p, d_cl.p, N, G,
                                                        ldexp(1.0, s1), d_acc.p, ctx->flag.p);
        CB_LAUNCH(ctx);
        k_pool_filter<<<C, 256, 0, ctx->stream>>>(d_acc.p, d_clptr.p, d_meanlib.p, G, ldexp(1.0, -s1), ldexp(1.0, -s2), min_mean,
                                                  d_pslot.p, d_pave.p, d_pinv.p, d_avefull.p, d_gf.p);
        CB_LAUNCH(ctx);
    }
    std::vector<int32_t> gf((size_t)C);
    CB_TRY(cb200_d2h(ctx, gf.data(), d_gf.p, (size_t)C * sizeof(int32_t), ctx->stream));
    int h_flag = 0;
    CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
    CB_CHECK(ctx, (h_flag & 1) == 0, "cb200_pooled_size_factors: a row index is outside [0, n_genes) or a count is negative / NaN");
    int gf_max = 0;
    for (int c = 0; c < C; ++c) {
        CB_CHECK(ctx, gf[(size_t)c] >= 1, "cb200_pooled_size_factors: no gene of cluster %d reaches min.mean = %g", c, min_mean);
        gf_max = gf[(size_t)c] > gf_max ? gf[(size_t)c] : gf_max;
    }
    const size_t smem = (size_t)((gf_max + 7) & ~7) * POOL_BYTES_PER_GENE + sizeof(PoolShared);
    CB_CHECK(ctx, smem <= 200 * 1024 && gf_max <= 65535, "cb200_pooled_size_factors: %d genes pass min.mean in one cluster; the shared-memory "
                                      "window holds at most %d (raise min.mean)", gf_max, (int)((200 * 1024 - sizeof(PoolShared)) / POOL_BYTES_PER_GENE));

    // ---- compact ring-ordered store of the kept entries
    CB_CUDA(ctx, d_kcnt.ensure((size_t)N));
    CB_CUDA(ctx, d_cptr.ensure((size_t)N + 1));
    int64_t n_kept = 0;
    {
        const int grid = cb_grid_for(N, 8, ctx->num_sms * 8);
        k_pool_compact<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->rowidx.p, ctx->x.p, ctx->blk_w.p, d_ring.p, d_cl.p,
                                                      d_pslot.p, N, G, ldexp(1.0, s2), nullptr, d_kcnt.p, nullptr, nullptr);
        CB_LAUNCH(ctx);
        CB_TRY(cb200_scan_i32_to_i64(ctx, d_kcnt.p, N, d_cptr.p));
        CB_TRY(cb200_read_back(ctx, &n_kept, d_cptr.p + N, sizeof(int64_t)));
        CB_CUDA(ctx, d_eslot.ensure((size_t)n_kept + 8));
        CB_CUDA(ctx, d_eq.ensure((size_t)n_kept + 8));
        k_pool_compact<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->rowidx.p, ctx->x.p, ctx->blk_w.p, d_ring.p, d_cl.p,
                                                      d_pslot.p, N, G, ldexp(1.0, s2), d_cptr.p, nullptr, d_eslot.p, d_eq.p);
        CB_LAUNCH(ctx);
    }

    // ---- step 3: work items = (cluster, size, run of ring positions)
    int64_t pairs = 0;
    for (int c = 0; c < C; ++c)
        pairs += (cl_ptr[(size_t)c + 1] - cl_ptr[(size_t)c]) * (ns_of[(size_t)c] + 1);
    int L = (int)(pairs / ((int64_t)ctx->num_sms * 24));
    L = L < 8 ? 8 : L > 128 ? 128 : L;
    std::vector<PoolItem> items;
    items.reserve((size_t)(pairs / L + (int64_t)C * (n_sizes + 1) + 16));
    for (int c = 0; c < C; ++c) {
        const int n = (int)(cl_ptr[(size_t)c + 1] - cl_ptr[(size_t)c]);
        for (int si = -1; si < ns_of[(size_t)c]; ++si) {
            const int size = si < 0 ? 1 : sz[(size_t)si];
            const int64_t off = (int64_t)(si < 0 ? n_sizes : si) * N + cl_ptr[(size_t)c];
            for (int p0 = 0; p0 < n; p0 += L)
                items.push_back(PoolItem{c, size, p0, p0 + L < n ? p0 + L : n, off});
        }
    }
    CB_CHECK(ctx, items.size() < (size_t)2147483647, "cb200_pooled_size_factors: too many work items");
    CB_CUDA(ctx, d_items.ensure(items.size()));
    CB_TRY(cb200_h2d(ctx, d_items.p, items.data(), items.size() * sizeof(PoolItem), ctx->stream));
    CB_CUDA(ctx, cudaFuncSetAttribute(k_pool_medians, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CB_CUDA(ctx, cudaEventRecord(ev[1], ctx->stream));
    {
        int per_sm = 1;
        CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pool_medians, POOL_THREADS, smem));
        per_sm = per_sm < 1 ? 1 : per_sm > 8 ? 8 : per_sm;
        int64_t grid = (int64_t)ctx->num_sms * per_sm;
        if (grid > (int64_t)items.size())
            grid = (int64_t)items.size();
        k_pool_medians<<<(unsigned)grid, POOL_THREADS, smem, ctx->stream>>>(
            d_items.p, (int)items.size(), d_cptr.p, d_eslot.p, d_eq.p, d_clptr.p, d_pinv.p, d_pave.p, d_gf.p, d_meanlib.p, G,
            gf_max, ldexp(1.0, -s2), d_bvals.p, ctx->flag.p);
        CB_LAUNCH(ctx);
    }
    CB_CUDA(ctx, cudaEventRecord(ev[2], ctx->stream));

    // ---- step 4
    k_pool_solve<<<C, 512, 0, ctx->stream>>>(d_clptr.p, d_sizes.p, d_ns.p, n_sizes, N, d_bvals.p, d_ring.p, ctx->blk_w.p,
                                             d_work.p, d_work.p + N, d_work.p + 2 * N, d_work.p + 3 * N, d_work.p + 4 * N, d_nf.p);
    CB_LAUNCH(ctx);
    CB_CUDA(ctx, cudaEventRecord(ev[3], ctx->stream));
    std::vector<double> nf((size_t)N), ave(CG);
    CB_TRY(cb200_d2h(ctx, nf.data(), d_nf.p, (size_t)N * sizeof(double), ctx->stream));
    CB_TRY(cb200_d2h(ctx, ave.data(), d_avefull.p, CG * sizeof(double), ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p, sizeof(int)));
    CB_CHECK(ctx, (h_flag & 12) == 0, "cb200_pooled_size_factors: internal error in the median selection (flag %d)", h_flag);
    const auto t_host = clk::now();

    // ---- step 5 (host, G-sized): reference cluster, median ratios of the pseudo-cells, final scaling
    int ref = 0;
    {
        std::vector<int64_t> nzc((size_t)C, 0);
        pool_parallel_for(C, [&](int c) {
            int64_t k = 0;
            for (int64_t g = 0; g < G; ++g)
                k += ave[(size_t)c * G + g] > 0.0 ? 1 : 0;
            nzc[(size_t)c] = k;
        });
        for (int c = 1; c < C; ++c)
            if (nzc[(size_t)c] > nzc[(size_t)ref])
                ref = c;
    }
    std::vector<double> resc((size_t)C, 1.0);
    std::vector<int> bad((size_t)C, 0);
    pool_parallel_for(C, [&](int c) {
        const double *cur = ave.data() + (size_t)c * G, *rp = ave.data() + (size_t)ref * G;
        double cs = 0.0, rs = 0.0;
        for (int64_t g = 0; g < G; ++g) {
            cs += cur[g];
            rs += rp[g];
        }
        std::vector<double> ratio;
        ratio.reserve((size_t)G);
        for (int64_t g = 0; g < G; ++g) {
            const double grand = (cur[g] / cs + rp[g] / rs) / 2.0 * (cs + rs) / 2.0;
            if (grand >= min_mean) {
                const double q = cur[g] / rp[g];
                if (q == q)   // na.rm
                    ratio.push_back(q);
            }
        }
        const double m = pool_median(ratio);
        if (!(m > 0.0) || !std::isfinite(m))
            bad[(size_t)c] = 1;
        resc[(size_t)c] = m;
    });
    for (int c = 0; c < C; ++c)
        CB_CHECK(ctx, bad[(size_t)c] == 0, "cb200_pooled_size_factors: inter-cluster rescaling factor for cluster %d is not strictly positive", c);
    double pos_sum = 0.0;
    int64_t n_pos = 0;
    for (int64_t i = 0; i < N; ++i) {
        const double v = nf[(size_t)i] * resc[(size_t)cl[(size_t)i]];
        sf_out[i] = v;
        if (v > 0.0) {
            pos_sum += v;
            ++n_pos;
        }
    }
    CB_CHECK(ctx, n_pos > 0, "cb200_pooled_size_factors: no positive size factor");
    const double scale = (double)n_pos / pos_sum;
    for (int64_t i = 0; i < N; ++i)
        sf_out[i] *= scale;
    if (n_nonpositive != nullptr)
        *n_nonpositive = N - n_pos;
    if (stage_ms_out != nullptr) {
        float a = 0.f, b = 0.f, c2 = 0.f;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        cudaEventElapsedTime(&c2, ev[2], ev[3]);
        stage_ms_out[5] = (double)gf_max;
        stage_ms_out[6] = (double)C;
        stage_ms_out[7] = (double)n_kept;
        stage_ms_out[0] = a;    // pseudo-cells + gene filter + compact store
        stage_ms_out[1] = b;    // pool medians
        stage_ms_out[2] = c2;   // circulant solves
        stage_ms_out[3] = std::chrono::duration<double, std::milli>(clk::now() - t_host).count();   // host rescaling
        stage_ms_out[4] = std::chrono::duration<double, std::milli>(clk::now() - t_begin).count();  // whole call
    }
    return 0;
}
