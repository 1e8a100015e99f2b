// cb200_knn.cu -- K6: exact Euclidean kNN in PC space.
//
// Replaces BiocNeighbors::findKNN(x, k, BNPARAM = KmknnParam(), get.distance = FALSE), the
// neighbour search inside bluster::makeSNNGraph() that /root/reference/R/
// makeGraphsAndClusters.R:83-85 (and :125-127, :284) calls.  Contract:
// for every cell the k nearest other cells ordered by (distance, index), distance =
// sqrt(sum_t (a_t - b_t)^2) accumulated sequentially in fp64 without FMA contraction.
//
// Structure
//   k_knn_prep      fp64 embedding -> padded fp32 copy, 0.5|r|^2, max |r|^2
//   k_knn_tiles     128 x 128 distance tiles (norm expansion, key = 0.5|r|^2 - q.r) with a
//                   fused per-query top-KP selection held in shared memory: the N x N
//                   distance matrix never exists in HBM.  KP > k candidates are kept.
//   k_knn_rerank    exact fp64 distances of the KP candidates in the reference's arithmetic,
//                   sort by (distance, index), emit k; prove with the a-priori error bound of
//                   the approximate keys that no non-candidate can enter the top k, else
//   k_knn_exact     flag the query for a full fp64 scan (rare).
// The result is therefore exact regardless of the precision of the tile kernel.
#include <cfloat>
#include <cmath>

#include <cstdlib>
#include <cstring>

#include "cb200_common.cuh"
#include "cb200_knn_tc.cuh"

#define KNN_THREADS 256
#define KNN_TN 128       // references per tile
#define KNN_MAX_DP 64    // padded embedding width
#define KNN_FB_CAP 2048  // candidates the exact-scan fallback can hold per query

__global__ void k_colmajor_to_rowmajor(const double *__restrict__ in, int64_t rows, int cols, double *__restrict__ out)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < rows * cols) {
        int64_t r = idx / cols;
        int c = (int)(idx % cols);
        out[idx] = in[r + rows * (int64_t)c];
    }
}

__global__ void k_knn_prep(const double *__restrict__ E, int64_t lde, int64_t n, int d, int DP, float *__restrict__ Ef,
                           float *__restrict__ hn, double *__restrict__ nrm2, unsigned long long *__restrict__ maxbits)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    double s = 0.0;
    for (int t = 0; t < d; ++t) {
        double v = E[i * lde + t];
        s += v * v;
        Ef[i * DP + t] = (float)v;
    }
    for (int t = d; t < DP; ++t)
        Ef[i * DP + t] = 0.f;
    hn[i] = (float)(0.5 * s);
    nrm2[i] = s;
    atomicMax(maxbits, (unsigned long long)__double_as_longlong(s));  // s >= 0: bit order == value order
}

// ---------------------------------------------------------------------------------------
// K6b: distance tiles + fused top-KP selection (fp32 SIMT tile engine)
// ---------------------------------------------------------------------------------------
// Per-row candidate store in shared memory: CAP = 2*KP slots.  A candidate is appended when
// its key beats the row's threshold tau (the KP-th best key seen so far, +inf until KP have
// been seen); a row holding more than KP entries is compacted by one warp (bitonic sort,
// keep KP, refresh tau).  Appends that find the store full are retried after compaction.
template <int KP, int TM>
struct KnnSmem {
    static constexpr int CAP = 2 * KP;
    static constexpr size_t q_floats = (size_t)KNN_MAX_DP * TM;
    static constexpr size_t r_floats = (size_t)KNN_MAX_DP * KNN_TN;
    static constexpr size_t bytes = (q_floats + r_floats + KNN_TN) * sizeof(float) + (size_t)TM * sizeof(float) +
                                    (size_t)TM * sizeof(int) + (size_t)TM * CAP * (sizeof(float) + sizeof(int));
};

template <int KP>
__device__ __forceinline__ void knn_compact_row(float *bkey, int *bidx, int *cnt_p, float *tau_p, int lane, bool force)
{
    constexpr int CAP = 2 * KP;
    constexpr int EPL = CAP / 32;
    int c = *cnt_p;
    if (c > CAP)
        c = CAP;
    if (!force && c <= KP)
        return;
    float key[EPL];
    int idx[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int p = e * 32 + lane;
        key[e] = p < c ? bkey[p] : FLT_MAX;
        idx[e] = p < c ? bidx[p] : 0x7fffffff;
    }
    warp_bitonic_sort<float, EPL>(key, idx, lane);
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int p = e * 32 + lane;
        if (p < KP) {
            bkey[p] = key[e];
            bidx[p] = idx[e];
        }
    }
    // tau = key at sorted position KP-1 (register (KP-1)/32 of lane 31)
    const float kth = __shfl_sync(0xffffffffu, key[(KP - 1) / 32], 31);
    __syncwarp();
    if (lane == 0) {
        *cnt_p = c < KP ? c : KP;
        *tau_p = c >= KP ? kth : FLT_MAX;
    }
    __syncwarp();
}

template <int KP, int TM>
__global__ void __launch_bounds__(KNN_THREADS, 1)
k_knn_tiles(const float *__restrict__ Ef, const float *__restrict__ hn, int64_t n, int DP, int64_t q_begin,
            int64_t q_end, int32_t *__restrict__ cand_idx, float *__restrict__ cand_tau)
{
    constexpr int CAP = 2 * KP;
    constexpr int RPT = TM / 16;  // query rows per thread: 8 (TM=128) or 4 (TM=64)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Qs = reinterpret_cast<float *>(smem_raw);         // [DP][TM]
    float *Rs = Qs + (size_t)KNN_MAX_DP * TM;                // [DP][KNN_TN]
    float *hn_s = Rs + (size_t)KNN_MAX_DP * KNN_TN;          // [KNN_TN]
    float *tau_s = hn_s + KNN_TN;                            // [TM]
    int *cnt_s = reinterpret_cast<int *>(tau_s + TM);        // [TM]
    float *bkey = reinterpret_cast<float *>(cnt_s + TM);     // [TM][CAP]
    int *bidx = reinterpret_cast<int *>(bkey + (size_t)TM * CAP);  // [TM][CAP]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t q0 = q_begin + (int64_t)blockIdx.x * TM;

    for (int i = tid; i < TM; i += KNN_THREADS) {
        tau_s[i] = FLT_MAX;
        cnt_s[i] = 0;
    }
    // query tile, transposed into [t][row]; rows beyond the range are zero
    const int dp4 = DP >> 2;
    for (int i = tid; i < TM * dp4; i += KNN_THREADS) {
        const int row = i % TM, c4 = i / TM;
        const int64_t g = q0 + row;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < q_end)
            v = *reinterpret_cast<const float4 *>(Ef + g * DP + c4 * 4);
        Qs[(c4 * 4 + 0) * TM + row] = v.x;
        Qs[(c4 * 4 + 1) * TM + row] = v.y;
        Qs[(c4 * 4 + 2) * TM + row] = v.z;
        Qs[(c4 * 4 + 3) * TM + row] = v.w;
    }
    __syncthreads();

    for (int64_t r0 = 0; r0 < n; r0 += KNN_TN) {
        for (int i = tid; i < KNN_TN * dp4; i += KNN_THREADS) {
            const int row = i % KNN_TN, c4 = i / KNN_TN;
            const int64_t g = r0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g < n)
                v = *reinterpret_cast<const float4 *>(Ef + g * DP + c4 * 4);
            Rs[(c4 * 4 + 0) * KNN_TN + row] = v.x;
            Rs[(c4 * 4 + 1) * KNN_TN + row] = v.y;
            Rs[(c4 * 4 + 2) * KNN_TN + row] = v.z;
            Rs[(c4 * 4 + 3) * KNN_TN + row] = v.w;
        }
        if (tid < KNN_TN)
            hn_s[tid] = (r0 + tid < n) ? hn[r0 + tid] : FLT_MAX;
        __syncthreads();

        float acc[RPT][8];
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                acc[i][j] = 0.f;
        for (int t = 0; t < DP; ++t) {
            float a[RPT], b[8];
            const float4 a0 = *reinterpret_cast<const float4 *>(Qs + t * TM + ty * 4);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            if (RPT == 8) {
                const float4 a1 = *reinterpret_cast<const float4 *>(Qs + t * TM + (TM / 2) + ty * 4);
                a[RPT - 4] = a1.x; a[RPT - 3] = a1.y; a[RPT - 2] = a1.z; a[RPT - 1] = a1.w;
            }
            const float4 b0 = *reinterpret_cast<const float4 *>(Rs + t * KNN_TN + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4 *>(Rs + t * KNN_TN + 64 + tx * 4);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }

        // selection; `done` marks the elements of this thread's micro-tile already stored
        unsigned long long done = 0ULL;
        int again;
        do {
            int ovf = 0;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int row = (i < 4) ? (ty * 4 + i) : ((TM / 2) + ty * 4 + (i - 4));
                const int64_t grow = q0 + row;
                if (grow >= q_end)
                    continue;
                const float tau = tau_s[row];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int bit = i * 8 + j;
                    if ((done >> bit) & 1ULL)
                        continue;
                    const int col = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4));
                    const int64_t gcol = r0 + col;
                    const float key = hn_s[col] - acc[i][j];
                    if (key < tau && gcol < n && gcol != grow) {
                        const int pos = atomicAdd(&cnt_s[row], 1);
                        if (pos < CAP) {
                            bkey[(size_t)row * CAP + pos] = key;
                            bidx[(size_t)row * CAP + pos] = (int)gcol;
                            done |= (1ULL << bit);
                        } else {
                            ovf = 1;
                        }
                    }
                }
            }
            __syncthreads();
            for (int row = w; row < TM; row += KNN_THREADS / 32)
                knn_compact_row<KP>(bkey + (size_t)row * CAP, bidx + (size_t)row * CAP, cnt_s + row, tau_s + row, lane,
                                    false);
            again = __syncthreads_or(ovf);
        } while (again);
    }

    // final: sort every row's store and write the KP best candidates + the row threshold
    for (int row = w; row < TM; row += KNN_THREADS / 32) {
        const int64_t grow = q0 + row;
        if (grow >= q_end)
            continue;
        knn_compact_row<KP>(bkey + (size_t)row * CAP, bidx + (size_t)row * CAP, cnt_s + row, tau_s + row, lane, true);
        const int c = cnt_s[row];
        int32_t *out = cand_idx + (grow - q_begin) * KP;
        for (int p = lane; p < KP; p += 32)
            out[p] = p < c ? bidx[(size_t)row * CAP + p] : -1;
        if (lane == 0)
            cand_tau[grow - q_begin] = tau_s[row];
    }
}

// ---------------------------------------------------------------------------------------
// K6c: exact re-rank.  Warp per query.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double exact_dist(const double *__restrict__ a, const double *__restrict__ b, int d)
{
    // the reference's arithmetic: sequential fp64, separate multiply and add, then sqrt
    double acc = 0.0;
    for (int t = 0; t < d; ++t) {
        const double dl = __dsub_rn(a[t], b[t]);
        acc = __dadd_rn(acc, __dmul_rn(dl, dl));
    }
    return __dsqrt_rn(acc);
}

template <int KP>
__global__ void __launch_bounds__(256)
k_knn_rerank(const double *__restrict__ E, int64_t lde, int d, int64_t q_begin, int64_t nq, int k,
             const int32_t *__restrict__ cand_idx, const float *__restrict__ cand_tau, int ntau, const double *__restrict__ nrm2,
             const unsigned long long *__restrict__ maxbits, double err_coef, const float *__restrict__ cand_err,
             const int32_t *__restrict__ qlist, const double *__restrict__ eperm, const int32_t *__restrict__ perm,
             const int32_t *__restrict__ qpos,
             int32_t *__restrict__ knn, double *__restrict__ knn_dk, int32_t *__restrict__ fb_list,
             int *__restrict__ fb_count)
{
    constexpr int EPL = KP / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double m2 = __longlong_as_double((long long)*maxbits);
    for (int64_t q = warp0; q < nq; q += nwarps) {
        // q runs over the candidate records; the tensor-core engine orders them by PC 1 (qlist)
        const int64_t grow = qlist != nullptr ? (int64_t)qlist[q] : q_begin + q;
        const int64_t orow = grow - q_begin;   // row of the result
        // tensor-core engine: the candidates are sorted POSITIONS and the rows are read from the copy of the
        // embedding in sorted order (same values; queries that follow each other share most candidates, so
        // the gathers hit L2); the original index -- the tie-break and the result -- comes from perm
        const double *qa = eperm != nullptr ? eperm + (int64_t)qpos[q] * d : E + grow * lde;
        double key[EPL];
        int idx[EPL];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            int ci = cand_idx[q * KP + e * 32 + lane];
            int oi = ci;
            if (eperm != nullptr && ci >= 0) {
                oi = perm[ci];
                if (oi < 0)
                    ci = -1;   // a padding position (only kept while fewer than KP references were seen)
            }
            if (ci >= 0) {
                if (eperm != nullptr) {
                    key[e] = exact_dist(qa, eperm + (int64_t)ci * d, d);
                    idx[e] = oi;
                } else {
                    key[e] = exact_dist(qa, E + (int64_t)ci * lde, d);
                    idx[e] = ci;
                }
            } else {
                key[e] = DBL_MAX;
                idx[e] = 0x7fffffff;
            }
        }
        warp_bitonic_sort<double, EPL>(key, idx, lane);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int p = e * 32 + lane;
            if (p < k)
                knn[orow * k + p] = idx[e];
        }
        // k-th exact distance sits at sorted position k-1
        double dk = 0.0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const double v = __shfl_sync(0xffffffffu, key[e], (k - 1) & 31);
            if (e == ((k - 1) >> 5))
                dk = v;
        }
        if (lane == 0) {
            knn_dk[orow] = dk;
            // exact key of the k-th candidate and the bound of |approx key - exact key|: the tensor-core
            // engine uses the canonical key 0.5 d^2 and records a bound per query, the fp32 engine the
            // key 0.5 d^2 - 0.5 |q|^2 with an a-priori bound
            double ek, bound;
            if (cand_err != nullptr) {
                ek = 0.5 * dk * dk;
                bound = (double)cand_err[q];
            } else {
                const double nq2 = nrm2[grow];
                ek = 0.5 * dk * dk - 0.5 * nq2;
                bound = err_coef * (nq2 + m2);
            }
            // every non-candidate has an approximate key >= the smallest of the ntau thresholds
            float tau = cand_tau[q * ntau];
            for (int u = 1; u < ntau; ++u)
                tau = fminf(tau, cand_tau[q * ntau + u]);
            const bool proven = (tau == FLT_MAX) || (ek + bound < (double)tau - bound);
            if (!proven) {
                const int pos = atomicAdd(fb_count, 1);
                fb_list[pos] = (int32_t)orow;
            }
        }
    }
}

// Exact scan for the flagged queries: every reference whose exact distance is <= the exact distance
// of the k-th candidate is collected, then ranked by (distance, index).  Two kernels so that a single
// flagged query still uses the whole GPU: blockIdx.y slices the references, the hits of a query are
// appended to its global buffer; then one block per query ranks them.
__global__ void __launch_bounds__(256)
k_knn_exact_collect(const double *__restrict__ E, int64_t lde, int d, int64_t n, int64_t q_begin,
                    const int32_t *__restrict__ fb_list, const double *__restrict__ knn_dk, int *__restrict__ gcnt,
                    double *__restrict__ gd, int32_t *__restrict__ gi)
{
    const int64_t q = fb_list[blockIdx.x];
    const int64_t grow = q_begin + q;
    const double thr = knn_dk[q];
    const double *qa = E + grow * lde;
    const int64_t per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * per, r1 = r0 + per < n ? r0 + per : n;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        if (r == grow)
            continue;
        const double dist = exact_dist(qa, E + r * lde, d);
        if (dist <= thr) {
            const int pos = atomicAdd(gcnt + blockIdx.x, 1);
            if (pos < KNN_FB_CAP) {
                gd[(size_t)blockIdx.x * KNN_FB_CAP + pos] = dist;
                gi[(size_t)blockIdx.x * KNN_FB_CAP + pos] = (int32_t)r;
            }
        }
    }
}

// The same scan for QB flagged queries at once: a thread walks one reference row and updates QB accumulators (the
// queries sit in shared memory, broadcast reads), so the embedding is read once per QB queries instead of once per
// query -- at k = 100 on 1 M points ~1 450 queries take this path and the per-query scan moved 580 GB.  Arithmetic and
// order of the sums are those of exact_dist (one accumulator per query, coordinates ascending).
template <int QB>
__global__ void __launch_bounds__(256)
k_knn_exact_collect_batch(const double *__restrict__ E, int64_t lde, int d, int64_t n, int64_t q_begin,
                          const int32_t *__restrict__ fb_list, int nfb, const double *__restrict__ knn_dk,
                          int *__restrict__ gcnt, double *__restrict__ gd, int32_t *__restrict__ gi)
{
    extern __shared__ double fbq[];   // [QB][d]
    __shared__ long long s_grow[QB];
    __shared__ double s_thr[QB];
    const int b0 = blockIdx.x * QB;
    const int nb = nfb - b0 < QB ? nfb - b0 : QB;
    for (int i = threadIdx.x; i < QB * d; i += blockDim.x) {
        const int b = i / d, t = i - b * d;
        fbq[i] = b < nb ? E[(q_begin + (int64_t)fb_list[b0 + b]) * lde + t] : 0.0;
    }
    if (threadIdx.x < QB) {
        const int b = threadIdx.x;
        s_grow[b] = b < nb ? q_begin + (int64_t)fb_list[b0 + b] : -1;
        s_thr[b] = b < nb ? knn_dk[fb_list[b0 + b]] : -1.0;
    }
    __syncthreads();
    const int64_t per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * per, r1 = r0 + per < n ? r0 + per : n;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        double acc[QB];
#pragma unroll
        for (int b = 0; b < QB; ++b)
            acc[b] = 0.0;
        const double *row = E + r * lde;
        for (int t = 0; t < d; ++t) {
            const double rv = row[t];
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                const double dl = __dsub_rn(fbq[b * d + t], rv);
                acc[b] = __dadd_rn(acc[b], __dmul_rn(dl, dl));
            }
        }
#pragma unroll
        for (int b = 0; b < QB; ++b) {
            const double dist = __dsqrt_rn(acc[b]);
            if (b < nb && r != s_grow[b] && dist <= s_thr[b]) {
                const int pos = atomicAdd(gcnt + b0 + b, 1);
                if (pos < KNN_FB_CAP) {
                    gd[(size_t)(b0 + b) * KNN_FB_CAP + pos] = dist;
                    gi[(size_t)(b0 + b) * KNN_FB_CAP + pos] = (int32_t)r;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_knn_exact_rank(int k, const int32_t *__restrict__ fb_list, const int *__restrict__ gcnt, const double *__restrict__ gd,
                 const int32_t *__restrict__ gi, int32_t *__restrict__ knn, int *__restrict__ err_flag)
{
    __shared__ double sd[KNN_FB_CAP];
    __shared__ int si[KNN_FB_CAP];
    const int64_t q = fb_list[blockIdx.x];
    const int c = gcnt[blockIdx.x];
    if (c > KNN_FB_CAP || c < k) {
        if (threadIdx.x == 0)
            atomicOr(err_flag, 1);
        return;
    }
    for (int e = threadIdx.x; e < c; e += blockDim.x) {
        sd[e] = gd[(size_t)blockIdx.x * KNN_FB_CAP + e];
        si[e] = gi[(size_t)blockIdx.x * KNN_FB_CAP + e];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < c; e += blockDim.x) {
        const double de = sd[e];
        const int ie = si[e];
        int rank = 0;
        for (int o = 0; o < c; ++o)
            rank += pair_less<double>(sd[o], si[o], de, ie) ? 1 : 0;
        if (rank < k)
            knn[q * k + rank] = ie;
    }
}

// ---------------------------------------------------------------------------------------
// General path for shapes the tile engines do not take (k > 112 -- rebuildModularity's default k = round(sqrt(N)),
// uwot's n_neighbors = floor(sqrt(N)) -- or more than 64 dimensions): exact, slow, any k <= KNN_BIG_MAXK.
// One block per query: the distances (reference arithmetic) to a strided sample give a radius that holds a few more
// than k references; all references within it are collected and ranked by (distance, index).  The radius is
// widened / narrowed until the count is in [k, KNN_BIG_CAP], so the result does not depend on the sample.
// ---------------------------------------------------------------------------------------
#define KNN_BIG_CAP 4096
#define KNN_BIG_MAXK 2048
__global__ void __launch_bounds__(256)
k_knn_bigk(const double *__restrict__ E, int64_t lde, int d, int64_t n, int64_t q_begin, int64_t nq, int k,
           int32_t *__restrict__ knn, int *__restrict__ err_flag)
{
    extern __shared__ __align__(16) unsigned char bk_smem[];
    double *sd = reinterpret_cast<double *>(bk_smem);                               // [KNN_BIG_CAP] sorted sample distances
    double *hd = sd + KNN_BIG_CAP;                                                   // [KNN_BIG_CAP] hits: distance
    int32_t *hx = reinterpret_cast<int32_t *>(hd + KNN_BIG_CAP);                     // [KNN_BIG_CAP] hits: index
    __shared__ int s_cnt, s_state;     // s_state: 0 = retry with s_tau, 1 = done, 2 = failed
    __shared__ double s_tau;
    const int tid = threadIdx.x;
    for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
        const int64_t grow = q_begin + q;
        const double *qa = E + grow * lde;
        // sample: S points at equal strides (all of them when n - 1 <= KNN_BIG_CAP)
        const int S = (int)(n - 1 < KNN_BIG_CAP ? n - 1 : KNN_BIG_CAP);
        for (int s = tid; s < KNN_BIG_CAP; s += blockDim.x) {
            double v = 1.0 / 0.0;
            if (s < S) {
                int64_t r = S == n - 1 ? s : (int64_t)(((__int128)s * (n - 1)) / S);
                if (r >= grow)
                    ++r;       // skip the query itself
                v = exact_dist(qa, E + r * lde, d);
            }
            sd[s] = v;
        }
        __syncthreads();
        // bitonic sort of the KNN_BIG_CAP sample distances (ascending)
        for (int size = 2; size <= KNN_BIG_CAP; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = tid; i < KNN_BIG_CAP / 2; i += blockDim.x) {
                    const int lo = 2 * i - (i & (stride - 1));
                    const int hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const double a = sd[lo], b = sd[hi];
                    if ((a > b) == up) {
                        sd[lo] = b;
                        sd[hi] = a;
                    }
                }
                __syncthreads();
            }
        }
        // thread 0 steers the radius: sample rank m, bracketed by ranks known to give too few / too many hits
        int m_lo = -1, m_hi = S, m = 0;
        if (tid == 0) {
            m = S == n - 1 ? k - 1 : (int)(((double)k * S) / (double)(n - 1) * 1.3) + 8;
            if (m > S - 1)
                m = S - 1;
            s_tau = sd[m];
            s_state = 0;
        }
        __syncthreads();
        for (int attempt = 0; attempt < 40; ++attempt) {
            const double tau = s_tau;
            if (tid == 0)
                s_cnt = 0;
            __syncthreads();
            for (int64_t r = tid; r < n; r += blockDim.x) {
                if (r == grow)
                    continue;
                const double dist = exact_dist(qa, E + r * lde, d);
                if (dist <= tau) {
                    const int pos = atomicAdd(&s_cnt, 1);
                    if (pos < KNN_BIG_CAP) {
                        hd[pos] = dist;
                        hx[pos] = (int32_t)r;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                const int c = s_cnt;
                if (c >= k && c <= KNN_BIG_CAP) {
                    s_state = 1;
                } else {
                    int m2;
                    if (c < k) {
                        m_lo = m;
                        m2 = m_hi == S ? (2 * m + 1 < S - 1 ? 2 * m + 1 : S - 1) : (m + m_hi + 1) / 2;
                    } else {
                        m_hi = m;
                        m2 = (m_lo + m) / 2;
                    }
                    if (m2 == m || m2 < 0 || m2 >= S || m2 <= m_lo || m2 >= m_hi) {
                        // the sample cannot separate k from KNN_BIG_CAP: the whole data set as a last resort, else ties
                        if (c < k && tau < 1.0 / 0.0 && n - 1 <= KNN_BIG_CAP)
                            s_tau = 1.0 / 0.0;
                        else
                            s_state = 2;
                    } else {
                        m = m2;
                        s_tau = sd[m];
                    }
                }
            }
            __syncthreads();
            if (s_state != 0)
                break;
        }
        if (s_state == 1) {
            const int c = s_cnt;
            for (int e = tid; e < c; e += blockDim.x) {
                const double de = hd[e];
                const int ie = hx[e];
                int rank = 0;
                for (int o = 0; o < c; ++o)
                    rank += pair_less<double>(hd[o], hx[o], de, ie) ? 1 : 0;
                if (rank < k)
                    knn[q * k + rank] = ie;
            }
        } else if (tid == 0) {
            atomicOr(err_flag, 1);
        }
        __syncthreads();
    }
}

// knn (row-major, 0-based) -> column-major 1-based (R layout)
__global__ void k_knn_to_r_layout(const int32_t *__restrict__ knn, int64_t nq, int k, int32_t *__restrict__ out)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < nq * k) {
        int64_t q = idx % nq;
        int c = (int)(idx / nq);
        out[idx] = knn[q * k + c] + 1;
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int KP, int TM>
static int run_tiles_simt(cb200_ctx *ctx, int64_t n, int DP, int64_t q_begin, int64_t q_end)
{
    const int64_t nq = q_end - q_begin;
    const size_t smem = KnnSmem<KP, TM>::bytes;
    CB_CUDA(ctx, cudaFuncSetAttribute(k_knn_tiles<KP, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CB_CUDA(ctx, ctx->cand_idx.ensure((size_t)nq * KP));
    CB_CUDA(ctx, ctx->cand_tau.ensure((size_t)nq));
    cb_stage_begin(ctx, ST_KNN_TILE);
    const unsigned grid = (unsigned)((nq + TM - 1) / TM);
    k_knn_tiles<KP, TM><<<grid, KNN_THREADS, smem, ctx->stream>>>(ctx->emb_f.p, ctx->hnorm.p, n, DP, q_begin, q_end,
                                                                   ctx->cand_idx.p, ctx->cand_tau.p);
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_KNN_TILE);
    return 0;
}

template <int KP>
static int run_rerank(cb200_ctx *ctx, const double *E, int64_t lde, int64_t n, int d, int k, int64_t q_begin,
                      int64_t nq, double err_coef, int ntau, const int32_t *qlist)
{
    cb_stage_begin(ctx, ST_KNN_RERANK);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 2, 0, sizeof(int), ctx->stream));
    const int rgrid = cb_grid_for(nq, 8, ctx->num_sms * 16);
    k_knn_rerank<KP><<<rgrid, 256, 0, ctx->stream>>>(E, lde, d, q_begin, nq, k, ctx->cand_idx.p, ctx->cand_tau.p,
                                                     ntau, ctx->nrm2.p, ctx->maxnrm.p, err_coef,
                                                     err_coef < 0.0 ? ctx->cand_err.p : nullptr, qlist,
                                                     err_coef < 0.0 ? ctx->knn_eperm.p : nullptr, ctx->knn_perm.p, ctx->knn_qpos.p,
                                                     ctx->knn.p, ctx->knn_dk.p,
                                                     ctx->fb_list.p, ctx->flag.p + 2);
    CB_LAUNCH(ctx);
    int h[2] = {0, 0};
    CB_TRY(cb200_read_back(ctx, h, ctx->flag.p + 2, 2 * sizeof(int)));
    CB_CHECK(ctx, (h[1] & 4) == 0, "cb200_knn: tensor-core tile kernel timed out on a barrier (protocol error)");
    ctx->tm.knn_fallback_queries = h[0];
    if (qlist != nullptr) {  // tensor-core engine: how much of the N x N tile grid survived the PC-1 pruning
        unsigned long long done = 0;
        CB_TRY(cb200_read_back(ctx, &done, ctx->maxnrm.p + 3, sizeof(done)));
        ctx->tm.knn_tiles_scanned = (int64_t)done;
        ctx->tm.knn_tiles_total = ((nq + 127) / 128) * ((n + 127) / 128);
        if (ctx->tm.knn_tiles_total < ctx->tm.knn_tiles_scanned)   // tiny inputs: the clusters are padded to whole tiles
            ctx->tm.knn_tiles_total = ctx->tm.knn_tiles_scanned;
    } else {
        ctx->tm.knn_tiles_scanned = ctx->tm.knn_tiles_total = ((nq + 127) / 128) * ((n + 127) / 128);
    }
    if (h[0] > 0) {
        const int nfb = h[0];
        int slices = (ctx->num_sms * 8 + nfb - 1) / nfb;   // enough blocks to fill the GPU whatever the count
        if ((int64_t)slices * 4096 > n)
            slices = (int)((n + 4095) / 4096);
        if (slices < 1)
            slices = 1;
        CB_CUDA(ctx, ctx->fb_cnt.ensure((size_t)nfb));
        CB_CUDA(ctx, ctx->fb_d.ensure((size_t)nfb * KNN_FB_CAP));
        CB_CUDA(ctx, ctx->fb_i.ensure((size_t)nfb * KNN_FB_CAP));
        CB_CUDA(ctx, cudaMemsetAsync(ctx->fb_cnt.p, 0, (size_t)nfb * sizeof(int), ctx->stream));
        if (nfb >= 8) {
            // many flagged queries: scan the references once per batch of queries
            const int qb = nfb >= 64 ? 16 : 4;
            const int nbatch = (nfb + qb - 1) / qb;
            int bs = (ctx->num_sms * 8 + nbatch - 1) / nbatch;
            if ((int64_t)bs * 4096 > n)
                bs = (int)((n + 4095) / 4096);
            if (bs < 1)
                bs = 1;
            const size_t fsm = (size_t)qb * d * sizeof(double);
            if (qb == 16)
                k_knn_exact_collect_batch<16><<<dim3((unsigned)nbatch, (unsigned)bs), 256, fsm, ctx->stream>>>(
                    E, lde, d, n, q_begin, ctx->fb_list.p, nfb, ctx->knn_dk.p, ctx->fb_cnt.p, ctx->fb_d.p, ctx->fb_i.p);
            else
                k_knn_exact_collect_batch<4><<<dim3((unsigned)nbatch, (unsigned)bs), 256, fsm, ctx->stream>>>(
                    E, lde, d, n, q_begin, ctx->fb_list.p, nfb, ctx->knn_dk.p, ctx->fb_cnt.p, ctx->fb_d.p, ctx->fb_i.p);
        } else {
            k_knn_exact_collect<<<dim3((unsigned)nfb, (unsigned)slices), 256, 0, ctx->stream>>>(
                E, lde, d, n, q_begin, ctx->fb_list.p, ctx->knn_dk.p, ctx->fb_cnt.p, ctx->fb_d.p, ctx->fb_i.p);
        }
        CB_LAUNCH(ctx);
        k_knn_exact_rank<<<nfb, 256, 0, ctx->stream>>>(k, ctx->fb_list.p, ctx->fb_cnt.p, ctx->fb_d.p, ctx->fb_i.p,
                                                       ctx->knn.p, ctx->flag.p + 3);
        CB_LAUNCH(ctx);
        CB_TRY(cb200_read_back(ctx, h, ctx->flag.p + 2, 2 * sizeof(int)));
        CB_CHECK(ctx, (h[1] & 1) == 0,
                 "cb200_knn: exact-scan fallback overflowed (more than %d ties); result not produced", KNN_FB_CAP);
    }
    cb_stage_end(ctx, ST_KNN_RERANK);
    return 0;
}

static int knn_core(cb200_ctx *ctx, const double *E, int64_t lde, int64_t n, int d, int k, int64_t q_begin,
                    int64_t q_end, int32_t *index_out)
{
    CB_CHECK(ctx, n >= 2, "cb200_knn: need at least 2 points");
    CB_CHECK(ctx, n < 2147483647LL, "cb200_knn: at most 2^31-2 points");
    CB_CHECK(ctx, d >= 1 && d <= 4096, "cb200_knn: 1 <= ndims <= 4096 required (got %d)", d);
    CB_CHECK(ctx, k >= 1 && (int64_t)k <= n - 1, "cb200_knn: need 1 <= k <= n-1 (k=%d, n=%lld)", k, (long long)n);
    CB_CHECK(ctx, k <= KNN_BIG_MAXK, "cb200_knn: k <= %d supported (got %d)", KNN_BIG_MAXK, k);
    CB_CHECK(ctx, q_begin >= 0 && q_begin < q_end && q_end <= n, "cb200_knn: bad query range");
    const int64_t nq = q_end - q_begin;
    CB_CUDA(ctx, ctx->knn.ensure((size_t)nq * k));
    ctx->have_knn = false;
    ctx->knn_E = E;   // for cb200_knn_distances
    ctx->knn_lde = lde;
    ctx->knn_d = d;
    if (k > 112 || d > KNN_MAX_DP) {
        // shapes outside the tile engines (rebuildModularity's k = round(sqrt(N)), wide embeddings): exact general path
        cb_stage_begin(ctx, ST_KNN_TILE);
        CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 3, 0, sizeof(int), ctx->stream));
        const size_t smem = (size_t)KNN_BIG_CAP * 20;
        CB_CUDA(ctx, cudaFuncSetAttribute(k_knn_bigk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = (int)(nq < (int64_t)ctx->num_sms * 2 ? nq : (int64_t)ctx->num_sms * 2);
        k_knn_bigk<<<grid, 256, smem, ctx->stream>>>(E, lde, d, n, q_begin, nq, k, ctx->knn.p, ctx->flag.p + 3);
        CB_LAUNCH(ctx);
        cb_stage_end(ctx, ST_KNN_TILE);
        int h_flag = 0;
        CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p + 3, sizeof(int)));
        CB_CHECK(ctx, h_flag == 0, "cb200_knn: more than %d references tie at the k-th distance of a query", KNN_BIG_CAP);
        ctx->tm.knn_engine = 3;
        ctx->tm.knn_fallback_queries = 0;
        ctx->tm.knn_tiles_scanned = ctx->tm.knn_tiles_total = 0;
        ctx->knn_nq = nq;
        ctx->knn_k = k;
        ctx->knn_ntotal = n;
        ctx->knn_qbegin = q_begin;
        ctx->have_knn = true;
        if (index_out != nullptr) {
            cb_stage_begin(ctx, ST_D2H);
            CB_CUDA(ctx, ctx->stage_i.ensure((size_t)nq * k));
            k_knn_to_r_layout<<<(unsigned)((nq * k + 255) / 256), 256, 0, ctx->stream>>>(ctx->knn.p, nq, k, ctx->stage_i.p);
            CB_LAUNCH(ctx);
            CB_TRY(cb200_d2h(ctx, index_out, ctx->stage_i.p, (size_t)nq * k * sizeof(int32_t), ctx->stream));
            cb_stage_end(ctx, ST_D2H);
            CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        return 0;
    }
    CB_CUDA(ctx, ctx->knn_dk.ensure((size_t)nq));
    CB_CUDA(ctx, ctx->fb_list.ensure((size_t)nq));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 3, 0, sizeof(int), ctx->stream));

    // tile engine: tensor cores (tcgen05) whenever the shape allows; CB200_KNN_ENGINE=simt forces the
    // fp32 CUDA-core engine (kept for d > 61 or k > 52, and as the cross-check in the tests)
    const char *eng = getenv("CB200_KNN_ENGINE");
    const bool want_simt = eng != nullptr && strcmp(eng, "simt") == 0;
    const bool use_tc = !want_simt && cb200_knn_tc_supported(d, k);
    int KP = 0, ntau = 1;
    double err_coef = 0.0;
    const int32_t *qlist = nullptr;  // order of the candidate records (tensor-core engine: sorted by PC 1)
    ctx->tm.knn_engine = use_tc ? 2 : 1;
    if (use_tc) {
        CB_TRY(cb200_knn_tiles_tc(ctx, E, lde, n, d, k, q_begin, q_end, &KP, &ntau, &qlist, &err_coef));
    } else {
        const int DP = (d + 3) & ~3;
        CB_CUDA(ctx, ctx->emb_f.ensure((size_t)n * DP));
        CB_CUDA(ctx, ctx->hnorm.ensure((size_t)n));
        CB_CUDA(ctx, ctx->nrm2.ensure((size_t)n));
        cb_stage_begin(ctx, ST_KNN_PREP);
        CB_CUDA(ctx, cudaMemsetAsync(ctx->maxnrm.p, 0, sizeof(unsigned long long), ctx->stream));
        k_knn_prep<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(E, lde, n, d, DP, ctx->emb_f.p, ctx->hnorm.p,
                                                                        ctx->nrm2.p, ctx->maxnrm.p);
        CB_LAUNCH(ctx);
        cb_stage_end(ctx, ST_KNN_PREP);
        // fp32 FMA chain of DP terms + input/half-norm roundings, see DESIGN.md "kNN error bound"
        err_coef = ldexp((double)(DP + 8), -24);
        if (k <= 24) {
            KP = 32;
            CB_TRY((run_tiles_simt<32, 128>(ctx, n, DP, q_begin, q_end)));
        } else if (k <= 52) {
            KP = 64;
            CB_TRY((run_tiles_simt<64, 128>(ctx, n, DP, q_begin, q_end)));
        } else {
            KP = 128;
            CB_TRY((run_tiles_simt<128, 64>(ctx, n, DP, q_begin, q_end)));
        }
    }
    if (KP == 32)
        CB_TRY(run_rerank<32>(ctx, E, lde, n, d, k, q_begin, nq, err_coef, ntau, qlist));
    else if (KP == 64)
        CB_TRY(run_rerank<64>(ctx, E, lde, n, d, k, q_begin, nq, err_coef, ntau, qlist));
    else
        CB_TRY(run_rerank<128>(ctx, E, lde, n, d, k, q_begin, nq, err_coef, ntau, qlist));
    ctx->knn_nq = nq;
    ctx->knn_k = k;
    ctx->knn_ntotal = n;
    ctx->knn_qbegin = q_begin;
    ctx->have_knn = true;

    if (index_out != nullptr) {
        cb_stage_begin(ctx, ST_D2H);
        CB_CUDA(ctx, ctx->stage_i.ensure((size_t)nq * k));
        k_knn_to_r_layout<<<(unsigned)((nq * k + 255) / 256), 256, 0, ctx->stream>>>(ctx->knn.p, nq, k, ctx->stage_i.p);
        CB_LAUNCH(ctx);
        CB_TRY(cb200_d2h(ctx, index_out, ctx->stage_i.p, (size_t)nq * k * sizeof(int32_t), ctx->stream));
        cb_stage_end(ctx, ST_D2H);
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

extern "C" int cb200_knn_device(cb200_ctx *ctx, const double *d_emb_rowmajor, int64_t n_total, int d, int k,
                                int64_t q_begin, int64_t q_end, int32_t *index_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, d_emb_rowmajor != nullptr, "cb200_knn_device: NULL embedding");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    return knn_core(ctx, d_emb_rowmajor, d, n_total, d, k, q_begin, q_end, index_out);
}

extern "C" int cb200_knn(cb200_ctx *ctx, const double *emb, int64_t n_total, int d, int k, int64_t q_begin,
                         int64_t q_end, int32_t *index_out)
{
    if (ctx == nullptr)
        return 1;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (emb == nullptr) {
        CB_CHECK(ctx, ctx->have_scores, "cb200_knn: no embedding given and no PCA scores in the context");
        CB_CHECK(ctx, n_total == ctx->N && ctx->N == ctx->N_total,
                 "cb200_knn: resident scores cover %lld of %lld cells; gather the embedding and use cb200_knn_device",
                 (long long)ctx->N, (long long)n_total);
        CB_CHECK(ctx, d <= ctx->d, "cb200_knn: ndims=%d exceeds the %d resident components", d, ctx->d);
        return knn_core(ctx, ctx->scores.p, ctx->d, n_total, d, k, q_begin, q_end, index_out);
    }
    CB_CHECK(ctx, n_total >= 2 && d >= 1, "cb200_knn: bad embedding shape");
    cb_stage_begin(ctx, ST_H2D);
    CB_CUDA(ctx, ctx->stage_d.ensure((size_t)n_total * d));
    CB_CUDA(ctx, ctx->emb.ensure((size_t)n_total * d));
    CB_TRY(cb200_h2d(ctx, ctx->stage_d.p, emb, (size_t)n_total * d * sizeof(double), ctx->stream));
    k_colmajor_to_rowmajor<<<(unsigned)((n_total * d + 255) / 256), 256, 0, ctx->stream>>>(ctx->stage_d.p, n_total, d,
                                                                                           ctx->emb.p);
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_H2D);
    return knn_core(ctx, ctx->emb.p, d, n_total, d, k, q_begin, q_end, index_out);
}
