// cb200_pca.cu -- K5: PCA of the HVG subset of the log-counts.
//
// Replaces scater::runPCA(sce, subset_row = hvgs, exprs_values = "logcounts",
// ncomponents = ndims) (/root/reference/R/doNormAndReduce.R:94-97) -> BiocSingular::runPCA:
// rank-d truncated SVD of the column-centred N x H matrix; scores = U D, rotation = V,
// varExplained = D^2/(N-1).
//
// Route: gather the H HVG rows of the CSC log-counts into a cell-major
// sparse sub-matrix, accumulate the H x H cross-product X^T X in fp64 (all-reduced across
// GPUs by the caller), centre it with the gene means already known from K3, take the top d
// eigenpairs of the covariance by Chebyshev-filtered block subspace iteration with
// Rayleigh-Ritz (fp64, no cuSOLVER), then project: scores = X V - 1 (mu^T V).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cb200_common.cuh"

int cb200_gram_tc(cb200_ctx *ctx, int nb, int *used);  // cb200_gram_tc.cu

// ---------------------------------------------------------------------------------------
// K5a: HVG sub-matrix extraction
// ---------------------------------------------------------------------------------------
__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = v;
}

__global__ void k_build_slot(const int32_t *__restrict__ hvg_idx, int H, int32_t *__restrict__ slot,
                             const double *__restrict__ gene_mean, double *__restrict__ mu)
{
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < H) {
        slot[hvg_idx[h]] = h;
        mu[h] = gene_mean[hvg_idx[h]];
    }
}

// warp per column: count entries that fall on an HVG row
__global__ void __launch_bounds__(256) k_hvg_count(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                   const int32_t *__restrict__ slot, int64_t n_cols,
                                                   int32_t *__restrict__ hcnt)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        int cnt = 0;
        int64_t i = b + lane;
        for (; i + 96 < e; i += 128) {
            int g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                g[u] = ld_stream_i(rowidx + i + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                cnt += (slot[g[u]] >= 0) ? 1 : 0;
        }
        for (; i < e; i += 32)
            cnt += (slot[rowidx[i]] >= 0) ? 1 : 0;
        cnt = warp_sum_i(cnt);
        if (lane == 0)
            hcnt[c] = cnt;
    }
}

// warp per column: ordered compaction of (slot, logcount) pairs
__global__ void __launch_bounds__(256)
k_hvg_fill(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, const double *__restrict__ logx,
           const int32_t *__restrict__ slot, int64_t n_cols, const int64_t *__restrict__ hptr,
           int32_t *__restrict__ hslot, double *__restrict__ hval)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        int64_t out = hptr[c];
        // 128 entries per pass: the four row-index loads, then the four slot look-ups, then the (few) value
        // loads are each issued back to back; the values of non-HVG rows are never read
        for (int64_t i0 = b; i0 < e; i0 += 128) {
            int g[4], s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + 32 * u + lane;
                g[u] = i < e ? ld_stream_i(rowidx + i) : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                s[u] = g[u] >= 0 ? slot[g[u]] : -1;
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = s[u] >= 0 ? logx[i0 + 32 * u + lane] : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned m = __ballot_sync(0xffffffffu, s[u] >= 0);
                if (s[u] >= 0) {
                    const int64_t o = out + __popc(m & ((1u << lane) - 1u));
                    hslot[o] = s[u];
                    hval[o] = v[u];
                }
                out += __popc(m);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// K5b (v1): sparse cross-product.  Warp per cell, every pair (a <= b) of its HVG entries adds
// y_a*y_b to the upper triangle (in slot order) with an fp64 reduction in L2.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gram_sparse(const int64_t *__restrict__ hptr, const int32_t *__restrict__ hslot, const double *__restrict__ hval,
              int64_t n_cols, int H, double *__restrict__ gram)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = hptr[c], e = hptr[c + 1];
        for (int64_t a = b; a < e; ++a) {
            const int sa = hslot[a];
            const double va = hval[a];
            for (int64_t q = a + lane; q < e; q += 32) {
                const int sb = hslot[q];
                const double vb = hval[q];
                const int r = sa < sb ? sa : sb;
                const int cc = sa < sb ? sb : sa;
                atomicAdd(gram + (int64_t)r * H + cc, va * vb);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// K5b (v2): tiled sparse cross-product.  HVG slots are numbered in ascending gene order, so the
// entries of a cell are sorted by slot and the entries that fall into a block of GB_BS slots are
// contiguous; bptr[c][blk] (uint16, relative to hptr[c]) marks the block boundaries.  A CTA owns
// one GB_BS x GB_BS tile (I <= J) of X^T X for a slice of the cells, accumulates the outer products
// of the I- and J-entries of every cell in shared memory (fp64 shared atomics), and flushes the
// tile once with fp64 reductions to L2.  Exact fp64 products; no dense operand is ever formed.
// ---------------------------------------------------------------------------------------
#define GB_BS 64

// warp per cell, lane per block: first entry with slot >= blk*GB_BS (binary search)
__global__ void __launch_bounds__(256)
k_hvg_blockptr(const int64_t *__restrict__ hptr, const int32_t *__restrict__ hslot, int64_t n_cols, int nb,
               uint16_t *__restrict__ bptr)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = hptr[c];
        const int m = (int)(hptr[c + 1] - b);
        for (int blk = lane; blk <= nb; blk += 32) {
            const int target = blk * GB_BS;
            int lo = 0, hi = m;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (hslot[b + mid] < target)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            bptr[c * (nb + 1) + blk] = (uint16_t)lo;
        }
    }
}

__global__ void __launch_bounds__(256)
k_gram_tiled(const int64_t *__restrict__ hptr, const int32_t *__restrict__ hslot, const double *__restrict__ hval,
             const uint16_t *__restrict__ bptr, int64_t n_cols, int nb, int H, double *__restrict__ gram)
{
    __shared__ double T[GB_BS][GB_BS + 1];
    // blockIdx.x enumerates the tiles (I <= J) in row-major order of the upper triangle
    int I = 0, rem = blockIdx.x;
    while (rem >= nb - I) {
        rem -= nb - I;
        ++I;
    }
    const int J = I + rem;
    for (int i = threadIdx.x; i < GB_BS * (GB_BS + 1); i += blockDim.x)
        (&T[0][0])[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t per = (n_cols + gridDim.y - 1) / gridDim.y;
    const int64_t c_begin = (int64_t)blockIdx.y * per;
    const int64_t c_end = c_begin + per < n_cols ? c_begin + per : n_cols;
    const int stride = nb + 1;
    for (int64_t c = c_begin + w; c < c_end; c += nw) {
        const uint16_t *bp = bptr + c * stride;
        const int i0 = bp[I], i1 = bp[I + 1], j0 = bp[J], j1 = bp[J + 1];
        const int nI = i1 - i0, nJ = j1 - j0;
        if (nI == 0 || nJ == 0)
            continue;
        const int64_t base = hptr[c];
        for (int ja = 0; ja < nJ; ja += 32) {
            const int nbj = nJ - ja < 32 ? nJ - ja : 32;
            int sJ = 0;
            double vJ = 0.0;
            if (lane < nbj) {
                sJ = hslot[base + j0 + ja + lane] - J * GB_BS;
                vJ = hval[base + j0 + ja + lane];
            }
            // lanes are split into G groups of nbp (a power of two >= nbj) lanes: group g takes I-entry
            // ia + g, lane (l % nbp) takes the J-entry
            int nbp = 1;
            while (nbp < nbj)
                nbp <<= 1;
            const int G = 32 / nbp;
            const int gsel = lane / nbp, bsel = lane & (nbp - 1);
            const double vb = __shfl_sync(0xffffffffu, vJ, bsel);
            const int sb = __shfl_sync(0xffffffffu, sJ, bsel);
            for (int ia = 0; ia < nI; ia += 32) {
                const int nai = nI - ia < 32 ? nI - ia : 32;
                int sI = 0;
                double vI = 0.0;
                if (lane < nai) {
                    sI = hslot[base + i0 + ia + lane] - I * GB_BS;
                    vI = hval[base + i0 + ia + lane];
                }
                for (int a0 = 0; a0 < nai; a0 += G) {
                    const int a = a0 + gsel;
                    const double va = __shfl_sync(0xffffffffu, vI, a & 31);
                    const int sa = __shfl_sync(0xffffffffu, sI, a & 31);
                    if (a < nai && bsel < nbj && (I != J || sa <= sb))
                        atomicAdd(&T[sa][sb], va * vb);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < GB_BS * GB_BS; i += blockDim.x) {
        const int r = i / GB_BS, cc = i % GB_BS;
        const int gr = I * GB_BS + r, gc = J * GB_BS + cc;
        const double v = T[r][cc];
        if (gr < H && gc < H && v != 0.0)
            atomicAdd(gram + (int64_t)gr * H + gc, v);
    }
}

// rotation in the caller's HVG order, column-major: out[h + H*j] = rot[perm[h]*d + j]
__global__ void k_rotation_out(const double *__restrict__ rot, const int32_t *__restrict__ perm, int H, int d,
                               double *__restrict__ out)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < (int64_t)H * d) {
        const int h = (int)(idx % H), j = (int)(idx / H);
        out[idx] = rot[(int64_t)perm[h] * d + j];
    }
}

// cov[i][j] = (G[min][max] - N mu_i mu_j) / (N - 1).  fixed != 0: G is the integer cross-product of the tensor-core
// engine, two 64-bit words per entry (cb200_gram_tc.cu): value = (lo + hi 2^16) 2^-38
__global__ void k_cov_from_gram(const double *__restrict__ gram, const double *__restrict__ mu, int H, double n_total,
                                double *__restrict__ cov, int fixed)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < (int64_t)H * H) {
        int i = (int)(idx / H), j = (int)(idx % H);
        int r = i < j ? i : j, c = i < j ? j : i;
        double g;
        if (fixed) {
            const unsigned long long *lo = reinterpret_cast<const unsigned long long *>(gram);
            const unsigned long long *hi = lo + (int64_t)H * H;
            g = ldexp((double)hi[(int64_t)r * H + c], 16 - 38) + ldexp((double)lo[(int64_t)r * H + c], -38);
        } else {
            g = gram[(int64_t)r * H + c];
        }
        cov[idx] = (g - n_total * mu[i] * mu[j]) / (n_total - 1.0);
    }
}

__global__ void k_gather_f64(const double *__restrict__ src, const int32_t *__restrict__ idx, int n,
                             double *__restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = src[idx[i]];
}

// ---------------------------------------------------------------------------------------
// K5c: dense fp64 building blocks of the eigen-solver
// ---------------------------------------------------------------------------------------
// O[M x N] (row-major, ldo) = alpha * A * B + beta * P + gamma * R   (P, R nullable, ld = ldo)
// A(i,k) = A[i*sai + k*sak], B(k,j) = B[k*sbk + j*sbj].  With gridDim.z > 1 the K range is
// split and slice z writes its partial product to O + z*M*ldo (epilogue terms must be off).
#define GM_BM 64
#define GM_BN 32
#define GM_BK 16
#define GM_TM 4
#define GM_TN 2
__global__ void __launch_bounds__(256)
k_gemm_f64(int M, int N, int K, const double *__restrict__ A, int64_t sai, int64_t sak, const double *__restrict__ B,
           int64_t sbk, int64_t sbj, double alpha, const double *__restrict__ P, double beta,
           const double *__restrict__ R, double gamma, double *__restrict__ O, int ldo)
{
    __shared__ double As[GM_BK][GM_BM + 2];
    __shared__ double Bs[GM_BK][GM_BN + 2];
    const int tid = threadIdx.x;
    const int tx = tid % (GM_BN / GM_TN);  // 0..15
    const int ty = tid / (GM_BN / GM_TN);  // 0..15
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    int kchunk = (K + gridDim.z - 1) / gridDim.z;
    kchunk = ((kchunk + GM_BK - 1) / GM_BK) * GM_BK;
    const int kbeg = blockIdx.z * kchunk;
    int kend = kbeg + kchunk;
    if (kend > K)
        kend = K;
    double acc[GM_TM][GM_TN];
#pragma unroll
    for (int i = 0; i < GM_TM; ++i)
#pragma unroll
        for (int j = 0; j < GM_TN; ++j)
            acc[i][j] = 0.0;

    for (int k0 = kbeg; k0 < kend; k0 += GM_BK) {
        for (int idx = tid; idx < GM_BM * GM_BK; idx += 256) {
            int mm, kk;
            if (sak == 1) {
                kk = idx % GM_BK;
                mm = idx / GM_BK;
            } else {
                mm = idx % GM_BM;
                kk = idx / GM_BM;
            }
            const int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < kend) ? A[(int64_t)gm * sai + (int64_t)gk * sak] : 0.0;
        }
        for (int idx = tid; idx < GM_BN * GM_BK; idx += 256) {
            int jj, kk;
            if (sbj == 1) {
                jj = idx % GM_BN;
                kk = idx / GM_BN;
            } else {
                kk = idx % GM_BK;
                jj = idx / GM_BK;
            }
            const int gn = n0 + jj, gk = k0 + kk;
            Bs[kk][jj] = (gn < N && gk < kend) ? B[(int64_t)gk * sbk + (int64_t)gn * sbj] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GM_BK; ++kk) {
            double a[GM_TM], b[GM_TN];
#pragma unroll
            for (int i = 0; i < GM_TM; ++i)
                a[i] = As[kk][ty * GM_TM + i];
#pragma unroll
            for (int j = 0; j < GM_TN; ++j)
                b[j] = Bs[kk][tx * GM_TN + j];
#pragma unroll
            for (int i = 0; i < GM_TM; ++i)
#pragma unroll
                for (int j = 0; j < GM_TN; ++j)
                    acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    double *Oz = O + (int64_t)blockIdx.z * M * ldo;
#pragma unroll
    for (int i = 0; i < GM_TM; ++i) {
        const int gm = m0 + ty * GM_TM + i;
        if (gm >= M)
            continue;
#pragma unroll
        for (int j = 0; j < GM_TN; ++j) {
            const int gn = n0 + tx * GM_TN + j;
            if (gn >= N)
                continue;
            const int64_t o = (int64_t)gm * ldo + gn;
            double v = alpha * acc[i][j];
            if (P != nullptr)
                v += beta * P[o];
            if (R != nullptr)
                v += gamma * R[o];
            Oz[o] = v;
        }
    }
}

// out[i] = sum_z part[z*n + i] in fixed order (deterministic split-K reduction)
__global__ void k_sum_partials(const double *__restrict__ part, int64_t n, int nz, double *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        double s = 0.0;
        for (int z = 0; z < nz; ++z)
            s += part[(int64_t)z * n + i];
        out[i] = s;
    }
}

// out[i] = alpha * sum_z part[z*n + i] + beta * P[i] + gamma * R[i]  (split-K epilogue, fixed order)
__global__ void k_combine_partials(const double *__restrict__ part, int64_t n, int nz, double alpha,
                                   const double *__restrict__ P, double beta, const double *__restrict__ R, double gamma,
                                   double *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        double s = 0.0;
        for (int z = 0; z < nz; ++z)
            s += part[(int64_t)z * n + i];
        double v = alpha * s;
        if (P != nullptr)
            v += beta * P[i];
        if (R != nullptr)
            v += gamma * R[i];
        out[i] = v;
    }
}

// deterministic start block: hash -> uniform(-1, 1)
__global__ void k_init_block(double *__restrict__ Q, int64_t n, uint64_t seed)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ULL + seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        Q[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    }
}

// Y = a*W + b*Q, elementwise
__global__ void k_axpby(const double *__restrict__ W, double a, const double *__restrict__ Q, double b, int64_t n,
                        double *__restrict__ Y)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n)
        Y[i] = a * W[i] + b * Q[i];
}

// one block per column j of the row-major H x b block: scale to unit 2-norm
__global__ void __launch_bounds__(256) k_colnorm_scale(double *__restrict__ Y, int H, int b)
{
    __shared__ double sm[8];
    __shared__ double inv;
    const int j = blockIdx.x;
    double s = 0.0;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        double v = Y[(int64_t)h * b + j];
        s += v * v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w)
            t += sm[w];
        inv = t > 0.0 ? 1.0 / sqrt(t) : 0.0;
    }
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += blockDim.x)
        Y[(int64_t)h * b + j] *= inv;
}

// res[j] = || W[:,j] - theta[j] Q[:,j] ||_2
__global__ void __launch_bounds__(256)
k_residual(const double *__restrict__ W, const double *__restrict__ Q, const double *__restrict__ theta, int H, int b,
           double *__restrict__ res)
{
    __shared__ double sm[8];
    const int j = blockIdx.x;
    const double th = theta[j];
    double s = 0.0;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        double v = W[(int64_t)h * b + j] - th * Q[(int64_t)h * b + j];
        s += v * v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w)
            t += sm[w];
        res[j] = sqrt(t);
    }
}

// Cholesky S = L L^T of a b x b SPD matrix, single block, L returned row-major in Lout
// (upper part zeroed).  Pivots below eps*max_diag are floored (rank-deficient blocks).
__global__ void __launch_bounds__(256) k_chol(const double *__restrict__ S, int b, double *__restrict__ Lout,
                                              int *__restrict__ flag)
{
    extern __shared__ double sh[];  // b*b
    __shared__ double piv_floor;
    const int tid = threadIdx.x;
    for (int i = tid; i < b * b; i += blockDim.x) {
        int r = i / b, c = i % b;
        sh[i] = 0.5 * (S[r * b + c] + S[c * b + r]);
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int i = 0; i < b; ++i)
            mx = fmax(mx, sh[i * b + i]);
        piv_floor = mx * 1e-14;
        if (!(mx > 0.0))
            piv_floor = 1e-300;
    }
    __syncthreads();
    for (int j = 0; j < b; ++j) {
        if (tid == 0) {
            double p = sh[j * b + j];
            if (!(p > piv_floor)) {
                p = piv_floor;
                atomicOr(flag, 2);
            }
            sh[j * b + j] = sqrt(p);
        }
        __syncthreads();
        const double ljj = sh[j * b + j];
        for (int i = j + 1 + tid; i < b; i += blockDim.x)
            sh[i * b + j] /= ljj;
        __syncthreads();
        // trailing update of the lower triangle: S[i][k] -= L[i][j] L[k][j], j < k <= i
        const int rem = b - j - 1;
        for (int t = tid; t < rem * rem; t += blockDim.x) {
            const int i = j + 1 + t / rem, k = j + 1 + t % rem;
            if (k <= i)
                sh[i * b + k] -= sh[i * b + j] * sh[k * b + j];
        }
        __syncthreads();
    }
    for (int i = tid; i < b * b; i += blockDim.x) {
        int r = i / b, c = i % b;
        Lout[i] = c <= r ? sh[i] : 0.0;
    }
}

// Q = Y L^{-T}: every row q solves q L^T = y by forward substitution.  Warp per row,
// lane (j % 32) owns q_j; L (row-major, lower) sits in shared memory.
#define TRSM_MAXB 128
__global__ void __launch_bounds__(256) k_trsm_rows(const double *__restrict__ Y, const double *__restrict__ L, int H,
                                                   int b, double *__restrict__ Q)
{
    extern __shared__ double shL[];  // b*b
    for (int i = threadIdx.x; i < b * b; i += blockDim.x)
        shL[i] = L[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int h = warp0; h < H; h += nwarps) {
        double q[TRSM_MAXB / 32];
#pragma unroll
        for (int u = 0; u < TRSM_MAXB / 32; ++u)
            q[u] = 0.0;
        for (int j = 0; j < b; ++j) {
            // sum_{i<j} q_i L[j][i]
            double part = 0.0;
#pragma unroll
            for (int u = 0; u < TRSM_MAXB / 32; ++u) {
                const int i = u * 32 + lane;
                if (i < j)
                    part += q[u] * shL[j * b + i];
            }
            part = warp_sum(part);
            const double qj = (Y[(int64_t)h * b + j] - part) / shL[j * b + j];
#pragma unroll
            for (int u = 0; u < TRSM_MAXB / 32; ++u)
                if (u * 32 + lane == j)
                    q[u] = qj;
        }
#pragma unroll
        for (int u = 0; u < TRSM_MAXB / 32; ++u) {
            const int i = u * 32 + lane;
            if (i < b)
                Q[(int64_t)h * b + i] = q[u];
        }
    }
}

// Eigen-decomposition of the symmetric b x b matrix T by one-sided (Hestenes) Jacobi in a
// single block: A <- T, V <- I, rotate column pairs until they are mutually orthogonal;
// then A = T V, lambda_j = v_j . (T v_j).  Output: theta sorted descending, Y row-major
// b x b with the matching eigenvectors in its columns.
__global__ void __launch_bounds__(1024) k_jacobi_eig(const double *__restrict__ T, int b, double *__restrict__ theta,
                                                     double *__restrict__ Yout, int *__restrict__ sweeps_out)
{
    extern __shared__ double sh[];
    const int ld = b | 1;
    double *A = sh;               // column-major, A[i + p*ld]
    double *V = sh + (size_t)ld * b;
    double *lam = V + (size_t)ld * b;  // b
    __shared__ int rotated;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    for (int i = tid; i < b * b; i += blockDim.x) {
        int r = i % b, c = i / b;
        A[r + c * ld] = 0.5 * (T[r * b + c] + T[c * b + r]);
        V[r + c * ld] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int np = (b + 1) & ~1;  // players (one dummy when b is odd)
    const double eps = 1e-12;  // relative off-diagonal threshold: eigenvectors good to ~1e-12
    // half-warp per column pair (all np/2 <= 64 pairs of a round at once); the squared column norms are
    // kept in shared memory and updated in closed form by every rotation (|a_p'|^2 = |a_p|^2 - t g,
    // |a_q'|^2 = |a_q|^2 + t g), recomputed at the start of every sweep, so a pair costs ONE reduction
    double *nrm = lam;   // b entries, free until the eigenvalues are formed below
    const int hw = tid >> 4, hl = tid & 15, nhw = blockDim.x >> 4;
    const unsigned hmask = 0xffffu << (16 * ((tid >> 4) & 1));
    int sweep = 0;
    for (; sweep < 40; ++sweep) {
        if (tid == 0)
            rotated = 0;
        for (int j = hw; j < b; j += nhw) {
            double s2 = 0.0;
            for (int i = hl; i < b; i += 16) {
                const double x = A[i + (size_t)j * ld];
                s2 += x * x;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1)
                s2 += __shfl_xor_sync(hmask, s2, o);
            if (hl == 0)
                nrm[j] = s2;
        }
        __syncthreads();
        for (int r = 0; r < np - 1; ++r) {
            for (int pr = hw; pr < np / 2; pr += nhw) {
                int p, q;
                if (pr == 0) {
                    p = np - 1;
                    q = r;
                } else {
                    p = (r + pr) % (np - 1);
                    q = (r - pr + (np - 1)) % (np - 1);
                }
                if (p > q) {
                    int t = p;
                    p = q;
                    q = t;
                }
                if (q >= b)
                    continue;  // dummy player
                double *ap = A + (size_t)p * ld, *aq = A + (size_t)q * ld;
                double ga = 0.0;
                for (int i = hl; i < b; i += 16)
                    ga += ap[i] * aq[i];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1)
                    ga += __shfl_xor_sync(hmask, ga, o);
                const double al = nrm[p], be = nrm[q];
                if (fabs(ga) > eps * sqrt(al * be) && fabs(ga) > 1e-300) {
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    double *vp = V + (size_t)p * ld, *vq = V + (size_t)q * ld;
                    for (int i = hl; i < b; i += 16) {
                        const double x = ap[i], y = aq[i];
                        ap[i] = c * x - s * y;
                        aq[i] = s * x + c * y;
                        const double u = vp[i], v = vq[i];
                        vp[i] = c * u - s * v;
                        vq[i] = s * u + c * v;
                    }
                    if (hl == 0) {
                        nrm[p] = al - t * ga;
                        nrm[q] = be + t * ga;
                        rotated = 1;
                    }
                }
            }
            __syncthreads();
        }
        const int any = rotated;
        __syncthreads();
        if (!any)
            break;
    }
    for (int j = w; j < b; j += nw) {
        double s = 0.0;
        for (int i = lane; i < b; i += 32)
            s += A[i + (size_t)j * ld] * V[i + (size_t)j * ld];
        s = warp_sum(s);
        if (lane == 0)
            lam[j] = s;
    }
    __syncthreads();
    for (int j = w; j < b; j += nw) {
        const double lj = lam[j];
        int rank = 0;
        for (int l = lane; l < b; l += 32) {
            const double ll = lam[l];
            rank += (ll > lj || (ll == lj && l < j)) ? 1 : 0;
        }
        rank = warp_sum_i(rank);
        if (lane == 0)
            theta[rank] = lj;
        for (int i = lane; i < b; i += 32)
            Yout[i * b + rank] = V[i + (size_t)j * ld];
    }
    if (tid == 0 && sweeps_out != nullptr)
        *sweeps_out = sweep;
}

// rot[h*d + j] = Q[h*b + j] (j < d) with a deterministic sign: the entry of largest
// magnitude in every column is made positive.  One block per column.
__global__ void __launch_bounds__(256) k_extract_rotation(const double *__restrict__ Q, int H, int b, int d,
                                                          double *__restrict__ rot)
{
    __shared__ double smv[256];
    __shared__ int smi[256];
    const int j = blockIdx.x;
    double best = -1.0;
    int bi = 0;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        double a = fabs(Q[(int64_t)h * b + j]);
        if (a > best) {
            best = a;
            bi = h;
        }
    }
    smv[threadIdx.x] = best;
    smi[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            double o = smv[threadIdx.x + s];
            int oi = smi[threadIdx.x + s];
            if (o > smv[threadIdx.x] || (o == smv[threadIdx.x] && oi < smi[threadIdx.x])) {
                smv[threadIdx.x] = o;
                smi[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    const double sign = Q[(int64_t)smi[0] * b + j] < 0.0 ? -1.0 : 1.0;
    for (int h = threadIdx.x; h < H; h += blockDim.x)
        rot[(int64_t)h * d + j] = sign * Q[(int64_t)h * b + j];
}

// muv[j] = sum_h mu[h] rot[h*d + j]
__global__ void __launch_bounds__(256) k_mu_dot_rot(const double *__restrict__ mu, const double *__restrict__ rot, int H,
                                                    int d, double *__restrict__ muv)
{
    __shared__ double sm[8];
    const int j = blockIdx.x;
    double s = 0.0;
    for (int h = threadIdx.x; h < H; h += blockDim.x)
        s += mu[h] * rot[(int64_t)h * d + j];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w)
            t += sm[w];
        muv[j] = t;
    }
}

// ---------------------------------------------------------------------------------------
// K5d: scores[c][j] = sum_a y_a rot[slot_a][j] - muv[j]; warp per cell, lanes over j
// ---------------------------------------------------------------------------------------
#define PROJ_MAXD 128
__global__ void __launch_bounds__(256)
k_project(const int64_t *__restrict__ hptr, const int32_t *__restrict__ hslot, const double *__restrict__ hval,
          int64_t n_cols, const double *__restrict__ rot, const double *__restrict__ muv, int d,
          double *__restrict__ scores)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = warp0; c < n_cols; c += nwarps) {
        const int64_t b = hptr[c], e = hptr[c + 1];
        double acc[PROJ_MAXD / 32];
#pragma unroll
        for (int u = 0; u < PROJ_MAXD / 32; ++u)
            acc[u] = 0.0;
        for (int64_t a = b; a < e; ++a) {
            const int s = hslot[a];
            const double v = hval[a];
            const double *r = rot + (int64_t)s * d;
#pragma unroll
            for (int u = 0; u < PROJ_MAXD / 32; ++u) {
                const int j = u * 32 + lane;
                if (j < d)
                    acc[u] = fma(v, r[j], acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < PROJ_MAXD / 32; ++u) {
            const int j = u * 32 + lane;
            if (j < d)
                scores[c * d + j] = acc[u] - muv[j];
        }
    }
}

// The same with the rotation blocked through shared memory: a CTA owns a contiguous range of cells and
// walks the HVG slots slab by slab (a slab = as many 64-slot blocks as fit ~200 KB of V rows); the entries
// of a cell that fall into the slab are found through the block pointers.  V is then read from shared
// memory (400 B per entry at d = 50) instead of L2 -- k_project runs at the L2 bandwidth.  The partial
// sums of a cell are added slab after slab by the same warp, so the result is deterministic.
template <int NU>   // NU = ceil(d / 32) accumulators per lane
__global__ void __launch_bounds__(1024)
k_project_slab(const int64_t *__restrict__ hptr, const int32_t *__restrict__ hslot, const double *__restrict__ hval,
               const uint16_t *__restrict__ bptr, int nb, int64_t n_cols, const double *__restrict__ rot,
               const double *__restrict__ muv, int H, int d, int blocks_per_slab, double *__restrict__ scores)
{
    extern __shared__ __align__(16) double vs[];   // [slab slots][d]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t per = (n_cols + gridDim.x - 1) / gridDim.x;
    const int64_t c0 = blockIdx.x * per, c1 = c0 + per < n_cols ? c0 + per : n_cols;
    for (int blk0 = 0; blk0 < nb; blk0 += blocks_per_slab) {
        const int blk1 = blk0 + blocks_per_slab < nb ? blk0 + blocks_per_slab : nb;
        const int s0 = blk0 * GB_BS, s1 = blk1 * GB_BS < H ? blk1 * GB_BS : H;
        __syncthreads();
        for (int i = threadIdx.x; i < (s1 - s0) * d; i += blockDim.x)
            vs[i] = rot[(int64_t)s0 * d + i];
        __syncthreads();
        for (int64_t c = c0 + w; c < c1; c += (blockDim.x >> 5)) {
            const int64_t b = hptr[c];
            const int e0 = bptr[c * (nb + 1) + blk0], e1 = bptr[c * (nb + 1) + blk1];
            double acc[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u)
                acc[u] = 0.0;
            for (int a0 = e0; a0 < e1; a0 += 32) {
                const int a = a0 + lane;
                int sl = 0;
                double v = 0.0;
                if (a < e1) {
                    sl = hslot[b + a] - s0;
                    v = hval[b + a];
                }
                const int cnt = e1 - a0 < 32 ? e1 - a0 : 32;
#pragma unroll 4
                for (int q = 0; q < cnt; ++q) {
                    const int sq = __shfl_sync(0xffffffffu, sl, q);
                    const double vq = __shfl_sync(0xffffffffu, v, q);
                    const double *r = vs + sq * d;
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        const int j = u * 32 + lane;
                        if (j < d)
                            acc[u] = fma(vq, r[j], acc[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int j = u * 32 + lane;
                if (j < d) {
                    const double prev = blk0 == 0 ? -muv[j] : scores[c * d + j];
                    scores[c * d + j] = prev + acc[u];
                }
            }
        }
    }
}

// row-major [rows x cols] -> column-major [rows x cols]
__global__ void k_rowmajor_to_colmajor(const double *__restrict__ in, int64_t rows, int cols, double *__restrict__ out)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < rows * cols) {
        int64_t r = idx % rows;
        int c = (int)(idx / rows);
        out[idx] = in[r * cols + c];
    }
}

// ---------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------
static int launch_gemm(cb200_ctx *ctx, int M, int N, int K, const double *A, int64_t sai, int64_t sak,
                       const double *B, int64_t sbk, int64_t sbj, double alpha, const double *P, double beta,
                       const double *R, double gamma, double *O, int ldo, int ksplit = 1)
{
    dim3 grid((N + GM_BN - 1) / GM_BN, (M + GM_BM - 1) / GM_BM, ksplit);
    k_gemm_f64<<<grid, 256, 0, ctx->stream>>>(M, N, K, A, sai, sak, B, sbk, sbj, alpha, P, beta, R, gamma, O, ldo);
    CB_LAUNCH(ctx);
    return 0;
}

// O(H x b) = alpha C Q + beta P + gamma R with the H x H covariance: split over K so that every SM has work
// (the plain launch has (b/32) x (H/64) = 64 CTAs for H = 2000, b = 64), deterministic reduction
static int launch_cq(cb200_ctx *ctx, int H, int b, const double *C, const double *Q, double alpha, const double *P,
                     double beta, const double *R, double gamma, double *O)
{
    const int ksplit = 6;
    CB_CUDA(ctx, ctx->egs.ensure((size_t)ksplit * H * b));
    CB_TRY(launch_gemm(ctx, H, b, H, C, H, 1, Q, b, 1, 1.0, nullptr, 0.0, nullptr, 0.0, ctx->egs.p, b, ksplit));
    k_combine_partials<<<(unsigned)(((int64_t)H * b + 255) / 256), 256, 0, ctx->stream>>>(ctx->egs.p, (int64_t)H * b, ksplit,
                                                                                        alpha, P, beta, R, gamma, O);
    CB_LAUNCH(ctx);
    return 0;
}

// S(b x b) = A^T B for row-major H x b blocks, deterministic split-K
static int launch_atb(cb200_ctx *ctx, const double *A, const double *B, int H, int b, double *S)
{
    const int ksplit = 32;
    CB_CUDA(ctx, ctx->es.ensure((size_t)ksplit * b * b));
    CB_TRY(launch_gemm(ctx, b, b, H, A, 1, b, B, b, 1, 1.0, nullptr, 0.0, nullptr, 0.0, ctx->es.p, b, ksplit));
    k_sum_partials<<<(b * b + 255) / 256, 256, 0, ctx->stream>>>(ctx->es.p, (int64_t)b * b, ksplit, S);
    CB_LAUNCH(ctx);
    return 0;
}

// Y <- orthonormal basis of span(Y) by Cholesky-QR; result in Qout (Y is overwritten as scratch)
static int chol_qr(cb200_ctx *ctx, double *Y, int H, int b, double *Qout)
{
    const size_t shb = (size_t)b * b * sizeof(double);
    CB_TRY(launch_atb(ctx, Y, Y, H, b, ctx->et.p));
    k_chol<<<1, 256, shb, ctx->stream>>>(ctx->et.p, b, ctx->et.p + (size_t)b * b, ctx->flag.p + 1);
    CB_LAUNCH(ctx);
    k_trsm_rows<<<ctx->num_sms, 256, shb, ctx->stream>>>(Y, ctx->et.p + (size_t)b * b, H, b, Qout);
    CB_LAUNCH(ctx);
    return 0;
}

extern "C" int cb200_pca_gram_local(cb200_ctx *ctx, const int32_t *hvg_idx, int n_hvg)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->have_logx && ctx->have_genestats,
             "cb200_pca: logcounts and gene statistics must be computed first");
    CB_CHECK(ctx, hvg_idx != nullptr && n_hvg >= 2, "cb200_pca: need at least 2 HVGs");
    CB_CHECK(ctx, n_hvg <= 32767, "cb200_pca: at most 32767 HVGs are supported (got %d)", n_hvg);
    {
        std::vector<char> seen((size_t)ctx->G, 0);
        for (int h = 0; h < n_hvg; ++h) {
            CB_CHECK(ctx, hvg_idx[h] >= 0 && hvg_idx[h] < ctx->G, "cb200_pca: hvg_idx[%d]=%d out of range", h,
                     hvg_idx[h]);
            CB_CHECK(ctx, !seen[hvg_idx[h]], "cb200_pca: hvg_idx contains gene %d twice", hvg_idx[h]);
            seen[hvg_idx[h]] = 1;
        }
    }
    CB_TRACE("gram_local: enter");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_TRACE("gram_local: setdevice");
    const int H = n_hvg;
    ctx->H = H;
    // Device-side slots are numbered in ascending gene order (entries of a cell then come sorted by
    // slot); perm[h] = slot of the caller's h-th HVG, used when the rotation is handed back.
    std::vector<int32_t> order((size_t)H), sorted_genes((size_t)H), perm((size_t)H);
    for (int h = 0; h < H; ++h)
        order[h] = h;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return hvg_idx[a] < hvg_idx[b]; });
    for (int s = 0; s < H; ++s) {
        sorted_genes[s] = hvg_idx[order[s]];
        perm[order[s]] = s;
    }
    CB_CUDA(ctx, ctx->hvg_perm.ensure((size_t)H));
    CB_CUDA(ctx, ctx->hvg_idx.ensure((size_t)H));
    CB_CUDA(ctx, ctx->slot.ensure((size_t)ctx->G));
    CB_CUDA(ctx, ctx->mu.ensure((size_t)H));
    CB_CUDA(ctx, ctx->hcnt.ensure((size_t)ctx->N));
    CB_CUDA(ctx, ctx->hptr.ensure((size_t)ctx->N + 1));
    CB_CUDA(ctx, ctx->gram.ensure((size_t)2 * H * H));   // two 64-bit words per entry for the integer engine

    CB_TRY(cb200_need_rowidx(ctx));
    cb_stage_begin(ctx, ST_HVG_EXTRACT);
    CB_TRY(cb200_write_small(ctx, ctx->hvg_idx.p, sorted_genes.data(), (size_t)H * sizeof(int32_t)));
    CB_TRY(cb200_write_small(ctx, ctx->hvg_perm.p, perm.data(), (size_t)H * sizeof(int32_t)));
    CB_TRACE("gram_local: h2d synced");
    k_fill_i32<<<(unsigned)((ctx->G + 255) / 256), 256, 0, ctx->stream>>>(ctx->slot.p, ctx->G, -1);
    CB_LAUNCH(ctx);
    k_build_slot<<<(H + 255) / 256, 256, 0, ctx->stream>>>(ctx->hvg_idx.p, H, ctx->slot.p, ctx->gene_mean.p,
                                                          ctx->mu.p);
    CB_LAUNCH(ctx);
    const int grid = cb_grid_for(ctx->N, 8, ctx->num_sms * 8);
    k_hvg_count<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->rowidx.p, ctx->slot.p, ctx->N, ctx->hcnt.p);
    CB_LAUNCH(ctx);
    CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->hcnt.p, ctx->N, ctx->hptr.p));
    int64_t nnz_hvg = 0;
    CB_TRY(cb200_read_back(ctx, &nnz_hvg, ctx->hptr.p + ctx->N, sizeof(int64_t)));
    ctx->nnz_hvg = nnz_hvg;
    CB_TRACE("gram_local: nnz_hvg read");
    CB_CUDA(ctx, ctx->hslot.ensure((size_t)nnz_hvg));
    CB_CUDA(ctx, ctx->hval.ensure((size_t)nnz_hvg));
    k_hvg_fill<<<grid, 256, 0, ctx->stream>>>(ctx->colptr.p, ctx->rowidx.p, ctx->logx.p, ctx->slot.p, ctx->N,
                                              ctx->hptr.p, ctx->hslot.p, ctx->hval.p);
    CB_LAUNCH(ctx);
    const int nb = (H + GB_BS - 1) / GB_BS;
    CB_CUDA(ctx, ctx->hbptr.ensure((size_t)ctx->N * (nb + 1)));
    k_hvg_blockptr<<<grid, 256, 0, ctx->stream>>>(ctx->hptr.p, ctx->hslot.p, ctx->N, nb, ctx->hbptr.p);
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_HVG_EXTRACT);
    CB_TRACE("gram_local: hvg stage launched");

    cb_stage_begin(ctx, ST_GRAM);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->gram.p, 0, (size_t)2 * H * H * sizeof(double), ctx->stream));
    // engine: tensor cores (error-free int8 digit products) by default; CB200_GRAM_ENGINE=tiled|atomic
    // selects the exact fp64 kernels (also used when a value is outside the fixed-point range)
    const char *geng = getenv("CB200_GRAM_ENGINE");
    int used_tc = 0;
    // (the choice must not depend on the shard: N_total, not N, so that every rank of a multi-GPU job agrees)
    if ((geng == nullptr && ctx->N_total >= 1024) || (geng != nullptr && strcmp(geng, "tc") == 0))
        CB_TRY(cb200_gram_tc(ctx, nb, &used_tc));
    ctx->tm.gram_engine = used_tc ? 3 : ((geng != nullptr && strcmp(geng, "atomic") == 0) ? 1 : 2);
    ctx->gram_fixed = used_tc != 0;
    if (used_tc) {
        // done
    } else if (geng != nullptr && strcmp(geng, "atomic") == 0) {
        // v1 (kept as a cross-check): one fp64 L2 reduction per product
        k_gram_sparse<<<grid, 256, 0, ctx->stream>>>(ctx->hptr.p, ctx->hslot.p, ctx->hval.p, ctx->N, H, ctx->gram.p);
        CB_LAUNCH(ctx);
    } else {
        // a cell holds at most 65535 HVG entries by construction (H <= 32767)
        const int n_tiles = nb * (nb + 1) / 2;
        int ysplit = (int)((4LL * ctx->num_sms + n_tiles - 1) / n_tiles);  // >= 4 CTAs per SM in flight
        const int64_t max_split = (ctx->N + 255) / 256;
        if (ysplit > max_split)
            ysplit = (int)max_split;
        if (ysplit < 1)
            ysplit = 1;
        dim3 ggrid((unsigned)n_tiles, (unsigned)ysplit);
        k_gram_tiled<<<ggrid, 256, 0, ctx->stream>>>(ctx->hptr.p, ctx->hslot.p, ctx->hval.p, ctx->hbptr.p, ctx->N, nb,
                                                     H, ctx->gram.p);
        CB_LAUNCH(ctx);
    }
    cb_stage_end(ctx, ST_GRAM);
    CB_TRACE("gram_local: gram done");
    ctx->have_gram = true;
    ctx->have_scores = false;
    return 0;
}

extern "C" int cb200_pca_finish(cb200_ctx *ctx, int d, double *scores, double *rotation, double *var_explained,
                                double *total_var)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->have_gram, "cb200_pca_finish: call cb200_pca_gram_local first");
    const int H = ctx->H;
    const int64_t Nt = ctx->N_total;
    CB_CHECK(ctx, d >= 1, "cb200_pca: ncomponents must be positive");
    const int max_rank = (int)((int64_t)H < Nt - 1 ? (int64_t)H : Nt - 1);
    CB_CHECK(ctx, d <= max_rank, "cb200_pca: ncomponents=%d exceeds min(n_hvg, n_cells-1)=%d", d, max_rank);
    CB_CHECK(ctx, d <= 88, "cb200_pca: at most 88 components are supported by the device eigen-solver (got %d)", d);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));

    // block size: d plus a guard band that sets the convergence rate
    int b = d + (d / 2 > 16 ? d / 2 : 16);
    b = (b + 7) & ~7;
    if (b > 104)
        b = 104;
    if (b > max_rank)
        b = max_rank;
    if (b < d)
        b = d;
    ctx->d = d;
    ctx->blk = b;
    const int64_t hb = (int64_t)H * b;

    CB_CUDA(ctx, ctx->cov.ensure((size_t)H * H));
    CB_CUDA(ctx, ctx->eq.ensure((size_t)hb));
    CB_CUDA(ctx, ctx->ew.ensure((size_t)hb));
    CB_CUDA(ctx, ctx->ey.ensure((size_t)hb));
    CB_CUDA(ctx, ctx->ex.ensure((size_t)hb));
    CB_CUDA(ctx, ctx->et.ensure((size_t)2 * b * b + 16));
    CB_CUDA(ctx, ctx->etheta.ensure((size_t)b));
    CB_CUDA(ctx, ctx->eres.ensure((size_t)b));
    CB_CUDA(ctx, ctx->emuv.ensure((size_t)d));
    CB_CUDA(ctx, ctx->rot.ensure((size_t)H * d));
    CB_CUDA(ctx, ctx->varexp.ensure((size_t)d));
    CB_CUDA(ctx, ctx->scores.ensure((size_t)ctx->N * d));
    CB_CUDA(ctx, ctx->stage_d.ensure((size_t)(ctx->N * d > (int64_t)H * d ? ctx->N * d : (int64_t)H * d)));

    const size_t jac_smem = ((size_t)(b | 1) * b * 2 + b) * sizeof(double);
    const size_t chol_smem = (size_t)b * b * sizeof(double);
    CB_CUDA(ctx, cudaFuncSetAttribute(k_jacobi_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jac_smem));
    CB_CUDA(ctx, cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem));
    CB_CUDA(ctx, cudaFuncSetAttribute(k_trsm_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem));

    cb_stage_begin(ctx, ST_EIG);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 1, 0, sizeof(int), ctx->stream));
    k_cov_from_gram<<<(unsigned)(((int64_t)H * H + 255) / 256), 256, 0, ctx->stream>>>(ctx->gram.p, ctx->mu.p, H,
                                                                                      (double)Nt, ctx->cov.p,
                                                                                      ctx->gram_fixed ? 1 : 0);
    CB_LAUNCH(ctx);
    // total variance of the HVG block = sum of the K3 variances of the selected genes
    k_gather_f64<<<(H + 255) / 256, 256, 0, ctx->stream>>>(ctx->gene_var.p, ctx->hvg_idx.p, H, ctx->ex.p);
    CB_LAUNCH(ctx);
    {
        std::vector<double> hv((size_t)H);
        CB_TRY(cb200_read_back(ctx, hv.data(), ctx->ex.p, (size_t)H * sizeof(double)));
        double tv = 0.0;
        for (int h = 0; h < H; ++h)
            tv += hv[h];
        ctx->total_var = tv;
    }

    double *Q = ctx->eq.p, *W = ctx->ew.p, *Y1 = ctx->ey.p, *Y2 = ctx->ex.p;
    const double *C = ctx->cov.p;
    const unsigned ew_grid = (unsigned)((hb + 255) / 256);
    k_init_block<<<ew_grid, 256, 0, ctx->stream>>>(Y1, hb, 0x5EEDULL);
    CB_LAUNCH(ctx);
    k_colnorm_scale<<<b, 256, 0, ctx->stream>>>(Y1, H, b);
    CB_LAUNCH(ctx);
    CB_TRY(chol_qr(ctx, Y1, H, b, Y2));
    CB_TRY(chol_qr(ctx, Y2, H, b, Q));

    const int max_outer = 300;
    const double tol = 2e-9;
    const bool trace_sweeps = getenv("CB200_TRACE") != nullptr;
    int degree_hi = 6;  // Chebyshev filter degree between Rayleigh-Ritz steps once the block has settled (CB200_EIG_DEGREE)
    if (const char *ed = getenv("CB200_EIG_DEGREE")) {
        const int v = atoi(ed);
        if (v >= 2 && v <= 16)
            degree_hi = v;
    }
    std::vector<double> th((size_t)b), res((size_t)b);
    int it = 0;
    bool converged = false;
    for (; it < max_outer; ++it) {
        // Rayleigh-Ritz: W = C Q, T = Q^T W, T = Y Theta Y^T, Q <- Q Y, W <- W Y
        CB_TRY(launch_cq(ctx, H, b, C, Q, 1.0, nullptr, 0.0, nullptr, 0.0, W));
        CB_TRY(launch_atb(ctx, Q, W, H, b, ctx->et.p));
        k_jacobi_eig<<<1, 1024, jac_smem, ctx->stream>>>(ctx->et.p, b, ctx->etheta.p, ctx->et.p + (size_t)b * b,
                                                         trace_sweeps ? ctx->flag.p + 1 : nullptr);
        CB_LAUNCH(ctx);
        if (trace_sweeps) {
            int sw = 0;
            CB_TRY(cb200_read_back(ctx, &sw, ctx->flag.p + 1, sizeof(int)));
            fprintf(stderr, "[cb200] eig iteration %d: %d Jacobi sweeps (b = %d)\n", it, sw, b);
            CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 1, 0, sizeof(int), ctx->stream));
        }
        const double *Yr = ctx->et.p + (size_t)b * b;
        CB_TRY(launch_gemm(ctx, H, b, b, Q, b, 1, Yr, b, 1, 1.0, nullptr, 0.0, nullptr, 0.0, Y1, b));
        CB_TRY(launch_gemm(ctx, H, b, b, W, b, 1, Yr, b, 1, 1.0, nullptr, 0.0, nullptr, 0.0, Y2, b));
        // now Y1 = Ritz vectors, Y2 = C * Ritz vectors
        k_residual<<<b, 256, 0, ctx->stream>>>(Y2, Y1, ctx->etheta.p, H, b, ctx->eres.p);
        CB_LAUNCH(ctx);
        CB_TRY(cb200_read_back(ctx, th.data(), ctx->etheta.p, (size_t)b * sizeof(double)));
        CB_TRY(cb200_read_back(ctx, res.data(), ctx->eres.p, (size_t)b * sizeof(double)));
        double worst = 0.0;
        for (int j = 0; j < d; ++j)
            worst = res[j] > worst ? res[j] : worst;
        CB_CHECK(ctx, th[0] == th[0] && worst == worst, "cb200_pca: eigen-solver produced NaN (iteration %d)", it);
        if (worst <= tol * fabs(th[0]) || !(th[0] > 0.0)) {
            converged = true;
            break;
        }
        // Chebyshev filter damping [0, cut]; cut = smallest Ritz value of the block.  Degree schedule: 3 while the
        // Ritz vectors are still rough, then `degree_hi` (default 6): the products are cheap next to the
        // Rayleigh-Ritz / QR work of an outer iteration, and once the columns are close to eigenvectors every column is
        // amplified by its own factor, so the column-normalised block stays well conditioned for Cholesky-QR.
        double cut = th[b - 1];
        if (!(cut > 1e-12 * th[0]))
            cut = 1e-12 * th[0];
        const double cc = 0.5 * cut, ee = 0.5 * cut;
        const int degree = it < 2 ? (degree_hi < 3 ? degree_hi : 3) : degree_hi;
        // X_k lives in bufs[k mod 4]: X0 = Ritz vectors (Y1), X1 -> W, X2 -> Q, X3 -> Y2, X4 -> Y1, ...
        double *bufs[4] = {Y1, W, Q, Y2};
        // X1 = (C X0 - c X0) / e  (C X0 is the Ritz block Y2)
        k_axpby<<<ew_grid, 256, 0, ctx->stream>>>(Y2, 1.0 / ee, Y1, -cc / ee, hb, W);
        CB_LAUNCH(ctx);
        for (int m = 2; m <= degree; ++m) {
            // X_m = 2 (C X_{m-1} - c X_{m-1}) / e - X_{m-2}
            double *xm1 = bufs[(m - 1) & 3], *xm2 = bufs[(m - 2) & 3], *xm = bufs[m & 3];
            CB_TRY(launch_cq(ctx, H, b, C, xm1, 2.0 / ee, xm1, -2.0 * cc / ee, xm2, -1.0, xm));
        }
        double *filtered = bufs[degree & 3];
        k_colnorm_scale<<<b, 256, 0, ctx->stream>>>(filtered, H, b);
        CB_LAUNCH(ctx);
        double *tmp = filtered == W ? Y1 : W;
        CB_TRY(chol_qr(ctx, filtered, H, b, tmp));
        CB_TRY(chol_qr(ctx, tmp, H, b, Q));
        if (degree >= 3) {
            // a floored Cholesky pivot means the filtered block was too ill-conditioned for Cholesky-QR:
            // continue with a gentler filter (6 -> 3 -> 2)
            int h_piv = 0;
            CB_TRY(cb200_read_back(ctx, &h_piv, ctx->flag.p + 1, sizeof(int)));
            if (h_piv != 0) {
                degree_hi = degree_hi > 3 ? 3 : 2;
                CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 1, 0, sizeof(int), ctx->stream));
            }
        }
    }
    ctx->tm.eig_iterations = it + (converged ? 1 : 0);
    // Ritz vectors of the last Rayleigh-Ritz step are in Y1, values in th
    k_extract_rotation<<<d, 256, 0, ctx->stream>>>(Y1, H, b, d, ctx->rot.p);
    CB_LAUNCH(ctx);
    CB_CUDA(ctx, cudaMemcpyAsync(ctx->varexp.p, ctx->etheta.p, (size_t)d * sizeof(double), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    k_mu_dot_rot<<<d, 256, 0, ctx->stream>>>(ctx->mu.p, ctx->rot.p, H, d, ctx->emuv.p);
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_EIG);

    cb_stage_begin(ctx, ST_PROJECT);
    {
        // V slab in shared memory: as many 64-slot blocks as fit 200 KB (CB200_PROJECT_ENGINE=l2: V from L2)
        const int nbp = (H + GB_BS - 1) / GB_BS;
        int bps = (int)((200 * 1024) / ((size_t)GB_BS * d * sizeof(double)));
        const char *peng = getenv("CB200_PROJECT_ENGINE");
        if (bps >= 1 && !(peng != nullptr && strcmp(peng, "l2") == 0)) {
            if (bps > nbp)
                bps = nbp;
            const size_t smem = (size_t)bps * GB_BS * d * sizeof(double);
            const int pgrid = (int)(ctx->N < ctx->num_sms ? ctx->N : ctx->num_sms);
#define CB_PROJECT_SLAB(NU)                                                                                          \
    do {                                                                                                             \
        CB_CUDA(ctx, cudaFuncSetAttribute(k_project_slab<NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_project_slab<NU><<<pgrid, 1024, smem, ctx->stream>>>(ctx->hptr.p, ctx->hslot.p, ctx->hval.p, ctx->hbptr.p, \
                                                               nbp, ctx->N, ctx->rot.p, ctx->emuv.p, H, d, bps,      \
                                                               ctx->scores.p);                                       \
    } while (0)
            if (d <= 32)
                CB_PROJECT_SLAB(1);
            else if (d <= 64)
                CB_PROJECT_SLAB(2);
            else
                CB_PROJECT_SLAB(4);
#undef CB_PROJECT_SLAB
        } else {
            const int grid = cb_grid_for(ctx->N, 8, ctx->num_sms * 8);
            k_project<<<grid, 256, 0, ctx->stream>>>(ctx->hptr.p, ctx->hslot.p, ctx->hval.p, ctx->N, ctx->rot.p,
                                                     ctx->emuv.p, d, ctx->scores.p);
        }
    }
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_PROJECT);
    ctx->have_scores = converged;   // scores of an unconverged solve are never handed to cb200_knn(NULL)

    if (scores != nullptr || rotation != nullptr)
        cb_stage_begin(ctx, ST_D2H);
    if (scores != nullptr) {
        k_rowmajor_to_colmajor<<<(unsigned)((ctx->N * d + 255) / 256), 256, 0, ctx->stream>>>(ctx->scores.p, ctx->N, d,
                                                                                             ctx->stage_d.p);
        CB_LAUNCH(ctx);
        CB_TRY(cb200_read_back(ctx, scores, ctx->stage_d.p, (size_t)ctx->N * d * sizeof(double)));
    }
    if (rotation != nullptr) {
        k_rotation_out<<<(unsigned)(((int64_t)H * d + 255) / 256), 256, 0, ctx->stream>>>(ctx->rot.p, ctx->hvg_perm.p,
                                                                                        H, d, ctx->stage_d.p);
        CB_LAUNCH(ctx);
        CB_TRY(cb200_read_back(ctx, rotation, ctx->stage_d.p, (size_t)H * d * sizeof(double)));
    }
    if (scores != nullptr || rotation != nullptr)
        cb_stage_end(ctx, ST_D2H);
    if (var_explained != nullptr) {
        for (int j = 0; j < d; ++j)
            var_explained[j] = th[j];
    }
    if (total_var != nullptr)
        *total_var = ctx->total_var;
    int h_flag = 0;
    CB_TRY(cb200_read_back(ctx, &h_flag, ctx->flag.p + 1, sizeof(int)));
    CB_CHECK(ctx, converged, "cb200_pca: eigen-solver did not converge in %d iterations (worst residual above %g)",
             max_outer, tol);
    (void)h_flag;  // floored Cholesky pivots only occur for rank-deficient blocks; residual test is the judge
    return 0;
}

extern "C" int cb200_pca(cb200_ctx *ctx, const int32_t *hvg_idx, int n_hvg, int d, double *scores, double *rotation,
                         double *var_explained, double *total_var)
{
    if (ctx == nullptr)
        return 1;
    CB_TRY(cb200_pca_gram_local(ctx, hvg_idx, n_hvg));
    return cb200_pca_finish(ctx, d, scores, rotation, var_explained, total_var);
}
