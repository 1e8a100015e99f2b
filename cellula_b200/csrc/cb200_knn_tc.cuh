// cb200_knn_tc.cuh -- pieces shared by the SIMT (cb200_knn.cu) and tensor-core (cb200_knn_tc.cu)
// kNN tile engines: the warp-wide bitonic sort of (key, index) pairs and the entry points of
// the tensor-core engine.
#pragma once

#include "cb200_common.cuh"

#ifdef __CUDACC__
template <typename K>
__device__ __forceinline__ bool pair_less(K ka, int ia, K kb, int ib)
{
    return ka < kb || (ka == kb && ia < ib);
}

// warp-wide bitonic sort of 32*EPL (key, idx) pairs, element i = e*32 + lane, ascending by
// (key, idx)
template <typename K, int EPL>
__device__ __forceinline__ void warp_bitonic_sort(K (&key)[EPL], int (&idx)[EPL], int lane)
{
    constexpr int NTOT = 32 * EPL;
#pragma unroll
    for (int k = 2; k <= NTOT; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int dj = j >> 5;
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    const int pe = e ^ dj;
                    if (pe > e) {
                        const bool up = (((e * 32 + lane) & k) == 0);
                        const bool gt = pair_less<K>(key[pe], idx[pe], key[e], idx[e]);  // v[e] > v[pe]
                        if (gt == up) {
                            K tk = key[e];
                            key[e] = key[pe];
                            key[pe] = tk;
                            int ti = idx[e];
                            idx[e] = idx[pe];
                            idx[pe] = ti;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    const bool up = (((e * 32 + lane) & k) == 0);
                    const K ok = __shfl_xor_sync(0xffffffffu, key[e], j);
                    const int oi = __shfl_xor_sync(0xffffffffu, idx[e], j);
                    const bool lower = ((lane & j) == 0);
                    const bool other_less = pair_less<K>(ok, oi, key[e], idx[e]);
                    // the lower lane of an ascending pair keeps the minimum
                    const bool take_other = (lower == up) ? other_less : !other_less;
                    if (take_other) {
                        key[e] = ok;
                        idx[e] = oi;
                    }
                }
            }
        }
    }
}
#endif

// tensor-core tile engine (cb200_knn_tc.cu)
bool cb200_knn_tc_supported(int d, int k);
int cb200_knn_tiles_tc(cb200_ctx *ctx, const double *E, int64_t lde, int64_t n, int d, int k, int64_t q_begin,
                       int64_t q_end, int *kp_out, int *ntau_out, const int32_t **qlist_out, double *err_coef_out);
