// cb200_snn.cu -- K7: shared-nearest-neighbour edge weighting.
//
// Replaces bluster::neighborsToSNNGraph(index, type) -> libscran build_snn_graph inside
// bluster::makeSNNGraph(), /root/reference/R/makeGraphsAndClusters.R:83-85 (:125-127, :284).
// Definition: L_j = [j, nn_1(j) .. nn_k(j)] with ranks 0..k;
// cells j and h < j are joined when L_j and L_h share an element;
//   rank    w = max(k - r/2, 1e-6),  r = min over shared s of rank_j(s) + rank_h(s)
//   number  w = max(|L_j & L_h|, 1e-6)
//   jaccard w = max(n / (2(k+1) - n), 1e-6),  n = |L_j & L_h|
//
// Kernels
//   k_rev_count / k_rev_fill / k_rev_sort   reverse lists hosts[s] = {(cell, rank)} of all
//                                           cells whose L contains s, sorted by cell
//   k_snn_merge<COUNT|EMIT>                 warp per cell j: the k+1 sorted lists
//                                           hosts[L_j[i]] are merged (lane i holds the head
//                                           of list i); every distinct h < j is one edge.
// Edges of cell j are emitted in libscran's order (first encounter while scanning
// i = 0..k, list i in ascending cell order), i.e. sorted by (first list that holds h, h).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cb200_common.cuh"

#define SNN_SLOTS 8  // at most 8 lists per lane: supports k + 1 <= 256 (ranks are 8 bits of a reverse-list entry)
#define SNN_INF 0xffffffffu

__global__ void k_index_to_rowmajor0(const int32_t *__restrict__ in, int64_t n, int k, int32_t *__restrict__ out,
                                     int *__restrict__ flag)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < n * k) {
        int64_t j = idx / k;
        int c = (int)(idx % k);
        int v = in[j + n * (int64_t)c] - 1;
        if (v < 0 || v >= n)
            atomicOr(flag, 1);
        out[idx] = v;
    }
}

// A neighbour index handed in by a caller (device pointer or host matrix) is not trusted: an entry outside [0, n) or
// equal to its own row would index out of bounds / break the one-entry-per-list invariant of the hash kernel.  Such
// entries are skipped by both passes (so everything stays in bounds) and reported through `flag` (1: range, 2: self).
__device__ __forceinline__ bool snn_entry_ok(int s, int64_t j, int i, int64_t n, int *__restrict__ flag)
{
    if (i == 0)
        return true;
    if (s < 0 || s >= n) {
        atomicOr(flag, 1);
        return false;
    }
    if (s == j) {
        atomicOr(flag, 2);
        return false;
    }
    return true;
}

__global__ void k_rev_count(const int32_t *__restrict__ nn, int64_t n, int k, int32_t *__restrict__ rcnt, int *__restrict__ flag)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int L = k + 1;
    if (idx < n * L) {
        int64_t j = idx / L;
        int i = (int)(idx % L);
        int s = i == 0 ? (int)j : nn[j * k + (i - 1)];
        if (snn_entry_ok(s, j, i, n, flag))
            atomicAdd(rcnt + s, 1);
    }
}

// duplicates inside a row of a caller-supplied index (thread per entry, compared with the entries before it)
__global__ void k_snn_check_dups(const int32_t *__restrict__ nn, int64_t n, int k, int *__restrict__ flag)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < n * k) {
        const int64_t j = idx / k;
        const int i = (int)(idx % k);
        const int v = nn[idx];
        if (v == (int)j)
            atomicOr(flag, 2);
        for (int t = 0; t < i; ++t)
            if (nn[j * k + t] == v) {
                atomicOr(flag, 4);
                break;
            }
    }
}

__global__ void k_rev_fill(const int32_t *__restrict__ nn, int64_t n, int k, const int64_t *__restrict__ rptr,
                           int32_t *__restrict__ rcur, uint32_t *__restrict__ hosts_u, int *__restrict__ flag)
{
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int L = k + 1;
    if (idx < n * L) {
        int64_t j = idx / L;
        int i = (int)(idx % L);
        int s = i == 0 ? (int)j : nn[j * k + (i - 1)];
        if (!snn_entry_ok(s, j, i, n, flag))
            return;
        int pos = atomicAdd(rcur + s, 1);
        hosts_u[rptr[s] + pos] = ((uint32_t)j << 8) | (uint32_t)i;
    }
}

// warp per list: rank sort (entries are unique, so rank = #smaller is a permutation)
// cut[c * L + i] = position of the entry (c, i) inside its sorted list = number of cells < c in that list
__global__ void __launch_bounds__(256) k_rev_sort(const int64_t *__restrict__ rptr, int64_t n,
                                                  const uint32_t *__restrict__ hosts_u, uint32_t *__restrict__ hosts,
                                                  int L, int32_t *__restrict__ cut)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t s = warp0; s < n; s += nwarps) {
        const int64_t b = rptr[s];
        const int len = (int)(rptr[s + 1] - b);
        if (len <= 32) {
            const uint32_t mine = lane < len ? hosts_u[b + lane] : SNN_INF;
            int rank = 0;
#pragma unroll
            for (int o = 0; o < 32; ++o) {
                const uint32_t other = __shfl_sync(0xffffffffu, mine, o);
                rank += other < mine ? 1 : 0;
            }
            if (lane < len) {
                hosts[b + rank] = mine;
                cut[(int64_t)(mine >> 8) * L + (mine & 0xffu)] = rank;
            }
        } else {
            for (int e = lane; e < len; e += 32) {
                const uint32_t mine = hosts_u[b + e];
                int rank = 0;
                for (int o = 0; o < len; ++o)
                    rank += hosts_u[b + o] < mine ? 1 : 0;
                hosts[b + rank] = mine;
                cut[(int64_t)(mine >> 8) * L + (mine & 0xffu)] = rank;
            }
        }
    }
}

// MODE 0: count edges of j (ecnt[j]) and per-list first-encounter counts (bcnt[j][i]).
// MODE 1: emit edges at eptr[j] + (prefix of bcnt[j][.])[imin] + running index.
// SLOTS lists per lane (k + 1 <= 32 * SLOTS).  Each lane stages the next SNN_CHUNK(SLOTS) entries
// of its lists in shared memory (loads issued back to back), so the merge loop -- one warp-min
// (redux.sync) and one shared-memory read per output edge -- has no dependent global load inside.
#define SNN_SMEM_WORDS 1056  // per warp: 32 lanes x (32 + 1 pad) words
template <int SLOTS>
struct SnnChunk {
    static constexpr int value = 32 / SLOTS;  // entries staged per list
};

template <int MODE, int SLOTS>
__global__ void __launch_bounds__(256)
k_snn_merge(const int32_t *__restrict__ nn, int64_t n, int k, int type, int64_t j_begin, int64_t j_end,
            const int64_t *__restrict__ rptr, const uint32_t *__restrict__ hosts, int32_t *__restrict__ ecnt,
            int32_t *__restrict__ bcnt, const int64_t *__restrict__ eptr, int32_t *__restrict__ efrom,
            int32_t *__restrict__ eto, double *__restrict__ ew, const int32_t *__restrict__ jlist,
            const int *__restrict__ nlist)
{
    constexpr int CH = SnnChunk<SLOTS>::value;
    __shared__ uint32_t stage[8][SNN_SMEM_WORDS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t *mine = stage[w] + lane * 33;  // this lane's SLOTS x CH staged entries (+1 pad: no bank conflicts)
    const int L = k + 1;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    // either every cell of [j_begin, j_end), or the cells j_begin + jlist[0 .. *nlist) (the ones k_snn_hash left)
    const int64_t n_items = jlist != nullptr ? (int64_t)*nlist : j_end - j_begin;
    for (int64_t it = warp0; it < n_items; it += nwarps) {
        const int64_t jl = jlist != nullptr ? (int64_t)jlist[it] : it;
        const int64_t j = j_begin + jl;
        int64_t gptr[SLOTS], gend[SLOTS];   // next entry not yet staged / end of the list
        int hpos[SLOTS], hcnt[SLOTS];       // position in / size of the staged chunk
        uint32_t head[SLOTS];
        int nfirst[SLOTS];  // MODE 0: edges first met in my list; MODE 1: next free slot of my bucket
        auto refill = [&](int u) {
            // stage the next CH entries of list u (independent loads, one latency)
            int c = 0;
#pragma unroll
            for (int e = 0; e < CH; ++e) {
                if (gptr[u] + e < gend[u]) {
                    mine[u * CH + e] = hosts[gptr[u] + e];
                    c = e + 1;
                }
            }
            gptr[u] += c;
            hcnt[u] = c;
            hpos[u] = 0;
        };
        auto load_head = [&](int u) {
            if (hpos[u] == hcnt[u] && gptr[u] < gend[u])
                refill(u);
            uint32_t hv = SNN_INF;
            if (hpos[u] < hcnt[u]) {
                hv = mine[u * CH + hpos[u]];
                if ((int64_t)(hv >> 8) >= j) {  // sorted by cell: nothing further can precede j
                    hv = SNN_INF;
                    gptr[u] = gend[u];
                    hcnt[u] = hpos[u];
                }
            }
            head[u] = hv;
        };
#pragma unroll
        for (int u = 0; u < SLOTS; ++u) {
            const int i = u * 32 + lane;
            gptr[u] = 0;
            gend[u] = 0;
            hpos[u] = hcnt[u] = 0;
            nfirst[u] = 0;
            if (i < L) {
                const int s = i == 0 ? (int)j : nn[j * k + (i - 1)];
                gptr[u] = rptr[s];
                gend[u] = rptr[s + 1];
            }
            load_head(u);
        }
        if (MODE == 1) {
            // exclusive prefix of the bucket counts over i = 0..k, bucket i = u*32 + lane
            int carry = 0;
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
                const int i = u * 32 + lane;
                const int c = i < L ? bcnt[jl * L + i] : 0;
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o)
                        inc += t;
                }
                nfirst[u] = carry + inc - c;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        const int64_t ebase = MODE == 1 ? eptr[jl] : 0;
        int total = 0;
        while (true) {
            uint32_t hmin = SNN_INF >> 8;
#pragma unroll
            for (int u = 0; u < SLOTS; ++u)
                hmin = min(hmin, head[u] >> 8);
            const uint32_t m = __reduce_min_sync(0xffffffffu, hmin);
            if (m == (SNN_INF >> 8))
                break;
            int rs = 0x7fffffff, nmatch = 0, imin = 0x7fffffff;
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
                if ((head[u] >> 8) == m && head[u] != SNN_INF) {
                    const int i = u * 32 + lane;
                    rs = min(rs, i + (int)(head[u] & 0xffu));
                    imin = min(imin, i);
                    nmatch += 1;
                    hpos[u] += 1;
                    load_head(u);
                }
            }
            rs = __reduce_min_sync(0xffffffffu, rs);
            imin = __reduce_min_sync(0xffffffffu, imin);
            nmatch = __reduce_add_sync(0xffffffffu, nmatch);
            // the lane/slot that owns bucket imin
            const int own_lane = imin & 31, own_u = imin >> 5;
            if (MODE == 0) {
#pragma unroll
                for (int u = 0; u < SLOTS; ++u)
                    if (u == own_u && lane == own_lane)
                        nfirst[u] += 1;
                total += 1;
            } else {
                int slot = 0;
#pragma unroll
                for (int u = 0; u < SLOTS; ++u) {
                    const int v = __shfl_sync(0xffffffffu, nfirst[u], own_lane);
                    if (u == own_u) {
                        slot = v;
                        if (lane == own_lane)
                            nfirst[u] += 1;
                    }
                }
                if (lane == 0) {
                    double wgt;
                    if (type == CB200_SNN_RANK)
                        wgt = (double)k - 0.5 * (double)rs;
                    else if (type == CB200_SNN_NUMBER)
                        wgt = (double)nmatch;
                    else
                        wgt = __ddiv_rn((double)nmatch, (double)(2 * L - nmatch));
                    if (wgt < 1e-6)
                        wgt = 1e-6;
                    const int64_t o = ebase + slot;
                    efrom[o] = (int32_t)(j + 1);
                    eto[o] = (int32_t)(m + 1);
                    ew[o] = wgt;
                }
            }
        }
        if (MODE == 0) {
            if (lane == 0)
                ecnt[jl] = total;
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
                const int i = u * 32 + lane;
                if (i < L)
                    bcnt[jl * L + i] = nfirst[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// K7b, main path: warp per cell j, hash table in shared memory, NO atomics.
//
// The entries (cell < j) of the lists hosts[L_j[0]], hosts[L_j[1]], ... are taken as ONE flat sequence -- list after
// list, ascending cell inside a list: libscran's scan order -- 32 consecutive entries per step, one per lane, read
// straight from global memory (the entry of the next step is in flight while this one is inserted).  Per step:
//   * __match_any_sync groups the lanes that hold the same cell h (it sits in several of the lists); the lowest
//     lane of a group acts for it with the group's count / smallest rank sum,
//   * the acting lanes probe the table (keys in shared memory).  The keys of one step are distinct, so a free slot
//     is claimed by WRITE-THEN-VERIFY (store, __syncwarp, re-read) instead of a CAS, and the per-partner value is
//     updated with plain shared-memory loads / stores,
//   * a key met for the first time gets the next output position (ballot + popcount keep the flat order), which IS
//     the emission order -- sorted by (first list that holds h, h) -- so no sort and no per-list counts are needed.
// Only ONE 16-bit value per partner is kept: the smallest rank sum (type rank) or the number of shared
// neighbours (number / jaccard).  The partners (h: 4 bytes, value: 2 bytes) go to the cell's slice of the partner
// store; k_snn_compact expands them into (from, to, weight) at their final offsets once the counts are scanned.
// Table sizes come in tiers (1 Ki ... 16 Ki slots); a cell goes to the smallest tier whose capacity covers the
// LENGTH of its flat sequence (an upper bound of its partners, so a table never overflows), and every tier is one
// launch over its own list of cells.  Cells beyond the largest tier are left to k_snn_merge.
// ---------------------------------------------------------------------------------------
#define SNN_HASH_EMPTY 0xffffffffu
#define SNN_TIERS 6       // 0..4: tables of 1 Ki .. 16 Ki slots chosen by the flat length; 5: 16 Ki-slot CTA table for longer sequences
#define SNN_LONG_MAXFLAT 65536
#define SNN_MAXL 256
template <int LOG_SLOTS>
struct SnnTier {
    static constexpr int log_slots = LOG_SLOTS, slots = 1 << LOG_SLOTS;
    static constexpr int cap = slots / 16 * 11;                        // load factor <= 0.6875
    // per warp: keys, positions, values, and the table of the k + 1 lists (start, prefix) padded to 32 lists
    __host__ __device__ static int bytes(int lpad) { return slots * 6 + cap * 2 + lpad * 8 + 16; }
    // small tables: two CTAs per SM (<= 110 KB each); larger ones: one CTA with as many warps as fit 225 KB
    static int warps(int lpad, int *ctas_per_sm)
    {
        const int b = bytes(lpad);
        const int w2 = (110 * 1024) / b > 16 ? 16 : (110 * 1024) / b;
        const int w1 = (225 * 1024) / b > 16 ? 16 : ((225 * 1024) / b < 1 ? 1 : (225 * 1024) / b);
        *ctas_per_sm = w2 >= 8 ? 2 : 1;
        return w2 >= 8 ? w2 : w1;
    }
};

template <int LOG_SLOTS>
__global__ void __launch_bounds__(512)
k_snn_flat(const int32_t *__restrict__ nn, int k, int type, int lpad, int64_t j_begin, const int64_t *__restrict__ rptr,
           const uint32_t *__restrict__ hosts, const int32_t *__restrict__ cut, const int32_t *__restrict__ jlist,
           const int *__restrict__ nlist, int32_t *__restrict__ ecnt, const int64_t *__restrict__ toff,
           uint32_t *__restrict__ tmp_h, uint16_t *__restrict__ tmp_v)
{
    using T = SnnTier<LOG_SLOTS>;
    extern __shared__ __align__(16) unsigned char snn_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    unsigned char *mine = snn_smem + (size_t)w * T::bytes(lpad);
    uint32_t *key = reinterpret_cast<uint32_t *>(mine);                          // [slots]
    uint32_t *lstart = key + T::slots;                                            // [lpad] first entry of list i in hosts
    uint32_t *lpre = lstart + lpad;                                               // [lpad] entries of the lists before i
    uint16_t *idx = reinterpret_cast<uint16_t *>(lpre + lpad);                    // [slots] slot -> output position
    uint16_t *val = idx + T::slots;                                               // [cap]   position -> value
    const int L = k + 1;
    const bool want_min = type == CB200_SNN_RANK;
    const int n_items = *nlist;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int it = blockIdx.x * n_warps + w; it < n_items; it += gridDim.x * n_warps) {
        const int64_t jl = jlist[it];
        const int64_t j = j_begin + jl;
        // the lists: start in hosts, number of entries with cell < j (the entry of j itself sits at position
        // cut[j][i] of the list of its i-th neighbour), exclusive prefix = start in the flat sequence
        int total = 0;
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            int len = 0;
            uint32_t lb = 0;
            if (i < L) {
                const int sidx = i == 0 ? (int)j : nn[j * k + (i - 1)];
                lb = (uint32_t)rptr[sidx];
                len = cut[j * L + i];
            }
            int inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t2 = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                    inc += t2;
            }
            if (i < L) {
                lstart[i] = lb;
                lpre[i] = (uint32_t)(total + inc - len);
            }
            total += __shfl_sync(0xffffffffu, inc, 31);
        }
        for (int e = lane * 4; e < T::slots; e += 128)
            *reinterpret_cast<uint4 *>(&key[e]) = make_uint4(SNN_HASH_EMPTY, SNN_HASH_EMPTY, SNN_HASH_EMPTY, SNN_HASH_EMPTY);
        __syncwarp();
        const int64_t tb = toff[jl];
        int base = 0;       // distinct partners so far = next output position
        int cur = 0;        // list that holds this lane's flat index (advances monotonically)
        auto fetch = [&](int f, uint32_t &ent, int &li) {
            ent = 0u;
            li = 0;
            if (f < total) {
                while (cur + 1 < L && lpre[cur + 1] <= (uint32_t)f)
                    ++cur;
                li = cur;
                ent = __ldg(hosts + lstart[cur] + ((uint32_t)f - lpre[cur]));
            }
        };
        uint32_t ent_n;
        int li_n;
        fetch(lane, ent_n, li_n);
        for (int f0 = 0; f0 < total; f0 += 32) {
            const uint32_t ent = ent_n;
            const int li = li_n;
            const bool valid = f0 + lane < total;
            fetch(f0 + 32 + lane, ent_n, li_n);              // next step's entry: in flight during the insertion
            const uint32_t h = valid ? (ent >> 8) : (0xff000000u | (uint32_t)lane);   // invalid lanes: unique dummies
            const int rs = li + (int)(ent & 0xffu);
            const unsigned grp = __match_any_sync(0xffffffffu, h);
            const bool leader = valid && (grp & lt_mask) == 0u;
            int v = want_min ? rs : __popc(grp);
            if (want_min && (grp & (grp - 1u)) != 0u)
                v = __reduce_min_sync(grp, rs);
            // probe; free slots are claimed by write-then-verify (the keys of the acting lanes are distinct)
            bool pending = leader, is_new = false;
            int slot = (int)((h * 2654435761u) >> (32 - T::log_slots));
            while (__any_sync(0xffffffffu, pending)) {
                const uint32_t kc = pending ? key[slot] : 0u;
                const bool tent = pending && kc == SNN_HASH_EMPTY;
                if (tent)
                    key[slot] = h;
                __syncwarp();
                if (pending) {
                    if (kc == h) {
                        pending = false;
                    } else if (tent && key[slot] == h) {
                        pending = false;
                        is_new = true;
                    } else {
                        slot = (slot + 1) & (T::slots - 1);
                    }
                }
            }
            const unsigned nm = __ballot_sync(0xffffffffu, is_new);
            if (is_new) {
                const int pos = base + __popc(nm & lt_mask);
                idx[slot] = (uint16_t)pos;
                val[pos] = (uint16_t)v;
                tmp_h[tb + pos] = h;
            } else if (leader) {
                const int pos = idx[slot];
                const int c = val[pos];
                val[pos] = (uint16_t)(want_min ? min(c, v) : c + v);
            }
            base += __popc(nm);
            __syncwarp();
        }
        if (lane == 0)
            ecnt[jl] = base;
        for (int p = lane; p < base; p += 32)
            tmp_v[tb + p] = val[p];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// K7b for the two largest table tiers: ONE table per CTA, all warps of the CTA work on the same cell.
// The warp-per-cell kernel above needs a private table per warp; at 8 Ki / 16 Ki slots that leaves 3 / 1 warps per
// SM (k = 100: every cell).  Here the 32 warps of a CTA take the 32-entry steps of the cell's flat sequence in
// turn.  Pass 1 inserts with atomicCAS on the key and records per partner the SMALLEST flat index it occurs at
// (atomicMin) and its value (atomicMin of the rank sum / atomicAdd of the count).  Pass 2 walks the flat sequence
// again: an entry is a partner's first occurrence iff its flat index equals that minimum; the first occurrences are
// compacted in flat order (per-step ballots + a prefix over the steps), which is libscran's emission order -- the
// same output, bit for bit, as the warp kernel's running counter.
// ---------------------------------------------------------------------------------------
#define SNN_CTA_THREADS 1024
template <int LOG_SLOTS>
struct SnnCtaTier {
    static constexpr int slots = 1 << LOG_SLOTS;
    static constexpr int cap = slots / 16 * 11;
    static constexpr int nsteps_max = (cap + 31) / 32;
    static constexpr int steps_per_warp = (nsteps_max + SNN_CTA_THREADS / 32 - 1) / (SNN_CTA_THREADS / 32);
    __host__ __device__ static int bytes(int lpad) { return slots * 12 + lpad * 8 + (nsteps_max + 8) * 4; }
};

template <int LOG_SLOTS>
__global__ void __launch_bounds__(SNN_CTA_THREADS)
k_snn_flat_cta(const int32_t *__restrict__ nn, int k, int type, int lpad, int64_t j_begin, const int64_t *__restrict__ rptr,
               const uint32_t *__restrict__ hosts, const int32_t *__restrict__ cut, const int32_t *__restrict__ jlist,
               const int *__restrict__ nlist, int32_t *__restrict__ ecnt, const int64_t *__restrict__ toff,
               uint32_t *__restrict__ tmp_h, uint16_t *__restrict__ tmp_v)
{
    using T = SnnCtaTier<LOG_SLOTS>;
    extern __shared__ __align__(16) unsigned char snn_smem[];
    uint32_t *key = reinterpret_cast<uint32_t *>(snn_smem);      // [slots]
    uint32_t *first = key + T::slots;                             // [slots] smallest flat index of the partner
    uint32_t *val = first + T::slots;                             // [slots] smallest rank sum / number of shared neighbours
    uint32_t *lstart = val + T::slots;                            // [lpad]
    uint32_t *lpre = lstart + lpad;                               // [lpad]
    int *cnt = reinterpret_cast<int *>(lpre + lpad);              // [nsteps_max + 1] first occurrences per step, then offsets
    int *misc = cnt + T::nsteps_max + 1;                          // [0] total
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = SNN_CTA_THREADS / 32;
    const int L = k + 1;
    const bool want_min = type == CB200_SNN_RANK;
    const int n_items = *nlist;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int64_t jl = jlist[it];
        const int64_t j = j_begin + jl;
        __syncthreads();   // the previous cell is completely written out
        if (w == 0) {
            int total = 0;
            for (int i0 = 0; i0 < L; i0 += 32) {
                const int i = i0 + lane;
                int len = 0;
                uint32_t lb = 0;
                if (i < L) {
                    const int sidx = i == 0 ? (int)j : nn[j * k + (i - 1)];
                    lb = (uint32_t)rptr[sidx];
                    len = cut[j * L + i];
                }
                int inc = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t2 = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o)
                        inc += t2;
                }
                if (i < L) {
                    lstart[i] = lb;
                    lpre[i] = (uint32_t)(total + inc - len);
                }
                total += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0)
                misc[0] = total;
        }
        {
            const uint32_t vinit = want_min ? 0xffffffffu : 0u;
            uint4 *k4 = reinterpret_cast<uint4 *>(key), *f4 = reinterpret_cast<uint4 *>(first), *v4 = reinterpret_cast<uint4 *>(val);
            for (int e = tid; e < T::slots / 4; e += SNN_CTA_THREADS) {
                k4[e] = make_uint4(SNN_HASH_EMPTY, SNN_HASH_EMPTY, SNN_HASH_EMPTY, SNN_HASH_EMPTY);
                f4[e] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                v4[e] = make_uint4(vinit, vinit, vinit, vinit);
            }
        }
        __syncthreads();
        const int total = misc[0];
        const int nsteps = (total + 31) >> 5;
        // pass 1a: every thread requests all its entries of the flat sequence first (independent loads in flight) ...
        uint32_t hreg[T::steps_per_warp];
        int sreg[T::steps_per_warp];
#pragma unroll
        for (int t = 0; t < T::steps_per_warp; ++t) {
            const int st = w + t * NW;
            const int f = st * 32 + lane;
            hreg[t] = 0u;
            sreg[t] = -1;
            if (st < nsteps && f < total) {
                int lo = 0, hi = L - 1;   // largest i with lpre[i] <= f (lists may be empty: equal prefixes)
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (lpre[mid] <= (uint32_t)f)
                        lo = mid;
                    else
                        hi = mid - 1;
                }
                hreg[t] = __ldg(hosts + lstart[lo] + ((uint32_t)f - lpre[lo]));   // raw entry: cell << 8 | rank
                sreg[t] = lo;                                                      // its list
            }
        }
        // ... pass 1b: then inserts them
#pragma unroll
        for (int t = 0; t < T::steps_per_warp; ++t) {
            if (sreg[t] >= 0) {
                const int f = (w + t * NW) * 32 + lane;
                const uint32_t h = hreg[t] >> 8;
                const int rs = sreg[t] + (int)(hreg[t] & 0xffu);
                int slot = (int)((h * 2654435761u) >> (32 - LOG_SLOTS));
                while (true) {
                    const uint32_t old = atomicCAS(key + slot, SNN_HASH_EMPTY, h);
                    if (old == SNN_HASH_EMPTY || old == h)
                        break;
                    slot = (slot + 1) & (T::slots - 1);
                }
                atomicMin(first + slot, (uint32_t)f);
                if (want_min)
                    atomicMin(val + slot, (uint32_t)rs);
                else
                    atomicAdd(val + slot, 1u);
                hreg[t] = h;
                sreg[t] = slot;
            }
        }
        __syncthreads();
        // pass 2a: first occurrences per step
        unsigned firstbits = 0u;
#pragma unroll
        for (int t = 0; t < T::steps_per_warp; ++t) {
            const int st = w + t * NW;
            const int f = st * 32 + lane;
            const bool isf = sreg[t] >= 0 && first[sreg[t]] == (uint32_t)f;
            const unsigned m = __ballot_sync(0xffffffffu, isf);
            if (isf)
                firstbits |= 1u << t;
            if (lane == 0 && st < nsteps)
                cnt[st] = __popc(m);
        }
        __syncthreads();
        // exclusive prefix over the steps (warp 0)
        if (w == 0) {
            constexpr int PER = (T::nsteps_max + 31) / 32;
            int loc[PER];
            int sum = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int st = lane * PER + q;
                loc[q] = st < nsteps ? cnt[st] : 0;
                sum += loc[q];
            }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t2 = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                    inc += t2;
            }
            int run = inc - sum;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int st = lane * PER + q;
                if (st < nsteps)
                    cnt[st] = run;
                run += loc[q];
            }
            if (lane == 31)
                misc[1] = inc;   // distinct partners
        }
        __syncthreads();
        // pass 2b: write the partners at their emission positions
        const int64_t tb = toff[jl];
#pragma unroll
        for (int t = 0; t < T::steps_per_warp; ++t) {
            const int st = w + t * NW;
            const bool isf = (firstbits >> t) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, isf);
            if (isf) {
                const int pos = cnt[st] + __popc(m & lt_mask);
                tmp_h[tb + pos] = hreg[t];
                tmp_v[tb + pos] = (uint16_t)val[sreg[t]];
            }
        }
        if (tid == 0)
            ecnt[jl] = misc[1];
    }
}

// tier of every cell from the length of its flat sequence; tier SNN_TIERS = left to the merge kernel
__global__ void k_snn_tiers(const int32_t *__restrict__ tcnt, const int64_t *__restrict__ rptr, const int32_t *__restrict__ nn,
                            int64_t j_begin, int64_t nj, int k, int32_t *__restrict__ lists, int *__restrict__ nlists,
                            unsigned char *__restrict__ flagged, int32_t *__restrict__ ecnt, int long_max)
{
    const int64_t jl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (jl >= nj)
        return;
    const int t = tcnt[jl];
    int tier = 0;
    while (tier < 5 && t > ((1 << (10 + tier)) / 16) * 11)
        ++tier;
    // beyond the 16 Ki-slot capacity: the long CTA tier when the sequence is short enough for it (its DISTINCT partners
    // almost always fit the table; the kernel hands the cell to the merge kernel if they do not), else the merge kernel
    if (tier == 5 && t > long_max)
        tier = SNN_TIERS;
    flagged[jl] = tier == SNN_TIERS ? (unsigned char)2 : (unsigned char)0;
    if (tier == SNN_TIERS)
        ecnt[jl] = 0;
    lists[(int64_t)tier * nj + atomicAdd(nlists + tier, 1)] = (int32_t)jl;
}

// upper bound of the partners of cell j: all list entries with a cell < j
__global__ void k_snn_upper(const int32_t *__restrict__ cut, int64_t j_begin, int64_t nj, int L, int32_t *__restrict__ tcnt)
{
    const int64_t jl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (jl < nj) {
        const int32_t *c = cut + (j_begin + jl) * L;
        int t = 0;
        for (int i = 0; i < L; ++i)
            t += c[i];
        tcnt[jl] = t;
    }
}

// edges of the cells [j_begin, j_end) from the partner slices kept by k_snn_flat: warp per cell, 32 consecutive
// edges per store instruction; the 2-byte value (smallest rank sum / shared count) becomes the fp64 weight here
__global__ void __launch_bounds__(256)
k_snn_compact(int64_t j_begin, int64_t j_end, int k, int type, const int32_t *__restrict__ ecnt,
              const unsigned char *__restrict__ flagged, const int64_t *__restrict__ eptr,
              const int64_t *__restrict__ toff, const uint32_t *__restrict__ tmp_h, const uint16_t *__restrict__ tmp_v,
              int32_t *__restrict__ efrom, int32_t *__restrict__ eto, double *__restrict__ ew, int64_t loc0)
{
    const int lane = threadIdx.x & 31;
    const int L = k + 1;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t j = j_begin + warp0; j < j_end; j += nwarps) {
        const int64_t jl = j - loc0;
        if (flagged[jl] != 0)
            continue;
        const int cnt = ecnt[jl];
        const int64_t ebase = eptr[jl], tb = toff[jl];
        for (int p = lane; p < cnt; p += 32) {
            const int v = tmp_v[tb + p];
            double wgt;
            if (type == CB200_SNN_RANK)
                wgt = (double)k - 0.5 * (double)v;
            else if (type == CB200_SNN_NUMBER)
                wgt = (double)v;
            else
                wgt = __ddiv_rn((double)v, (double)(2 * L - v));
            if (wgt < 1e-6)
                wgt = 1e-6;
            efrom[ebase + p] = (int32_t)(j + 1);
            eto[ebase + p] = (int32_t)(tmp_h[tb + p] + 1);
            ew[ebase + p] = wgt;
        }
    }
}

template <int LOG_SLOTS>
static int launch_snn_flat(cb200_ctx *ctx, const int32_t *d_nn, int k, int type, int64_t j_begin, int64_t nj, int tier)
{
    using T = SnnTier<LOG_SLOTS>;
    const int lpad = (k + 1 + 31) & ~31;
    int ctas_per_sm = 1;
    const int warps = T::warps(lpad, &ctas_per_sm);
    const size_t smem = (size_t)T::bytes(lpad) * warps;
    CB_CUDA(ctx, cudaFuncSetAttribute(k_snn_flat<LOG_SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = ctx->num_sms * ctas_per_sm;
    k_snn_flat<LOG_SLOTS><<<grid, 32 * warps, smem, ctx->stream>>>(
        d_nn, k, type, lpad, j_begin, ctx->rptr.p, ctx->hosts.p, ctx->snn_cut.p, ctx->snn_jlist.p + (int64_t)tier * nj,
        ctx->snn_nlists.p + tier, ctx->ecnt.p, ctx->snn_toff.p, ctx->snn_tmp_h.p,
        reinterpret_cast<uint16_t *>(ctx->snn_tmp_rc.p));
    CB_LAUNCH(ctx);
    return 0;
}

// The same CTA-per-cell scheme for sequences LONGER than the largest table's capacity (k = 100: 28 % of the cells,
// half of the edges; the k-way merge kernel spends one warp-iteration per edge on them).  The capacity bounds the
// DISTINCT partners, which are a fraction of the flat length (duplicates: a partner sits in several lists), so the
// 16 Ki-slot table nearly always suffices; the kernel counts the distinct keys while inserting and hands the cell to
// the merge kernel's list if they exceed the capacity.  Nothing is cached in registers (up to 2 048 steps per cell):
// pass 2 re-reads the entries (L1/L2 hits) and re-probes; the first-occurrence masks of the steps are kept in shared
// memory so that only first occurrences are visited by the write-out pass.
__global__ void __launch_bounds__(SNN_CTA_THREADS)
k_snn_flat_cta_long(const int32_t *__restrict__ nn, int k, int type, int lpad, int64_t j_begin, const int64_t *__restrict__ rptr,
                    const uint32_t *__restrict__ hosts, const int32_t *__restrict__ cut, const int32_t *__restrict__ jlist,
                    const int *__restrict__ nlist, int32_t *__restrict__ ecnt, const int64_t *__restrict__ toff,
                    uint32_t *__restrict__ tmp_h, uint16_t *__restrict__ tmp_v, int32_t *__restrict__ merge_list,
                    int *__restrict__ merge_count, unsigned char *__restrict__ flagged, int cap)
{
    constexpr int LOG_SLOTS = 14, SLOTS = 1 << LOG_SLOTS;   // cap <= SLOTS / 16 * 11 distinct partners
    constexpr int MAXSTEPS = SNN_LONG_MAXFLAT / 32;
    extern __shared__ __align__(16) unsigned char snn_smem[];
    uint32_t *key = reinterpret_cast<uint32_t *>(snn_smem);
    uint32_t *first = key + SLOTS;
    uint32_t *val = first + SLOTS;
    uint32_t *lstart = val + SLOTS;
    uint32_t *lpre = lstart + lpad;
    int *cnt = reinterpret_cast<int *>(lpre + lpad);              // [MAXSTEPS + 1]
    uint32_t *fmask = reinterpret_cast<uint32_t *>(cnt + MAXSTEPS + 1);   // [MAXSTEPS] first-occurrence lanes of every step
    int *misc = reinterpret_cast<int *>(fmask + MAXSTEPS);        // [0] total, [1] distinct (final), [2] overflow, [3] distinct (running)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = SNN_CTA_THREADS / 32;
    const int L = k + 1;
    const bool want_min = type == CB200_SNN_RANK;
    const int n_items = *nlist;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int64_t jl = jlist[it];
        const int64_t j = j_begin + jl;
        __syncthreads();
        if (w == 0) {
            int total = 0;
            for (int i0 = 0; i0 < L; i0 += 32) {
                const int i = i0 + lane;
                int len = 0;
                uint32_t lb = 0;
                if (i < L) {
                    const int sidx = i == 0 ? (int)j : nn[j * k + (i - 1)];
                    lb = (uint32_t)rptr[sidx];
                    len = cut[j * L + i];
                }
                int inc = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t2 = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o)
                        inc += t2;
                }
                if (i < L) {
                    lstart[i] = lb;
                    lpre[i] = (uint32_t)(total + inc - len);
                }
                total += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) {
                misc[0] = total;
                misc[2] = 0;
                misc[3] = 0;
            }
        }
        {
            const uint32_t vinit = want_min ? 0xffffffffu : 0u;
            uint4 *k4 = reinterpret_cast<uint4 *>(key), *f4 = reinterpret_cast<uint4 *>(first), *v4 = reinterpret_cast<uint4 *>(val);
            for (int e = tid; e < SLOTS / 4; e += SNN_CTA_THREADS) {
                k4[e] = make_uint4(SNN_HASH_EMPTY, SNN_HASH_EMPTY, SNN_HASH_EMPTY, SNN_HASH_EMPTY);
                f4[e] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                v4[e] = make_uint4(vinit, vinit, vinit, vinit);
            }
        }
        __syncthreads();
        const int total = misc[0];
        const int nsteps = (total + 31) >> 5;
        // entry f of the flat sequence: (cell << 8 | rank, list)
        auto entry = [&](int f, int &li) -> uint32_t {
            int lo = 0, hi = L - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (lpre[mid] <= (uint32_t)f)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            li = lo;
            return __ldg(hosts + lstart[lo] + ((uint32_t)f - lpre[lo]));
        };
        // pass 1: insert, counting the distinct keys
        for (int st = w; st < nsteps; st += NW) {
            if (*reinterpret_cast<volatile int *>(misc + 2) != 0)
                break;
            const int f = st * 32 + lane;
            bool is_new = false;
            if (f < total) {
                int li;
                const uint32_t ent = entry(f, li);
                const uint32_t h = ent >> 8;
                const int rs = li + (int)(ent & 0xffu);
                int slot = (int)((h * 2654435761u) >> (32 - LOG_SLOTS));
                while (true) {
                    const uint32_t old = atomicCAS(key + slot, SNN_HASH_EMPTY, h);
                    if (old == SNN_HASH_EMPTY) {
                        is_new = true;
                        break;
                    }
                    if (old == h)
                        break;
                    slot = (slot + 1) & (SLOTS - 1);
                }
                atomicMin(first + slot, (uint32_t)f);
                if (want_min)
                    atomicMin(val + slot, (uint32_t)rs);
                else
                    atomicAdd(val + slot, 1u);
            }
            const unsigned nm = __ballot_sync(0xffffffffu, is_new);
            if (lane == 0 && nm != 0u) {
                const int c = __popc(nm);
                if (atomicAdd(misc + 3, c) + c > cap)
                    *reinterpret_cast<volatile int *>(misc + 2) = 1;   // at most 32 warps x 32 keys land after this: the table holds them
            }
        }
        __syncthreads();
        if (misc[2] != 0) {   // more distinct partners than the table may hold: the merge kernel takes the cell
            if (tid == 0) {
                merge_list[atomicAdd(merge_count, 1)] = (int32_t)jl;
                flagged[jl] = (unsigned char)2;
                ecnt[jl] = 0;
            }
            continue;
        }
        // pass 2a: first occurrences per step
        for (int st = w; st < nsteps; st += NW) {
            const int f = st * 32 + lane;
            bool isf = false;
            if (f < total) {
                int li;
                const uint32_t h = entry(f, li) >> 8;
                int slot = (int)((h * 2654435761u) >> (32 - LOG_SLOTS));
                while (key[slot] != h)
                    slot = (slot + 1) & (SLOTS - 1);
                isf = first[slot] == (uint32_t)f;
            }
            const unsigned m = __ballot_sync(0xffffffffu, isf);
            if (lane == 0) {
                cnt[st] = __popc(m);
                fmask[st] = m;
            }
        }
        __syncthreads();
        if (w == 0) {   // exclusive prefix over the steps
            constexpr int PER = MAXSTEPS / 32;
            int sum = 0;
            for (int q = 0; q < PER; ++q) {
                const int st = lane * PER + q;
                sum += st < nsteps ? cnt[st] : 0;
            }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t2 = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                    inc += t2;
            }
            int run = inc - sum;
            for (int q = 0; q < PER; ++q) {
                const int st = lane * PER + q;
                if (st < nsteps) {
                    const int c = cnt[st];
                    cnt[st] = run;
                    run += c;
                }
            }
            if (lane == 31)
                misc[1] = inc;
        }
        __syncthreads();
        // pass 2b: the first occurrences go to their emission positions
        const int64_t tb = toff[jl];
        for (int st = w; st < nsteps; st += NW) {
            const unsigned m = fmask[st];
            if ((m >> lane) & 1u) {
                const int f = st * 32 + lane;
                int li;
                const uint32_t h = entry(f, li) >> 8;
                int slot = (int)((h * 2654435761u) >> (32 - LOG_SLOTS));
                while (key[slot] != h)
                    slot = (slot + 1) & (SLOTS - 1);
                const int pos = cnt[st] + __popc(m & lt_mask);
                tmp_h[tb + pos] = h;
                tmp_v[tb + pos] = (uint16_t)val[slot];
            }
        }
        if (tid == 0)
            ecnt[jl] = misc[1];
    }
}

static int launch_snn_flat_cta_long(cb200_ctx *ctx, const int32_t *d_nn, int k, int type, int64_t j_begin, int64_t nj, int tier)
{
    const int lpad = (k + 1 + 31) & ~31;
    const size_t smem = (size_t)(1 << 14) * 12 + (size_t)lpad * 8 + (size_t)(SNN_LONG_MAXFLAT / 32 + 1) * 4 +
                        (size_t)(SNN_LONG_MAXFLAT / 32) * 4 + 32;
    CB_CUDA(ctx, cudaFuncSetAttribute(k_snn_flat_cta_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int cap = (1 << 14) / 16 * 11;
    if (const char *ec = getenv("CB200_SNN_LONG_CAP")) {   // tests: force the hand-over to the merge kernel
        const int v = atoi(ec);
        if (v >= 1 && v < cap)
            cap = v;
    }
    k_snn_flat_cta_long<<<ctx->num_sms, SNN_CTA_THREADS, smem, ctx->stream>>>(
        d_nn, k, type, lpad, j_begin, ctx->rptr.p, ctx->hosts.p, ctx->snn_cut.p, ctx->snn_jlist.p + (int64_t)tier * nj,
        ctx->snn_nlists.p + tier, ctx->ecnt.p, ctx->snn_toff.p, ctx->snn_tmp_h.p,
        reinterpret_cast<uint16_t *>(ctx->snn_tmp_rc.p), ctx->snn_jlist.p + (int64_t)SNN_TIERS * nj,
        ctx->snn_nlists.p + SNN_TIERS, ctx->snn_flag.p, cap);
    CB_LAUNCH(ctx);
    return 0;
}

template <int LOG_SLOTS>
static int launch_snn_flat_cta(cb200_ctx *ctx, const int32_t *d_nn, int k, int type, int64_t j_begin, int64_t nj, int tier)
{
    using T = SnnCtaTier<LOG_SLOTS>;
    const int lpad = (k + 1 + 31) & ~31;
    const size_t smem = (size_t)T::bytes(lpad);
    CB_CUDA(ctx, cudaFuncSetAttribute(k_snn_flat_cta<LOG_SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_snn_flat_cta<LOG_SLOTS>, SNN_CTA_THREADS, smem));
    const int grid = ctx->num_sms * (per_sm < 1 ? 1 : per_sm);
    k_snn_flat_cta<LOG_SLOTS><<<grid, SNN_CTA_THREADS, smem, ctx->stream>>>(
        d_nn, k, type, lpad, j_begin, ctx->rptr.p, ctx->hosts.p, ctx->snn_cut.p, ctx->snn_jlist.p + (int64_t)tier * nj,
        ctx->snn_nlists.p + tier, ctx->ecnt.p, ctx->snn_toff.p, ctx->snn_tmp_h.p,
        reinterpret_cast<uint16_t *>(ctx->snn_tmp_rc.p));
    CB_LAUNCH(ctx);
    return 0;
}

// [j_begin, j_end) is a sub-range of the range the per-cell arrays (ecnt, bcnt, eptr) were sized for;
// loc = j_begin - (first cell of that range)
template <int MODE>
static void launch_snn_merge(cb200_ctx *ctx, int grid, const int32_t *d_nn, int64_t n, int k, int type, int64_t j_begin,
                             int64_t j_end, int64_t loc, const int32_t *jlist = nullptr, const int *nlist = nullptr)
{
    const int32_t *nn = d_nn;
    int32_t *ecnt = ctx->ecnt.p + loc, *bcnt = ctx->bcnt.p + loc * (k + 1);
    int64_t *eptr = ctx->eptr.p + loc;
    if (k + 1 <= 32)
        k_snn_merge<MODE, 1><<<grid, 256, 0, ctx->stream>>>(nn, n, k, type, j_begin, j_end, ctx->rptr.p, ctx->hosts.p,
                                                           ecnt, bcnt, eptr, ctx->efrom.p, ctx->eto.p, ctx->ew8.p, jlist, nlist);
    else if (k + 1 <= 64)
        k_snn_merge<MODE, 2><<<grid, 256, 0, ctx->stream>>>(nn, n, k, type, j_begin, j_end, ctx->rptr.p, ctx->hosts.p,
                                                           ecnt, bcnt, eptr, ctx->efrom.p, ctx->eto.p, ctx->ew8.p, jlist, nlist);
    else if (k + 1 <= 128)
        k_snn_merge<MODE, 4><<<grid, 256, 0, ctx->stream>>>(nn, n, k, type, j_begin, j_end, ctx->rptr.p, ctx->hosts.p,
                                                           ecnt, bcnt, eptr, ctx->efrom.p, ctx->eto.p, ctx->ew8.p, jlist, nlist);
    else
        k_snn_merge<MODE, 8><<<grid, 256, 0, ctx->stream>>>(nn, n, k, type, j_begin, j_end, ctx->rptr.p, ctx->hosts.p,
                                                           ecnt, bcnt, eptr, ctx->efrom.p, ctx->eto.p, ctx->ew8.p, jlist, nlist);
}

// cell boundaries of SNN_CHUNKS pieces with about equal numbers of edges: out[2c] = first local cell of
// piece c, out[2c+1] = its first edge; out[2*SNN_CHUNKS..] = (nj, ne)
#define SNN_CHUNKS 8
__global__ void k_snn_chunk_bounds(const int64_t *__restrict__ eptr, int64_t nj, int64_t *__restrict__ out)
{
    const int c = threadIdx.x;
    if (c > SNN_CHUNKS)
        return;
    const int64_t ne = eptr[nj];
    int64_t j = nj;
    if (c < SNN_CHUNKS) {
        const int64_t target = (int64_t)(((__int128)ne * c) / SNN_CHUNKS);
        int64_t lo = 0, hi = nj;  // first j with eptr[j] >= target
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (eptr[mid] < target)
                lo = mid + 1;
            else
                hi = mid;
        }
        j = lo;
    }
    out[2 * c] = j;
    out[2 * c + 1] = eptr[j];
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// reverse lists + counting pass: how many edges the cells [j_begin, j_end) emit
static int snn_count(cb200_ctx *ctx, const int32_t *d_nn, int64_t n, int k, int type, int64_t j_begin, int64_t j_end,
                     int64_t *n_edges)
{
    CB_CHECK(ctx, n >= 2 && n < (1LL << 24), "cb200_snn: 2 <= n < 2^24 cells supported (got %lld)", (long long)n);
    CB_CHECK(ctx, k >= 1 && k + 1 <= 32 * SNN_SLOTS, "cb200_snn: 1 <= k <= %d supported (got %d)", 32 * SNN_SLOTS - 1, k);
    CB_CHECK(ctx, type >= 0 && type <= 2, "cb200_snn: type must be 0 (rank), 1 (number) or 2 (jaccard)");
    CB_CHECK(ctx, j_begin >= 0 && j_begin < j_end && j_end <= n, "cb200_snn: bad cell range");
    const int L = k + 1;
    const int64_t nj = j_end - j_begin;
    CB_CUDA(ctx, ctx->rcnt.ensure((size_t)n));
    CB_CUDA(ctx, ctx->rptr.ensure((size_t)n + 1));
    CB_CUDA(ctx, ctx->rcur.ensure((size_t)n));
    CB_CUDA(ctx, ctx->hosts_u.ensure((size_t)n * L));
    CB_CUDA(ctx, ctx->hosts.ensure((size_t)n * L));
    CB_CUDA(ctx, ctx->ecnt.ensure((size_t)nj));
    CB_CUDA(ctx, ctx->eptr.ensure((size_t)nj + 1));
    CB_CUDA(ctx, ctx->bcnt.ensure((size_t)nj * L));
    CB_CUDA(ctx, ctx->snn_flag.ensure((size_t)nj));
    CB_CUDA(ctx, ctx->snn_cut.ensure((size_t)n * L));
    CB_CUDA(ctx, ctx->snn_bounds.ensure(2 * (SNN_CHUNKS + 1)));

    cb_stage_begin(ctx, ST_SNN_REVLIST);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->rcnt.p, 0, (size_t)n * sizeof(int32_t), ctx->stream));
    CB_CUDA(ctx, cudaMemsetAsync(ctx->rcur.p, 0, (size_t)n * sizeof(int32_t), ctx->stream));
    const unsigned g1 = (unsigned)((n * L + 255) / 256);
    CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p + 3, 0, sizeof(int), ctx->stream));
    k_rev_count<<<g1, 256, 0, ctx->stream>>>(d_nn, n, k, ctx->rcnt.p, ctx->flag.p + 3);
    CB_LAUNCH(ctx);
    CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->rcnt.p, n, ctx->rptr.p));
    k_rev_fill<<<g1, 256, 0, ctx->stream>>>(d_nn, n, k, ctx->rptr.p, ctx->rcur.p, ctx->hosts_u.p, ctx->flag.p + 3);
    CB_LAUNCH(ctx);
    const int g2 = cb_grid_for(n, 8, ctx->num_sms * 16);
    k_rev_sort<<<g2, 256, 0, ctx->stream>>>(ctx->rptr.p, n, ctx->hosts_u.p, ctx->hosts.p, L, ctx->snn_cut.p);
    CB_LAUNCH(ctx);
    cb_stage_end(ctx, ST_SNN_REVLIST);
    {   // the edge kernels index the reverse lists with the entries of d_nn: stop here if any is unusable
        int h_bad = 0;
        CB_TRY(cb200_read_back(ctx, &h_bad, ctx->flag.p + 3, sizeof(int)));
        CB_CHECK(ctx, (h_bad & 1) == 0, "cb200_snn: neighbour index out of range [0, n)");
        CB_CHECK(ctx, (h_bad & 2) == 0, "cb200_snn: a row of the neighbour index names its own cell");
    }

    cb_stage_begin(ctx, ST_SNN_EDGES);
    // hash kernel for (nearly) every cell, merge kernel for the flagged ones (too many partners for the
    // table); CB200_SNN_ENGINE=merge runs everything through the merge kernel (cross-check)
    const char *seng = getenv("CB200_SNN_ENGINE");
    ctx->snn_use_hash = !(seng != nullptr && strcmp(seng, "merge") == 0);
    const int g3 = cb_grid_for(nj, 8, ctx->num_sms * 16);
    if (ctx->snn_use_hash) {
        // slices of the partner store: scan of the per-cell upper bounds (= lengths of the flat sequences)
        CB_CUDA(ctx, ctx->snn_tcnt.ensure((size_t)nj));
        CB_CUDA(ctx, ctx->snn_toff.ensure((size_t)nj + 1));
        k_snn_upper<<<(unsigned)((nj + 255) / 256), 256, 0, ctx->stream>>>(ctx->snn_cut.p, j_begin, nj, L, ctx->snn_tcnt.p);
        CB_LAUNCH(ctx);
        CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->snn_tcnt.p, nj, ctx->snn_toff.p));
        // one list of cells per table tier (+ one for the merge kernel)
        CB_CUDA(ctx, ctx->snn_jlist.ensure((size_t)nj * (SNN_TIERS + 1)));
        CB_CUDA(ctx, ctx->snn_nlists.ensure(SNN_TIERS + 1));
        CB_CUDA(ctx, cudaMemsetAsync(ctx->snn_nlists.p, 0, (SNN_TIERS + 1) * sizeof(int), ctx->stream));
        // the two largest tiers and the long tier: one table per CTA (the warp-per-cell kernel has room for 3 / 1 tables per
        // SM there); CB200_SNN_CTA=0 keeps the warp-per-cell kernel and leaves the long sequences to the merge kernel
        const char *scta = getenv("CB200_SNN_CTA");
        const bool use_cta = !(scta != nullptr && strcmp(scta, "0") == 0);
        k_snn_tiers<<<(unsigned)((nj + 255) / 256), 256, 0, ctx->stream>>>(ctx->snn_tcnt.p, ctx->rptr.p, d_nn, j_begin, nj, k,
                                                                         ctx->snn_jlist.p, ctx->snn_nlists.p, ctx->snn_flag.p,
                                                                         ctx->ecnt.p, use_cta ? SNN_LONG_MAXFLAT : 0);
        CB_LAUNCH(ctx);
        struct { int64_t ttot; int nl[SNN_TIERS + 1]; } hs;
        CB_TRY(cb200_read_back(ctx, &hs.ttot, ctx->snn_toff.p + nj, sizeof(int64_t)));
        CB_TRY(cb200_read_back(ctx, hs.nl, ctx->snn_nlists.p, sizeof(hs.nl)));
        if (getenv("CB200_TRACE") != nullptr)
            fprintf(stderr, "[cb200] snn: flat entries %lld, cells per tier %d %d %d %d %d, long tier %d, merge kernel %d\n",
                    (long long)hs.ttot, hs.nl[0], hs.nl[1], hs.nl[2], hs.nl[3], hs.nl[4], hs.nl[5], hs.nl[6]);
        CB_CUDA(ctx, ctx->snn_tmp_h.ensure((size_t)hs.ttot));
        CB_CUDA(ctx, ctx->snn_tmp_rc.ensure((size_t)(hs.ttot + 1) / 2));        // 2-byte values
        if (hs.nl[0] > 0) CB_TRY((launch_snn_flat<10>(ctx, d_nn, k, type, j_begin, nj, 0)));
        if (hs.nl[1] > 0) CB_TRY((launch_snn_flat<11>(ctx, d_nn, k, type, j_begin, nj, 1)));
        if (hs.nl[2] > 0) CB_TRY((launch_snn_flat<12>(ctx, d_nn, k, type, j_begin, nj, 2)));
        if (hs.nl[3] > 0) CB_TRY(use_cta ? (launch_snn_flat_cta<13>(ctx, d_nn, k, type, j_begin, nj, 3))
                                         : (launch_snn_flat<13>(ctx, d_nn, k, type, j_begin, nj, 3)));
        if (hs.nl[4] > 0) CB_TRY(use_cta ? (launch_snn_flat_cta<14>(ctx, d_nn, k, type, j_begin, nj, 4))
                                         : (launch_snn_flat<14>(ctx, d_nn, k, type, j_begin, nj, 4)));
        if (hs.nl[5] > 0) CB_TRY(launch_snn_flat_cta_long(ctx, d_nn, k, type, j_begin, nj, 5));   // only filled when use_cta
        ctx->snn_n_merge = hs.nl[SNN_TIERS];
        launch_snn_merge<0>(ctx, g3, d_nn, n, k, type, j_begin, j_end, 0, ctx->snn_jlist.p + (int64_t)SNN_TIERS * nj,
                            ctx->snn_nlists.p + SNN_TIERS);
    } else {
        launch_snn_merge<0>(ctx, g3, d_nn, n, k, type, j_begin, j_end, 0);
    }
    CB_LAUNCH(ctx);
    if (getenv("CB200_TRACE") != nullptr && ctx->snn_use_hash) {   // how many cells the hash table could not hold
        std::vector<unsigned char> hf((size_t)nj);
        std::vector<int32_t> hc((size_t)nj);
        CB_CUDA(ctx, cudaMemcpyAsync(hf.data(), ctx->snn_flag.p, (size_t)nj, cudaMemcpyDeviceToHost, ctx->stream));
        CB_CUDA(ctx, cudaMemcpyAsync(hc.data(), ctx->ecnt.p, (size_t)nj * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int64_t nf = 0, ef = 0, et = 0;
        for (int64_t q = 0; q < nj; ++q) {
            nf += hf[q];
            et += hc[q];
            if (hf[q])
                ef += hc[q];
        }
        fprintf(stderr, "[cb200] snn: %lld of %lld cells flagged for the merge kernel, holding %lld of %lld edges\n",
                (long long)nf, (long long)nj, (long long)ef, (long long)et);
    }
    CB_TRY(cb200_scan_i32_to_i64(ctx, ctx->ecnt.p, nj, ctx->eptr.p));
    k_snn_chunk_bounds<<<1, 32, 0, ctx->stream>>>(ctx->eptr.p, nj, ctx->snn_bounds.p);
    CB_LAUNCH(ctx);
    CB_CUDA(ctx, cudaMemcpyAsync(ctx->snn_bounds_h, ctx->snn_bounds.p, 2 * (SNN_CHUNKS + 1) * sizeof(int64_t),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t ne = ctx->snn_bounds_h[2 * SNN_CHUNKS + 1];
    ctx->snn_nn = d_nn;
    ctx->snn_n = n;
    ctx->snn_k = k;
    ctx->snn_type = type;
    ctx->snn_jbegin = j_begin;
    ctx->snn_jend = j_end;
    ctx->snn_counted = true;
    ctx->n_edges = ne;
    *n_edges = ne;
    return 0;
}

// emission pass of the counted range, in SNN_CHUNKS pieces of about equal size; with host outputs
// every piece is copied out (paced background copies) while the next one is being emitted
static int snn_emit(cb200_ctx *ctx, int32_t *from, int32_t *to, double *weight)
{
    CB_CHECK(ctx, ctx->snn_counted, "cb200_snn_emit: no counted SNN range in the context (call cb200_snn_count_device first)");
    const int64_t ne = ctx->n_edges;
    CB_CUDA(ctx, ctx->efrom.ensure((size_t)ne));
    CB_CUDA(ctx, ctx->eto.ensure((size_t)ne));
    CB_CUDA(ctx, ctx->ew8.ensure((size_t)ne));
    const bool out = from != nullptr && to != nullptr && weight != nullptr;
    const int64_t *bd = ctx->snn_bounds_h;
    if (ctx->snn_use_hash) {
        // the few cells left to the merge kernel first, in one launch over their list (their edges land at
        // their final offsets); the pieces below then only compact and copy
        const int64_t nj = ctx->snn_jend - ctx->snn_jbegin;
        const int gm = cb_grid_for(nj, 8, ctx->num_sms * 16);
        launch_snn_merge<1>(ctx, gm, ctx->snn_nn, ctx->snn_n, ctx->snn_k, ctx->snn_type, ctx->snn_jbegin, ctx->snn_jend, 0,
                            ctx->snn_jlist.p + (int64_t)SNN_TIERS * nj, ctx->snn_nlists.p + SNN_TIERS);
        CB_LAUNCH(ctx);
    }
    for (int c = 0; c < SNN_CHUNKS; ++c) {
        const int64_t ja = bd[2 * c], jb = bd[2 * c + 2], ea = bd[2 * c + 1], eb = bd[2 * c + 3];
        if (jb <= ja)
            continue;
        const int g3 = cb_grid_for(jb - ja, 8, ctx->num_sms * 16);
        if (ctx->snn_use_hash) {
            k_snn_compact<<<g3, 256, 0, ctx->stream>>>(ctx->snn_jbegin + ja, ctx->snn_jbegin + jb, ctx->snn_k, ctx->snn_type,
                                                       ctx->ecnt.p, ctx->snn_flag.p, ctx->eptr.p, ctx->snn_toff.p,
                                                       ctx->snn_tmp_h.p, reinterpret_cast<const uint16_t *>(ctx->snn_tmp_rc.p),
                                                       ctx->efrom.p, ctx->eto.p, ctx->ew8.p, ctx->snn_jbegin);
        } else {
            launch_snn_merge<1>(ctx, g3, ctx->snn_nn, ctx->snn_n, ctx->snn_k, ctx->snn_type, ctx->snn_jbegin + ja,
                                ctx->snn_jbegin + jb, ja);
        }
        CB_LAUNCH(ctx);
        if (out && eb > ea) {
            CB_TRY(cb200_copy_async(ctx, from + ea, ctx->efrom.p + ea, (size_t)(eb - ea) * sizeof(int32_t)));
            CB_TRY(cb200_copy_async(ctx, to + ea, ctx->eto.p + ea, (size_t)(eb - ea) * sizeof(int32_t)));
            CB_TRY(cb200_copy_async(ctx, weight + ea, ctx->ew8.p + ea, (size_t)(eb - ea) * sizeof(double)));
        }
    }
    cb_stage_end(ctx, ST_SNN_EDGES);
    ctx->snn_counted = false;
    if (out)
        CB_TRY(cb200_copies_wait(ctx));
    return 0;
}

static int snn_core(cb200_ctx *ctx, const int32_t *d_nn, int64_t n, int k, int type, int64_t j_begin, int64_t j_end,
                    int64_t *n_edges)
{
    CB_TRY(snn_count(ctx, d_nn, n, k, type, j_begin, j_end, n_edges));
    return snn_emit(ctx, nullptr, nullptr, nullptr);
}

static int snn_fetch(cb200_ctx *ctx, int64_t ne, int32_t **from, int32_t **to, double **weight)
{
    *from = nullptr;
    *to = nullptr;
    *weight = nullptr;
    const size_t cnt = ne > 0 ? (size_t)ne : 1;
    CB_CUDA(ctx, cudaMallocHost(reinterpret_cast<void **>(from), cnt * sizeof(int32_t)));
    CB_CUDA(ctx, cudaMallocHost(reinterpret_cast<void **>(to), cnt * sizeof(int32_t)));
    CB_CUDA(ctx, cudaMallocHost(reinterpret_cast<void **>(weight), cnt * sizeof(double)));
    cb_stage_begin(ctx, ST_D2H);
    if (ne > 0) {
        CB_TRY(cb200_d2h(ctx, *from, ctx->efrom.p, (size_t)ne * sizeof(int32_t), ctx->stream));
        CB_TRY(cb200_d2h(ctx, *to, ctx->eto.p, (size_t)ne * sizeof(int32_t), ctx->stream));
        CB_TRY(cb200_d2h(ctx, *weight, ctx->ew8.p, (size_t)ne * sizeof(double), ctx->stream));
    }
    cb_stage_end(ctx, ST_D2H);
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Copies the edges left in HBM by the last cb200_snn* call into caller-owned buffers (R vectors,
// pinned staging, ...): no allocation, no intermediate copy.
extern "C" int cb200_fetch_edges(cb200_ctx *ctx, int64_t n_edges, int32_t *from, int32_t *to, double *weight)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, n_edges == ctx->n_edges, "cb200_fetch_edges: the context holds %lld edges, not %lld",
             (long long)ctx->n_edges, (long long)n_edges);
    CB_CHECK(ctx, from != nullptr && to != nullptr && weight != nullptr, "cb200_fetch_edges: NULL output");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_stage_begin(ctx, ST_D2H);
    if (n_edges > 0) {
        CB_TRY(cb200_d2h(ctx, from, ctx->efrom.p, (size_t)n_edges * sizeof(int32_t), ctx->stream));
        CB_TRY(cb200_d2h(ctx, to, ctx->eto.p, (size_t)n_edges * sizeof(int32_t), ctx->stream));
        CB_TRY(cb200_d2h(ctx, weight, ctx->ew8.p, (size_t)n_edges * sizeof(double), ctx->stream));
    }
    cb_stage_end(ctx, ST_D2H);
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// order-independent digest of the edge list resident in HBM: sum over the edges of a 64-bit hash of
// (from, to, bit pattern of the weight), modulo 2^64.  Shards add up (sum over ranks), so an N-GPU run can be
// compared with the single-GPU run without moving 8 GB of edges.
__device__ __forceinline__ unsigned long long cs_mix(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) k_edge_checksum(const int32_t *__restrict__ ef, const int32_t *__restrict__ et,
                                                       const double *__restrict__ ew, int64_t ne,
                                                       unsigned long long *__restrict__ out)
{
    unsigned long long s = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += stride) {
        const unsigned long long key = ((unsigned long long)(uint32_t)ef[e] << 32) | (unsigned long long)(uint32_t)et[e];
        s += cs_mix(cs_mix(key) ^ (unsigned long long)__double_as_longlong(ew[e]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s != 0)
        atomicAdd(out, s);
}

extern "C" int cb200_edges_checksum(cb200_ctx *ctx, int64_t *n_edges, uint64_t *checksum)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, n_edges != nullptr && checksum != nullptr, "cb200_edges_checksum: NULL output argument");
    CB_CHECK(ctx, ctx->efrom.p != nullptr || ctx->n_edges == 0, "cb200_edges_checksum: no edge list in the context");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d = reinterpret_cast<unsigned long long *>(ctx->maxnrm.p);
    CB_CUDA(ctx, cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->n_edges > 0) {
        k_edge_checksum<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->efrom.p, ctx->eto.p, ctx->ew8.p, ctx->n_edges, d);
        CB_LAUNCH(ctx);
    }
    unsigned long long h = 0;
    CB_TRY(cb200_read_back(ctx, &h, d, sizeof(h)));
    *n_edges = ctx->n_edges;
    *checksum = (uint64_t)h;
    return 0;
}

extern "C" int cb200_free_edges(int32_t *from, int32_t *to, double *weight)
{
    if (from != nullptr)
        cudaFreeHost(from);
    if (to != nullptr)
        cudaFreeHost(to);
    if (weight != nullptr)
        cudaFreeHost(weight);
    return 0;
}

extern "C" int cb200_snn_device(cb200_ctx *ctx, const int32_t *d_index_rowmajor0, int64_t n_total, int k, int type,
                                int64_t j_begin, int64_t j_end, int64_t *n_edges, int32_t **from, int32_t **to,
                                double **weight)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, d_index_rowmajor0 != nullptr && n_edges != nullptr, "cb200_snn_device: NULL argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_TRY(snn_core(ctx, d_index_rowmajor0, n_total, k, type, j_begin, j_end, n_edges));
    if (from == nullptr || to == nullptr || weight == nullptr) {  // edges stay resident in HBM
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    return snn_fetch(ctx, *n_edges, from, to, weight);
}

extern "C" int cb200_snn(cb200_ctx *ctx, const int32_t *index, int64_t n_total, int k, int type, int64_t j_begin,
                         int64_t j_end, int64_t *n_edges, int32_t **from, int32_t **to, double **weight)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, n_edges != nullptr && from != nullptr && to != nullptr && weight != nullptr,
             "cb200_snn: NULL output argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int32_t *d_nn = nullptr;
    if (index == nullptr) {
        CB_CHECK(ctx, ctx->have_knn && ctx->knn_nq == n_total && ctx->knn_ntotal == n_total && ctx->knn_k == k,
                 "cb200_snn: no index given and the resident kNN result does not cover all %lld cells with k=%d",
                 (long long)n_total, k);
        d_nn = ctx->knn.p;
    } else {
        CB_CHECK(ctx, n_total >= 2 && k >= 1, "cb200_snn: bad index shape");
        cb_stage_begin(ctx, ST_H2D);
        CB_CUDA(ctx, ctx->stage_i.ensure((size_t)n_total * k));
        CB_CUDA(ctx, ctx->index0.ensure((size_t)n_total * k));
        CB_TRY(cb200_h2d(ctx, ctx->stage_i.p, index, (size_t)n_total * k * sizeof(int32_t), ctx->stream));
        CB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
        k_index_to_rowmajor0<<<(unsigned)((n_total * k + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->stage_i.p, n_total, k, ctx->index0.p, ctx->flag.p);
        CB_LAUNCH(ctx);
        k_snn_check_dups<<<(unsigned)((n_total * k + 255) / 256), 256, 0, ctx->stream>>>(ctx->index0.p, n_total, k, ctx->flag.p);
        CB_LAUNCH(ctx);
        cb_stage_end(ctx, ST_H2D);
        int h_flag = 0;
        CB_CUDA(ctx, cudaMemcpyAsync(&h_flag, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CB_CHECK(ctx, (h_flag & 1) == 0, "cb200_snn: neighbour index out of range [1, n]");
        CB_CHECK(ctx, (h_flag & 6) == 0, "cb200_snn: a row of the neighbour index names its own cell or the same cell twice");
        d_nn = ctx->index0.p;
    }
    CB_TRY(snn_core(ctx, d_nn, n_total, k, type, j_begin, j_end, n_edges));
    return snn_fetch(ctx, *n_edges, from, to, weight);
}

// Two-phase form for callers that own the result buffers (R vectors, pinned staging): count, let the
// caller size its arrays, then emit straight into them with the copies overlapping the emission.
extern "C" int cb200_snn_count_device(cb200_ctx *ctx, const int32_t *d_index_rowmajor0, int64_t n_total, int k,
                                      int type, int64_t j_begin, int64_t j_end, int64_t *n_edges)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, n_edges != nullptr, "cb200_snn_count_device: NULL argument");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int32_t *d_nn = d_index_rowmajor0;
    if (d_nn == nullptr) {
        CB_CHECK(ctx, ctx->have_knn && ctx->knn_nq == n_total && ctx->knn_ntotal == n_total && ctx->knn_k == k,
                 "cb200_snn_count_device: no index given and the resident kNN result does not cover all %lld cells "
                 "with k=%d", (long long)n_total, k);
        d_nn = ctx->knn.p;
    }
    return snn_count(ctx, d_nn, n_total, k, type, j_begin, j_end, n_edges);
}

extern "C" int cb200_snn_emit(cb200_ctx *ctx, int32_t *from, int32_t *to, double *weight)
{
    if (ctx == nullptr)
        return 1;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_TRY(snn_emit(ctx, from, to, weight));
    if (from == nullptr)
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cb200_snn_resident(cb200_ctx *ctx, int k, int type, int64_t *n_edges)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, n_edges != nullptr, "cb200_snn_resident: NULL output argument");
    CB_CHECK(ctx, ctx->have_knn && ctx->knn_nq == ctx->knn_ntotal && ctx->knn_k == k,
             "cb200_snn_resident: the resident kNN result does not cover all cells with k=%d", k);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_TRY(snn_core(ctx, ctx->knn.p, ctx->knn_ntotal, k, type, 0, ctx->knn_ntotal, n_edges));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
