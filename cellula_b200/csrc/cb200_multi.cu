// cb200_multi.cu -- the hot path over several B200s of one node from ONE host process: what an R session needs
// (`cb200_init(n_gpus)`, single-threaded caller, ncclCommInitAll).
//
// A cb200_multi owns one cb200_ctx per GPU (devices 0 .. n_gpus-1).  Every entry point fans out over one host
// thread per GPU; GPU r holds the contiguous block of cells [bounds[r], bounds[r+1]) (dgCMatrix columns) and runs
// the same rank-local stages as the single-GPU calls (*_local / *_finish, cb200_knn_device, cb200_snn_count_device
// + cb200_snn_emit).  Between the stages the threads perform the exchanges of SURVEY.md 8e with NCCL over
// NVLink, directly on the library's device buffers and on each context's own stream:
//     size factors  all-reduce  [sum of library sizes, #cells]          2 doubles
//     gene stats    all-reduce  fixed-point [sum y, sum y^2]            2 G int64  (exact: integer sum)
//     PCA           all-reduce  fixed-point H x H cross-product         2 H^2 int64 (exact)
//                   all-gather  scores                                  N x d doubles
//     kNN           all-gather  neighbour index                         N x k int32
//     SNN           none: rank r emits the edges of a contiguous range of cells j; the ranges are balanced by
//                   work (boundaries at N sqrt(r / R)) and the concatenation in rank order is the single-GPU list
// Because the two reductions are integer sums, the results are bit-identical to the single-GPU results.
// NCCL is loaded with dlopen at cb200_multi_init, so the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdarg>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "cb200_common.cuh"

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

bool load_nccl(NcclApi &api, std::string &err)
{
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle != nullptr)
            break;
    }
    if (api.handle == nullptr) {
        err = std::string("cb200_multi_init: cannot load NCCL (libnccl.so.2): ") + dlerror();
        return false;
    }
#define CB_SYM(field, name)                                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));                 \
    if (api.field == nullptr) {                                                                 \
        err = std::string("cb200_multi_init: NCCL symbol missing: ") + name;                   \
        return false;                                                                           \
    }
    CB_SYM(CommInitAll, "ncclCommInitAll")
    CB_SYM(CommDestroy, "ncclCommDestroy")
    CB_SYM(AllReduce, "ncclAllReduce")
    CB_SYM(Broadcast, "ncclBroadcast")
    CB_SYM(GroupStart, "ncclGroupStart")
    CB_SYM(GroupEnd, "ncclGroupEnd")
    CB_SYM(GetErrorString, "ncclGetErrorString")
#undef CB_SYM
    return true;
}
}  // namespace

struct cb200_multi {
    int R = 0;
    std::vector<cb200_ctx *> ctx;
    std::vector<ncclComm_t> comm;
    NcclApi nccl;
    std::string err;
    int64_t G = 0, N = 0, nnz = 0;
    std::vector<int64_t> bounds;      // cells of rank r: [bounds[r], bounds[r+1])
    std::vector<int64_t> jbounds;     // cells whose SNN edges rank r emits
    std::vector<int64_t> nnz_off;     // first non-zero of rank r's shard
    std::vector<double *> emb_full;   // per rank: row-major N x d (gathered scores or uploaded embedding)
    std::vector<int32_t *> knn_full;  // per rank: row-major N x k, 0-based
    std::vector<size_t> emb_cap, knn_cap;
    int emb_d = 0, knn_k = 0;
    bool have_emb = false, have_knn = false;
    std::vector<int64_t> edge_cnt;    // counted edges per rank (between snn_count and snn_emit)
    int H = 0;
};

static std::string g_multi_init_error;

static int multi_fail(cb200_multi *m, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (m != nullptr)
        m->err = buf;
    else
        g_multi_init_error = buf;
    return 1;
}

// runs fn(rank) on one thread per GPU; the first failing rank's message becomes the error of the call
static int run_all(cb200_multi *m, const std::function<int(int)> &fn)
{
    std::vector<int> rc((size_t)m->R, 0);
    if (m->R == 1) {
        rc[0] = fn(0);
    } else {
        std::vector<std::thread> th;
        for (int r = 0; r < m->R; ++r)
            th.emplace_back([&, r] {
                cudaSetDevice(m->ctx[(size_t)r]->device);
                rc[(size_t)r] = fn(r);
            });
        for (auto &t : th)
            t.join();
    }
    for (int r = 0; r < m->R; ++r)
        if (rc[(size_t)r] != 0) {
            if (rc[(size_t)r] != 2)   // 2: the message has been set by the rank itself (NCCL failure)
                m->err = std::string("[GPU ") + std::to_string(r) + "] " + cb200_last_error(m->ctx[(size_t)r]);
            return 1;
        }
    return 0;
}

#define CBM_NCCL(m, call)                                                                            \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess) {                                                                    \
            (m)->err = std::string(#call " failed: ") + (m)->nccl.GetErrorString(r__);              \
            return 2;                                                                                \
        }                                                                                            \
    } while (0)

static int allreduce(cb200_multi *m, int r, void *buf, size_t count, ncclDataType_t type)
{
    if (m->R == 1)
        return 0;
    CBM_NCCL(m, m->nccl.AllReduce(buf, buf, count, type, ncclSum, m->comm[(size_t)r], m->ctx[(size_t)r]->stream));
    return 0;
}

// concatenation of the ranks' row blocks in rank order: recv[off_q .. off_q + cnt_q) <- rank q's block
static int allgather_rows(cb200_multi *m, int r, const void *send, void *recv, const std::vector<int64_t> &bounds,
                          size_t row_elems, ncclDataType_t type, size_t elem_bytes)
{
    cb200_ctx *c = m->ctx[(size_t)r];
    if (m->R == 1) {
        if (send != recv)
            CB_CUDA(c, cudaMemcpyAsync(recv, send, (size_t)(bounds[1] - bounds[0]) * row_elems * elem_bytes,
                                       cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    CBM_NCCL(m, m->nccl.GroupStart());
    for (int q = 0; q < m->R; ++q) {
        const size_t cnt = (size_t)(bounds[(size_t)q + 1] - bounds[(size_t)q]) * row_elems;
        char *dst = static_cast<char *>(recv) + (size_t)bounds[(size_t)q] * row_elems * elem_bytes;
        ncclResult_t rc = m->nccl.Broadcast(q == r ? send : dst, dst, cnt, type, q, m->comm[(size_t)r], c->stream);
        if (rc != ncclSuccess) {
            m->nccl.GroupEnd();
            m->err = std::string("ncclBroadcast failed: ") + m->nccl.GetErrorString(rc);
            return 2;
        }
    }
    CBM_NCCL(m, m->nccl.GroupEnd());
    return 0;
}

extern "C" const char *cb200_multi_last_error(const cb200_multi *m)
{
    return m != nullptr ? m->err.c_str() : g_multi_init_error.c_str();
}

extern "C" int cb200_multi_init(int n_gpus, cb200_multi **out)
{
    if (out == nullptr)
        return multi_fail(nullptr, "cb200_multi_init: out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return multi_fail(nullptr, "cb200_multi_init: no CUDA device available; this library has no CPU path");
    if (n_gpus <= 0)
        n_gpus = count;
    if (n_gpus > count)
        return multi_fail(nullptr, "cb200_multi_init: %d GPUs requested, %d present", n_gpus, count);
    cb200_multi *m = new cb200_multi();
    m->R = n_gpus;
    for (int r = 0; r < n_gpus; ++r) {
        cb200_ctx *c = nullptr;
        if (cb200_init(r, nullptr, &c) != 0) {
            std::string msg = cb200_last_error(nullptr);
            for (auto *p : m->ctx)
                cb200_free(p);
            delete m;
            return multi_fail(nullptr, "%s", msg.c_str());
        }
        m->ctx.push_back(c);
    }
    if (n_gpus > 1) {
        std::string err;
        if (!load_nccl(m->nccl, err)) {
            for (auto *p : m->ctx)
                cb200_free(p);
            delete m;
            return multi_fail(nullptr, "%s", err.c_str());
        }
        m->comm.resize((size_t)n_gpus);
        std::vector<int> devs((size_t)n_gpus);
        for (int r = 0; r < n_gpus; ++r)
            devs[(size_t)r] = r;
        ncclResult_t rc = m->nccl.CommInitAll(m->comm.data(), n_gpus, devs.data());
        if (rc != ncclSuccess) {
            std::string msg = std::string("cb200_multi_init: ncclCommInitAll: ") + m->nccl.GetErrorString(rc);
            for (auto *p : m->ctx)
                cb200_free(p);
            delete m;
            return multi_fail(nullptr, "%s", msg.c_str());
        }
    }
    m->emb_full.assign((size_t)n_gpus, nullptr);
    m->knn_full.assign((size_t)n_gpus, nullptr);
    m->emb_cap.assign((size_t)n_gpus, 0);
    m->knn_cap.assign((size_t)n_gpus, 0);
    m->edge_cnt.assign((size_t)n_gpus, 0);
    *out = m;
    return 0;
}

extern "C" int cb200_multi_free(cb200_multi *m)
{
    if (m == nullptr)
        return 0;
    for (int r = 0; r < m->R; ++r) {
        cudaSetDevice(m->ctx[(size_t)r]->device);
        cudaStreamSynchronize(m->ctx[(size_t)r]->stream);
        if (m->emb_full[(size_t)r] != nullptr)
            cudaFree(m->emb_full[(size_t)r]);
        if (m->knn_full[(size_t)r] != nullptr)
            cudaFree(m->knn_full[(size_t)r]);
    }
    for (auto c : m->comm)
        m->nccl.CommDestroy(c);
    for (auto *c : m->ctx)
        cb200_free(c);
    delete m;
    return 0;
}

extern "C" int cb200_multi_n_gpus(const cb200_multi *m) { return m != nullptr ? m->R : 0; }

extern "C" int cb200_multi_stage_timings(cb200_multi *m, int rank, cb200_timing *out)
{
    if (m == nullptr || rank < 0 || rank >= m->R)
        return 1;
    cudaSetDevice(m->ctx[(size_t)rank]->device);
    return cb200_stage_timings(m->ctx[(size_t)rank], out);
}

static void set_bounds(cb200_multi *m, int64_t n)
{
    m->N = n;
    m->bounds.assign((size_t)m->R + 1, 0);
    const int64_t base = n / m->R, rem = n % m->R;
    for (int r = 0; r < m->R; ++r)
        m->bounds[(size_t)r + 1] = m->bounds[(size_t)r] + base + (r < rem ? 1 : 0);
    // SNN: a cell j is joined to partners h < j only, so its work grows like j: equal work = boundaries at N sqrt(r/R)
    m->jbounds.assign((size_t)m->R + 1, 0);
    for (int r = 0; r < m->R; ++r)
        m->jbounds[(size_t)r] = (int64_t)llround((double)n * sqrt((double)r / (double)m->R));
    m->jbounds[(size_t)m->R] = n;
    for (int r = 1; r <= m->R; ++r) {
        int64_t lo = m->jbounds[(size_t)r - 1] + 1, hi = n - (m->R - r);
        if (m->jbounds[(size_t)r] < lo)
            m->jbounds[(size_t)r] = lo;
        if (m->jbounds[(size_t)r] > hi)
            m->jbounds[(size_t)r] = hi;
    }
}

template <typename P>
static int multi_load(cb200_multi *m, int64_t n_genes, int64_t n_cells, const P *colptr, const int32_t *rowidx,
                      const double *values, bool i32)
{
    if (m == nullptr)
        return 1;
    if (colptr == nullptr || rowidx == nullptr || values == nullptr)
        return multi_fail(m, "cb200_multi_load_csc: NULL input");
    if (n_cells < m->R)
        return multi_fail(m, "cb200_multi_load_csc: fewer cells (%lld) than GPUs (%d)", (long long)n_cells, m->R);
    set_bounds(m, n_cells);
    m->G = n_genes;
    m->nnz = (int64_t)colptr[n_cells];
    m->nnz_off.assign((size_t)m->R + 1, 0);
    for (int r = 0; r <= m->R; ++r)
        m->nnz_off[(size_t)r] = (int64_t)colptr[m->bounds[(size_t)r]];
    m->have_emb = m->have_knn = false;
    return run_all(m, [&](int r) -> int {
        const int64_t c0 = m->bounds[(size_t)r], c1 = m->bounds[(size_t)r + 1];
        std::vector<int64_t> cp((size_t)(c1 - c0 + 1));   // the shard's column pointers, rebased to 0
        const int64_t off = (int64_t)colptr[c0];
        for (int64_t c = c0; c <= c1; ++c)
            cp[(size_t)(c - c0)] = (int64_t)colptr[c] - off;
        (void)i32;
        return cb200_load_csc(m->ctx[(size_t)r], n_genes, c1 - c0, cp.data(), rowidx + off, values + off, c0, n_cells);
    });
}

extern "C" int cb200_multi_load_csc(cb200_multi *m, int64_t n_genes, int64_t n_cells, const int64_t *colptr,
                                    const int32_t *rowidx, const double *values)
{
    return multi_load<int64_t>(m, n_genes, n_cells, colptr, rowidx, values, false);
}
extern "C" int cb200_multi_load_csc_i32(cb200_multi *m, int64_t n_genes, int64_t n_cells, const int32_t *colptr,
                                        const int32_t *rowidx, const double *values)
{
    return multi_load<int32_t>(m, n_genes, n_cells, colptr, rowidx, values, true);
}

extern "C" int cb200_multi_size_factors(cb200_multi *m, const double *sf_in, double *sf_out)
{
    if (m == nullptr)
        return 1;
    return run_all(m, [&](int r) -> int {
        cb200_ctx *c = m->ctx[(size_t)r];
        const int64_t c0 = m->bounds[(size_t)r];
        CB_TRY(cb200_size_factors_local(c, sf_in != nullptr ? sf_in + c0 : nullptr));
        CB_TRY(allreduce(m, r, c->part.p, 2, ncclDouble));
        double part[2];
        CB_TRY(cb200_read_back(c, part, c->part.p, sizeof(part)));
        return cb200_size_factors_finish(c, part[0], (int64_t)llround(part[1]), sf_out != nullptr ? sf_out + c0 : nullptr);
    });
}

extern "C" int cb200_multi_lognorm_genestats(cb200_multi *m, double *logx_out, double *gene_mean, double *gene_var)
{
    if (m == nullptr)
        return 1;
    return run_all(m, [&](int r) -> int {
        cb200_ctx *c = m->ctx[(size_t)r];
        CB_TRY(cb200_lognorm_genestats_local(c));
        CB_TRY(allreduce(m, r, c->gene_acc.p, (size_t)(2 * c->G), ncclInt64));   // fixed-point sums: exact
        CB_TRY(cb200_genestats_finish(c, r == 0 ? gene_mean : nullptr, r == 0 ? gene_var : nullptr));
        if (logx_out != nullptr)
            CB_TRY(cb200_fetch_logcounts(c, logx_out + m->nnz_off[(size_t)r]));
        return 0;
    });
}

static int ensure_emb(cb200_multi *m, int r, int d)
{
    const size_t need = (size_t)m->N * d;
    if (m->emb_cap[(size_t)r] < need) {
        cb200_ctx *c = m->ctx[(size_t)r];
        if (m->emb_full[(size_t)r] != nullptr)
            cudaFree(m->emb_full[(size_t)r]);
        m->emb_full[(size_t)r] = nullptr;
        CB_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&m->emb_full[(size_t)r]), need * sizeof(double)));
        m->emb_cap[(size_t)r] = need;
    }
    return 0;
}

extern "C" int cb200_multi_pca(cb200_multi *m, const int32_t *hvg_idx, int n_hvg, int d, double *scores, double *rotation,
                               double *var_explained, double *total_var)
{
    if (m == nullptr)
        return 1;
    m->H = n_hvg;
    int rc = run_all(m, [&](int r) -> int {
        cb200_ctx *c = m->ctx[(size_t)r];
        const int64_t c0 = m->bounds[(size_t)r], nl = m->bounds[(size_t)r + 1] - c0;
        CB_TRY(cb200_pca_gram_local(c, hvg_idx, n_hvg));
        if (c->gram_fixed)
            CB_TRY(allreduce(m, r, c->gram.p, (size_t)2 * n_hvg * n_hvg, ncclInt64));   // exact integer cross-product
        else
            CB_TRY(allreduce(m, r, c->gram.p, (size_t)n_hvg * n_hvg, ncclDouble));
        std::vector<double> local;
        if (scores != nullptr)
            local.resize((size_t)nl * d);
        CB_TRY(cb200_pca_finish(c, d, scores != nullptr ? local.data() : nullptr, r == 0 ? rotation : nullptr,
                                r == 0 ? var_explained : nullptr, r == 0 ? total_var : nullptr));
        if (scores != nullptr)   // the shard's rows of every column of the N x d column-major result
            for (int t = 0; t < d; ++t)
                memcpy(scores + (size_t)t * (size_t)m->N + (size_t)c0, local.data() + (size_t)t * (size_t)nl,
                       (size_t)nl * sizeof(double));
        // every GPU gets the whole embedding (row-major N x d): it is small (cfg3: 0.5 GB) and kNN needs all of it
        CB_TRY(ensure_emb(m, r, d));
        CB_TRY(allgather_rows(m, r, c->scores.p, m->emb_full[(size_t)r], m->bounds, (size_t)d, ncclDouble, sizeof(double)));
        return 0;
    });
    if (rc == 0) {
        m->emb_d = d;
        m->have_emb = true;
        m->have_knn = false;
    }
    return rc;
}

__global__ void k_multi_colmajor_to_rowmajor(const double *__restrict__ in, int64_t n, int d, double *__restrict__ out)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < n * d) {
        const int64_t i = idx / d;
        const int t = (int)(idx % d);
        out[idx] = in[i + n * (int64_t)t];
    }
}

extern "C" int cb200_multi_knn(cb200_multi *m, const double *emb, int64_t n_total, int d, int k, int32_t *index_out)
{
    if (m == nullptr)
        return 1;
    if (emb == nullptr) {
        if (!m->have_emb)
            return multi_fail(m, "cb200_multi_knn: no embedding given and no PCA scores in the context");
        if (n_total != m->N || d > m->emb_d)
            return multi_fail(m, "cb200_multi_knn: the resident scores are %lld x %d", (long long)m->N, m->emb_d);
    } else {
        if (n_total < 2 * m->R || d < 1)
            return multi_fail(m, "cb200_multi_knn: bad embedding shape");
        if (n_total != m->N)
            set_bounds(m, n_total);
    }
    int rc = run_all(m, [&](int r) -> int {
        cb200_ctx *c = m->ctx[(size_t)r];
        const int64_t c0 = m->bounds[(size_t)r], c1 = m->bounds[(size_t)r + 1], nl = c1 - c0;
        int ld = m->emb_d;
        if (emb != nullptr) {   // every GPU uploads the whole (small) matrix and converts it to row-major
            CB_TRY(ensure_emb(m, r, d));
            CB_CUDA(c, c->stage_d.ensure((size_t)n_total * d));
            CB_TRY(cb200_h2d(c, c->stage_d.p, emb, (size_t)n_total * d * sizeof(double), c->stream));
            k_multi_colmajor_to_rowmajor<<<(unsigned)((n_total * d + 255) / 256), 256, 0, c->stream>>>(c->stage_d.p, n_total, d,
                                                                                                    m->emb_full[(size_t)r]);
            CB_LAUNCH(c);
            ld = d;
        }
        std::vector<int32_t> local;
        if (index_out != nullptr)
            local.resize((size_t)nl * k);
        // cb200_knn_device takes a row-major matrix with leading dimension = its d; the resident scores have
        // emb_d columns of which the first d are used, so the truncated case goes through a strided view
        if (ld != d)
            return cb200_fail(c, "cb200_multi_knn: ndims (%d) smaller than the resident components (%d): pass the matrix", d, ld);
        CB_TRY(cb200_knn_device(c, m->emb_full[(size_t)r], n_total, d, k, c0, c1, index_out != nullptr ? local.data() : nullptr));
        if (index_out != nullptr)
            for (int t = 0; t < k; ++t)
                memcpy(index_out + (size_t)t * (size_t)n_total + (size_t)c0, local.data() + (size_t)t * (size_t)nl,
                       (size_t)nl * sizeof(int32_t));
        const size_t need = (size_t)n_total * k;
        if (m->knn_cap[(size_t)r] < need) {
            if (m->knn_full[(size_t)r] != nullptr)
                cudaFree(m->knn_full[(size_t)r]);
            m->knn_full[(size_t)r] = nullptr;
            CB_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&m->knn_full[(size_t)r]), need * sizeof(int32_t)));
            m->knn_cap[(size_t)r] = need;
        }
        CB_TRY(allgather_rows(m, r, c->knn.p, m->knn_full[(size_t)r], m->bounds, (size_t)k, ncclInt32, sizeof(int32_t)));
        return 0;
    });
    if (rc == 0) {
        if (emb != nullptr) {
            m->emb_d = d;
            m->have_emb = true;
        }
        m->knn_k = k;
        m->have_knn = true;
    }
    return rc;
}

__global__ void k_multi_index_to_rowmajor0(const int32_t *__restrict__ in, int64_t n, int k, int32_t *__restrict__ out,
                                           int *__restrict__ flag)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx < n * k) {
        const int64_t j = idx / k;
        const int cc = (int)(idx % k);
        const int v = in[j + n * (int64_t)cc] - 1;
        if (v < 0 || v >= n)
            atomicOr(flag, 1);
        out[idx] = v;
    }
}

// phase 1: reverse lists + counting on every GPU for its range of cells j; total edge count
extern "C" int cb200_multi_snn_count(cb200_multi *m, const int32_t *index, int64_t n_total, int k, int type, int64_t *n_edges)
{
    if (m == nullptr || n_edges == nullptr)
        return 1;
    if (index == nullptr) {
        if (!m->have_knn || n_total != m->N || k != m->knn_k)
            return multi_fail(m, "cb200_multi_snn: no index given and the resident kNN result is not %lld x %d",
                              (long long)n_total, k);
    } else if (n_total != m->N) {
        if (n_total < 2 * m->R)
            return multi_fail(m, "cb200_multi_snn: bad index shape");
        set_bounds(m, n_total);
    }
    int rc = run_all(m, [&](int r) -> int {
        cb200_ctx *c = m->ctx[(size_t)r];
        if (index != nullptr) {
            const size_t need = (size_t)n_total * k;
            if (m->knn_cap[(size_t)r] < need) {
                if (m->knn_full[(size_t)r] != nullptr)
                    cudaFree(m->knn_full[(size_t)r]);
                m->knn_full[(size_t)r] = nullptr;
                CB_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&m->knn_full[(size_t)r]), need * sizeof(int32_t)));
                m->knn_cap[(size_t)r] = need;
            }
            CB_CUDA(c, c->stage_i.ensure(need));
            CB_TRY(cb200_h2d(c, c->stage_i.p, index, need * sizeof(int32_t), c->stream));
            CB_CUDA(c, cudaMemsetAsync(c->flag.p, 0, sizeof(int), c->stream));
            k_multi_index_to_rowmajor0<<<(unsigned)((need + 255) / 256), 256, 0, c->stream>>>(c->stage_i.p, n_total, k,
                                                                                             m->knn_full[(size_t)r], c->flag.p);
            CB_LAUNCH(c);
            int h_flag = 0;
            CB_TRY(cb200_read_back(c, &h_flag, c->flag.p, sizeof(int)));
            CB_CHECK(c, h_flag == 0, "cb200_multi_snn: neighbour index out of range [1, n]");
        }
        int64_t ne = 0;
        CB_TRY(cb200_snn_count_device(c, m->knn_full[(size_t)r], n_total, k, type, m->jbounds[(size_t)r],
                                      m->jbounds[(size_t)r + 1], &ne));
        m->edge_cnt[(size_t)r] = ne;
        return 0;
    });
    if (rc != 0)
        return rc;
    if (index != nullptr) {
        m->knn_k = k;
        m->have_knn = true;
    }
    int64_t tot = 0;
    for (int r = 0; r < m->R; ++r)
        tot += m->edge_cnt[(size_t)r];
    *n_edges = tot;
    return 0;
}

// phase 2: every GPU emits its edges at its offset of the caller's arrays (rank order = ascending j = the
// single-GPU emission order); NULL pointers keep the edges in HBM
extern "C" int cb200_multi_snn_emit(cb200_multi *m, int32_t *from, int32_t *to, double *weight)
{
    if (m == nullptr)
        return 1;
    std::vector<int64_t> off((size_t)m->R + 1, 0);
    for (int r = 0; r < m->R; ++r)
        off[(size_t)r + 1] = off[(size_t)r] + m->edge_cnt[(size_t)r];
    return run_all(m, [&](int r) -> int {
        cb200_ctx *c = m->ctx[(size_t)r];
        const int64_t o = off[(size_t)r];
        if (from == nullptr)
            return cb200_snn_emit(c, nullptr, nullptr, nullptr);
        return cb200_snn_emit(c, from + o, to + o, weight + o);
    });
}

extern "C" int cb200_multi_edges_checksum(cb200_multi *m, int64_t *n_edges, uint64_t *checksum)
{
    if (m == nullptr || n_edges == nullptr || checksum == nullptr)
        return 1;
    std::vector<int64_t> ne((size_t)m->R, 0);
    std::vector<uint64_t> cs((size_t)m->R, 0);
    int rc = run_all(m, [&](int r) -> int { return cb200_edges_checksum(m->ctx[(size_t)r], &ne[(size_t)r], &cs[(size_t)r]); });
    if (rc != 0)
        return rc;
    *n_edges = 0;
    *checksum = 0;
    for (int r = 0; r < m->R; ++r) {
        *n_edges += ne[(size_t)r];
        *checksum += cs[(size_t)r];
    }
    return 0;
}
