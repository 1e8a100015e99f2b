// cb200_ctx.cu -- context lifetime, error reporting, stage timing, counts upload (a1) and
// the device-wide exclusive scan used by the compaction steps.
#include <cstdarg>
#include <cstring>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "cb200_common.cuh"

static std::string g_init_error;

int cb200_fail(cb200_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx != nullptr)
        ctx->err = buf;
    else
        g_init_error = buf;
    return 1;
}

extern "C" const char *cb200_version(void) { return "cellula_b200 0.1 sm_100a"; }

extern "C" const char *cb200_last_error(const cb200_ctx *ctx)
{
    return ctx != nullptr ? ctx->err.c_str() : g_init_error.c_str();
}

extern "C" int cb200_init(int device, void *stream, cb200_ctx **out)
{
    if (out == nullptr)
        return cb200_fail(nullptr, "cb200_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return cb200_fail(nullptr, "cb200_init: no CUDA device available (%s); this library has no CPU path",
                          cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return cb200_fail(nullptr, "cb200_init: device %d out of range (have %d)", device, count);
    e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return cb200_fail(nullptr, "cb200_init: cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess)
        return cb200_fail(nullptr, "cb200_init: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return cb200_fail(nullptr,
                          "cb200_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                          device, prop.major, prop.minor);
    cb200_ctx *ctx = new cb200_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (stream != nullptr) {
        ctx->stream = reinterpret_cast<cudaStream_t>(stream);
        ctx->own_stream = false;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return cb200_fail(nullptr, "cb200_init: cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        ctx->own_stream = true;
    }
    for (int s = 0; s < ST_COUNT_; ++s) {
        cudaEventCreate(&ctx->ev[s][0]);
        cudaEventCreate(&ctx->ev[s][1]);
        ctx->ev_used[s] = false;
    }
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    if (ctx->part.ensure(2) != cudaSuccess || ctx->flag.ensure(4) != cudaSuccess ||
        ctx->maxnrm.ensure(4) != cudaSuccess) {
        delete ctx;
        return cb200_fail(nullptr, "cb200_init: device allocation failed");
    }
    cudaMemsetAsync(ctx->flag.p, 0, 4 * sizeof(int), ctx->stream);
    if (cudaMallocHost(reinterpret_cast<void **>(&ctx->snn_bounds_h), 64 * sizeof(int64_t)) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void **>(&ctx->pin_scratch), CB200_PIN_SCRATCH) != cudaSuccess) {
        delete ctx;
        return cb200_fail(nullptr, "cb200_init: pinned allocation failed");
    }
    *out = ctx;
    return 0;
}

extern "C" int cb200_free(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 0;
    cudaSetDevice(ctx->device);
    // background work first: the copy worker and the row-index uploader still read / write the buffers freed below
    cb200_copies_shutdown(ctx);
    cb200_stage_shutdown(ctx);
    if (ctx->up_stream != nullptr)
        cudaStreamSynchronize(ctx->up_stream);
    cudaStreamSynchronize(ctx->stream);
    ctx->colptr.release(); ctx->rowidx.release(); ctx->x.release();
    ctx->libsize.release(); ctx->sf.release(); ctx->part.release(); ctx->flag.release();
    ctx->logx.release(); ctx->gene_acc.release(); ctx->gs_split.release(); ctx->gs_queue.release(); ctx->gs_tab_y.release(); ctx->gs_tab_q.release(); ctx->gs_ovf_entry.release(); ctx->gs_ovf_cell.release(); ctx->gs_ovf_count.release(); ctx->blk_id.release(); ctx->blk_cells.release(); ctx->blk_ptr.release(); ctx->blk_part.release(); ctx->blk_w.release(); ctx->gene_mean.release(); ctx->gene_var.release();
    ctx->hvg_idx.release(); ctx->hvg_perm.release(); ctx->hbptr.release(); ctx->gram_img.release(); ctx->slot.release(); ctx->hcnt.release(); ctx->hptr.release();
    ctx->hslot.release(); ctx->hval.release(); ctx->gram.release(); ctx->cov.release();
    ctx->mu.release(); ctx->eq.release(); ctx->ew.release(); ctx->ey.release(); ctx->ex.release();
    ctx->es.release(); ctx->et.release(); ctx->etheta.release(); ctx->eres.release();
    ctx->emuv.release(); ctx->egs.release(); ctx->rot.release(); ctx->scores.release(); ctx->varexp.release();
    ctx->emb.release(); ctx->stage_d.release(); ctx->emb_f.release(); ctx->knn_img.release(); ctx->knn_eperm.release(); ctx->knn_perm.release(); ctx->knn_bkt.release(); ctx->knn_hist.release();
    ctx->knn_qlist.release(); ctx->knn_qpos.release(); ctx->knn_cursor.release(); ctx->knn_qoff.release();
    ctx->knn_tile_lo.release(); ctx->knn_tile_hi.release(); ctx->hnorm.release();
    ctx->knn_cen.release(); ctx->knn_km_sums.release(); ctx->knn_km_cnt.release(); ctx->knn_dqc.release();
    ctx->knn_cl.release(); ctx->knn_rho.release(); ctx->knn_cl_rhomax.release(); ctx->knn_cl_tile_begin.release();
    ctx->knn_pad_before.release(); ctx->knn_cl_order.release();
    ctx->nrm2.release(); ctx->maxnrm.release(); ctx->cand_idx.release(); ctx->cand_tau.release(); ctx->cand_err.release();
    ctx->knn.release(); ctx->knn_dk.release(); ctx->fb_list.release(); ctx->fb_cnt.release(); ctx->fb_d.release(); ctx->fb_i.release(); ctx->stage_i.release();
    ctx->index0.release(); ctx->rcnt.release(); ctx->rptr.release(); ctx->rcur.release();
    ctx->hosts_u.release(); ctx->hosts.release(); ctx->ecnt.release(); ctx->eptr.release();
    ctx->bcnt.release(); ctx->efrom.release(); ctx->eto.release(); ctx->ew8.release();
    ctx->qc_mask.release(); ctx->qc_sum.release(); ctx->qc_det.release();
    ctx->diag_lab.release(); ctx->diag_i.release(); ctx->diag_ef.release(); ctx->diag_d.release(); ctx->diag_ew.release(); ctx->diag_cnt.release();
    ctx->scan_blk.release(); ctx->snn_bounds.release(); ctx->snn_flag.release(); ctx->snn_jlist.release(); ctx->snn_nlists.release(); ctx->snn_tcnt.release(); ctx->snn_toff.release(); ctx->snn_tmp_h.release(); ctx->snn_tmp_rc.release(); ctx->snn_cut.release();
    if (ctx->snn_bounds_h != nullptr)
        cudaFreeHost(ctx->snn_bounds_h);
    if (ctx->pin_scratch != nullptr)
        cudaFreeHost(ctx->pin_scratch);
    for (int s = 0; s < ST_COUNT_; ++s) {
        cudaEventDestroy(ctx->ev[s][0]);
        cudaEventDestroy(ctx->ev[s][1]);
    }
    if (ctx->up_stream != nullptr) {
        cudaStreamDestroy(ctx->up_stream);
        cudaEventDestroy(ctx->ev_x);
        cudaEventDestroy(ctx->ev_rowidx);
    }
    if (ctx->own_stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

// bytes: word copies when everything is 8- or 4-byte aligned, else byte by byte
__global__ void k_copy_bytes(void *__restrict__ dst, const void *__restrict__ src, size_t bytes)
{
    const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    const size_t align = (size_t)dst | (size_t)src | bytes;
    if ((align & 7) == 0) {
        for (size_t i = i0; i < bytes / 8; i += step)
            static_cast<unsigned long long *>(dst)[i] = static_cast<const unsigned long long *>(src)[i];
    } else if ((align & 3) == 0) {
        for (size_t i = i0; i < bytes / 4; i += step)
            static_cast<unsigned int *>(dst)[i] = static_cast<const unsigned int *>(src)[i];
    } else {
        for (size_t i = i0; i < bytes; i += step)
            static_cast<unsigned char *>(dst)[i] = static_cast<const unsigned char *>(src)[i];
    }
}

int cb200_read_back(cb200_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0)
        return 0;
    if (bytes <= CB200_PIN_SCRATCH / 2) {
        const unsigned grid = (unsigned)((bytes / 8 + 255) / 256 > 64 ? 64 : (bytes / 8 + 255) / 256 + 1);
        k_copy_bytes<<<grid, 256, 0, ctx->stream>>>(ctx->pin_scratch, src, bytes);
        CB_LAUNCH(ctx);
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(dst, ctx->pin_scratch, bytes);
        return 0;
    }
    CB_TRY(cb200_d2h(ctx, dst, src, bytes, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int cb200_write_small(cb200_ctx *ctx, void *dev_dst, const void *src, size_t bytes)
{
    if (bytes == 0)
        return 0;
    if (bytes <= CB200_PIN_SCRATCH / 2) {
        unsigned char *stage = ctx->pin_scratch + CB200_PIN_SCRATCH / 2;
        memcpy(stage, src, bytes);
        const unsigned grid = (unsigned)((bytes / 8 + 255) / 256 > 64 ? 64 : (bytes / 8 + 255) / 256 + 1);
        k_copy_bytes<<<grid, 256, 0, ctx->stream>>>(dev_dst, stage, bytes);
        CB_LAUNCH(ctx);
        CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    CB_TRY(cb200_h2d(ctx, dev_dst, src, bytes, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------
// paced background device->host copies
// ---------------------------------------------------------------------------------------
struct CopyJob {
    void *dst;
    const void *src;
    size_t bytes;
    cudaEvent_t after;
};
struct CopyWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<CopyJob> q;
    bool stop = false;
    int pending = 0;
    cudaError_t err = cudaSuccess;
    int device = 0;
    cb200_ctx *ctx = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t chunk_ev[2] = {nullptr, nullptr};
};
static constexpr size_t COPY_CHUNK = 32u << 20;  // 0.6 ms of PCIe; ONE chunk in flight: the copy engine serves another
                                                 // stream's transfer only when this stream's queue has run dry

// Waits for an event without sitting in a blocking driver call (which would hold up the API calls
// of the thread that runs the pipeline).
static cudaError_t poll_event(cudaEvent_t ev)
{
    for (;;) {
        const cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady)
            return e;
        std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
}

static void copy_worker_main(CopyWorker *w)
{
    cudaSetDevice(w->device);
    for (;;) {
        CopyJob job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->stop || !w->q.empty(); });
            if (w->q.empty())
                return;
            job = w->q.front();
            w->q.pop_front();
        }
        cudaError_t e = cudaStreamWaitEvent(w->stream, job.after, 0);
        size_t i = 0;
        if (e == cudaSuccess && !cb200_host_is_pinned(job.dst)) {
            // pageable destination (an R vector): device -> pinned ring -> caller memory with the memcpy threads;
            // chunks are 16 MB, so the short reads of the stages are not starved either
            if (cb200_d2h_worker(w->ctx, job.dst, job.src, job.bytes, w->stream) != 0)
                e = cudaErrorUnknown;
            job.bytes = 0;
        }
        for (size_t off = 0; off < job.bytes && e == cudaSuccess; off += COPY_CHUNK, ++i) {
            const size_t nb = job.bytes - off < COPY_CHUNK ? job.bytes - off : COPY_CHUNK;
            if (i >= 1)
                e = poll_event(w->chunk_ev[(i - 1) & 1]);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(static_cast<char *>(job.dst) + off, static_cast<const char *>(job.src) + off, nb,
                                    cudaMemcpyDeviceToHost, w->stream);
            if (e == cudaSuccess)
                e = cudaEventRecord(w->chunk_ev[i & 1], w->stream);
        }
        if (i >= 1 && e == cudaSuccess)
            e = poll_event(w->chunk_ev[(i - 1) & 1]);
        cudaEventDestroy(job.after);
        {
            std::lock_guard<std::mutex> lk(w->mu);
            if (e != cudaSuccess && w->err == cudaSuccess)
                w->err = e;
            w->pending -= 1;
        }
        w->cv_done.notify_all();
    }
}

int cb200_copy_async(cb200_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (ctx->copier == nullptr) {
        CopyWorker *w = new CopyWorker();
        w->device = ctx->device;
        w->ctx = ctx;
        cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i)
            e = cudaEventCreateWithFlags(&w->chunk_ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            delete w;
            return cb200_fail(ctx, "cb200_copy_async: %s", cudaGetErrorString(e));
        }
        w->th = std::thread(copy_worker_main, w);
        ctx->copier = w;
    }
    CopyJob job{dst, src, bytes, nullptr};
    CB_CUDA(ctx, cudaEventCreateWithFlags(&job.after, cudaEventDisableTiming));
    CB_CUDA(ctx, cudaEventRecord(job.after, ctx->stream));
    {
        std::lock_guard<std::mutex> lk(ctx->copier->mu);
        ctx->copier->q.push_back(job);
        ctx->copier->pending += 1;
    }
    ctx->copier->cv.notify_one();
    return 0;
}

int cb200_copies_wait(cb200_ctx *ctx)
{
    CopyWorker *w = ctx->copier;
    if (w == nullptr)
        return 0;
    cudaError_t e;
    {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv_done.wait(lk, [&] { return w->pending == 0; });
        e = w->err;
        w->err = cudaSuccess;
    }
    CB_CHECK(ctx, e == cudaSuccess, "background copy failed: %s", cudaGetErrorString(e));
    return 0;
}

void cb200_copies_shutdown(cb200_ctx *ctx)
{
    CopyWorker *w = ctx->copier;
    if (w == nullptr)
        return;
    {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv_done.wait(lk, [&] { return w->pending == 0; });
        w->stop = true;
    }
    w->cv.notify_all();
    w->th.join();
    cudaStreamDestroy(w->stream);
    cudaEventDestroy(w->chunk_ev[0]);
    cudaEventDestroy(w->chunk_ev[1]);
    delete w;
    ctx->copier = nullptr;
}

extern "C" int cb200_synchronize(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 1;
    CB_TRY(cb200_need_rowidx(ctx));   // the upload of the row indices runs on its own stream
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cb200_reset_timings(cb200_ctx *ctx)
{
    if (ctx == nullptr)
        return 1;
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    for (int s = 0; s < ST_COUNT_; ++s)
        ctx->ev_used[s] = false;
    return 0;
}

extern "C" int cb200_stage_timings(cb200_ctx *ctx, cb200_timing *out)
{
    if (ctx == nullptr || out == nullptr)
        return 1;
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double *slots[ST_COUNT_] = {&ctx->tm.h2d_ms,        &ctx->tm.colsum_ms,      &ctx->tm.lognorm_ms,
                                &ctx->tm.hvg_extract_ms, &ctx->tm.gram_ms,        &ctx->tm.eig_ms,
                                &ctx->tm.project_ms,     &ctx->tm.knn_prep_ms,    &ctx->tm.knn_tile_ms,
                                &ctx->tm.knn_rerank_ms,  &ctx->tm.snn_revlist_ms, &ctx->tm.snn_edges_ms,
                                &ctx->tm.d2h_ms};
    for (int s = 0; s < ST_COUNT_; ++s) {
        if (ctx->ev_used[s]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev[s][0], ctx->ev[s][1]) == cudaSuccess)
                *slots[s] = (double)ms;
        }
    }
    *out = ctx->tm;
    return 0;
}

extern "C" int cb200_device_ptr(cb200_ctx *ctx, int which, void **ptr, int64_t *bytes)
{
    if (ctx == nullptr || ptr == nullptr || bytes == nullptr)
        return 1;
    switch (which) {
    case CB200_BUF_LIBSIZE_PARTIAL: *ptr = ctx->part.p; *bytes = 2 * sizeof(double); break;
    case CB200_BUF_GENE_ACC: *ptr = ctx->gene_acc.p; *bytes = 2 * ctx->G * (int64_t)sizeof(double); break;
    case CB200_BUF_GRAM:
        *ptr = ctx->gram.p;
        *bytes = ((int64_t)ctx->H * ctx->H) * (int64_t)sizeof(double) * (ctx->gram_fixed ? 2 : 1);
        break;
    case CB200_BUF_SCORES: *ptr = ctx->scores.p; *bytes = ctx->N * (int64_t)ctx->d * (int64_t)sizeof(double); break;
    case CB200_BUF_KNN: *ptr = ctx->knn.p; *bytes = ctx->knn_nq * (int64_t)ctx->knn_k * (int64_t)sizeof(int32_t); break;
    case CB200_BUF_SIZE_FACTORS: *ptr = ctx->sf.p; *bytes = ctx->N * (int64_t)sizeof(double); break;
    case CB200_BUF_LOGCOUNTS: *ptr = ctx->logx.p; *bytes = ctx->nnz * (int64_t)sizeof(double); break;
    case CB200_BUF_BLOCK_PART: *ptr = ctx->blk_part.p; *bytes = (2 * (int64_t)ctx->n_blocks + 1) * (int64_t)sizeof(double); break;
    case CB200_BUF_BLOCK_ACC: *ptr = ctx->gene_acc.p; *bytes = 2 * ctx->G * (int64_t)ctx->n_blocks * (int64_t)sizeof(double); break;
    default: return cb200_fail(ctx, "cb200_device_ptr: unknown buffer id %d", which);
    }
    if (*ptr == nullptr)
        return cb200_fail(ctx, "cb200_device_ptr: buffer %d has not been produced yet", which);
    return 0;
}

// ---------------------------------------------------------------------------------------
// a1: counts upload
// ---------------------------------------------------------------------------------------
__global__ void k_widen_colptr(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = (int64_t)in[i];
}

static int load_common(cb200_ctx *ctx, int64_t n_genes, int64_t n_cells, int64_t nnz,
                       const int32_t *rowidx, const double *values, int64_t cell_offset,
                       int64_t n_cells_total)
{
    CB_CHECK(ctx, n_genes > 0 && n_genes < 2147483647LL, "cb200_load_csc: n_genes must be in [1, 2^31)");
    CB_CHECK(ctx, n_cells > 0, "cb200_load_csc: n_cells must be positive");
    CB_CHECK(ctx, cell_offset >= 0 && cell_offset + n_cells <= n_cells_total,
             "cb200_load_csc: shard [%lld, %lld) does not fit n_cells_total=%lld", (long long)cell_offset,
             (long long)(cell_offset + n_cells), (long long)n_cells_total);
    // a background copy-out of the previous matrix's results may still be reading buffers that are re-used below
    CB_TRY(cb200_copies_wait(ctx));
    CB_TRY(cb200_upload_join(ctx));
    ctx->G = n_genes;
    ctx->N = n_cells;
    ctx->nnz = nnz;
    ctx->cell_offset = cell_offset;
    ctx->N_total = n_cells_total;
    CB_CUDA(ctx, ctx->rowidx.ensure((size_t)nnz + 16));   // slack: the bulk copies of the gene-sum kernel round
    CB_CUDA(ctx, ctx->x.ensure((size_t)nnz + 16));        // their source ranges out to 16 bytes
    // values first: size factors and log-counts need only them (and the column pointers), so that part of
    // the path -- and the copy-out of the log-counts -- runs while the row indices are still arriving
    if (ctx->up_stream == nullptr) {
        CB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
        CB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_x, cudaEventDisableTiming));
        CB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_rowidx, cudaEventDisableTiming));
    }
    // (pageable host memory -- R vectors -- goes through the pinned staging ring, cb200_stage.cu)
    CB_TRY(cb200_h2d(ctx, ctx->x.p, values, (size_t)nnz * sizeof(double), ctx->stream));
    CB_CUDA(ctx, cudaEventRecord(ctx->ev_x, ctx->stream));
    CB_CUDA(ctx, cudaStreamWaitEvent(ctx->up_stream, ctx->ev_x, 0));
    if ((size_t)nnz * sizeof(int32_t) >= (64u << 20) && !cb200_host_is_pinned(rowidx)) {
        // the indices keep arriving on a thread of their own after this call has returned
        CB_TRY(cb200_upload_background(ctx, ctx->rowidx.p, rowidx, (size_t)nnz * sizeof(int32_t), ctx->up_stream,
                                       ctx->ev_rowidx));
    } else {
        CB_TRY(cb200_h2d(ctx, ctx->rowidx.p, rowidx, (size_t)nnz * sizeof(int32_t), ctx->up_stream));
        CB_CUDA(ctx, cudaEventRecord(ctx->ev_rowidx, ctx->up_stream));
    }
    ctx->rowidx_pending = true;
    ctx->logx_values_only = false;
    ctx->loaded = true;
    ctx->gs_split_ready = ctx->gs_tables_ready = false;
    ctx->n_blocks = 0;
    ctx->have_sf = ctx->have_logx = ctx->have_genestats = ctx->have_gram = ctx->have_scores = false;
    ctx->have_knn = false;
    return 0;
}

int cb200_need_rowidx(cb200_ctx *ctx)
{
    if (ctx->rowidx_pending) {
        CB_TRY(cb200_upload_join(ctx));   // the event is recorded by the uploader thread when the last chunk is queued
        CB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_rowidx, 0));
        ctx->rowidx_pending = false;
    }
    return 0;
}

extern "C" int cb200_load_csc(cb200_ctx *ctx, int64_t n_genes, int64_t n_cells, const int64_t *colptr,
                              const int32_t *rowidx, const double *values, int64_t cell_offset,
                              int64_t n_cells_total)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, colptr != nullptr && rowidx != nullptr && values != nullptr, "cb200_load_csc: NULL input");
    CB_CHECK(ctx, n_cells > 0, "cb200_load_csc: n_cells must be positive");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t base = colptr[0];
    int64_t nnz = colptr[n_cells] - base;
    CB_CHECK(ctx, base == 0, "cb200_load_csc: colptr[0] must be 0 (rebase the shard's column pointers)");
    CB_CHECK(ctx, nnz >= 0, "cb200_load_csc: colptr is not non-decreasing");
    for (int64_t c = 0; c < n_cells; ++c)
        CB_CHECK(ctx, colptr[c + 1] >= colptr[c], "cb200_load_csc: colptr is not non-decreasing (column %lld)", (long long)c);
    cb_stage_begin(ctx, ST_H2D);
    CB_CUDA(ctx, ctx->colptr.ensure((size_t)n_cells + 1));
    CB_CUDA(ctx, cudaMemcpyAsync(ctx->colptr.p, colptr, (size_t)(n_cells + 1) * sizeof(int64_t),
                                 cudaMemcpyHostToDevice, ctx->stream));
    CB_TRY(load_common(ctx, n_genes, n_cells, nnz, rowidx, values, cell_offset, n_cells_total));
    cb_stage_end(ctx, ST_H2D);
    return 0;
}

extern "C" int cb200_load_csc_i32(cb200_ctx *ctx, int64_t n_genes, int64_t n_cells, const int32_t *colptr,
                                  const int32_t *rowidx, const double *values, int64_t cell_offset,
                                  int64_t n_cells_total)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, colptr != nullptr && rowidx != nullptr && values != nullptr, "cb200_load_csc_i32: NULL input");
    CB_CHECK(ctx, n_cells > 0, "cb200_load_csc_i32: n_cells must be positive");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    CB_CHECK(ctx, colptr[0] == 0, "cb200_load_csc_i32: p[0] must be 0");
    int64_t nnz = colptr[n_cells];
    CB_CHECK(ctx, nnz >= 0, "cb200_load_csc_i32: p is not non-decreasing");
    for (int64_t c = 0; c < n_cells; ++c)
        CB_CHECK(ctx, colptr[c + 1] >= colptr[c], "cb200_load_csc_i32: p is not non-decreasing (column %lld)", (long long)c);
    cb_stage_begin(ctx, ST_H2D);
    CB_CUDA(ctx, ctx->colptr.ensure((size_t)n_cells + 1));
    CB_CUDA(ctx, ctx->stage_i.ensure((size_t)n_cells + 1));
    CB_CUDA(ctx, cudaMemcpyAsync(ctx->stage_i.p, colptr, (size_t)(n_cells + 1) * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, ctx->stream));
    k_widen_colptr<<<(unsigned)((n_cells + 1 + 255) / 256), 256, 0, ctx->stream>>>(ctx->stage_i.p, n_cells + 1,
                                                                                    ctx->colptr.p);
    CB_LAUNCH(ctx);
    CB_TRY(load_common(ctx, n_genes, n_cells, nnz, rowidx, values, cell_offset, n_cells_total));
    cb_stage_end(ctx, ST_H2D);
    return 0;
}

extern "C" int cb200_export_csc(cb200_ctx *ctx, int64_t *colptr, int32_t *rowidx, double *values)
{
    if (ctx == nullptr)
        return 1;
    CB_CHECK(ctx, ctx->loaded, "cb200_export_csc: no count matrix in the context");
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (colptr != nullptr)
        CB_CUDA(ctx, cudaMemcpyAsync(colptr, ctx->colptr.p, (size_t)(ctx->N + 1) * sizeof(int64_t),
                                     cudaMemcpyDeviceToHost, ctx->stream));
    CB_TRY(cb200_need_rowidx(ctx));
    if (rowidx != nullptr)
        CB_CUDA(ctx, cudaMemcpyAsync(rowidx, ctx->rowidx.p, (size_t)ctx->nnz * sizeof(int32_t),
                                     cudaMemcpyDeviceToHost, ctx->stream));
    if (values != nullptr)
        CB_CUDA(ctx, cudaMemcpyAsync(values, ctx->x.p, (size_t)ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------
// exclusive scan int32 -> int64 (three launches; n up to a few million)
// ---------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int64_t block_exclusive_scan_i64(int64_t v, int64_t *total)
{
    // blockDim.x == SCAN_THREADS
    __shared__ int64_t warp_tot[SCAN_THREADS / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    if (lane == 31)
        warp_tot[w] = inc;
    __syncthreads();
    int64_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < SCAN_THREADS / 32; ++i) {
        int64_t t = warp_tot[i];
        if (i < w)
            base += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const int32_t *__restrict__ in, int64_t n,
                                                                 int64_t *__restrict__ tile_sum)
{
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        if (idx < n)
            s += in[idx];
    }
    int64_t tot;
    block_exclusive_scan_i64(s, &tot);
    if (threadIdx.x == 0)
        tile_sum[blockIdx.x] = tot;
}

// single block: exclusive scan of the tile sums in place (n_tiles arbitrary)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_offsets(int64_t *__restrict__ tile_sum, int64_t n_tiles,
                                                                    int64_t *__restrict__ grand_total)
{
    int64_t carry = 0;
    for (int64_t b = 0; b < n_tiles; b += SCAN_THREADS) {
        int64_t idx = b + threadIdx.x;
        int64_t v = idx < n_tiles ? tile_sum[idx] : 0;
        int64_t tot;
        int64_t ex = block_exclusive_scan_i64(v, &tot);
        if (idx < n_tiles)
            tile_sum[idx] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0)
        *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_finish(const int32_t *__restrict__ in, int64_t n,
                                                              const int64_t *__restrict__ tile_off,
                                                              int64_t *__restrict__ out)
{
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t vals[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        vals[i] = idx < n ? (int64_t)in[idx] : 0;
        s += vals[i];
    }
    int64_t tot;
    int64_t ex = block_exclusive_scan_i64(s, &tot) + tile_off[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        if (idx < n)
            out[idx] = ex;
        ex += vals[i];
    }
}

int cb200_scan_i32_to_i64(cb200_ctx *ctx, const int32_t *d_in, int64_t n, int64_t *d_out)
{
    if (n <= 0) {
        CB_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(int64_t), ctx->stream));
        return 0;
    }
    int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    CB_CUDA(ctx, ctx->scan_blk.ensure((size_t)n_tiles + 1));
    k_scan_tile_sums<<<(unsigned)n_tiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, ctx->scan_blk.p);
    CB_LAUNCH(ctx);
    k_scan_tile_offsets<<<1, SCAN_THREADS, 0, ctx->stream>>>(ctx->scan_blk.p, n_tiles, d_out + n);
    CB_LAUNCH(ctx);
    k_scan_finish<<<(unsigned)n_tiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, ctx->scan_blk.p, d_out);
    CB_LAUNCH(ctx);
    return 0;
}
