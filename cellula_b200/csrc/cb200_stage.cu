// cb200_stage.cu -- host boundary: copies between PAGEABLE host memory and HBM through a pinned staging ring.
//
// The reference's host is R: every input (the dgCMatrix slots i / p / x, a reducedDim matrix) and every output
// (allocVector'ed numeric / integer vectors) is ordinary pageable memory (hard part 4).  A plain
// cudaMemcpy from such memory is staged by the driver through one small internal buffer by one thread; at cfg3
// (23.5 GB in, 24.7 GB out) that is the whole run time.  Here instead:
//   * a pool of host threads copies 16 MB chunks between the caller's memory and a ring of pinned buffers
//     (several threads per chunk: one core's memcpy does not fill a PCIe 5 x16 link),
//   * every chunk is ONE cudaMemcpyAsync between the pinned buffer and HBM, so the DMA of chunk c overlaps the
//     host-side memcpy of chunk c + 1.
// Page-locked caller memory (bench.py's pinned leg, a host that registered its buffers) is recognised with
// cudaPointerGetAttributes and copied directly.  Registering R's memory on the fly (cudaHostRegister) costs more
// than the copy itself and is not attempted.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "cb200_common.cuh"

namespace {
constexpr size_t STAGE_CHUNK = 16u << 20;
constexpr int STAGE_RING = 4;

// fixed pool of threads that split one memcpy between them; one job at a time (host DRAM bandwidth is the limit)
class MemcpyPool {
public:
    explicit MemcpyPool(int n_threads) : n_(n_threads)
    {
        for (int t = 1; t < n_; ++t)
            th_.emplace_back([this, t] { loop(t); });
    }
    ~MemcpyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto &t : th_)
            t.join();
    }
    void copy(void *dst, const void *src, size_t bytes)
    {
        if (bytes < (1u << 20) || n_ <= 1) {
            memcpy(dst, src, bytes);
            return;
        }
        std::lock_guard<std::mutex> job(job_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = static_cast<char *>(dst);
            src_ = static_cast<const char *>(src);
            bytes_ = bytes;
            left_ = n_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        piece(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return left_ == 0; });
    }
    int threads() const { return n_; }

private:
    void piece(int t)
    {
        const size_t per = ((bytes_ + n_ - 1) / n_ + 4095) & ~(size_t)4095;
        const size_t a = std::min(bytes_, per * (size_t)t), b = std::min(bytes_, a + per);
        if (b > a)
            memcpy(dst_ + a, src_ + a, b - a);
    }
    void loop(int t)
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_)
                    return;
            }
            piece(t);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --left_;
            }
            done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_, done_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0;
    int left_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};
}  // namespace

struct StagePool {
    MemcpyPool pool;
    // one ring per concurrent user: [0] the calling thread (cb200_h2d / cb200_d2h), [1] the background upload of the
    // row indices, [2] the background device->host copy worker
    unsigned char *buf[3][STAGE_RING] = {};
    cudaEvent_t ev[3][STAGE_RING] = {};
    bool ev_armed[3][STAGE_RING] = {};
    std::thread uploader;
    std::atomic<int> upload_rc{0};
    explicit StagePool(int n) : pool(n) {}
};

static int stage_threads()
{
    const char *e = getenv("CB200_STAGE_THREADS");
    if (e != nullptr && atoi(e) > 0)
        return atoi(e);
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(2u, std::min(8u, hw / 2));
}

static int stage_pool(cb200_ctx *ctx, int ring, StagePool **out)
{
    if (ctx->stage == nullptr)
        ctx->stage = new StagePool(stage_threads());
    StagePool *sp = ctx->stage;
    if (sp->buf[ring][0] == nullptr) {
        for (int i = 0; i < STAGE_RING; ++i) {
            CB_CUDA(ctx, cudaMallocHost(reinterpret_cast<void **>(&sp->buf[ring][i]), STAGE_CHUNK));
            CB_CUDA(ctx, cudaEventCreateWithFlags(&sp->ev[ring][i], cudaEventDisableTiming));
            sp->ev_armed[ring][i] = false;
        }
    }
    *out = sp;
    return 0;
}

bool cb200_host_is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();   // pre-11.0 behaviour for unregistered memory: clear the sticky error
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

static cudaError_t wait_event_polite(cudaEvent_t ev)
{
    for (;;) {
        const cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady)
            return e;
        std::this_thread::yield();
    }
}

// host -> device.  Pinned source: one asynchronous copy (the source must stay valid until the stream has passed it).
// Pageable source: staged; returns once the last chunk has left the caller's memory (the DMA may still be running).
static int h2d_ring(cb200_ctx *ctx, int ring, void *dev_dst, const void *host_src, size_t bytes, cudaStream_t stream)
{
    StagePool *sp;
    CB_TRY(stage_pool(ctx, ring, &sp));
    size_t i = 0;
    for (size_t off = 0; off < bytes; off += STAGE_CHUNK, ++i) {
        const size_t nb = std::min(STAGE_CHUNK, bytes - off);
        const int b = (int)(i % STAGE_RING);
        if (sp->ev_armed[ring][b])
            CB_CUDA(ctx, wait_event_polite(sp->ev[ring][b]));
        sp->pool.copy(sp->buf[ring][b], static_cast<const char *>(host_src) + off, nb);
        CB_CUDA(ctx, cudaMemcpyAsync(static_cast<char *>(dev_dst) + off, sp->buf[ring][b], nb, cudaMemcpyHostToDevice, stream));
        CB_CUDA(ctx, cudaEventRecord(sp->ev[ring][b], stream));
        sp->ev_armed[ring][b] = true;
    }
    return 0;
}

int cb200_h2d(cb200_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0)
        return 0;
    if (bytes < (4u << 20) || cb200_host_is_pinned(host_src)) {
        CB_CUDA(ctx, cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, stream));
        return 0;
    }
    return h2d_ring(ctx, 0, dev_dst, host_src, bytes, stream);
}

// device -> host, ordered behind the work already in `stream`.  Pinned destination: asynchronous (the caller
// synchronises).  Pageable destination: complete on return.
static int d2h_ring(cb200_ctx *ctx, int ring, void *host_dst, const void *dev_src, size_t bytes, cudaStream_t stream)
{
    StagePool *sp;
    CB_TRY(stage_pool(ctx, ring, &sp));
    const size_t n_chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    // chunk c is in flight (device -> ring) while chunk c - 1 is copied out by the host threads
    for (size_t c = 0; c <= n_chunks; ++c) {
        if (c < n_chunks) {
            const size_t off = c * STAGE_CHUNK, nb = std::min(STAGE_CHUNK, bytes - off);
            const int b = (int)(c % STAGE_RING);
            CB_CUDA(ctx, cudaMemcpyAsync(sp->buf[ring][b], static_cast<const char *>(dev_src) + off, nb, cudaMemcpyDeviceToHost,
                                         stream));
            CB_CUDA(ctx, cudaEventRecord(sp->ev[ring][b], stream));
            sp->ev_armed[ring][b] = false;   // consumed below, before the buffer is reused (STAGE_RING >= 2)
        }
        if (c >= 1) {
            const size_t off = (c - 1) * STAGE_CHUNK, nb = std::min(STAGE_CHUNK, bytes - off);
            const int b = (int)((c - 1) % STAGE_RING);
            CB_CUDA(ctx, wait_event_polite(sp->ev[ring][b]));
            sp->pool.copy(static_cast<char *>(host_dst) + off, sp->buf[ring][b], nb);
        }
    }
    return 0;
}

int cb200_d2h(cb200_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0)
        return 0;
    if (bytes < (4u << 20) || cb200_host_is_pinned(host_dst)) {
        CB_CUDA(ctx, cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, stream));
        return 0;
    }
    return d2h_ring(ctx, 0, host_dst, dev_src, bytes, stream);
}

// used by the background copy worker (its own thread, its own ring)
int cb200_d2h_worker(cb200_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes, cudaStream_t stream)
{
    return d2h_ring(ctx, 2, host_dst, dev_src, bytes, stream);
}

// Upload that keeps running after the call has returned (the row indices of a dgCMatrix behind its values): a
// thread stages the pageable source through ring 1 into `stream` and records `done` when the last chunk is queued.
// cb200_upload_join must be called before anything waits on `done`.
int cb200_upload_background(cb200_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes, cudaStream_t stream,
                            cudaEvent_t done)
{
    StagePool *sp;
    CB_TRY(stage_pool(ctx, 1, &sp));
    CB_TRY(cb200_upload_join(ctx));
    sp->upload_rc = 0;
    const int device = ctx->device;
    sp->uploader = std::thread([=] {
        cudaSetDevice(device);
        int rc = h2d_ring(ctx, 1, dev_dst, host_src, bytes, stream);
        if (rc == 0 && cudaEventRecord(done, stream) != cudaSuccess)
            rc = 1;
        ctx->stage->upload_rc = rc;
    });
    return 0;
}

int cb200_upload_join(cb200_ctx *ctx)
{
    if (ctx->stage != nullptr && ctx->stage->uploader.joinable()) {
        ctx->stage->uploader.join();
        CB_CHECK(ctx, ctx->stage->upload_rc == 0, "background upload of the row indices failed");
    }
    return 0;
}

void cb200_stage_shutdown(cb200_ctx *ctx)
{
    StagePool *sp = ctx->stage;
    if (sp == nullptr)
        return;
    if (sp->uploader.joinable())
        sp->uploader.join();
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < STAGE_RING; ++i)
            if (sp->buf[r][i] != nullptr) {
                cudaFreeHost(sp->buf[r][i]);
                cudaEventDestroy(sp->ev[r][i]);
            }
    delete sp;
    ctx->stage = nullptr;
}

int cb200_stage_thread_count(cb200_ctx *ctx)
{
    StagePool *sp;
    if (stage_pool(ctx, 0, &sp) != 0)
        return 0;
    return sp->pool.threads();
}
