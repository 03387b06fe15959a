// staging.cu -- host->device path of trajectory chunks (row F3 of SURVEY.md section 8).
//
// The reference pulls 1000-frame chunks out of HDF5 into NumPy arrays
// (mdlmc/IO/trajectory_parser.py:296,322); here such a chunk has to reach HBM while the kernels of
// the previous chunk run.  A DMA engine only overlaps with kernels when the source is page-locked,
// so:
//   * cmd_host_alloc / cmd_host_free hand out page-locked buffers -- a reader that fills those
//     (runtime.pinned_empty on the Python side) is copied from directly;
//   * every other (pageable) pointer goes through a library-owned ring of CMD_RING_SLOTS page-locked
//     slots: a few host threads copy piece i+1 into its slot while the DMA engine moves piece i and
//     the SMs work on the chunk before.  cmd_h2d_staged never blocks on the GPU except to recycle a
//     slot that is still in flight.
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

#define CMD_RING_SLOTS 3
#define CMD_RING_SLOT_BYTES ((size_t)8 << 20)
#define CMD_RING_MAX_THREADS 8

namespace {

// a tiny persistent pool: run(n, fn) executes fn(0..n-1), the caller takes part 0
class CopyPool {
public:
    explicit CopyPool(int helpers) : stop_(false), gen_(0), pending_(0)
    {
        for (int i = 0; i < helpers; i++) workers_.emplace_back([this, i] { loop(i + 1); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    int parts() const { return (int)workers_.size() + 1; }
    void copy(char *dst, const char *src, size_t bytes)
    {
        const int p = bytes < ((size_t)1 << 20) ? 1 : parts();
        if (p == 1) { memcpy(dst, src, bytes); return; }
        {
            std::lock_guard<std::mutex> l(m_);
            dst_ = dst; src_ = src; bytes_ = bytes; nparts_ = p;
            pending_ = p - 1;
            gen_++;
        }
        cv_.notify_all();
        part(0);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
    }

private:
    void part(int i)
    {
        const size_t per = (bytes_ / nparts_ + 63) & ~(size_t)63;
        const size_t lo = per * i < bytes_ ? per * i : bytes_;
        const size_t hi = i == nparts_ - 1 ? bytes_ : (lo + per < bytes_ ? lo + per : bytes_);
        if (hi > lo) memcpy(dst_ + lo, src_ + lo, hi - lo);
    }
    void loop(int id)
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                if (id >= nparts_) continue;
            }
            part(id);
            {
                std::lock_guard<std::mutex> l(m_);
                pending_--;
            }
            done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    bool stop_;
    unsigned long long gen_;
    int pending_, nparts_ = 1;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0;
};

struct Ring {
    void *slot[CMD_RING_SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[CMD_RING_SLOTS] = {0, 0, 0};
    bool busy[CMD_RING_SLOTS] = {false, false, false};
    int next = 0;
    CopyPool *pool = nullptr;
    unsigned long long staged_bytes = 0, direct_bytes = 0;
};

Ring &ring()
{
    static Ring r;
    return r;
}

int ring_ready()
{
    Ring &r = ring();
    if (r.slot[0]) return CMD_OK;
    for (int i = 0; i < CMD_RING_SLOTS; i++) {
        if (cudaHostAlloc(&r.slot[i], CMD_RING_SLOT_BYTES, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            for (int k = 0; k < i; k++) { cudaFreeHost(r.slot[k]); r.slot[k] = nullptr; }
            return cmd_set_error(CMD_ENOMEM, "cudaHostAlloc of the %zu-byte staging ring failed",
                                 CMD_RING_SLOTS * CMD_RING_SLOT_BYTES);
        }
        CMD_CUDA(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
    }
    int threads = 4;
    if (const char *e = getenv("CMDLMC_B200_STAGE_THREADS")) threads = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && threads > hw) threads = hw;
    if (threads < 1) threads = 1;
    if (threads > CMD_RING_MAX_THREADS) threads = CMD_RING_MAX_THREADS;
    r.pool = new CopyPool(threads - 1);
    return CMD_OK;
}

bool is_page_locked(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

}  // namespace

// dst (device) <- src (host), `bytes`, ordered on `stream`.  Page-locked sources are copied from
// directly; pageable ones go through the ring.  On return the caller may reuse `src` only in the
// pageable case (page-locked sources are read asynchronously, like cudaMemcpyAsync).
int cmd_h2d_staged(void *dst, const void *src, size_t bytes, cudaStream_t stream)
{
    if (!bytes) return CMD_OK;
    Ring &r = ring();
    if (is_page_locked(src)) {
        r.direct_bytes += bytes;
        CMD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return CMD_OK;
    }
    int rc = ring_ready();
    if (rc) return rc;
    for (size_t off = 0; off < bytes; off += CMD_RING_SLOT_BYTES) {
        const size_t nb = bytes - off < CMD_RING_SLOT_BYTES ? bytes - off : CMD_RING_SLOT_BYTES;
        const int s = r.next;
        r.next = (r.next + 1) % CMD_RING_SLOTS;
        if (r.busy[s]) CMD_CUDA(cudaEventSynchronize(r.ev[s]));   // its last DMA has to be done
        r.pool->copy((char *)r.slot[s], (const char *)src + off, nb);
        CMD_CUDA(cudaMemcpyAsync((char *)dst + off, r.slot[s], nb, cudaMemcpyHostToDevice, stream));
        CMD_CUDA(cudaEventRecord(r.ev[s], stream));
        r.busy[s] = true;
    }
    r.staged_bytes += bytes;
    return CMD_OK;
}

void cmd_staging_shutdown()
{
    Ring &r = ring();
    for (int i = 0; i < CMD_RING_SLOTS; i++) {
        if (r.busy[i]) cudaEventSynchronize(r.ev[i]);
        if (r.ev[i]) cudaEventDestroy(r.ev[i]);
        if (r.slot[i]) cudaFreeHost(r.slot[i]);
        r.slot[i] = nullptr; r.ev[i] = 0; r.busy[i] = false;
    }
    delete r.pool;
    r.pool = nullptr;
    r.next = 0;
}

extern "C" int cmd_host_alloc(size_t bytes, void **out)
{
    CMD_REQUIRE_INIT();
    if (!out || !bytes) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (cudaHostAlloc(out, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return cmd_set_error(CMD_ENOMEM, "cudaHostAlloc of %zu bytes failed", bytes);
    }
    return CMD_OK;
}

extern "C" int cmd_host_free(void *p)
{
    if (!p) return CMD_OK;
    CMD_CUDA(cudaFreeHost(p));
    return CMD_OK;
}

extern "C" int cmd_staging_stats(uint64_t *staged_bytes, uint64_t *direct_bytes, int *threads)
{
    Ring &r = ring();
    if (staged_bytes) *staged_bytes = r.staged_bytes;
    if (direct_bytes) *direct_bytes = r.direct_bytes;
    if (threads) *threads = r.pool ? r.pool->parts() : 0;
    return CMD_OK;
}
