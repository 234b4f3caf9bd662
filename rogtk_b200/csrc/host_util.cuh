// host_util.cuh -- host-side helpers shared by the orchestration code: error plumbing and a stream-ordered device buffer.
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include <string>

namespace rogtk {

struct CudaFail {
    cudaError_t e;
    const char *what;
};

#define CK(call)                                                                 \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) throw CudaFail{_e, #call};                        \
    } while (0)

struct ArgFail {
    int code;
    std::string msg;
};

// The memory pool stream-ordered allocations of this thread come from: every context owns one (created in rogtk_gpu_create,
// destroyed with the context), so a context neither pins freed memory in the device's default pool for the rest of the process
// nor trims other users' cached memory when it goes away.  Entry points set it for their duration (PoolScope).
inline thread_local cudaMemPool_t tl_pool = nullptr;
struct PoolScope {
    cudaMemPool_t prev;
    explicit PoolScope(cudaMemPool_t p) : prev(tl_pool) { tl_pool = p; }
    ~PoolScope() { tl_pool = prev; }
};
inline cudaError_t pool_malloc(void **p, size_t n, cudaStream_t s)
{
    return tl_pool ? cudaMallocFromPoolAsync(p, n, tl_pool, s) : cudaMallocAsync(p, n, s);
}

// stream-ordered device buffer
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t st = nullptr;
    DevBuf() = default;
    DevBuf(size_t n, cudaStream_t s) { alloc(n, s); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes), st(o.st) { o.p = nullptr; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        release();
        p = o.p; bytes = o.bytes; st = o.st; o.p = nullptr;
        return *this;
    }
    void alloc(size_t n, cudaStream_t s)
    {
        release();
        st = s;
        bytes = n;
        CK(pool_malloc(&p, n ? n : 16, s));
    }
    void release()
    {
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
    }
    void *detach()
    {
        void *q = p;
        p = nullptr;
        return q;
    }
    ~DevBuf() { release(); }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// Small device->host reads (counters the host sizes the next allocation with) that do not touch the copy engines: a tiny
// kernel stores the words into mapped pinned memory.  A cudaMemcpy of 8 bytes queued while another stream runs a bulk D2H
// waits for that whole transfer (measured: ~4 ms per read inside the chunked host path), this does not.
__global__ void k_peek(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int n_words)
{
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
}

struct Peek {
    static constexpr size_t SLOT_BYTES = 4096;
    uint8_t *slot;        // mapped pinned, SLOT_BYTES (owned by the context)
    cudaStream_t st;
    size_t used = 0;
    struct Item { void *h; size_t off, bytes; };
    Item items[8];
    int n = 0;
    int64_t *launches;    // launch counter of the context (may be null)
    Peek(uint8_t *slot_, cudaStream_t st_, int64_t *launches_ = nullptr) : slot(slot_), st(st_), launches(launches_) {}
    // bytes: multiple of 4, source 4-byte aligned
    void add(const void *dsrc, size_t bytes, void *hdst)
    {
        if (n == 8 || used + bytes > SLOT_BYTES || (bytes & 3)) throw ArgFail{-1, "internal: peek slot overflow"};
        k_peek<<<1, 64, 0, st>>>(reinterpret_cast<const uint32_t *>(dsrc), reinterpret_cast<uint32_t *>(slot + used), (int)(bytes / 4));
        if (launches) ++*launches;
        items[n++] = Item{hdst, used, bytes};
        used += (bytes + 15) & ~(size_t)15;
    }
    void sync()
    {
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < n; i++) memcpy(items[i].h, slot + items[i].off, items[i].bytes);
        n = 0;
        used = 0;
    }
};

}  // namespace rogtk
