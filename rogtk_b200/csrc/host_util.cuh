// host_util.cuh -- host-side helpers shared by the orchestration code: error plumbing and a stream-ordered device buffer.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include <cstdlib>

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

// Device memory of a context: a caching allocator over cudaMalloc, private to the context.
//
// Why not cudaMallocAsync: measured on B200 (tools/csr_probe.py, 200 k UMIs), a stream-ordered pool that is asked for a new mix of
// sizes (a CSR batch after plain batches, chunks of different sizes) re-maps its physical pages between virtual ranges for the
// next several calls -- the same kernels then take 38-77 ms per batch instead of 27 ms, with the time showing up inside whichever
// stage happens to allocate.  Here a freed block keeps its address and size and is handed out again as a whole (no splitting, no
// remapping): sizes are rounded to 1/8-octave classes so batches of similar size hit the same blocks, a request may take a cached
// block up to 25 % larger, and cudaMalloc is only reached while the working set is still growing.  If cudaMalloc fails, every
// cached block is returned to the driver and the request retried once.
//
// Ordering: free(p, s) records an event on s; a later alloc on another stream waits for it (same stream: program order suffices).
// Entry points set the cache of the calling thread for their duration (CacheScope).
//
// Several contexts may live in one process (one per Polars worker thread): when cudaMalloc fails, the idle blocks
// of EVERY cache on the device are released before the request is retried (registry below; never two cache locks held at once), and
// ROGTK_DEVICE_CACHE_BYTES bounds the idle bytes a cache keeps (largest blocks go first).
struct DevCache;
struct DevCacheRegistry {
    std::mutex mu;
    std::vector<DevCache *> all;
    static DevCacheRegistry &get()
    {
        static DevCacheRegistry r;
        return r;
    }
};
struct DevCache {
    struct Block {
        void *p;
        size_t sz;
        cudaEvent_t ev;
        cudaStream_t st;
    };
    std::mutex mu;
    std::multimap<size_t, Block> idle;            // by size
    std::unordered_map<void *, Block> live;
    size_t idle_bytes = 0, live_bytes = 0, peak_bytes = 0;
    size_t idle_limit = ~(size_t)0;
    int64_t n_malloc = 0, n_reuse = 0, n_released_for_others = 0;
    int device = 0;

    explicit DevCache(int dev = 0) : device(dev)
    {
        if (const char *e = getenv("ROGTK_DEVICE_CACHE_BYTES")) idle_limit = (size_t)strtoull(e, nullptr, 10);
        DevCacheRegistry &r = DevCacheRegistry::get();
        std::lock_guard<std::mutex> lk(r.mu);
        r.all.push_back(this);
    }
    // cudaMalloc failed: hand idle blocks of the caches on this device back to the driver, largest first, until about twice the
    // request has been freed (everything when need == 0); the caller holds no cache lock
    static void release_all_idle(int dev, size_t need = 0)
    {
        DevCacheRegistry &r = DevCacheRegistry::get();
        std::lock_guard<std::mutex> lk(r.mu);
        cudaDeviceSynchronize();
        size_t freed = 0;
        while (true) {
            DevCache *best = nullptr;
            size_t best_sz = 0;
            for (DevCache *c : r.all) {
                if (c->device != dev) continue;
                std::lock_guard<std::mutex> cl(c->mu);
                if (!c->idle.empty() && std::prev(c->idle.end())->first > best_sz) {
                    best = c;
                    best_sz = std::prev(c->idle.end())->first;
                }
            }
            if (!best) break;
            {
                std::lock_guard<std::mutex> cl(best->mu);
                if (best->idle.empty()) continue;
                auto last = std::prev(best->idle.end());
                cudaEventDestroy(last->second.ev);
                cudaFree(last->second.p);
                freed += last->second.sz;
                best->idle_bytes -= last->second.sz;
                best->idle.erase(last);
                best->n_released_for_others++;
            }
            if (need && freed >= 2 * need) break;
        }
    }
    static size_t size_class(size_t n)
    {
        if (n < 512) return 512;
        int top = 63 - __builtin_clzll((unsigned long long)n);
        size_t step = n <= (1u << 20) ? ((size_t)1 << top) : ((size_t)1 << (top - 3));   // small: powers of two; large: 1/8 octave
        return (n + step - 1) & ~(step - 1);
    }
    cudaError_t alloc(void **out, size_t n, cudaStream_t s)
    {
        cudaError_t e = alloc_once(out, n, s);
        if (e != cudaErrorMemoryAllocation) return e;
        cudaGetLastError();
        release_all_idle(device, size_class(n));       // enough for this request ...
        e = alloc_once(out, n, s);
        if (e != cudaErrorMemoryAllocation) return e;
        cudaGetLastError();
        release_all_idle(device);                      // ... or everything (fragmented address space)
        return alloc_once(out, n, s);
    }
    cudaError_t alloc_once(void **out, size_t n, cudaStream_t s)
    {
        const size_t sz = size_class(n);
        std::lock_guard<std::mutex> lk(mu);
        auto it = idle.lower_bound(sz);
        if (it != idle.end() && it->first <= sz + (sz > (1u << 20) ? sz / 4 : 0)) {
            Block b = it->second;
            idle.erase(it);
            idle_bytes -= b.sz;
            if (b.st != s) {
                cudaError_t e = cudaStreamWaitEvent(s, b.ev, 0);
                if (e != cudaSuccess) return e;
            }
            live_bytes += b.sz;
            live.emplace(b.p, b);
            n_reuse++;
            *out = b.p;
            return cudaSuccess;
        }
        Block b{nullptr, sz, nullptr, s};
        cudaError_t e = cudaMalloc(&b.p, sz);
        if (e != cudaSuccess) return e;
        e = cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            cudaFree(b.p);
            return e;
        }
        n_malloc++;
        live_bytes += sz;
        peak_bytes = std::max(peak_bytes, live_bytes + idle_bytes);
        live.emplace(b.p, b);
        *out = b.p;
        return cudaSuccess;
    }
    void free(void *p, cudaStream_t s)
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = live.find(p);
        if (it == live.end()) return;   // not ours (never happens for pointers this library handed out)
        Block b = it->second;
        live.erase(it);
        live_bytes -= b.sz;
        b.st = s;
        cudaEventRecord(b.ev, s);
        idle_bytes += b.sz;
        idle.emplace(b.sz, b);
        while (idle_bytes > idle_limit && !idle.empty()) {   // bounded cache: the largest idle block goes back to the driver
            auto last = std::prev(idle.end());
            cudaEventSynchronize(last->second.ev);
            cudaEventDestroy(last->second.ev);
            cudaFree(last->second.p);
            idle_bytes -= last->second.sz;
            idle.erase(last);
        }
    }
    void release_idle_locked()
    {
        cudaDeviceSynchronize();
        for (auto &kv : idle) {
            cudaEventDestroy(kv.second.ev);
            cudaFree(kv.second.p);
        }
        idle.clear();
        idle_bytes = 0;
    }
    void release_idle()
    {
        std::lock_guard<std::mutex> lk(mu);
        release_idle_locked();
    }
    ~DevCache()
    {
        {
            DevCacheRegistry &r = DevCacheRegistry::get();
            std::lock_guard<std::mutex> lk(r.mu);
            r.all.erase(std::remove(r.all.begin(), r.all.end(), this), r.all.end());
        }
        release_idle();
        for (auto &kv : live) {   // blocks the caller never freed go with the context
            cudaEventDestroy(kv.second.ev);
            cudaFree(kv.second.p);
        }
    }
};

inline thread_local DevCache *tl_cache = nullptr;
struct CacheScope {
    DevCache *prev;
    explicit CacheScope(DevCache *c) : prev(tl_cache) { tl_cache = c; }
    ~CacheScope() { tl_cache = prev; }
};
inline cudaError_t dev_alloc(void **p, size_t n, cudaStream_t s)
{
    if (!tl_cache) return cudaErrorInvalidResourceHandle;   // an entry point forgot its CacheScope
    return tl_cache->alloc(p, n, s);
}
inline void dev_free(void *p, cudaStream_t s)
{
    if (tl_cache) tl_cache->free(p, s);
}

// stream-ordered device buffer
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t st = nullptr;
    DevCache *cache = nullptr;   // where p came from (a buffer may outlive the CacheScope it was allocated under)
    DevBuf() = default;
    DevBuf(size_t n, cudaStream_t s) { alloc(n, s); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes), st(o.st), cache(o.cache) { o.p = nullptr; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        release();
        p = o.p; bytes = o.bytes; st = o.st; cache = o.cache; o.p = nullptr;
        return *this;
    }
    void alloc(size_t n, cudaStream_t s)
    {
        release();
        st = s;
        bytes = n;
        cache = tl_cache;
        CK(dev_alloc(&p, n ? n : 16, s));
    }
    void release()
    {
        if (p && cache) cache->free(p, st);
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
