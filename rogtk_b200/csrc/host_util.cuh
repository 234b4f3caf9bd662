// host_util.cuh -- host-side helpers shared by the orchestration code: error plumbing and a stream-ordered device buffer.
#pragma once
#include <cuda_runtime.h>
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
        CK(cudaMallocAsync(&p, n ? n : 16, s));
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


}  // namespace rogtk
