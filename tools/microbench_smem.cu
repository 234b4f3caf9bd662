// tools/microbench_smem.cu -- how fast are random-address shared-memory atomics vs plain accesses on B200?
// Justifies the table protocol of rogtk_b200/csrc/kernels_smem.cuh (plain RMW + tag election instead of atomics).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_smem tools/microbench_smem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NT = 256, SLOTS = 8192, ITERS = 2000;

__device__ __forceinline__ uint32_t rnd(uint32_t &s) { s = s * 1664525u + 1013904223u; return (s ^ (s >> 15)) * 0x2C1B3C6Du; }

template <int MODE>
__global__ void __launch_bounds__(NT, 2) k(uint32_t *out)
{
    extern __shared__ __align__(16) uint8_t raw[];
    unsigned long long *k64 = reinterpret_cast<unsigned long long *>(raw);
    uint32_t *v32 = reinterpret_cast<uint32_t *>(k64 + SLOTS);
    for (int i = threadIdx.x; i < SLOTS; i += NT) { k64[i] = ~0ull; v32[i] = 0; }
    __syncthreads();
    uint32_t s = blockIdx.x * NT + threadIdx.x + 1, acc = 0;
    for (int it = 0; it < ITERS; it++) {
        uint32_t h = rnd(s) >> 19;   // 13 bits
        if (MODE == 0) atomicAdd(&v32[h], 1u);                                   // RED-style add
        if (MODE == 1) { uint32_t v = v32[h]; v32[h] = v + 1; }                  // plain RMW (racy, for speed only)
        if (MODE == 2) acc += (uint32_t)atomicCAS(&k64[h], ~0ull, (unsigned long long)h);   // CAS.64
        if (MODE == 3) acc += (uint32_t)k64[h];                                 // LDS.64
        if (MODE == 4) acc += atomicAdd(&v32[h], 1u);                           // add with return
        if (MODE == 5) { acc += (uint32_t)k64[h]; reinterpret_cast<uint8_t *>(&v32[h])[3] = (uint8_t)threadIdx.x;
                         uint32_t v = v32[h]; v32[h] = v + 1; }                 // probe + tag + RMW (the protocol's mix)
        if (MODE == 6) { acc += (uint32_t)k64[h]; atomicAdd(&v32[h], 1u); }      // probe + atomic add
        if (MODE == 7) { unsigned m = __match_any_sync(0xffffffffu, h); acc += m; }
        if (MODE == 8) { s ^= atomicAdd(&v32[h], 1u) & 0xff; }                      // dependent atomicAdd chain (latency)
        if (MODE == 9) { s ^= v32[h] & 0xff; }                                      // dependent LDS chain (latency)
        if (MODE == 10) { s ^= (uint32_t)atomicCAS(&k64[h], ~0ull, (unsigned long long)h) & 0xff; }   // dependent CAS chain
        if (MODE == 11) { uint32_t v = v32[h]; unsigned long long kk = k64[h]; if ((uint32_t)kk != h) atomicAdd(&v32[h], 256u); if (v & 1) atomicOr(&v32[h], 2u); }
    }
    if (acc == 0x12345678u) out[0] = acc + v32[threadIdx.x];
}

template <int MODE>
void run(const char *name, int sms)
{
    uint32_t *d;
    cudaMalloc(&d, 1024);
    size_t sm = SLOTS * 12;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = sms * 2;
    k<MODE><<<grid, NT, sm>>>(d);
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) k<MODE><<<grid, NT, sm>>>(d);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double ops = 5.0 * grid * NT * (double)ITERS;
    printf("%-34s %8.2f Gops/s  %6.3f ops/ns/SM  (%s)\n", name, ops / ms / 1e6, ops / ms / 1e6 / sms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int sms = p.multiProcessorCount;
    run<0>("atomicAdd u32 (no return)", sms);
    run<4>("atomicAdd u32 (return)", sms);
    run<1>("plain LDS+STS RMW u32", sms);
    run<2>("atomicCAS u64", sms);
    run<3>("LDS.64", sms);
    run<5>("LDS.64 + tag STS.8 + RMW", sms);
    run<6>("LDS.64 + atomicAdd", sms);
    run<7>("match_any u32", sms);
    run<8>("dependent atomicAdd chain", sms);
    run<9>("dependent LDS chain", sms);
    run<10>("dependent CAS.64 chain", sms);
    run<11>("LDS.64+LDS.32 then RED add/or", sms);
    return 0;
}
