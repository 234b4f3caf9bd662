// scan.cuh -- exclusive prefix sum over int64 (group sizes -> offsets).  Three small kernels; the arrays scanned
// here are per-group (<= a few million entries), far off the hot path.
#pragma once
#include "common.cuh"

namespace rogtk {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;                        // per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;   // 2048 per block

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total, int64_t *smem /* >= 33 */)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (unsigned)d) x += y;
    }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int64_t s = lane < (blockDim.x >> 5) ? smem[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (unsigned)d) s += y;
        }
        smem[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    int64_t warp_base = warp ? smem[warp - 1] : 0;
    *total = smem[(blockDim.x >> 5) - 1];
    __syncthreads();
    return warp_base + x - v;
}

template <typename InT>
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block_sums(const InT *in, int64_t n, int64_t *block_sums)
{
    __shared__ int64_t sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) s += (int64_t)in[base + i];
    int64_t tot;
    block_exclusive_scan(s, &tot, sm);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(int64_t *block_sums, int64_t nb)
{
    __shared__ int64_t sm[33];
    int64_t carry = 0;
    for (int64_t b0 = 0; b0 < nb; b0 += SCAN_BLOCK) {
        int64_t i = b0 + threadIdx.x;
        int64_t v = i < nb ? block_sums[i] : 0;
        int64_t tot;
        int64_t ex = block_exclusive_scan(v, &tot, sm);
        if (i < nb) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) block_sums[nb] = carry;
}

template <typename InT>
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_final(const InT *in, int64_t n, const int64_t *block_sums, int64_t nb, int64_t *out /* n+1 */)
{
    __shared__ int64_t sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? (int64_t)in[base + i] : 0;
        s += v[i];
    }
    int64_t tot;
    int64_t ex = block_exclusive_scan(s, &tot, sm) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_sums[nb];
}

// host launcher: out[0..n] = exclusive scan of in[0..n) ; scratch must hold (n/SCAN_TILE + 2) int64
template <typename InT>
inline void exclusive_scan(const InT *in, int64_t n, int64_t *out, int64_t *scratch, cudaStream_t st, int64_t *launches)
{
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 0) nb = 1;
    k_scan_block_sums<InT><<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(in, n, scratch);
    k_scan_sums<<<1, SCAN_BLOCK, 0, st>>>(scratch, nb);
    k_scan_final<InT><<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(in, n, scratch, nb, out);
    if (launches) *launches += 3;
}

inline size_t scan_scratch_bytes(int64_t n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE + 2) * sizeof(int64_t); }

}  // namespace rogtk
