// common.cuh -- shared device helpers for the rogtk_b200 k-mer engine (sm_100a).
//
// Domain recap (reference src/fracture.rs:105-158 + SURVEY.md Appendix A): a k-mer of K bases is the base-4
// big-endian number of its bases (A=0,C=1,G=2,T=3); Exts is one byte, low nibble = left neighbour bits, high
// nibble = right neighbour bits; counts are saturating u16.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rogtk {

// ------------------------------------------------------------------------------------------------------------
// string column access (Arrow Utf8 / LargeUtf8 after view-gather), reference src/expressions.rs:734-740
// ------------------------------------------------------------------------------------------------------------
struct ColView {
    const uint8_t *data;
    const void *offsets;      // int32 or int64, n_rows+1
    const uint8_t *validity;  // bitmap or nullptr
    int64_t validity_bit_offset;
    int64_t n_rows;
    int64_t row_begin;  // first row of this view (offsets / validity are indexed with absolute rows)
    int off64;  // 1: int64 offsets
};

__device__ __forceinline__ int64_t col_off(const ColView &c, int64_t i)
{
    return c.off64 ? reinterpret_cast<const int64_t *>(c.offsets)[i]
                   : (int64_t) reinterpret_cast<const int32_t *>(c.offsets)[i];
}
__device__ __forceinline__ bool col_valid(const ColView &c, int64_t i)
{
    if (!c.validity) return true;
    int64_t b = i + c.validity_bit_offset;
    return (c.validity[b >> 3] >> (b & 7)) & 1;
}

// ------------------------------------------------------------------------------------------------------------
// base coding.  code = ((c>>1)^(c>>2))&3 maps A/a->0 C/c->1 G/g->2 T/t->3; validity is checked
// separately because every other byte also lands on some code.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t base_code(uint32_t c) { return ((c >> 1) ^ (c >> 2)) & 3u; }
__device__ __forceinline__ bool base_ok(uint32_t c)
{
    uint32_t u = c & 0xDFu;  // to_uppercase for ASCII letters (src/fracture.rs:221)
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

// 4 ASCII bytes (little-endian word, byte 0 = first base) -> 8 packed bits, first base in the top 2 bits.
__device__ __forceinline__ uint32_t pack4(uint32_t w)
{
    uint32_t x = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    return (x * 0x40100401u) >> 24;
}
// all four bytes in {A,C,G,T,a,c,g,t}?
__device__ __forceinline__ bool valid4(uint32_t w)
{
    uint32_t u = w & 0xDFDFDFDFu;
    uint32_t x = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    // expected upper-case letter for each code via byte permute of "ACGT"
    uint32_t t = x | (x >> 12);  // nibble0=c0 nibble1=c2 nibble2=c1 nibble3=c3
    uint32_t e = __byte_perm(0x54474341u, 0u, t & 0x3333u);      // bytes: L[c0] L[c2] L[c1] L[c3]
    uint32_t v = __byte_perm(u, 0u, 0x3120u);                     // bytes: u0 u2 u1 u3
    return e == v;
}

// ------------------------------------------------------------------------------------------------------------
// k-mer keys.  Key<1> = one 64-bit word (K <= 32), Key<2> = two words (K <= 64; w[1] = high word).
// ------------------------------------------------------------------------------------------------------------
template <int W>
struct Key;

template <>
struct Key<1> {
    uint64_t lo;
    __device__ __forceinline__ static Key zero() { return Key{0ull}; }
    __device__ __forceinline__ void push(uint32_t b, int kbits)
    {
        lo = (lo << 2) | b;
        if (kbits < 64) lo &= (1ull << kbits) - 1ull;
    }
    __device__ __forceinline__ uint32_t top_base(int kbits) const { return (uint32_t)(lo >> (kbits - 2)) & 3u; }
    __device__ __forceinline__ bool operator==(const Key &o) const { return lo == o.lo; }
    __device__ __forceinline__ bool operator<(const Key &o) const { return lo < o.lo; }
    __device__ __forceinline__ uint64_t hash() const
    {
        uint64_t z = lo;
        z = (z ^ (z >> 32)) * 0xD6E8FEB86659FD93ull;
        z = (z ^ (z >> 32)) * 0xD6E8FEB86659FD93ull;
        return z ^ (z >> 32);
    }
    __device__ __forceinline__ uint64_t hi_word() const { return 0ull; }
};

template <>
struct Key<2> {
    uint64_t lo, hi;
    __device__ __forceinline__ static Key zero() { return Key{0ull, 0ull}; }
    __device__ __forceinline__ void push(uint32_t b, int kbits)
    {
        hi = (hi << 2) | (lo >> 62);
        lo = (lo << 2) | b;
        if (kbits < 128) hi &= (1ull << (kbits - 64)) - 1ull;  // kbits in (64,128)
    }
    __device__ __forceinline__ uint32_t top_base(int kbits) const { return (uint32_t)(hi >> (kbits - 66)) & 3u; }
    __device__ __forceinline__ bool operator==(const Key &o) const { return lo == o.lo && hi == o.hi; }
    __device__ __forceinline__ bool operator<(const Key &o) const { return hi < o.hi || (hi == o.hi && lo < o.lo); }
    __device__ __forceinline__ uint64_t hash() const
    {
        uint64_t z = lo ^ (hi * 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 32)) * 0xD6E8FEB86659FD93ull;
        z ^= hi;
        z = (z ^ (z >> 32)) * 0xD6E8FEB86659FD93ull;
        return z ^ (z >> 32);
    }
    __device__ __forceinline__ uint64_t hi_word() const { return hi; }
};

// SplitMix64 (bench/test data generator and UMI sharding)
__host__ __device__ __forceinline__ uint64_t splitmix_mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t splitmix_out(uint64_t seed, uint64_t c)
{
    return splitmix_mix(seed + (c + 1ull) * 0x9E3779B97F4A7C15ull);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

}  // namespace rogtk
