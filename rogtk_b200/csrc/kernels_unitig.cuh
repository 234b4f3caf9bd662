// kernels_unitig.cuh -- SURVEY 8(f-2): unitig compaction of the pruned node sets (the `Compression` assembly method,
// reference src/fracture.rs:351-383 -> debruijn::compression::compress_graph with SimpleCompress(saturating_add))
// and the plugin's contig selection (min_length filter + longest contig, src/fracture.rs:258-277).
//
// Input: a device-resident batch of the hot path (per group: ascending k-mers, exts, counts).  Per group:
//   fix_exts   : an ext bit survives iff the neighbour k-mer it names is a node of the group;
//   link i -> j: i has exactly one right ext, j is that neighbour, j has exactly one left ext, j != i;
//   unitig     : maximal chain of links; sequence = first k-mer + last base of every further node (K-1 overlap merge);
//                count = saturating u16 sum of the node counts; exts = (left nibble of the head, right nibble of the tail).
// The SET of unitigs of a group does not depend on traversal order.  What does (SURVEY D-3): the order in which the reference
// lists them (boomphf slot order), hence which of several equally long unitigs `max_by_key` returns, and where a pure cycle
// is cut.  Here unitigs are listed by ascending head k-mer, a cycle is cut at its smallest k-mer, and both situations are
// flagged per group so that a host can re-run exactly those groups through the reference code.
#pragma once
#include "common.cuh"
#include "kernels_post.cuh"

namespace rogtk {

constexpr uint8_t UF_HAS_IN = 1, UF_VISITED = 2, UF_HEAD = 4, UF_CYCLIC = 8;

struct UnitigArgs {
    int64_t n, n_groups;
    const int64_t *offsets;   // [n_groups+1]
    const uint64_t *lo, *hi;  // hi == nullptr when K <= 32
    const uint8_t *exts;
    const uint16_t *counts;
    int K;
    uint8_t *fx;      // [n] exts after fix_exts
    int32_t *nxt;     // [n] linked successor, index inside the group, or -1
    uint8_t *flags;   // [n] UF_*
    int32_t *hlen;    // [n] nodes of the unitig headed by i (heads only)
    uint16_t *hcnt;   // [n] its count
    uint8_t *hext;    // [n] its exts
};

__device__ __forceinline__ int64_t group_of(const int64_t *offsets, int64_t n_groups, int64_t i)
{
    int64_t a = 0, b = n_groups;   // last g with offsets[g] <= i
    while (b - a > 1) {
        const int64_t m = (a + b) >> 1;
        if (offsets[m] <= i) a = m; else b = m;
    }
    return a;
}

// fix_exts + the unique successor (if the fixed right nibble has exactly one bit).  Only right neighbours are searched: a left
// ext bit c of node j stems from an observation "c then j", which is also an observation "i = c+j[:-1] then last base of j",
// so j's surviving left bits are exactly {first base of i : i is a node and one of its surviving right exts names j}; they
// are OR-ed into j's byte from i's side (k_unitig_fixleft intersects with j's own left nibble for batches of foreign origin).
// fx is addressed as 32-bit words (four nodes per word) for the atomics and must be zero on entry.
__global__ void k_unitig_links(UnitigArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int64_t g = group_of(a.offsets, a.n_groups, i);
    const int64_t o = a.offsets[g], n = a.offsets[g + 1] - o;
    const int kbits = 2 * a.K;
    const uint64_t xl = a.lo[i], xh = a.hi ? a.hi[i] : 0ull;
    const uint32_t e = a.exts[i];
    const int top = kbits - 2;   // bit position of the first base
    const uint32_t first = (uint32_t)((top >= 64 ? xh >> (top - 64) : xl >> top) & 3ull);
    uint32_t *fxw = reinterpret_cast<uint32_t *>(a.fx);
    uint32_t f = 0;
    int64_t succ = -1;
    for (uint32_t b = 0; b < 4; b++) {
        if (!((e >> (4 + b)) & 1u)) continue;   // right neighbour: drop the first base, append b
        uint64_t yl = (xl << 2) | b, yh = (xh << 2) | (xl >> 62);
        if (kbits < 64) { yl &= (1ull << kbits) - 1ull; yh = 0; }
        else if (kbits == 64) yh = 0;
        else if (kbits < 128) yh &= (1ull << (kbits - 64)) - 1ull;
        const int64_t j = find_in_group(a.lo + o, a.hi ? a.hi + o : nullptr, n, yh, yl);
        if (j >= 0) {
            f |= 16u << b;
            succ = j;
            const int64_t jj = o + j;
            atomicOr(fxw + (jj >> 2), (1u << first) << (8 * (int)(jj & 3)));
        }
    }
    if (f) atomicOr(fxw + (i >> 2), f << (8 * (int)(i & 3)));
    a.nxt[i] = __popc(f >> 4) == 1 ? (int32_t)succ : -1;
}

__global__ void k_unitig_fixleft(UnitigArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t f = a.fx[i];
    a.fx[i] = (uint8_t)((f & 0xf0u) | (f & a.exts[i] & 15u));
}

// keep the link only if the successor has exactly one (fixed) left ext and is not the node itself; mark linked-to nodes
__global__ void k_unitig_next(UnitigArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int32_t s = a.nxt[i];
    if (s < 0) return;
    const int64_t o = a.offsets[group_of(a.offsets, a.n_groups, i)];
    const int64_t j = o + s;
    if (j == i || __popc(a.fx[j] & 15u) != 1) {
        a.nxt[i] = -1;
        return;
    }
    a.flags[j] = UF_HAS_IN;   // one writer: j's only left neighbour
}

__device__ __forceinline__ uint32_t sat16(uint32_t x) { return x > 65535u ? 65535u : x; }

// chains: a node nobody links to heads a unitig; walk it
__global__ void k_unitig_heads(UnitigArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (a.flags[i] & UF_HAS_IN) return;
    const int64_t o = a.offsets[group_of(a.offsets, a.n_groups, i)];
    int64_t cur = i;
    int32_t len = 1;
    uint32_t cnt = a.counts[i];
    while (a.nxt[cur] >= 0) {
        cur = o + a.nxt[cur];
        a.flags[cur] = UF_HAS_IN | UF_VISITED;
        len++;
        cnt = sat16(cnt + a.counts[cur]);
    }
    a.flags[i] = UF_HEAD | UF_VISITED;
    a.hlen[i] = len;
    a.hcnt[i] = (uint16_t)cnt;
    a.hext[i] = (uint8_t)((a.fx[i] & 15u) | (a.fx[cur] & 0xf0u));
}

// what is left are pure cycles (every member is linked to): the member with the smallest k-mer heads the unitig
__global__ void k_unitig_cycles(UnitigArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (a.flags[i] != UF_HAS_IN) return;   // visited by a chain, or a head
    const int64_t g = group_of(a.offsets, a.n_groups, i);
    const int64_t o = a.offsets[g], n = a.offsets[g + 1] - o;
    int64_t cur = i, last = i;
    int32_t len = 0;
    uint32_t cnt = 0;
    bool smallest = true;
    do {
        if (cur < i) {
            smallest = false;
            break;
        }
        len++;
        cnt = sat16(cnt + a.counts[cur]);
        last = cur;
        cur = o + a.nxt[cur];
    } while (cur != i && len <= n);
    if (!smallest) return;
    // (flags of the other members stay UF_HAS_IN; only heads are read later)
    a.hlen[i] = len;
    a.hcnt[i] = (uint16_t)cnt;
    a.hext[i] = (uint8_t)((a.fx[i] & 15u) | (a.fx[last] & 0xf0u));
    a.flags[i] = UF_HAS_IN | UF_HEAD | UF_CYCLIC;
}

__global__ void k_unitig_headflags(const uint8_t *flags, int64_t n, int32_t *is_head)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_head[i] = (flags[i] & UF_HEAD) ? 1 : 0;
}

// per unitig (rank of its head among the heads = ascending head k-mer inside a group): head node, sequence length, count, exts
__global__ void k_unitig_compact(UnitigArgs a, const int64_t *upos, int64_t *uhead, int32_t *ulen, uint16_t *ucnt, uint8_t *uext,
                                 uint8_t *ucyc)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n || !(a.flags[i] & UF_HEAD)) return;
    const int64_t u = upos[i];
    uhead[u] = i;
    ulen[u] = a.K + a.hlen[i] - 1;
    ucnt[u] = a.hcnt[i];
    uext[u] = a.hext[i];
    ucyc[u] = (a.flags[i] & UF_CYCLIC) ? 1 : 0;
}

// ASCII writer that stores aligned 8-byte words
struct SeqWriter {
    uint8_t *dst;
    uint64_t buf = 0;
    int nb = 0;
    __device__ explicit SeqWriter(uint8_t *p) : dst(p) {}
    __device__ __forceinline__ void put(uint32_t code)
    {
        const uint8_t ch = (uint8_t)((0x54474341u >> (8u * code)) & 0xffu);   // "ACGT"
        if ((reinterpret_cast<uintptr_t>(dst) & 7u) && nb == 0) {   // unaligned prologue: single bytes
            *dst++ = ch;
            return;
        }
        buf |= (uint64_t)ch << (8 * nb);
        if (++nb == 8) {
            *reinterpret_cast<uint64_t *>(dst) = buf;
            dst += 8;
            buf = 0;
            nb = 0;
        }
    }
    __device__ __forceinline__ void flush()
    {
        for (int t = 0; t < nb; t++) *dst++ = (uint8_t)(buf >> (8 * t));
        nb = 0;
    }
};

// one thread per unitig: first k-mer in full, then the last base of every further node
__global__ void k_unitig_emit(UnitigArgs a, int64_t n_unitigs, const int64_t *uhead, const int32_t *ulen, const int64_t *seq_off,
                              uint8_t *seq)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unitigs) return;
    const int64_t h = uhead[u];
    const int64_t o = a.offsets[group_of(a.offsets, a.n_groups, h)];
    SeqWriter w(seq + seq_off[u]);
    const uint64_t xl = a.lo[h], xh = a.hi ? a.hi[h] : 0ull;
    for (int p = 0; p < a.K; p++) {
        const int sh = 2 * (a.K - 1 - p);
        w.put((uint32_t)((sh >= 64 ? xh >> (sh - 64) : xl >> sh) & 3ull));
    }
    int64_t cur = h;
    for (int32_t t = a.K; t < ulen[u]; t++) {
        cur = o + a.nxt[cur];
        w.put((uint32_t)(a.lo[cur] & 3ull));
    }
    w.flush();
}

// the plugin's selection: contigs of at least min_length bases, the longest one (src/fracture.rs:258-277).  Ties in length are
// flagged (the reference's pick depends on its own listing order); here the tied unitig with the smallest head k-mer is named.
__global__ void k_unitig_largest(int64_t n_groups, const int64_t *uoff, const int32_t *ulen, const uint8_t *ucyc, int64_t min_length,
                                 int64_t *largest, uint8_t *lflags)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int64_t best = -1;
    int32_t best_len = -1;
    uint8_t tie = 0;
    for (int64_t u = uoff[g]; u < uoff[g + 1]; u++) {
        const int32_t l = ulen[u];
        if ((int64_t)l < min_length) continue;
        if (l > best_len) {
            best_len = l;
            best = u;
            tie = 0;
        } else if (l == best_len) {
            tie = 1;
        }
    }
    largest[g] = best;
    lflags[g] = best < 0 ? 0 : (uint8_t)(tie | (ucyc[best] ? 2 : 0));
}

}  // namespace rogtk
