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
    for (uint32_t m = e >> 4; m; m &= m - 1u) {   // right neighbours (t-th set bit per iteration, so a warp's lanes search together)
        const uint32_t b = (uint32_t)__ffs((int)m) - 1u;   // drop the first base, append b
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


// ---------------------------------------------------------------------------------------------------------------------------------
// Fused per-group variant: one CTA holds a whole group (<= NMAX nodes) in shared memory -- neighbour searches, link building,
// list ranking by pointer jumping (ceil(log2 n) rounds instead of serial chain walks), per-unitig sums, and the ASCII
// sequences staged in shared memory and written out as aligned words.  Two passes of the same body: EMIT = false counts the
// unitigs and contig bases of every group (-> two scans give every group its slice of the outputs), EMIT = true writes them.
// Used when every group of the batch fits (the per-UMI workloads); otherwise the node-parallel kernels above run.
// ---------------------------------------------------------------------------------------------------------------------------------
struct UnitigGroupArgs {
    int64_t n_groups;
    const int64_t *offsets;
    const uint64_t *lo, *hi;
    const uint8_t *exts;
    const uint16_t *counts;
    int K;
    int n_min;              // this launch takes the groups with n_min < nodes <= NMAX (and the empty ones when n_min == 0)
    // EMIT = false
    int32_t *g_nu;          // [n_groups] unitigs of the group
    int64_t *g_nb;          // [n_groups] contig bases of the group
    // EMIT = true
    const int64_t *uoff;    // [n_groups+1] exclusive scan of g_nu
    const int64_t *sbase;   // [n_groups+1] exclusive scan of g_nb
    int64_t *seq_offsets;   // [n_unitigs]
    uint8_t *seq;
    int32_t *ulen;
    uint16_t *ucnt;
    uint8_t *uext, *ucyc;
};

template <int W, int NMAX>
constexpr size_t unitig_group_smem()
{
    return (size_t)NMAX * (8 * W + 4 + 4 + 4 + 1 + 2 + 2 + 2 + 1 + 2) + 64 * 4;
}

__global__ void k_max_group(const int64_t *offsets, int64_t n_groups, unsigned long long *out)
{
    unsigned long long m = 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long d = (unsigned long long)(offsets[g + 1] - offsets[g]);
        m = d > m ? d : m;
    }
    for (int d = 16; d; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, d);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// exclusive block scan of two values per thread (NT threads); ws: 2 * 32 words
template <int NT>
__device__ __forceinline__ void block_scan_pair(uint32_t a, uint32_t b, uint32_t &ea, uint32_t &eb, uint32_t &ta, uint32_t &tb, uint32_t *ws)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t xa = a, xb = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ya = __shfl_up_sync(0xffffffffu, xa, d), yb = __shfl_up_sync(0xffffffffu, xb, d);
        if (lane >= (unsigned)d) { xa += ya; xb += yb; }
    }
    __syncthreads();   // ws may still be read from a previous call
    if (lane == 31) { ws[warp] = xa; ws[32 + warp] = xb; }
    __syncthreads();
    uint32_t ba = 0, bb = 0;
    ta = 0; tb = 0;
    for (int w = 0; w < NT / 32; w++) {
        const uint32_t sa = ws[w], sb = ws[32 + w];
        if (w < (int)warp) { ba += sa; bb += sb; }
        ta += sa; tb += sb;
    }
    ea = ba + xa - a;
    eb = bb + xb - b;
}

template <int W, int NMAX, int NT, bool EMIT>
__global__ void __launch_bounds__(NT) k_unitig_group(UnitigGroupArgs a)
{
    extern __shared__ __align__(16) uint8_t ug_raw[];
    uint64_t *s_lo = reinterpret_cast<uint64_t *>(ug_raw);
    uint64_t *s_hi = s_lo + NMAX;                                                   // only touched when W == 2
    uint32_t *s_pr0 = reinterpret_cast<uint32_t *>(s_lo + (W == 2 ? 2 : 1) * NMAX);   // (head candidate | distance << 16), double buffered
    uint32_t *s_pr1 = s_pr0 + NMAX;
    uint32_t *s_sum = s_pr1 + NMAX;                                                 // per head: count sum, later its start in the staging
    uint32_t *s_fxw = s_sum + NMAX;                                                 // fixed exts, one byte per node (NMAX / 4 words)
    uint16_t *s_cnt = reinterpret_cast<uint16_t *>(s_fxw + NMAX / 4);
    int16_t *s_nxt = reinterpret_cast<int16_t *>(s_cnt + NMAX);
    int16_t *s_pred = s_nxt + NMAX;                                                 // linking predecessor; later: nodes of the unitig (heads)
    uint8_t *s_ex = reinterpret_cast<uint8_t *>(s_pred + NMAX);                     // exts; later flags: 1 head, 2 cyclic
    uint8_t *s_stage = s_ex + NMAX;                                                 // 2 * NMAX bytes of contig text
    uint32_t *s_ws = reinterpret_cast<uint32_t *>(s_stage + 2 * NMAX);
    uint8_t *s_fx = reinterpret_cast<uint8_t *>(s_fxw);
    const int tid = threadIdx.x;
    const int K = a.K, kbits = 2 * a.K;

    auto find = [&](int n, uint64_t khi, uint64_t klo) -> int {
        int lo_ = 0, hi_ = n;
        while (lo_ < hi_) {
            const int m = (lo_ + hi_) >> 1;
            const uint64_t mh = W == 2 ? s_hi[m] : 0ull, ml = s_lo[m];
            if (mh == khi && ml == klo) return m;
            if (mh < khi || (mh == khi && ml < klo)) lo_ = m + 1; else hi_ = m;
        }
        return -1;
    };

    for (int64_t g = blockIdx.x; g < a.n_groups; g += gridDim.x) {
        const int64_t o = a.offsets[g];
        const int n = (int)(a.offsets[g + 1] - o);
        if (n > NMAX || n <= a.n_min) {   // another launch's group (the host excludes batches with n > the largest NMAX)
            if (n == 0 && a.n_min == 0 && !EMIT && tid == 0) {
                a.g_nu[g] = 0;
                a.g_nb[g] = 0;
            }
            continue;
        }
        __syncthreads();   // shared arrays of the previous group are no longer read
        // ---- 1. stage the group ---------------------------------------------------------------------------------------
        for (int i = tid; i < n; i += NT) {
            s_lo[i] = a.lo[o + i];
            if (W == 2) s_hi[i] = a.hi[o + i];
            s_ex[i] = a.exts[o + i];
            s_cnt[i] = a.counts[o + i];
            s_pred[i] = -1;
            s_sum[i] = 0;
        }
        for (int i = tid; i < (n + 3) / 4; i += NT) s_fxw[i] = 0;
        __syncthreads();
        // ---- 2. fix_exts (right side searched, left side OR-ed in from the linking node) + unique successor ---------------------
        for (int i = tid; i < n; i += NT) {
            const uint64_t xl = s_lo[i], xh = W == 2 ? s_hi[i] : 0ull;
            const uint32_t e = s_ex[i];
            const int top = kbits - 2;
            const uint32_t first = (uint32_t)((top >= 64 ? xh >> (top - 64) : xl >> top) & 3ull);
            uint32_t f = 0;
            int succ = -1;
            for (uint32_t m = e >> 4; m; m &= m - 1u) {   // t-th set bit per iteration: the lanes of a warp search together
                const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                uint64_t yl = (xl << 2) | b, yh = (xh << 2) | (xl >> 62);
                if (kbits < 64) { yl &= (1ull << kbits) - 1ull; yh = 0; }
                else if (kbits == 64) yh = 0;
                else if (kbits < 128) yh &= (1ull << (kbits - 64)) - 1ull;
                const int j = find(n, yh, yl);
                if (j >= 0) {
                    f |= 16u << b;
                    succ = j;
                    atomicOr(&s_fxw[j >> 2], (1u << first) << (8 * (j & 3)));
                }
            }
            if (f) atomicOr(&s_fxw[i >> 2], f << (8 * (i & 3)));
            s_nxt[i] = __popc(f >> 4) == 1 ? (int16_t)succ : (int16_t)-1;
        }
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const uint32_t f = s_fx[i];
            s_fx[i] = (uint8_t)((f & 0xf0u) | (f & s_ex[i] & 15u));
        }
        __syncthreads();
        // ---- 3. links: the successor must have exactly one left ext and differ from the node ------------------------------------------
        for (int i = tid; i < n; i += NT) {
            const int j = s_nxt[i];
            if (j < 0) continue;
            if (j == i || __popc(s_fx[j] & 15u) != 1) s_nxt[i] = -1;
            else s_pred[j] = (int16_t)i;   // one writer: j's only left neighbour
        }
        __syncthreads();
        // ---- 4. list ranking by pointer jumping: (node reached, distance to it); heads are fixed points -----------------------------
        for (int i = tid; i < n; i += NT) {
            const int p = s_pred[i];
            s_pr0[i] = p >= 0 ? ((uint32_t)p | (1u << 16)) : (uint32_t)i;
            s_ex[i] = p >= 0 ? 0 : 1;   // flags from here on: bit 0 head, bit 1 cyclic
        }
        __syncthreads();
        uint32_t *pin = s_pr0, *pout = s_pr1;
        for (int span = 1; span < n; span <<= 1) {
            for (int i = tid; i < n; i += NT) {
                const uint32_t v = pin[i], w = pin[v & 0xffffu];
                pout[i] = (w & 0xffffu) | ((((v >> 16) + (w >> 16)) & 0xffffu) << 16);
            }
            __syncthreads();
            uint32_t *t = pin; pin = pout; pout = t;
        }
        // ---- 5. what does not end at a head is a pure cycle: its smallest member becomes the head and ranks the others ----------------
        for (int i = tid; i < n; i += NT) {
            if (s_ex[pin[i] & 0xffffu] & 1u) continue;   // chain node (or head)
            s_ex[i] |= 4u;                               // cycle member (own byte; bit 0 of other bytes is stable here)
            int cur = i, len = 0;
            bool smallest = true;
            do {
                if (cur < i) { smallest = false; break; }
                len++;
                cur = s_nxt[cur];
            } while (cur != i && len <= n);
            if (!smallest) continue;
            cur = i;
            for (int r = 0; r < len; r++) {   // (only this thread writes the members' entries; nobody reads them before the barrier)
                pout[cur] = (uint32_t)i | ((uint32_t)r << 16);
                cur = s_nxt[cur];
            }
            pout[i] |= 0x80000000u;   // marks the cycle head for the merge below
        }
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            if (!(s_ex[i] & 4u)) continue;
            const uint32_t v = pout[i];
            pin[i] = v & 0x7fffffffu;
            if (v & 0x80000000u) s_ex[i] = 7;   // head of a cycle
        }
        __syncthreads();
        // ---- 6. per unitig: nodes, saturating count sum, exts of the tail ------------------------------------------------------------------
        for (int i = tid; i < n; i += NT) pout[i] = 0;   // per head: number of nodes
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const uint32_t v = pin[i];
            const int h = (int)(v & 0xffffu), r = (int)(v >> 16);
            atomicAdd(&s_sum[h], (uint32_t)s_cnt[i]);
            atomicMax(&pout[h], (uint32_t)r + 1u);
            const int nx = s_nxt[i];
            if (nx < 0 || nx == h) s_pred[h] = (int16_t)(s_fx[i] & 0xf0u);   // the tail (one per unitig) leaves its right nibble
        }
        __syncthreads();
        // ---- 7. order of the unitigs = order of their heads; every thread owns a contiguous range of nodes -------------------------------
        const int per = (n + NT - 1) / NT;
        const int c0 = tid * per < n ? tid * per : n, c1 = c0 + per < n ? c0 + per : n;
        uint32_t my_heads = 0, my_bases = 0;
        for (int i = c0; i < c1; i++)
            if (s_ex[i] & 1u) {
                my_heads++;
                my_bases += (uint32_t)K + pout[i] - 1u;
            }
        uint32_t eh, eb, th, tb;
        block_scan_pair<NT>(my_heads, my_bases, eh, eb, th, tb, s_ws);
        if (!EMIT) {
            if (tid == 0) {
                a.g_nu[g] = (int32_t)th;
                a.g_nb[g] = (int64_t)tb;
            }
            continue;
        }
        const int64_t u0 = a.uoff[g], sb = a.sbase[g];
        const bool staged = tb <= 2u * (uint32_t)NMAX;
        {
            uint32_t hb = eh, bb = eb;
            for (int i = c0; i < c1; i++) {
                if (!(s_ex[i] & 1u)) continue;
                const uint32_t len = (uint32_t)K + pout[i] - 1u;
                const int64_t u = u0 + hb;
                a.seq_offsets[u] = sb + bb;
                a.ulen[u] = (int32_t)len;
                a.ucnt[u] = (uint16_t)(s_sum[i] > 65535u ? 65535u : s_sum[i]);
                a.uext[u] = (uint8_t)((s_fx[i] & 15u) | ((uint32_t)s_pred[i] & 0xf0u));
                a.ucyc[u] = (s_ex[i] & 2u) ? 1 : 0;
                s_sum[i] = bb;   // start of the unitig's text inside the group's slice
                hb++;
                bb += len;
            }
        }
        __syncthreads();
        // ---- 8. text: the head writes its k-mer, every other node its last base; staged, then stored as aligned words -------------------
        uint8_t *text = staged ? s_stage : a.seq + sb;
        for (int i = tid; i < n; i += NT) {
            const uint32_t v = pin[i];
            const int h = (int)(v & 0xffffu), r = (int)(v >> 16);
            const uint32_t start = s_sum[h];
            if (r == 0) {
                const uint64_t xl = s_lo[i], xh = W == 2 ? s_hi[i] : 0ull;
                for (int p = 0; p < K; p++) {
                    const int sh = 2 * (K - 1 - p);
                    const uint32_t code = (uint32_t)((sh >= 64 ? xh >> (sh - 64) : xl >> sh) & 3ull);
                    text[start + p] = (uint8_t)((0x54474341u >> (8u * code)) & 0xffu);
                }
            } else {
                text[start + K - 1 + r] = (uint8_t)((0x54474341u >> (8u * (uint32_t)(s_lo[i] & 3ull))) & 0xffu);
            }
        }
        if (staged) {
            __syncthreads();
            uint8_t *dst = a.seq + sb;
            const uint32_t head = (uint32_t)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u);   // bytes before the first aligned word
            const uint32_t hb = head < tb ? head : tb;
            if ((uint32_t)tid < hb) dst[tid] = s_stage[tid];
            const uint32_t nw = (tb - hb) / 4;
            uint32_t *dw = reinterpret_cast<uint32_t *>(dst + hb);
            for (uint32_t w = tid; w < nw; w += NT) {
                const uint8_t *s = s_stage + hb + 4 * w;
                dw[w] = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24);
            }
            const uint32_t done = hb + 4 * nw;
            if ((uint32_t)tid < tb - done) dst[done + tid] = s_stage[done + tid];
        }
    }
}

// ---- device side: one string per group out of a device-resident contig batch ---------------------------------------------------
// only_largest: the unitig `largest[g]` (already filtered by min_length); otherwise every unitig of at least min_length bases,
// joined with '\n' in listing order
__global__ void k_contig_row_len(int64_t ng, const int64_t *uoff, const int64_t *soff, const int64_t *largest, int64_t min_length,
                                 int only_largest, int64_t *row_len, uint8_t *row_many)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ng) return;
    int64_t len = 0, cnt = 0;
    if (only_largest) {
        const int64_t u = largest[g];
        if (u >= 0) len = soff[u + 1] - soff[u];
    } else {
        for (int64_t u = uoff[g]; u < uoff[g + 1]; u++) {
            const int64_t l = soff[u + 1] - soff[u];
            if (l >= min_length) {
                len += l + (cnt ? 1 : 0);
                cnt++;
            }
        }
    }
    row_len[g] = len;
    if (row_many) row_many[g] = cnt > 1 ? 8 : 0;
}
// one warp per group
__global__ void k_contig_row_fill(int64_t ng, const int64_t *uoff, const int64_t *soff, const uint8_t *seq, const int64_t *largest,
                                  int64_t min_length, int only_largest, const int64_t *row_off, uint8_t *out)
{
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= ng) return;
    const unsigned lane = threadIdx.x & 31u;
    int64_t o = row_off[g];
    if (only_largest) {
        const int64_t u = largest[g];
        if (u < 0) return;
        const int64_t s = soff[u], l = soff[u + 1] - s;
        for (int64_t j = lane; j < l; j += 32) out[o + j] = seq[s + j];
        return;
    }
    bool first = true;
    for (int64_t u = uoff[g]; u < uoff[g + 1]; u++) {
        const int64_t s = soff[u], l = soff[u + 1] - s;
        if (l < min_length) continue;
        if (!first) {
            if (lane == 0) out[o] = (uint8_t)'\n';
            o++;
        }
        first = false;
        for (int64_t j = lane; j < l; j += 32) out[o + j] = seq[s + j];
        o += l;
    }
}


}  // namespace rogtk
