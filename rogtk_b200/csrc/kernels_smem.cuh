// kernels_smem.cuh -- the hot kernel: one CTA builds one UMI group's pruned de Bruijn node set entirely in shared
// memory (ingest -> 2-bit pack -> rolling k-mers -> open-addressing table -> min-coverage compaction -> bucket sort
// -> emit).  Replaces debruijn::filter::filter_kmers + CountFilter as called at reference src/fracture.rs:110-116.
//
// Table protocol (no shared-memory atomics on the common path): the CTA advances in rounds, every thread carrying at
// most one observation (k-mer, exts) per round.
//   P  probe (read only): first slot on the linear-probe path that holds the key or is EMPTY
//   -- barrier A --
//   C  claim: threads that stopped on EMPTY store their key (plain store, last writer wins); every thread stores its
//      thread id into the tag byte of the slot's value word
//   -- barrier B --
//   V  verify: key differs  -> lost the slot, continue probing at the next slot next round
//              key == mine, tag == mine -> sole elected owner this round: plain read-modify-write of count/exts
//              key == mine, tag != mine -> a duplicate of this k-mer (or a stray tag) won: the observation is applied
//                                          with a shared-memory atomic in the next round's C phase
// Keys never change once verified, slots never return to EMPTY, so lookups stay consistent; counts are exact because
// every observation is applied exactly once, either by the elected plain RMW (phase V) or atomically (phase C), and
// the two phases are separated by barriers.
#pragma once
#include <vector>

#include "common.cuh"
#include "host_util.cuh"
#include "kernels_table.cuh"

namespace rogtk {

constexpr int SM_NT = 256;            // threads per CTA == max rows per group on this path == tag range (8 bits)
constexpr int SM_MAX_BASES = 16384;   // bases staged per group (including <=15 bytes of alignment slack)
constexpr int SM_PK_WORDS = SM_MAX_BASES / 16 + 4;
constexpr int SM_TBITS = 13;          // 8192 slots
constexpr int SM_TSIZE = 1 << SM_TBITS;
constexpr int SM_LIMIT = SM_TSIZE * 3 / 4 - 8;   // max distinct k-mers per group on this path (sort scratch lives above them)
constexpr int SM_SORT_BUCKETS = 2048;
constexpr int SM_SORT_BITS = 11;

template <int W>
constexpr size_t smem_bytes()
{
    return (size_t)SM_TSIZE * sizeof(typename KeySlot<W>::T) + (size_t)SM_TSIZE * 4 + (size_t)SM_PK_WORDS * 4 +
           (size_t)(SM_NT + 1) * 4 * 2 + 40 * 4 + SM_NT + 64;
}

struct SmemArgs {
    ColView col;
    const uint8_t *data_end;         // one past the last readable byte of col.data
    const int64_t *group_offsets;
    int64_t n_groups;
    int K;
    uint32_t min_cov;
    unsigned long long *ticket;      // dynamic group scheduler
    unsigned long long *n_obs;
    int64_t *nsurv;                  // [n_groups]
    int64_t *pool_start;             // [n_groups], -1 = not produced here
    unsigned long long *pool_cursor;
    int64_t pool_cap;
    uint64_t *pool_lo, *pool_hi;
    uint8_t *pool_exts;
    uint16_t *pool_counts;
    int32_t *deferred_ids;           // groups handed to the HBM-table path
    unsigned long long *n_deferred;
};

__device__ __forceinline__ uint32_t hash32(const Key<1> &k)
{
    uint32_t h = (uint32_t)k.lo * 0x9E3779B1u + (uint32_t)(k.lo >> 32) * 0x85EBCA77u;
    h ^= h >> 15;
    return h * 0x2C1B3C6Du;
}
__device__ __forceinline__ uint32_t hash32(const Key<2> &k)
{
    uint32_t h = (uint32_t)k.lo * 0x9E3779B1u + (uint32_t)(k.lo >> 32) * 0x85EBCA77u;
    h ^= h >> 15;
    h = h * 0x2C1B3C6Du + (uint32_t)k.hi * 0xC2B2AE3Du + (uint32_t)(k.hi >> 32) * 0x27D4EB2Fu;
    h ^= h >> 16;
    return h * 0x165667B1u;
}

// 32 bases starting at base index q of a packed stream (16 bases per u32 word, first base in the top bits),
// returned top-aligned.  Reads words q>>4 .. (q>>4)+2.
__device__ __forceinline__ uint64_t window64(const uint32_t *pk, uint32_t q)
{
    const uint32_t w = q >> 4, s = 2u * (q & 15u);
    const uint32_t x0 = pk[w], x1 = pk[w + 1], x2 = pk[w + 2];
    const uint32_t r0 = __funnelshift_l(x1, x0, s), r1 = __funnelshift_l(x2, x1, s);
    return ((uint64_t)r0 << 32) | r1;
}
__device__ __forceinline__ uint32_t base_of(const uint32_t *pk, uint32_t q) { return (pk[q >> 4] >> (30u - 2u * (q & 15u))) & 3u; }

__device__ __forceinline__ void load_window(const uint32_t *pk, uint32_t q, int K, Key<1> &key)
{
    key.lo = window64(pk, q) >> (64 - 2 * K);
}
__device__ __forceinline__ void load_window(const uint32_t *pk, uint32_t q, int K, Key<2> &key)
{
    const uint64_t xh = window64(pk, q), xl = window64(pk, q + 32);
    const int sh = 128 - 2 * K;   // 0..62
    key.hi = xh >> sh;
    key.lo = sh ? ((xl >> sh) | (xh << (64 - sh))) : xl;
}

template <int W>
__device__ __forceinline__ uint32_t sort_bucket(const Key<W> &k, int kbits);
template <>
__device__ __forceinline__ uint32_t sort_bucket<1>(const Key<1> &k, int kbits)
{
    return kbits > SM_SORT_BITS ? (uint32_t)(k.lo >> (kbits - SM_SORT_BITS)) : (uint32_t)k.lo;
}
template <>
__device__ __forceinline__ uint32_t sort_bucket<2>(const Key<2> &k, int kbits)
{
    const int sh = kbits - SM_SORT_BITS;   // kbits > 64 here
    return sh >= 64 ? (uint32_t)(k.hi >> (sh - 64)) : (uint32_t)((k.hi << (64 - sh)) | (k.lo >> sh));
}

// saturating-u16 count increment + exts OR on a value word, keeping the tag byte
__device__ __forceinline__ uint32_t val_apply(uint32_t v, uint32_t exts)
{
    uint32_t c = v & 0xffffu;
    c += (c != 0xffffu);
    return (v & 0xffff0000u) | c | (exts << 16);
}
__device__ __forceinline__ void val_apply_atomic(uint32_t *vp, uint32_t exts)
{
    uint32_t old = *reinterpret_cast<volatile uint32_t *>(vp);
    if ((old & 0xffffu) < 0xfe00u) {   // < 256 concurrent adders cannot carry out of the count field
        atomicAdd(vp, 1u);
        if ((exts << 16) & ~old) atomicOr(vp, exts << 16);
        return;
    }
    while (true) {
        uint32_t nw = val_apply(old, exts);
        uint32_t prev = atomicCAS(vp, old, nw);
        if (prev == old) return;
        old = prev;
    }
}

// block-wide exclusive scan of one u32 per thread (SM_NT threads); ws = 9 words of shared scratch
__device__ __forceinline__ uint32_t block_scan_u32(uint32_t v, uint32_t *total, uint32_t *ws)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (unsigned)d) x += y;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SM_NT / 32; w++) {
        uint32_t s = ws[w];
        if (w < (int)warp) base += s;
        tot += s;
    }
    *total = tot;
    __syncthreads();
    return base + x - v;
}

template <int W>
__global__ void __launch_bounds__(SM_NT, (W == 1 ? 2 : 1)) k_smem_count(SmemArgs a)
{
    using S = KeySlot<W>;
    using KT = typename S::T;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    KT *tkey = reinterpret_cast<KT *>(smem_raw);
    uint32_t *tval = reinterpret_cast<uint32_t *>(tkey + SM_TSIZE);
    uint32_t *pk = tval + SM_TSIZE;
    uint32_t *rstart = pk + SM_PK_WORDS;
    uint32_t *cum = rstart + SM_NT + 1;
    uint32_t *ws = cum + SM_NT + 1;                      // 40 words of scratch: [0..8] scan, [16..] misc
    uint8_t *bad = reinterpret_cast<uint8_t *>(ws + 40);
    volatile uint32_t *misc = ws + 16;                   // [0],[1] ticket lo/hi  [2] special val  [3] abort flag  [4],[5] pool base
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u, warp = tid >> 5;
    const int K = a.K, kbits = 2 * K;
    const uint32_t thr = a.min_cov > 1u ? a.min_cov : 1u;
    unsigned long long cta_obs = 0;

    while (true) {
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = atomicAdd(a.ticket, 1ull);
            misc[0] = (uint32_t)t;
            misc[1] = (uint32_t)(t >> 32);
            misc[2] = 0;
            misc[3] = 0;
        }
        __syncthreads();
        const int64_t g = (int64_t)(((unsigned long long)misc[1] << 32) | misc[0]);
        if (g >= a.n_groups) break;
        const int64_t r0 = a.group_offsets[g], r1 = a.group_offsets[g + 1];
        const int64_t nrows = r1 - r0;
        if (nrows <= 0) continue;
        const int64_t b0 = col_off(a.col, r0), b1 = col_off(a.col, r1);
        if (b1 <= b0) continue;
        const uintptr_t addr0 = reinterpret_cast<uintptr_t>(a.col.data + b0);
        const uint32_t slack = (uint32_t)(addr0 & 15u);
        const int64_t span64 = (b1 - b0) + slack;
        if (nrows > SM_NT || span64 > SM_MAX_BASES) {   // too many rows / bases for this path
            if (tid == 0) a.deferred_ids[atomicAdd(a.n_deferred, 1ull)] = (int32_t)g;
            continue;
        }
        const uint32_t span = (uint32_t)span64;
        const int nr = (int)nrows;

        // ---- per-read base offsets inside the staged span, null rows ----------------------------------------------
        for (int r = tid; r <= nr; r += SM_NT) rstart[r] = (uint32_t)(col_off(a.col, r0 + r) - b0) + slack;
        if (tid < nr) bad[tid] = col_valid(a.col, r0 + tid) ? 0 : 1;
        __syncthreads();

        // ---- ingest: 16-byte chunks -> one packed u32 (16 bases, first base in the top bits) --------------------------
        {
            const uint8_t *a0 = reinterpret_cast<const uint8_t *>(addr0 - slack);
            const int nchunks = (int)((span + 15u) >> 4);
            uint4 v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = tid + i * SM_NT;
                v[i] = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
                if (c < nchunks) {
                    const uint8_t *p = a0 + 16 * c;
                    if (p >= a.col.data && p + 16 <= a.data_end) {
                        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                     : "l"(p));
                    } else {   // chunk straddles the buffer edge: guarded byte loads
                        uint32_t wv[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};
                        for (int t = 0; t < 16; t++) {
                            const uint8_t *pb = p + t;
                            if (pb >= a.col.data && pb < a.data_end)
                                wv[t >> 2] = (wv[t >> 2] & ~(0xffu << (8 * (t & 3)))) | ((uint32_t)*pb << (8 * (t & 3)));
                        }
                        v[i] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = tid + i * SM_NT;
                if (c < nchunks) {
                    pk[c] = (pack4(v[i].x) << 24) | (pack4(v[i].y) << 16) | (pack4(v[i].z) << 8) | pack4(v[i].w);
                    const bool ok = valid4(v[i].x) & valid4(v[i].y) & valid4(v[i].z) & valid4(v[i].w);
                    if (!ok) {   // attribute every bad byte to its read (reference src/fracture.rs:222 drops that read)
                        const uint32_t wv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                        for (int t = 0; t < 16; t++) {
                            const uint32_t q = 16u * c + t;
                            if (q < slack || q >= span) continue;
                            if (base_ok((wv[t >> 2] >> (8 * (t & 3))) & 0xffu)) continue;
                            int lo = 0, hi = nr;   // last r with rstart[r] <= q
                            while (hi - lo > 1) {
                                int m = (lo + hi) >> 1;
                                if (rstart[m] <= q) lo = m; else hi = m;
                            }
                            bad[lo] = 1;
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- observations per read, prefix sums ---------------------------------------------------------------------------
        uint32_t total;
        {
            uint32_t nk = 0;
            if (tid < nr) {
                const uint32_t len = rstart[tid + 1] - rstart[tid];
                if (!bad[tid] && len >= (uint32_t)K) nk = len - K + 1;
            }
            const uint32_t ex = block_scan_u32(nk, &total, ws);
            cum[tid] = ex;
            if (tid == 0) cum[SM_NT] = total;
        }
        if (total == 0) continue;   // (the loop-top barrier orders the cum[] writes before reuse)

        // ---- table sized to the group: power of two >= 2 * observations, at most SM_TSIZE -------------------------------------
        int tbits = 8;
        while (tbits < SM_TBITS && (1u << tbits) < 2u * total) tbits++;
        const uint32_t tsize = 1u << tbits, tmask = tsize - 1u;
        {
            // keys = EMPTY (all ones), vals = 0
            uint4 *kq = reinterpret_cast<uint4 *>(tkey);
            const int nkq = (int)(tsize * sizeof(KT) / 16);
            for (int j = tid; j < nkq; j += SM_NT) kq[j] = make_uint4(~0u, ~0u, ~0u, ~0u);
            uint4 *vq = reinterpret_cast<uint4 *>(tval);
            const int nvq = (int)(tsize / 4);
            for (int j = tid; j < nvq; j += SM_NT) vq[j] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();

        // ---- counting rounds ------------------------------------------------------------------------------------------------
        const uint32_t per = (total + SM_NT - 1) / SM_NT;
        uint32_t oi = (uint32_t)tid * per;                       // next observation index of this thread
        const uint32_t oe = (oi + per < total) ? oi + per : total;
        int r = 0;
        uint32_t p = 0, nkr = 0, q = 0, len = 0;
        bool fresh = true;
        if (oi < oe) {   // last r with cum[r] <= oi and a non-empty read
            int lo = 0, hi = nr;
            while (hi - lo > 1) {
                int m = (lo + hi) >> 1;
                if (cum[m] <= oi) lo = m; else hi = m;
            }
            r = lo;
            p = oi - cum[r];
            nkr = cum[r + 1] - cum[r];
            len = rstart[r + 1] - rstart[r];
            q = rstart[r] + p;
        }
        Key<W> key = Key<W>::zero();
        uint32_t ex = 0, nb = 0, prev_top = 0, h = 0;
        bool have = false, dup = false, claim = false;
        int occupied = 0, claimed_prev = 0;
        bool aborted = false;
        while (true) {
            if (!have && oi < oe) {
                // ---- fetch the next observation: rolling update inside a read, rebuild at read starts -------------------------
                if (p == nkr) {
                    do { r++; } while (cum[r + 1] == cum[r]);
                    p = 0;
                    nkr = cum[r + 1] - cum[r];
                    len = rstart[r + 1] - rstart[r];
                    q = rstart[r];
                    fresh = true;
                }
                uint32_t left;
                if (fresh) {
                    load_window(pk, q, K, key);
                    left = p ? (1u << base_of(pk, q - 1)) : 0u;
                    fresh = false;
                } else {
                    key.push(nb, kbits);
                    left = 1u << prev_top;
                }
                uint32_t right = 0;
                if (p + K < len) {
                    nb = base_of(pk, q + K);
                    right = 16u << nb;
                }
                ex = left | right;
                prev_top = key.top_base(kbits);
                p++;
                q++;
                oi++;
                if (S::pack(key) == S::empty()) {
                    // the all-ones k-mer is the table's EMPTY marker: counted in a side word
                    val_apply_atomic(const_cast<uint32_t *>(&misc[2]), ex);
                } else {
                    h = hash32(key) >> (32 - tbits);
                    have = true;
                }
            }
            // ---- P: probe ---------------------------------------------------------------------------------------------------
            claim = false;
            if (have && !dup) {
                const KT kp = S::pack(key);
                uint32_t guard = 0;
                while (true) {
                    const KT cur = tkey[h];
                    if (cur == kp) break;
                    if (cur == S::empty()) {
                        claim = true;
                        break;
                    }
                    h = (h + 1) & tmask;
                    if (++guard > tsize) {   // cannot happen while occupied <= limit; never spin forever
                        misc[3] = 1;
                        have = false;
                        break;
                    }
                }
            }
            occupied += __syncthreads_count(claimed_prev);   // barrier A
            claimed_prev = 0;
            if (occupied > SM_LIMIT || misc[3]) {
                aborted = true;
                break;
            }
            // ---- C: claim + tag, and deferred duplicates -------------------------------------------------------------------------
            if (have) {
                if (dup) {
                    val_apply_atomic(&tval[h], ex);
                    have = false;
                    dup = false;
                } else {
                    if (claim) tkey[h] = S::pack(key);
                    reinterpret_cast<uint8_t *>(&tval[h])[3] = (uint8_t)tid;
                }
            }
            const int any = __syncthreads_or(have || (oi < oe));   // barrier B
            // ---- V: verify ----------------------------------------------------------------------------------------------------------
            if (have) {
                const KT kp = S::pack(key);
                const KT cur = claim ? tkey[h] : kp;
                if (cur == kp) {
                    claimed_prev = claim;   // duplicates that claimed together over-count: the occupancy bound stays safe
                    const uint32_t v = tval[h];
                    if ((v >> 24) == (uint32_t)tid) {
                        tval[h] = val_apply(v, ex);
                        have = false;
                    } else {
                        dup = true;
                    }
                } else {
                    h = (h + 1) & tmask;
                }
            }
            if (!any) break;
        }
        __syncthreads();
        if (aborted) {   // too many distinct k-mers for a shared-memory table: HBM-table path
            if (tid == 0) a.deferred_ids[atomicAdd(a.n_deferred, 1ull)] = (int32_t)g;
            continue;
        }

        // ---- min-coverage prune + in-place compaction (4 slots per thread per pass; destinations never pass sources) -----------
        uint32_t nS = 0;
        for (uint32_t base = 0; base < tsize; base += 4 * SM_NT) {
            KT k4[4];
            uint32_t v4[4];
            uint32_t cnt = 0;
            bool keep[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint32_t slot = base + 4u * tid + t;
                keep[t] = false;
                if (slot < tsize) {
                    k4[t] = tkey[slot];
                    v4[t] = tval[slot];
                    const uint32_t c = v4[t] & 0xffffu;
                    keep[t] = (k4[t] != S::empty()) && c >= thr;
                }
                cnt += keep[t];
            }
            uint32_t tot;
            uint32_t off = nS + block_scan_u32(cnt, &tot, ws);   // contains the barriers that order reads before writes
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (keep[t]) {
                    tkey[off] = k4[t];
                    tval[off] = v4[t];
                    off++;
                }
            }
            nS += tot;
            __syncthreads();
        }
        {
            const uint32_t sv = misc[2];
            if ((sv & 0xffffu) >= thr) {   // the all-ones k-mer survives
                if (tid == 0) {
                    tkey[nS] = S::empty();
                    tval[nS] = sv;
                }
                nS++;
            }
        }
        __syncthreads();
        if (nS == 0) {
            cta_obs += total;
            continue;
        }
        if (tid == 0) {
            unsigned long long pb = atomicAdd(a.pool_cursor, (unsigned long long)nS);
            misc[4] = (uint32_t)pb;
            misc[5] = (uint32_t)(pb >> 32);
        }
        __syncthreads();
        const int64_t pb = (int64_t)(((unsigned long long)misc[5] << 32) | misc[4]);
        if (pb + (int64_t)nS > a.pool_cap) {   // survivor pool exhausted: let the HBM path (own pool) take the group
            if (tid == 0) a.deferred_ids[atomicAdd(a.n_deferred, 1ull)] = (int32_t)g;
            continue;
        }

        // ---- ascending order: counting sort on the top SM_SORT_BITS bits, then fix the (tiny) buckets --------------------------------
        uint32_t *hist = tval + nS;                                         // SM_SORT_BUCKETS words (nS <= SM_LIMIT)
        uint16_t *order = reinterpret_cast<uint16_t *>(tkey + nS);          // nS entries
        for (int j = tid; j < SM_SORT_BUCKETS; j += SM_NT) hist[j] = 0;
        __syncthreads();
        for (uint32_t j = tid; j < nS; j += SM_NT) atomicAdd(&hist[sort_bucket<W>(S::unpack(tkey[j]), kbits)], 1u);
        __syncthreads();
        {
            constexpr int PER = SM_SORT_BUCKETS / SM_NT;
            uint32_t loc[PER], s = 0;
#pragma unroll
            for (int t = 0; t < PER; t++) {
                loc[t] = hist[tid * PER + t];
                s += loc[t];
            }
            uint32_t tot;
            uint32_t e0 = block_scan_u32(s, &tot, ws);
#pragma unroll
            for (int t = 0; t < PER; t++) {
                hist[tid * PER + t] = e0;
                e0 += loc[t];
            }
        }
        __syncthreads();
        for (uint32_t j = tid; j < nS; j += SM_NT) {
            const uint32_t pos = atomicAdd(&hist[sort_bucket<W>(S::unpack(tkey[j]), kbits)], 1u);
            order[pos] = (uint16_t)j;
        }
        __syncthreads();
        for (int b = tid; b < SM_SORT_BUCKETS; b += SM_NT) {   // hist[b] is now the END of bucket b
            const uint32_t s = b ? hist[b - 1] : 0u, e = hist[b];
            for (uint32_t i = s + 1; i < e; i++) {             // insertion sort by key
                const uint16_t oi2 = order[i];
                const KT ki = tkey[oi2];
                uint32_t j = i;
                while (j > s && tkey[order[j - 1]] > ki) {
                    order[j] = order[j - 1];
                    j--;
                }
                order[j] = oi2;
            }
        }
        __syncthreads();

        // ---- emit (coalesced) ------------------------------------------------------------------------------------------------------------
        for (uint32_t j = tid; j < nS; j += SM_NT) {
            const uint32_t idx = order[j];
            const Key<W> k = S::unpack(tkey[idx]);
            const uint32_t v = tval[idx];
            a.pool_lo[pb + j] = k.lo;
            if (W == 2) a.pool_hi[pb + j] = k.hi_word();
            a.pool_exts[pb + j] = (uint8_t)(v >> 16);
            a.pool_counts[pb + j] = (uint16_t)(v & 0xffffu);
        }
        if (tid == 0) {
            a.nsurv[g] = (int64_t)nS;
            a.pool_start[g] = pb;
        }
        cta_obs += total;
    }
    if (tid == 0 && cta_obs) atomicAdd(a.n_obs, cta_obs);
}

// upper bound of the observations of a column: sum over rows of max(0, len-K+1) (validity ignored)
__global__ void k_obs_upper_bound(ColView col, int K, unsigned long long *out)
{
    unsigned long long s = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < col.n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t len = col_off(col, r + 1) - col_off(col, r);
        if (len >= K) s += (unsigned long long)(len - K + 1);
    }
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// -----------------------------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------------------------
struct SmemResult {
    const int32_t *deferred_ids = nullptr;
    int64_t n_deferred = 0;
    int64_t n_pool_groups = 0;
    const int64_t *pool_start = nullptr;
    const uint64_t *pool_lo = nullptr, *pool_hi = nullptr;
    const uint8_t *pool_exts = nullptr;
    const uint16_t *pool_counts = nullptr;
    std::vector<DevBuf> keep;
    float ms_kernel = 0.f;
};

inline void smem_path_configure()
{
    CK(cudaFuncSetAttribute(k_smem_count<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<1>()));
    CK(cudaFuncSetAttribute(k_smem_count<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<2>()));
}

inline void run_smem_path(cudaStream_t st, int sm_count, size_t workspace_bytes, const ColView &col, const uint8_t *data_end,
                          const int64_t *d_group_offsets, int64_t n_groups, int K, uint32_t min_cov, unsigned long long *d_nobs,
                          int64_t *d_nsurv, SmemResult &res, int64_t *launches, cudaEvent_t ev_a, cudaEvent_t ev_b)
{
    const int W = K > 32 ? 2 : 1;
    DevBuf ctr(64, st);
    CK(cudaMemsetAsync(ctr.p, 0, 64, st));
    unsigned long long *d_ctr = ctr.as<unsigned long long>();   // [0] ticket [1] pool cursor [2] n_deferred [3] obs upper bound
    k_obs_upper_bound<<<sm_count * 4, 256, 0, st>>>(col, K, d_ctr + 3);
    unsigned long long ub = 0;
    CK(cudaMemcpyAsync(&ub, d_ctr + 3, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // a k-mer needs min_cov observations to survive, so survivors <= observations / min_cov
    const size_t entry = (W == 2 ? 19 : 11);
    int64_t pool_cap = (int64_t)(ub / (min_cov > 1 ? min_cov : 1)) + 1;
    pool_cap = std::min<int64_t>(pool_cap, (int64_t)(workspace_bytes / 2 / entry));
    pool_cap = std::max<int64_t>(pool_cap, 1024);
    res.keep.emplace_back((size_t)pool_cap * 8, st);
    uint64_t *p_lo = res.keep.back().as<uint64_t>();
    uint64_t *p_hi = nullptr;
    if (W == 2) {
        res.keep.emplace_back((size_t)pool_cap * 8, st);
        p_hi = res.keep.back().as<uint64_t>();
    }
    res.keep.emplace_back((size_t)pool_cap, st);
    uint8_t *p_exts = res.keep.back().as<uint8_t>();
    res.keep.emplace_back((size_t)pool_cap * 2, st);
    uint16_t *p_counts = res.keep.back().as<uint16_t>();
    res.keep.emplace_back((size_t)n_groups * 8, st);
    int64_t *p_start = res.keep.back().as<int64_t>();
    CK(cudaMemsetAsync(p_start, 0xff, (size_t)n_groups * 8, st));
    res.keep.emplace_back((size_t)n_groups * 4, st);
    int32_t *d_ids = res.keep.back().as<int32_t>();

    SmemArgs a{};
    a.col = col;
    a.data_end = data_end;
    a.group_offsets = d_group_offsets;
    a.n_groups = n_groups;
    a.K = K;
    a.min_cov = min_cov;
    a.ticket = d_ctr;
    a.n_obs = d_nobs;
    a.nsurv = d_nsurv;
    a.pool_start = p_start;
    a.pool_cursor = d_ctr + 1;
    a.pool_cap = pool_cap;
    a.pool_lo = p_lo;
    a.pool_hi = p_hi;
    a.pool_exts = p_exts;
    a.pool_counts = p_counts;
    a.deferred_ids = d_ids;
    a.n_deferred = d_ctr + 2;
    CK(cudaEventRecord(ev_a, st));
    if (W == 1) {
        const unsigned grid = (unsigned)std::min<int64_t>(n_groups, (int64_t)sm_count * 2);
        k_smem_count<1><<<grid, SM_NT, smem_bytes<1>(), st>>>(a);
    } else {
        const unsigned grid = (unsigned)std::min<int64_t>(n_groups, (int64_t)sm_count);
        k_smem_count<2><<<grid, SM_NT, smem_bytes<2>(), st>>>(a);
    }
    CK(cudaEventRecord(ev_b, st));
    if (launches) *launches += 2;
    unsigned long long h_ctr[3];
    CK(cudaMemcpyAsync(h_ctr, d_ctr, 24, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&res.ms_kernel, ev_a, ev_b));
    res.n_deferred = (int64_t)h_ctr[2];
    res.n_pool_groups = n_groups - res.n_deferred;
    res.deferred_ids = d_ids;
    res.pool_start = p_start;
    res.pool_lo = p_lo;
    res.pool_hi = p_hi;
    res.pool_exts = p_exts;
    res.pool_counts = p_counts;
    res.keep.push_back(std::move(ctr));
}

}  // namespace rogtk
