// kernels_smem.cuh -- the hot kernel: one CTA builds one UMI group's pruned de Bruijn node set entirely in shared
// memory (ingest -> 2-bit pack -> k-mer windows -> open-addressing table -> min-coverage compaction into a pool), plus
// the per-group ordering kernels that turn the pool into the dense, ascending output.  Together they replace
// debruijn::filter::filter_kmers + CountFilter as called at reference src/fracture.rs:110-116.
//
// Table protocol: open addressing with linear probing in shared memory, keys claimed with one 64/128-bit
// atomicCAS, values updated with one 32-bit atomicAdd (count in bits 8..31, exts in bits 0..7, so the add can never
// carry into the exts and a group of <= 16384 bases cannot overflow 24 bits); the add returns the old word and an
// atomicOr follows only when this observation brings a new exts bit.  Measured on B200 (tools/microbench_smem.cu,
// profiles/r01_microbench_smem.txt): random-address shared atomicAdd sustains 16 ops/ns/SM, more than a plain
// LDS+STS read-modify-write, so no barrier-synchronised election scheme is needed; warps run the whole counting
// phase without a single __syncthreads.
//
// The same kernel body, instantiated with MULTI = true, takes the groups that exceed one table or one staged span: the k-mer
// space is cut into P ranges and the group's reads are re-staged slab by slab once per range.  A count of
// (count << 8) cannot overflow there either: a multi-pass group has at most SM_PMAX * SM_PASS_OBS = 131 072 bases.
#pragma once
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "host_util.cuh"
#include "kernels_table.cuh"

namespace rogtk {

constexpr int SM_NT = 384;            // threads per CTA, one-word keys (2 CTAs per SM)
constexpr int SM_NT2 = 640;           // two-word keys: the 160 KB table leaves one CTA per SM, so it gets more warps
template <int W>
constexpr int sm_nt() { return W == 2 ? SM_NT2 : SM_NT; }
template <int W>
constexpr int sm_nwarp() { return sm_nt<W>() / 32; }
// words of the 64-word scratch `ws`: block-scan partials, misc, group id (MULTI), per-warp queue leftovers
constexpr int WS_SCAN = 0, WS_MISC = 24, WS_GID = 40, WS_LEFT = 42;
constexpr int SM_MAXR = 256;          // max rows per group on this path
constexpr int SM_MAX_BASES = 16384;   // bases staged per group (including <=15 bytes of alignment slack)
constexpr int SM_PK_WORDS = SM_MAX_BASES / 16 + 8;   // 1 word of front pad, tail pad for window over-reads
constexpr int SM_UNIT = 128;          // windows per work unit (one warp, 4 consecutive windows per lane)
constexpr int SM_QCAP = 64;           // per-warp queue of collided observations (drained 32 at a time)
constexpr int SM_MAX_UNITS = SM_MAX_BASES / SM_UNIT + SM_MAXR;   // work units of one group
constexpr int SM_TBITS = 13;          // 8192 slots
constexpr int SM_TSIZE = 1 << SM_TBITS;
// slots of the table a CTA can hold.  One-word keys: 8192 (12 B per slot, two CTAs per SM).  Two-word keys have the SM to
// themselves, and K = 64 makes every read error spoil 64 windows, so the table takes what fits: 9728 slots of 20 B (not a power
// of two: home slot = umulhi(hash, slots), wrap-around by comparison).
template <int W>
constexpr int table_slots() { return W == 2 ? 9728 : SM_TSIZE; }
template <int W>
__device__ __forceinline__ uint32_t home_slot(uint32_t hash, uint32_t tsize, int tshift) { return W == 2 ? __umulhi(hash, tsize) : hash >> tshift; }
template <int W>
__device__ __forceinline__ uint32_t next_slot(uint32_t h, uint32_t tsize)
{
    if (W == 2) return h + 1u == tsize ? 0u : h + 1u;
    return (h + 1u) & (tsize - 1u);
}
template <int W>
__device__ __forceinline__ uint32_t prev_slot(uint32_t h, uint32_t tsize)
{
    if (W == 2) return h == 0u ? tsize - 1u : h - 1u;
    return (h - 1u) & (tsize - 1u);
}
constexpr int SM_SORT_BITS = 10;
constexpr int SM_SORT_BUCKETS = 1 << SM_SORT_BITS;
constexpr int SM_POOL_CHUNK = 65536;   // survivors a CTA reserves from the pool at a time (a group needs room for SM_TSIZE + 1)
constexpr int SM_MAX_BUCKET_WARP = 512;   // survivors sharing a sort prefix beyond this (all-pairs ranking gets quadratic): the item goes to the bitonic kernel
// Multi-pass variant for groups that do not fit one table / one staged span (MULTI = true): the k-mer space is cut into
// P equal ranges (range of a k-mer = floor(top 32 bits * P / 2^32)); pass p re-reads the group's reads in slabs of
// <= SM_MAX_BASES bases and counts only the k-mers of range p, so every pass needs a table for about 1/P of the distinct
// k-mers and its survivors form one ascending segment of the group's output.
constexpr int SM_PMAX = 8;                    // passes per group at most (larger groups: one CTA would become the tail of the launch)
constexpr int64_t SM_PASS_OBS = 16384;        // bases per pass the first attempt plans for (doubles the passes on overflow)

struct SegRec {          // survivors of one pass of a multi-pass group
    int64_t pool_start;
    int32_t n;
    int32_t prefix;      // survivors of the earlier passes of the group
};
template <int W>
constexpr int scratch_bytes()   // pk + per-warp queues; reused by the sort (histogram + order)
{
    return SM_PK_WORDS * 4 + sm_nwarp<W>() * SM_QCAP * (int)(sizeof(typename KeySlot<W>::T) + 4);
}

template <int W>
constexpr size_t smem_bytes()
{
    return (size_t)table_slots<W>() * sizeof(typename KeySlot<W>::T) + (size_t)table_slots<W>() * 4 + (size_t)scratch_bytes<W>() +
           (size_t)(SM_MAXR + 1) * 4 * 2 + 64 * 4 + SM_MAXR + SM_MAX_UNITS * 2 + 64;
}

struct SmemArgs {
    ColView col;
    const uint8_t *data_begin;       // first readable byte of the data buffer (col.data may be a virtual base below it)
    const uint8_t *data_end;         // one past the last readable byte
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
    unsigned long long *phase_cycles;   // optional [16]: [0..7] per-phase SM cycles summed over CTAs, [8..15] event counts of k_walk_count (diagnostics)
    // MULTI only: the groups to process (tickets index this list), one SegRec per (list entry, pass), passes used per entry
    const int32_t *group_list;
    int64_t n_list;
    SegRec *segs;                    // [n_list * SM_PMAX]
    int32_t *seg_count;              // [n_list], 0 = the group was not finished here
};

// top 32 bits of a k-mer of kbits bits (shorter k-mers are left-aligned)
__device__ __forceinline__ uint32_t key_top32(const Key<1> &k, int kbits)
{
    return kbits >= 32 ? (uint32_t)(k.lo >> (kbits - 32)) : (uint32_t)k.lo << (32 - kbits);
}
__device__ __forceinline__ uint32_t key_top32(const Key<2> &k, int kbits)
{
    const int sh = kbits - 32;   // 34 .. 96
    return sh >= 64 ? (uint32_t)(k.hi >> (sh - 64)) : (uint32_t)((k.hi << (64 - sh)) | (k.lo >> sh));
}
// range (pass) of a k-mer among P equal ranges of the k-mer space; monotone in the k-mer
__device__ __forceinline__ uint32_t pass_of(uint32_t top32, uint32_t P) { return __umulhi(top32, P); }
// first top-32 value of range p: smallest x with floor(x * P / 2^32) == p
__device__ __forceinline__ uint32_t pass_first(uint32_t p, uint32_t P)
{
    return (uint32_t)((((uint64_t)p << 32) + P - 1u) / P);
}

// multiplicative hash, slot = top bits.  On the synthetic shapes it probes 1.32 slots per observation vs 1.31 for a
// two-round mixer, for half the instructions.
__device__ __forceinline__ uint32_t hash32(uint32_t lo, uint32_t hi) { return lo * 0x9E3779B1u + hi * 0x85EBCA77u; }
__device__ __forceinline__ uint32_t hash32(const Key<1> &k) { return hash32((uint32_t)k.lo, (uint32_t)(k.lo >> 32)); }
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
// base at index q; q == -1 (as uint32) reads the front pad word
__device__ __forceinline__ uint32_t base_of(const uint32_t *pk, uint32_t q)
{
    return (pk[(int32_t)q >> 4] >> (30u - 2u * (q & 15u))) & 3u;
}

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

// block-wide exclusive scan of one u32 per thread (NW warps); ws = NW words of shared scratch
template <int NW>
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
    for (int w = 0; w < NW; w++) {
        uint32_t s = ws[w];
        if (w < (int)warp) base += s;
        tot += s;
    }
    *total = tot;
    __syncthreads();
    return base + x - v;
}

// ---- ingest helpers kept out of line: they run only for chunks at the edge of the data buffer / with a bad byte ------------
__device__ __noinline__ uint4 load_chunk_guarded(const uint8_t *p, const uint8_t *lo, const uint8_t *hi)
{
    uint32_t wv[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};
    for (int t = 0; t < 16; t++) {
        const uint8_t *pb = p + t;
        if (pb >= lo && pb < hi) wv[t >> 2] = (wv[t >> 2] & ~(0xffu << (8 * (t & 3)))) | ((uint32_t)*pb << (8 * (t & 3)));
    }
    return make_uint4(wv[0], wv[1], wv[2], wv[3]);
}
// attribute every bad byte of a chunk to its read (reference src/fracture.rs:222 drops that read)
__device__ __noinline__ void mark_bad_reads(uint4 v, uint32_t qbase, uint32_t slack, uint32_t span, const uint32_t *rstart, int nr,
                                            uint8_t *bad)
{
    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
    for (int t = 0; t < 16; t++) {
        const uint32_t q = qbase + t;
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

// ---- the table ------------------------------------------------------------------------------------------------------------------
// count in bits 8..31, exts in bits 0..7: the add can never carry into the exts; the returned old word tells whether this
// observation brings a new exts bit.
__device__ __forceinline__ void val_add(uint32_t *vp, uint32_t ex)
{
    const uint32_t old = atomicAdd(vp, 256u);
    if (ex & ~old) atomicOr(vp, ex);
}

template <int W>
struct WarpQueue {   // collided observations of one warp, waiting for their next probe
    typename KeySlot<W>::T *key;
    uint32_t *hx;    // slot | exts << 16
    uint32_t n;      // entries (warp-uniform)
    int budget;      // drain steps left before the table is declared full (warp-uniform)
};

// One probe step for up to 32 queued observations (one per lane, full width).  Resolved ones are applied, the others are
// pushed back.  Returns false when the probe bound is exceeded (table hopelessly full).
template <int W>
__device__ __forceinline__ void queue_drain_step(WarpQueue<W> &q, typename KeySlot<W>::T *tkey, uint32_t *tval, uint32_t tsize,
                                                 uint32_t lane)
{
    using S = KeySlot<W>;
    using KT = typename S::T;
    const uint32_t take = q.n < 32u ? q.n : 32u;
    const uint32_t base = q.n - take;
    bool pend = false;
    KT kp = S::empty();
    uint32_t hx = 0;
    if (lane < take) {
        kp = q.key[base + lane];
        hx = q.hx[base + lane];
        uint32_t h = next_slot<W>(hx & 0xffffu, tsize);
        KT cur = tkey[h];
        if (cur == S::empty()) cur = atomicCAS(&tkey[h], S::empty(), kp);
        if (cur == kp || cur == S::empty()) {
            val_add(&tval[h], hx >> 16);
        } else {
            pend = true;
            hx = (hx & 0xffff0000u) | h;
        }
    }
    __syncwarp();
    const uint32_t m = __ballot_sync(0xffffffffu, pend);
    if (pend) {
        const uint32_t i = base + __popc(m & ((1u << lane) - 1u));
        q.key[i] = kp;
        q.hx[i] = hx;
    }
    q.n = base + __popc(m);
    __syncwarp();
}

template <int W>
__device__ __forceinline__ void queue_push(WarpQueue<W> &q, bool pend, typename KeySlot<W>::T kp, uint32_t h, uint32_t ex,
                                           typename KeySlot<W>::T *tkey, uint32_t *tval, uint32_t tsize, uint32_t lane)
{
    const uint32_t m = __ballot_sync(0xffffffffu, pend);
    if (m == 0) return;
    if (pend) {
        const uint32_t i = q.n + __popc(m & ((1u << lane) - 1u));
        q.key[i] = kp;
        q.hx[i] = h | (ex << 16);
    }
    q.n += __popc(m);
    __syncwarp();
    // keep fewer than 32 entries queued, so that the next push (<= 32 entries) always fits in SM_QCAP
    while (q.n >= 32u && q.budget > 0) {
        queue_drain_step<W>(q, tkey, tval, tsize, lane);
        q.budget--;
    }
    if (q.n >= 32u) q.n = 0;   // budget exhausted: the caller sees budget <= 0 and gives the group up
}

#define PHASE_MARK(idx)                                                  \
    do {                                                                 \
        if (a.phase_cycles && tid == 0) {                                \
            const long long _t = clock64();                              \
            atomicAdd(a.phase_cycles + (idx), (unsigned long long)(_t - t_phase)); \
            t_phase = _t;                                                \
        }                                                                \
    } while (0)

// KF: the K the wide window path is specialised for (32 with one-word keys, 64 with two-word keys), 0 = generic path only
template <int W, int KF, bool MULTI>
__global__ void __launch_bounds__(sm_nt<W>(), (W == 1 ? 2 : 1)) k_smem_count(SmemArgs a)
{
    constexpr int NT = sm_nt<W>(), NW = sm_nwarp<W>();
    constexpr bool K32 = KF == 32, K64 = KF == 64;
    static_assert(NW <= WS_MISC && WS_LEFT + NW <= 64, "scratch words: scan partials and queue leftovers are per warp");
    static_assert(KF == 0 || (KF == 32 && W == 1) || (KF == 64 && W == 2), "wide path: K = 32 or K = 64");
    using S = KeySlot<W>;
    using KT = typename S::T;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    KT *tkey = reinterpret_cast<KT *>(smem_raw);
    uint32_t *tval = reinterpret_cast<uint32_t *>(tkey + table_slots<W>());
    uint32_t *scratch = tval + table_slots<W>();
    uint32_t *pk = scratch + 1;                                                  // pk[-1] is a readable pad word
    uint8_t *qbytes = reinterpret_cast<uint8_t *>(scratch + SM_PK_WORDS);        // per-warp queues (W=1: 8 B keys)
    uint32_t *rstart = scratch + scratch_bytes<W>() / 4;
    uint32_t *ucum = rstart + SM_MAXR + 1;               // work units before read r
    uint32_t *ws = ucum + SM_MAXR + 1;                     // 64 words, see WS_*
    uint8_t *bad = reinterpret_cast<uint8_t *>(ws + 64);
    uint16_t *unit_read = reinterpret_cast<uint16_t *>(bad + SM_MAXR);   // read index of every work unit
    volatile uint32_t *misc = ws + WS_MISC;   // [0],[1] ticket [2] all-ones k-mer value [3] abort [4],[5] pool base [6] big bucket [7] unit counter [8..15] r0,r1,b0,b1
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const int K = a.K;
    // CountFilter compares the SATURATED u16 count: nothing can reach a min_coverage above 65535
    const uint32_t thr = a.min_cov > 65535u ? 0xffffffffu : (a.min_cov > 1u ? a.min_cov : 1u);
    const int64_t n_items = MULTI ? a.n_list : a.n_groups;
    unsigned long long cta_obs = 0;
    long long t_phase = clock64();
    WarpQueue<W> wq;
    wq.key = reinterpret_cast<KT *>(qbytes) + warp * SM_QCAP;
    wq.hx = reinterpret_cast<uint32_t *>(reinterpret_cast<KT *>(qbytes) + NW * SM_QCAP) + warp * SM_QCAP;
    wq.n = 0;
    wq.budget = 0;

    // thread 0 runs one group ahead: ticket, row range and byte range of the next group are fetched while the current one is
    // being counted, so their global-memory latencies are off the critical path.
    unsigned long long pf_g = 0;
    int64_t pf_gid = 0, pf_r0 = 0, pf_r1 = 0, pf_b0 = 0, pf_b1 = 0;
    int pf_stage = 0;   // 0: have pf_g   1: + row range   2: + byte range
    int64_t cta_pool_pos = 0, cta_pool_end = 0;   // this CTA's chunk of the survivor pool (thread 0)
    if (tid == 0) pf_g = atomicAdd(a.ticket, 1ull);
    auto pf_advance = [&]() {   // thread 0 only
        if ((int64_t)pf_g >= n_items) {
            pf_stage = 2;
            return;
        }
        if (pf_stage == 0) {
            pf_gid = MULTI ? (int64_t)a.group_list[pf_g] : (int64_t)pf_g;
            pf_r0 = a.group_offsets[pf_gid];
            pf_r1 = a.group_offsets[pf_gid + 1];
            pf_stage = 1;
        } else if (pf_stage == 1) {
            pf_b0 = col_off(a.col, pf_r0);
            pf_b1 = col_off(a.col, pf_r1);
            pf_stage = 2;
        }
    };
    auto defer = [&](int64_t gid) {   // hand a group to the next path
        if (tid == 0) a.deferred_ids[atomicAdd(a.n_deferred, 1ull)] = (int32_t)gid;
    };

    // the group being worked on; MULTI keeps it over several iterations of the loop (one slab of one pass each)
    int64_t item = 0, g = 0, r0 = 0, r1 = 0, b0 = 0, b1 = 0;
    bool in_group = false;             // MULTI: a group is in progress
    int P = 1, pass = 0;               // MULTI: number of passes, current pass
    int64_t srow = 0;                  // MULTI: first row of the next slab
    uint32_t gprefix = 0, gobs = 0;    // MULTI: survivors of the finished passes, observations of the group

    while (true) {
        __syncthreads();
        if (!MULTI || !in_group) {
            if (tid == 0) {
                while (pf_stage < 2) pf_advance();
                misc[0] = (uint32_t)pf_g;
                misc[1] = (uint32_t)(pf_g >> 32);
                misc[2] = 0;
                misc[3] = 0;
                misc[6] = 0;
                misc[7] = 0;
                misc[8] = (uint32_t)pf_r0;  misc[9] = (uint32_t)((uint64_t)pf_r0 >> 32);
                misc[10] = (uint32_t)pf_r1; misc[11] = (uint32_t)((uint64_t)pf_r1 >> 32);
                misc[12] = (uint32_t)pf_b0; misc[13] = (uint32_t)((uint64_t)pf_b0 >> 32);
                misc[14] = (uint32_t)pf_b1; misc[15] = (uint32_t)((uint64_t)pf_b1 >> 32);
                if (MULTI) {
                    ws[WS_GID] = (uint32_t)pf_gid;
                    ws[WS_GID + 1] = (uint32_t)((uint64_t)pf_gid >> 32);
                }
            }
            __syncthreads();
            item = (int64_t)(((unsigned long long)misc[1] << 32) | misc[0]);
            if (item >= n_items) break;
            g = MULTI ? (int64_t)(((uint64_t)ws[WS_GID + 1] << 32) | ws[WS_GID]) : item;
            r0 = (int64_t)(((uint64_t)misc[9] << 32) | misc[8]);
            r1 = (int64_t)(((uint64_t)misc[11] << 32) | misc[10]);
            b0 = (int64_t)(((uint64_t)misc[13] << 32) | misc[12]);
            b1 = (int64_t)(((uint64_t)misc[15] << 32) | misc[14]);
            if (tid == 0) {   // take the next ticket now; its metadata follows at later points of this iteration
                pf_g = atomicAdd(a.ticket, 1ull);
                pf_stage = 0;
            }
            if (r1 - r0 <= 0) continue;
            if (b1 <= b0) continue;
            if (MULTI) {
                const int64_t ub = b1 - b0;   // every observation starts at its own base
                const int64_t pass_obs = K > 32 ? SM_PASS_OBS / 2 : SM_PASS_OBS;
                if (ub > SM_PASS_OBS * SM_PMAX) {   // too much work for one CTA: HBM-table path
                    defer(g);
                    continue;
                }
                P = 1;
                // (<= 4^K <= 4096 distinct k-mers when K <= 6; long k-mers: every read error spoils K windows, plan half the bases per pass)
                if (2 * K > 12) P = (int)((ub + pass_obs - 1) / pass_obs);
                if (P > SM_PMAX) P = SM_PMAX;
                in_group = true;
                pass = 0;
                srow = r0;
                gprefix = 0;
                gobs = 0;
            }
        }

        // ---- the rows staged in this iteration: the whole group, or (MULTI) the next slab of it -----------------------
        int64_t s0 = r0, s1 = r1, sb0 = b0, sb1 = b1;
        if (MULTI) {
            s0 = srow;
            sb0 = col_off(a.col, s0);
            const uint32_t slk = (uint32_t)(reinterpret_cast<uintptr_t>(a.col.data + sb0) & 15u);
            bool fits = false;   // rows s0 .. s0+tid fit one slab (monotone in tid: the offsets are validated as monotone)
            if (tid < SM_MAXR && s0 + tid + 1 <= r1) fits = col_off(a.col, s0 + tid + 1) - sb0 + (int64_t)slk <= (int64_t)SM_MAX_BASES;
            const int cnt = __syncthreads_count(fits);
            if (cnt == 0) {   // a single read longer than a slab
                defer(g);
                in_group = false;
                continue;
            }
            s1 = s0 + cnt;
            sb1 = col_off(a.col, s1);
            if (tid == 0) {
                misc[7] = 0;
                if (s0 == r0) {
                    misc[2] = 0;
                    misc[3] = 0;
                }
            }
            if (s0 == r0) {   // a pass starts: empty table of full size
                uint4 *kq = reinterpret_cast<uint4 *>(tkey);
                for (int j = tid; j < (int)(table_slots<W>() * sizeof(KT) / 16); j += NT) kq[j] = make_uint4(~0u, ~0u, ~0u, ~0u);
                uint4 *vq = reinterpret_cast<uint4 *>(tval);
                for (int j = tid; j < table_slots<W>() / 4; j += NT) vq[j] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        const int64_t nrows = s1 - s0;
        const uintptr_t addr0 = reinterpret_cast<uintptr_t>(a.col.data + sb0);
        const uint32_t slack = (uint32_t)(addr0 & 15u);
        const int64_t span64 = (sb1 - sb0) + slack;
        if (!MULTI && (nrows > SM_MAXR || span64 > SM_MAX_BASES)) {   // too many rows / bases for one staged span
            defer(g);
            continue;
        }
        const uint32_t span = (uint32_t)span64;
        const int nr = (int)nrows;

        // ---- per-read base offsets inside the staged span, null rows ----------------------------------------------
        for (int r = tid; r <= nr; r += NT) rstart[r] = (uint32_t)(col_off(a.col, s0 + r) - sb0) + slack;
        if (tid < nr) bad[tid] = col_valid(a.col, s0 + tid) ? 0 : 1;
        __syncthreads();

        PHASE_MARK(0);
        // ---- ingest: 16-byte chunks -> one packed u32 (16 bases, first base in the top bits) --------------------------
        {
            const uint8_t *a0 = reinterpret_cast<const uint8_t *>(addr0 - slack);
            const int nchunks = (int)((span + 15u) >> 4);
            uint4 v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = tid + i * NT;
                v[i] = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
                if (c < nchunks) {
                    const uint8_t *p = a0 + 16 * c;
                    if (p >= a.data_begin && p + 16 <= a.data_end) {
                        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                     : "l"(p));
                    } else {
                        v[i] = load_chunk_guarded(p, a.data_begin, a.data_end);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = tid + i * NT;
                if (c < nchunks) {
                    pk[c] = (pack4(v[i].x) << 24) | (pack4(v[i].y) << 16) | (pack4(v[i].z) << 8) | pack4(v[i].w);
                    const bool ok = valid4(v[i].x) & valid4(v[i].y) & valid4(v[i].z) & valid4(v[i].w);
                    if (!ok) mark_bad_reads(v[i], 16u * c, slack, span, rstart, nr, bad);
                }
            }
        }
        __syncthreads();

        PHASE_MARK(1);
        // ---- windows per read -> work units of SM_UNIT windows, prefix sums ------------------------------------------------------
        uint32_t total, nunits;
        {
            uint32_t nk = 0;
            if (tid < nr) {
                const uint32_t len = rstart[tid + 1] - rstart[tid];
                if (!bad[tid] && len >= (uint32_t)K) nk = len - K + 1;
            }
            block_scan_u32<NW>(nk, &total, ws);
            const uint32_t ex = block_scan_u32<NW>((nk + SM_UNIT - 1) / SM_UNIT, &nunits, ws);
            if (tid < SM_MAXR) ucum[tid] = ex;
            if (tid == 0) ucum[SM_MAXR] = nunits;
            for (uint32_t j = 0; j < (nk + SM_UNIT - 1) / SM_UNIT; j++) unit_read[ex + j] = (uint16_t)tid;
        }
        if (!MULTI && total == 0) continue;   // (the loop-top barrier orders the shared writes before reuse)

        // ---- table sized to the group: power of two >= 2 * observations, at most SM_TSIZE (MULTI: always SM_TSIZE) ---------------
        int tbits = MULTI ? SM_TBITS : 8;
        if (!MULTI)
            while (tbits < SM_TBITS && (1u << tbits) < 2u * total) tbits++;
        // (two-word keys: the largest table is not a power of two, see table_slots)
        const uint32_t tsize = (MULTI || (W == 2 && 2u * total > (uint32_t)SM_TSIZE)) ? (uint32_t)table_slots<W>() : 1u << tbits;
        const int tshift = 32 - tbits;
        if (!MULTI) {
            // keys = EMPTY (all ones), vals = 0
            uint4 *kq = reinterpret_cast<uint4 *>(tkey);
            const int nkq = (int)(tsize * sizeof(KT) / 16);
            for (int j = tid; j < nkq; j += NT) kq[j] = make_uint4(~0u, ~0u, ~0u, ~0u);
            uint4 *vq = reinterpret_cast<uint4 *>(tval);
            const int nvq = (int)(tsize / 4);
            for (int j = tid; j < nvq; j += NT) vq[j] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();

        PHASE_MARK(2);
        // ---- counting: warps fetch units of SM_UNIT windows of one read; no block barriers ---------------------------------------------
        if (tid == 0 && pf_stage == 0) pf_advance();   // next group's row range (its ticket has long arrived)
        // MULTI: this pass takes the k-mers of range `pass` (pass_of)
        if (total != 0) {
            wq.n = 0;
            wq.budget = 64 + (int)(total >> 1);   // drain steps allowed per warp: far above any non-full table's need
            uint32_t *special = const_cast<uint32_t *>(&misc[2]);
            const uint32_t unit_ctr_addr = (uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&misc[7]));
            // first probe + claim of one observation; returns true when it must queue for further probing
            auto first_probe = [&](KT kp, uint32_t h, uint32_t ex, KT cur) -> bool {
                if (kp == S::empty()) {   // the all-ones k-mer is the EMPTY marker: counted in a side word
                    val_add(special, ex);
                    return false;
                }
                if (cur == S::empty()) cur = atomicCAS(&tkey[h], S::empty(), kp);
                if (cur == kp || cur == S::empty()) {
                    val_add(&tval[h], ex);
                    return false;
                }
                return true;
            };
            while (true) {
                uint32_t unit = 0;
                // (inline PTX: for a warp-uniform address the compiler wraps atomicAdd in a vote/popc/shuffle aggregation
                //  sequence -- 17 instructions per unit for an add that only lane 0 performs)
                if (lane == 0)
                    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(unit) : "r"(unit_ctr_addr));
                unit = __shfl_sync(0xffffffffu, unit, 0);
                if (unit >= nunits || misc[3] || wq.budget <= 0) break;
                const int lo = unit_read[unit];   // read of this unit
                const uint32_t rs = rstart[lo];
                const uint32_t nk = rstart[lo + 1] - rs - (uint32_t)K + 1u;
                const uint32_t p0 = (unit - ucum[lo]) * SM_UNIT;
                const uint32_t rem = nk - p0;
                const uint32_t ulen = rem < (uint32_t)SM_UNIT ? rem : (uint32_t)SM_UNIT;   // windows of this unit
                // The 4-windows-per-lane path costs the same for any number of windows, the generic path costs one pass per 32
                // windows: short read tails (269 = 128 + 128 + 13) take the generic path, everything else the wide one.
                if (K32 && ulen > 48u) {
                    // ---- K == 32: lane owns R <= 4 consecutive windows p0+R*lane ..; five packed words hold all their bases ---------
                    const uint32_t R = (ulen + 31u) >> 5;
                    const uint32_t j0 = R * lane;                        // first window of this lane inside the unit
                    const uint32_t p = p0 + (j0 < ulen ? j0 : 0u);       // idle lanes read the unit's first words (in bounds)
                    const uint32_t q = rs + p;
                    const int32_t w = (int32_t)(q >> 4);
                    const uint32_t s = 2u * (q & 15u);
                    const uint32_t xm = pk[w - 1], x0 = pk[w], x1 = pk[w + 1], x2 = pk[w + 2], x3 = pk[w + 3];
                    uint32_t kh[4], kl[4], ex[4], h[4];
                    bool act[4];
                    kh[0] = __funnelshift_l(x1, x0, s);
                    kl[0] = __funnelshift_l(x2, x1, s);
                    const uint32_t tail = __funnelshift_l(x3, x2, s);        // bases q+32 .. q+47
                    const uint32_t lb0 = __funnelshift_l(x0, xm, s) & 3u;    // base q-1
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t nb = (tail >> (30 - 2 * u)) & 3u;     // base q+32+u: right neighbour of window u
                        const uint32_t lb = u ? (kh[u - 1] >> 30) : lb0;     // left neighbour of window u
                        const bool has_l = u ? true : (p != 0u);
                        const bool has_r = (p + u + 1u) < nk;
                        act[u] = (uint32_t)u < R && j0 + (uint32_t)u < ulen;
                        if (MULTI) act[u] = act[u] && pass_of(kh[u], (uint32_t)P) == (uint32_t)pass;
                        ex[u] = (has_l ? (1u << lb) : 0u) | (has_r ? (16u << nb) : 0u);
                        h[u] = home_slot<W>(hash32(kl[u], kh[u]), tsize, tshift);
                        if (u < 3) {
                            kh[u + 1] = __funnelshift_l(kl[u], kh[u], 2);
                            kl[u + 1] = (kl[u] << 2) | nb;
                        }
                    }
                    const bool sparse = MULTI && P > 1;   // few windows of the unit belong to this pass
                    KT cur[4];
                    if (!sparse) {
#pragma unroll
                        for (int u = 0; u < 4; u++) cur[u] = act[u] ? tkey[h[u]] : S::empty();
                    }
                    // the all-ones k-mer is the table's EMPTY marker: divert it to the side word first (one warp vote, rarely taken)
                    {
                        bool sp = false;
#pragma unroll
                        for (int u = 0; u < 4; u++) sp |= act[u] && (kh[u] & kl[u]) == 0xffffffffu;
                        if (__any_sync(0xffffffffu, sp)) {
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                if (act[u] && (kh[u] & kl[u]) == 0xffffffffu) {
                                    val_add(special, ex[u]);
                                    act[u] = false;
                                }
                            }
                        }
                    }
                    uint32_t pm = 0;   // bit u: window u must queue (it collided, or the pass is sparse)
                    if (!sparse) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const KT kp = S::pack(Key<W>{((uint64_t)kh[u] << 32) | kl[u]});
                            KT c = cur[u];
                            if (act[u] && c == S::empty()) c = atomicCAS(&tkey[h[u]], S::empty(), kp);
                            const bool ok = (c == kp) | (c == S::empty());   // hit, or the claim went through
                            if (act[u] && ok) val_add(&tval[h[u]], ex[u]);
                            pm |= (act[u] && !ok) ? (1u << u) : 0u;
                        }
                    } else {
                        // probing inline would run the insert code for a fraction of the lanes: the windows of this pass are
                        // queued instead (as if their probe of slot h-1 had failed) and the queue is drained 32 wide
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            pm |= act[u] ? (1u << u) : 0u;
                            h[u] = prev_slot<W>(h[u], tsize);
                        }
                    }
                    __syncwarp();
                    bool pend[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) pend[u] = (pm >> u) & 1u;
                    // collided observations of the four windows -> queue, with one warp scan when they fit at once
                    const uint32_t c = (uint32_t)__popc(pm);
                    uint32_t incl = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= (uint32_t)d) incl += y;
                    }
                    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                    if (tot != 0u && wq.n + tot <= (uint32_t)SM_QCAP) {
                        uint32_t i = wq.n + incl - c;
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (pend[u]) {
                                wq.key[i] = S::pack(Key<W>{((uint64_t)kh[u] << 32) | kl[u]});
                                wq.hx[i] = h[u] | (ex[u] << 16);
                                i++;
                            }
                        }
                        wq.n += tot;
                        __syncwarp();
                        while (wq.n >= 32u && wq.budget > 0) {
                            queue_drain_step<W>(wq, tkey, tval, tsize, lane);
                            wq.budget--;
                        }
                        if (wq.n >= 32u) wq.n = 0;   // budget exhausted (checked by the caller)
                    } else if (tot != 0u) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const KT kp = S::pack(Key<W>{((uint64_t)kh[u] << 32) | kl[u]});
                            queue_push<W>(wq, pend[u], kp, h[u], ex[u], tkey, tval, tsize, lane);
                        }
                    }
                } else if (K64 && ulen > 48u) {
                    // ---- K == 64: the same scheme with four-word k-mers; seven packed words hold the bases of a lane's windows -------
                    const uint32_t R = (ulen + 31u) >> 5;
                    const uint32_t j0 = R * lane;
                    const uint32_t p = p0 + (j0 < ulen ? j0 : 0u);
                    const uint32_t q = rs + p;
                    const int32_t w = (int32_t)(q >> 4);
                    const uint32_t s = 2u * (q & 15u);
                    const uint32_t xm = pk[w - 1], x0 = pk[w], x1 = pk[w + 1], x2 = pk[w + 2], x3 = pk[w + 3], x4 = pk[w + 4], x5 = pk[w + 5];
                    uint32_t k3 = __funnelshift_l(x1, x0, s), k2 = __funnelshift_l(x2, x1, s), k1 = __funnelshift_l(x3, x2, s),
                             k0 = __funnelshift_l(x4, x3, s);                // bases q .. q+63, first base in the top bits of k3
                    const uint32_t tail = __funnelshift_l(x5, x4, s);        // bases q+64 .. q+79
                    uint32_t lb = __funnelshift_l(x0, xm, s) & 3u;           // base q-1
                    KT kp[4];
                    uint32_t ex[4], h[4];
                    bool act[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t nb = (tail >> (30 - 2 * u)) & 3u;     // right neighbour of window u
                        const bool has_l = u ? true : (p != 0u);
                        const bool has_r = (p + u + 1u) < nk;
                        act[u] = (uint32_t)u < R && j0 + (uint32_t)u < ulen;
                        if (MULTI) act[u] = act[u] && pass_of(k3, (uint32_t)P) == (uint32_t)pass;
                        ex[u] = (has_l ? (1u << lb) : 0u) | (has_r ? (16u << nb) : 0u);
                        const Key<2> key{((uint64_t)k1 << 32) | k0, ((uint64_t)k3 << 32) | k2};
                        h[u] = home_slot<W>(hash32(key), tsize, tshift);
                        kp[u] = KeySlot<2>::pack(key);
                        if ((k3 & k2 & k1 & k0) == 0xffffffffu && act[u]) {   // the all-ones k-mer is the EMPTY marker: side word
                            val_add(special, ex[u]);
                            act[u] = false;
                        }
                        lb = k3 >> 30;                                       // left neighbour of the next window
                        k3 = __funnelshift_l(k2, k3, 2);
                        k2 = __funnelshift_l(k1, k2, 2);
                        k1 = __funnelshift_l(k0, k1, 2);
                        k0 = (k0 << 2) | nb;
                    }
                    const bool sparse = MULTI && P > 1;
                    uint32_t pm = 0;   // bit u: window u must queue (it collided, or the pass is sparse)
                    if (!sparse) {
                        KT cur[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) cur[u] = act[u] ? tkey[h[u]] : S::empty();
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            KT c = cur[u];
                            if (act[u] && c == S::empty()) c = atomicCAS(&tkey[h[u]], S::empty(), kp[u]);
                            const bool ok = (c == kp[u]) | (c == S::empty());
                            if (act[u] && ok) val_add(&tval[h[u]], ex[u]);
                            pm |= (act[u] && !ok) ? (1u << u) : 0u;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            pm |= act[u] ? (1u << u) : 0u;
                            h[u] = prev_slot<W>(h[u], tsize);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 4; u++) queue_push<W>(wq, (pm >> u) & 1u, kp[u], h[u], ex[u], tkey, tval, tsize, lane);
                } else {
                    // ---- generic: 32 windows per pass, one per lane ------------------------------------------------------------------
                    for (uint32_t pb = 0; pb < ulen; pb += 32u) {
                        const uint32_t p = p0 + pb + lane;
                        bool pend = false;
                        KT kp = S::empty();
                        uint32_t h = 0, ex = 0;
                        if (pb + lane < ulen) {
                            const uint32_t q = rs + p;
                            Key<W> key = Key<W>::zero();
                            load_window(pk, q, K, key);
                            if (!MULTI || pass_of(key_top32(key, 2 * K), (uint32_t)P) == (uint32_t)pass) {
                                const uint32_t lb = base_of(pk, q - 1u), rb = base_of(pk, q + (uint32_t)K);
                                // Exts of iter_kmer_exts with seq_exts = Exts::empty()
                                ex = (p ? (1u << lb) : 0u) | ((p + 1u < nk) ? (16u << rb) : 0u);
                                h = home_slot<W>(hash32(key), tsize, tshift);
                                kp = S::pack(key);
                                if (MULTI && P > 1 && !(kp == S::empty())) {   // sparse pass: queue, drain 32 wide
                                    pend = true;
                                    h = prev_slot<W>(h, tsize);
                                } else {
                                    pend = first_probe(kp, h, ex, tkey[h]);
                                }
                            }
                        }
                        __syncwarp();
                        queue_push<W>(wq, pend, kp, h, ex, tkey, tval, tsize, lane);
                    }
                }
            }
            if (wq.budget <= 0) misc[3] = 1;   // never spin on a full table: the HBM path takes the group
            if (tid == 0 && pf_stage == 1) pf_advance();   // next group's byte range
            if (lane == 0) ws[WS_LEFT + warp] = wq.n;   // leftovers (< 32 per warp) are finished by the whole CTA below
        }
        if (tid == 0) {
            // pool space for whatever this table can yield (<= SM_TSIZE + 1 survivors): one global atomic per SM_POOL_CHUNK, not per group
            if (cta_pool_pos + (int64_t)(table_slots<W>() + 1) > cta_pool_end) {
                cta_pool_pos = (int64_t)atomicAdd(a.pool_cursor, (unsigned long long)SM_POOL_CHUNK);
                cta_pool_end = cta_pool_pos + (int64_t)SM_POOL_CHUNK;
            }
            misc[4] = (uint32_t)cta_pool_pos;
            misc[5] = (uint32_t)((uint64_t)cta_pool_pos >> 32);
            misc[6] = 0;   // survivors written so far
        }
        __syncthreads();
        if (total != 0 && !misc[3]) {
            // every thread takes at most one leftover observation and probes until it lands
            const uint32_t wsrc = (uint32_t)tid >> 5, isrc = (uint32_t)tid & 31u;   // thread t <- entry isrc of warp wsrc
            if (wsrc < (uint32_t)NW && isrc < ws[WS_LEFT + wsrc]) {
                const KT kp = (reinterpret_cast<KT *>(qbytes) + wsrc * SM_QCAP)[isrc];
                const uint32_t hx = (reinterpret_cast<uint32_t *>(reinterpret_cast<KT *>(qbytes) + NW * SM_QCAP) + wsrc * SM_QCAP)[isrc];
                uint32_t h = hx & 0xffffu;
                uint32_t guard = 0;
                while (true) {
                    h = next_slot<W>(h, tsize);
                    KT cur = tkey[h];
                    if (cur == S::empty()) cur = atomicCAS(&tkey[h], S::empty(), kp);
                    if (cur == kp || cur == S::empty()) {
                        val_add(&tval[h], hx >> 16);
                        break;
                    }
                    if (++guard > tsize) {
                        misc[3] = 1;
                        break;
                    }
                }
            }
        }
        __syncthreads();
        if (misc[3]) {   // too many distinct k-mers for a shared-memory table
            if (MULTI && P < SM_PMAX && 2 * K > 12) {   // start the group over with twice the passes
                P = 2 * P < SM_PMAX ? 2 * P : SM_PMAX;
                pass = 0;
                srow = r0;
                gprefix = 0;
                gobs = 0;
                continue;
            }
            defer(g);   // HBM-table path
            in_group = false;
            continue;
        }
        if (MULTI) {
            if (pass == 0) gobs += total;
            srow = s1;
            if (srow < r1) continue;   // next slab of this pass
        }

        PHASE_MARK(3);
        // ---- min-coverage prune + compaction in one pass: survivors go to the pool unsorted (k_sort_small / k_sort_big order them
        //      while building the dense output, at far higher occupancy than this kernel's 2 CTAs/SM).  A thread keeps its slots'
        //      value words in registers, a warp reserves its range of the group's pool slice with one shared atomic. ----------------
        const int64_t pb = (int64_t)(((unsigned long long)misc[5] << 32) | misc[4]);
        if (pb + (int64_t)(table_slots<W>() + 1) > a.pool_cap) {   // survivor pool exhausted: the HBM path (own pool) takes the group
            defer(g);
            in_group = false;
            continue;
        }
        auto emit = [&](int64_t pos, KT kq, uint32_t v) {
            const Key<W> k = S::unpack(kq);
            a.pool_lo[pos] = k.lo;
            if (W == 2) a.pool_hi[pos] = k.hi_word();
            a.pool_exts[pos] = (uint8_t)v;
            a.pool_counts[pos] = (uint16_t)((v >> 8) > 65535u ? 65535u : (v >> 8));
        };
        {
            constexpr int CIT = (table_slots<W>() / 4 + NT - 1) / NT;
            const uint4 *tval4 = reinterpret_cast<const uint4 *>(tval);
            uint4 vv[CIT];
            uint32_t cnt = 0;
#pragma unroll
            for (int it = 0; it < CIT; it++) {
                const uint32_t j = (uint32_t)tid + (uint32_t)it * NT;
                vv[it] = make_uint4(0u, 0u, 0u, 0u);
                if (j < tsize / 4) {
                    vv[it] = tval4[j];
                    cnt += ((vv[it].x >> 8) >= thr) + ((vv[it].y >> 8) >= thr) + ((vv[it].z >> 8) >= thr) + ((vv[it].w >> 8) >= thr);
                }
            }
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += y;
            }
            uint32_t wbase = 0;
            if (lane == 31u && incl) wbase = atomicAdd(const_cast<uint32_t *>(&misc[6]), incl);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            int64_t pos = pb + (int64_t)(wbase + incl - cnt);
            if (cnt) {
#pragma unroll
                for (int it = 0; it < CIT; it++) {
                    const uint32_t j = (uint32_t)tid + (uint32_t)it * NT;
                    const uint32_t c[4] = {vv[it].x, vv[it].y, vv[it].z, vv[it].w};
                    if (((c[0] >> 8) >= thr) | ((c[1] >> 8) >= thr) | ((c[2] >> 8) >= thr) | ((c[3] >> 8) >= thr)) {
#pragma unroll
                        for (int t = 0; t < 4; t++)
                            if ((c[t] >> 8) >= thr) emit(pos++, tkey[4 * j + t], c[t]);
                    }
                }
            }
            if (tid == 0) {
                const uint32_t sv = misc[2];
                if ((sv >> 8) >= thr) emit(pb + (int64_t)atomicAdd(const_cast<uint32_t *>(&misc[6]), 1u), S::empty(), sv);
            }
        }
        __syncthreads();
        const uint32_t n_out = misc[6];
        if (tid == 0) cta_pool_pos += n_out;
        if (!MULTI && n_out == 0) {
            cta_obs += total;
            continue;
        }
        if (n_out > (uint32_t)SM_TSIZE) {   // (a full table plus the side word) more than the ordering kernels hold: HBM path
            defer(g);
            in_group = false;
            continue;
        }
        PHASE_MARK(4);
        if (!MULTI) {
            if (tid == 0) {
                a.nsurv[g] = (int64_t)n_out;
                a.pool_start[g] = pb;
            }
            cta_obs += total;
        } else {
            if (tid == 0) a.segs[item * SM_PMAX + pass] = SegRec{pb, (int32_t)n_out, (int32_t)gprefix};
            gprefix += n_out;
            pass++;
            srow = r0;
            if (pass == P) {   // group complete; its pool_start stays -1, the segments carry the survivors
                if (tid == 0) {
                    a.seg_count[item] = pass;
                    a.nsurv[g] = (int64_t)gprefix;
                }
                cta_obs += gobs;
                in_group = false;
            }
        }
        PHASE_MARK(5);
    }
    if (tid == 0 && cta_obs) atomicAdd(a.n_obs, cta_obs);
}

// upper bound of the observations of a column: sum over rows of max(0, len-K+1) (validity ignored)
__global__ void k_obs_upper_bound(ColView col, int K, unsigned long long *out)
{
    unsigned long long s = 0;
    for (int64_t r = col.row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < col.row_begin + col.n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        int64_t len = col_off(col, r + 1) - col_off(col, r);
        if (len >= K) s += (unsigned long long)(len - K + 1);
    }
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// -----------------------------------------------------------------------------------------------------------------------------
// Ordering + dense output.  filter_kmers emits every group's survivors in ascending k-mer order; the counting
// kernel leaves them unsorted in the pool.  k_sort_small orders one group per CTA pass in 32 KB of shared memory (counting
// sort on the top 10 k-mer bits + insertion sort inside the tiny buckets) and writes the dense, group-ordered arrays;
// groups with more than sort_small_max() survivors or a crowded bucket go to k_sort_big (bitonic sort in shared memory).
// -----------------------------------------------------------------------------------------------------------------------------
constexpr int SORT_NT = 128;
template <int W>
constexpr int sort_small_max() { return W == 2 ? 1408 : 1536; }   // 12 (W=1) / 11 (W=2) survivors per thread in registers
constexpr int SORT_BIG_NT = 256;
constexpr int SORT_BIG_MAX = SM_TSIZE;   // a shared-memory table never yields more survivors than it has slots

struct SortArgs {
    int64_t n_groups;
    const int64_t *pool_start;     // -1: group not in the pool (deferred to the HBM path)
    const int64_t *kmer_offsets;   // dense output offsets
    const uint64_t *p_lo, *p_hi;
    const uint8_t *p_exts;
    const uint16_t *p_counts;
    uint64_t *o_lo, *o_hi;
    uint8_t *o_exts;
    uint16_t *o_counts;
    int32_t *big_ids;              // groups left to k_sort_big
    unsigned long long *n_big;
    int kbits;
    // segments of multi-pass groups (k_sort_big<W, true>): item i = pass i % SM_PMAX of list entry i / SM_PMAX
    const SegRec *segs;
    const int32_t *seg_count;
    const int32_t *seg_list;
    int seg_stride;                // segment slots per list entry (SM_PMAX for multi-pass groups, PT_PMAX for partitioned ones)
    uint32_t seg_total, seg_base;  // seg_total != 0: slot p is range seg_base + p of seg_total equal ranges of the k-mer space (merge of owned ranges)
};

// Counting sort of one work item per CTA pass: survivors are read once into registers, counted per bucket (top 10 k-mer bits),
// scattered from the registers into shared memory in bucket order, and every survivor then ranks itself inside its tiny bucket
// and is written straight to its final position (4 block barriers, no index indirection).  SEG = false: work items are groups.  SEG = true: segments of multi-pass
// groups (item i = pass i % SM_PMAX of list entry i / SM_PMAX); the k-mers of a segment lie in 1/P of the k-mer space, so the
// buckets subdivide that range.  LIST = true: the items are those named in in_list (what a smaller instantiation passed on).
// Two instantiations: CAP = sort_small_max() with 128 threads (23 KB and 64 registers, 8 CTAs/SM: the headline shape, ~1000 survivors per
// group) and CAP = 8192 with 512 threads (two CTAs/SM: min_coverage <= 1 keeps ~4500 k-mers per group; the bitonic kernel that
// used to take those groups was 10x slower).  Items that do not fit, or have a crowded bucket, go to the output list.
template <int W, bool SEG, int CAP, int NT, bool LIST>
__global__ void __launch_bounds__(NT, (NT == 128 && W == 1 ? 8 : 1)) k_sort_small(SortArgs a, int64_t n_items, const int32_t *in_list)
{
    constexpr int EPT = (CAP + NT - 1) / NT;   // survivors a thread keeps in registers between the passes
    extern __shared__ __align__(16) uint8_t sort_raw[];
    uint64_t *s_lo = reinterpret_cast<uint64_t *>(sort_raw);                 // bucket-ordered keys
    uint64_t *s_hi = s_lo + CAP;                                             // only touched when W == 2
    uint32_t *s_val = reinterpret_cast<uint32_t *>(s_lo + (W == 2 ? 2 : 1) * CAP);   // exts | count << 8, bucket-ordered
    uint32_t *s_hist = s_val + CAP;
    uint32_t *s_ws = s_hist + SM_SORT_BUCKETS;
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int64_t it = LIST ? (int64_t)in_list[w] : w;
        int64_t ps, o;
        int n;
        const int kb = a.kbits;
        uint32_t seg_np = 1, seg_first = 0;   // SEG: passes of the group, first top-32 value of this segment's k-mer range
        if (SEG) {
            const int64_t e = it / a.seg_stride;
            const int np = a.seg_count[e];
            if ((int)(it % a.seg_stride) >= np) continue;
            const SegRec r = a.segs[it];
            ps = r.pool_start;
            o = a.kmer_offsets[a.seg_list[e]] + r.prefix;
            n = r.n;
            seg_np = a.seg_total ? a.seg_total : (uint32_t)np;
            seg_first = pass_first((uint32_t)(it % a.seg_stride) + a.seg_base, seg_np);
        } else {
            ps = a.pool_start[it];
            if (ps < 0) continue;
            o = a.kmer_offsets[it];
            n = (int)(a.kmer_offsets[it + 1] - o);
        }
        if (n == 0) continue;
        if (n > CAP) {
            if (tid == 0) a.big_ids[atomicAdd(a.n_big, 1ull)] = (int32_t)it;
            continue;
        }
        auto bucket_of = [&](uint64_t lo, uint64_t hi) -> uint32_t {
            if (SEG) {   // position inside the segment's k-mer range, scaled to the bucket count (monotone in the k-mer)
                const uint32_t t = W == 2 ? key_top32(Key<2>{lo, hi}, kb) : key_top32(Key<1>{lo}, kb);
                const uint64_t b = ((uint64_t)(t - seg_first) * seg_np) >> (32 - SM_SORT_BITS);
                return b < (uint64_t)SM_SORT_BUCKETS ? (uint32_t)b : (uint32_t)(SM_SORT_BUCKETS - 1);
            }
            if (W == 2) return sort_bucket<2>(Key<2>{lo, hi}, kb);
            return sort_bucket<1>(Key<1>{lo}, kb);
        };
        // ---- survivors of the item -> registers; bucket histogram ---------------------------------------------------------------
        uint64_t klo[EPT], khi[W == 2 ? EPT : 1];
        uint32_t kv[EPT];
        __syncthreads();   // previous item's shared data no longer in use
        if (tid == 0) s_flag = 0;
        for (int j = tid; j < SM_SORT_BUCKETS; j += NT) s_hist[j] = 0;
#pragma unroll
        for (int t = 0; t < EPT; t++) {
            const int j = tid + t * NT;
            if (j < n) {
                klo[t] = a.p_lo[ps + j];
                if (W == 2) khi[t] = a.p_hi[ps + j];
                kv[t] = (uint32_t)a.p_exts[ps + j] | ((uint32_t)a.p_counts[ps + j] << 8);
            }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < EPT; t++)
            if (tid + t * NT < n) atomicAdd(&s_hist[bucket_of(klo[t], W == 2 ? khi[t] : 0ull)], 1u);
        __syncthreads();
        {   // exclusive scan of the bucket sizes (PER per thread)
            constexpr int PER = SM_SORT_BUCKETS / NT;
            uint32_t loc[PER], sum = 0, mx = 0;
#pragma unroll
            for (int t = 0; t < PER; t++) {
                loc[t] = s_hist[tid * PER + t];
                sum += loc[t];
                mx = loc[t] > mx ? loc[t] : mx;
            }
            if (mx > (uint32_t)SM_MAX_BUCKET_WARP) s_flag = 2;
            uint32_t x = sum;
            const unsigned lane = tid & 31u, warp = tid >> 5;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (unsigned)d) x += y;
            }
            if (lane == 31) s_ws[warp] = x;
            __syncthreads();
            uint32_t base = 0;
            for (unsigned ww = 0; ww < warp; ww++) base += s_ws[ww];
            uint32_t e0 = base + x - sum;
#pragma unroll
            for (int t = 0; t < PER; t++) {
                s_hist[tid * PER + t] = e0;
                e0 += loc[t];
            }
        }
        __syncthreads();
        if (s_flag == 2) {   // a very crowded bucket (low-complexity k-mers): bitonic kernel
            if (tid == 0) a.big_ids[atomicAdd(a.n_big, 1ull)] = (int32_t)it;
            continue;
        }
        // ---- registers -> shared memory in bucket order (arrival order inside a bucket) ----------------------------------------------
#pragma unroll
        for (int t = 0; t < EPT; t++) {
            if (tid + t * NT < n) {
                const uint32_t pos = atomicAdd(&s_hist[bucket_of(klo[t], W == 2 ? khi[t] : 0ull)], 1u);
                s_lo[pos] = klo[t];
                if (W == 2) s_hi[pos] = khi[t];
                s_val[pos] = kv[t];
            }
        }
        __syncthreads();
        // s_hist[b] is now the END of bucket b.
        // Deep UMIs: every true k-mer drags ~3K single-error variants along that share its leading bases, i.e. its bucket (dozens to
        // hundreds of members).  Such a bucket is ranked like the tiny ones -- all pairs: consecutive threads hold members of the
        // same bucket, so the inner loop's shared-memory reads are broadcasts.  (An earlier version had one warp order each crowded
        // bucket by odd-even transposition: `size` dependent phases per bucket; 39 ms of BASELINE configs[2]'s 205 ms.)
        // Every survivor ranks itself inside its (tiny) bucket -- final position = bucket start + number of bucket members with a
        // smaller k-mer (k-mers are distinct) -- and goes straight to the dense output; members of a warp-ordered bucket are in place.
        for (int j = tid; j < n; j += NT) {
            const uint64_t lj = s_lo[j], hj = (W == 2) ? s_hi[j] : 0ull;
            const uint32_t b = bucket_of(lj, hj);
            const uint32_t st = b ? s_hist[b - 1] : 0u, en = s_hist[b];
            uint32_t r = st;
            if (en - st > 1u) {
                for (uint32_t i = st; i < en; i++) {
                    const uint64_t li = s_lo[i], hi_i = (W == 2) ? s_hi[i] : 0ull;
                    r += (hi_i < hj || (hi_i == hj && li < lj)) ? 1u : 0u;
                }
            }
            a.o_lo[o + r] = lj;
            if (W == 2) a.o_hi[o + r] = hj;
            const uint32_t v = s_val[j];
            a.o_exts[o + r] = (uint8_t)v;
            a.o_counts[o + r] = (uint16_t)(v >> 8);
        }
    }
}
template <int W, int CAP>
constexpr size_t sort_count_smem() { return (size_t)CAP * (8 * W + 4) + SM_SORT_BUCKETS * 4 + 32 * 4; }
constexpr int SORT_MID_NT = 512;
constexpr int SORT_MID_CAP = SM_TSIZE;   // every single-table group fits; 100 KB of shared memory for 8-byte keys: two CTAs per SM

// SEG = false: one work item per group listed in big_ids.  SEG = true: big_ids lists segments of multi-pass groups; the
// segments of a group hold ascending k-mer ranges, so ordering each one in place orders the group.
template <int W, bool SEG>
__global__ void __launch_bounds__(SORT_BIG_NT) k_sort_big(SortArgs a, int64_t n_big)
{
    extern __shared__ __align__(16) uint8_t sraw[];
    uint64_t *s_lo = reinterpret_cast<uint64_t *>(sraw);
    uint64_t *s_hi = s_lo + SORT_BIG_MAX;                                    // only touched when W == 2
    uint32_t *s_val = reinterpret_cast<uint32_t *>(s_lo + (W == 2 ? 2 : 1) * SORT_BIG_MAX);
    const int tid = threadIdx.x;
    for (int64_t i = blockIdx.x; i < n_big; i += gridDim.x) {
        int64_t ps, o;
        int n;
        if (SEG) {
            const int64_t it = a.big_ids[i];
            const SegRec r = a.segs[it];
            ps = r.pool_start;
            o = a.kmer_offsets[a.seg_list[it / a.seg_stride]] + r.prefix;
            n = r.n;
        } else {
            const int64_t g = a.big_ids[i];
            ps = a.pool_start[g];
            o = a.kmer_offsets[g];
            n = (int)(a.kmer_offsets[g + 1] - o);
        }
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        __syncthreads();
        for (int j = tid; j < np2; j += SORT_BIG_NT) {
            const bool in = j < n;
            s_lo[j] = in ? a.p_lo[ps + j] : ~0ull;
            if (W == 2) s_hi[j] = in ? a.p_hi[ps + j] : ~0ull;
            // padding sorts after every real entry: real k-mers equal to the all-ones key keep val != 0xffffffff
            s_val[j] = in ? ((uint32_t)a.p_exts[ps + j] | ((uint32_t)a.p_counts[ps + j] << 8)) : 0xffffffffu;
        }
        __syncthreads();
        for (int k = 2; k <= np2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < np2; t += SORT_BIG_NT) {
                    const int p = t ^ j;
                    if (p > t) {
                        const bool up = (t & k) == 0;
                        const uint64_t la = s_lo[t], lb = s_lo[p];
                        const uint64_t ha = (W == 2) ? s_hi[t] : 0ull, hb = (W == 2) ? s_hi[p] : 0ull;
                        const uint32_t va = s_val[t], vb = s_val[p];
                        // order by (hi, lo, padding flag)
                        const bool a_gt_b = ha > hb || (ha == hb && (la > lb || (la == lb && va == 0xffffffffu && vb != 0xffffffffu)));
                        if (a_gt_b == up) {
                            s_lo[t] = lb; s_lo[p] = la;
                            if (W == 2) { s_hi[t] = hb; s_hi[p] = ha; }
                            s_val[t] = vb; s_val[p] = va;
                        }
                    }
                }
                __syncthreads();
            }
        }
        for (int j = tid; j < n; j += SORT_BIG_NT) {
            a.o_lo[o + j] = s_lo[j];
            if (W == 2) a.o_hi[o + j] = s_hi[j];
            const uint32_t v = s_val[j];
            a.o_exts[o + j] = (uint8_t)v;
            a.o_counts[o + j] = (uint16_t)(v >> 8);
        }
    }
}

template <int W>
constexpr size_t sort_big_smem() { return (size_t)SORT_BIG_MAX * (8 * W + 4); }

// host launcher: pool (unsorted, completion order) -> dense, group-ordered, k-mer-ascending arrays.  Three tiers per kind of
// work item (groups; segments of multi-pass groups): small counting sort -> mid counting sort -> bitonic sort.
template <int W>
inline void run_sort_gather_w(cudaStream_t st, uint8_t *peek_slot, int sm_count, SortArgs a, int64_t n_groups, int64_t n_seg_groups,
                              int64_t *launches)
{
    constexpr int SMALL = sort_small_max<W>();
    const int64_t n_seg_items = n_seg_groups * (int64_t)a.seg_stride;
    // lists: [0] groups small->mid, [1] segments small->mid, [2] groups mid->bitonic, [3] segments mid->bitonic
    DevBuf ctr(32, st), l0((size_t)n_groups * 4 + 16, st), l1((size_t)n_seg_items * 4 + 16, st);
    CK(cudaMemsetAsync(ctr.p, 0, 32, st));
    unsigned long long *d_n = ctr.as<unsigned long long>();
    SortArgs g = a, s = a;
    g.big_ids = l0.as<int32_t>(); g.n_big = d_n;
    s.big_ids = l1.as<int32_t>(); s.n_big = d_n + 1;
    if (n_groups > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>(n_groups, (int64_t)sm_count * 16);
        k_sort_small<W, false, SMALL, SORT_NT, false><<<grid, SORT_NT, sort_count_smem<W, SMALL>(), st>>>(g, n_groups, nullptr);
        if (launches) *launches += 1;
    }
    if (n_seg_items > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>(n_seg_items, (int64_t)sm_count * 16);
        k_sort_small<W, true, SMALL, SORT_NT, false><<<grid, SORT_NT, sort_count_smem<W, SMALL>(), st>>>(s, n_seg_items, nullptr);
        if (launches) *launches += 1;
    }
    unsigned long long n_mid[2] = {0, 0};
    Peek pk(peek_slot, st, launches);
    pk.add(d_n, 16, n_mid);
    pk.sync();
    if (!n_mid[0] && !n_mid[1]) return;
    DevBuf l2((size_t)n_mid[0] * 4 + 16, st), l3((size_t)n_mid[1] * 4 + 16, st);
    SortArgs g2 = g, s2 = s;
    g2.big_ids = l2.as<int32_t>(); g2.n_big = d_n + 2;
    s2.big_ids = l3.as<int32_t>(); s2.n_big = d_n + 3;
    if (n_mid[0]) {
        const unsigned grid = (unsigned)std::min<int64_t>((int64_t)n_mid[0], (int64_t)sm_count);
        k_sort_small<W, false, SORT_MID_CAP, SORT_MID_NT, true>
            <<<grid, SORT_MID_NT, sort_count_smem<W, SORT_MID_CAP>(), st>>>(g2, (int64_t)n_mid[0], g.big_ids);
        if (launches) *launches += 1;
    }
    if (n_mid[1]) {
        const unsigned grid = (unsigned)std::min<int64_t>((int64_t)n_mid[1], (int64_t)sm_count);
        k_sort_small<W, true, SORT_MID_CAP, SORT_MID_NT, true>
            <<<grid, SORT_MID_NT, sort_count_smem<W, SORT_MID_CAP>(), st>>>(s2, (int64_t)n_mid[1], s.big_ids);
        if (launches) *launches += 1;
    }
    unsigned long long n_big[2] = {0, 0};
    pk.add(d_n + 2, 16, n_big);
    pk.sync();
    if (n_big[0]) {
        const unsigned grid = (unsigned)std::min<int64_t>((int64_t)n_big[0], (int64_t)sm_count * 2);
        k_sort_big<W, false><<<grid, SORT_BIG_NT, sort_big_smem<W>(), st>>>(g2, (int64_t)n_big[0]);
        if (launches) *launches += 1;
    }
    if (n_big[1]) {
        const unsigned grid = (unsigned)std::min<int64_t>((int64_t)n_big[1], (int64_t)sm_count * 2);
        k_sort_big<W, true><<<grid, SORT_BIG_NT, sort_big_smem<W>(), st>>>(s2, (int64_t)n_big[1]);
        if (launches) *launches += 1;
    }
}

inline void run_sort_gather(cudaStream_t st, uint8_t *peek_slot, int sm_count, int K, int64_t n_groups, const int64_t *pool_start, const int64_t *kmer_offsets,
                            const uint64_t *p_lo, const uint64_t *p_hi, const uint8_t *p_exts, const uint16_t *p_counts, uint64_t *o_lo,
                            uint64_t *o_hi, uint8_t *o_exts, uint16_t *o_counts, int64_t *launches, const SegRec *segs = nullptr,
                            const int32_t *seg_count = nullptr, const int32_t *seg_list = nullptr, int64_t n_seg_groups = 0,
                            int seg_stride = SM_PMAX, bool groups_too = true, uint32_t seg_total = 0, uint32_t seg_base = 0)
{
    SortArgs a{};
    a.n_groups = n_groups;
    a.pool_start = pool_start;
    a.kmer_offsets = kmer_offsets;
    a.p_lo = p_lo; a.p_hi = p_hi; a.p_exts = p_exts; a.p_counts = p_counts;
    a.o_lo = o_lo; a.o_hi = o_hi; a.o_exts = o_exts; a.o_counts = o_counts;
    a.kbits = 2 * K;
    a.segs = segs;
    a.seg_count = seg_count;
    a.seg_list = seg_list;
    a.seg_stride = seg_stride;
    a.seg_total = seg_total;
    a.seg_base = seg_base;
    if (K > 32) run_sort_gather_w<2>(st, peek_slot, sm_count, a, groups_too ? n_groups : 0, n_seg_groups, launches);
    else run_sort_gather_w<1>(st, peek_slot, sm_count, a, groups_too ? n_groups : 0, n_seg_groups, launches);
}

}  // namespace rogtk
#include "kernels_walk.cuh"
namespace rogtk {

// -----------------------------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------------------------
// knobs read once per context (rogtk_gpu_create), never per call
struct SmemTuning {
    int walk_rl_bits = 0;          // 0 (default): k_smem_count counts the first tier.  3..6: the experimental link-walk kernel
                                   // k_walk_count with ranges of 1 << bits windows per lane (measured slower)
    int64_t debug_pool_cap = 0;    // tests: force a tiny survivor pool
};

struct SmemResult {
    const int32_t *deferred_ids = nullptr;
    int64_t n_deferred = 0;
    int64_t n_pool_groups = 0;
    const int64_t *pool_start = nullptr;
    const uint64_t *pool_lo = nullptr, *pool_hi = nullptr;
    const uint8_t *pool_exts = nullptr;
    const uint16_t *pool_counts = nullptr;
    unsigned long long *pool_cursor = nullptr;   // the partition tier appends its segments to the same pool
    int64_t pool_cap = 0;
    std::vector<DevBuf> keep;
    float ms_kernel = 0.f;
    // multi-pass groups: list entries, their segments (SM_PMAX per entry) and passes used
    int64_t n_seg_groups = 0;
    const int32_t *seg_list = nullptr;
    const SegRec *segs = nullptr;
    const int32_t *seg_count = nullptr;
};

inline void smem_path_configure()
{
    auto cfg = [](auto kernel, size_t bytes) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    };
    cfg(k_smem_count<1, 32, false>, smem_bytes<1>());
    cfg(k_smem_count<1, 0, false>, smem_bytes<1>());
    cfg(k_smem_count<2, 0, false>, smem_bytes<2>());
    cfg(k_smem_count<2, 64, false>, smem_bytes<2>());
    cfg(k_smem_count<1, 32, true>, smem_bytes<1>());
    cfg(k_smem_count<1, 0, true>, smem_bytes<1>());
    cfg(k_smem_count<2, 0, true>, smem_bytes<2>());
    cfg(k_smem_count<2, 64, true>, smem_bytes<2>());
    cfg(k_walk_count<1, false>, walk_smem_bytes<1>());
    cfg(k_walk_count<2, false>, walk_smem_bytes<2>());
    cfg(k_walk_count<1, true>, walk_smem_bytes<1>());
    cfg(k_walk_count<2, true>, walk_smem_bytes<2>());
    CK(cudaFuncSetAttribute(k_sort_small<1, false, SORT_MID_CAP, SORT_MID_NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sort_count_smem<1, SORT_MID_CAP>()));
    CK(cudaFuncSetAttribute(k_sort_small<1, true, SORT_MID_CAP, SORT_MID_NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sort_count_smem<1, SORT_MID_CAP>()));
    CK(cudaFuncSetAttribute(k_sort_small<2, false, SORT_MID_CAP, SORT_MID_NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sort_count_smem<2, SORT_MID_CAP>()));
    CK(cudaFuncSetAttribute(k_sort_small<2, true, SORT_MID_CAP, SORT_MID_NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sort_count_smem<2, SORT_MID_CAP>()));
    CK(cudaFuncSetAttribute(k_sort_big<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_big_smem<1>()));
    CK(cudaFuncSetAttribute(k_sort_big<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_big_smem<2>()));
    CK(cudaFuncSetAttribute(k_sort_big<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_big_smem<1>()));
    CK(cudaFuncSetAttribute(k_sort_big<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_big_smem<2>()));
}

inline void run_smem_path(cudaStream_t st, uint8_t *peek_slot, int sm_count, size_t workspace_bytes, const ColView &col, const uint8_t *data_begin,
                          const uint8_t *data_end,
                          const int64_t *d_group_offsets, int64_t n_groups, int K, uint32_t min_cov, unsigned long long *d_nobs,
                          int64_t *d_nsurv, SmemResult &res, int64_t *launches, cudaEvent_t ev_a, cudaEvent_t ev_b,
                          unsigned long long *phase_out /* host [16] or nullptr */, bool multipass = true, const SmemTuning &tune = SmemTuning())
{
    const int W = K > 32 ? 2 : 1;
    const bool walk = tune.walk_rl_bits >= WK_MIN_RL_BITS;
    // CTAs the first-tier kernel runs at once (each keeps one chunk of the survivor pool open)
    const int64_t first_ctas = (int64_t)sm_count * (walk ? walk_min_blocks_rt(W) : (W == 1 ? 2 : 1));
    DevBuf ctr(256, st);
    CK(cudaMemsetAsync(ctr.p, 0, 256, st));
    unsigned long long *d_ctr = ctr.as<unsigned long long>();   // [0] ticket [1] pool cursor [2] n_deferred [3] obs upper bound
    k_obs_upper_bound<<<sm_count * 4, 256, 0, st>>>(col, K, d_ctr + 3);
    unsigned long long ub = 0;
    Peek pk(peek_slot, st, launches);
    pk.add(d_ctr + 3, 8, &ub);
    pk.sync();
    // a k-mer needs min_cov observations to survive, so survivors <= observations / min_cov
    const size_t entry = (W == 2 ? 19 : 11);
    int64_t pool_cap = (int64_t)(ub / (min_cov > 1 ? min_cov : 1)) + 1;
    pool_cap = std::min<int64_t>(pool_cap, (int64_t)(workspace_bytes / 2 / entry));
    // per-CTA chunking leaves unused tails: one open chunk per running CTA, up to SM_TSIZE + 1 entries at every chunk switch
    pool_cap += pool_cap / 4 + (std::min<int64_t>(n_groups, first_ctas) * 2 + (int64_t)sm_count * 2) * SM_POOL_CHUNK;
    pool_cap = std::max<int64_t>(pool_cap, 1024);
    if (tune.debug_pool_cap) pool_cap = std::max<int64_t>(1024, tune.debug_pool_cap);   // tests: exhaust the pool
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
    a.data_begin = data_begin;
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
    a.phase_cycles = phase_out ? d_ctr + 8 : nullptr;
    CK(cudaEventRecord(ev_a, st));
    if (walk) {
        const unsigned grid = (unsigned)std::min<int64_t>(n_groups, first_ctas);
        if (phase_out) {   // diagnostics build of the same kernel: phase cycles + event counts
            if (W == 1) k_walk_count<1, true><<<grid, WK_NT, walk_smem_bytes<1>(), st>>>(a, tune.walk_rl_bits);
            else k_walk_count<2, true><<<grid, WK_NT, walk_smem_bytes<2>(), st>>>(a, tune.walk_rl_bits);
        } else {
            if (W == 1) k_walk_count<1, false><<<grid, WK_NT, walk_smem_bytes<1>(), st>>>(a, tune.walk_rl_bits);
            else k_walk_count<2, false><<<grid, WK_NT, walk_smem_bytes<2>(), st>>>(a, tune.walk_rl_bits);
        }
    } else if (W == 1) {
        const unsigned grid = (unsigned)std::min<int64_t>(n_groups, (int64_t)sm_count * 2);
        if (K == 32) k_smem_count<1, 32, false><<<grid, sm_nt<1>(), smem_bytes<1>(), st>>>(a);
        else k_smem_count<1, 0, false><<<grid, sm_nt<1>(), smem_bytes<1>(), st>>>(a);
    } else {
        const unsigned grid = (unsigned)std::min<int64_t>(n_groups, (int64_t)sm_count);
        if (K == 64) k_smem_count<2, 64, false><<<grid, sm_nt<2>(), smem_bytes<2>(), st>>>(a);
        else k_smem_count<2, 0, false><<<grid, sm_nt<2>(), smem_bytes<2>(), st>>>(a);
    }
    CK(cudaEventRecord(ev_b, st));
    if (launches) *launches += 2;
    unsigned long long h_ctr[3];
    pk.add(d_ctr, 24, h_ctr);
    if (phase_out) pk.add(d_ctr + 8, 128, phase_out);
    pk.sync();
    CK(cudaEventElapsedTime(&res.ms_kernel, ev_a, ev_b));
    res.n_deferred = (int64_t)h_ctr[2];
    res.deferred_ids = d_ids;
    if (res.n_deferred > 0 && multipass) {
        // second launch: the deferred groups, several passes over slabs of their reads; what it cannot finish either
        // (very large groups, single reads longer than a slab) is listed again for the HBM-table path
        const int64_t n_a = res.n_deferred;
        res.keep.emplace_back((size_t)n_a * 4, st);
        int32_t *d_ids2 = res.keep.back().as<int32_t>();
        res.keep.emplace_back((size_t)n_a * SM_PMAX * sizeof(SegRec), st);
        SegRec *d_segs = res.keep.back().as<SegRec>();
        res.keep.emplace_back((size_t)n_a * 4, st);
        int32_t *d_segc = res.keep.back().as<int32_t>();
        CK(cudaMemsetAsync(d_segc, 0, (size_t)n_a * 4, st));
        a.group_list = d_ids;
        a.n_list = n_a;
        a.segs = d_segs;
        a.seg_count = d_segc;
        a.ticket = d_ctr + 4;
        a.deferred_ids = d_ids2;
        a.n_deferred = d_ctr + 5;
        a.phase_cycles = nullptr;
        if (W == 1) {
            const unsigned grid = (unsigned)std::min<int64_t>(n_a, (int64_t)sm_count * 2);
            if (K == 32) k_smem_count<1, 32, true><<<grid, sm_nt<1>(), smem_bytes<1>(), st>>>(a);
            else k_smem_count<1, 0, true><<<grid, sm_nt<1>(), smem_bytes<1>(), st>>>(a);
        } else {
            const unsigned grid = (unsigned)std::min<int64_t>(n_a, (int64_t)sm_count);
            if (K == 64) k_smem_count<2, 64, true><<<grid, sm_nt<2>(), smem_bytes<2>(), st>>>(a);
            else k_smem_count<2, 0, true><<<grid, sm_nt<2>(), smem_bytes<2>(), st>>>(a);
        }
        if (launches) *launches += 1;
        unsigned long long n_b = 0;
        pk.add(d_ctr + 5, 8, &n_b);
        pk.sync();
        res.n_seg_groups = n_a;
        res.seg_list = d_ids;
        res.segs = d_segs;
        res.seg_count = d_segc;
        res.deferred_ids = d_ids2;
        res.n_deferred = (int64_t)n_b;
    }
    res.n_pool_groups = n_groups - res.n_deferred;
    res.pool_start = p_start;
    res.pool_lo = p_lo;
    res.pool_hi = p_hi;
    res.pool_exts = p_exts;
    res.pool_counts = p_counts;
    res.pool_cursor = d_ctr + 1;
    res.pool_cap = pool_cap;
    res.keep.push_back(std::move(ctr));
}

}  // namespace rogtk
