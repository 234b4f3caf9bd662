// kernels_table.cuh -- HBM-resident per-group open-addressing tables: the spill path.
//
// Groups whose distinct k-mers do not fit a shared-memory table (kernels_smem.cuh), or every group when
// ROGTK_FORCE_SPILL is set, are counted here.  Same semantics as debruijn::filter::filter_kmers + CountFilter at
// reference src/fracture.rs:110-116: every window of every valid read is one observation; per distinct k-mer
// count saturates at u16 and exts are OR-ed; survivors (count >= min_coverage) are emitted as records, sorted by
// (group, k-mer) afterwards.
#pragma once
#include "common.cuh"

namespace rogtk {

template <int W>
struct KeySlot;
template <>
struct KeySlot<1> {
    using T = unsigned long long;
    __device__ __forceinline__ static T empty() { return ~0ull; }
    __device__ __forceinline__ static T pack(const Key<1> &k) { return k.lo; }
    __device__ __forceinline__ static Key<1> unpack(T s) { return Key<1>{(uint64_t)s}; }
};
template <>
struct KeySlot<2> {
    using T = unsigned __int128;
    __device__ __forceinline__ static T empty() { return ~(unsigned __int128)0; }
    __device__ __forceinline__ static T pack(const Key<2> &k) { return ((unsigned __int128)k.hi << 64) | k.lo; }
    __device__ __forceinline__ static Key<2> unpack(T s) { return Key<2>{(uint64_t)s, (uint64_t)(s >> 64)}; }
};

// value word of the HBM tables: bits 0..23 observation count (clamped later to u16), bits 24..31 exts
constexpr uint32_t TV_COUNT_MASK = 0x00FFFFFFu;
constexpr uint32_t TV_COUNT_STOP = 0x00800000u;

struct SpillRec {       // one survivor, sortable by (group rank, k-mer)
    uint64_t hi, lo;
    uint32_t grank;     // rank of the group inside the spill list (ascending group index)
    uint32_t val;       // count (low 16) | exts << 16
};
struct SpillRecLess {
    __device__ __forceinline__ bool operator()(const SpillRec &a, const SpillRec &b) const
    {
        if (a.grank != b.grank) return a.grank < b.grank;
        if (a.hi != b.hi) return a.hi < b.hi;
        return a.lo < b.lo;
    }
};

struct SpillArgs {
    ColView col;
    const int64_t *group_offsets;  // all groups
    const int32_t *ids;            // spill list (ascending group indices) or nullptr = identity
    int64_t n_list;                // groups in this wave
    int64_t list_base;             // rank of the first group of this wave inside the whole list
    const int64_t *row_off;        // [n_list+1] exclusive scan of rows per listed group (wave-local)
    const int64_t *slot_off;       // [n_list+1] exclusive scan of table capacities (wave-local)
    uint32_t *special;             // [n_list] value word of the all-ones k-mer (the table's EMPTY sentinel)
    void *keys;                    // KeySlot<W>::T [slot_off[n_list]]
    uint32_t *vals;                // [slot_off[n_list]]
    int K;
    uint32_t min_cov;
    unsigned long long *n_obs;     // global accumulator
    // extraction
    SpillRec *pool;
    int64_t pool_cap;
    int64_t *nsurv;                // [n_groups] survivors per group (global index)
    int count_only;                // pass 1: only count survivors per warp
    int64_t *warp_count;           // [n_extract_warps] pass 1 output
    const int64_t *warp_base;      // [n_extract_warps + 1] exclusive scan of warp_count (pass 2 input)
};

// per listed group: rows and table capacity (power of two >= 1.5x the observation upper bound).  A warp sums the rows of
// an entry; gridDim.y warps-worth of blocks share the rows of very large entries (one global-spectrum group has every row).
__global__ void k_spill_plan_obs(ColView col, const int64_t *group_offsets, const int32_t *ids, int64_t n_list,
                                 int64_t list_base, int K, unsigned long long *ub /* [n_list], zeroed */)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n_list; i += n_warps) {
        const int64_t g = ids ? ids[list_base + i] : list_base + i;
        const int64_t r0 = group_offsets[g], r1 = group_offsets[g + 1];
        unsigned long long s = 0;
        for (int64_t r = r0 + (int64_t)blockIdx.y * 32 + lane; r < r1; r += (int64_t)gridDim.y * 32) {
            const int64_t len = col_off(col, r + 1) - col_off(col, r);
            if (len >= K) s += (unsigned long long)(len - K + 1);
        }
        for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0 && s) atomicAdd(ub + i, s);
    }
}
__global__ void k_spill_plan_caps(const int64_t *group_offsets, const int32_t *ids, int64_t n_list, int64_t list_base,
                                  const unsigned long long *ub, int64_t *rows, int64_t *caps)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    const int64_t g = ids ? ids[list_base + i] : list_base + i;
    const int64_t u = (int64_t)ub[i];
    const int64_t want = u + (u >> 1) + 8;
    // at least 32 slots: every table then starts at a multiple of 32, so the 32 slots a warp extracts per step belong to one
    // group and the survivor records come out grouped by listed group (what the segmented sort relies on)
    int64_t cap = 32;
    while (cap < want) cap <<= 1;
    rows[i] = group_offsets[g + 1] - group_offsets[g];
    caps[i] = cap;
}

template <int W>
__device__ __forceinline__ void table_insert_global(typename KeySlot<W>::T *keys, uint32_t *vals, uint64_t mask,
                                                    const Key<W> &key, uint32_t exts, uint32_t *special)
{
    using S = KeySlot<W>;
    const typename S::T kp = S::pack(key);
    uint32_t *vp;
    if (kp == S::empty()) {
        vp = special;
    } else {
        uint64_t h = key.hash() & mask;
        while (true) {
            typename S::T cur = keys[h];
            if (cur == kp) break;
            if (cur == S::empty()) {
                typename S::T prev = atomicCAS(&keys[h], S::empty(), kp);
                if (prev == S::empty() || prev == kp) break;
            }
            h = (h + 1) & mask;
        }
        vp = vals + h;
    }
    uint32_t v = *reinterpret_cast<volatile uint32_t *>(vp);
    if ((v & TV_COUNT_MASK) < TV_COUNT_STOP) atomicAdd(vp, 1u);
    uint32_t need = (exts << 24) & ~v;
    if (need) atomicOr(vp, need);
}

constexpr int SPILL_TILE = 2048;                 // bases staged per warp per step
constexpr int SPILL_TILE_WORDS = SPILL_TILE / 16;
constexpr int SPILL_WARPS = 8;

// one warp per read
template <int W>
__global__ void __launch_bounds__(SPILL_WARPS * 32) k_spill_insert(SpillArgs a)
{
    __shared__ uint32_t tiles[SPILL_WARPS][SPILL_TILE_WORDS];
    using S = KeySlot<W>;
    const unsigned lane = lane_id();
    uint32_t *tile = tiles[threadIdx.x >> 5];
    const int K = a.K, kbits = 2 * K;
    const int64_t total_rows = a.row_off[a.n_list];
    const int64_t n_warps = (int64_t)gridDim.x * SPILL_WARPS;
    for (int64_t wr = (int64_t)blockIdx.x * SPILL_WARPS + (threadIdx.x >> 5); wr < total_rows; wr += n_warps) {
        // listed group of this row: last i with row_off[i] <= wr
        int64_t lo = 0, hi = a.n_list;
        while (hi - lo > 1) {
            int64_t m = (lo + hi) >> 1;
            if (a.row_off[m] <= wr) lo = m; else hi = m;
        }
        const int64_t i = lo;
        const int64_t g = a.ids ? a.ids[a.list_base + i] : a.list_base + i;
        const int64_t row = a.group_offsets[g] + (wr - a.row_off[i]);
        if (!col_valid(a.col, row)) continue;
        const int64_t off = col_off(a.col, row);
        const int64_t len = col_off(a.col, row + 1) - off;
        if (len < K) continue;   // contributes nothing whether valid or not
        const uint8_t *s = a.col.data + off;
        // validate the whole read first (src/fracture.rs:222: any bad byte drops the read)
        bool ok = true;
        for (int64_t j = lane; j < len; j += 32) ok &= base_ok(s[j]);
        if (!__all_sync(0xffffffffu, ok)) continue;
        const int64_t nk = len - K + 1;
        if (lane == 0) atomicAdd(a.n_obs, (unsigned long long)nk);
        typename S::T *keys = reinterpret_cast<typename S::T *>(a.keys) + a.slot_off[i];
        uint32_t *vals = a.vals + a.slot_off[i];
        const uint64_t mask = (uint64_t)(a.slot_off[i + 1] - a.slot_off[i]) - 1;
        int64_t w0 = 0;  // first base staged in the tile
        while (true) {
            const int64_t wend = (w0 + SPILL_TILE < len) ? w0 + SPILL_TILE : len;
            __syncwarp();
            for (int j = lane; j < SPILL_TILE_WORDS; j += 32) {
                int64_t b0 = w0 + 16 * j;
                uint32_t word = 0;
                if (b0 < wend) {
#pragma unroll
                    for (int t = 0; t < 16; t++) {
                        uint32_t c = (b0 + t < wend) ? base_code(s[b0 + t]) : 0u;
                        word |= c << (30 - 2 * t);
                    }
                }
                tile[j] = word;
            }
            __syncwarp();
            // positions whose window and right neighbour (when it exists) lie inside the tile
            const int64_t p_begin = (w0 == 0) ? 0 : w0 + 1;
            int64_t p_end = (wend == len) ? nk : (wend - K);   // exclusive
            const int64_t npos = p_end - p_begin;
            const int64_t per = (npos + 31) / 32;
            int64_t p = p_begin + (int64_t)lane * per;
            const int64_t pe = (p + per < p_end) ? p + per : p_end;
            if (p < pe) {
                Key<W> key = Key<W>::zero();
                auto base_at = [&](int64_t q) -> uint32_t {  // absolute base index -> code
                    int64_t t = q - w0;
                    return (tile[t >> 4] >> (30 - 2 * (int)(t & 15))) & 3u;
                };
                for (int t = 0; t < K - 1; t++) key.push(base_at(p + t), kbits);
                uint32_t left = (p > 0) ? (1u << base_at(p - 1)) : 0u;
                for (; p < pe; p++) {
                    key.push(base_at(p + K - 1), kbits);
                    uint32_t right = (p + K < len) ? (16u << base_at(p + K)) : 0u;
                    table_insert_global<W>(keys, vals, mask, key, left | right, a.special + i);
                    left = 1u << key.top_base(kbits);
                }
            }
            if (wend == len) break;
            w0 = p_end - 1;
        }
    }
}

// Survivors -> record pool, per-group survivor counts.  Every warp owns SPILL_EXTRACT_SLOTS consecutive slots.  Pass 1
// (count_only) writes the number of survivors of each warp; after an exclusive scan pass 2 writes the records at
// warp_base + running rank: no global cursor (a single cursor serialised ~15 M same-address atomics and made this kernel
// 10 ms for 7 GB of table, profiles/r01 notes), deterministic pool order.
constexpr int SPILL_EXTRACT_NT = 256;
constexpr int SPILL_EXTRACT_SLOTS = 2048;   // per warp
template <int W>
__global__ void __launch_bounds__(SPILL_EXTRACT_NT) k_spill_extract(SpillArgs a)
{
    using S = KeySlot<W>;
    const int64_t total = a.slot_off[a.n_list];
    const unsigned lane = lane_id();
    const int64_t gw = (int64_t)blockIdx.x * (SPILL_EXTRACT_NT / 32) + (threadIdx.x >> 5);
    const int64_t w0 = gw * SPILL_EXTRACT_SLOTS;
    if (w0 >= total) return;
    int64_t i0 = 0;   // listed group of the warp's first slot
    if (lane == 0) {
        int64_t lo = 0, hi = a.n_list;
        while (hi - lo > 1) {
            int64_t m = (lo + hi) >> 1;
            if (a.slot_off[m] <= w0) lo = m; else hi = m;
        }
        i0 = lo;
    }
    i0 = __shfl_sync(0xffffffffu, i0, 0);
    int64_t running = a.count_only ? 0 : a.warp_base[gw];
    const int kbits = 2 * a.K;
    for (int64_t slot = w0 + lane; slot - lane < w0 + SPILL_EXTRACT_SLOTS && slot - lane < total; slot += 32) {
        bool keep = false, keep_special = false;
        SpillRec rec{}, rec2{};
        int64_t g = -1;
        if (slot < total) {
            int64_t i = i0;
            while (i + 1 < a.n_list && a.slot_off[i + 1] <= slot) i++;
            g = a.ids ? a.ids[a.list_base + i] : a.list_base + i;
            // a claimed slot always carries a count >= 1 (the claimer adds right after its CAS), an empty one has val == 0:
            // the counting pass needs the value words only, the key is read for survivors
            const uint32_t v = a.vals[slot];
            {
                uint32_t c = v & TV_COUNT_MASK;
                if (c > 65535u) c = 65535u;
                if (c >= a.min_cov && c > 0) {
                    const Key<W> k = a.count_only ? Key<W>::zero() : S::unpack(reinterpret_cast<const typename S::T *>(a.keys)[slot]);
                    rec.hi = k.hi_word();
                    rec.lo = k.lo;
                    rec.grank = (uint32_t)(a.list_base + i);
                    rec.val = c | ((v >> 24) << 16);
                    keep = true;
                }
            }
            // the all-ones k-mer lives outside the table; emitted by the thread of the group's first slot
            if (slot == a.slot_off[i]) {
                const uint32_t v = a.special[i];
                uint32_t c = v & TV_COUNT_MASK;
                if (c > 65535u) c = 65535u;
                if (c >= a.min_cov && c > 0) {
                    rec2.lo = kbits >= 64 ? ~0ull : ((1ull << kbits) - 1ull);
                    rec2.hi = kbits >= 128 ? ~0ull : (kbits > 64 ? ((1ull << (kbits - 64)) - 1ull) : 0ull);
                    rec2.grank = (uint32_t)(a.list_base + i);
                    rec2.val = c | ((v >> 24) << 16);
                    keep_special = true;
                }
            }
            if (lane == 31) i0 = i;   // the next 32 slots start in this group or later
        }
        i0 = __shfl_sync(0xffffffffu, i0, 31);
        const unsigned m = __ballot_sync(0xffffffffu, keep), m2 = __ballot_sync(0xffffffffu, keep_special);
        const unsigned n_new = __popc(m) + __popc(m2);
        if (n_new == 0) continue;
        if (!a.count_only) {
            if (keep) a.pool[running + __popc(m & ((1u << lane) - 1u))] = rec;
            if (keep_special) a.pool[running + __popc(m) + __popc(m2 & ((1u << lane) - 1u))] = rec2;
            // survivors per group: one add per warp step when all its survivors share a group (the common case)
            const int leader = m ? __ffs(m) - 1 : __ffs(m2) - 1;
            const int64_t g_lead = __shfl_sync(0xffffffffu, g, leader);
            const bool uniform = __all_sync(0xffffffffu, (!keep && !keep_special) || g == g_lead);
            if (uniform) {
                if ((int)lane == leader) atomicAdd(reinterpret_cast<unsigned long long *>(a.nsurv + g_lead), (unsigned long long)n_new);
            } else {
                if (keep) atomicAdd(reinterpret_cast<unsigned long long *>(a.nsurv + g), 1ull);
                if (keep_special) atomicAdd(reinterpret_cast<unsigned long long *>(a.nsurv + g), 1ull);
            }
        }
        running += n_new;
    }
    if (a.count_only && lane == 0) a.warp_count[gw] = running;
}

__global__ void k_add_i64(int64_t *p, int64_t n, int64_t v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += v;
}

__global__ void k_gather_i64(const int64_t *src, const int32_t *ids, int64_t list_base, int64_t n, int64_t *dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[ids ? ids[list_base + i] : list_base + i];
}

// survivor records -> key columns + identity permutation (for the segmented sort), and the gather that applies a permutation
__global__ void k_rec_split(const SpillRec *recs, int64_t n, uint64_t *lo, uint64_t *hi, uint32_t *idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lo[i] = recs[i].lo;
    if (hi) hi[i] = recs[i].hi;
    idx[i] = (uint32_t)i;
}
__global__ void k_gather_u64(const uint64_t *src, const uint32_t *idx, int64_t n, uint64_t *dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_rec_gather(const SpillRec *src, const uint32_t *idx, int64_t n, SpillRec *dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void k_spill_scatter(const SpillRec *recs, int64_t n_recs, const int32_t *ids, int64_t list_base, const int64_t *list_rec_off,
                                const int64_t *kmer_offsets, uint64_t *out_lo, uint64_t *out_hi, uint8_t *out_exts,
                                uint16_t *out_counts)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_recs) return;
    SpillRec rec = recs[r];
    int64_t g = ids ? ids[rec.grank] : (int64_t)rec.grank;
    int64_t dst = kmer_offsets[g] + (r - list_rec_off[rec.grank - list_base]);
    out_lo[dst] = rec.lo;
    if (out_hi) out_hi[dst] = rec.hi;
    out_exts[dst] = (uint8_t)(rec.val >> 16);
    out_counts[dst] = (uint16_t)(rec.val & 0xffffu);
}

}  // namespace rogtk
