// kernels_part.cuh -- partition-then-count: the tier for groups beyond the multi-pass kernel (more than SM_PMAX * SM_PASS_OBS
// bases: UMIs with thousands of reads, BASELINE configs[2]).  Same semantics as debruijn::filter::filter_kmers + CountFilter at
// reference src/fracture.rs:110-116, which itself walks the k-mer space bucket range by bucket range.
//
// Round 1 sent these groups to open-addressing tables in HBM (kernels_table.cuh): one L2/HBM atomic per observation on random
// addresses, a scan of 2x obs slots to extract, a library sort of the records -- 1.3 x 10^10 obs/s on configs[2], where they hold
// most of the observations.  Here the k-mer space of a group is cut into P equal ranges (the multi-pass kernel's pass_of) so that
// one range's distinct k-mers fit a shared-memory table, and
//   k_part_scatter  stages slabs of the group's reads exactly like the hot kernel (128-bit loads, 2-bit pack) and runs twice:
//                   HIST = true counts the slab's observations per range (-> exact bucket sizes: the true k-mers of a deep UMI
//                   carry thousands of observations each, so range loads are far from uniform and no fixed slack fits them);
//                   HIST = false appends (k-mer, exts) records to the range buckets in HBM: a slab's records of one range form one
//                   contiguous run reserved with ONE global atomic per (slab, range), not one per observation;
//   k_part_count    one CTA per (group, range): reads the bucket with coalesced loads, counts it in a shared-memory table, prunes,
//                   and leaves the survivors in the pool as one segment (SegRec), exactly what a pass of the multi-pass kernel
//                   leaves -- ranges ascend, so ordering every segment in place orders the group (k_sort_small<W, true>).
// A group whose bucket or table overflows (extremely repetitive reads) is handed to the HBM tables, which hold anything.
#pragma once

namespace rogtk {

constexpr int PT_PMAX = 128;             // ranges per group at most (per-UMI groups; a batch holds thousands of them)
constexpr int64_t PT_TARGET = 8192;      // observations per range the plan aims at (table: 8192 slots; about half of the observations repeat a k-mer)
// One very large group (the un-grouped column of the global-spectrum mode, SURVEY 8e row 2: 10^8 observations in one group) gets
// up to PT_PBIG ranges of PT_TARGET_BIG observations -- most k-mers of such a group are seen once or twice, so a range's table must
// hold about as many k-mers as the range has observations -- and every scatter item covers PT_SUB_BIG slabs, so that an item's
// records of one range still form a run (about 16 x 16384 / P records) behind ONE global atomic.
constexpr int PT_PBIG = 49152;     // 192 KB of shared-memory bins: one scatter CTA per SM
constexpr int64_t PT_TARGET_BIG = 4096;
constexpr int PT_SUB_BIG = 16;
constexpr int PT_NT = 384;
constexpr int PT_NW = PT_NT / 32;
constexpr int PT_MAX_BASES = 16384;      // staged bases per slab (incl. alignment slack)
constexpr int PT_PK_WORDS = PT_MAX_BASES / 16 + 8;

struct PartArgs {
    ColView col;
    const uint8_t *data_begin, *data_end;
    const int64_t *group_offsets;
    const int32_t *ids;          // listed groups
    int64_t n_list;
    const int32_t *np;           // [n_list] ranges of entry e
    const int32_t *rows_per_slab;// [n_list]
    int sub_slabs;               // slabs per scatter item (1, or PT_SUB_BIG for a wave holding one large group)
    int pcap;                    // ranges per entry at most in this wave: stride of slab_cnt, size of the shared-memory bins
    int stage_cap;               // scatter: records of an item staged in shared memory and written out as whole runs (0: written directly)
    const int64_t *slab_off;     // [n_list+1] first slab item of entry e
    const int64_t *bucket_off;   // [n_list+1] first bucket of entry e
    const int32_t *item_entry;   // [n_items] entry of scatter item i   } one load instead of a binary search over slab_off / bucket_off:
    const int32_t *bucket_entry; // [n_buckets] entry of bucket b       } 16 dependent L2 round trips per item on a CTA that has the SM to itself
    const int64_t *rec_off;      // [n_buckets+1] first record slot of bucket b
    unsigned int *cursor;        // [n_buckets] HIST: records of the bucket; scatter: records appended so far
    uint32_t *slab_cnt;          // [n_items * pcap] observations per (item, range), written by HIST, read by the scatter
    uint64_t *rec_lo, *rec_hi;
    uint8_t *rec_ex;
    int32_t *failed;             // [n_list]
    int K;
    uint32_t min_cov;
    unsigned long long *ticket;
    // count kernel
    unsigned long long *pool_cursor;
    int64_t pool_cap;
    uint64_t *pool_lo, *pool_hi;
    uint8_t *pool_exts;
    uint16_t *pool_counts;
    SegRec *segs;                // [n_list * pcap]
    // merge mode (k_part_count<W, true>): the records of range b are the slices [m_bnd[r * m_stride + b], m_bnd[r * m_stride + b + 1])
    // of m_runs ascending runs (k-mer, partial count, exts) laid out one after the other in m_lo / m_hi / m_cnt / m_ext
    const uint64_t *m_lo, *m_hi;
    const uint16_t *m_cnt;
    const uint8_t *m_ext;
    const int64_t *m_bnd;
    int m_runs, m_stride;
    // prefix kernel
    int32_t *seg_count;          // [n_list]
    int64_t *nsurv;              // [n_groups]
    unsigned long long *n_obs;
    int32_t *failed_ids;         // groups handed to the HBM tables
    unsigned long long *n_failed;
};

// per listed group: observation upper bound and longest read, accumulated by `parts` warps per group (a batch of thousands of
// groups: one warp each; one group of millions of reads: thousands of warps), ...
__global__ void k_part_plan_rows(ColView col, const int64_t *group_offsets, const int32_t *ids, int64_t n_list, int K, int parts,
                                 unsigned long long *obs_acc, int *maxlen_acc)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n_list * parts; u += n_warps) {
        const int64_t e = u / parts, part = u % parts;
        const int64_t g = ids[e];
        const int64_t g0 = group_offsets[g], g1 = group_offsets[g + 1];
        const int64_t per = (g1 - g0 + parts - 1) / parts;
        const int64_t r0 = g0 + part * per, r1 = r0 + per < g1 ? r0 + per : g1;
        unsigned long long obs = 0;
        long long maxlen = 1;
        for (int64_t r = r0 + lane; r < r1; r += 32) {
            const int64_t len = col_off(col, r + 1) - col_off(col, r);
            if (len >= K) obs += (unsigned long long)(len - K + 1);
            maxlen = len > maxlen ? len : maxlen;
        }
        for (int d = 16; d; d >>= 1) {
            obs += __shfl_xor_sync(0xffffffffu, obs, d);
            const long long o = __shfl_xor_sync(0xffffffffu, maxlen, d);
            maxlen = o > maxlen ? o : maxlen;
        }
        if (lane == 0 && r0 < r1) {
            atomicAdd(&obs_acc[e], obs);
            atomicMax(&maxlen_acc[e], (int)(maxlen > 0x7fffffffll ? 0x7fffffffll : maxlen));
        }
    }
}
// ... -> ranges, rows per slab, scatter items, record capacity
__global__ void k_part_plan(const int64_t *group_offsets, const int32_t *ids, int64_t n_list, int K, const unsigned long long *obs_acc,
                            const int *maxlen_acc, int32_t *np, int32_t *rows_per_slab, int64_t *slabs, int64_t *bucket_cap)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_list) return;
    const int64_t g = ids[e];
    const int64_t r0 = group_offsets[g], r1 = group_offsets[g + 1];
    const unsigned long long obs = obs_acc[e];
    const long long maxlen = maxlen_acc[e] > 1 ? maxlen_acc[e] : 1;
    int64_t P = ((int64_t)obs + PT_TARGET - 1) / PT_TARGET;
    // rows per slab: every slab of rps rows fits the staged span whatever the reads' lengths (0 = a read longer than a slab)
    int64_t rps = (PT_MAX_BASES - 32) / maxlen;
    rps = rps > SM_MAXR ? SM_MAXR : rps;
    int64_t sub = 1;
    if (P > PT_PMAX && 2 * K > 24) {   // a large group (beyond 128 x 8192 observations)
        P = ((int64_t)obs + PT_TARGET_BIG - 1) / PT_TARGET_BIG;
        sub = PT_SUB_BIG;
        if (P > PT_PBIG) rps = 0;      // more than the bins hold: the HBM tables take it (callers bound their rounds below this)
    }
    P = P < 1 ? 1 : (P > PT_PBIG ? PT_PBIG : P);
    if (sub == 1 && P > PT_PMAX) P = PT_PMAX;
    if (2 * K <= 12) P = 1;   // at most 4096 distinct k-mers
    np[e] = (int32_t)P;
    rows_per_slab[e] = (int32_t)rps;
    slabs[e] = rps > 0 ? (r1 - r0 + rps * sub - 1) / (rps * sub) : 0;   // scatter items
    bucket_cap[e] = (int64_t)obs;   // records of the whole entry at most (wave planning)
}
__host__ __device__ inline size_t part_scatter_smem(int pcap) { return (size_t)PT_PK_WORDS * 4 + (size_t)(SM_MAXR + 1) * 4 + (size_t)pcap * 4 + 64 * 4 + SM_MAXR + 64; }
// ... plus the staging area of the scatter pass: per-range start / fill words and `cap` records
template <int W>
constexpr int part_stage_cap() { return W == 2 ? 11264 : PT_MAX_BASES; }
template <int W>
inline size_t part_scatter_smem_staged(int pcap)
{
    return ((part_scatter_smem(pcap) + 15) & ~(size_t)15) + (size_t)(2 * pcap + 2) * 4 + 8 + (size_t)part_stage_cap<W>() * (8 * W + 1);
}

// one slab of one listed group per iteration: stage, bin the observations by range, append the bins to the range buckets
// (the HIST pass runs two CTAs of 384 threads per SM; the scatter pass holds the staging area or the bins of a large group, has the
//  SM to itself and runs 768 threads -- with 384 its sweep was latency-bound: 62 -> 53 ms on configs[2] from the staging alone)
constexpr int PT_NT_SCATTER = 768;
template <int W, bool HIST>
__global__ void __launch_bounds__(HIST ? PT_NT : PT_NT_SCATTER, HIST ? 2 : 1) k_part_scatter(PartArgs a, int64_t n_items)
{
    constexpr int NT = HIST ? PT_NT : PT_NT_SCATTER, NWARP = NT / 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *scratch = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *pk = scratch + 1;
    uint32_t *rstart = scratch + PT_PK_WORDS;
    uint32_t *bin = rstart + SM_MAXR + 1;         // [pcap] HIST: observations of the item per range; scatter: next record slot of the
                                                  //        item's run inside the range bucket
    uint32_t *ws = bin + a.pcap;                  // 64 words
    uint8_t *bad = reinterpret_cast<uint8_t *>(ws + 64);
    // staging area (scatter pass, a.stage_cap > 0): an item's records are binned in shared memory first and leave as whole runs --
    // written directly, the 32 lanes of a store hit 32 different sectors, and the pass ran at the L2's sector-write rate
    uint8_t *stg = smem_raw + ((part_scatter_smem(a.pcap) + 15) & ~(size_t)15);
    uint32_t *soff = reinterpret_cast<uint32_t *>(stg);            // [pcap + 1] first staged slot of range p
    uint32_t *lfill = soff + a.pcap + 1;                             // [pcap] next staged slot of range p
    uint64_t *s_lo = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(lfill + a.pcap) + 7) & ~(uintptr_t)7);
    uint64_t *s_hi = s_lo + (W == 2 ? a.stage_cap : 0);
    uint8_t *s_ex = reinterpret_cast<uint8_t *>(s_hi + a.stage_cap);
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const int K = a.K;
    while (true) {
        __syncthreads();
        if (tid == 0) {
            const unsigned long long t = atomicAdd(a.ticket, 1ull);
            ws[0] = (uint32_t)t;
            ws[1] = (uint32_t)(t >> 32);
        }
        __syncthreads();
        const int64_t item = (int64_t)(((unsigned long long)ws[1] << 32) | ws[0]);
        if (item >= n_items) break;
        const int64_t e = a.item_entry[item];
        const int64_t g = a.ids[e];
        const int rps = a.rows_per_slab[e];
        const int P = a.np[e];
        const int64_t r1g = a.group_offsets[g + 1];
        const int64_t bucket0 = a.bucket_off[e];
        for (int p = tid; p < P; p += NT) {
            if (HIST) {
                bin[p] = 0;
            } else {   // the item's run inside every range bucket: one global atomic per (item, range)
                const uint32_t c = a.slab_cnt[item * a.pcap + p];
                bin[p] = c ? atomicAdd(&a.cursor[bucket0 + p], c) : 0u;
                if (a.stage_cap) soff[p + 1] = c;
            }
        }
        bool staged = false;
        if (!HIST && a.stage_cap) {   // staged slots: exclusive prefix of the per-range counts (P <= 128: one warp)
            __syncthreads();
            if (warp == 0) {
                uint32_t c[4], sum = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int p = (int)lane * 4 + t;
                    c[t] = p < P ? soff[p + 1] : 0u;
                    sum += c[t];
                }
                uint32_t incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += y;
                }
                uint32_t run = incl - sum;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int p = (int)lane * 4 + t;
                    if (p < P) {
                        soff[p] = run;       // (soff[p + 1] of the last range is rewritten below)
                        lfill[p] = run;
                    }
                    run += c[t];
                }
                if (lane == 31u) ws[2] = incl;
            }
            __syncthreads();
            if (tid == 0) soff[P] = ws[2];
            staged = ws[2] <= (uint32_t)a.stage_cap;
        }
        for (int sub = 0; sub < a.sub_slabs; sub++) {
            const int64_t r0 = a.group_offsets[g] + ((item - a.slab_off[e]) * a.sub_slabs + sub) * rps;
            if (r0 >= r1g) break;
            const int64_t r1 = r0 + rps < r1g ? r0 + rps : r1g;
            const int nr = (int)(r1 - r0);
            const int64_t sb0 = col_off(a.col, r0), sb1 = col_off(a.col, r1);
            const uintptr_t addr0 = reinterpret_cast<uintptr_t>(a.col.data + sb0);
            const uint32_t slack = (uint32_t)(addr0 & 15u);
            const uint32_t span = (uint32_t)((sb1 - sb0) + slack);
            __syncthreads();   // the previous slab's staged bases are no longer read
            for (int r = tid; r <= nr; r += NT) rstart[r] = (uint32_t)(col_off(a.col, r0 + r) - sb0) + slack;
            if (tid < nr) bad[tid] = col_valid(a.col, r0 + tid) ? 0 : 1;
            __syncthreads();
            // ---- ingest: 16-byte chunks -> one packed u32 (16 bases, first base in the top bits) --------------------------
            {
                const uint8_t *a0 = reinterpret_cast<const uint8_t *>(addr0 - slack);
                const int nchunks = (int)((span + 15u) >> 4);
                for (int c = tid; c < nchunks; c += NT) {
                    const uint8_t *p = a0 + 16 * c;
                    uint4 v;
                    if (p >= a.data_begin && p + 16 <= a.data_end) {
                        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
                    } else {
                        v = load_chunk_guarded(p, a.data_begin, a.data_end);
                    }
                    pk[c] = (pack4(v.x) << 24) | (pack4(v.y) << 16) | (pack4(v.z) << 8) | pack4(v.w);
                    const bool ok = valid4(v.x) & valid4(v.y) & valid4(v.z) & valid4(v.w);
                    if (!ok) mark_bad_reads(v, 16u * c, slack, span, rstart, nr, bad);
                }
            }
            __syncthreads();
            // ---- one sweep over the windows (a warp per read): HIST counts per range, the scatter writes ----------------------------------
            for (int r = (int)warp; r < nr; r += NWARP) {
                const uint32_t rs = rstart[r], len = rstart[r + 1] - rs;
                if (bad[r] || len < (uint32_t)K) continue;
                const uint32_t nk = len - (uint32_t)K + 1u;
                for (uint32_t p = lane; p < nk; p += 32u) {
                    const uint32_t q = rs + p;
                    Key<W> key = Key<W>::zero();
                    load_window(pk, q, K, key);
                    const uint32_t range = P > 1 ? pass_of(key_top32(key, 2 * K), (uint32_t)P) : 0u;
                    if (HIST) {
                        atomicAdd(&bin[range], 1u);
                    } else {
                        const uint32_t lb = base_of(pk, q - 1u), rb = base_of(pk, q + (uint32_t)K);
                        const uint32_t ex = (p ? (1u << lb) : 0u) | ((p + 1u < nk) ? (16u << rb) : 0u);
                        if (staged) {
                            const uint32_t slot = atomicAdd(&lfill[range], 1u);
                            s_lo[slot] = key.lo;
                            if (W == 2) s_hi[slot] = key.hi_word();
                            s_ex[slot] = (uint8_t)ex;
                        } else {
                            const int64_t at = a.rec_off[bucket0 + range] + atomicAdd(&bin[range], 1u);
                            a.rec_lo[at] = key.lo;
                            if (W == 2) a.rec_hi[at] = key.hi_word();
                            a.rec_ex[at] = (uint8_t)ex;
                        }
                    }
                }
            }
        }
        if (HIST) {
            __syncthreads();
            for (int p = tid; p < P; p += NT) {
                const uint32_t c = bin[p];
                a.slab_cnt[item * a.pcap + p] = c;
                if (c) atomicAdd(&a.cursor[bucket0 + p], c);
            }
        } else if (staged) {   // a warp per range: the run leaves as coalesced stores
            __syncthreads();
            for (int p = (int)warp; p < P; p += NWARP) {
                const uint32_t s0 = soff[p], n = soff[p + 1] - s0;
                const int64_t dst = a.rec_off[bucket0 + p] + bin[p];
                for (uint32_t i = lane; i < n; i += 32u) {
                    a.rec_lo[dst + i] = s_lo[s0 + i];
                    if (W == 2) a.rec_hi[dst + i] = s_hi[s0 + i];
                    a.rec_ex[dst + i] = s_ex[s0 + i];
                }
            }
        }
    }
}

template <int W>
constexpr size_t part_count_smem() { return (size_t)table_slots<W>() * (sizeof(typename KeySlot<W>::T) + 4) + 64 * 4; }

// one (group, range) bucket per iteration: count it in shared memory, prune, leave the survivors in the pool as one segment
// MERGE = true: the same table fed by slices of already counted ascending runs -- partial counts add up (saturating at 65535 when
// written out), exts are OR-ed: the merge step of the global-spectrum exchange (spectrum_nccl.cuh), one list entry, bucket = range.
template <int W, bool MERGE>
__global__ void __launch_bounds__(PT_NT, (W == 1 ? 2 : 1)) k_part_count(PartArgs a, int64_t n_buckets)
{
    constexpr int NT = PT_NT, TS = table_slots<W>();
    using S = KeySlot<W>;
    using KT = typename S::T;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    KT *tkey = reinterpret_cast<KT *>(smem_raw);
    uint32_t *tval = reinterpret_cast<uint32_t *>(tkey + TS);
    uint32_t *ws = tval + TS;   // [0],[1] ticket [2] all-ones k-mer value [3] abort [4],[5] pool base [6] survivors
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t thr = a.min_cov > 65535u ? 0xffffffffu : (a.min_cov > 1u ? a.min_cov : 1u);
    int64_t cta_pool_pos = 0, cta_pool_end = 0;
    while (true) {
        __syncthreads();
        if (tid == 0) {
            const unsigned long long t = atomicAdd(a.ticket, 1ull);
            ws[0] = (uint32_t)t;
            ws[1] = (uint32_t)(t >> 32);
            ws[2] = 0;
            ws[3] = 0;
            ws[6] = 0;
        }
        __syncthreads();
        const int64_t b = (int64_t)(((unsigned long long)ws[1] << 32) | ws[0]);
        if (b >= n_buckets) break;
        const int64_t e = MERGE ? 0 : (int64_t)a.bucket_entry[b];
        const int pass = MERGE ? (int)b : (int)(b - a.bucket_off[e]);
        if (a.failed[e]) continue;
        int64_t rec0 = 0;
        uint32_t n = 0;
        if (MERGE) {
            int64_t tot = 0;
            for (int r = 0; r < a.m_runs; r++) tot += a.m_bnd[(int64_t)r * a.m_stride + b + 1] - a.m_bnd[(int64_t)r * a.m_stride + b];
            if (tot > (int64_t)(TS - TS / 8)) {   // more records than the table takes for sure: the caller falls back to its sort
                if (tid == 0) a.failed[e] = 1;
                continue;
            }
            n = (uint32_t)tot;
        } else {
            rec0 = a.rec_off[b];
            n = (uint32_t)(a.rec_off[b + 1] - rec0);
        }
        // table sized to the bucket
        uint32_t tsize = (uint32_t)TS;
        if (W == 1) {
            int tb = 8;
            while (tb < SM_TBITS && (1u << tb) < 2u * n) tb++;
            tsize = 1u << tb;
        }
        {
            uint4 *kq = reinterpret_cast<uint4 *>(tkey);
            for (int j = tid; j < (int)(tsize * sizeof(KT) / 16); j += NT) kq[j] = make_uint4(~0u, ~0u, ~0u, ~0u);
            uint4 *vq = reinterpret_cast<uint4 *>(tval);
            for (int j = tid; j < (int)(tsize / 4); j += NT) vq[j] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        // (measured and dropped: merging equal k-mers inside a warp with __match_any_sync before the probe -- a range of a deep UMI
        //  holds a handful of true k-mers that carry most of its records -- made this kernel 2.4x slower: 62 -> 147 ms on configs[2])
        auto insert = [&](const Key<W> &key, uint32_t ex, uint32_t weight) {
            const KT kp = S::pack(key);
            if (kp == S::empty()) {   // the all-ones k-mer is the table's EMPTY marker: side word
                const uint32_t old = atomicAdd(&ws[2], weight << 8);
                if (ex & ~old) atomicOr(&ws[2], ex);
                return;
            }
            uint32_t h = W == 2 ? __umulhi(hash32(key), tsize) : (hash32(key) >> (32 - __ffs((int)tsize) + 1));
            uint32_t guard = 0;
            while (true) {
                KT cur = tkey[h];
                if (cur == S::empty()) cur = atomicCAS(&tkey[h], S::empty(), kp);
                if (cur == kp || cur == S::empty()) {
                    const uint32_t old = atomicAdd(&tval[h], weight << 8);
                    if (ex & ~old) atomicOr(&tval[h], ex);
                    break;
                }
                h = h + 1u == tsize ? 0u : h + 1u;
                if (++guard > tsize) {
                    ws[3] = 1;
                    break;
                }
            }
        };
        if (MERGE) {
            // (partial counts are at most 65535 and at most 257 runs meet in a table word: the 24-bit count field cannot overflow)
            for (int r = 0; r < a.m_runs; r++) {
                const int64_t s0 = a.m_bnd[(int64_t)r * a.m_stride + b], s1 = a.m_bnd[(int64_t)r * a.m_stride + b + 1];
                for (int64_t i = s0 + tid; i < s1; i += NT) {
                    Key<W> key = Key<W>::zero();
                    key.lo = a.m_lo[i];
                    if constexpr (W == 2) key.hi = a.m_hi[i];
                    insert(key, a.m_ext[i], a.m_cnt[i]);
                }
            }
        } else {
            // four records per thread in flight: with one, the loop ran at the latency of its global loads
            for (uint32_t i0 = tid; i0 < n; i0 += 4u * NT) {
                uint64_t lo4[4], hi4[4];
                uint32_t ex4[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)u * NT;
                    lo4[u] = hi4[u] = 0ull;
                    ex4[u] = 0u;
                    if (i < n) {
                        lo4[u] = a.rec_lo[rec0 + i];
                        if constexpr (W == 2) hi4[u] = a.rec_hi[rec0 + i];
                        ex4[u] = a.rec_ex[rec0 + i];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (i0 + (uint32_t)u * NT < n) {
                        Key<W> key = Key<W>::zero();
                        key.lo = lo4[u];
                        if constexpr (W == 2) key.hi = hi4[u];
                        insert(key, ex4[u], 1u);
                    }
                }
            }
        }
        if (tid == 0) {
            if (cta_pool_pos + (int64_t)(TS + 1) > cta_pool_end) {
                cta_pool_pos = (int64_t)atomicAdd(a.pool_cursor, (unsigned long long)SM_POOL_CHUNK);
                cta_pool_end = cta_pool_pos + (int64_t)SM_POOL_CHUNK;
            }
            ws[4] = (uint32_t)cta_pool_pos;
            ws[5] = (uint32_t)((uint64_t)cta_pool_pos >> 32);
        }
        __syncthreads();
        const int64_t pb = (int64_t)(((unsigned long long)ws[5] << 32) | ws[4]);
        if (ws[3] || pb + (int64_t)(TS + 1) > a.pool_cap) {   // table full / pool exhausted: the group is handed on
            if (tid == 0) a.failed[e] = 1;
            continue;
        }
        // prune + compact (as in k_smem_count)
        {
            const uint32_t nq = tsize / 4;
            const uint4 *tval4 = reinterpret_cast<const uint4 *>(tval);
            for (uint32_t j0 = 0; j0 < nq; j0 += NT) {
                const uint32_t j = j0 + (uint32_t)tid;
                uint4 vv = make_uint4(0u, 0u, 0u, 0u);
                if (j < nq) vv = tval4[j];
                const uint32_t c[4] = {vv.x, vv.y, vv.z, vv.w};
                uint32_t cntk = ((c[0] >> 8) >= thr) + ((c[1] >> 8) >= thr) + ((c[2] >> 8) >= thr) + ((c[3] >> 8) >= thr);
                uint32_t incl = cntk;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += y;
                }
                uint32_t wbase = 0;
                if (lane == 31u && incl) wbase = atomicAdd(&ws[6], incl);
                wbase = __shfl_sync(0xffffffffu, wbase, 31);
                int64_t pos = pb + (int64_t)(wbase + incl - cntk);
                if (cntk) {
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if ((c[t] >> 8) >= thr) {
                            const Key<W> k = S::unpack(tkey[4 * j + t]);
                            a.pool_lo[pos] = k.lo;
                            if (W == 2) a.pool_hi[pos] = k.hi_word();
                            a.pool_exts[pos] = (uint8_t)c[t];
                            a.pool_counts[pos] = (uint16_t)((c[t] >> 8) > 65535u ? 65535u : (c[t] >> 8));
                            pos++;
                        }
                }
            }
            if (tid == 0) {
                const uint32_t sv = ws[2];
                if ((sv >> 8) >= thr) {
                    const int64_t pos = pb + (int64_t)atomicAdd(&ws[6], 1u);
                    a.pool_lo[pos] = ~0ull;
                    if (W == 2) a.pool_hi[pos] = ~0ull;
                    a.pool_exts[pos] = (uint8_t)sv;
                    a.pool_counts[pos] = (uint16_t)((sv >> 8) > 65535u ? 65535u : (sv >> 8));
                }
            }
        }
        __syncthreads();
        const uint32_t n_out = ws[6];
        if (tid == 0) {
            cta_pool_pos += n_out;
            if (n_out > (uint32_t)SM_TSIZE) a.failed[e] = 1;   // more than the ordering kernels hold
            a.segs[e * a.pcap + pass] = SegRec{pb, (int32_t)n_out, 0};
        }
    }
}

// item -> entry and bucket -> entry maps of a wave (a warp per entry)
__global__ void k_part_maps(const int64_t *slab_off, const int64_t *bucket_off, int64_t n_list, int32_t *item_entry, int32_t *bucket_entry)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp; e < n_list; e += n_warps) {
        for (int64_t i = slab_off[e] + lane; i < slab_off[e + 1]; i += 32) item_entry[i] = (int32_t)e;
        for (int64_t b = bucket_off[e] + lane; b < bucket_off[e + 1]; b += 32) bucket_entry[b] = (int32_t)e;
    }
}

// per listed group: prefix of its segments, survivor total, observations; failed groups go to the next list
__global__ void k_part_finish(PartArgs a)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n_list) return;
    const int64_t g = a.ids[e];
    const int P = a.np[e];
    if (a.failed[e] || a.rows_per_slab[e] <= 0) {
        a.seg_count[e] = 0;
        a.failed_ids[atomicAdd(a.n_failed, 1ull)] = (int32_t)g;
        return;
    }
    int64_t pre = 0;
    unsigned long long obs = 0;
    for (int p = 0; p < P; p++) {
        SegRec &r = a.segs[e * a.pcap + p];
        r.prefix = (int32_t)pre;
        pre += r.n;
        obs += (unsigned long long)(a.rec_off[a.bucket_off[e] + p + 1] - a.rec_off[a.bucket_off[e] + p]);
    }
    a.seg_count[e] = P;
    a.nsurv[g] = pre;
    if (obs) atomicAdd(a.n_obs, obs);
}

// the same for a wave holding ONE large group (thousands of segments): one block, chunked scan
__global__ void __launch_bounds__(1024) k_part_finish_one(PartArgs a)
{
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const int64_t g = a.ids[0];
    const int P = a.np[0];
    if (a.failed[0] || a.rows_per_slab[0] <= 0) {
        if (tid == 0) {
            a.seg_count[0] = 0;
            a.failed_ids[atomicAdd(a.n_failed, 1ull)] = (int32_t)g;
        }
        return;
    }
    const int per = (P + 1023) / 1024;
    const int p0 = tid * per, p1 = p0 + per < P ? p0 + per : P;
    long long sum = 0;
    for (int p = p0; p < p1; p++) sum += a.segs[p].n;
    part[tid] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const long long y = tid >= d ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += y;
        __syncthreads();
    }
    long long pre = part[tid] - sum;
    for (int p = p0; p < p1; p++) {
        a.segs[p].prefix = (int32_t)pre;
        pre += a.segs[p].n;
    }
    if (tid == 1023) {
        a.seg_count[0] = P;
        a.nsurv[g] = part[1023];
        if (a.rec_off) {
            const unsigned long long obs = (unsigned long long)(a.rec_off[a.bucket_off[0] + P] - a.rec_off[a.bucket_off[0]]);
            if (obs) atomicAdd(a.n_obs, obs);
        }
        if (part[1023] > 0x7fffffffll) {   // (prefixes are 32-bit)
            a.seg_count[0] = 0;
            a.failed_ids[atomicAdd(a.n_failed, 1ull)] = (int32_t)g;
        }
    }
}

// merge mode: first record of every (run, range): bnd[r * stride + j] = first index in run r (ascending k-mers, [run_off[r],
// run_off[r+1]) of the concatenated arrays) whose range is >= base + j, j in [0, n_ranges]
template <int W>
__global__ void k_merge_bounds(const uint64_t *lo, const uint64_t *hi, const int64_t *run_off, int n_runs, int kbits, uint32_t total, uint32_t base,
                               int n_ranges, int64_t *bnd)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)n_ranges + 1;
    if (i >= (int64_t)n_runs * stride) return;
    const int r = (int)(i / stride), j = (int)(i % stride);
    int64_t a = run_off[r], b = run_off[r + 1];
    if (j == n_ranges) {
        bnd[i] = b;
        return;
    }
    const uint32_t want = base + (uint32_t)j;
    while (a < b) {
        const int64_t m = (a + b) >> 1;
        Key<W> key = Key<W>::zero();
        key.lo = lo[m];
        if constexpr (W == 2) key.hi = hi[m];
        if (pass_of(key_top32(key, kbits), total) < want) a = m + 1; else b = m;
    }
    bnd[i] = a;
}

}  // namespace rogtk

// -----------------------------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------------------------
namespace rogtk {

struct PartSegSet {   // segments of a run of listed groups: seg slot = (entry - first) * stride + range
    const int32_t *seg_list = nullptr;
    const SegRec *segs = nullptr;
    const int32_t *seg_count = nullptr;
    int64_t n_entries = 0;
    int stride = PT_PMAX;
};

struct PartResult {
    int64_t n_entries = 0;            // listed groups
    std::vector<PartSegSet> sets;
    const int32_t *failed_ids = nullptr;
    int64_t n_failed = 0;
    std::vector<DevBuf> keep;
};

inline void part_path_configure()
{
    auto cfg = [](auto kernel, size_t bytes) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    };
    cfg(k_part_scatter<1, true>, part_scatter_smem(PT_PBIG));
    cfg(k_part_scatter<2, true>, part_scatter_smem(PT_PBIG));
    cfg(k_part_scatter<1, false>, std::max(part_scatter_smem(PT_PBIG), part_scatter_smem_staged<1>(PT_PMAX)));
    cfg(k_part_scatter<2, false>, std::max(part_scatter_smem(PT_PBIG), part_scatter_smem_staged<2>(PT_PMAX)));
    cfg(k_part_count<1, false>, part_count_smem<1>());
    cfg(k_part_count<2, false>, part_count_smem<2>());
    cfg(k_part_count<1, true>, part_count_smem<1>());
    cfg(k_part_count<2, true>, part_count_smem<2>());
}

// d_ids: n_list listed groups (device).  Survivors go to the shared pool (pool_* / pool_cursor of the shared-memory path).
template <int W>
inline void run_part_path_w(cudaStream_t st, uint8_t *peek_slot, int sm_count, size_t workspace_bytes, const ColView &col, const uint8_t *data_begin,
                            const uint8_t *data_end, const int64_t *d_group_offsets, const int32_t *d_ids, int64_t n_list, int K, uint32_t min_cov,
                            unsigned long long *d_nobs, int64_t *d_nsurv, unsigned long long *pool_cursor, int64_t pool_cap, uint64_t *pool_lo,
                            uint64_t *pool_hi, uint8_t *pool_exts, uint16_t *pool_counts, PartResult &res, int64_t *launches)
{
    res.n_entries = n_list;
    if (n_list == 0) return;
    auto blocks = [](int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); };
    res.keep.emplace_back((size_t)n_list * 4, st);
    int32_t *d_segc = res.keep.back().as<int32_t>();
    res.keep.emplace_back((size_t)n_list * 4, st);
    int32_t *d_failed_ids = res.keep.back().as<int32_t>();
    CK(cudaMemsetAsync(d_segc, 0, (size_t)n_list * 4, st));
    DevBuf np((size_t)n_list * 4, st), rps((size_t)n_list * 4, st), slabs((size_t)n_list * 8, st), bcap((size_t)n_list * 8, st),
        failed((size_t)n_list * 4, st), ctr(64, st);
    CK(cudaMemsetAsync(failed.p, 0, (size_t)n_list * 4, st));
    CK(cudaMemsetAsync(ctr.p, 0, 64, st));
    unsigned long long *d_ctr = ctr.as<unsigned long long>();   // [0] scatter ticket [1] count ticket [2] n_failed
    {
        DevBuf acc((size_t)n_list * 12, st);   // u64 observations, then int longest read, per entry
        CK(cudaMemsetAsync(acc.p, 0, (size_t)n_list * 12, st));
        unsigned long long *obs_acc = acc.as<unsigned long long>();
        int *maxlen_acc = reinterpret_cast<int *>(obs_acc + n_list);
        const int64_t warps = (int64_t)sm_count * 16 * 8;
        const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(4096, warps / n_list));
        k_part_plan_rows<<<(unsigned)std::min<int64_t>(blocks(n_list * parts * 32, 256), (int64_t)sm_count * 16), 256, 0, st>>>(
            col, d_group_offsets, d_ids, n_list, K, parts, obs_acc, maxlen_acc);
        k_part_plan<<<blocks(n_list, 128), 128, 0, st>>>(d_group_offsets, d_ids, n_list, K, obs_acc, maxlen_acc, np.as<int32_t>(), rps.as<int32_t>(),
                                                         slabs.as<int64_t>(), bcap.as<int64_t>());
        if (launches) *launches += 2;
    }
    // waves bounded by the workspace: records of a wave = sum over its entries of np * capacity
    std::vector<int32_t> h_np((size_t)n_list);
    std::vector<int64_t> h_bcap((size_t)n_list);
    CK(cudaMemcpyAsync(h_np.data(), np.p, (size_t)n_list * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_bcap.data(), bcap.p, (size_t)n_list * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const size_t rec_bytes = (W == 2 ? 17 : 9);
    const int64_t budget = std::max<int64_t>(1 << 20, (int64_t)(workspace_bytes / 2 / rec_bytes));
    int64_t e0 = 0;
    while (e0 < n_list) {
        // a wave: a run of per-UMI entries (<= PT_PMAX ranges each) within the record budget, or ONE large group on its own
        const bool big = h_np[(size_t)e0] > PT_PMAX;
        int64_t e1 = e0, recs = 0;
        while (e1 < n_list && (e1 == e0 || (!big && h_np[(size_t)e1] <= PT_PMAX && recs + h_bcap[(size_t)e1] <= budget))) {
            recs += h_bcap[(size_t)e1];
            e1++;
        }
        const int64_t nl = e1 - e0;
        const int pcap = big ? h_np[(size_t)e0] : PT_PMAX;
        res.keep.emplace_back((size_t)nl * pcap * sizeof(SegRec), st);
        SegRec *d_segs = res.keep.back().as<SegRec>();
        PartSegSet set;
        set.seg_list = d_ids + e0; set.segs = d_segs; set.seg_count = d_segc + e0; set.n_entries = nl; set.stride = pcap;
        res.sets.push_back(set);
        DevBuf slab_off((size_t)(nl + 1) * 8, st), bucket_off((size_t)(nl + 1) * 8, st), scr(scan_scratch_bytes(nl), st);
        exclusive_scan<int64_t>(slabs.as<int64_t>() + e0, nl, slab_off.as<int64_t>(), scr.as<int64_t>(), st, launches);
        exclusive_scan<int32_t>(np.as<int32_t>() + e0, nl, bucket_off.as<int64_t>(), scr.as<int64_t>(), st, launches);
        int64_t h_tot[2] = {0, 0};
        Peek pk(peek_slot, st, launches);
        pk.add(slab_off.as<int64_t>() + nl, 8, &h_tot[0]);
        pk.add(bucket_off.as<int64_t>() + nl, 8, &h_tot[1]);
        pk.sync();
        const int64_t n_slabs = h_tot[0], n_buckets = h_tot[1];
        DevBuf rec_off((size_t)(n_buckets + 1) * 8, st), scr2(scan_scratch_bytes(n_buckets), st), cursor((size_t)std::max<int64_t>(n_buckets, 1) * 4, st),
            slab_cnt((size_t)std::max<int64_t>(n_slabs, 1) * pcap * 4, st);
        CK(cudaMemsetAsync(cursor.p, 0, cursor.bytes, st));
        DevBuf item_entry((size_t)std::max<int64_t>(n_slabs, 1) * 4, st), bucket_entry((size_t)std::max<int64_t>(n_buckets, 1) * 4, st);
        k_part_maps<<<(unsigned)std::min<int64_t>(blocks(nl * 32, 256), (int64_t)sm_count * 16), 256, 0, st>>>(
            slab_off.as<int64_t>(), bucket_off.as<int64_t>(), nl, item_entry.as<int32_t>(), bucket_entry.as<int32_t>());
        if (launches) *launches += 1;
        PartArgs a{};
        a.item_entry = item_entry.as<int32_t>(); a.bucket_entry = bucket_entry.as<int32_t>();
        a.col = col; a.data_begin = data_begin; a.data_end = data_end; a.group_offsets = d_group_offsets;
        a.ids = d_ids + e0; a.n_list = nl;
        a.np = np.as<int32_t>() + e0; a.rows_per_slab = rps.as<int32_t>() + e0;
        a.sub_slabs = big ? PT_SUB_BIG : 1; a.pcap = pcap;
        a.slab_off = slab_off.as<int64_t>(); a.bucket_off = bucket_off.as<int64_t>(); a.rec_off = rec_off.as<int64_t>();
        a.cursor = cursor.as<unsigned int>();
        a.slab_cnt = slab_cnt.as<uint32_t>();
        a.failed = failed.as<int32_t>() + e0;
        a.K = K; a.min_cov = min_cov;
        a.pool_cursor = pool_cursor; a.pool_cap = pool_cap; a.pool_lo = pool_lo; a.pool_hi = pool_hi; a.pool_exts = pool_exts; a.pool_counts = pool_counts;
        a.segs = d_segs; a.seg_count = d_segc + e0; a.nsurv = d_nsurv; a.n_obs = d_nobs;
        a.failed_ids = d_failed_ids; a.n_failed = d_ctr + 2;
        CK(cudaMemsetAsync(d_ctr, 0, 16, st));
        const size_t ssmem = part_scatter_smem(pcap);
        const unsigned sgrid = (unsigned)std::min<int64_t>(std::max<int64_t>(n_slabs, 1), (int64_t)sm_count * 2);
        a.ticket = d_ctr;
        if (n_slabs > 0) k_part_scatter<W, true><<<sgrid, PT_NT, ssmem, st>>>(a, n_slabs);
        // exact bucket sizes -> record offsets; the cursors restart at zero for the scatter
        exclusive_scan<uint32_t>(cursor.as<uint32_t>(), n_buckets, rec_off.as<int64_t>(), scr2.as<int64_t>(), st, launches);
        int64_t n_recs = 0;
        pk.add(rec_off.as<int64_t>() + n_buckets, 8, &n_recs);
        pk.sync();
        CK(cudaMemsetAsync(cursor.p, 0, cursor.bytes, st));
        CK(cudaMemsetAsync(d_ctr, 0, 8, st));
        DevBuf rlo((size_t)std::max<int64_t>(n_recs, 1) * 8, st), rhi(W == 2 ? (size_t)std::max<int64_t>(n_recs, 1) * 8 : 0, st),
            rex((size_t)std::max<int64_t>(n_recs, 1), st);
        a.rec_lo = rlo.as<uint64_t>(); a.rec_hi = W == 2 ? rhi.as<uint64_t>() : nullptr; a.rec_ex = rex.as<uint8_t>();
        if (n_slabs > 0) {
            // per-UMI waves stage an item's records in shared memory (one CTA per SM); one large group has too many ranges for runs
            a.stage_cap = big ? 0 : part_stage_cap<W>();
            const unsigned grid2 = (unsigned)std::min<int64_t>(n_slabs, (int64_t)sm_count);
            k_part_scatter<W, false><<<grid2, PT_NT_SCATTER, big ? ssmem : part_scatter_smem_staged<W>(pcap), st>>>(a, n_slabs);
        }
        if (n_buckets > 0) {
            a.ticket = d_ctr + 1;
            const unsigned grid = (unsigned)std::min<int64_t>(n_buckets, (int64_t)sm_count * (W == 1 ? 2 : 1));
            k_part_count<W, false><<<grid, PT_NT, part_count_smem<W>(), st>>>(a, n_buckets);
        }
        if (big) k_part_finish_one<<<1, 1024, 0, st>>>(a);
        else k_part_finish<<<blocks(nl, 128), 128, 0, st>>>(a);
        if (launches) *launches += 5;
        // (the wave's buffers are released in stream order when they go out of scope)
        e0 = e1;
    }
    unsigned long long nf = 0;
    Peek pk(peek_slot, st, launches);
    pk.add(d_ctr + 2, 8, &nf);
    pk.sync();
    res.failed_ids = d_failed_ids;
    res.n_failed = (int64_t)nf;
    res.keep.push_back(std::move(ctr));
}

// Merge of R ascending runs of counted k-mers that all lie in ranges [base, base + n_ranges) of `total` equal ranges of the k-mer
// space: per range one shared-memory table (k_part_count<W, true>), survivors ordered by the segment sort.  Returns false when a
// range does not fit a table (skewed k-mers): nothing is written and the caller uses its generic sort-based merge.
// d_run_off: device [n_runs + 1].  Outputs (allocated here): o_* with n_out survivors, ascending.
template <int W>
inline bool merge_sorted_runs_w(cudaStream_t st, uint8_t *peek_slot, int sm_count, const uint64_t *m_lo, const uint64_t *m_hi, const uint16_t *m_cnt,
                                const uint8_t *m_ext, const int64_t *d_run_off, int n_runs, int64_t n_in, int K, uint32_t min_cov, uint32_t total,
                                uint32_t base, int n_ranges, DevBuf &o_lo, DevBuf &o_hi, DevBuf &o_exts, DevBuf &o_counts, int64_t &n_out,
                                int64_t *launches)
{
    auto blocks = [](int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); };
    const int64_t stride = (int64_t)n_ranges + 1;
    DevBuf bnd((size_t)n_runs * stride * 8, st), segs((size_t)n_ranges * sizeof(SegRec), st), misc(256, st), boff(16, st);
    CK(cudaMemsetAsync(misc.p, 0, 256, st));
    k_merge_bounds<W><<<blocks((int64_t)n_runs * stride, 256), 256, 0, st>>>(m_lo, m_hi, d_run_off, n_runs, 2 * K, total, base, n_ranges, bnd.as<int64_t>());
    // words of misc: [0] ticket [1] pool cursor [2] n_failed [3] nsurv [4] n_obs (unused) ; 32-bit: np, rows_per_slab, failed, seg_count, ids, failed_ids
    unsigned long long *d_ctr = misc.as<unsigned long long>();
    int32_t *d_i32 = reinterpret_cast<int32_t *>(d_ctr + 8);
    const int32_t h_i32[6] = {n_ranges, 1, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(d_i32, h_i32, sizeof(h_i32), cudaMemcpyHostToDevice, st));
    const int64_t h_boff[2] = {0, n_ranges};
    CK(cudaMemcpyAsync(boff.p, h_boff, 16, cudaMemcpyHostToDevice, st));
    const unsigned grid = (unsigned)std::min<int64_t>(n_ranges, (int64_t)sm_count * (W == 1 ? 2 : 1));
    const int64_t pool_cap = n_in + n_in / 4 + (int64_t)(grid + 2) * (SM_POOL_CHUNK + table_slots<W>() + 1);
    DevBuf p_lo((size_t)pool_cap * 8, st), p_hi(W == 2 ? (size_t)pool_cap * 8 : 0, st), p_exts((size_t)pool_cap, st), p_counts((size_t)pool_cap * 2, st);
    PartArgs a{};
    a.n_list = 1;
    a.ids = d_i32 + 4; a.np = d_i32; a.rows_per_slab = d_i32 + 1; a.failed = d_i32 + 2; a.seg_count = d_i32 + 3; a.failed_ids = d_i32 + 5;
    a.bucket_off = boff.as<int64_t>();
    a.pcap = n_ranges;
    a.min_cov = min_cov; a.K = K;
    a.ticket = d_ctr; a.pool_cursor = d_ctr + 1; a.n_failed = d_ctr + 2; a.nsurv = reinterpret_cast<int64_t *>(d_ctr + 3); a.n_obs = d_ctr + 4;
    a.pool_cap = pool_cap; a.pool_lo = p_lo.as<uint64_t>(); a.pool_hi = W == 2 ? p_hi.as<uint64_t>() : nullptr;
    a.pool_exts = p_exts.as<uint8_t>(); a.pool_counts = p_counts.as<uint16_t>();
    a.segs = segs.as<SegRec>();
    a.m_lo = m_lo; a.m_hi = m_hi; a.m_cnt = m_cnt; a.m_ext = m_ext; a.m_bnd = bnd.as<int64_t>(); a.m_runs = n_runs; a.m_stride = (int)stride;
    k_part_count<W, true><<<grid, PT_NT, part_count_smem<W>(), st>>>(a, n_ranges);
    k_part_finish_one<<<1, 1024, 0, st>>>(a);
    if (launches) *launches += 3;
    unsigned long long nf = 0;
    int64_t ns = 0;
    Peek pk(peek_slot, st, launches);
    pk.add(d_ctr + 2, 8, &nf);
    pk.add(d_ctr + 3, 8, &ns);
    pk.sync();
    if (nf) return false;
    n_out = ns;
    const size_t no = (size_t)std::max<int64_t>(n_out, 1);
    o_lo.alloc(no * 8, st);
    if (W == 2) o_hi.alloc(no * 8, st);
    o_exts.alloc(no, st);
    o_counts.alloc(no * 2, st);
    DevBuf koff(16, st);
    const int64_t h_koff[2] = {0, n_out};
    CK(cudaMemcpyAsync(koff.p, h_koff, 16, cudaMemcpyHostToDevice, st));
    if (n_out > 0)
        run_sort_gather(st, peek_slot, sm_count, K, 1, nullptr, koff.as<int64_t>(), a.pool_lo, a.pool_hi, a.pool_exts, a.pool_counts, o_lo.as<uint64_t>(),
                        W == 2 ? o_hi.as<uint64_t>() : nullptr, o_exts.as<uint8_t>(), o_counts.as<uint16_t>(), launches, a.segs, a.seg_count, a.ids, 1,
                        n_ranges, false, total, base);
    CK(cudaStreamSynchronize(st));   // h_* staging arrays above are on this frame
    return true;
}

}  // namespace rogtk
