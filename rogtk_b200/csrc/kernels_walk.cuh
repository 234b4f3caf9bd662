// kernels_walk.cuh -- round-2 counting kernel for groups that fit one staged span: successor-link walk.
//
// Same contract as k_smem_count<W, KF, false> (kernels_smem.cuh): one CTA builds one UMI group's node set in shared memory and
// leaves the min-coverage survivors unsorted in the pool; together with the ordering kernels it replaces
// debruijn::filter::filter_kmers + CountFilter as called at reference src/fracture.rs:110-116.  What changed is how an
// observation finds its table slot.  Round 1 paid hash + LDS.64 + 64-bit compare + ATOMS for every window although two
// thirds of the windows repeat a k-mer whose predecessor window was looked up one step earlier.  Here
//
//   * a slot is 8 bytes whatever K is: a value word (count << 8 | exts) and a meta word holding a 14-bit REFERENCE to the
//     staged position of the k-mer's first sighting (the packed reads stay in shared memory, so the k-mer itself need not be
//     stored), a 2-bit fingerprint, and a 16-bit successor LINK (slot of the k-mer that followed this one, tagged with the
//     base that led there).  8192 slots are 64 KB: three CTAs per SM for one-word k-mers, and two-word k-mers (K = 64) fit
//     two CTAs per SM with 12288 slots instead of one CTA with a 160 KB table;
//   * a lane owns a RANGE of up to RL consecutive windows of one read.  It looks the first one up by hash and then WALKS:
//     next base -> link of the current slot -> one 16-bit load, one compare, one ATOMS.ADD on the target.  A link names its
//     target k-mer exactly (it was stored by a thread that saw both k-mers side by side), so a walked window needs no hash,
//     no key compare and no exts update: the two neighbour bits of that transition were set by the thread that created the link;
//   * where the walk breaks (first sightings -- every read error spoils K windows -- and links not made yet) the rest of the
//     range is handed to the warp, which inserts the leftover windows of all its lanes densely, 32 consecutive windows per step
//     (hash, probe, claim with one 32-bit CAS, own exts), and stores the links between neighbouring windows.
//
// Exts bookkeeping (left bit = previous base unless the window starts the read, right bit = next base unless it ends
// it).  Hashed windows (range heads, dense steps) set both of their own bits.  A walked window sets nothing: its left bit and
// its predecessor's right bit are exactly the bits the link's creator set.  That leaves the right bit of a window whose successor
// is hashed by somebody else: the last window of a completed range sets it itself, and the first window of a leftover run sets it
// on the predecessor slot it inherited from the walker.
#pragma once
// (included by kernels_smem.cuh after its device code: SmemArgs, window helpers, block_scan_u32, ingest helpers)

namespace rogtk {

constexpr int WK_NT = 384;
constexpr int WK_NW = WK_NT / 32;
constexpr int WK_MAX_BASES = 16368;                 // staged bases incl. alignment slack: references are 14 bits, 0x3fff|fp 3 = EMPTY
constexpr int WK_PK_WORDS = 16384 / 16 + 8;         // 1 word of front pad, tail pad for window over-reads
constexpr int WK_MIN_RL_BITS = 3;                   // ranges of at least 8 windows
constexpr int WK_MAX_RANGES = 16384 / (1 << WK_MIN_RL_BITS) + SM_MAXR;
constexpr uint32_t WK_EMPTY = 0xffffffffu;
constexpr uint32_t WK_NOLINK = 0xffffu;

template <int KW>
constexpr int walk_slots() { return KW == 2 ? 12288 : 8192; }
template <int KW>
constexpr int walk_min_blocks() { return KW == 2 ? 2 : 3; }
inline int walk_min_blocks_rt(int kw) { return kw == 2 ? walk_min_blocks<2>() : walk_min_blocks<1>(); }
template <int KW>
constexpr size_t walk_smem_bytes()
{
    return (size_t)walk_slots<KW>() * 8 + (size_t)WK_PK_WORDS * 4 + (size_t)(SM_MAXR + 1) * 4 * 2 + 64 * 4 + SM_MAXR + WK_MAX_RANGES + 64;
}

// ---- shared-memory access by 32-bit shared-space address (the generic-pointer forms made the compiler rebuild the window base
//      inside the walk loop: 36 instructions per walked window instead of 19) ---------------------------------------------------
template <int OFF>
__device__ __forceinline__ uint32_t sld32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t sld16(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sst16(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u16 [%0+%1], %2;" ::"r"(a), "n"(OFF), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ uint32_t satom_add(uint32_t a, uint32_t v)
{
    uint32_t o;
    asm volatile("atom.shared.add.u32 %0, [%1+%2], %3;" : "=r"(o) : "r"(a), "n"(OFF), "r"(v) : "memory");
    return o;
}
template <int OFF>
__device__ __forceinline__ void sred_add(uint32_t a, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sred_or(uint32_t a, uint32_t v)
{
    asm volatile("red.shared.or.b32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t satom_cas(uint32_t a, uint32_t cmp, uint32_t v)
{
    uint32_t o;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "r"(a), "r"(cmp), "r"(v) : "memory");
    return o;
}
// base at staged index y (first base of a word in its top bits)
__device__ __forceinline__ uint32_t wk_base(uint32_t pk_sa, uint32_t y)
{
    return (sld32<0>(pk_sa + ((y >> 2) & ~3u)) >> ((~(y << 1)) & 30u)) & 3u;
}
// k-mer of the window at staged base q; xm/x0 (the words before / at q) are handed back for the left neighbour
template <int KW>
__device__ __forceinline__ Key<KW> wk_window(uint32_t pk_sa, uint32_t q, int K, uint32_t &lb)
{
    const uint32_t wa = pk_sa + ((q >> 2) & ~3u), sh = 2u * (q & 15u);
    const uint32_t xm = sld32<-4>(wa), x0 = sld32<0>(wa), x1 = sld32<4>(wa), x2 = sld32<8>(wa);
    lb = __funnelshift_l(x0, xm, sh) & 3u;   // base q - 1
    const uint64_t xh = ((uint64_t)__funnelshift_l(x1, x0, sh) << 32) | __funnelshift_l(x2, x1, sh);
    Key<KW> key = Key<KW>::zero();
    if constexpr (KW == 1) {
        key.lo = xh >> (64 - 2 * K);
    } else {
        const uint32_t x3 = sld32<12>(wa), x4 = sld32<16>(wa);
        const uint64_t xl = ((uint64_t)__funnelshift_l(x3, x2, sh) << 32) | __funnelshift_l(x4, x3, sh);
        const int s2 = 128 - 2 * K;   // 0..62
        key.hi = xh >> s2;
        key.lo = s2 ? ((xl >> s2) | (xh << (64 - s2))) : xl;
    }
    return key;
}

// hash lookup-or-claim of the window at staged base q, then count + own exts (left bit iff has_l, right bit iff has_r).
// Returns the shared address of the slot's meta word (its value word sits TS * 4 bytes below); `claimed` tells a first sighting.
template <int KW, int TS, bool STATS>
__device__ __forceinline__ uint32_t wk_insert(uint32_t pk_sa, uint32_t tmeta_sa, uint32_t tsize, int K, uint32_t q, bool has_l, bool has_r,
                                              bool &claimed, uint32_t abort_sa, uint32_t *st_probe, uint32_t *st_cmp)
{
    uint32_t lb;
    const Key<KW> key = wk_window<KW>(pk_sa, q, K, lb);
    uint32_t ex = has_l ? (1u << lb) : 0u;
    if (has_r) ex |= 16u << wk_base(pk_sa, q + (uint32_t)K);
    const uint32_t h = hash32(key);
    const uint32_t end_sa = tmeta_sa + (tsize << 2);
    uint32_t ma = tmeta_sa + (__umulhi(h, tsize) << 2);
    const uint32_t mine = 0xffff0000u | (q << 2) | (h & 3u);
    claimed = false;
    uint32_t guard = 0;
    while (true) {
        uint32_t m = sld32<0>(ma);
        if (STATS) ++*st_probe;
        if ((m & 0xffffu) == 0xffffu) {   // empty: claim it with this window as the k-mer's reference
            const uint32_t old = satom_cas(ma, WK_EMPTY, mine);
            if (old == WK_EMPTY) {
                claimed = true;
                break;
            }
            m = old;
        }
        if (((m ^ mine) & 3u) == 0u) {   // fingerprint agrees: compare the k-mers (the slot's one is read through its reference)
            uint32_t dummy;
            if (STATS) ++*st_cmp;
            const Key<KW> other = wk_window<KW>(pk_sa, (m & 0xffffu) >> 2, K, dummy);
            if (other == key) break;
        }
        ma += 4u;
        if (ma == end_sa) ma = tmeta_sa;
        if (++guard > tsize) {   // table full: the group goes to the next tier
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(abort_sa), "r"(1u) : "memory");
            break;
        }
    }
    const uint32_t old = satom_add<-TS * 4>(ma, 256u);
    if (ex & ~old) sred_or<-TS * 4>(ma, ex);
    return ma;
}

template <int KW, bool STATS>
__global__ void __launch_bounds__(WK_NT, walk_min_blocks<KW>()) k_walk_count(SmemArgs a, int rl_bits)
{
    uint32_t st_walk = 0, st_dense = 0, st_head = 0, st_probe = 0, st_dsteps = 0, st_wsteps = 0, st_cmp = 0, st_broken = 0;   // STATS only
    constexpr int NT = WK_NT, NW = WK_NW, TS = walk_slots<KW>();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *tval = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *tmeta = tval + TS;
    uint32_t *scratch = tmeta + TS;
    uint32_t *pk = scratch + 1;                       // pk[-1] is a readable pad word
    uint32_t *rstart = scratch + WK_PK_WORDS;
    uint32_t *rcum = rstart + SM_MAXR + 1;            // ranges before read r
    uint32_t *ws = rcum + SM_MAXR + 1;                // 64 words, see WS_*
    uint8_t *bad = reinterpret_cast<uint8_t *>(ws + 64);
    uint8_t *rng_read = bad + SM_MAXR;                // read of every range
    volatile uint32_t *misc = ws + WS_MISC;   // [0],[1] ticket [3] abort [4],[5] pool base [6] survivors [7] block counter [8..15] r0,r1,b0,b1
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const int K = a.K;
    const uint32_t RL = 1u << rl_bits;
    const uint32_t thr = a.min_cov > 65535u ? 0xffffffffu : (a.min_cov > 1u ? a.min_cov : 1u);
    const int64_t n_items = a.n_groups;
    unsigned long long cta_obs = 0;
    long long t_phase = clock64();

    unsigned long long pf_g = 0;
    int64_t pf_r0 = 0, pf_r1 = 0, pf_b0 = 0, pf_b1 = 0;
    int pf_stage = 0;
    int64_t cta_pool_pos = 0, cta_pool_end = 0;
    if (tid == 0) pf_g = atomicAdd(a.ticket, 1ull);
    auto pf_advance = [&]() {   // thread 0 only: metadata of the next group, one dependent load per call
        if ((int64_t)pf_g >= n_items) {
            pf_stage = 2;
            return;
        }
        if (pf_stage == 0) {
            pf_r0 = a.group_offsets[pf_g];
            pf_r1 = a.group_offsets[pf_g + 1];
            pf_stage = 1;
        } else if (pf_stage == 1) {
            pf_b0 = col_off(a.col, pf_r0);
            pf_b1 = col_off(a.col, pf_r1);
            pf_stage = 2;
        }
    };
    auto defer = [&](int64_t gid) {
        if (tid == 0) a.deferred_ids[atomicAdd(a.n_deferred, 1ull)] = (int32_t)gid;
    };

    while (true) {
        __syncthreads();
        if (tid == 0) {
            while (pf_stage < 2) pf_advance();
            misc[0] = (uint32_t)pf_g;
            misc[1] = (uint32_t)(pf_g >> 32);
            misc[3] = 0;
            misc[6] = 0;
            misc[7] = 0;
            misc[8] = (uint32_t)pf_r0;  misc[9] = (uint32_t)((uint64_t)pf_r0 >> 32);
            misc[10] = (uint32_t)pf_r1; misc[11] = (uint32_t)((uint64_t)pf_r1 >> 32);
            misc[12] = (uint32_t)pf_b0; misc[13] = (uint32_t)((uint64_t)pf_b0 >> 32);
            misc[14] = (uint32_t)pf_b1; misc[15] = (uint32_t)((uint64_t)pf_b1 >> 32);
        }
        __syncthreads();
        const int64_t g = (int64_t)(((unsigned long long)misc[1] << 32) | misc[0]);
        if (g >= n_items) break;
        const int64_t r0 = (int64_t)(((uint64_t)misc[9] << 32) | misc[8]);
        const int64_t r1 = (int64_t)(((uint64_t)misc[11] << 32) | misc[10]);
        const int64_t b0 = (int64_t)(((uint64_t)misc[13] << 32) | misc[12]);
        const int64_t b1 = (int64_t)(((uint64_t)misc[15] << 32) | misc[14]);
        if (tid == 0) {
            pf_g = atomicAdd(a.ticket, 1ull);
            pf_stage = 0;
        }
        if (r1 - r0 <= 0) continue;
        if (b1 <= b0) continue;
        const int64_t nrows = r1 - r0;
        const uintptr_t addr0 = reinterpret_cast<uintptr_t>(a.col.data + b0);
        const uint32_t slack = (uint32_t)(addr0 & 15u);
        const int64_t span64 = (b1 - b0) + slack;
        if (nrows > SM_MAXR || span64 > WK_MAX_BASES) {   // too many rows / bases for one staged span: multi-pass tier
            defer(g);
            continue;
        }
        const uint32_t span = (uint32_t)span64;
        const int nr = (int)nrows;

        for (int r = tid; r <= nr; r += NT) rstart[r] = (uint32_t)(col_off(a.col, r0 + r) - b0) + slack;
        if (tid < nr) bad[tid] = col_valid(a.col, r0 + tid) ? 0 : 1;
        __syncthreads();

        PHASE_MARK(0);
        // ---- ingest: 16-byte chunks -> one packed u32 (16 bases, first base in the top bits) --------------------------
        {
            const uint8_t *a0 = reinterpret_cast<const uint8_t *>(addr0 - slack);
            const int nchunks = (int)((span + 15u) >> 4);
            uint4 v[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
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
            for (int i = 0; i < 3; i++) {
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
        // ---- windows per read -> ranges of RL windows, prefix sums, read of every range ---------------------------------
        uint32_t total, nranges;
        {
            uint32_t nk = 0;
            if (tid < nr) {
                const uint32_t len = rstart[tid + 1] - rstart[tid];
                if (!bad[tid] && len >= (uint32_t)K) nk = len - K + 1;
            }
            block_scan_u32<NW>(nk, &total, ws);
            const uint32_t nrg = (nk + RL - 1u) >> rl_bits;
            const uint32_t ex = block_scan_u32<NW>(nrg, &nranges, ws);
            if (tid < SM_MAXR) rcum[tid] = ex;
            for (uint32_t j = 0; j < nrg; j++) rng_read[ex + j] = (uint8_t)tid;
        }
        if (total == 0) continue;   // (the loop-top barrier orders the shared writes before reuse)
        // table sized to the group: about two slots per observation, at most TS
        uint32_t tsize = 2u * total + 64u;
        tsize = tsize > (uint32_t)TS ? (uint32_t)TS : (tsize & ~3u);
        {
            uint4 *vq = reinterpret_cast<uint4 *>(tval);
            uint4 *mq = reinterpret_cast<uint4 *>(tmeta);
            for (int j = tid; j < (int)(tsize / 4); j += NT) {
                vq[j] = make_uint4(0u, 0u, 0u, 0u);
                mq[j] = make_uint4(WK_EMPTY, WK_EMPTY, WK_EMPTY, WK_EMPTY);
            }
        }
        __syncthreads();

        PHASE_MARK(2);
        if (tid == 0 && pf_stage == 0) pf_advance();
        // ---- counting: warps fetch blocks of 32 ranges ------------------------------------------------------------------------------
        {
            const uint32_t ctr_addr = (uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&misc[7]));
            const uint32_t abort_sa = (uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&misc[3]));
            const uint32_t pk_sa = (uint32_t)__cvta_generic_to_shared(pk);
            const uint32_t tmeta_sa = (uint32_t)__cvta_generic_to_shared(tmeta);
            const uint32_t km1 = (uint32_t)K - 1u;
            while (true) {
                uint32_t blk = 0;
                if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(blk) : "r"(ctr_addr));
                blk = __shfl_sync(0xffffffffu, blk, 0);
                if (blk * 32u >= nranges || sld32<0>(abort_sa)) break;
                const uint32_t x = blk * 32u + lane;
                const bool have = x < nranges;
                // q: next window (staged index of its first base), qe: end of the range, ma: meta word of the slot of window q - 1
                uint32_t q = 0, qe = 0, ma = tmeta_sa, rs = 0, rl = 0;
                bool alive = false;
                if (have) {
                    const uint32_t r = rng_read[x];
                    rs = rstart[r];
                    const uint32_t nk = rstart[r + 1] - rs - km1;
                    const uint32_t w0 = (x - rcum[r]) << rl_bits;
                    const uint32_t w1 = w0 + RL < nk ? w0 + RL : nk;
                    rl = rs + nk - 1u;          // start of the read's last window
                    q = rs + w0;
                    qe = rs + w1;
                    bool clm;
                    ma = wk_insert<KW, TS, STATS>(pk_sa, tmeta_sa, tsize, K, q, q > rs, q < rl, clm, abort_sa, &st_probe, &st_cmp);   // head of the range: by hash
                    if (STATS) st_head++;
                    q++;
                    alive = true;
                }
                // ---- walk: one link per window ------------------------------------------------------------------------------
                while (true) {
                    const bool go = alive && q < qe;
                    if (!__any_sync(0xffffffffu, go)) break;
                    if (STATS && lane == 0) st_wsteps++;
                    if (go) {
                        const uint32_t b = wk_base(pk_sa, q + km1);
                        const uint32_t lk = sld16<2>(ma);
                        if (lk != WK_NOLINK && ((lk ^ b) & 3u) == 0u) {
                            ma = tmeta_sa + (lk & 0xfffcu);   // link = slot << 2 | base: the slot's byte offset is in place
                            sred_add<-TS * 4>(ma, 256u);
                            q++;
                            if (STATS) st_walk++;
                        } else {
                            alive = false;
                            if (STATS) st_broken++;
                        }
                    }
                }
                // a completed range: the right neighbour bit of its last window (the next window is somebody else's range head)
                if (have && alive && q - 1u < rl) sred_or<-TS * 4>(ma, 16u << wk_base(pk_sa, q + km1));
                // ---- leftovers of the lanes whose walk broke: dense insertion, 32 consecutive windows per step -----------------
                const uint32_t rem = (have && !alive) ? qe - q : 0u;
                uint32_t incl = rem;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += y;
                }
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t excl = incl - rem;
                const uint32_t qrs = q | (rs << 16);   // both below 2^14
                uint32_t carry = 0;
                for (uint32_t base = 0; base < tot; base += 32u) {
                    const uint32_t idx = base + lane;
                    const bool act = idx < tot;
                    uint32_t o = 0;   // owner: the last lane whose leftovers start at or before idx
#pragma unroll
                    for (int d = 16; d; d >>= 1) {
                        const uint32_t t = __shfl_sync(0xffffffffu, excl, o + (uint32_t)d);
                        if (t <= idx) o += (uint32_t)d;
                    }
                    const uint32_t off = idx - __shfl_sync(0xffffffffu, excl, o);
                    const uint32_t oqrs = __shfl_sync(0xffffffffu, qrs, o);
                    const uint32_t orl = __shfl_sync(0xffffffffu, rl, o);
                    const uint32_t opred = __shfl_sync(0xffffffffu, ma, o);
                    const uint32_t myq = (oqrs & 0xffffu) + off;
                    uint32_t mm = tmeta_sa;
                    bool clm = false;
                    if (act) mm = wk_insert<KW, TS, STATS>(pk_sa, tmeta_sa, tsize, K, myq, myq > (oqrs >> 16), myq < orl, clm, abort_sa, &st_probe, &st_cmp);
                    if (STATS) {
                        st_dense += act;
                        if (lane == 0) st_dsteps++;
                    }
                    uint32_t prev = __shfl_up_sync(0xffffffffu, mm, 1);
                    if (lane == 0) prev = carry;
                    if (act) {
                        const uint32_t pred = off == 0u ? opred : prev;
                        const uint32_t b = wk_base(pk_sa, myq + km1);
                        // the predecessor was walked (or is a range head): this transition's right bit is not set by its own thread
                        if (off == 0u) sred_or<-TS * 4>(pred, 16u << b);
                        // a first sighting (mostly a read error) does not take over a link that already leads somewhere
                        if (!clm || sld16<2>(pred) == WK_NOLINK) sst16<2>(pred, (mm - tmeta_sa) | b);
                    }
                    carry = __shfl_sync(0xffffffffu, mm, 31);
                }
            }
            if (tid == 0 && pf_stage == 1) pf_advance();
        }
        if (tid == 0) {
            if (cta_pool_pos + (int64_t)(TS + 1) > cta_pool_end) {
                cta_pool_pos = (int64_t)atomicAdd(a.pool_cursor, (unsigned long long)SM_POOL_CHUNK);
                cta_pool_end = cta_pool_pos + (int64_t)SM_POOL_CHUNK;
            }
            misc[4] = (uint32_t)cta_pool_pos;
            misc[5] = (uint32_t)((uint64_t)cta_pool_pos >> 32);
            misc[6] = 0;
        }
        __syncthreads();
        if (misc[3]) {   // table full
            defer(g);
            continue;
        }

        PHASE_MARK(3);
        // ---- min-coverage prune + compaction: survivors go to the pool unsorted; the k-mer is read back through the reference ------
        const int64_t pb = (int64_t)(((unsigned long long)misc[5] << 32) | misc[4]);
        if (pb + (int64_t)(TS + 1) > a.pool_cap) {
            defer(g);
            continue;
        }
        {
            constexpr int CIT = (TS / 4 + NT - 1) / NT;
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
                        for (int t = 0; t < 4; t++) {
                            if ((c[t] >> 8) >= thr) {
                                Key<KW> k = Key<KW>::zero();
                                load_window(pk, (tmeta[4u * j + t] & 0xffffu) >> 2, K, k);
                                a.pool_lo[pos] = k.lo;
                                if (KW == 2) a.pool_hi[pos] = k.hi_word();
                                a.pool_exts[pos] = (uint8_t)c[t];
                                a.pool_counts[pos] = (uint16_t)((c[t] >> 8) > 65535u ? 65535u : (c[t] >> 8));
                                pos++;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t n_out = misc[6];
        if (tid == 0) cta_pool_pos += n_out;
        if (n_out > (uint32_t)SM_TSIZE) {   // more than the ordering kernels hold: next tier
            defer(g);
            continue;
        }
        cta_obs += total;
        if (n_out == 0) continue;
        PHASE_MARK(4);
        if (tid == 0) {
            a.nsurv[g] = (int64_t)n_out;
            a.pool_start[g] = pb;
        }
        PHASE_MARK(5);
    }
    if (tid == 0 && cta_obs) atomicAdd(a.n_obs, cta_obs);
    if (STATS && a.phase_cycles) {
        const uint32_t v[8] = {st_walk, st_dense, st_head, st_probe, st_dsteps, st_wsteps, st_cmp, st_broken};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t x = v[i];
            for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
            if (lane == 0 && x) atomicAdd(a.phase_cycles + 8 + i, (unsigned long long)x);
        }
    }
}

}  // namespace rogtk
