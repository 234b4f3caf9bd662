// kernels_post.cuh -- kernels either side of the counting step: input validation, Utf8View gather (ingest), node statistics (reference src/fracture.rs:124-141) and the successor CSR that stands in
// for BaseGraph::finish()'s neighbour lookup (reference src/fracture.rs:343-348), plus the synthetic
// read generator used by bench.py / tests.
#pragma once
#include "common.cuh"

namespace rogtk {

// ---------------------------------------------------------------------------------------------------------------
// input validation: a malformed column must come back as ROGTK_ERR_INVALID_ARG, never as an out-of-bounds access
// err bits: 1 offsets decrease / negative, 2 offsets exceed the data buffer, 4 group_offsets not monotone / out of range
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_validate(ColView col, int64_t data_len, const int64_t *group_offsets, int64_t n_groups, int *err)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int e = 0;
    for (int64_t r = t; r < col.n_rows; r += stride) {
        const int64_t a = col_off(col, col.row_begin + r), b = col_off(col, col.row_begin + r + 1);
        if (a < 0 || b < a) e |= 1;
        if (b > data_len) e |= 2;
    }
    for (int64_t g = t; g < n_groups; g += stride) {
        const int64_t a = group_offsets[g], b = group_offsets[g + 1];
        if (a < col.row_begin || b < a || b > col.row_begin + col.n_rows) e |= 4;
    }
    if (e) atomicOr(err, e);
}

// ---------------------------------------------------------------------------------------------------------------
// Arrow Utf8View -> contiguous (LargeUtf8-like) column
// view = { u32 len; if len<=12: 12 inline bytes; else u32 prefix, u32 buffer_index, u32 offset }
// ---------------------------------------------------------------------------------------------------------------
struct View16 {
    uint32_t len, a, b, c;
};

// lengths of the valid rows; an out-of-line view that names a buffer the column does not have, or bytes beyond that buffer's
// length, sets *err (the call returns ROGTK_ERR_INVALID_ARG before anything is gathered)
__global__ void k_view_len(const View16 *views, int64_t n_rows, const uint8_t *validity, int64_t vbo, int n_buffers,
                           const int64_t *buf_len, int64_t *lens, int *err)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    bool ok = true;
    if (validity) {
        int64_t b = r + vbo;
        ok = (validity[b >> 3] >> (b & 7)) & 1;
    }
    const View16 v = views[r];
    if (ok && v.len > 12u) {
        // (lengths are int32 in Arrow: a negative one shows up here as >= 2^31)
        if (v.len >= 0x80000000u || v.b >= (uint32_t)n_buffers || (int64_t)v.c + (int64_t)v.len > buf_len[v.b]) {
            atomicOr(err, 8);
            ok = false;
        }
    }
    lens[r] = ok ? (int64_t)v.len : 0;
}

// One warp per row.  Out-of-line rows are copied with 16-byte loads: the source is read in aligned 16-byte chunks (the two
// chunks that straddle a destination chunk are funnel-shifted together byte-wise), the destination is written in aligned
// 16-byte chunks wherever a whole chunk belongs to the row; the ragged head and tail bytes go one by one.
__device__ __forceinline__ uint4 shift_bytes(uint4 a, uint4 b, unsigned s)   // bytes s .. s+15 of the 32 bytes a:b (s < 16)
{
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const unsigned ws = s >> 2, bs = (s & 3u) * 8u;
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {   // select w[ws + i], w[ws + i + 1] without dynamic register indexing
            if ((int)ws == j) {
                lo = w[j + i];
                hi = w[j + i + 1 < 8 ? j + i + 1 : 7];
            }
        }
        o[i] = bs ? __funnelshift_r(lo, hi, bs) : lo;
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void k_view_gather(const View16 *views, int64_t n_rows, const uint8_t *const *bufs, const int64_t *buf_len,
                              const int64_t *offsets, uint8_t *out)
{
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const unsigned lane = lane_id();
    const int64_t o = offsets[r];
    const int64_t len = offsets[r + 1] - o;   // 0 for null rows
    if (len == 0) return;
    const View16 v = views[r];
    if (v.len <= 12) {
        const uint8_t *src = reinterpret_cast<const uint8_t *>(&views[r]) + 4;
        if (lane < len) out[o + lane] = src[lane];
        return;
    }
    const uint8_t *base = bufs[v.b];
    const uint8_t *src = base + v.c;
    const int64_t blen = buf_len[v.b];
    uint8_t *dst = out + o;
    // destination chunks [d0, d1): 16-byte aligned, entirely inside the row
    const int64_t head = (int64_t)((16u - (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
    const int64_t h = head < len ? head : len;
    const int64_t nchunk = (len - h) >> 4;
    for (int64_t j = lane; j < h; j += 32) dst[j] = src[j];
    if (nchunk > 0) {
        const uint8_t *s0 = src + h;                                   // source byte of the first aligned destination chunk
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(s0) & 15u);
        const uint4 *sa = reinterpret_cast<const uint4 *>(s0 - sh);    // aligned (buffers come from cudaMalloc / the staged copy: base 256-aligned)
        const bool base_aligned = (reinterpret_cast<uintptr_t>(base) & 15u) == 0;
        for (int64_t c = lane; c < nchunk; c += 32) {
            uint4 val;
            const int64_t sb = (int64_t)(reinterpret_cast<const uint8_t *>(sa + c) - base);   // offset of the first load in the buffer
            if (base_aligned && sb >= 0 && sb + (sh ? 32 : 16) <= blen) {   // every loaded byte lies inside the buffer
                const uint4 a = sa[c];
                val = sh ? shift_bytes(a, sa[c + 1], sh) : a;
            } else {   // edge of the buffer / unaligned buffer: byte loads
                uint32_t w[4] = {0, 0, 0, 0};
                for (int t = 0; t < 16; t++) w[t >> 2] |= (uint32_t)s0[c * 16 + t] << (8 * (t & 3));
                val = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *reinterpret_cast<uint4 *>(dst + h + c * 16) = val;
        }
    }
    for (int64_t j = h + nchunk * 16 + lane; j < len; j += 32) dst[j] = src[j];
}

// ---------------------------------------------------------------------------------------------------------------
// per-group outputs derived from the dense result
// ---------------------------------------------------------------------------------------------------------------
// node / terminal / isolated counts per group (reference src/fracture.rs:124-141); one warp per group
__global__ void k_stats(int64_t n_groups, const int64_t *kmer_offsets, const uint8_t *exts, uint32_t *stats)
{
    const unsigned lane = lane_id();
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += nw) {
        const int64_t o = kmer_offsets[g], n = kmer_offsets[g + 1] - o;
        uint32_t term = 0, iso = 0;
        for (int64_t j = lane; j < n; j += 32) {
            uint32_t e = exts[o + j];
            bool l0 = (e & 0x0fu) == 0, r0 = (e & 0xf0u) == 0;
            term += (l0 || r0);
            iso += (l0 && r0);
        }
        for (int d = 16; d; d >>= 1) {
            term += __shfl_xor_sync(0xffffffffu, term, d);
            iso += __shfl_xor_sync(0xffffffffu, iso, d);
        }
        if (lane == 0) {
            stats[3 * g] = (uint32_t)n;
            stats[3 * g + 1] = term;
            stats[3 * g + 2] = iso;
        }
    }
}

// successor lookup: for survivor i and right-ext base b, j with kmers[j] == extend_right(kmers[i], b) inside the group
__device__ __forceinline__ int64_t find_in_group(const uint64_t *lo, const uint64_t *hi, int64_t n, uint64_t khi, uint64_t klo)
{
    int64_t a = 0, b = n;
    while (a < b) {
        int64_t m = (a + b) >> 1;
        uint64_t mh = hi ? hi[m] : 0ull, ml = lo[m];
        if (mh == khi && ml == klo) return m;
        if (mh < khi || (mh == khi && ml < klo)) a = m + 1; else b = m;
    }
    return -1;
}

// pass 0 (cols == nullptr): degree per node; pass 1: fill.  thread per node; group found by binary search.
__global__ void k_csr(int64_t n_kmers, int64_t n_groups, const int64_t *kmer_offsets, const uint64_t *lo, const uint64_t *hi,
                      const uint8_t *exts, int K, int32_t *degree, const int64_t *row_ptr, int32_t *cols)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_kmers) return;
    int64_t a = 0, b = n_groups;   // last g with kmer_offsets[g] <= i
    while (b - a > 1) {
        int64_t m = (a + b) >> 1;
        if (kmer_offsets[m] <= i) a = m; else b = m;
    }
    const int64_t o = kmer_offsets[a], n = kmer_offsets[a + 1] - o;
    const int kbits = 2 * K;
    const uint64_t xl = lo[i], xh = hi ? hi[i] : 0ull;
    const uint32_t e = exts[i] >> 4;
    int deg = 0;
    int64_t w = cols ? row_ptr[i] : 0;
    for (uint32_t m = e; m; m &= m - 1u) {   // ascending set bits; t-th bit per iteration keeps a warp's lanes together
        const uint32_t bb = (uint32_t)__ffs((int)m) - 1u;
        uint64_t yl = (xl << 2) | bb, yh = (xh << 2) | (xl >> 62);
        if (kbits < 64) { yl &= (1ull << kbits) - 1ull; yh = 0; }
        else if (kbits == 64) yh = 0;
        else if (kbits < 128) yh &= (1ull << (kbits - 64)) - 1ull;
        int64_t j = find_in_group(lo + o, hi ? hi + o : nullptr, n, yh, yl);
        if (j >= 0) {
            if (cols) cols[w++] = (int32_t)j;
            deg++;
        }
    }
    if (!cols) degree[i] = deg;
}


// Per-group variant: one CTA takes one group at a time and stages its ascending k-mers in shared memory together with a table of
// bucket starts (bucket = top 11 k-mer bits; the group's k-mers are sorted, so a bucket is a run of about one entry): the
// successor of a node is found by one table load and a scan of its bucket instead of a 20-step binary search for the group and a
// 10-step one for the k-mer in global memory (k_csr: ~250 instructions per lookup, issue-bound).  Groups of more than CSR_CAP
// nodes (not the per-UMI workloads) search in global memory with the same kernel.  FILL = false: edges per group (-> exclusive
// scan = every group's first CSR entry).  FILL = true: tiles of NT nodes, a block scan of the degrees gives every node its row;
// row_ptr and the successor indices are written tile by tile.
constexpr int CSR_NT = 256;
constexpr int CSR_CAP = 2048;
constexpr int CSR_BBITS = 11;
template <int W>
__device__ __forceinline__ uint32_t csr_bucket(uint64_t klo, uint64_t khi, int kbits)
{
    if (kbits <= CSR_BBITS) return (uint32_t)klo;
    const int sh = kbits - CSR_BBITS;
    if (W == 1) return (uint32_t)(klo >> sh);
    return sh >= 64 ? (uint32_t)(khi >> (sh - 64)) : (uint32_t)((khi << (64 - sh)) | (klo >> sh));
}
template <int W, bool FILL>
__global__ void __launch_bounds__(CSR_NT) k_csr_group(int64_t n_groups, const int64_t *kmer_offsets, const uint64_t *lo, const uint64_t *hi,
                                                     const uint8_t *exts, int K, int64_t *group_edges, const int64_t *group_base,
                                                     int64_t *row_ptr, int32_t *cols)
{
    __shared__ uint64_t s_lo[CSR_CAP];
    __shared__ uint64_t s_hi[W == 2 ? CSR_CAP : 1];
    __shared__ uint16_t s_start[1 << CSR_BBITS];
    __shared__ uint32_t s_ws[CSR_NT / 32 + 1];
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u, warp = tid >> 5;
    const int kbits = 2 * K;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int64_t o = kmer_offsets[g];
        const int n = (int)(kmer_offsets[g + 1] - o);
        if (n == 0) {
            if (!FILL && tid == 0) group_edges[g] = 0;
            continue;
        }
        const bool staged = n <= CSR_CAP;
        __syncthreads();   // the previous group's shared data is no longer read
        if (staged) {
            for (int j = tid; j < (1 << CSR_BBITS); j += CSR_NT) s_start[j] = 0xffffu;
            for (int j = tid; j < n; j += CSR_NT) {
                s_lo[j] = lo[o + j];
                if (W == 2) s_hi[j] = hi[o + j];
            }
            __syncthreads();
            for (int j = tid; j < n; j += CSR_NT) {   // the first node of every bucket
                const uint32_t bk = csr_bucket<W>(s_lo[j], W == 2 ? s_hi[j] : 0ull, kbits);
                if (j == 0 || csr_bucket<W>(s_lo[j - 1], W == 2 ? s_hi[j - 1] : 0ull, kbits) != bk) s_start[bk] = (uint16_t)j;
            }
            __syncthreads();
        }
        auto find = [&](uint64_t yh, uint64_t yl) -> int {
            if (!staged) return (int)find_in_group(lo + o, W == 2 ? hi + o : nullptr, n, yh, yl);
            const uint32_t bk = csr_bucket<W>(yl, yh, kbits);
            uint32_t j = s_start[bk];
            if (j == 0xffffu) return -1;
            for (; j < (uint32_t)n; j++) {   // ascending inside the bucket
                const uint64_t ml = s_lo[j], mh = W == 2 ? s_hi[j] : 0ull;
                if (mh == yh && ml == yl) return (int)j;
                if (mh > yh || (mh == yh && ml > yl)) return -1;
            }
            return -1;
        };
        int64_t run = FILL ? group_base[g] : 0;   // CSR entries before the current tile
        uint32_t total = 0;
        for (int t0 = 0; t0 < n; t0 += CSR_NT) {
            const int i = t0 + tid;
            int s0 = -1, s1 = -1, s2 = -1, s3 = -1;   // successors in base order (registers: no indexed array)
            int deg = 0;
            if (i < n) {
                const uint64_t xl = staged ? s_lo[i] : lo[o + i], xh = W == 2 ? (staged ? s_hi[i] : hi[o + i]) : 0ull;
                const uint32_t e = exts[o + i] >> 4;
#pragma unroll
                for (uint32_t bb = 0; bb < 4; bb++) {
                    if (!(e & (1u << bb))) continue;
                    uint64_t yl = (xl << 2) | bb, yh = (xh << 2) | (xl >> 62);
                    if (kbits < 64) { yl &= (1ull << kbits) - 1ull; yh = 0; }
                    else if (kbits == 64) yh = 0;
                    else if (kbits < 128) yh &= (1ull << (kbits - 64)) - 1ull;
                    const int j = find(yh, yl);
                    if (j >= 0) {
                        if (bb == 0) s0 = j; else if (bb == 1) s1 = j; else if (bb == 2) s2 = j; else s3 = j;
                        deg++;
                    }
                }
            }
            if (!FILL) {
                total += (uint32_t)deg;
                continue;
            }
            // exclusive block scan of the degrees
            uint32_t x = (uint32_t)deg;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (unsigned)d) x += y;
            }
            if (lane == 31) s_ws[warp] = x;
            __syncthreads();
            uint32_t base = 0, tile = 0;
#pragma unroll
            for (int w = 0; w < CSR_NT / 32; w++) {
                const uint32_t sw = s_ws[w];
                if (w < (int)warp) base += sw;
                tile += sw;
            }
            __syncthreads();
            if (i < n) {
                int64_t w0 = run + (int64_t)(base + x - (uint32_t)deg);
                row_ptr[o + i] = w0;
                if (s0 >= 0) cols[w0++] = s0;
                if (s1 >= 0) cols[w0++] = s1;
                if (s2 >= 0) cols[w0++] = s2;
                if (s3 >= 0) cols[w0++] = s3;
            }
            run += tile;
        }
        if (!FILL) {
            for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
            if (lane == 0) s_ws[warp] = total;
            __syncthreads();
            if (tid == 0) {
                uint32_t t = 0;
                for (int w = 0; w < CSR_NT / 32; w++) t += s_ws[w];
                group_edges[g] = (int64_t)t;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// threshold re-use: the reference re-counts all reads for every (k, min_coverage) cell of
// sweep_assembly_params_expr (src/expressions.rs:901-941) and every probe of optimize_assembly (src/fracture_opt.rs:243-263),
// although only K_eff and the threshold matter.  Pruning is a pure filter on the counts (exts are never censored),
// so a batch counted once at a low threshold yields every higher threshold exactly.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_thr_flags(const uint16_t *counts, int64_t n, uint32_t thr, int32_t *keep)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = (uint32_t)counts[i] >= thr ? 1 : 0;
}
__global__ void k_thr_offsets(const int64_t *old_off, int64_t n_groups, const int64_t *pos, int64_t *new_off)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= n_groups) new_off[g] = pos[old_off[g]];
}
__global__ void k_thr_compact(int64_t n, const int32_t *keep, const int64_t *pos, const uint64_t *lo, const uint64_t *hi,
                              const uint8_t *exts, const uint16_t *counts, uint64_t *o_lo, uint64_t *o_hi, uint8_t *o_exts,
                              uint16_t *o_counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int64_t d = pos[i];
    o_lo[d] = lo[i];
    if (o_hi) o_hi[d] = hi[i];
    o_exts[d] = exts[i];
    o_counts[d] = counts[i];
}

// ---------------------------------------------------------------------------------------------------------------
// global k-mer spectrum (one group spanning every rank).  Distinct k-mers are owned by the rank whose
// range of first-4-base buckets (the reference's bucket()) contains them, so the rank-ordered concatenation of
// the per-rank results is globally ascending like filter_kmers' bucket order.
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int spectrum_owner(uint64_t lo, uint64_t hi, int kbits, int n_ranks)
{
    uint32_t b;   // first 4 bases = top 8 bits of the k-mer (k-mers shorter than 4 bases: left-aligned)
    if (kbits > 64) b = (uint32_t)(hi >> (kbits - 64 - 8 >= 0 ? kbits - 72 : 0)) & 0xffu;
    else if (kbits >= 8) b = (uint32_t)(lo >> (kbits - 8)) & 0xffu;
    else b = (uint32_t)(lo << (8 - kbits)) & 0xffu;
    if (kbits > 64 && kbits < 72) b = (uint32_t)(((hi << (72 - kbits)) | (lo >> (kbits - 8))) & 0xffu);
    return (int)((b * (uint32_t)n_ranks) >> 8);
}

// bounds[d] = index of the first k-mer (ascending input) owned by a rank >= d, d in [0, n_ranks]
__global__ void k_spectrum_bounds(const uint64_t *lo, const uint64_t *hi, int64_t n, int kbits, int n_ranks, int64_t *bounds)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > n_ranks) return;
    int64_t a = 0, b = n;
    while (a < b) {
        const int64_t m = (a + b) >> 1;
        if (spectrum_owner(lo[m], hi ? hi[m] : 0ull, kbits, n_ranks) < d) a = m + 1; else b = m;
    }
    bounds[d] = a;
}

struct SpecRec {   // (k-mer, partial count, exts) received from one source rank
    uint64_t hi, lo;
    uint32_t val;   // count (low 16, saturated) | exts << 16
    uint32_t pad;
};
struct SpecRecLess {
    __device__ __forceinline__ bool operator()(const SpecRec &a, const SpecRec &b) const
    {
        return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
    }
};
__global__ void k_spec_pack(const uint64_t *lo, const uint64_t *hi, const uint16_t *counts, const uint8_t *exts, int64_t n, SpecRec *r)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    r[i] = SpecRec{hi ? hi[i] : 0ull, lo[i], (uint32_t)counts[i] | ((uint32_t)exts[i] << 16), 0u};
}
// equal k-mers are adjacent after the sort: the first of each run sums the partial counts (saturating u16) and ORs the exts
__global__ void k_spec_reduce(SpecRec *r, int64_t n, uint32_t min_cov, int32_t *keep)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SpecRec me = r[i];
    int k = 0;
    if (i == 0 || r[i - 1].hi != me.hi || r[i - 1].lo != me.lo) {
        uint32_t c = me.val & 0xffffu, e = me.val >> 16;
        for (int64_t j = i + 1; j < n && r[j].hi == me.hi && r[j].lo == me.lo; j++) {
            c += r[j].val & 0xffffu;
            e |= r[j].val >> 16;
        }
        if (c > 65535u) c = 65535u;
        r[i].val = c | (e << 16);
        k = (c >= min_cov && c > 0u) ? 1 : 0;
    }
    keep[i] = k;
}
__global__ void k_spec_emit(const SpecRec *r, int64_t n, const int32_t *keep, const int64_t *pos, uint64_t *o_lo, uint64_t *o_hi,
                            uint8_t *o_exts, uint16_t *o_counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int64_t d = pos[i];
    o_lo[d] = r[i].lo;
    if (o_hi) o_hi[d] = r[i].hi;
    o_exts[d] = (uint8_t)(r[i].val >> 16);
    o_counts[d] = (uint16_t)(r[i].val & 0xffffu);
}

// ---------------------------------------------------------------------------------------------------------------
// synthetic reads: per UMI u (global index) s_u = splitmix_out(seed, u);
//   template base i      = (splitmix_out(s_u, i>>5) >> 2*(i&31)) & 3
//   read r start         = splitmix_out(s_u, 2^20 + 4096 r) mod (T-L+1)
//   read r base j error  : v = 16-bit field (j&3) of splitmix_out(s_u, 2^20 + 4096 r + 1 + (j>>2));
//                          if v < err_per_64k: base = (base + 1 + v%3) & 3
// The same function in numpy: rogtk_b200/synth.py.
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint8_t synth_base(uint64_t seed, int64_t umi, int32_t r, int32_t j, int32_t L, int32_t T,
                                                       uint32_t err)
{
    const uint64_t su = splitmix_out(seed, (uint64_t)umi);
    const uint64_t rc = (1ull << 20) + 4096ull * (uint64_t)r;
    const uint32_t start = (uint32_t)(splitmix_out(su, rc) % (uint64_t)(T - L + 1));
    const uint32_t i = start + (uint32_t)j;
    uint32_t b = (uint32_t)(splitmix_out(su, i >> 5) >> (2 * (i & 31))) & 3u;
    const uint32_t v = (uint32_t)(splitmix_out(su, rc + 1 + (uint64_t)(j >> 2)) >> (16 * (j & 3))) & 0xffffu;
    if (v < err) b = (b + 1 + v % 3u) & 3u;
    return (uint8_t)("ACGT"[b]);
}

// the same generator with a per-UMI read count (BASELINE configs[2]: skewed groups): group_offsets is given, one thread per 4 bases
__global__ void k_synth_var(uint64_t seed, int64_t umi_base, int64_t n_umis, const int64_t *group_offsets, int32_t L, int32_t T, uint32_t err,
                            uint8_t *data, int64_t *offsets)
{
    const int64_t n_reads = group_offsets[n_umis];
    const int64_t total = n_reads * (int64_t)L;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= n_reads) offsets[t] = t * (int64_t)L;
    const int64_t q0 = t * 4;
    if (q0 >= total) return;
    uint8_t out[4] = {0, 0, 0, 0};
    int64_t cur_read = -1, u = 0;
    for (int x = 0; x < 4; x++) {
        const int64_t q = q0 + x;
        if (q >= total) break;
        const int64_t read = q / L;
        if (read != cur_read) {   // last u with group_offsets[u] <= read
            int64_t a = 0, b = n_umis;
            while (b - a > 1) {
                const int64_t m = (a + b) >> 1;
                if (group_offsets[m] <= read) a = m; else b = m;
            }
            u = a;
            cur_read = read;
        }
        out[x] = synth_base(seed, umi_base + u, (int32_t)(read - group_offsets[u]), (int32_t)(q - read * L), L, T, err);
    }
    for (int x = 0; x < 4 && q0 + x < total; x++) data[q0 + x] = out[x];
}

__global__ void k_synth(uint64_t seed, int64_t umi_base, int64_t n_umis, int32_t R, int32_t L, int32_t T, uint32_t err,
                        uint8_t *data, int64_t *offsets, int64_t *group_offsets, int32_t read_base, const int64_t *umi_ids)
{
    const int64_t total = n_umis * (int64_t)R * L;
    const int64_t q0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int64_t n_reads = n_umis * (int64_t)R;
    // offsets / group offsets (threads double as writers for the small arrays)
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= n_reads) offsets[t] = t * (int64_t)L;
    if (t <= n_umis) group_offsets[t] = t * (int64_t)R;
    if (q0 >= total) return;
    uint8_t out[4];
#pragma unroll
    for (int x = 0; x < 4; x++) {
        int64_t q = q0 + x;
        if (q < total) {
            int64_t read = q / L;
            int32_t j = (int32_t)(q - read * L);
            int64_t u = read / R;
            int32_t r = (int32_t)(read - u * R);
            out[x] = synth_base(seed, umi_ids ? umi_ids[u] : umi_base + u, r + read_base, j, L, T, err);
        } else {
            out[x] = 0;
        }
    }
    if (q0 + 3 < total) {
        *reinterpret_cast<uint32_t *>(data + q0) = out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
    } else {
        for (int x = 0; x < 4 && q0 + x < total; x++) data[q0 + x] = out[x];
    }
}

}  // namespace rogtk
