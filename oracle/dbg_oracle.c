/*
 * oracle/dbg_oracle.c -- CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY, see
 * dbg_oracle.h for the rules and the "parity unpinned" statement).
 *
 * What it restates, per group (one Polars group == one plugin call, src/expressions.rs:695-762):
 *   - null rows dropped                                   src/expressions.rs:737-740
 *   - to_uppercase + "all bytes in ACGT else drop read"   src/fracture.rs:218-229
 *   - DnaString::from_dna_string 2-bit packing            src/fracture.rs:223
 *   - (DnaString, Exts::empty(), ())                      src/fracture.rs:238-240
 *   - k -> Kmer4/8/16/32/64 dispatch                      src/fracture.rs:246-256
 *   - filter_kmers + CountFilter                          src/fracture.rs:110-116
 *   - terminal / isolated statistics                      src/fracture.rs:124-141
 *   - successor lookup of BaseGraph (r_edges after fix_exts), SURVEY A.7/A.8
 */
#include "dbg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    uint64_t *words; /* DnaString storage */
    int64_t len;
} packed_read;

typedef struct {
    int64_t n, cap, n_obs;
    uint64_t *lo, *hi;
    uint8_t *exts;
    uint16_t *counts;
} group_result;

typedef struct {
    void *obs32, *tmp32, *obs64, *tmp64, *obs128, *tmp128;
    size_t obs_cap32, obs_cap64, obs_cap128;
    uint64_t *words;
    size_t words_cap;
    packed_read *reads;
    size_t reads_cap;
} scratch_t;

static void result_push(group_result *r, uint64_t lo, uint64_t hi, uint8_t e, uint16_t c)
{
    if (r->n == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 256;
        r->lo = realloc(r->lo, (size_t)r->cap * 8);
        r->hi = realloc(r->hi, (size_t)r->cap * 8);
        r->exts = realloc(r->exts, (size_t)r->cap);
        r->counts = realloc(r->counts, (size_t)r->cap * 2);
    }
    r->lo[r->n] = lo;
    r->hi[r->n] = hi;
    r->exts[r->n] = e;
    r->counts[r->n] = c;
    r->n++;
}

#define KMER_T uint32_t
#define SFX 32
#include "dbg_oracle_impl.inc"
#undef KMER_T
#undef SFX
#define KMER_T uint64_t
#define SFX 64
#include "dbg_oracle_impl.inc"
#undef KMER_T
#undef SFX
#define KMER_T unsigned __int128
#define SFX 128
#include "dbg_oracle_impl.inc"
#undef KMER_T
#undef SFX

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int oracle_k_eff(int64_t k)
{
    /* src/fracture.rs:246-256 */
    if (k < 0) return -1;
    if (k <= 4) return 4;
    if (k <= 8) return 8;
    if (k <= 16) return 16;
    if (k <= 32) return 32;
    if (k <= 64) return 64;
    return -1;
}

int64_t oracle_estimate_k(const int64_t *offsets, const uint8_t *validity, int64_t n_rows)
{
    /* src/fracture.rs:24-54.  Nulls never reach estimate_k (expressions.rs:737-740). */
    uint64_t total = 0, count = 0;
    for (int64_t r = 0; r < n_rows; r++) {
        if (validity && !((validity[r >> 3] >> (r & 7)) & 1)) continue;
        int64_t len = offsets[r + 1] - offsets[r];
        if (len > 0) {
            total += (uint64_t)len;
            count++;
        }
    }
    if (count == 0) return 31;
    double mean = (double)total / (double)count;
    /* f64::round = half away from zero; values are positive */
    uint64_t k = (uint64_t)round(mean / 3.0);
    /* `k - 1` on usize 0 wraps in release builds -> clamp gives 63 */
    if (k % 2 == 0) k = k - 1;
    if (k < 11) return 11;
    if (k > 63) return 63;
    return (int64_t)k;
}

/* base code per SURVEY A.1; 0xff = invalid.  to_uppercase() makes lower-case ACGT valid. */
static inline int base_code(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

static void process_group(const uint8_t *data, const int64_t *offsets, const uint8_t *validity,
                          int64_t row_lo, int64_t row_hi, int K, uint32_t min_cov, group_result *res,
                          scratch_t *sc)
{
    /* 1. validate + pack (src/fracture.rs:218-229) */
    size_t need_words = 0;
    for (int64_t r = row_lo; r < row_hi; r++) need_words += (size_t)((offsets[r + 1] - offsets[r] + 31) / 32) + 1;
    if (need_words > sc->words_cap) {
        sc->words_cap = need_words + need_words / 4 + 64;
        sc->words = realloc(sc->words, sc->words_cap * 8);
    }
    if ((size_t)(row_hi - row_lo) > sc->reads_cap) {
        sc->reads_cap = (size_t)(row_hi - row_lo) * 2 + 16;
        sc->reads = realloc(sc->reads, sc->reads_cap * sizeof(packed_read));
    }
    int64_t n_reads = 0;
    size_t wpos = 0;
    for (int64_t r = row_lo; r < row_hi; r++) {
        if (validity && !((validity[r >> 3] >> (r & 7)) & 1)) continue; /* null row */
        const uint8_t *s = data + offsets[r];
        const int64_t len = offsets[r + 1] - offsets[r];
        uint64_t *w = sc->words + wpos;
        const int64_t nw = (len + 31) / 32;
        memset(w, 0, (size_t)nw * 8);
        int ok = 1;
        for (int64_t i = 0; i < len; i++) {
            int c = base_code(s[i]);
            if (c < 0) {
                ok = 0;
                break;
            }
            w[i >> 5] |= (uint64_t)c << (62 - 2 * (i & 31));
        }
        if (!ok) continue; /* "Skipping sequence with invalid characters" */
        sc->reads[n_reads].words = w;
        sc->reads[n_reads].len = len;
        n_reads++;
        wpos += (size_t)nw;
    }
    /* an empty-after-filtering group returns no contigs before any k-mer work (fracture.rs:233-236) */
    if (n_reads == 0) return;
    if (K <= 16)
        filter_kmers32(sc->reads, n_reads, K, min_cov, res, sc);
    else if (K <= 32)
        filter_kmers64(sc->reads, n_reads, K, min_cov, res, sc);
    else
        filter_kmers128(sc->reads, n_reads, K, min_cov, res, sc);
}

static int cmp128(uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo)
{
    if (ahi != bhi) return ahi < bhi ? -1 : 1;
    if (alo != blo) return alo < blo ? -1 : 1;
    return 0;
}

/* index of (hi,lo) inside a group's ascending survivor list, or -1 */
static int64_t find_kmer(const uint64_t *lo, const uint64_t *hi, int64_t n, uint64_t khi, uint64_t klo)
{
    int64_t a = 0, b = n;
    while (a < b) {
        int64_t m = (a + b) / 2;
        int c = cmp128(hi ? hi[m] : 0, lo[m], khi, klo);
        if (c == 0) return m;
        if (c < 0) a = m + 1; else b = m;
    }
    return -1;
}

int oracle_count_prune_batch(const uint8_t *data, const int64_t *offsets, const uint8_t *validity,
                             int64_t n_rows, const int64_t *group_offsets, int64_t n_groups, int64_t k,
                             uint32_t min_coverage, int exact_k, int emit_csr, int n_threads,
                             oracle_batch *out)
{
    memset(out, 0, sizeof(*out));
    if (n_groups < 0 || n_rows < 0) return -1;
    int K;
    if (exact_k) {
        if (k < 1 || k > 64) return -1;
        K = (int)k;
    } else {
        K = oracle_k_eff(k);
    }
    out->n_groups = n_groups;
    out->k_eff = K;
    out->kmer_offsets = calloc((size_t)n_groups + 1, 8);
    out->stats = calloc((size_t)n_groups * 3 + 1, 4);
    if (K < 0) { /* k > 64: Ok(Vec::new()) for every group (src/fracture.rs:212-215) */
        out->kmers_lo = malloc(8);
        out->exts = malloc(1);
        out->counts = malloc(2);
        out->csr_row_ptr = calloc(1, 8);
        out->csr_col = malloc(4);
        return 0;
    }
    group_result *res = calloc((size_t)n_groups + 1, sizeof(group_result));
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
#pragma omp parallel num_threads(n_threads)
    {
        scratch_t sc;
        memset(&sc, 0, sizeof(sc));
#pragma omp for schedule(dynamic, 8)
        for (int64_t g = 0; g < n_groups; g++)
            process_group(data, offsets, validity, group_offsets[g], group_offsets[g + 1], K, min_coverage,
                          &res[g], &sc);
        free(sc.obs32); free(sc.tmp32); free(sc.obs64); free(sc.tmp64); free(sc.obs128); free(sc.tmp128);
        free(sc.words); free(sc.reads);
    }
    int64_t total = 0, nobs = 0;
    for (int64_t g = 0; g < n_groups; g++) {
        out->kmer_offsets[g] = total;
        total += res[g].n;
        nobs += res[g].n_obs;
    }
    out->kmer_offsets[n_groups] = total;
    out->n_kmers = total;
    out->n_observations = nobs;
    out->kmers_lo = malloc((size_t)total * 8 + 8);
    out->kmers_hi = K > 32 ? malloc((size_t)total * 8 + 8) : NULL;
    out->exts = malloc((size_t)total + 1);
    out->counts = malloc((size_t)total * 2 + 2);
    for (int64_t g = 0; g < n_groups; g++) {
        const int64_t o = out->kmer_offsets[g], n = res[g].n;
        if (n) {
            memcpy(out->kmers_lo + o, res[g].lo, (size_t)n * 8);
            if (out->kmers_hi) memcpy(out->kmers_hi + o, res[g].hi, (size_t)n * 8);
            memcpy(out->exts + o, res[g].exts, (size_t)n);
            memcpy(out->counts + o, res[g].counts, (size_t)n * 2);
        }
        /* src/fracture.rs:124-141 */
        uint32_t term = 0, iso = 0;
        for (int64_t i = 0; i < n; i++) {
            uint8_t e = res[g].exts[i];
            int l0 = (e & 0x0f) == 0, r0 = (e & 0xf0) == 0;
            term += (uint32_t)(l0 || r0);
            iso += (uint32_t)(l0 && r0);
        }
        out->stats[3 * g] = (uint32_t)n;
        out->stats[3 * g + 1] = term;
        out->stats[3 * g + 2] = iso;
        free(res[g].lo); free(res[g].hi); free(res[g].exts); free(res[g].counts);
    }
    free(res);
    if (emit_csr) {
        /* successors among survivors: kmer.extend_right(b) for each right bit b ascending */
        out->csr_row_ptr = calloc((size_t)total + 1, 8);
        out->csr_col = malloc((size_t)total * 4 * 4 + 4);
        int64_t ne = 0;
        const int kbits = 2 * K;
        for (int64_t g = 0; g < n_groups; g++) {
            const int64_t o = out->kmer_offsets[g], n = out->kmer_offsets[g + 1] - o;
            const uint64_t *lo = out->kmers_lo + o;
            const uint64_t *hi = out->kmers_hi ? out->kmers_hi + o : NULL;
            for (int64_t i = 0; i < n; i++) {
                out->csr_row_ptr[o + i] = ne;
                unsigned __int128 x = ((unsigned __int128)(hi ? hi[i] : 0) << 64) | lo[i];
                for (int b = 0; b < 4; b++) {
                    if (!((out->exts[o + i] >> (4 + b)) & 1)) continue;
                    unsigned __int128 y = (x << 2) | (unsigned)b;
                    if (kbits < 128) y &= (((unsigned __int128)1) << kbits) - 1;
                    int64_t j = find_kmer(lo, hi, n, (uint64_t)(y >> 64), (uint64_t)y);
                    if (j >= 0) out->csr_col[ne++] = (int32_t)j;
                }
            }
        }
        out->csr_row_ptr[total] = ne;
    }
    return 0;
}

void oracle_free_batch(oracle_batch *b)
{
    if (!b) return;
    free(b->kmer_offsets); free(b->kmers_lo); free(b->kmers_hi); free(b->exts); free(b->counts);
    free(b->stats); free(b->csr_row_ptr); free(b->csr_col);
    memset(b, 0, sizeof(*b));
}
