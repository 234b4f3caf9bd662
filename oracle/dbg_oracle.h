/*
 * oracle/dbg_oracle.h -- CPU restatement of rogtk's per-UMI de Bruijn k-mer count/prune path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rogtk_b200/ (the product) may include, link, import or
 * execute this code.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at the k-mer level.  The arithmetic restated here lives in the
 * un-vendored crates debruijn ^0.3.4 (Cargo.toml:51) and boomphf ^0.5 (Cargo.toml:55); their source is
 * not in /root/reference and no Rust toolchain exists in this image, so the restatement follows the
 * published algorithm of debruijn::filter::filter_kmers + CountFilter and is
 * anchored on the reference's own call sites (src/fracture.rs:110-116, :218-229, :238-256) and on the
 * contig-level expectations of the reference's unit tests (src/fracture.rs:593-679), see
 * tests/test_oracle.py.  Two independent implementations (this sort-based one and the dict-based one in
 * oracle/dbg_oracle.py) must agree bit-for-bit before either is trusted.
 */
#ifndef ROGTK_DBG_ORACLE_H
#define ROGTK_DBG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_batch {
    int64_t n_groups;
    int64_t n_kmers;         /* total surviving k-mers over all groups */
    int64_t n_observations;  /* sum over valid reads of max(0, len - K + 1) */
    int32_t k_eff;           /* K actually used */
    int32_t _pad;
    int64_t *kmer_offsets;   /* n_groups + 1 */
    uint64_t *kmers_lo;      /* low 64 bits of each k-mer value (base-4 big-endian number) */
    uint64_t *kmers_hi;      /* high 64 bits; NULL unless K > 32 */
    uint8_t *exts;           /* low nibble = left neighbour bits, high nibble = right, bit = base code */
    uint16_t *counts;        /* saturating u16 */
    uint32_t *stats;         /* 3 per group: node_count, terminal_count, isolated_count */
    int64_t *csr_row_ptr;    /* n_kmers + 1 (global), successors among survivors of the same group */
    int32_t *csr_col;        /* index of the successor inside its group */
} oracle_batch;

/* src/fracture.rs:246-256 -- the type dispatch: K_eff = min{4,8,16,32,64} >= k ; -1 if k > 64 */
int oracle_k_eff(int64_t k);

/* src/fracture.rs:24-54 -- estimate_k over rows [0,n_rows) (nulls are removed by the caller,
 * src/expressions.rs:737-740, so validity==NULL rows are all present) */
int64_t oracle_estimate_k(const int64_t *offsets, const uint8_t *validity, int64_t n_rows);

/*
 * filter_kmers + CountFilter for many groups (src/fracture.rs:105-158 applied per group).
 *   data/offsets : Arrow LargeUtf8-style buffers (int64 offsets, n_rows + 1)
 *   validity     : Arrow validity bitmap (LSB first) or NULL; null rows are dropped (expressions.rs:737-740)
 *   group_offsets: n_groups + 1 row boundaries; rows of one group are contiguous
 *   k            : user k; rounded with oracle_k_eff unless exact_k != 0 (non-reference extension)
 *   n_threads    : <=0 -> all cores (one group per task, dynamic schedule: Polars' per-group rayon fan-out)
 * returns 0, or -1 on bad arguments (k > 64 in exact mode, k < 1 in exact mode, ...).
 * k > 64 in reference mode returns an empty batch (src/fracture.rs:212-215).
 */
int oracle_count_prune_batch(const uint8_t *data, const int64_t *offsets, const uint8_t *validity,
                             int64_t n_rows, const int64_t *group_offsets, int64_t n_groups,
                             int64_t k, uint32_t min_coverage, int exact_k, int emit_csr, int n_threads,
                             oracle_batch *out);

void oracle_free_batch(oracle_batch *b);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
