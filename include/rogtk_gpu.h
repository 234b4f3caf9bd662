/*
 * include/rogtk_gpu.h -- C ABI of the B200-native per-UMI de Bruijn k-mer count / prune engine.
 *
 * This is the drop-in boundary for ONE path of tzeitim/rogtk: everything that
 *     debruijn::filter::filter_kmers::<K,_,_,_,_>(seq_tuples, &Box::new(CountFilter::new(min_coverage)), true, true, 4)
 * does inside analyze_dbg (reference src/fracture.rs:105-158, call at :110-116), for many groups (UMIs)
 * per call, on one B200.  The Rust host (src/fracture.rs, src/expressions.rs, src/fracture_opt.rs) keeps
 * its signatures; only the filter_kmers call is replaced by a call into this library (INTEGRATION.md shows
 * the Rust FFI stub).  No torch / Python types appear here: plain pointers and sizes only.
 *
 * Semantics per group (restated in oracle/dbg_oracle.c):
 *   - null rows are dropped                                        (src/expressions.rs:737-740)
 *   - a read is used iff every byte is one of ACGTacgt             (src/fracture.rs:218-229)
 *   - K = 4/8/16/32/64, the smallest >= k                          (src/fracture.rs:246-256)
 *   - every window of K bases is one observation; exts.left = bit(previous base) (absent at read start),
 *     exts.right = bit(next base) << 4 (absent at read end); no reverse-complement canonicalisation
 *   - per distinct k-mer: count = min(65535, #observations), exts = OR over observations
 *   - survivors: count >= min_coverage, emitted in ascending k-mer order (k-mer read as a base-4
 *     big-endian number, A<C<G<T) -- exactly the three vectors filter_kmers passes to BoomHashMap2::new.
 */
#ifndef ROGTK_GPU_H
#define ROGTK_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ROGTK_GPU_ABI_VERSION 2

/* error codes (0 = success; never throws, never aborts) */
#define ROGTK_OK 0
#define ROGTK_ERR_INVALID_ARG (-1)
#define ROGTK_ERR_CUDA (-2)
#define ROGTK_ERR_OOM (-3)
#define ROGTK_ERR_UNSUPPORTED (-4)
#define ROGTK_ERR_INTERNAL (-5)

/* string column layouts (Arrow C Data Interface buffers of the Series handed to the plugin,
 * reference src/expressions.rs:734 `inputs[0].str()`) */
#define ROGTK_UTF8_I32 0  /* Arrow Utf8: int32 offsets + one data buffer */
#define ROGTK_UTF8_I64 1  /* Arrow LargeUtf8: int64 offsets + one data buffer */
#define ROGTK_UTF8_VIEW 2 /* Arrow Utf8View (Polars >= 0.43 String): 16-byte views + N data buffers */

/* flags */
#define ROGTK_EMIT_CSR (1u << 0)         /* successor adjacency among survivors (BaseGraph r_edges, fracture.rs:343-348) */
#define ROGTK_EMIT_STATS (1u << 1)       /* node / terminal / isolated counts (fracture.rs:124-141) */
#define ROGTK_EXACT_K (1u << 2)          /* NON-REFERENCE extension: use k itself (1..64) instead of K_eff rounding */
#define ROGTK_INPUT_ON_DEVICE (1u << 3)  /* all rogtk_strcol / group_offsets pointers are device pointers */
#define ROGTK_OUTPUT_ON_DEVICE (1u << 4) /* leave the result arrays in device memory */
#define ROGTK_FORCE_SPILL (1u << 5)      /* testing: bypass the shared-memory tables, use the HBM tables only */
#define ROGTK_NO_MULTIPASS (1u << 6)     /* testing: groups beyond one shared-memory table go straight to the HBM tables */
#define ROGTK_NO_PARTITION (1u << 8)     /* testing: groups beyond the multi-pass kernel skip the partition-then-count tier (HBM tables take them) */
#define ROGTK_UNITIG_NODE_PARALLEL (1u << 7) /* testing: rogtk_gpu_compress_unitigs uses the node-parallel kernels for every batch */

typedef struct rogtk_gpu_ctx rogtk_gpu_ctx;

/* One chunk of a string column.  A multi-chunk Series is passed as several calls or concatenated by the caller. */
typedef struct rogtk_strcol {
    int32_t layout;                 /* ROGTK_UTF8_* */
    int32_t n_buffers;              /* number of data buffers (1 for I32/I64) */
    int64_t n_rows;
    const uint8_t *validity;        /* Arrow validity bitmap, LSB first, or NULL (= all valid) */
    int64_t validity_bit_offset;    /* bit index of row 0 inside validity (Arrow array offset) */
    const void *offsets_or_views;   /* int32[n_rows+1] | int64[n_rows+1] | view[n_rows] (16 B each) */
    const uint8_t *const *data;     /* data[n_buffers] */
    const int64_t *data_len;        /* byte length of each data buffer */
} rogtk_strcol;

/* Result of one batch.  All arrays are callee-allocated (pinned host memory, or device memory with
 * ROGTK_OUTPUT_ON_DEVICE) and released with rogtk_gpu_free_batch. */
typedef struct rogtk_graph_batch {
    int64_t n_groups;
    int64_t n_kmers;         /* total survivors */
    int64_t n_edges;         /* total CSR entries (0 unless ROGTK_EMIT_CSR) */
    int64_t n_observations;  /* sum over used reads of max(0, len-K+1) */
    int32_t k_eff;           /* K used (after rounding); -1 when k > 64 (empty result, fracture.rs:212-215) */
    int32_t on_device;       /* 1 if the arrays below live in device memory */
    int64_t *kmer_offsets;   /* [n_groups+1] survivors of group g are [kmer_offsets[g], kmer_offsets[g+1]) */
    uint64_t *kmers_lo;      /* [n_kmers] low 64 bits of the k-mer value, ascending inside each group */
    uint64_t *kmers_hi;      /* [n_kmers] high 64 bits; NULL unless k_eff > 32 */
    uint8_t *exts;           /* [n_kmers] debruijn::Exts byte: low nibble left, high nibble right, bit = base code */
    uint16_t *counts;        /* [n_kmers] saturating u16 */
    int64_t *csr_row_ptr;    /* [n_kmers+1] or NULL */
    int32_t *csr_col;        /* [n_edges] successor index inside the group, or NULL */
    uint32_t *stats;         /* [3*n_groups] node, terminal, isolated counts, or NULL */
    float ms_total;          /* device time of the whole pipeline (CUDA events on the context stream) */
    float ms_count;          /* device time of the counting phase (shared-memory path + HBM-table path) */
    float ms_hot_kernel;     /* device time of the k_smem_count launch alone (the dominant kernel; roofline numerator) */
    int32_t n_multipass_groups; /* groups counted in several shared-memory passes (more distinct k-mers / bases than one table holds) */
    int64_t n_spilled_groups;/* groups that went through the HBM-table path */
    float ms_sort;           /* device time of the ordering kernels (pool -> dense, ascending output) */
    float ms_csr;            /* device time of the successor CSR (0 unless ROGTK_EMIT_CSR) */
    float ms_stats;          /* device time of the node statistics (0 unless ROGTK_EMIT_STATS) */
    float _pad1;
    void *_owner;            /* internal */
} rogtk_graph_batch;

/* ---- lifecycle -------------------------------------------------------------------------------------------- */
int rogtk_gpu_abi_version(void);
int rogtk_gpu_device_count(void);
/* workspace_bytes: upper bound for transient HBM tables + pools (0 = 25% of free device memory) */
rogtk_gpu_ctx *rogtk_gpu_create(int device, size_t workspace_bytes, int *err);
void rogtk_gpu_destroy(rogtk_gpu_ctx *ctx);
/* hand the context's cached device memory (its private stream-ordered pool) and cached page-locked blocks back to the system */
int rogtk_gpu_trim(rogtk_gpu_ctx *ctx);
const char *rogtk_gpu_last_error(rogtk_gpu_ctx *ctx);

/* ---- host helpers that mirror reference host logic (no GPU needed) ------------------------------------------ */
/* src/fracture.rs:246-256; returns -1 for k > 64 */
int rogtk_gpu_k_eff(int64_t k);
/* src/fracture.rs:24-54 over the non-null rows of a (host) column with int64 offsets */
int64_t rogtk_gpu_estimate_k(const int64_t *offsets, const uint8_t *validity, int64_t n_rows);
/* UMI -> GPU shard: splitmix64(umi_id) mod n_shards */
int32_t rogtk_gpu_shard_of(uint64_t umi_id, int32_t n_shards);

/* ---- the hot path --------------------------------------------------------------------------------------------
 * Replaces, for n_groups groups at once, the filter_kmers call at reference src/fracture.rs:110-116.
 *   reads          one column chunk; rows of a group are contiguous
 *   group_offsets  [n_groups+1] row boundaries (a List(String) column's offsets, or {0,n_rows} for the legacy
 *                  one-group-per-call plugin form, src/expressions.rs:695)
 *   k, min_coverage as in AssemblyKwargs (src/expressions.rs:669-687)
 * Thread-safe: calls on one ctx are serialised internally; use one ctx per host thread for concurrency
 * (Polars calls the plugin concurrently from its rayon workers). */
int rogtk_gpu_count_prune_batch(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets,
                                int64_t n_groups, int64_t k, uint32_t min_coverage, uint32_t flags,
                                rogtk_graph_batch *out);
void rogtk_gpu_free_batch(rogtk_gpu_ctx *ctx, rogtk_graph_batch *batch);

/* ---- one host frame on several GPUs of this process (BASELINE configs[3]) --------------------------------------------
 * ctxs: one context per device (rogtk_gpu_create(device d, ..)).  Whole groups go to device rogtk_gpu_shard_of(umi_ids[g], n_ctx)
 * (umi_ids NULL: the group index is hashed); every device gets its groups as one gathered page-locked sub-column, runs
 * rogtk_gpu_count_prune_batch on its own context from its own host thread, and the results come back in ORIGINAL group order
 * (host arrays, released with rogtk_gpu_free_batch(ctxs[0], out)).  No collective, no peer access: groups are independent.
 * shard_of_group (optional, [n_groups]) receives the device every group ran on.  Offsets layouts only. */
int rogtk_gpu_count_prune_multi(rogtk_gpu_ctx *const *ctxs, int n_ctx, const rogtk_strcol *reads, const int64_t *group_offsets,
                                int64_t n_groups, const uint64_t *umi_ids, int64_t k, uint32_t min_coverage, uint32_t flags,
                                rogtk_graph_batch *out, int32_t *shard_of_group);

/* ---- threshold re-use -------------------------------------------------------------------------------------------
 * `in` is a device-resident batch (ROGTK_OUTPUT_ON_DEVICE) counted at min_coverage <= min_coverage2.  Produces the batch the
 * hot path would return for min_coverage2 without touching the reads again: what sweep_assembly_params_expr
 * (src/expressions.rs:901-941) and optimize_assembly's W/E moves (src/fracture_opt.rs:163-218) need per extra threshold. */
int rogtk_gpu_rethreshold(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *in, uint32_t min_coverage2, uint32_t flags,
                          rogtk_graph_batch *out);

/* ---- unitig compaction + contig selection ------------------------------------------------------------------------
 * The `Compression` assembly method: replaces compress_graph(true, SimpleCompress(saturating_add), graph, None) and the
 * contig loop of reference src/fracture.rs:351-383, and the plugin's min_length filter + longest-contig choice
 * (src/fracture.rs:258-277; only_largest is hard-wired to true by the expression, src/expressions.rs:742-751).
 * `in` is a device-resident batch of the hot path.  Per group the SET of unitigs (sequence, saturating-summed count, exts)
 * equals the reference's; their ORDER is by ascending first k-mer here (the reference lists them in boomphf slot order), and a
 * pure cycle is cut at its smallest k-mer.  largest_flags marks the groups where that matters for the plugin's single answer:
 * bit 0 = several contigs share the maximal length (the reference's max_by_key keeps the last one in ITS order), bit 1 = the
 * chosen contig is a pure cycle (its rotation depends on the reference's start node).  For every other group the sequence
 * of unitig `largest[g]` is byte-for-byte what assemble_sequences_expr returns. */
typedef struct rogtk_contig_batch {
    int64_t n_groups;
    int64_t n_unitigs;       /* unitigs of all groups */
    int64_t n_bases;         /* bytes in seq */
    int32_t k_eff;
    int32_t on_device;
    int64_t *unitig_offsets; /* [n_groups+1] unitigs of group g: [unitig_offsets[g], unitig_offsets[g+1]) */
    int64_t *seq_offsets;    /* [n_unitigs+1] bytes of unitig u in seq (Arrow LargeUtf8 layout) */
    uint8_t *seq;            /* ASCII A/C/G/T */
    uint16_t *counts;        /* [n_unitigs] saturating sum of the node counts (src/fracture.rs:361) */
    uint8_t *exts;           /* [n_unitigs] left nibble of the first node, right nibble of the last (after fix_exts) */
    uint8_t *cyclic;         /* [n_unitigs] 1 = pure cycle */
    int64_t *largest;        /* [n_groups] unitig the plugin returns for the group (>= min_length, longest), or -1 */
    uint8_t *largest_flags;  /* [n_groups] bit 0: tie in length, bit 1: cyclic */
    float ms_total;          /* device time (CUDA events on the context stream) */
    float _pad0;
    void *_owner;            /* internal */
} rogtk_contig_batch;

int rogtk_gpu_compress_unitigs(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *in, int64_t min_length, uint32_t flags,
                               rogtk_contig_batch *out);
void rogtk_gpu_free_contigs(rogtk_gpu_ctx *ctx, rogtk_contig_batch *batch);

/* ---- whole-expression entries -------------------------------------------------------
 * One call per frame for what the reference does with one plugin call per group.  They compose the device primitives above
 * (count once per K_eff, rogtk_gpu_rethreshold per min_coverage, unitigs on the device) with the host restatement of the
 * path-finding methods (rogtk_b200/csrc/pathfind_host.h follows src/djfind.rs line by line; its f64 weights use the host's libm
 * log, as Rust's f64::ln does).  Results come back as Arrow LargeUtf8 buffers (int64 offsets + bytes), one row per group. */
#define ROGTK_METHOD_COMPRESSION 0          /* AssemblyMethod::Compression        (src/djfind.rs:19-28) */
#define ROGTK_METHOD_SHORTEST_PATH 1        /* AssemblyMethod::ShortestPath{..}: both anchors required */
#define ROGTK_METHOD_SHORTEST_PATH_AUTO 2   /* AssemblyMethod::ShortestPathAuto */

/* AssemblyKwargs (src/expressions.rs:669-687) without the fields the reference ignores (export_graphs, prefix and the three
 * optimisation leftovers).  Anchors are NUL-terminated; with per-group anchors (assemble_sequences_with_anchors_expr,
 * src/expressions.rs:770-849: first row of inputs[1] / inputs[2]) pass anchor_offsets/_data arrays instead (LargeUtf8 layout, one
 * row per group) and leave the scalars NULL. */
typedef struct rogtk_assembly_params {
    int64_t k;
    uint32_t min_coverage;
    int32_t method;                     /* ROGTK_METHOD_* */
    const char *start_anchor;           /* NULL = None */
    const char *end_anchor;
    int64_t min_length;                 /* 0 = None */
    int32_t only_largest;               /* both expressions hard-wire Some(true) (src/expressions.rs:748); fracture_sequences passes its own */
    int32_t auto_k;                     /* k = estimate_k(group) per group (src/fracture.rs:203-209); host-resident input only */
    const int64_t *start_anchor_offsets;/* [n_groups+1] or NULL */
    const uint8_t *start_anchor_data;
    const int64_t *end_anchor_offsets;
    const uint8_t *end_anchor_data;
    const int32_t *node_rank;           /* optional, shortest-path methods: node_rank[i] = the reference's graph node id (boomphf
                                           slot order) of survivor i inside its group; NULL = ascending k-mer order.  Needs the
                                           survivor layout of a previous call with the same arguments; ties only. */
    int32_t host_threads;               /* worker threads of the host path search; 0 = all cores */
    int32_t _pad;
} rogtk_assembly_params;

/* one string per row (Arrow LargeUtf8); flags per row: bit 0 = several longest contigs of equal length, bit 1 = the contig is a
 * pure cycle (both: the reference's choice depends on boomphf order, see rogtk_contig_batch), bit 2 = the group used a k of its
 * own (auto_k) */
typedef struct rogtk_string_batch {
    int64_t n_rows;
    int64_t n_bytes;
    int64_t *offsets;        /* [n_rows+1] */
    uint8_t *data;           /* [n_bytes] */
    uint8_t *flags;          /* [n_rows] */
    float ms_device;         /* device time (CUDA events) */
    float ms_host;           /* host time of the path search / drivers */
    int64_t n_count_passes;  /* counting passes over the reads this call made (the reference makes one per probe) */
    void *_owner;            /* internal */
} rogtk_string_batch;
void rogtk_gpu_free_strings(rogtk_gpu_ctx *ctx, rogtk_string_batch *b);

/* assemble_sequences_expr / assemble_sequences_with_anchors_expr / fracture_sequences for n_groups groups at once: per group the
 * contigs joined with '\n' ("" when there is none: k > 64, no valid reads, anchors or path missing, nothing >= min_length).
 * Errors the reference raises for bad method / anchor combinations (src/expressions.rs:700-731) come back as
 * ROGTK_ERR_INVALID_ARG with the reference's message in rogtk_gpu_last_error.  Compression copies only the contigs back
 * (about 1.2 KB per UMI of the headline shape instead of the 10.5 KB of node arrays). */
int rogtk_gpu_assemble_batch(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups,
                             const rogtk_assembly_params *params, uint32_t flags, rogtk_string_batch *out);

/* sweep_assembly_params_expr (src/expressions.rs:880-955): Struct{k, min_coverage, contig_length} rows, cells in the reference's
 * order (k outer, min_coverage inner), group-major: row g * n_cells + c.  The reads are counted once per distinct K_eff of the k
 * range; every min_coverage comes from rogtk_gpu_rethreshold. */
typedef struct rogtk_sweep_params {
    int64_t k_start, k_end, k_step, cov_start, cov_end, cov_step;
    int32_t method;
    int32_t host_threads;
    const char *start_anchor, *end_anchor;
} rogtk_sweep_params;
typedef struct rogtk_sweep_result {
    int64_t n_groups, n_cells;
    int64_t *k, *min_coverage, *contig_length;   /* [n_groups * n_cells] */
    float ms_device, ms_host;
    int64_t n_count_passes;
    void *_owner;
} rogtk_sweep_result;
int rogtk_gpu_sweep_batch(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups,
                          const rogtk_sweep_params *params, uint32_t flags, rogtk_sweep_result *out);
void rogtk_gpu_free_sweep(rogtk_gpu_ctx *ctx, rogtk_sweep_result *r);

/* optimize_assembly_expr (src/fracture_opt.rs:120-367): the (k, min_coverage) hill-climb per group, Struct{contig, k, min_coverage,
 * length, input_sequences}; a group without a result gets the reference's zero-filled row (src/fracture_opt.rs:342-354).  All groups
 * advance one iteration at a time; every (K_eff, min_coverage) cell any group asks for is evaluated once for the whole batch. */
typedef struct rogtk_optimize_params {
    int64_t start_k, start_min_coverage;
    int64_t max_iterations;             /* <= 0 = the reference's default 50 */
    int32_t explore_k, prioritize_length;
    int32_t method;
    int32_t host_threads;
    const char *start_anchor, *end_anchor;   /* both required (src/fracture_opt.rs:298-302) */
} rogtk_optimize_params;
typedef struct rogtk_optimize_result {
    rogtk_string_batch contig;          /* one row per group */
    uint32_t *k, *min_coverage, *length, *input_sequences;   /* [n_groups] */
    int64_t n_cells_evaluated;          /* distinct (K_eff, min_coverage) cells computed */
    int64_t n_probes;                   /* assemble_and_check calls the reference would have made (each a full recount there) */
    void *_owner;
} rogtk_optimize_result;
int rogtk_gpu_optimize_batch(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups,
                             const rogtk_optimize_params *params, uint32_t flags, rogtk_optimize_result *out);
void rogtk_gpu_free_optimize(rogtk_gpu_ctx *ctx, rogtk_optimize_result *r);

/* host-only: the shortest-path methods on node arrays the caller already holds (one group; CSR as rogtk_graph_batch emits it with
 * csr_row_ptr rebased to the group).  Returns the number of bytes written to `out` (at most cap; the contig is n_nodes + K - 1
 * bases), 0 when there is no contig, -needed when cap is too small.  start/end NULL = shortest_path_auto. */
int64_t rogtk_host_shortest_path(int k_eff, int32_t n_nodes, const uint64_t *kmers_lo, const uint64_t *kmers_hi, const uint16_t *counts,
                                 const int64_t *csr_row_ptr, const int32_t *csr_col, const int32_t *node_rank, const char *start_anchor,
                                 const char *end_anchor, uint8_t *out, int64_t cap);

/* ---- global k-mer spectrum: one group spanning the reads of every rank (BASELINE configs[4]) -----------------------
 * The reference has no multi-process analogue: it is the un-grouped call of the same expression (one filter_kmers over all
 * reads).  Protocol per rank:
 *   1. rogtk_gpu_count_prune_batch(reads of this rank, group_offsets = {0, n_rows}, min_coverage = 1, OUTPUT_ON_DEVICE)
 *      -> ascending distinct k-mers of the local reads with partial (u16-saturated) counts and exts
 *   2. rogtk_gpu_spectrum_split: where the ascending list changes owner (owner = range of first-4-base buckets)
 *   3. all-to-all of the four arrays (NCCL; done by the host, e.g. torch.distributed.all_to_all_single)
 *   4. rogtk_gpu_spectrum_merge: sum partial counts (saturating), OR exts, apply min_coverage, ascending output (one group)
 * Rank-ordered concatenation of the results equals the single-process answer. */
int rogtk_gpu_spectrum_owner(uint64_t kmer_lo, uint64_t kmer_hi, int k_eff, int n_ranks);
int rogtk_gpu_spectrum_split(rogtk_gpu_ctx *ctx, const uint64_t *d_kmers_lo, const uint64_t *d_kmers_hi, int64_t n, int k_eff,
                             int n_ranks, int64_t *h_bounds /* [n_ranks+1] host */);
int rogtk_gpu_spectrum_merge(rogtk_gpu_ctx *ctx, const uint64_t *d_kmers_lo, const uint64_t *d_kmers_hi, const uint16_t *d_counts,
                             const uint8_t *d_exts, int64_t n, int k_eff, uint32_t min_coverage, uint32_t flags,
                             rogtk_graph_batch *out);

/* The exchange inside the library: NCCL point-to-point over NVLink / NVSwitch, reached through dlopen("libnccl.so.2") (no link-time
 * dependency; single-GPU users need no NCCL).  One communicator per context:
 *   rank 0: rogtk_gpu_nccl_unique_id(id) -> the 128 bytes reach every rank by any means (rogtk_b200/spectrum.py uses a TCP socket)
 *   all:    rogtk_gpu_nccl_init(ctx, id, rank, n_ranks) ... rogtk_gpu_nccl_finalize(ctx)
 * rogtk_gpu_spectrum_exchange = steps 2-4 above in one call: `local` is this rank's device-resident one-group batch (step 1),
 * `carry` (optional) the device-resident share of the spectrum this rank owns from earlier rounds.  Per peer ONE grouped
 * ncclSend/ncclRecv of a packed stream (k-mer words, [high words,] u16 partial counts, exts bytes: 11 / 19 bytes per record) after
 * a size exchange; then the merge.  Without a communicator (or n_ranks == 1) it is the single-rank merge. */
typedef struct rogtk_exchange_stats {
    int64_t n_sent, n_received;          /* records that left / reached this rank over NCCL */
    int64_t bytes_sent, bytes_received;  /* payload bytes over NCCL (blocks are padded to 16 bytes) */
    float ms_exchange;                   /* device time of the grouped send/recv (CUDA events on the context stream) */
    int32_t n_ranks;
} rogtk_exchange_stats;
int rogtk_gpu_nccl_unique_id(uint8_t *id128);
int rogtk_gpu_nccl_init(rogtk_gpu_ctx *ctx, const uint8_t *id128, int rank, int n_ranks);
int rogtk_gpu_nccl_finalize(rogtk_gpu_ctx *ctx);
int rogtk_gpu_spectrum_exchange(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *local, const rogtk_graph_batch *carry, uint32_t min_coverage,
                                uint32_t flags, rogtk_graph_batch *out, rogtk_exchange_stats *stats);
/* merge of several device-resident shares of the spectrum this rank owns (one-group batches as rogtk_gpu_spectrum_exchange leaves
 * them with ROGTK_OUTPUT_ON_DEVICE; at most 64) into one: partial counts add (saturating), exts OR, min_coverage on the way out.  No
 * communication.  For drivers that keep the carried share in levels (rogtk_b200/spectrum.py) instead of re-merging all of it
 * every round. */
int rogtk_gpu_spectrum_merge_runs(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *const *parts, int n_parts, uint32_t min_coverage,
                                  uint32_t flags, rogtk_graph_batch *out);

/* ---- bench / test utilities (not part of the reference surface) ------------------------------------------------ */
/* Synthetic reads of SURVEY 8(d) generated directly in HBM: n_umis groups x reads_per_umi reads of read_len bases
 * drawn from a per-UMI random template of template_len bases with iid substitution errors (err_per_64k / 65536).
 * d_data must hold n_umis*reads_per_umi*read_len bytes, d_offsets n_umis*reads_per_umi+1 int64,
 * d_group_offsets n_umis+1 int64 (all device pointers).  umi_base = global index of the first UMI. */
int rogtk_gpu_synth_reads(rogtk_gpu_ctx *ctx, uint64_t seed, int64_t umi_base, int64_t n_umis, int32_t reads_per_umi,
                          int32_t read_len, int32_t template_len, uint32_t err_per_64k, uint8_t *d_data,
                          int64_t *d_offsets, int64_t *d_group_offsets);
/* the same, reads read_base .. read_base + reads_per_umi - 1 of every template (ranks of the global-spectrum mode draw different reads
   of the same templates); d_umi_ids (device, [n_umis], or NULL) names the UMIs instead of umi_base + i (a rank's hash shard) */
int rogtk_gpu_synth_reads_from(rogtk_gpu_ctx *ctx, uint64_t seed, int64_t umi_base, int64_t n_umis, int32_t reads_per_umi, int32_t read_base,
                               int32_t read_len, int32_t template_len, uint32_t err_per_64k, uint8_t *d_data, int64_t *d_offsets,
                               int64_t *d_group_offsets, const int64_t *d_umi_ids);
/* the same generator with a read count per UMI (BASELINE configs[2], skewed groups): d_group_offsets [n_umis+1] is an input */
int rogtk_gpu_synth_reads_var(rogtk_gpu_ctx *ctx, uint64_t seed, int64_t umi_base, int64_t n_umis, const int64_t *d_group_offsets,
                              int64_t n_reads, int32_t read_len, int32_t template_len, uint32_t err_per_64k, uint8_t *d_data,
                              int64_t *d_offsets);
/* raw device/pinned memory for harnesses that must not depend on torch */
void *rogtk_gpu_device_alloc(rogtk_gpu_ctx *ctx, size_t bytes);
void rogtk_gpu_device_free(rogtk_gpu_ctx *ctx, void *p);
void *rogtk_gpu_pinned_alloc(rogtk_gpu_ctx *ctx, size_t bytes);
void rogtk_gpu_pinned_free(rogtk_gpu_ctx *ctx, void *p);
int rogtk_gpu_memcpy_h2d(rogtk_gpu_ctx *ctx, void *dst, const void *src, size_t bytes);
int rogtk_gpu_memcpy_d2h(rogtk_gpu_ctx *ctx, void *dst, const void *src, size_t bytes);
/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
int64_t rogtk_gpu_launch_count(rogtk_gpu_ctx *ctx);
/* diagnostics: enable per-phase clock64 accounting inside k_smem_count and read the cycles of the last batch
 * (0 meta, 1 ingest, 2 prefix+clear, 3 count, 4 prune+reserve, 5 compact), summed over CTAs; with the link-walk kernel
 * also event counts: 8 walked windows, 9 densely inserted windows, 10 range heads, 11 probe steps, 12 dense warp steps,
 * 13 walk warp steps, 14 k-mer compares through a reference, 15 ranges whose walk broke.  out16: 16 words */
int rogtk_gpu_debug_phase_cycles(rogtk_gpu_ctx *ctx, int enable, uint64_t *out16);
/* device-memory cache of the context: out7 = cudaMalloc calls, reuses of a cached block, idle bytes, live bytes, peak bytes, idle blocks,
   times its idle blocks were released because an allocation on the device failed */
int rogtk_gpu_debug_cache_stats(rogtk_gpu_ctx *ctx, int64_t *out7);
/* write a >L2 scratch buffer (L2 flush between timed iterations) */
int rogtk_gpu_flush_l2(rogtk_gpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
