// multi_host.cuh -- one host frame on several GPUs of one process (BASELINE configs[3]: "100M reads UMI-hash-sharded across
// 2/4/8 B200, per-UMI pruning, no collective").  Included at the end of rogtk_gpu.cu.
//
// Groups are independent (every plugin call of the reference sees one group; nothing in reference src/fracture.rs crosses groups), so
// the frame is cut by rogtk_gpu_shard_of(umi_id) while it is staged: every device receives whole groups as one contiguous
// page-locked sub-column (gathered by host threads), runs the ordinary batch path on its own context and stream from its own host
// thread, and the per-device results are scattered back to the original group order.  No device ever talks to another one.
#pragma once

namespace {

struct SubColumn {   // the groups of one device, gathered into page-locked Arrow buffers
    std::vector<int64_t> groups;          // original indices, ascending
    uint8_t *data = nullptr;
    int64_t *offs = nullptr, *goff = nullptr;
    uint8_t *valid = nullptr;
    int64_t n_rows = 0, n_bytes = 0;
};

}  // namespace

extern "C" int rogtk_gpu_count_prune_multi(rogtk_gpu_ctx *const *ctxs, int n_ctx, const rogtk_strcol *reads, const int64_t *group_offsets,
                                           int64_t n_groups, const uint64_t *umi_ids, int64_t k, uint32_t min_coverage, uint32_t flags,
                                           rogtk_graph_batch *out, int32_t *shard_of_group)
{
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return ROGTK_ERR_INVALID_ARG;
    rogtk_gpu_ctx *c0 = ctxs[0];
    if (!reads || !out || n_groups < 0 || (n_groups > 0 && !group_offsets)) {
        set_error(c0, "null or negative argument");
        return ROGTK_ERR_INVALID_ARG;
    }
    for (int d = 0; d < n_ctx; d++)
        if (!ctxs[d]) {
            set_error(c0, "null context in the list");
            return ROGTK_ERR_INVALID_ARG;
        }
    memset(out, 0, sizeof(*out));
    int rc = ROGTK_OK;
    std::vector<SubColumn> sub((size_t)n_ctx);
    std::vector<rogtk_graph_batch> part((size_t)n_ctx);
    std::vector<char> part_live((size_t)n_ctx, 0);
    BatchOwner *own = new BatchOwner();
    out->_owner = own;
    try {
        if (flags & (ROGTK_INPUT_ON_DEVICE | ROGTK_OUTPUT_ON_DEVICE)) throw ArgFail{ROGTK_ERR_UNSUPPORTED, "the multi-GPU entry takes and returns host buffers"};
        if (reads->layout != ROGTK_UTF8_I32 && reads->layout != ROGTK_UTF8_I64)
            throw ArgFail{ROGTK_ERR_UNSUPPORTED, "the multi-GPU entry takes the offsets layouts (cast a view column first)"};
        if (reads->n_buffers != 1 || !reads->data || !reads->data_len || !reads->offsets_or_views)
            throw ArgFail{ROGTK_ERR_INVALID_ARG, "offset layouts need exactly one data buffer"};
        const bool o64 = reads->layout == ROGTK_UTF8_I64;
        auto off_at = [&](int64_t row) -> int64_t {
            return o64 ? ((const int64_t *)reads->offsets_or_views)[row] : (int64_t)((const int32_t *)reads->offsets_or_views)[row];
        };
        const int nt = c0->stage_threads;
        // the host indexes with every offset it copies by: check them all here (rows and groups)
        {
            int64_t prev = n_groups ? group_offsets[0] : 0;
            bool ok = prev >= 0;
            for (int64_t g = 1; g <= n_groups && ok; g++) {
                ok = group_offsets[g] >= prev;
                prev = group_offsets[g];
            }
            if (!ok || prev > reads->n_rows) throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: group_offsets not monotone or outside the rows;"};
            std::atomic<int> bad{0};
            const int64_t lim = reads->data_len[0];
            host_parallel(reads->n_rows, nt, 1 << 16, [&](int64_t r0, int64_t r1) {
                for (int64_t r = r0; r < r1; r++) {
                    const int64_t a = off_at(r), b = off_at(r + 1);
                    if (a < 0 || b < a || b > lim) bad = 1;
                }
            });
            if (bad) throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: string offsets not monotone or beyond the data buffer;"};
        }
        // ---- shard: whole groups by hash of their UMI id -----------------------------------------------------------------------------
        std::vector<int32_t> shard((size_t)n_groups);
        host_parallel(n_groups, nt, 1 << 16, [&](int64_t g0, int64_t g1) {
            for (int64_t g = g0; g < g1; g++) shard[(size_t)g] = rogtk_gpu_shard_of(umi_ids ? umi_ids[g] : (uint64_t)g, n_ctx);
        });
        if (shard_of_group) memcpy(shard_of_group, shard.data(), (size_t)n_groups * 4);
        for (int64_t g = 0; g < n_groups; g++) sub[(size_t)shard[(size_t)g]].groups.push_back(g);
        // ---- per device: gather -> count/prune on its own context, all devices at once -----------------------------------------------
        std::vector<std::thread> th;
        std::vector<int> drc((size_t)n_ctx, ROGTK_OK);
        std::vector<std::string> derr((size_t)n_ctx);
        const int nt_dev = std::max(1, nt / n_ctx);
        for (int d = 0; d < n_ctx; d++)
            th.emplace_back([&, d]() {
                rogtk_gpu_ctx *cx = ctxs[d];
                SubColumn &sc = sub[(size_t)d];
                try {
                    const int64_t ng = (int64_t)sc.groups.size();
                    std::vector<int64_t> rpre((size_t)ng + 1, 0), bpre((size_t)ng + 1, 0);
                    for (int64_t i = 0; i < ng; i++) {
                        const int64_t g = sc.groups[(size_t)i];
                        rpre[(size_t)i + 1] = rpre[(size_t)i] + (group_offsets[g + 1] - group_offsets[g]);
                        bpre[(size_t)i + 1] = bpre[(size_t)i] + (off_at(group_offsets[g + 1]) - off_at(group_offsets[g]));
                    }
                    sc.n_rows = rpre[(size_t)ng];
                    sc.n_bytes = bpre[(size_t)ng];
                    sc.data = (uint8_t *)pinned_get(cx, (size_t)sc.n_bytes + 32);
                    sc.offs = (int64_t *)pinned_get(cx, (size_t)(sc.n_rows + 1) * 8);
                    sc.goff = (int64_t *)pinned_get(cx, (size_t)(ng + 1) * 8);
                    if (reads->validity) {
                        sc.valid = (uint8_t *)pinned_get(cx, (size_t)(sc.n_rows + 7) / 8 + 8);
                        memset(sc.valid, 0, (size_t)(sc.n_rows + 7) / 8 + 8);
                    }
                    memcpy(sc.goff, rpre.data(), (size_t)(ng + 1) * 8);
                    host_parallel(ng, nt_dev, 256, [&](int64_t i0, int64_t i1) {
                        for (int64_t i = i0; i < i1; i++) {
                            const int64_t g = sc.groups[(size_t)i];
                            const int64_t r0 = group_offsets[g], r1 = group_offsets[g + 1];
                            const int64_t b0 = off_at(r0), b1 = off_at(r1);
                            memcpy(sc.data + bpre[(size_t)i], reads->data[0] + b0, (size_t)(b1 - b0));
                            for (int64_t r = r0; r <= r1; r++) sc.offs[rpre[(size_t)i] + (r - r0)] = bpre[(size_t)i] + (off_at(r) - b0);
                        }
                    });
                    if (reads->validity) {   // (bits of neighbouring groups share bytes: one thread)
                        for (int64_t i = 0; i < ng; i++) {
                            const int64_t g = sc.groups[(size_t)i];
                            for (int64_t r = group_offsets[g]; r < group_offsets[g + 1]; r++) {
                                const int64_t b = r + reads->validity_bit_offset;
                                if ((reads->validity[b >> 3] >> (b & 7)) & 1) {
                                    const int64_t q = rpre[(size_t)i] + (r - group_offsets[g]);
                                    sc.valid[q >> 3] |= (uint8_t)(1u << (q & 7));
                                }
                            }
                        }
                    }
                    rogtk_strcol col{};
                    col.layout = ROGTK_UTF8_I64;
                    col.n_buffers = 1;
                    col.n_rows = sc.n_rows;
                    col.validity = sc.valid;
                    col.offsets_or_views = sc.offs;
                    const uint8_t *dp = sc.data;
                    const int64_t dl = sc.n_bytes;
                    col.data = &dp;
                    col.data_len = &dl;
                    drc[(size_t)d] = rogtk_gpu_count_prune_batch(cx, &col, sc.goff, ng, k, min_coverage, flags, &part[(size_t)d]);
                    if (drc[(size_t)d] == ROGTK_OK) part_live[(size_t)d] = 1;
                    else derr[(size_t)d] = cx->last_error;
                } catch (...) {
                    drc[(size_t)d] = fail_code(cx);
                    derr[(size_t)d] = cx->last_error;
                }
            });
        for (auto &t : th) t.join();
        for (int d = 0; d < n_ctx; d++)
            if (drc[(size_t)d] != ROGTK_OK) throw ArgFail{drc[(size_t)d], "device " + std::to_string(d) + ": " + derr[(size_t)d]};
        // ---- results back in the original group order -----------------------------------------------------------------------------
        const int K = part[0].k_eff;
        const bool has_csr = flags & ROGTK_EMIT_CSR, has_stats = flags & ROGTK_EMIT_STATS;
        std::vector<int64_t> local_of((size_t)n_groups);
        for (int d = 0; d < n_ctx; d++)
            for (size_t i = 0; i < sub[(size_t)d].groups.size(); i++) local_of[(size_t)sub[(size_t)d].groups[i]] = (int64_t)i;
        int64_t *f_off = (int64_t *)pinned_get(c0, (size_t)(n_groups + 1) * 8);
        own->ptrs.push_back(f_off);
        std::vector<int64_t> e_off(has_csr ? (size_t)n_groups + 1 : 1, 0);
        f_off[0] = 0;
        for (int64_t g = 0; g < n_groups; g++) {
            const rogtk_graph_batch &p = part[(size_t)shard[(size_t)g]];
            const int64_t i = local_of[(size_t)g];
            const int64_t n = K < 0 ? 0 : p.kmer_offsets[i + 1] - p.kmer_offsets[i];
            f_off[g + 1] = f_off[g] + n;
            if (has_csr && K >= 0) e_off[(size_t)g + 1] = e_off[(size_t)g] + (p.csr_row_ptr[p.kmer_offsets[i + 1]] - p.csr_row_ptr[p.kmer_offsets[i]]);
        }
        const int64_t nk = f_off[n_groups], ne = has_csr ? e_off[(size_t)n_groups] : 0;
        auto alloc = [&](size_t bytes) -> void * {
            void *q = pinned_get(c0, std::max<size_t>(bytes, 16));
            own->ptrs.push_back(q);
            return q;
        };
        out->n_groups = n_groups;
        out->n_kmers = nk;
        out->n_edges = ne;
        out->k_eff = K;
        out->kmer_offsets = f_off;
        if (K >= 0) {
            out->kmers_lo = (uint64_t *)alloc((size_t)nk * 8);
            out->kmers_hi = K > 32 ? (uint64_t *)alloc((size_t)nk * 8) : nullptr;
            out->exts = (uint8_t *)alloc((size_t)nk);
            out->counts = (uint16_t *)alloc((size_t)nk * 2);
            if (has_csr) {
                out->csr_row_ptr = (int64_t *)alloc((size_t)(nk + 1) * 8);
                out->csr_col = (int32_t *)alloc((size_t)ne * 4);
                out->csr_row_ptr[nk] = ne;
            }
        }
        if (has_stats) out->stats = (uint32_t *)alloc((size_t)n_groups * 12 + 16);
        host_parallel(n_groups, nt, 1024, [&](int64_t g0, int64_t g1) {
            for (int64_t g = g0; g < g1; g++) {
                const rogtk_graph_batch &p = part[(size_t)shard[(size_t)g]];
                const int64_t i = local_of[(size_t)g];
                if (has_stats && p.stats) memcpy(out->stats + 3 * g, p.stats + 3 * i, 12);
                else if (has_stats) memset(out->stats + 3 * g, 0, 12);
                if (K < 0) continue;
                const int64_t s = p.kmer_offsets[i], n = p.kmer_offsets[i + 1] - s, t = f_off[g];
                memcpy(out->kmers_lo + t, p.kmers_lo + s, (size_t)n * 8);
                if (K > 32) memcpy(out->kmers_hi + t, p.kmers_hi + s, (size_t)n * 8);
                memcpy(out->exts + t, p.exts + s, (size_t)n);
                memcpy(out->counts + t, p.counts + s, (size_t)n * 2);
                if (has_csr) {
                    const int64_t es = p.csr_row_ptr[s], shift = e_off[(size_t)g] - es;
                    for (int64_t j = 0; j < n; j++) out->csr_row_ptr[t + j] = p.csr_row_ptr[s + j] + shift;
                    memcpy(out->csr_col + e_off[(size_t)g], p.csr_col + es, (size_t)(p.csr_row_ptr[s + n] - es) * 4);
                }
            }
        });
        for (int d = 0; d < n_ctx; d++) {
            const rogtk_graph_batch &p = part[(size_t)d];
            out->n_observations += p.n_observations;
            out->n_spilled_groups += p.n_spilled_groups;
            out->n_multipass_groups += p.n_multipass_groups;
            out->ms_total = std::max(out->ms_total, p.ms_total);
            out->ms_count = std::max(out->ms_count, p.ms_count);
            out->ms_hot_kernel = std::max(out->ms_hot_kernel, p.ms_hot_kernel);
            out->ms_sort = std::max(out->ms_sort, p.ms_sort);
            out->ms_csr = std::max(out->ms_csr, p.ms_csr);
            out->ms_stats = std::max(out->ms_stats, p.ms_stats);
        }
    } catch (...) {
        rc = fail_code(c0);
    }
    for (int d = 0; d < n_ctx; d++) {
        if (part_live[(size_t)d]) rogtk_gpu_free_batch(ctxs[d], &part[(size_t)d]);
        SubColumn &sc = sub[(size_t)d];
        for (void *q : {(void *)sc.data, (void *)sc.offs, (void *)sc.goff, (void *)sc.valid})
            if (q) pinned_put(ctxs[d], q);
    }
    if (rc != ROGTK_OK) {
        for (void *q : own->ptrs) pinned_put(c0, q);
        delete own;
        memset(out, 0, sizeof(*out));
    }
    return rc;
}
