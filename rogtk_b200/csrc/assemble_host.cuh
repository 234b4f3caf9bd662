// assemble_host.cuh -- whole-expression entries of the C ABI (include/rogtk_gpu.h): what the reference's plugin functions do per
// group, for a whole frame per call.  Included at the end of rogtk_gpu.cu (it uses the context internals).
//
//   rogtk_gpu_assemble_batch   assemble_sequences_expr (reference src/expressions.rs:695-762), assemble_sequences_with_anchors_expr
//                              (:770-849), fracture_sequences (src/fracture.rs:506-564) on top of assemble_sequences (:188-280)
//   rogtk_gpu_sweep_batch      sweep_assembly_params_expr (src/expressions.rs:880-955)
//   rogtk_gpu_optimize_batch   optimize_assembly / optimize_assembly_expr (src/fracture_opt.rs:120-367)
//
// Counting stays on the device and happens once per distinct K_eff (the reference recounts all reads for every probe); every
// further min_coverage is a rogtk_gpu_rethreshold of that batch; Compression runs on the device (rogtk_gpu_compress_unitigs) and
// only the chosen contigs are copied back; the path-finding methods run on the host over the node arrays + successor CSR
// (pathfind_host.h), one group per task on a pool of threads.
#pragma once
#include <atomic>
#include <set>
#include <thread>

#include "pathfind_host.h"

namespace {

struct HostStrings {   // one string per group, LargeUtf8 layout, plain host memory
    std::vector<int64_t> off;
    ByteVec data;                 // (pathfind_device.cuh: resize() without a zero-fill)
    std::vector<uint8_t> flags;
    std::string get(int64_t g) const { return std::string(data.begin() + off[(size_t)g], data.begin() + off[(size_t)g + 1]); }
    int64_t len(int64_t g) const { return off[(size_t)g + 1] - off[(size_t)g]; }
};

struct RefError {   // a condition the reference reports as PolarsError::ComputeError / PyValueError
    std::string msg;
};

// Compression on a device-resident node batch -> one string per group (into `hs`, or lengths only)
void compression_rows(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *db, int64_t min_length, bool only_largest, HostStrings *hs,
                      std::vector<int64_t> *lengths_only, float *ms_device)
{
    const int64_t ng = db->n_groups;
    rogtk_contig_batch cb;
    const int rc = rogtk_gpu_compress_unitigs(ctx, db, min_length, ROGTK_OUTPUT_ON_DEVICE, &cb);
    if (rc != ROGTK_OK) throw ArgFail{rc, ctx->last_error};
    struct Free {
        rogtk_gpu_ctx *c;
        rogtk_contig_batch *b;
        ~Free() { rogtk_gpu_free_contigs(c, b); }
    } fr{ctx, &cb};
    if (ms_device) *ms_device += cb.ms_total;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t ngn = (size_t)std::max<int64_t>(ng, 1);
    DevBuf rlen(ngn * 8, st), roff((ngn + 1) * 8, st), many(ngn, st), scr(scan_scratch_bytes(ng), st);
    CK(cudaEventRecord(ctx->ev[0], st));
    if (ng > 0)
        k_contig_row_len<<<blocks_for(ng, 256), 256, 0, st>>>(ng, cb.unitig_offsets, cb.seq_offsets, cb.largest, min_length, only_largest ? 1 : 0,
                                                              rlen.as<int64_t>(), many.as<uint8_t>());
    ctx->launches++;
    if (lengths_only) {
        lengths_only->assign((size_t)ng, 0);
        CK(cudaMemcpyAsync(lengths_only->data(), rlen.p, (size_t)ng * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev[3], st));
        CK(cudaStreamSynchronize(st));
    } else {
        exclusive_scan<int64_t>(rlen.as<int64_t>(), ng, roff.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
        int64_t nbytes = 0;
        Peek pk(ctx->peek_slot, st, &ctx->launches);
        pk.add(roff.as<int64_t>() + ng, 8, &nbytes);
        pk.sync();
        DevBuf rows((size_t)std::max<int64_t>(nbytes, 1), st);
        if (ng > 0)
            k_contig_row_fill<<<blocks_for(ng * 32, 256), 256, 0, st>>>(ng, cb.unitig_offsets, cb.seq_offsets, cb.seq, cb.largest, min_length,
                                                                       only_largest ? 1 : 0, roff.as<int64_t>(), rows.as<uint8_t>());
        ctx->launches++;
        CK(cudaEventRecord(ctx->ev[3], st));
        hs->off.assign((size_t)ng + 1, 0);
        hs->data.assign((size_t)nbytes, 0);
        hs->flags.assign((size_t)ng, 0);
        std::vector<uint8_t> many_h((size_t)ng, 0);
        CK(cudaMemcpyAsync(hs->off.data(), roff.p, (size_t)(ng + 1) * 8, cudaMemcpyDeviceToHost, st));
        if (nbytes) CK(cudaMemcpyAsync(hs->data.data(), rows.p, (size_t)nbytes, cudaMemcpyDeviceToHost, st));
        if (ng) CK(cudaMemcpyAsync(hs->flags.data(), cb.largest_flags, (size_t)ng, cudaMemcpyDeviceToHost, st));
        if (ng) CK(cudaMemcpyAsync(many_h.data(), many.p, (size_t)ng, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int64_t g = 0; g < ng; g++) hs->flags[(size_t)g] = only_largest ? hs->flags[(size_t)g] : many_h[(size_t)g];
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
    if (ms_device) *ms_device += ms;
}

struct Anchors {   // scalar anchors or one pair per group
    const char *s = nullptr, *e = nullptr;
    const int64_t *so = nullptr, *eo = nullptr;
    const uint8_t *sd = nullptr, *ed = nullptr;
    bool per_group() const { return so != nullptr; }
    std::string start(int64_t g) const { return so ? std::string((const char *)sd + so[g], (size_t)(so[g + 1] - so[g])) : std::string(s ? s : ""); }
    std::string end(int64_t g) const { return eo ? std::string((const char *)ed + eo[g], (size_t)(eo[g + 1] - eo[g])) : std::string(e ? e : ""); }
};

template <typename F>
void parallel_groups(int64_t ng, int threads, F &&fn)
{
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min<int>(nt, (int)std::min<int64_t>(std::max<int64_t>(1, ng / 16), 256)));
    if (nt == 1) {
        for (int64_t g = 0; g < ng; g++) fn(g);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    std::atomic<bool> failed{false};
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&]() {
            try {
                while (true) {
                    const int64_t g0 = next.fetch_add(64);
                    if (g0 >= ng) break;
                    for (int64_t g = g0; g < std::min<int64_t>(ng, g0 + 64); g++) fn(g);
                }
            } catch (...) {
                failed = true;
            }
        });
    for (auto &t : pool) t.join();
    if (failed) throw std::bad_alloc();
}

// path-finding methods over a HOST-resident node batch with CSR -> one string per group (min_length applied as assemble_sequences does)
void pathfind_rows(const rogtk_graph_batch *hb, int method, const Anchors &an, int64_t min_length, const int32_t *node_rank, int threads,
                   HostStrings &hs)
{
    const int64_t ng = hb->n_groups;
    std::vector<std::string> rows((size_t)ng);
    parallel_groups(ng, threads, [&](int64_t g) {
        const int64_t o = hb->kmer_offsets[g], n = hb->kmer_offsets[g + 1] - o;
        if (n <= 0 || hb->k_eff < 1) return;
        rogtk_host::GroupGraph gg;
        gg.K = hb->k_eff;
        gg.n = (int32_t)n;
        gg.lo = hb->kmers_lo + o;
        gg.hi = hb->kmers_hi ? hb->kmers_hi + o : nullptr;
        gg.exts = hb->exts + o;
        gg.counts = hb->counts + o;
        gg.row_ptr = hb->csr_row_ptr + o;
        gg.col = hb->csr_col;
        gg.rank = node_rank ? node_rank + o : nullptr;
        std::string c = method == ROGTK_METHOD_SHORTEST_PATH ? rogtk_host::shortest_path_contig(gg, an.start(g), an.end(g))
                                                            : rogtk_host::shortest_path_auto_contig(gg);
        if ((int64_t)c.size() >= min_length) rows[(size_t)g] = std::move(c);   // src/fracture.rs:258-262
    });
    hs.off.assign((size_t)ng + 1, 0);
    for (int64_t g = 0; g < ng; g++) hs.off[(size_t)g + 1] = hs.off[(size_t)g] + (int64_t)rows[(size_t)g].size();
    hs.data.resize((size_t)hs.off[(size_t)ng]);
    hs.flags.assign((size_t)ng, 0);
    for (int64_t g = 0; g < ng; g++)
        if (!rows[(size_t)g].empty()) memcpy(hs.data.data() + hs.off[(size_t)g], rows[(size_t)g].data(), rows[(size_t)g].size());
}

// the reference's method / anchor validation (src/expressions.rs:700-731, :811-829; src/djfind.rs:31-58)
void check_method(int method, bool has_start, bool has_end, bool per_group_anchors)
{
    if (per_group_anchors) {
        if (method == ROGTK_METHOD_COMPRESSION) throw RefError{"compression method is not supported with dynamic anchors; use shortest_path"};
        if (method == ROGTK_METHOD_SHORTEST_PATH_AUTO)
            throw RefError{"shortest_path_auto method is not supported with dynamic anchors; use shortest_path"};
        if (method != ROGTK_METHOD_SHORTEST_PATH) throw RefError{"Invalid assembly method for dynamic anchors. Must be 'shortest_path'"};
        return;
    }
    switch (method) {
    case ROGTK_METHOD_COMPRESSION:
        if (has_start || has_end) throw RefError{"Anchor sequences should not be provided for compression method"};
        break;
    case ROGTK_METHOD_SHORTEST_PATH:
        if (!(has_start && has_end)) throw RefError{"Both start_anchor and end_anchor are required for shortest_path method"};
        break;
    case ROGTK_METHOD_SHORTEST_PATH_AUTO:
        if (has_start || has_end) throw RefError{"Anchor sequences should not be provided for shortest_path_auto method"};
        break;
    default:
        throw RefError{"Invalid assembly method. Must be 'compression', 'shortest_path', or 'shortest_path_auto'"};
    }
}

// A counted batch kept on the device (min_coverage of the count in `base_cov`) + the cells derived from it
struct CountedK {
    rogtk_graph_batch db{};
    uint32_t base_cov = 1;
    bool live = false;
};

// one (K_eff, min_coverage) cell for every group: strings (or only their lengths) from a device batch counted at <= cov
void evaluate_cell(rogtk_gpu_ctx *ctx, CountedK &ck, uint32_t cov, int method, const Anchors &an, int64_t min_length, bool only_largest,
                   const int32_t *node_rank, int threads, HostStrings *hs, std::vector<int64_t> *lengths_only, float *ms_device, float *ms_host)
{
    const int64_t ng = ck.db.n_groups;
    if (!ck.live || ck.db.k_eff < 0) {   // k > 64: no contigs (src/fracture.rs:212-215)
        if (lengths_only) lengths_only->assign((size_t)ng, 0);
        if (hs) {
            hs->off.assign((size_t)ng + 1, 0);
            hs->data.clear();
            hs->flags.assign((size_t)ng, 0);
        }
        return;
    }
    const uint32_t eff = cov > 1u ? cov : 1u;
    if (method == ROGTK_METHOD_COMPRESSION) {
        if (eff > (ck.base_cov > 1u ? ck.base_cov : 1u)) {
            rogtk_graph_batch t;
            const int rc = rogtk_gpu_rethreshold(ctx, &ck.db, eff, ROGTK_OUTPUT_ON_DEVICE, &t);
            if (rc != ROGTK_OK) throw ArgFail{rc, ctx->last_error};
            if (ms_device) *ms_device += t.ms_total;
            try {
                compression_rows(ctx, &t, min_length, only_largest, hs, lengths_only, ms_device);
            } catch (...) {
                rogtk_gpu_free_batch(ctx, &t);
                throw;
            }
            rogtk_gpu_free_batch(ctx, &t);
        } else {
            compression_rows(ctx, &ck.db, min_length, only_largest, hs, lengths_only, ms_device);
        }
        return;
    }
    if (!node_rank && !ctx->host_pathfind && ng >= ctx->pathfind_min_groups) {
        // path finding on the device: one thread per group runs the procedure of pathfind_core.h on the batch's nodes + CSR
        rogtk_graph_batch t;
        const int rc = rogtk_gpu_rethreshold(ctx, &ck.db, eff, ROGTK_EMIT_CSR | ROGTK_OUTPUT_ON_DEVICE, &t);
        if (rc != ROGTK_OK) throw ArgFail{rc, ctx->last_error};
        if (ms_device) *ms_device += t.ms_total;
        HostStrings local;
        HostStrings &dst = hs ? *hs : local;
        try {
            DevAnchors da;
            da.s = an.s; da.e = an.e; da.so = an.so; da.eo = an.eo; da.sd = an.sd; da.ed = an.ed;
            pathfind_rows_device(ctx, &t, method, da, min_length, dst.off, dst.data, ms_device);
            dst.flags.assign((size_t)ng, 0);
        } catch (...) {
            rogtk_gpu_free_batch(ctx, &t);
            throw;
        }
        rogtk_gpu_free_batch(ctx, &t);
        if (lengths_only) {
            lengths_only->assign((size_t)ng, 0);
            for (int64_t g = 0; g < ng; g++) (*lengths_only)[(size_t)g] = dst.len(g);
        }
        return;
    }
    // path finding on the host (a boomphf node rank was given, or ROGTK_HOST_PATHFIND): node arrays + CSR of this threshold come back
    rogtk_graph_batch hb;
    const int rc = rogtk_gpu_rethreshold(ctx, &ck.db, eff, ROGTK_EMIT_CSR, &hb);
    if (rc != ROGTK_OK) throw ArgFail{rc, ctx->last_error};
    if (ms_device) *ms_device += hb.ms_total;
    const auto t0 = std::chrono::steady_clock::now();
    HostStrings local;
    HostStrings &dst = hs ? *hs : local;
    try {
        pathfind_rows(&hb, method, an, min_length, node_rank, threads, dst);
    } catch (...) {
        rogtk_gpu_free_batch(ctx, &hb);
        throw;
    }
    rogtk_gpu_free_batch(ctx, &hb);
    if (ms_host) *ms_host += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (lengths_only) {
        lengths_only->assign((size_t)ng, 0);
        for (int64_t g = 0; g < ng; g++) (*lengths_only)[(size_t)g] = dst.len(g);
    }
}

void count_k(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups, int64_t k, uint32_t cov,
             uint32_t flags, CountedK &ck, float *ms_device)
{
    const uint32_t f = (flags & (ROGTK_INPUT_ON_DEVICE | ROGTK_EXACT_K | ROGTK_FORCE_SPILL | ROGTK_NO_MULTIPASS)) | ROGTK_OUTPUT_ON_DEVICE;
    const int rc = rogtk_gpu_count_prune_batch(ctx, reads, group_offsets, n_groups, k, cov, f, &ck.db);
    if (rc != ROGTK_OK) throw ArgFail{rc, ctx->last_error};
    ck.base_cov = cov;
    ck.live = true;
    if (ms_device) *ms_device += ck.db.ms_total;
}

// rows (non-null strings) of every group: sequences.len() after the plugin's .flatten() (src/expressions.rs:737-740)
void group_row_counts(const rogtk_strcol *c, const int64_t *goff, int64_t ng, std::vector<uint32_t> &out)
{
    out.assign((size_t)ng, 0);
    for (int64_t g = 0; g < ng; g++) {
        uint32_t n = 0;
        for (int64_t r = goff[g]; r < goff[g + 1]; r++) {
            if (c->validity) {
                const int64_t b = r + c->validity_bit_offset;
                if (!((c->validity[b >> 3] >> (b & 7)) & 1)) continue;
            }
            n++;
        }
        out[(size_t)g] = n;
    }
}
int64_t row_len_host(const rogtk_strcol *c, int64_t r)
{
    if (c->layout == ROGTK_UTF8_I32) return (int64_t)((const int32_t *)c->offsets_or_views)[r + 1] - ((const int32_t *)c->offsets_or_views)[r];
    if (c->layout == ROGTK_UTF8_I64) return ((const int64_t *)c->offsets_or_views)[r + 1] - ((const int64_t *)c->offsets_or_views)[r];
    return (int64_t)((const uint32_t *)c->offsets_or_views)[4 * r];
}
// estimate_k of one group (src/fracture.rs:24-54) over its non-null rows
int64_t estimate_k_group(const rogtk_strcol *c, int64_t r0, int64_t r1)
{
    uint64_t total = 0, count = 0, rows = 0;
    for (int64_t r = r0; r < r1; r++) {
        if (c->validity) {
            const int64_t b = r + c->validity_bit_offset;
            if (!((c->validity[b >> 3] >> (b & 7)) & 1)) continue;
        }
        rows++;
        const int64_t len = row_len_host(c, r);
        if (len > 0) {
            total += (uint64_t)len;
            count++;
        }
    }
    (void)rows;
    if (count == 0) return 31;
    const double mean = (double)total / (double)count;
    uint64_t k = (uint64_t)std::round(mean / 3.0);
    if (k % 2 == 0) k = k - 1;   // usize wrap for k == 0 as in a release build -> clamps to 63
    if (k < 11) return 11;
    if (k > 63) return 63;
    return (int64_t)k;
}

void hand_over_strings(rogtk_gpu_ctx *ctx, const HostStrings &hs, rogtk_string_batch *out, BatchOwner *own)
{
    const int64_t ng = (int64_t)hs.off.size() - 1;
    out->n_rows = ng;
    out->n_bytes = hs.off.empty() ? 0 : hs.off.back();
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    int64_t *o = (int64_t *)pinned_get(ctx, (size_t)(ng + 1) * 8);
    own->ptrs.push_back(o);
    uint8_t *d = (uint8_t *)pinned_get(ctx, (size_t)std::max<int64_t>(out->n_bytes, 1));
    own->ptrs.push_back(d);
    uint8_t *f = (uint8_t *)pinned_get(ctx, (size_t)std::max<int64_t>(ng, 1));
    own->ptrs.push_back(f);
    memcpy(o, hs.off.data(), (size_t)(ng + 1) * 8);
    if (out->n_bytes)   // (hundreds of MB for a large frame: all staging threads copy)
        host_parallel(out->n_bytes, ctx->stage_threads, 8 << 20, [&](int64_t x0, int64_t x1) { memcpy(d + x0, hs.data.data() + x0, (size_t)(x1 - x0)); });
    if (ng) memcpy(f, hs.flags.data(), (size_t)ng);
    out->offsets = o;
    out->data = d;
    out->flags = f;
}

void release_owner(rogtk_gpu_ctx *ctx, void *owner)
{
    if (!owner) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    BatchOwner *own = reinterpret_cast<BatchOwner *>(owner);
    for (void *p : own->ptrs) pinned_put(ctx, p);
    delete own;
}

int ref_fail(rogtk_gpu_ctx *ctx)   // call inside catch (...): like fail_code, plus the reference's own error conditions
{
    try {
        throw;
    } catch (const RefError &e) {
        set_error(ctx, "%s", e.msg.c_str());
        return ROGTK_ERR_INVALID_ARG;
    } catch (...) {
        return fail_code(ctx);
    }
}

}  // namespace

extern "C" void rogtk_gpu_free_strings(rogtk_gpu_ctx *ctx, rogtk_string_batch *b)
{
    if (!ctx || !b) return;
    release_owner(ctx, b->_owner);
    memset(b, 0, sizeof(*b));
}

extern "C" int rogtk_gpu_assemble_batch(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups,
                                        const rogtk_assembly_params *p, uint32_t flags, rogtk_string_batch *out)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    if (!reads || !p || !out || n_groups < 0 || (n_groups > 0 && !group_offsets)) {
        set_error(ctx, "null or negative argument");
        return ROGTK_ERR_INVALID_ARG;
    }
    memset(out, 0, sizeof(*out));
    BatchOwner *own = new BatchOwner();
    out->_owner = own;
    int rc = ROGTK_OK;
    std::vector<CountedK> counted;
    try {
        Anchors an;
        an.s = p->start_anchor; an.e = p->end_anchor;
        an.so = p->start_anchor_offsets; an.sd = p->start_anchor_data;
        an.eo = p->end_anchor_offsets; an.ed = p->end_anchor_data;
        if ((an.so != nullptr) != (an.eo != nullptr)) throw RefError{"assemble_sequences_with_anchors requires 3 inputs: sequences, start_anchor, end_anchor"};
        check_method(p->method, p->start_anchor != nullptr, p->end_anchor != nullptr, an.per_group());
        const bool in_dev = flags & ROGTK_INPUT_ON_DEVICE;
        if (p->auto_k && in_dev) throw ArgFail{ROGTK_ERR_UNSUPPORTED, "auto_k needs host-resident offsets"};
        const int64_t min_length = p->min_length > 0 ? p->min_length : 0;
        // k per group (auto_k) -> classes of equal K_eff; one counting pass per class
        std::vector<int64_t> kg;
        std::vector<int64_t> classes;
        if (p->auto_k) {
            kg.resize((size_t)n_groups);
            std::set<int64_t> ks;
            for (int64_t g = 0; g < n_groups; g++) {
                kg[(size_t)g] = estimate_k_group(reads, group_offsets[g], group_offsets[g + 1]);
                ks.insert((int64_t)k_eff_host(kg[(size_t)g]));
            }
            classes.assign(ks.begin(), ks.end());
        } else {
            classes.push_back(p->k);
        }
        if (p->method == ROGTK_METHOD_COMPRESSION && !p->auto_k) {
            // large host-resident input: the overlapped path (H2D of the next chunk and D2H of the previous one under the compute),
            // unitigs on the device per chunk, only the contigs come back -- straight into the page-locked blocks handed to the caller
            ChunkPost post;
            post.min_length = min_length;
            post.only_largest = p->only_largest != 0;
            rogtk_graph_batch info;
            bool ran = false;
            {
                std::lock_guard<std::mutex> lk(ctx->mu);
                CacheScope cache_scope(ctx->cache);
                ran = assemble_compression_chunked(ctx, reads, group_offsets, n_groups, p->k, p->min_coverage, flags, post, &info);
                if (ran) {
                    for (HostArr *h : {&post.h_soff, &post.h_sdata, &post.h_sflags})
                        if (h->p) own->ptrs.push_back(h->p);
                }
            }
            if (ran) {
                out->n_rows = n_groups;
                out->n_bytes = post.n_bytes;
                out->offsets = (int64_t *)post.h_soff.p;
                out->data = post.h_sdata.p;
                out->flags = post.h_sflags.p;
                out->ms_device = info.ms_total + post.ms_unitig;
                out->n_count_passes = 1;
                if (!out->data) {   // no contig at all: still a valid (empty) data buffer
                    std::lock_guard<std::mutex> lk(ctx->mu);
                    CacheScope cache_scope(ctx->cache);
                    out->data = (uint8_t *)pinned_get(ctx, 16);
                    own->ptrs.push_back(out->data);
                }
                return ROGTK_OK;
            }
        }
        HostStrings result;
        result.off.assign((size_t)n_groups + 1, 0);
        result.flags.assign((size_t)n_groups, 0);
        std::vector<HostStrings> per_class(classes.size());
        for (size_t ci = 0; ci < classes.size(); ci++) {
            counted.emplace_back();
            CountedK &ck = counted.back();
            out->n_count_passes++;
            if (p->method != ROGTK_METHOD_COMPRESSION && !p->node_rank && !ctx->host_pathfind && n_groups >= ctx->pathfind_min_groups) {
                // path finding on the device: nodes + successor CSR stay in HBM, one thread per group searches (pathfind_device.cuh)
                rogtk_graph_batch t;
                const uint32_t f = (flags & (ROGTK_INPUT_ON_DEVICE | ROGTK_EXACT_K | ROGTK_FORCE_SPILL | ROGTK_NO_MULTIPASS)) | ROGTK_EMIT_CSR |
                                   ROGTK_OUTPUT_ON_DEVICE;
                const int rc2 = rogtk_gpu_count_prune_batch(ctx, reads, group_offsets, n_groups, classes[ci], p->min_coverage, f, &t);
                if (rc2 != ROGTK_OK) throw ArgFail{rc2, ctx->last_error};
                out->ms_device += t.ms_total;
                try {
                    per_class[ci].off.assign((size_t)n_groups + 1, 0);
                    per_class[ci].flags.assign((size_t)n_groups, 0);
                    if (t.k_eff >= 0) {
                        DevAnchors da;
                        da.s = an.s; da.e = an.e; da.so = an.so; da.eo = an.eo; da.sd = an.sd; da.ed = an.ed;
                        pathfind_rows_device(ctx, &t, p->method, da, min_length, per_class[ci].off, per_class[ci].data, &out->ms_device);
                    }
                } catch (...) {
                    rogtk_gpu_free_batch(ctx, &t);
                    throw;
                }
                rogtk_gpu_free_batch(ctx, &t);
                continue;
            }
            if (p->method != ROGTK_METHOD_COMPRESSION) {
                // path finding on the host (a boomphf node rank was given, or ROGTK_HOST_PATHFIND): node arrays + successor CSR straight
                // to host memory (chunked, copies overlapped)
                rogtk_graph_batch hb;
                const uint32_t f = (flags & (ROGTK_INPUT_ON_DEVICE | ROGTK_EXACT_K | ROGTK_FORCE_SPILL | ROGTK_NO_MULTIPASS)) | ROGTK_EMIT_CSR;
                const int rc2 = rogtk_gpu_count_prune_batch(ctx, reads, group_offsets, n_groups, classes[ci], p->min_coverage, f, &hb);
                if (rc2 != ROGTK_OK) throw ArgFail{rc2, ctx->last_error};
                out->ms_device += hb.ms_total;
                const auto t0 = std::chrono::steady_clock::now();
                try {
                    if (hb.k_eff < 0) {
                        per_class[ci].off.assign((size_t)n_groups + 1, 0);
                        per_class[ci].flags.assign((size_t)n_groups, 0);
                    } else {
                        pathfind_rows(&hb, p->method, an, min_length, p->node_rank, p->host_threads, per_class[ci]);
                    }
                } catch (...) {
                    rogtk_gpu_free_batch(ctx, &hb);
                    throw;
                }
                rogtk_gpu_free_batch(ctx, &hb);
                out->ms_host += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
                continue;
            }
            count_k(ctx, reads, group_offsets, n_groups, classes[ci], p->min_coverage, flags, ck, &out->ms_device);
            evaluate_cell(ctx, ck, p->min_coverage, p->method, an, min_length, p->only_largest != 0, p->node_rank, p->host_threads, &per_class[ci],
                          nullptr, &out->ms_device, &out->ms_host);
            rogtk_gpu_free_batch(ctx, &ck.db);
            ck.live = false;
        }
        if (classes.size() == 1) {
            result = std::move(per_class[0]);
            if (p->auto_k)
                for (auto &f : result.flags) f |= 4;
        } else {   // every group takes the row of its own class
            std::vector<int> cls((size_t)n_groups, 0);
            for (int64_t g = 0; g < n_groups; g++) {
                const int64_t K = k_eff_host(kg[(size_t)g]);
                cls[(size_t)g] = (int)(std::lower_bound(classes.begin(), classes.end(), K) - classes.begin());
                result.off[(size_t)g + 1] = result.off[(size_t)g] + per_class[(size_t)cls[(size_t)g]].len(g);
            }
            result.data.resize((size_t)result.off[(size_t)n_groups]);
            for (int64_t g = 0; g < n_groups; g++) {
                const HostStrings &src = per_class[(size_t)cls[(size_t)g]];
                if (src.len(g)) memcpy(result.data.data() + result.off[(size_t)g], src.data.data() + src.off[(size_t)g], (size_t)src.len(g));
                result.flags[(size_t)g] = src.flags[(size_t)g] | 4;
            }
        }
        hand_over_strings(ctx, result, out, own);
    } catch (...) {
        rc = ref_fail(ctx);
    }
    for (auto &ck : counted)
        if (ck.live) rogtk_gpu_free_batch(ctx, &ck.db);
    if (rc != ROGTK_OK) {
        release_owner(ctx, own);
        memset(out, 0, sizeof(*out));
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// sweep_assembly_params_expr
// ---------------------------------------------------------------------------------------------------------------------------------
extern "C" void rogtk_gpu_free_sweep(rogtk_gpu_ctx *ctx, rogtk_sweep_result *r)
{
    if (!ctx || !r) return;
    release_owner(ctx, r->_owner);
    memset(r, 0, sizeof(*r));
}

extern "C" int rogtk_gpu_sweep_batch(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups,
                                     const rogtk_sweep_params *p, uint32_t flags, rogtk_sweep_result *out)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    if (!reads || !p || !out || n_groups < 0 || (n_groups > 0 && !group_offsets)) {
        set_error(ctx, "null or negative argument");
        return ROGTK_ERR_INVALID_ARG;
    }
    memset(out, 0, sizeof(*out));
    BatchOwner *own = new BatchOwner();
    out->_owner = own;
    int rc = ROGTK_OK;
    CountedK ck;
    try {
        if (p->k_start < 0 || p->k_end < 0 || p->cov_start < 0 || p->cov_end < 0) throw RefError{"negative sweep bound"};
        if (p->k_step <= 0 || p->cov_step <= 0) throw RefError{"assertion failed: step != 0"};   // Rust's step_by(0) panics
        check_method(p->method, p->start_anchor != nullptr, p->end_anchor != nullptr, false);
        Anchors an;
        an.s = p->start_anchor; an.e = p->end_anchor;
        std::vector<int64_t> ks, covs;
        for (int64_t k = p->k_start; k <= p->k_end; k += p->k_step) ks.push_back(k);
        for (int64_t c = p->cov_start; c <= p->cov_end; c += p->cov_step) covs.push_back(c);
        const int64_t n_cells = (int64_t)(ks.size() * covs.size());
        out->n_groups = n_groups;
        out->n_cells = n_cells;
        const size_t total = (size_t)std::max<int64_t>(1, n_groups * n_cells);
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            CacheScope cache_scope(ctx->cache);
            out->k = (int64_t *)pinned_get(ctx, total * 8); own->ptrs.push_back(out->k);
            out->min_coverage = (int64_t *)pinned_get(ctx, total * 8); own->ptrs.push_back(out->min_coverage);
            out->contig_length = (int64_t *)pinned_get(ctx, total * 8); own->ptrs.push_back(out->contig_length);
        }
        // cells grouped by K_eff: one count (at the smallest min_coverage of the range), one evaluation per (K_eff, min_coverage)
        std::map<int, std::vector<size_t>> by_K;
        for (size_t i = 0; i < ks.size(); i++) by_K[k_eff_host(ks[i])].push_back(i);
        std::vector<int64_t> lens;
        for (auto &kv : by_K) {
            if (covs.empty()) break;
            const int K = kv.first;
            const uint32_t base = (uint32_t)std::min<int64_t>(covs.front(), 0xffffffffll);
            if (K > 0) {
                count_k(ctx, reads, group_offsets, n_groups, K, base, flags, ck, &out->ms_device);
                out->n_count_passes++;
            } else {
                ck.db = rogtk_graph_batch{};
                ck.db.n_groups = n_groups;
                ck.db.k_eff = -1;
                ck.live = false;
            }
            for (size_t ci = 0; ci < covs.size(); ci++) {
                const uint32_t cov = (uint32_t)std::min<int64_t>(covs[ci], 0xffffffffll);
                // assemble_sequences(.., Some(true), None, Some(false), ..): largest contig, no min_length (src/expressions.rs:913-923)
                evaluate_cell(ctx, ck, cov, p->method, an, 0, true, nullptr, p->host_threads, nullptr, &lens, &out->ms_device, &out->ms_host);
                for (const size_t ki : kv.second) {
                    const int64_t cell = (int64_t)(ki * covs.size() + ci);
                    for (int64_t g = 0; g < n_groups; g++) {
                        const int64_t row = g * n_cells + cell;
                        out->k[row] = ks[ki];
                        out->min_coverage[row] = covs[ci];
                        out->contig_length[row] = lens[(size_t)g];
                    }
                }
            }
            if (ck.live) {
                rogtk_gpu_free_batch(ctx, &ck.db);
                ck.live = false;
            }
        }
    } catch (...) {
        rc = ref_fail(ctx);
    }
    if (ck.live) rogtk_gpu_free_batch(ctx, &ck.db);
    if (rc != ROGTK_OK) {
        release_owner(ctx, own);
        memset(out, 0, sizeof(*out));
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// optimize_assembly_expr
// ---------------------------------------------------------------------------------------------------------------------------------
extern "C" void rogtk_gpu_free_optimize(rogtk_gpu_ctx *ctx, rogtk_optimize_result *r)
{
    if (!ctx || !r) return;
    release_owner(ctx, r->contig._owner);
    release_owner(ctx, r->_owner);
    memset(r, 0, sizeof(*r));
}

namespace {

struct ParamPoint {
    int64_t k, cov;
    bool operator<(const ParamPoint &o) const { return k < o.k || (k == o.k && cov < o.cov); }
};
struct OptResult {   // AssemblyResult (src/fracture_opt.rs:35-66)
    std::string contig;
    ParamPoint params{0, 0};
    size_t length = 0;
    bool has_anchors = false;
    bool valid = false;
};
struct ExplorationPath {
    ParamPoint params;
    size_t length;
    size_t steps_without_improvement;
};
struct OptState {   // optimize_assembly's locals for one group
    std::set<ParamPoint> tested;
    OptResult best_anchored, best_length, returned;
    bool done = false, early = false;
    std::vector<ExplorationPath> paths;
    std::vector<std::pair<ParamPoint, size_t>> probes;   // this iteration's new parameter points and the path each came from
};

bool apply_direction(int dir, ParamPoint pt, ParamPoint &o)   // Direction::apply (src/fracture_opt.rs:76-105)
{
    switch (dir) {
    case 0: if (pt.cov > 1) { o = ParamPoint{pt.k, pt.cov - 1}; return true; } return false;          // West
    case 1: o = ParamPoint{pt.k, pt.cov + 1}; return true;                                              // East
    case 2: if (pt.k > 4) { o = ParamPoint{pt.k - 1, pt.cov}; return true; } return false;            // North
    default: if (pt.k < 64) { o = ParamPoint{pt.k + 1, pt.cov}; return true; } return false;          // South
    }
}

}  // namespace

extern "C" int rogtk_gpu_optimize_batch(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups,
                                        const rogtk_optimize_params *p, uint32_t flags, rogtk_optimize_result *out)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    if (!reads || !p || !out || n_groups < 0 || (n_groups > 0 && !group_offsets)) {
        set_error(ctx, "null or negative argument");
        return ROGTK_ERR_INVALID_ARG;
    }
    memset(out, 0, sizeof(*out));
    BatchOwner *own = new BatchOwner(), *own_s = new BatchOwner();
    out->_owner = own;
    out->contig._owner = own_s;
    int rc = ROGTK_OK;
    std::map<int, CountedK> counted;   // K_eff -> batch counted once at min_coverage 1
    try {
        if (flags & ROGTK_INPUT_ON_DEVICE) throw ArgFail{ROGTK_ERR_UNSUPPORTED, "optimize needs host-resident input (input_sequences is counted on the host)"};
        if (!p->start_anchor) throw RefError{"start_anchor is required"};
        if (!p->end_anchor) throw RefError{"end_anchor is required"};
        // AssemblyMethod::from_str(method, Some(start), Some(end)) (src/fracture_opt.rs:304-308): anchors are always passed
        check_method(p->method, true, true, false);
        Anchors an;
        an.s = p->start_anchor; an.e = p->end_anchor;
        const std::string sa(p->start_anchor), ea(p->end_anchor);
        const int64_t max_it = p->max_iterations > 0 ? p->max_iterations : 50;
        const int n_dir = p->explore_k ? 4 : 2;
        std::vector<uint32_t> n_seq;
        group_row_counts(reads, group_offsets, n_groups, n_seq);
        // memo of evaluated cells: (K_eff, max(cov, 1)) -> the largest contig of every group
        std::map<std::pair<int, int64_t>, HostStrings> memo;
        auto need_cell = [&](ParamPoint pt) {
            const int K = pt.k > 64 ? -1 : k_eff_host(pt.k);
            const int64_t cov = pt.cov > 1 ? pt.cov : 1;
            const auto key = std::make_pair(K, K < 0 ? 0 : cov);
            if (memo.count(key)) return;
            HostStrings hs;
            if (K < 0) {
                hs.off.assign((size_t)n_groups + 1, 0);
                hs.flags.assign((size_t)n_groups, 0);
            } else {
                CountedK &ck = counted[K];
                if (!ck.live) {
                    count_k(ctx, reads, group_offsets, n_groups, K, 1, flags, ck, &out->contig.ms_device);
                    out->contig.n_count_passes++;
                }
                // assemble_and_check: only_largest, no min_length (src/fracture_opt.rs:254-264)
                evaluate_cell(ctx, ck, (uint32_t)std::min<int64_t>(cov, 0xffffffffll), p->method, an, 0, true, nullptr, p->host_threads, &hs, nullptr,
                              &out->contig.ms_device, &out->contig.ms_host);
            }
            memo.emplace(key, std::move(hs));
            out->n_cells_evaluated++;
        };
        auto probe = [&](int64_t g, ParamPoint pt) -> OptResult {   // assemble_and_check + AssemblyResult::new
            const int K = pt.k > 64 ? -1 : k_eff_host(pt.k);
            const HostStrings &hs = memo.at(std::make_pair(K, K < 0 ? (int64_t)0 : (pt.cov > 1 ? pt.cov : (int64_t)1)));
            OptResult r;
            r.contig = hs.get(g);
            r.params = pt;
            r.length = r.contig.size();
            r.has_anchors = r.contig.find(sa) != std::string::npos && r.contig.find(ea) != std::string::npos;
            r.valid = true;
            out->n_probes++;
            return r;
        };
        std::vector<OptState> st((size_t)n_groups);
        const ParamPoint start{p->start_k, p->start_min_coverage};
        need_cell(start);
        for (int64_t g = 0; g < n_groups; g++) {
            OptState &s = st[(size_t)g];
            s.tested.insert(start);
            const OptResult cur = probe(g, start);
            if (cur.has_anchors) s.best_anchored = cur;
            if (!s.best_length.valid || cur.length > s.best_length.length) s.best_length = cur;
            s.paths.push_back(ExplorationPath{cur.params, cur.length, 0});
        }
        for (int64_t it = 0; it < max_it; it++) {
            // the parameter points this iteration tests depend only on the paths and on what was tested before, not on results
            bool any = false;
            for (int64_t g = 0; g < n_groups; g++) {
                OptState &s = st[(size_t)g];
                s.probes.clear();
                if (s.done) continue;
                for (size_t pi = 0; pi < s.paths.size(); pi++)
                    for (int d = 0; d < n_dir; d++) {
                        ParamPoint np;
                        if (!apply_direction(d, s.paths[pi].params, np)) continue;
                        if (s.tested.count(np)) continue;
                        s.tested.insert(np);
                        s.probes.emplace_back(np, pi);
                        need_cell(np);
                        any = true;
                    }
            }
            if (!any) break;
            for (int64_t g = 0; g < n_groups; g++) {
                OptState &s = st[(size_t)g];
                if (s.done) continue;
                std::vector<ExplorationPath> new_paths;
                for (const auto &pr : s.probes) {
                    const OptResult r = probe(g, pr.first);
                    if (r.has_anchors && (!s.best_anchored.valid || r.length > s.best_anchored.length)) s.best_anchored = r;
                    if (!s.best_length.valid || r.length > s.best_length.length) s.best_length = r;
                    if (r.has_anchors && !p->prioritize_length) {   // early return (src/fracture_opt.rs:186-189)
                        s.returned = r;
                        s.done = s.early = true;
                        break;
                    }
                    if (!r.contig.empty()) {
                        const ExplorationPath &from = s.paths[pr.second];
                        new_paths.push_back(ExplorationPath{pr.first, r.length, r.length > from.length ? 0 : from.steps_without_improvement + 1});
                    }
                }
                if (s.done) continue;
                if (new_paths.empty()) {
                    s.done = true;
                    continue;
                }
                // select_promising_paths: stable sort by (length desc, steps asc), keep 4
                std::stable_sort(new_paths.begin(), new_paths.end(), [](const ExplorationPath &a, const ExplorationPath &b) {
                    if (a.length != b.length) return a.length > b.length;
                    return a.steps_without_improvement < b.steps_without_improvement;
                });
                if (new_paths.size() > 4) new_paths.resize(4);
                s.paths = std::move(new_paths);
            }
        }
        // rows
        HostStrings rows;
        rows.off.assign((size_t)n_groups + 1, 0);
        rows.flags.assign((size_t)n_groups, 0);
        std::vector<const OptResult *> pick((size_t)n_groups, nullptr);
        for (int64_t g = 0; g < n_groups; g++) {
            OptState &s = st[(size_t)g];
            const OptResult *r = s.early ? &s.returned : (p->prioritize_length ? &s.best_length : &s.best_anchored);
            if (!r->valid) r = nullptr;
            pick[(size_t)g] = r;
            rows.off[(size_t)g + 1] = rows.off[(size_t)g] + (r ? (int64_t)r->contig.size() : 0);
        }
        rows.data.resize((size_t)rows.off[(size_t)n_groups]);
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            CacheScope cache_scope(ctx->cache);
            const size_t nb = (size_t)std::max<int64_t>(n_groups, 1) * 4;
            out->k = (uint32_t *)pinned_get(ctx, nb); own->ptrs.push_back(out->k);
            out->min_coverage = (uint32_t *)pinned_get(ctx, nb); own->ptrs.push_back(out->min_coverage);
            out->length = (uint32_t *)pinned_get(ctx, nb); own->ptrs.push_back(out->length);
            out->input_sequences = (uint32_t *)pinned_get(ctx, nb); own->ptrs.push_back(out->input_sequences);
        }
        for (int64_t g = 0; g < n_groups; g++) {
            const OptResult *r = pick[(size_t)g];
            if (r && !r->contig.empty()) memcpy(rows.data.data() + rows.off[(size_t)g], r->contig.data(), r->contig.size());
            // Ok(Some(result)) row, or the zero-filled row (src/fracture_opt.rs:326-354)
            out->k[g] = r ? (uint32_t)r->params.k : 0u;
            out->min_coverage[g] = r ? (uint32_t)r->params.cov : 0u;
            out->length[g] = r ? (uint32_t)r->length : 0u;
            out->input_sequences[g] = n_seq[(size_t)g];
        }
        const float msd = out->contig.ms_device, msh = out->contig.ms_host;
        const int64_t ncp = out->contig.n_count_passes;
        hand_over_strings(ctx, rows, &out->contig, own_s);
        out->contig.ms_device = msd;
        out->contig.ms_host = msh;
        out->contig.n_count_passes = ncp;
    } catch (...) {
        rc = ref_fail(ctx);
    }
    for (auto &kv : counted)
        if (kv.second.live) rogtk_gpu_free_batch(ctx, &kv.second.db);
    if (rc != ROGTK_OK) {
        release_owner(ctx, own);
        release_owner(ctx, own_s);
        memset(out, 0, sizeof(*out));
    }
    return rc;
}

extern "C" int64_t rogtk_host_shortest_path(int k_eff, int32_t n_nodes, const uint64_t *kmers_lo, const uint64_t *kmers_hi, const uint16_t *counts,
                                            const int64_t *csr_row_ptr, const int32_t *csr_col, const int32_t *node_rank, const char *start_anchor,
                                            const char *end_anchor, uint8_t *out, int64_t cap)
{
    if (k_eff < 1 || k_eff > 64 || n_nodes < 0 || (n_nodes > 0 && (!kmers_lo || !counts || !csr_row_ptr)) || (k_eff > 32 && n_nodes > 0 && !kmers_hi) ||
        ((start_anchor == nullptr) != (end_anchor == nullptr)))
        return 0;
    try {
        rogtk_host::GroupGraph gg;
        gg.K = k_eff;
        gg.n = n_nodes;
        gg.lo = kmers_lo;
        gg.hi = k_eff > 32 ? kmers_hi : nullptr;
        gg.counts = counts;
        gg.row_ptr = csr_row_ptr;
        gg.col = csr_col;
        gg.rank = node_rank;
        const std::string c = start_anchor ? rogtk_host::shortest_path_contig(gg, start_anchor, end_anchor) : rogtk_host::shortest_path_auto_contig(gg);
        if ((int64_t)c.size() > cap) return -(int64_t)c.size();
        if (!c.empty() && out) memcpy(out, c.data(), c.size());
        return (int64_t)c.size();
    } catch (...) {
        return 0;
    }
}
