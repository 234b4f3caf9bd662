// pathfind_device.cuh -- the shortest-path assembly methods on the device: ONE THREAD PER GROUP runs the
// container-free procedure of pathfind_core.h -- the same source the host entries execute, so the result is petgraph's first-pop
// Dijkstra with Rust's heap order, the 1e-9 back-trace and the automatic method's scoring, bit for bit (edge weights come from a table
// filled on the host with libm's log).  The procedure is sequential per group by nature (a heap); the parallelism is across groups:
// a frame has 10^5..10^6 of them.  Working arrays live in HBM (32 B per node + 24 B per edge), groups go in waves bounded by the
// workspace.  Included by rogtk_gpu.cu after the context definition.
#pragma once
#include "pathfind_core.h"
#include "pathfind_host.h"

namespace {

// byte buffer whose resize() does not zero-fill (the bytes are written right after: a 300 MB memset is a tenth of the call)
template <typename T>
struct NoInitAlloc : std::allocator<T> {
    template <typename U>
    struct rebind {
        using other = NoInitAlloc<U>;
    };
    template <typename U>
    void construct(U *p) noexcept
    {
        ::new (static_cast<void *>(p)) U;
    }
    template <typename U, typename... A>
    void construct(U *p, A &&...a)
    {
        ::new (static_cast<void *>(p)) U(std::forward<A>(a)...);
    }
};
using ByteVec = std::vector<uint8_t, NoInitAlloc<uint8_t>>;

struct PathArgs {
    int64_t g0, g1;                 // groups of this wave
    const int64_t *kmer_offsets;
    const uint64_t *lo, *hi;
    const uint16_t *counts;
    const int64_t *row_ptr;
    const int32_t *col;
    int K, method;
    const char *sa, *ea;            // scalar anchors (device) ...
    int sl, el;
    const int64_t *so, *eo;         // ... or one pair per group (LargeUtf8 layout, device)
    const uint8_t *sd, *ed;
    const double *wtab;
    int64_t o0, e0;                 // first node / first edge of the wave
    int32_t *list_a, *list_b, *path, *best_path;
    rogtk_pf::NodeRec *node;
    rogtk_pf::EdgeRec *edge;
    rogtk_pf::HeapItem *heap;
    char *out;                      // slot of group g: kmer_offsets[g] + g * K  (n + K - 1 bytes at most)
    int32_t *out_len;               // [n_groups]
    int64_t min_length;
    int64_t max_nodes;              // larger groups are left to the host (one thread would take seconds for 10^4..10^5 nodes)
};
constexpr int64_t PF_DEVICE_MAX_NODES = 4096;

__global__ void __launch_bounds__(128) k_pathfind(PathArgs a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = a.g0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < a.g1; g += stride) {
        const int64_t o = a.kmer_offsets[g];
        const int64_t n = a.kmer_offsets[g + 1] - o;
        if (n <= 0 || n > a.max_nodes) {
            a.out_len[g] = 0;
            continue;
        }
        const int64_t ln = o - a.o0, le = a.row_ptr[o] - a.e0, gi = g - a.g0;
        const rogtk_pf::View v{a.K, (int32_t)n, a.lo + o, a.hi ? a.hi + o : nullptr, a.counts + o, a.row_ptr + o, a.col, a.wtab};
        const rogtk_pf::Scratch s{a.node + ln, a.edge + le, a.heap + le + gi, a.path + ln + gi, a.best_path + ln + gi, a.list_a + ln, a.list_b + ln};
        char *out = a.out + o + g * a.K;
        int32_t len;
        if (a.method == ROGTK_METHOD_SHORTEST_PATH) {
            const char *sa = a.so ? reinterpret_cast<const char *>(a.sd) + a.so[g] : a.sa;
            const char *ea = a.eo ? reinterpret_cast<const char *>(a.ed) + a.eo[g] : a.ea;
            const int sl = a.so ? (int)(a.so[g + 1] - a.so[g]) : a.sl;
            const int el = a.eo ? (int)(a.eo[g + 1] - a.eo[g]) : a.el;
            len = rogtk_pf::anchored_contig(v, s, sa, sl, ea, el, out);
        } else {
            len = rogtk_pf::auto_contig(v, s, out);
        }
        a.out_len[g] = (int64_t)len >= a.min_length ? len : 0;   // src/fracture.rs:258-262
    }
}

// edges before every group: dst[g] = row_ptr[kmer_offsets[g]], g in [0, n_groups]
__global__ void k_group_edge_base(const int64_t *kmer_offsets, const int64_t *row_ptr, int64_t n_groups, int64_t *dst)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= n_groups) dst[g] = row_ptr[kmer_offsets[g]];
}

struct DevAnchors {   // what pathfind_rows_device needs from the caller's anchors (host pointers)
    const char *s = nullptr, *e = nullptr;
    const int64_t *so = nullptr, *eo = nullptr;
    const uint8_t *sd = nullptr, *ed = nullptr;
};

// db: device-resident batch WITH its successor CSR.  One string per group into off / data (host vectors), "" where there is no contig.
void pathfind_rows_device(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *db, int method, const DevAnchors &an, int64_t min_length,
                          std::vector<int64_t> &off, ByteVec &data, float *ms_device)
{
    const int64_t ng = db->n_groups, N = db->n_kmers;
    const int K = db->k_eff;
    off.assign((size_t)ng + 1, 0);
    data.clear();
    if (ng == 0 || N == 0 || K < 1) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (!ctx->d_wtab) {   // -ln(sum / 2) for every sum of two u16 counts, from the host's libm (pathfind_host.h)
        CK(cudaMalloc(&ctx->d_wtab, (2 * 65535 + 1) * sizeof(double)));
        CK(cudaMemcpy(ctx->d_wtab, rogtk_host::edge_weight_table(), (2 * 65535 + 1) * sizeof(double), cudaMemcpyHostToDevice));
    }
    // group sizes on the host: nodes from kmer_offsets, edges from row_ptr at the group starts
    std::vector<int64_t> h_off((size_t)ng + 1), h_eb((size_t)ng + 1);
    {
        DevBuf eb((size_t)(ng + 1) * 8, st);
        k_group_edge_base<<<blocks_for(ng + 1, 256), 256, 0, st>>>(db->kmer_offsets, db->csr_row_ptr, ng, eb.as<int64_t>());
        ctx->launches++;
        CK(cudaMemcpyAsync(h_off.data(), db->kmer_offsets, (size_t)(ng + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_eb.data(), eb.p, (size_t)(ng + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    // anchors
    DevBuf d_sa, d_ea, d_so, d_eo, d_sd, d_ed;
    PathArgs a{};
    auto up = [&](DevBuf &b, const void *src, size_t bytes) {
        b.alloc(std::max<size_t>(bytes, 16), st);
        if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
    };
    if (method == ROGTK_METHOD_SHORTEST_PATH) {
        if (an.so) {
            up(d_so, an.so, (size_t)(ng + 1) * 8);
            up(d_eo, an.eo, (size_t)(ng + 1) * 8);
            up(d_sd, an.sd, (size_t)an.so[ng]);
            up(d_ed, an.ed, (size_t)an.eo[ng]);
            a.so = d_so.as<int64_t>(); a.eo = d_eo.as<int64_t>(); a.sd = d_sd.as<uint8_t>(); a.ed = d_ed.as<uint8_t>();
        } else {
            a.sl = an.s ? (int)strlen(an.s) : 0;
            a.el = an.e ? (int)strlen(an.e) : 0;
            up(d_sa, an.s, (size_t)a.sl);
            up(d_ea, an.e, (size_t)a.el);
            a.sa = d_sa.as<char>(); a.ea = d_ea.as<char>();
        }
    }
    a.kmer_offsets = db->kmer_offsets; a.lo = db->kmers_lo; a.hi = K > 32 ? db->kmers_hi : nullptr; a.counts = db->counts;
    a.row_ptr = db->csr_row_ptr; a.col = db->csr_col; a.K = K; a.method = method; a.wtab = ctx->d_wtab; a.min_length = min_length;
    a.max_nodes = PF_DEVICE_MAX_NODES;
    const size_t out_bytes = (size_t)N + (size_t)ng * (size_t)K + 16;
    DevBuf d_out(out_bytes, st), d_len((size_t)ng * 4, st);
    a.out = d_out.as<char>(); a.out_len = d_len.as<int32_t>();
    CK(cudaEventRecord(ctx->ev[12], st));
    // waves: 32 B per node + 24 B per edge + per-group slack within half the workspace
    const int64_t budget = std::max<int64_t>(64 << 20, (int64_t)(ctx->workspace_bytes / 2));
    int64_t g0 = 0;
    while (g0 < ng) {
        int64_t g1 = g0, bytes = 0;
        while (g1 < ng) {
            const int64_t n = h_off[(size_t)g1 + 1] - h_off[(size_t)g1], m = h_eb[(size_t)g1 + 1] - h_eb[(size_t)g1];
            const int64_t need = 32 * n + 24 * m + 64;   // (groups the kernel skips keep their place: the arrays are indexed by node / edge offset)
            if (g1 > g0 && bytes + need > budget) break;
            bytes += need;
            g1++;
        }
        const int64_t nw = h_off[(size_t)g1] - h_off[(size_t)g0], mw = h_eb[(size_t)g1] - h_eb[(size_t)g0], gw = g1 - g0;
        const size_t nn = (size_t)std::max<int64_t>(nw, 1), mm = (size_t)std::max<int64_t>(mw, 1), gg = (size_t)gw + 2;
        DevBuf node(nn * sizeof(rogtk_pf::NodeRec), st), edge(mm * sizeof(rogtk_pf::EdgeRec), st), list_a(nn * 4, st), list_b(nn * 4, st),
            path((nn + gg) * 4, st), best_path((nn + gg) * 4, st), heap((mm + gg) * sizeof(rogtk_pf::HeapItem), st);
        a.g0 = g0; a.g1 = g1; a.o0 = h_off[(size_t)g0]; a.e0 = h_eb[(size_t)g0];
        a.node = node.as<rogtk_pf::NodeRec>(); a.edge = edge.as<rogtk_pf::EdgeRec>(); a.list_a = list_a.as<int32_t>(); a.list_b = list_b.as<int32_t>();
        a.path = path.as<int32_t>(); a.best_path = best_path.as<int32_t>(); a.heap = heap.as<rogtk_pf::HeapItem>();
        const unsigned grid = (unsigned)std::min<int64_t>((gw + 127) / 128, (int64_t)ctx->sm_count * 16);
        k_pathfind<<<grid, 128, 0, st>>>(a);
        ctx->launches++;
        CK(cudaStreamSynchronize(st));   // (the wave's buffers are reused by the next one)
        g0 = g1;
    }
    CK(cudaEventRecord(ctx->ev[13], st));
    // strings to the host: lengths, then the slots that hold something
    std::vector<int32_t> h_len((size_t)ng);
    CK(cudaMemcpyAsync(h_len.data(), d_len.p, (size_t)ng * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // groups beyond PF_DEVICE_MAX_NODES: their arrays come to the host and the same procedure runs there, one group per core
    std::vector<int64_t> big;
    for (int64_t g = 0; g < ng; g++)
        if (h_off[(size_t)g + 1] - h_off[(size_t)g] > PF_DEVICE_MAX_NODES) big.push_back(g);
    std::vector<std::string> big_rows(big.size());
    if (!big.empty()) {
        struct HostGroup {
            std::vector<uint64_t> lo, hi;
            std::vector<uint16_t> counts;
            std::vector<int64_t> row_ptr;
            std::vector<int32_t> col;
        };
        std::vector<HostGroup> hg(big.size());
        for (size_t b = 0; b < big.size(); b++) {
            const int64_t g = big[b], o = h_off[(size_t)g], n = h_off[(size_t)g + 1] - o, e0 = h_eb[(size_t)g], m = h_eb[(size_t)g + 1] - e0;
            HostGroup &x = hg[b];
            x.lo.resize((size_t)n); x.counts.resize((size_t)n); x.row_ptr.resize((size_t)n + 1); x.col.resize((size_t)std::max<int64_t>(m, 1));
            CK(cudaMemcpyAsync(x.lo.data(), db->kmers_lo + o, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
            if (K > 32) {
                x.hi.resize((size_t)n);
                CK(cudaMemcpyAsync(x.hi.data(), db->kmers_hi + o, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
            }
            CK(cudaMemcpyAsync(x.counts.data(), db->counts + o, (size_t)n * 2, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(x.row_ptr.data(), db->csr_row_ptr + o, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
            if (m) CK(cudaMemcpyAsync(x.col.data(), db->csr_col + e0, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        host_parallel((int64_t)big.size(), ctx->stage_threads, 1, [&](int64_t x0, int64_t x1) {
            for (int64_t b = x0; b < x1; b++) {
                HostGroup &x = hg[(size_t)b];
                const int64_t g = big[(size_t)b], base = x.row_ptr[0];
                for (auto &v : x.row_ptr) v -= base;   // col holds this group's rows only
                rogtk_host::GroupGraph gg;
                gg.K = K; gg.n = (int32_t)x.lo.size(); gg.lo = x.lo.data(); gg.hi = K > 32 ? x.hi.data() : nullptr; gg.counts = x.counts.data();
                gg.row_ptr = x.row_ptr.data(); gg.col = x.col.data();
                std::string c;
                if (method == ROGTK_METHOD_SHORTEST_PATH) {
                    const std::string sa = an.so ? std::string((const char *)an.sd + an.so[g], (size_t)(an.so[g + 1] - an.so[g])) : std::string(an.s ? an.s : "");
                    const std::string ea = an.eo ? std::string((const char *)an.ed + an.eo[g], (size_t)(an.eo[g + 1] - an.eo[g])) : std::string(an.e ? an.e : "");
                    c = rogtk_host::shortest_path_contig(gg, sa, ea);
                } else {
                    c = rogtk_host::shortest_path_auto_contig(gg);
                }
                if ((int64_t)c.size() >= min_length) big_rows[(size_t)b] = std::move(c);
            }
        });
        for (size_t b = 0; b < big.size(); b++) h_len[(size_t)big[b]] = (int32_t)big_rows[b].size();
    }
    for (int64_t g = 0; g < ng; g++) off[(size_t)g + 1] = off[(size_t)g] + h_len[(size_t)g];
    data.resize((size_t)off[(size_t)ng]);
    if (!data.empty()) {
        uint8_t *h_out = (uint8_t *)pinned_get(ctx, out_bytes);
        try {
            CK(cudaMemcpyAsync(h_out, d_out.p, out_bytes, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            host_parallel(ng, ctx->stage_threads, 4096, [&](int64_t x0, int64_t x1) {
                for (int64_t g = x0; g < x1; g++)
                    if (h_len[(size_t)g] && h_off[(size_t)g + 1] - h_off[(size_t)g] <= PF_DEVICE_MAX_NODES)
                        memcpy(data.data() + off[(size_t)g], h_out + h_off[(size_t)g] + g * K, (size_t)h_len[(size_t)g]);
            });
        } catch (...) {
            pinned_put(ctx, h_out);
            throw;
        }
        pinned_put(ctx, h_out);
        for (size_t b = 0; b < big.size(); b++)
            if (!big_rows[b].empty()) memcpy(data.data() + off[(size_t)big[b]], big_rows[b].data(), big_rows[b].size());
    }
    if (ms_device) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        *ms_device += ms;
    }
    CK(cudaGetLastError());
}

}  // namespace
