// spectrum_nccl.cuh -- the exchange step of the global k-mer-spectrum mode inside the library (BASELINE configs[4]).
// Included at the end of rogtk_gpu.cu.
//
// The reference has no multi-process form (its closest call is one un-grouped filter_kmers over all reads, reference
// src/fracture.rs:110-116).  Protocol per rank and round:
//   local count (one group, min_coverage 1, result in HBM) -> rogtk_gpu_spectrum_exchange:
//     split the ascending local spectrum at the owner boundaries (owner = contiguous range of first-4-base buckets)
//     -> sizes to every peer, then ONE grouped ncclSend/ncclRecv per peer of a packed byte stream
//        (per destination: n k-mer words, [n high words,] n u16 partial counts, n exts bytes = 11 / 19 bytes per record)
//     -> received slices + the share carried from earlier rounds -> sort, sum partial counts (saturating), OR exts, min_coverage.
// NCCL is reached through dlopen("libnccl.so.2") (the copy already in the process, e.g. PyTorch's, when there is one): a single-GPU
// user needs no NCCL installed, and the library itself links against nothing but cudart.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; every call goes through the table below

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return;
        }
        auto sym = [&](const char *s) -> void * {
            void *p = dlsym(api.handle, s);
            if (!p) api.error = std::string("libnccl lacks ") + s;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    });
    return api;
}

#define NCK(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t _r = (call);                                                                          \
        if (_r != ncclSuccess) throw ArgFail{ROGTK_ERR_CUDA, std::string("NCCL: ") + #call + ": " + nccl_api().GetErrorString(_r)}; \
    } while (0)

struct NcclState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};
std::mutex g_nccl_mu;
std::map<rogtk_gpu_ctx *, NcclState> g_nccl;   // per context (kept outside the context struct: NCCL stays optional)

NcclState *nccl_state(rogtk_gpu_ctx *ctx)
{
    std::lock_guard<std::mutex> l(g_nccl_mu);
    auto it = g_nccl.find(ctx);
    return it == g_nccl.end() ? nullptr : &it->second;
}

size_t spec_rec_bytes(int k_eff) { return (k_eff > 32 ? 16u : 8u) + 3u; }
size_t spec_block_bytes(int64_t n, int k_eff) { return (((size_t)n * spec_rec_bytes(k_eff)) + 15u) & ~(size_t)15u; }   // 16-byte aligned blocks

}  // namespace

extern "C" int rogtk_gpu_nccl_unique_id(uint8_t *id128)
{
    NcclApi &api = nccl_api();
    if (!id128 || !api.error.empty()) return ROGTK_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (api.GetUniqueId(&id) != ncclSuccess) return ROGTK_ERR_CUDA;
    static_assert(NCCL_UNIQUE_ID_BYTES == 128, "unique id size");
    memcpy(id128, id.internal, 128);
    return ROGTK_OK;
}

extern "C" int rogtk_gpu_nccl_init(rogtk_gpu_ctx *ctx, const uint8_t *id128, int rank, int n_ranks)
{
    if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    try {
        NcclApi &api = nccl_api();
        if (!api.error.empty()) throw ArgFail{ROGTK_ERR_UNSUPPORTED, api.error};
        CK(cudaSetDevice(ctx->device));
        ncclUniqueId id;
        memcpy(id.internal, id128, 128);
        NcclState st;
        st.rank = rank;
        st.world = n_ranks;
        NCK(api.CommInitRank(&st.comm, n_ranks, id, rank));
        std::lock_guard<std::mutex> l(g_nccl_mu);
        g_nccl[ctx] = st;
    } catch (...) {
        return fail_code(ctx);
    }
    return ROGTK_OK;
}

extern "C" int rogtk_gpu_nccl_finalize(rogtk_gpu_ctx *ctx)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    std::lock_guard<std::mutex> l(g_nccl_mu);
    auto it = g_nccl.find(ctx);
    if (it == g_nccl.end()) return ROGTK_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (it->second.comm) nccl_api().CommDestroy(it->second.comm);
    g_nccl.erase(it);
    return ROGTK_OK;
}

// this rank owns the first-4-base buckets [b0, b1) (spectrum_owner); cut each into m equal ranges of about 3000 records
static void plan_merge_ranges(MergeRuns &runs, int64_t n_merge, int K, int rank, int world)
{
    if (2 * K <= 24 || n_merge <= 0) return;
    const uint32_t b0 = (uint32_t)((256u * (uint32_t)rank + (uint32_t)world - 1u) / (uint32_t)world);
    const uint32_t b1 = (uint32_t)((256u * (uint32_t)(rank + 1) + (uint32_t)world - 1u) / (uint32_t)world);
    const int64_t want = (n_merge + 2999) / 3000;
    const uint32_t nb = b1 > b0 ? b1 - b0 : 1u;
    uint32_t m = (uint32_t)((want + nb - 1) / nb);
    m = m < 1u ? 1u : m;
    if ((int64_t)m * nb <= (1 << 22)) {
        runs.total = 256u * m;
        runs.base = b0 * m;
        runs.n_ranges = (int)(nb * m);
    }
}

// Merge of several device-resident shares of the spectrum this rank owns (each ascending, one group, same K): what a driver that
// keeps its carried share in LEVELS calls -- a level is merged with the next one only when it has grown to its size, so a record
// is re-read O(log rounds) times instead of once per round.  No communication.
extern "C" int rogtk_gpu_spectrum_merge_runs(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *const *parts, int n_parts, uint32_t min_coverage,
                                             uint32_t flags, rogtk_graph_batch *out)
{
    if (!ctx || !parts || n_parts < 1 || n_parts > 64 || !out) return ROGTK_ERR_INVALID_ARG;
    const int K = parts[0]->k_eff;
    for (int i = 0; i < n_parts; i++)
        if (!parts[i] || !parts[i]->on_device || parts[i]->n_groups != 1 || parts[i]->k_eff != K || K < 1 || K > 64) return ROGTK_ERR_INVALID_ARG;
    const bool two = K > 32;
    NcclState *ns = nccl_state(ctx);
    const int world = ns ? ns->world : 1, rank = ns ? ns->rank : 0;
    DevBuf m_lo, m_hi, m_cnt, m_ext;
    MergeRuns runs;
    int64_t n_merge = 0, n_obs = 0;
    try {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CacheScope cache_scope(ctx->cache);
        CK(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        for (int i = 0; i < n_parts; i++) {
            n_merge += parts[i]->n_kmers;
            n_obs += parts[i]->n_observations;
        }
        const size_t nm = (size_t)std::max<int64_t>(n_merge, 1);
        m_lo.alloc(nm * 8, st);
        if (two) m_hi.alloc(nm * 8, st);
        m_cnt.alloc(nm * 2, st);
        m_ext.alloc(nm, st);
        int64_t at = 0;
        runs.run_off.push_back(0);
        for (int i = 0; i < n_parts; i++) {
            const int64_t c = parts[i]->n_kmers;
            runs.run_off.push_back(at + c);
            if (c == 0) continue;
            CK(cudaMemcpyAsync(m_lo.as<uint64_t>() + at, parts[i]->kmers_lo, (size_t)c * 8, cudaMemcpyDeviceToDevice, st));
            if (two) CK(cudaMemcpyAsync(m_hi.as<uint64_t>() + at, parts[i]->kmers_hi, (size_t)c * 8, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(m_cnt.as<uint16_t>() + at, parts[i]->counts, (size_t)c * 2, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(m_ext.as<uint8_t>() + at, parts[i]->exts, (size_t)c, cudaMemcpyDeviceToDevice, st));
            at += c;
        }
        CK(cudaStreamSynchronize(st));
    } catch (...) {
        return fail_code(ctx);
    }
    plan_merge_ranges(runs, n_merge, K, rank, world);
    const int rc = spectrum_merge_impl(ctx, m_lo.as<uint64_t>(), two ? m_hi.as<uint64_t>() : nullptr, m_cnt.as<uint16_t>(), m_ext.as<uint8_t>(), n_merge, K,
                                       min_coverage, flags, out, &runs);
    if (rc == ROGTK_OK) out->n_observations = n_obs;
    return rc;
}

extern "C" int rogtk_gpu_spectrum_exchange(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *local, const rogtk_graph_batch *carry,
                                           uint32_t min_coverage, uint32_t flags, rogtk_graph_batch *out, rogtk_exchange_stats *stats)
{
    if (!ctx || !local || !out || !local->on_device || local->n_groups != 1 || local->k_eff < 1 || local->k_eff > 64 ||
        (carry && (!carry->on_device || carry->k_eff != local->k_eff)))
        return ROGTK_ERR_INVALID_ARG;
    const int K = local->k_eff;
    const bool two = K > 32;
    NcclState *ns = nccl_state(ctx);
    const int world = ns ? ns->world : 1, rank = ns ? ns->rank : 0;
    DevBuf m_lo, m_hi, m_cnt, m_ext;   // what the merge reads: carry + every received slice
    int64_t n_merge = 0;
    MergeRuns runs;                    // ... as ascending runs inside this rank's share of the k-mer space
    rogtk_exchange_stats xs{};
    try {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CacheScope cache_scope(ctx->cache);
        CK(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        const int64_t n = local->n_kmers;
        std::vector<int64_t> bounds((size_t)world + 1, 0);
        bounds[(size_t)world] = n;
        if (world > 1 && n > 0) {
            DevBuf b((size_t)(world + 1) * 8, st);
            k_spectrum_bounds<<<blocks_for(world + 1, 128), 128, 0, st>>>(local->kmers_lo, two ? local->kmers_hi : nullptr, n, 2 * K, world, b.as<int64_t>());
            ctx->launches++;
            CK(cudaMemcpyAsync(bounds.data(), b.p, (size_t)(world + 1) * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        } else if (world > 1) {
            std::fill(bounds.begin(), bounds.end(), 0);
        }
        std::vector<int64_t> send_n((size_t)world), recv_n((size_t)world, 0);
        for (int d = 0; d < world; d++) send_n[(size_t)d] = bounds[(size_t)d + 1] - bounds[(size_t)d];
        recv_n[(size_t)rank] = send_n[(size_t)rank];
        cudaEvent_t e0 = ctx->ev[12], e1 = ctx->ev[13];
        DevBuf sendbuf, recvbuf;
        std::vector<size_t> soff((size_t)world + 1, 0), roff((size_t)world + 1, 0);
        if (world > 1) {
            NcclApi &api = nccl_api();
            // sizes first: one int64 to and from every peer
            DevBuf d_sn((size_t)world * 8, st), d_rn((size_t)world * 8, st);
            CK(cudaMemcpyAsync(d_sn.p, send_n.data(), (size_t)world * 8, cudaMemcpyHostToDevice, st));
            NCK(api.GroupStart());
            for (int p = 0; p < world; p++) {
                if (p == rank) continue;
                NCK(api.Send(d_sn.as<int64_t>() + p, 1, ncclInt64, p, ns->comm, st));
                NCK(api.Recv(d_rn.as<int64_t>() + p, 1, ncclInt64, p, ns->comm, st));
            }
            NCK(api.GroupEnd());
            std::vector<int64_t> tmp((size_t)world, 0);
            CK(cudaMemcpyAsync(tmp.data(), d_rn.p, (size_t)world * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (int p = 0; p < world; p++)
                if (p != rank) recv_n[(size_t)p] = tmp[(size_t)p];
            // pack: per destination one block [k-mer words | high words | counts | exts]
            for (int p = 0; p < world; p++) {
                soff[(size_t)p + 1] = soff[(size_t)p] + (p == rank ? 0 : spec_block_bytes(send_n[(size_t)p], K));
                roff[(size_t)p + 1] = roff[(size_t)p] + (p == rank ? 0 : spec_block_bytes(recv_n[(size_t)p], K));
            }
            sendbuf.alloc(std::max<size_t>(soff[(size_t)world], 16), st);
            recvbuf.alloc(std::max<size_t>(roff[(size_t)world], 16), st);
            for (int p = 0; p < world; p++) {
                const int64_t c = send_n[(size_t)p], b0 = bounds[(size_t)p];
                if (p == rank || c == 0) continue;
                uint8_t *blk = sendbuf.as<uint8_t>() + soff[(size_t)p];
                CK(cudaMemcpyAsync(blk, local->kmers_lo + b0, (size_t)c * 8, cudaMemcpyDeviceToDevice, st));
                size_t o = (size_t)c * 8;
                if (two) {
                    CK(cudaMemcpyAsync(blk + o, local->kmers_hi + b0, (size_t)c * 8, cudaMemcpyDeviceToDevice, st));
                    o += (size_t)c * 8;
                }
                CK(cudaMemcpyAsync(blk + o, local->counts + b0, (size_t)c * 2, cudaMemcpyDeviceToDevice, st));
                o += (size_t)c * 2;
                CK(cudaMemcpyAsync(blk + o, local->exts + b0, (size_t)c, cudaMemcpyDeviceToDevice, st));
            }
            CK(cudaEventRecord(e0, st));
            NCK(api.GroupStart());
            for (int p = 0; p < world; p++) {
                if (p == rank) continue;
                const size_t sb = soff[(size_t)p + 1] - soff[(size_t)p], rb = roff[(size_t)p + 1] - roff[(size_t)p];
                if (sb) NCK(api.Send(sendbuf.as<uint8_t>() + soff[(size_t)p], sb, ncclUint8, p, ns->comm, st));
                if (rb) NCK(api.Recv(recvbuf.as<uint8_t>() + roff[(size_t)p], rb, ncclUint8, p, ns->comm, st));
                xs.bytes_sent += (int64_t)sb;
                xs.bytes_received += (int64_t)rb;
            }
            NCK(api.GroupEnd());
            CK(cudaEventRecord(e1, st));
        }
        // merge input: carry, own slice, received slices
        const int64_t n_carry = carry ? carry->n_kmers : 0;
        n_merge = n_carry;
        for (int p = 0; p < world; p++) n_merge += recv_n[(size_t)p];
        const size_t nm = (size_t)std::max<int64_t>(n_merge, 1);
        m_lo.alloc(nm * 8, st);
        if (two) m_hi.alloc(nm * 8, st);
        m_cnt.alloc(nm * 2, st);
        m_ext.alloc(nm, st);
        int64_t at = 0;
        runs.run_off.push_back(0);
        auto append = [&](const uint64_t *lo, const uint64_t *hi, const uint16_t *cnt, const uint8_t *ext, int64_t c) {
            runs.run_off.push_back(at + c);
            if (c == 0) return;
            CK(cudaMemcpyAsync(m_lo.as<uint64_t>() + at, lo, (size_t)c * 8, cudaMemcpyDeviceToDevice, st));
            if (two) CK(cudaMemcpyAsync(m_hi.as<uint64_t>() + at, hi, (size_t)c * 8, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(m_cnt.as<uint16_t>() + at, cnt, (size_t)c * 2, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(m_ext.as<uint8_t>() + at, ext, (size_t)c, cudaMemcpyDeviceToDevice, st));
            at += c;
        };
        if (n_carry) append(carry->kmers_lo, carry->kmers_hi, carry->counts, carry->exts, n_carry);
        {
            const int64_t b0 = bounds[(size_t)rank];
            append(local->kmers_lo + b0, two ? local->kmers_hi + b0 : nullptr, local->counts + b0, local->exts + b0, send_n[(size_t)rank]);
        }
        for (int p = 0; p < world; p++) {
            const int64_t c = recv_n[(size_t)p];
            if (p == rank || c == 0) continue;
            const uint8_t *blk = recvbuf.as<uint8_t>() + roff[(size_t)p];
            const uint64_t *lo = reinterpret_cast<const uint64_t *>(blk);
            const uint64_t *hi = two ? lo + c : nullptr;
            const uint16_t *cnt = reinterpret_cast<const uint16_t *>(blk + (size_t)c * (two ? 16 : 8));
            const uint8_t *ext = reinterpret_cast<const uint8_t *>(cnt + c);
            append(lo, hi, cnt, ext, c);
        }
        CK(cudaStreamSynchronize(st));
        if (world > 1) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            xs.ms_exchange = ms;
        }
        xs.n_sent = n - send_n[(size_t)rank];
        xs.n_received = n_merge - n_carry - send_n[(size_t)rank];
        xs.n_ranks = world;
    } catch (...) {
        return fail_code(ctx);
    }
    if (stats) *stats = xs;
    plan_merge_ranges(runs, n_merge, K, rank, world);
    // (takes the context lock itself)
    return spectrum_merge_impl(ctx, m_lo.as<uint64_t>(), two ? m_hi.as<uint64_t>() : nullptr, m_cnt.as<uint16_t>(), m_ext.as<uint8_t>(), n_merge, K,
                               min_coverage, flags, out, &runs);
}
