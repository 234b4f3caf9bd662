// rogtk_gpu.cu -- C ABI (include/rogtk_gpu.h) and batch orchestration of the B200 k-mer count/prune engine.
//
// Replaces, for many groups per call, the filter_kmers call of reference src/fracture.rs:110-116.  There is no CPU
// fallback in this file: every entry point that computes needs a CUDA device and fails with ROGTK_ERR_CUDA otherwise.
#include "../../include/rogtk_gpu.h"

#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <cuda/std/tuple>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <atomic>
#include <condition_variable>
#include <exception>
#include <mutex>
#include <thread>
#include <string>
#include <chrono>
#include <vector>

#include <unistd.h>

#include "common.cuh"
#include "host_util.cuh"
#include "kernels_post.cuh"
#include "kernels_smem.cuh"
#include "kernels_table.cuh"
#include "kernels_unitig.cuh"
#include "scan.cuh"
#include "kernels_part.cuh"

using namespace rogtk;

// -----------------------------------------------------------------------------------------------------------------
// context
// -----------------------------------------------------------------------------------------------------------------
struct rogtk_gpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevCache *cache = nullptr;      // this context's device memory (host_util.cuh: DevCache / CacheScope)
    cudaStream_t copy_stream = nullptr;
    cudaStream_t out_stream = nullptr;
    size_t chunk_bytes = (size_t)256 << 20;   // read bytes per chunk of the overlapped host->host path
    cudaEvent_t ev[14] = {};   // 0 start, 1/2 counting, 3 end, 4/5 hot kernel, 6/7 ordering, 8/9 stats, 10/11 CSR, 12/13 unitigs
    size_t workspace_bytes = 0;
    std::mutex mu;
    std::string last_error;
    int64_t launches = 0;
    uint8_t *peek_slot = nullptr;   // mapped pinned words the counter reads land in (host_util.cuh: Peek)
    double *d_wtab = nullptr;       // edge-weight table of the device path search (pathfind_device.cuh)
    bool host_pathfind = false;     // ROGTK_HOST_PATHFIND: run the shortest-path methods on the host cores instead
    int64_t pathfind_min_groups = 16384;   // frames with fewer groups search on the host (ROGTK_PATHFIND_MIN_GROUPS): one device thread per
                                           // group takes ~70 us per node whatever the frame size, a host core ~50 us per GROUP
    int sm_count = 148;
    size_t smem_optin = 0;
    void *flush_buf = nullptr;
    bool phase_enabled = false;
    SmemTuning tune;                // environment knobs, read once in rogtk_gpu_create
    int stage_threads = 8;          // host threads that copy pageable / view input into the pinned bounce buffers (ROGTK_STAGE_THREADS)
    std::mutex pin_mu;              // the pinned cache is also used by the staging thread of the chunked path
    unsigned long long phase_cycles[16] = {};
    size_t flush_bytes = 0;
    // cache of pinned host blocks (size -> pointers) so steady-state calls do not pay cudaHostAlloc
    std::multimap<size_t, void *> pinned_free;
    std::map<void *, size_t> pinned_size;
    size_t pinned_cached = 0;        // bytes sitting in pinned_free
    size_t pinned_cache_cap = 0;     // above this, returned blocks are released (smallest first); set in rogtk_gpu_create
};

namespace {

void set_error(rogtk_gpu_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
}

// maps the exception in flight to an error code + last_error text; call inside catch (...)
int fail_code(rogtk_gpu_ctx *ctx)
{
    try {
        throw;
    } catch (const ArgFail &a) {
        set_error(ctx, "%s", a.msg.c_str());
        return a.code < 0 ? a.code : ROGTK_ERR_INTERNAL;
    } catch (const CudaFail &f) {
        set_error(ctx, "CUDA error in %s: %s", f.what, cudaGetErrorString(f.e));
        cudaGetLastError();
        return f.e == cudaErrorMemoryAllocation ? ROGTK_ERR_OOM : ROGTK_ERR_CUDA;
    } catch (const std::bad_alloc &) {
        set_error(ctx, "host out of memory");
        return ROGTK_ERR_OOM;
    } catch (...) {
        set_error(ctx, "unexpected internal error");
        return ROGTK_ERR_INTERNAL;
    }
}

void *pinned_get(rogtk_gpu_ctx *ctx, size_t bytes)
{
    std::lock_guard<std::mutex> plk(ctx->pin_mu);
    if (bytes == 0) bytes = 16;
    auto it = ctx->pinned_free.lower_bound(bytes);
    if (it != ctx->pinned_free.end() && it->first <= bytes + bytes / 4 + 4096) {
        void *p = it->second;
        ctx->pinned_cached -= it->first;
        ctx->pinned_free.erase(it);
        return p;
    }
    void *p = nullptr;
    CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    ctx->pinned_size[p] = bytes;
    return p;
}
void pinned_put(rogtk_gpu_ctx *ctx, void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> plk(ctx->pin_mu);
    auto it = ctx->pinned_size.find(p);
    if (it == ctx->pinned_size.end()) return;
    ctx->pinned_free.emplace(it->second, p);
    ctx->pinned_cached += it->second;
    // a long-lived context fed with many batch sizes must not pile up pinned memory: keep the large blocks (expensive to
    // page-lock again), release the small ones first
    while (ctx->pinned_cached > ctx->pinned_cache_cap && !ctx->pinned_free.empty()) {
        auto sm = ctx->pinned_free.begin();
        ctx->pinned_cached -= sm->first;
        ctx->pinned_size.erase(sm->second);
        cudaFreeHost(sm->second);
        ctx->pinned_free.erase(sm);
    }
}

struct BatchOwner {
    bool on_device = false;
    std::vector<void *> ptrs;
};

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); }

// ---- record ordering: LSD radix sort (cub, library call off the hot path) over exactly the key bits in use ---------------------
struct SpillKey1 {   // K <= 32: (group rank, k-mer)
    __host__ __device__ ::cuda::std::tuple<uint32_t &, uint64_t &> operator()(SpillRec &r) const { return {r.grank, r.lo}; }
};
struct SpillKey2 {   // K > 32: (group rank, hi, lo)
    __host__ __device__ ::cuda::std::tuple<uint32_t &, uint64_t &, uint64_t &> operator()(SpillRec &r) const { return {r.grank, r.hi, r.lo}; }
};
struct SpecKey1 {
    __host__ __device__ ::cuda::std::tuple<uint64_t &> operator()(SpecRec &r) const { return {r.lo}; }
};
struct SpecKey2 {
    __host__ __device__ ::cuda::std::tuple<uint64_t &, uint64_t &> operator()(SpecRec &r) const { return {r.hi, r.lo}; }
};
// orders n records by (group rank, k-mer); the result is in `recs` (which may be replaced by a new buffer).  One listed group
// (a giant UMI group, the global-spectrum local count): LSD radix sort over the k-mer's 2K bits.  Several groups: the
// extraction leaves the records grouped by listed group already (slot order), so only the k-mers inside every group need
// ordering: cub::DeviceSegmentedSort on the key columns with the record index as value (segments = seg_off[0..n_seg]), two
// stable passes for two-word keys, then one gather of the records (library calls off the hot path).
template <int W>
void sort_spill_records(cudaStream_t st, DevBuf &recs, int64_t n, int K, int64_t n_seg, const int64_t *seg_off)
{
    size_t tmp_bytes = 0;
    if (n > (int64_t)0xffffffffll) throw ArgFail{ROGTK_ERR_UNSUPPORTED, "more than 2^32 survivor records in one wave"};
    if (n_seg > 1) {
        DevBuf lo((size_t)n * 8, st), lo2((size_t)n * 8, st), hi(W == 2 ? (size_t)n * 8 : 0, st), idx((size_t)n * 4, st), idx2((size_t)n * 4, st);
        const unsigned gb = blocks_for(n, 256);
        k_rec_split<<<gb, 256, 0, st>>>(recs.as<SpillRec>(), n, lo.as<uint64_t>(), W == 2 ? hi.as<uint64_t>() : nullptr, idx.as<uint32_t>());
        CK(cub::DeviceSegmentedSort::StableSortPairs(nullptr, tmp_bytes, lo.as<uint64_t>(), lo2.as<uint64_t>(), idx.as<uint32_t>(),
                                                     idx2.as<uint32_t>(), n, n_seg, seg_off, seg_off + 1, st));
        DevBuf tmp(tmp_bytes, st);
        CK(cub::DeviceSegmentedSort::StableSortPairs(tmp.p, tmp_bytes, lo.as<uint64_t>(), lo2.as<uint64_t>(), idx.as<uint32_t>(),
                                                     idx2.as<uint32_t>(), n, n_seg, seg_off, seg_off + 1, st));
        const uint32_t *order = idx2.as<uint32_t>();
        if (W == 2) {   // second (more significant) key word, stable
            k_gather_u64<<<gb, 256, 0, st>>>(hi.as<uint64_t>(), idx2.as<uint32_t>(), n, lo.as<uint64_t>());
            size_t tb2 = 0;
            CK(cub::DeviceSegmentedSort::StableSortPairs(nullptr, tb2, lo.as<uint64_t>(), lo2.as<uint64_t>(), idx2.as<uint32_t>(),
                                                         idx.as<uint32_t>(), n, n_seg, seg_off, seg_off + 1, st));
            DevBuf tmp2(tb2, st);
            CK(cub::DeviceSegmentedSort::StableSortPairs(tmp2.p, tb2, lo.as<uint64_t>(), lo2.as<uint64_t>(), idx2.as<uint32_t>(),
                                                         idx.as<uint32_t>(), n, n_seg, seg_off, seg_off + 1, st));
            order = idx.as<uint32_t>();
        }
        DevBuf sorted((size_t)n * sizeof(SpillRec), st);
        k_rec_gather<<<gb, 256, 0, st>>>(recs.as<SpillRec>(), order, n, sorted.as<SpillRec>());
        recs = std::move(sorted);
        return;
    }
    DevBuf sorted((size_t)n * sizeof(SpillRec), st);
    const SpillRec *in = recs.as<SpillRec>();
    SpillRec *out = sorted.as<SpillRec>();
    if (W == 1) {
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, n, SpillKey1(), 0, 2 * K, st));
        DevBuf tmp(tmp_bytes, st);
        CK(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, in, out, n, SpillKey1(), 0, 2 * K, st));
    } else {
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, n, SpillKey2(), 0, 2 * K, st));
        DevBuf tmp(tmp_bytes, st);
        CK(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, in, out, n, SpillKey2(), 0, 2 * K, st));
    }
    recs = std::move(sorted);
}
inline void sort_spec_records(cudaStream_t st, const SpecRec *in, SpecRec *out, int64_t n, int K)
{
    size_t tmp_bytes = 0;
    if (K <= 32) {
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, n, SpecKey1(), 0, 2 * K, st));
        DevBuf tmp(tmp_bytes, st);
        CK(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, in, out, n, SpecKey1(), 0, 2 * K, st));
    } else {
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, n, SpecKey2(), 0, 2 * K, st));
        DevBuf tmp(tmp_bytes, st);
        CK(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, in, out, n, SpecKey2(), 0, 2 * K, st));
    }
}

int k_eff_host(int64_t k)
{
    if (k < 0) return -1;
    if (k <= 4) return 4;
    if (k <= 8) return 8;
    if (k <= 16) return 16;
    if (k <= 32) return 32;
    if (k <= 64) return 64;
    return -1;
}

// ---- spill path (HBM tables) ---------------------------------------------------------------------------------------
struct SpillWave {
    DevBuf recs;       // sorted SpillRec
    DevBuf rec_off;    // [n_list+1] exclusive scan of survivors per listed group
    int64_t list_base = 0, n_list = 0, n_recs = 0;
};

template <int W>
void run_spill(rogtk_gpu_ctx *ctx, const ColView &col, const int64_t *d_group_offsets, const int32_t *d_ids, int64_t n_list_total,
               int K, uint32_t min_cov, unsigned long long *d_nobs, int64_t *d_nsurv, std::vector<SpillWave> &waves)
{
    cudaStream_t st = ctx->stream;
    if (n_list_total == 0) return;
    // capacities of every listed group -> host decides the waves
    DevBuf rows(n_list_total * 8, st), caps(n_list_total * 8, st), ub(n_list_total * 8, st);
    CK(cudaMemsetAsync(ub.p, 0, ub.bytes, st));
    {
        // a warp per listed group; when the list is short its rows are shared by up to 1024 warps (gridDim.y)
        const int64_t per_entry = std::max<int64_t>(1, col.n_rows / std::max<int64_t>(1, n_list_total));
        const unsigned gy = (unsigned)std::min<int64_t>(1024, std::max<int64_t>(1, per_entry / 2048));
        const unsigned gx = (unsigned)std::min<int64_t>(blocks_for(n_list_total * 32, 256), (int64_t)ctx->sm_count * 32);
        k_spill_plan_obs<<<dim3(gx, gy), 256, 0, st>>>(col, d_group_offsets, d_ids, n_list_total, 0, K, ub.as<unsigned long long>());
        k_spill_plan_caps<<<blocks_for(n_list_total, 256), 256, 0, st>>>(d_group_offsets, d_ids, n_list_total, 0,
                                                                          ub.as<unsigned long long>(), rows.as<int64_t>(),
                                                                          caps.as<int64_t>());
        ctx->launches += 2;
    }
    std::vector<int64_t> h_caps((size_t)n_list_total);
    CK(cudaMemcpyAsync(h_caps.data(), caps.p, n_list_total * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const size_t slot_bytes = sizeof(typename KeySlot<W>::T) + 4;
    const int64_t budget_slots = std::max<int64_t>(1 << 16, (int64_t)(ctx->workspace_bytes / 2 / slot_bytes));
    int64_t i0 = 0;
    while (i0 < n_list_total) {
        int64_t i1 = i0, slots = 0;
        while (i1 < n_list_total && (i1 == i0 || slots + h_caps[i1] <= budget_slots)) slots += h_caps[i1++];
        const int64_t nl = i1 - i0;
        SpillWave wv;
        wv.list_base = i0;
        wv.n_list = nl;
        DevBuf row_off((nl + 1) * 8, st), slot_off((nl + 1) * 8, st), scratch(scan_scratch_bytes(nl), st);
        exclusive_scan<int64_t>(rows.as<int64_t>() + i0, nl, row_off.as<int64_t>(), scratch.as<int64_t>(), st, &ctx->launches);
        exclusive_scan<int64_t>(caps.as<int64_t>() + i0, nl, slot_off.as<int64_t>(), scratch.as<int64_t>(), st, &ctx->launches);
        int64_t h_tot[2];
        Peek pk(ctx->peek_slot, st, &ctx->launches);
        pk.add(row_off.as<int64_t>() + nl, 8, &h_tot[0]);
        pk.add(slot_off.as<int64_t>() + nl, 8, &h_tot[1]);
        pk.sync();
        const int64_t total_rows = h_tot[0], total_slots = h_tot[1];
        DevBuf keys((size_t)total_slots * sizeof(typename KeySlot<W>::T), st), vals((size_t)total_slots * 4, st),
            special((size_t)nl * 4, st);
        CK(cudaMemsetAsync(keys.p, 0xff, keys.bytes, st));
        CK(cudaMemsetAsync(vals.p, 0, vals.bytes, st));
        CK(cudaMemsetAsync(special.p, 0, special.bytes, st));
        SpillArgs a{};
        a.col = col;
        a.group_offsets = d_group_offsets;
        a.ids = d_ids;
        a.n_list = nl;
        a.list_base = i0;
        a.row_off = row_off.as<int64_t>();
        a.slot_off = slot_off.as<int64_t>();
        a.special = special.as<uint32_t>();
        a.keys = keys.p;
        a.vals = vals.as<uint32_t>();
        a.K = K;
        a.min_cov = min_cov;
        a.n_obs = d_nobs;
        a.nsurv = d_nsurv;
        if (total_rows > 0) {
            const int64_t want_blocks = (total_rows + SPILL_WARPS - 1) / SPILL_WARPS;
            const unsigned grid = (unsigned)std::min<int64_t>(want_blocks, (int64_t)ctx->sm_count * 16);
            k_spill_insert<W><<<grid, SPILL_WARPS * 32, 0, st>>>(a);
            ctx->launches++;
        }
        // pass 1: survivors per extraction warp -> exclusive scan -> exact pool size and every warp's first record
        const int64_t n_xw = (total_slots + SPILL_EXTRACT_SLOTS - 1) / SPILL_EXTRACT_SLOTS;
        const unsigned x_grid = blocks_for(n_xw, SPILL_EXTRACT_NT / 32);
        DevBuf wcount((size_t)std::max<int64_t>(n_xw, 1) * 8, st), wbase((size_t)(n_xw + 1) * 8, st), wscr(scan_scratch_bytes(n_xw), st);
        CK(cudaMemsetAsync(wcount.p, 0, wcount.bytes, st));
        a.warp_count = wcount.as<int64_t>();
        a.warp_base = wbase.as<int64_t>();
        a.count_only = 1;
        k_spill_extract<W><<<x_grid, SPILL_EXTRACT_NT, 0, st>>>(a);
        exclusive_scan<int64_t>(wcount.as<int64_t>(), n_xw, wbase.as<int64_t>(), wscr.as<int64_t>(), st, &ctx->launches);
        int64_t h_cnt = 0;
        Peek pk2(ctx->peek_slot, st, &ctx->launches);
        pk2.add(wbase.as<int64_t>() + n_xw, 8, &h_cnt);
        pk2.sync();
        wv.n_recs = h_cnt;
        wv.recs.alloc((size_t)std::max<int64_t>(1, wv.n_recs) * sizeof(SpillRec), st);
        a.pool = wv.recs.as<SpillRec>();
        a.pool_cap = wv.n_recs;
        a.count_only = 0;
        k_spill_extract<W><<<x_grid, SPILL_EXTRACT_NT, 0, st>>>(a);
        ctx->launches += 2;
        // survivors per listed group -> record offsets of the wave (the records are in slot order, i.e. grouped by listed group)
        DevBuf ns((size_t)nl * 8, st);
        k_gather_i64<<<blocks_for(nl, 256), 256, 0, st>>>(d_nsurv, d_ids, i0, nl, ns.as<int64_t>());
        ctx->launches++;
        wv.rec_off.alloc((size_t)(nl + 1) * 8, st);
        exclusive_scan<int64_t>(ns.as<int64_t>(), nl, wv.rec_off.as<int64_t>(), scratch.as<int64_t>(), st, &ctx->launches);
        if (wv.n_recs > 1) {
            sort_spill_records<W>(st, wv.recs, wv.n_recs, K, nl, wv.rec_off.as<int64_t>());
            ctx->launches += 4;
        }
        waves.push_back(std::move(wv));
        i0 = i1;
    }
}

}  // namespace

// -----------------------------------------------------------------------------------------------------------------
// lifecycle
// -----------------------------------------------------------------------------------------------------------------
extern "C" int rogtk_gpu_abi_version(void) { return ROGTK_GPU_ABI_VERSION; }

extern "C" int rogtk_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" void rogtk_gpu_destroy(rogtk_gpu_ctx *ctx);

extern "C" rogtk_gpu_ctx *rogtk_gpu_create(int device, size_t workspace_bytes, int *err)
{
    auto fail = [&](int code) -> rogtk_gpu_ctx * {
        if (err) *err = code;
        return nullptr;
    };
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(ROGTK_ERR_CUDA);   // no CPU fallback: the product path needs a GPU
    }
    if (device < 0 || device >= n) return fail(ROGTK_ERR_INVALID_ARG);
    rogtk_gpu_ctx *ctx = new rogtk_gpu_ctx();
    try {
        ctx->device = device;
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {   // the library holds sm_100a code only
            rogtk_gpu_destroy(ctx);
            return fail(ROGTK_ERR_UNSUPPORTED);   // sm_100a only
        }
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = prop.sharedMemPerBlockOptin;
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
        if (const char *e = getenv("ROGTK_CHUNK_BYTES")) ctx->chunk_bytes = (size_t)std::max(4096ll, atoll(e));
        {
            const unsigned hw = std::thread::hardware_concurrency();
            ctx->stage_threads = (int)std::min(16u, std::max(1u, hw / 2));
            if (const char *e = getenv("ROGTK_STAGE_THREADS")) ctx->stage_threads = std::max(1, atoi(e));
            if (const char *e = getenv("ROGTK_HOST_PATHFIND")) ctx->host_pathfind = atoi(e) != 0;
            if (const char *e = getenv("ROGTK_PATHFIND_MIN_GROUPS")) ctx->pathfind_min_groups = atoll(e);
        }
        if (const char *e = getenv("ROGTK_DEBUG_POOL_CAP")) ctx->tune.debug_pool_cap = atoll(e);   // tests: exhaust the survivor pool
        if (const char *e = getenv("ROGTK_WALK_RL")) {   // experiments: 0 = round-1 counting kernel, 3..6 = log2(windows per lane range)
            const int v = atoi(e);
            ctx->tune.walk_rl_bits = v <= 0 ? 0 : std::min(std::max(v, WK_MIN_RL_BITS), 6);
        }
        {   // cache of page-locked result blocks: a quarter of the host's memory unless ROGTK_PINNED_CACHE_BYTES says otherwise
            const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
            ctx->pinned_cache_cap = pages > 0 && psz > 0 ? (size_t)pages * (size_t)psz / 4 : (size_t)16 << 30;
            if (const char *e = getenv("ROGTK_PINNED_CACHE_BYTES")) ctx->pinned_cache_cap = (size_t)std::max(0ll, atoll(e));
        }
        for (auto &e : ctx->ev) CK(cudaEventCreate(&e));
        CK(cudaHostAlloc((void **)&ctx->peek_slot, Peek::SLOT_BYTES, cudaHostAllocMapped | cudaHostAllocPortable));
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        ctx->workspace_bytes = workspace_bytes ? workspace_bytes : free_b / 4;
        ctx->cache = new DevCache(device);
        smem_path_configure();
        part_path_configure();
    } catch (const CudaFail &f) {
        fprintf(stderr, "rogtk_gpu_create: %s: %s\n", f.what, cudaGetErrorString(f.e));
        cudaGetLastError();
        rogtk_gpu_destroy(ctx);   // releases whatever was created so far
        return fail(ROGTK_ERR_CUDA);
    }
    if (err) *err = ROGTK_OK;
    return ctx;
}

extern "C" void rogtk_gpu_destroy(rogtk_gpu_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->pinned_size) cudaFreeHost(kv.first);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    if (ctx->peek_slot) cudaFreeHost(ctx->peek_slot);
    if (ctx->d_wtab) cudaFree(ctx->d_wtab);
    if (ctx->cache) {   // the context's device memory goes away with it (nobody else's cached memory is touched)
        cudaDeviceSynchronize();
        delete ctx->cache;
    }
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
    delete ctx;
}

extern "C" int rogtk_gpu_trim(rogtk_gpu_ctx *ctx)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->cache) ctx->cache->release_idle();
    {
        std::lock_guard<std::mutex> plk(ctx->pin_mu);
        for (auto &kv : ctx->pinned_free) {
            ctx->pinned_size.erase(kv.second);
            cudaFreeHost(kv.second);
        }
        ctx->pinned_free.clear();
        ctx->pinned_cached = 0;
    }
    return ROGTK_OK;
}

extern "C" const char *rogtk_gpu_last_error(rogtk_gpu_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" int rogtk_gpu_k_eff(int64_t k) { return k_eff_host(k); }

extern "C" int64_t rogtk_gpu_estimate_k(const int64_t *offsets, const uint8_t *validity, int64_t n_rows)
{
    // reference src/fracture.rs:24-54
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
    const double mean = (double)total / (double)count;
    uint64_t k = (uint64_t)std::round(mean / 3.0);
    if (k % 2 == 0) k = k - 1;   // usize wrap for k == 0 as in a release build -> clamps to 63
    if (k < 11) return 11;
    if (k > 63) return 63;
    return (int64_t)k;
}

extern "C" int32_t rogtk_gpu_shard_of(uint64_t umi_id, int32_t n_shards)
{
    if (n_shards <= 1) return 0;
    return (int32_t)(splitmix_mix(umi_id + 0x9E3779B97F4A7C15ull) % (uint64_t)n_shards);
}

// -----------------------------------------------------------------------------------------------------------------
// the hot path
// -----------------------------------------------------------------------------------------------------------------
namespace {

struct Staged {   // device-resident contiguous column
    ColView col{};
    const int64_t *group_offsets = nullptr;
    const uint8_t *data_begin = nullptr;   // first / one-past-last readable byte of the data buffer
    const uint8_t *data_end = nullptr;
    std::vector<DevBuf> keep;
};

const void *to_device(rogtk_gpu_ctx *ctx, Staged &s, const void *host, size_t bytes, bool on_device)
{
    if (on_device || host == nullptr) return host;
    s.keep.emplace_back(bytes, ctx->stream);
    if (bytes) CK(cudaMemcpyAsync(s.keep.back().p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return s.keep.back().p;
}

void stage_input(rogtk_gpu_ctx *ctx, const rogtk_strcol *c, const int64_t *group_offsets, int64_t n_groups, bool on_device, Staged &s)
{
    cudaStream_t st = ctx->stream;
    const int64_t n = c->n_rows;
    const size_t vbytes = c->validity ? (size_t)((n + c->validity_bit_offset + 7) / 8) : 0;
    s.col.n_rows = n;
    s.col.validity = (const uint8_t *)to_device(ctx, s, c->validity, vbytes, on_device);
    s.col.validity_bit_offset = c->validity_bit_offset;
    s.group_offsets = (const int64_t *)to_device(ctx, s, group_offsets, (size_t)(n_groups + 1) * 8, on_device);
    if (c->layout == ROGTK_UTF8_I32 || c->layout == ROGTK_UTF8_I64) {
        if (c->n_buffers != 1 || !c->data || !c->data_len) throw ArgFail{ROGTK_ERR_INVALID_ARG, "offset layouts need exactly one data buffer"};
        const bool o64 = c->layout == ROGTK_UTF8_I64;
        s.col.off64 = o64;
        s.col.offsets = to_device(ctx, s, c->offsets_or_views, (size_t)(n + 1) * (o64 ? 8 : 4), on_device);
        s.col.data = (const uint8_t *)to_device(ctx, s, c->data[0], (size_t)c->data_len[0], on_device);
        s.data_begin = s.col.data;
        s.data_end = s.col.data + c->data_len[0];
        return;
    }
    if (c->layout != ROGTK_UTF8_VIEW) throw ArgFail{ROGTK_ERR_INVALID_ARG, "unknown string layout"};
    // Utf8View: gather on the device into a contiguous LargeUtf8-like column
    const View16 *views = (const View16 *)to_device(ctx, s, c->offsets_or_views, (size_t)n * 16, on_device);
    if (c->n_buffers < 0 || (c->n_buffers > 0 && (!c->data || !c->data_len))) throw ArgFail{ROGTK_ERR_INVALID_ARG, "view layout: null buffer table"};
    const size_t nb = (size_t)std::max(1, c->n_buffers);
    // buffer table for the device: nb pointers, then nb lengths (a null / empty variadic buffer keeps its index, with length 0)
    std::vector<uint64_t> h_tab(2 * nb, 0);
    for (int b = 0; b < c->n_buffers; b++) {
        const int64_t len = c->data[b] ? c->data_len[b] : 0;
        if (len < 0) throw ArgFail{ROGTK_ERR_INVALID_ARG, "view layout: negative buffer length"};
        h_tab[b] = (uint64_t) reinterpret_cast<uintptr_t>(to_device(ctx, s, c->data[b], (size_t)len, on_device));
        h_tab[nb + b] = (uint64_t)len;
    }
    s.keep.emplace_back(h_tab.size() * 8, st);
    const uint8_t *const *d_bufs = s.keep.back().as<const uint8_t *>();
    const int64_t *d_buf_len = s.keep.back().as<int64_t>() + nb;
    CK(cudaMemcpyAsync(s.keep.back().p, h_tab.data(), h_tab.size() * 8, cudaMemcpyHostToDevice, st));
    DevBuf lens((size_t)std::max<int64_t>(1, n) * 8, st), scratch(scan_scratch_bytes(n), st), verr(16, st);
    CK(cudaMemsetAsync(verr.p, 0, 16, st));
    s.keep.emplace_back((size_t)(n + 1) * 8, st);
    int64_t *offs = s.keep.back().as<int64_t>();
    k_view_len<<<blocks_for(n, 256), 256, 0, st>>>(views, n, s.col.validity, s.col.validity_bit_offset, c->n_buffers, d_buf_len,
                                                   lens.as<int64_t>(), verr.as<int>());
    exclusive_scan<int64_t>(lens.as<int64_t>(), n, offs, scratch.as<int64_t>(), st, &ctx->launches);
    int64_t total = 0;
    int h_verr = 0;
    CK(cudaMemcpyAsync(&total, offs + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_verr, verr.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));   // also keeps h_tab alive long enough
    if (h_verr) throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: a string view names a missing data buffer or bytes beyond its end;"};
    s.keep.emplace_back((size_t)total, st);
    uint8_t *data = s.keep.back().as<uint8_t>();
    k_view_gather<<<blocks_for(n * 32, 256), 256, 0, st>>>(views, n, d_bufs, d_buf_len, offs, data);
    ctx->launches += 2;
    s.col.off64 = 1;
    s.col.offsets = offs;
    s.col.data = data;
    s.data_begin = data;
    s.data_end = data + total;
}

void fill_empty_batch(rogtk_gpu_ctx *ctx, int64_t n_groups, bool out_dev, uint32_t flags, rogtk_graph_batch *out, BatchOwner *own)
{
    // k > 64: every group yields nothing (reference src/fracture.rs:212-215)
    const size_t ob = (size_t)(n_groups + 1) * 8;
    if (out_dev) {
        void *p;
        CK(dev_alloc(&p, ob, ctx->stream));
        CK(cudaMemsetAsync(p, 0, ob, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        out->kmer_offsets = (int64_t *)p;
        own->ptrs.push_back(p);
    } else {
        void *p = pinned_get(ctx, ob);
        memset(p, 0, ob);
        out->kmer_offsets = (int64_t *)p;
        own->ptrs.push_back(p);
        if (flags & ROGTK_EMIT_STATS) {
            void *q = pinned_get(ctx, (size_t)n_groups * 12 + 16);
            memset(q, 0, (size_t)n_groups * 12 + 16);
            out->stats = (uint32_t *)q;
            own->ptrs.push_back(q);
        }
    }
}

// Successor CSR of a dense batch: row_ptr[n_kmers + 1], cols[n_edges].  Per-UMI batches (every group fits a CTA-sized search) use
// the per-group shared-memory kernel; a batch with a very large group (global spectrum, giant UMIs) keeps the node-parallel one.
void emit_csr(rogtk_gpu_ctx *ctx, int64_t n_groups, int64_t n_kmers, const int64_t *off, const uint64_t *lo, const uint64_t *hi,
              const uint8_t *exts, int K, DevBuf &rowptr, DevBuf &cols, int64_t &n_edges)
{
    cudaStream_t st = ctx->stream;
    rowptr.alloc((size_t)(n_kmers + 1) * 8, st);
    n_edges = 0;
    if (n_kmers == 0 || n_groups == 0) {
        CK(cudaMemsetAsync(rowptr.p, 0, 8, st));
        cols.alloc(16, st);
        return;
    }
    unsigned long long max_n = 0;
    {
        DevBuf mx(8, st);
        CK(cudaMemsetAsync(mx.p, 0, 8, st));
        k_max_group<<<ctx->sm_count * 4, 256, 0, st>>>(off, n_groups, mx.as<unsigned long long>());
        Peek pk(ctx->peek_slot, st, &ctx->launches);
        pk.add(mx.p, 8, &max_n);
        pk.sync();
        ctx->launches++;
    }
    Peek pk(ctx->peek_slot, st, &ctx->launches);
    if (max_n <= 16ull * CSR_CAP) {
        DevBuf ge((size_t)n_groups * 8, st), gb((size_t)(n_groups + 1) * 8, st), scr(scan_scratch_bytes(n_groups), st);
        const unsigned grid = (unsigned)std::min<int64_t>(n_groups, (int64_t)ctx->sm_count * 8);
        if (K > 32) k_csr_group<2, false><<<grid, CSR_NT, 0, st>>>(n_groups, off, lo, hi, exts, K, ge.as<int64_t>(), nullptr, nullptr, nullptr);
        else k_csr_group<1, false><<<grid, CSR_NT, 0, st>>>(n_groups, off, lo, hi, exts, K, ge.as<int64_t>(), nullptr, nullptr, nullptr);
        exclusive_scan<int64_t>(ge.as<int64_t>(), n_groups, gb.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
        pk.add(gb.as<int64_t>() + n_groups, 8, &n_edges);
        pk.sync();
        cols.alloc((size_t)std::max<int64_t>(n_edges, 1) * 4, st);
        if (K > 32)
            k_csr_group<2, true><<<grid, CSR_NT, 0, st>>>(n_groups, off, lo, hi, exts, K, nullptr, gb.as<int64_t>(), rowptr.as<int64_t>(), cols.as<int32_t>());
        else
            k_csr_group<1, true><<<grid, CSR_NT, 0, st>>>(n_groups, off, lo, hi, exts, K, nullptr, gb.as<int64_t>(), rowptr.as<int64_t>(), cols.as<int32_t>());
        CK(cudaMemcpyAsync(rowptr.as<int64_t>() + n_kmers, gb.as<int64_t>() + n_groups, 8, cudaMemcpyDeviceToDevice, st));
        ctx->launches += 2;
        return;
    }
    DevBuf deg((size_t)n_kmers * 4, st), scr(scan_scratch_bytes(n_kmers), st);
    k_csr<<<blocks_for(n_kmers, 256), 256, 0, st>>>(n_kmers, n_groups, off, lo, hi, exts, K, deg.as<int32_t>(), nullptr, nullptr);
    exclusive_scan<int32_t>(deg.as<int32_t>(), n_kmers, rowptr.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
    pk.add(rowptr.as<int64_t>() + n_kmers, 8, &n_edges);
    pk.sync();
    cols.alloc((size_t)std::max<int64_t>(n_edges, 1) * 4, st);
    k_csr<<<blocks_for(n_kmers, 256), 256, 0, st>>>(n_kmers, n_groups, off, lo, hi, exts, K, nullptr, rowptr.as<int64_t>(), cols.as<int32_t>());
    ctx->launches += 2;
}

struct DeviceResult {   // dense, group-ordered result of one pipeline run, still in HBM
    DevBuf off, lo, hi, exts, counts, stats, rowptr, cols;
    int64_t n_kmers = 0, n_edges = 0, n_obs = 0, n_spill = 0, n_multi = 0;
    float ms_hot = 0.f;
    bool has_stats = false, has_csr = false;   // ev[8..11] were recorded
};

// The device pipeline on ctx->stream: count (shared-memory path, then HBM-table path for deferred groups), scan, gather,
// optional statistics / CSR.  Records ctx->ev[0] (start), ev[1]/ev[2] (counting), ev[3] (end); returns without a final sync.
void run_pipeline(rogtk_gpu_ctx *ctx, const Staged &in, int64_t n_groups, int K, uint32_t min_coverage, uint32_t flags,
                  DeviceResult &r)
{
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], st));
    DevBuf nsurv((size_t)(n_groups + 1) * 8, st), misc(64, st);
    CK(cudaMemsetAsync(nsurv.p, 0, nsurv.bytes, st));
    CK(cudaMemsetAsync(misc.p, 0, 64, st));
    unsigned long long *d_nobs = misc.as<unsigned long long>();
    {   // reject malformed offsets before any kernel dereferences them
        int *d_err = reinterpret_cast<int *>(misc.as<unsigned long long>() + 4);
        const int64_t data_len = (int64_t)(in.data_end - in.col.data);   // virtual-base aware: offsets are absolute
        k_validate<<<ctx->sm_count * 4, 256, 0, st>>>(in.col, data_len, in.group_offsets, n_groups, d_err);
        ctx->launches++;
        int h_err = 0;
        Peek pk(ctx->peek_slot, st, &ctx->launches);
        pk.add(d_err, 4, &h_err);
        pk.sync();
        if (h_err) {
            char msg[160];
            snprintf(msg, sizeof(msg), "malformed input:%s%s%s", (h_err & 1) ? " string offsets not monotone;" : "",
                     (h_err & 2) ? " string offsets exceed the data buffer;" : "",
                     (h_err & 4) ? " group_offsets not monotone or outside the rows;" : "");
            throw ArgFail{ROGTK_ERR_INVALID_ARG, msg};
        }
    }

    // ---- counting -------------------------------------------------------------------------------------------
    CK(cudaEventRecord(ctx->ev[1], st));
    SmemResult sm;            // survivors of the shared-memory path (pool, per-group start) + deferred groups
    std::vector<SpillWave> waves;
    const int32_t *d_spill_ids = nullptr;
    int64_t n_spill = n_groups;
    if (!(flags & ROGTK_FORCE_SPILL) && n_groups > 0) {
        run_smem_path(ctx->stream, ctx->peek_slot, ctx->sm_count, ctx->workspace_bytes, in.col, in.data_begin, in.data_end, in.group_offsets,
                      n_groups, K, min_coverage, d_nobs, nsurv.as<int64_t>(), sm, &ctx->launches, ctx->ev[4], ctx->ev[5],
                      ctx->phase_enabled ? ctx->phase_cycles : nullptr, !(flags & ROGTK_NO_MULTIPASS), ctx->tune);
        d_spill_ids = sm.deferred_ids;
        n_spill = sm.n_deferred;
    }
    // groups beyond the multi-pass kernel: partition-then-count in shared memory (kernels_part.cuh); what overflows there goes on
    PartResult part;
    if (n_spill > 0 && sm.pool_cursor && !(flags & (ROGTK_FORCE_SPILL | ROGTK_NO_MULTIPASS | ROGTK_NO_PARTITION))) {
        if (K <= 32)
            run_part_path_w<1>(st, ctx->peek_slot, ctx->sm_count, ctx->workspace_bytes, in.col, in.data_begin, in.data_end, in.group_offsets, d_spill_ids,
                               n_spill, K, min_coverage, d_nobs, nsurv.as<int64_t>(), sm.pool_cursor, sm.pool_cap, const_cast<uint64_t *>(sm.pool_lo),
                               const_cast<uint64_t *>(sm.pool_hi), const_cast<uint8_t *>(sm.pool_exts), const_cast<uint16_t *>(sm.pool_counts), part,
                               &ctx->launches);
        else
            run_part_path_w<2>(st, ctx->peek_slot, ctx->sm_count, ctx->workspace_bytes, in.col, in.data_begin, in.data_end, in.group_offsets, d_spill_ids,
                               n_spill, K, min_coverage, d_nobs, nsurv.as<int64_t>(), sm.pool_cursor, sm.pool_cap, const_cast<uint64_t *>(sm.pool_lo),
                               const_cast<uint64_t *>(sm.pool_hi), const_cast<uint8_t *>(sm.pool_exts), const_cast<uint16_t *>(sm.pool_counts), part,
                               &ctx->launches);
        d_spill_ids = part.failed_ids;
        n_spill = part.n_failed;
    }
    if (n_spill > 0) {
        if (K <= 32)
            run_spill<1>(ctx, in.col, in.group_offsets, d_spill_ids, n_spill, K, min_coverage, d_nobs, nsurv.as<int64_t>(), waves);
        else
            run_spill<2>(ctx, in.col, in.group_offsets, d_spill_ids, n_spill, K, min_coverage, d_nobs, nsurv.as<int64_t>(), waves);
    }
    CK(cudaEventRecord(ctx->ev[2], st));

    // ---- dense, group-ordered output ----------------------------------------------------------------------------
    DevBuf scratch(scan_scratch_bytes(n_groups), st);
    r.off.alloc((size_t)(n_groups + 1) * 8, st);
    exclusive_scan<int64_t>(nsurv.as<int64_t>(), n_groups, r.off.as<int64_t>(), scratch.as<int64_t>(), st, &ctx->launches);
    int64_t n_kmers = 0;
    unsigned long long h_nobs = 0;
    Peek pk(ctx->peek_slot, st, &ctx->launches);
    pk.add(r.off.as<int64_t>() + n_groups, 8, &n_kmers);
    pk.add(d_nobs, 8, &h_nobs);
    pk.sync();
    const size_t nk = (size_t)std::max<int64_t>(n_kmers, 1);
    r.lo.alloc(nk * 8, st);
    if (K > 32) r.hi.alloc(nk * 8, st);
    r.exts.alloc(nk, st);
    r.counts.alloc(nk * 2, st);
    uint64_t *p_hi = K > 32 ? r.hi.as<uint64_t>() : nullptr;
    CK(cudaEventRecord(ctx->ev[6], st));
    if (sm.n_pool_groups > 0)
        run_sort_gather(st, ctx->peek_slot, ctx->sm_count, K, n_groups, sm.pool_start, r.off.as<int64_t>(), sm.pool_lo, sm.pool_hi, sm.pool_exts,
                        sm.pool_counts, r.lo.as<uint64_t>(), p_hi, r.exts.as<uint8_t>(), r.counts.as<uint16_t>(), &ctx->launches,
                        sm.segs, sm.seg_count, sm.seg_list, sm.n_seg_groups);
    for (const PartSegSet &ps : part.sets)   // segments of the partitioned groups (same pool; ps.stride segment slots per group)
        run_sort_gather(st, ctx->peek_slot, ctx->sm_count, K, n_groups, sm.pool_start, r.off.as<int64_t>(), sm.pool_lo, sm.pool_hi, sm.pool_exts,
                        sm.pool_counts, r.lo.as<uint64_t>(), p_hi, r.exts.as<uint8_t>(), r.counts.as<uint16_t>(), &ctx->launches, ps.segs,
                        ps.seg_count, ps.seg_list, ps.n_entries, ps.stride, false);
    for (auto &wv : waves) {
        if (wv.n_recs == 0) continue;
        k_spill_scatter<<<blocks_for(wv.n_recs, 256), 256, 0, st>>>(wv.recs.as<SpillRec>(), wv.n_recs, d_spill_ids, wv.list_base,
                                                                    wv.rec_off.as<int64_t>(), r.off.as<int64_t>(),
                                                                    r.lo.as<uint64_t>(), p_hi, r.exts.as<uint8_t>(),
                                                                    r.counts.as<uint16_t>());
        ctx->launches++;
    }

    CK(cudaEventRecord(ctx->ev[7], st));
    // ---- optional node statistics and successor CSR ---------------------------------------------------------------------
    int64_t n_edges = 0;
    if (flags & ROGTK_EMIT_STATS) {
        CK(cudaEventRecord(ctx->ev[8], st));
        r.has_stats = true;
        r.stats.alloc((size_t)n_groups * 12 + 16, st);
        if (n_groups > 0) {
            k_stats<<<ctx->sm_count * 8, 256, 0, st>>>(n_groups, r.off.as<int64_t>(), r.exts.as<uint8_t>(), r.stats.as<uint32_t>());
            ctx->launches++;
        }
        CK(cudaEventRecord(ctx->ev[9], st));
    }
    if (flags & ROGTK_EMIT_CSR) {
        CK(cudaEventRecord(ctx->ev[10], st));
        r.has_csr = true;
        emit_csr(ctx, n_groups, n_kmers, r.off.as<int64_t>(), r.lo.as<uint64_t>(), p_hi, r.exts.as<uint8_t>(), K, r.rowptr, r.cols, n_edges);
        CK(cudaEventRecord(ctx->ev[11], st));
    }
    CK(cudaEventRecord(ctx->ev[3], st));
    r.n_kmers = n_kmers;
    r.n_edges = n_edges;
    r.n_obs = (int64_t)h_nobs;
    r.n_spill = n_spill;
    r.n_multi = sm.n_seg_groups - (sm.n_seg_groups ? sm.n_deferred : 0) + (part.n_entries - part.n_failed);
    r.ms_hot = sm.ms_kernel;
}

// growable pinned host array of the final result
struct HostArr {
    uint8_t *p = nullptr;
    size_t cap = 0;   // bytes
};
void host_reserve(rogtk_gpu_ctx *ctx, HostArr &a, size_t need, size_t used, cudaStream_t drain)
{
    if (need <= a.cap) return;
    size_t ncap = std::max(need + need / 4 + 4096, a.cap + a.cap / 2);
    uint8_t *np_ = (uint8_t *)pinned_get(ctx, ncap);
    {
        std::lock_guard<std::mutex> plk(ctx->pin_mu);
        ncap = ctx->pinned_size[np_];
    }
    if (a.p) {
        CK(cudaStreamSynchronize(drain));   // copies into the old block must have landed
        if (used) memcpy(np_, a.p, used);
        pinned_put(ctx, a.p);
    }
    a.p = np_;
    a.cap = ncap;
}

void compress_unitigs_impl(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *in, int64_t min_length, uint32_t flags, rogtk_contig_batch *out,
                           BatchOwner *own);

// What the chunked path does with a chunk's device-resident node arrays instead of copying them to the host: the Compression
// method (unitigs + the expression's contig choice) on the device, and only the resulting strings travel back.
struct ChunkPost {
    int64_t min_length = 0;
    bool only_largest = true;
    HostArr h_soff, h_sdata, h_sflags;   // [n_groups+1] int64 | bytes | [n_groups] flags (page-locked, handed to the caller)
    int64_t n_bytes = 0;
    float ms_unitig = 0.f;
};

// one string per group out of a device-resident contig batch (roff: [ng+1] chunk-local offsets)
void contig_rows_device(rogtk_gpu_ctx *ctx, const rogtk_contig_batch &cb, int64_t min_length, bool only_largest, DevBuf &roff, DevBuf &rows,
                        DevBuf &rflags, int64_t &nbytes)
{
    cudaStream_t st = ctx->stream;
    const int64_t ng = cb.n_groups;
    const size_t ngn = (size_t)std::max<int64_t>(ng, 1);
    DevBuf rlen(ngn * 8, st), scr(scan_scratch_bytes(ng), st);
    roff.alloc((ngn + 1) * 8, st);
    rflags.alloc(ngn, st);
    if (ng > 0)
        k_contig_row_len<<<blocks_for(ng, 256), 256, 0, st>>>(ng, cb.unitig_offsets, cb.seq_offsets, cb.largest, min_length, only_largest ? 1 : 0,
                                                              rlen.as<int64_t>(), only_largest ? nullptr : rflags.as<uint8_t>());
    if (only_largest && ng > 0) CK(cudaMemcpyAsync(rflags.p, cb.largest_flags, (size_t)ng, cudaMemcpyDeviceToDevice, st));
    exclusive_scan<int64_t>(rlen.as<int64_t>(), ng, roff.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
    nbytes = 0;
    Peek pk(ctx->peek_slot, st, &ctx->launches);
    pk.add(roff.as<int64_t>() + ng, 8, &nbytes);
    pk.sync();
    rows.alloc((size_t)std::max<int64_t>(nbytes, 1), st);
    if (ng > 0)
        k_contig_row_fill<<<blocks_for(ng * 32, 256), 256, 0, st>>>(ng, cb.unitig_offsets, cb.seq_offsets, cb.seq, cb.largest, min_length,
                                                                   only_largest ? 1 : 0, roff.as<int64_t>(), rows.as<uint8_t>());
    ctx->launches += 2;
}

// ---- host staging of the chunked path ------------------------------------------------------------------------------------------
bool host_ptr_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
// fn(begin, end) over [0, n) on up to nt threads (the calling thread takes the first slice)
template <typename F>
void host_parallel(int64_t n, int nt, int64_t min_per_thread, F &&fn)
{
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n / std::max<int64_t>(1, min_per_thread)));
    if (nt <= 1) {
        fn((int64_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    std::exception_ptr err;
    std::mutex em;
    for (int t = 1; t < nt; t++)
        th.emplace_back([&, t]() {
            try {
                fn(n * t / nt, n * (t + 1) / nt);
            } catch (...) {
                std::lock_guard<std::mutex> l(em);
                err = std::current_exception();
            }
        });
    try {
        fn((int64_t)0, n / nt);
    } catch (...) {
        std::lock_guard<std::mutex> l(em);
        err = std::current_exception();
    }
    for (auto &t : th) t.join();
    if (err) std::rethrow_exception(err);
}

// Host Arrow buffers in, host arrays out, for batches much larger than ctx->chunk_bytes: the groups are cut into chunks of
// ~chunk_bytes of read data; H2D of chunk c+1 and D2H of chunk c-1 run on their own streams while chunk c is being counted.
// Page-locked offset-layout input is copied straight from the caller's buffers.  Anything else -- PAGEABLE buffers (what Polars and
// pyarrow hand over: a cudaMemcpyAsync from pageable memory is staged synchronously by the driver and stalls the pipeline) and the
// Utf8View layout (rows scattered over several data buffers) -- goes through two page-locked bounce buffers that a staging thread
// fills one chunk ahead with ctx->stage_threads copy threads; views are gathered into the contiguous offsets layout on the way, so the
// device sees the same column either way.
void run_chunked(rogtk_gpu_ctx *ctx, const rogtk_strcol *c, const int64_t *group_offsets, int64_t n_groups, int K,
                 uint32_t min_coverage, uint32_t flags, rogtk_graph_batch *out, BatchOwner *own, ChunkPost *post = nullptr)
{
    const bool is_view = c->layout == ROGTK_UTF8_VIEW;
    const bool o64 = is_view || c->layout == ROGTK_UTF8_I64;   // (gathered views get int64 offsets)
    const int osz = o64 ? 8 : 4;
    const int nt = ctx->stage_threads;
    const View16 *views = is_view ? (const View16 *)c->offsets_or_views : nullptr;
    auto off_at = [&](int64_t row) -> int64_t {
        return c->layout == ROGTK_UTF8_I64 ? ((const int64_t *)c->offsets_or_views)[row] : (int64_t)((const int32_t *)c->offsets_or_views)[row];
    };
    auto row_valid = [&](int64_t r) -> bool {
        if (!c->validity) return true;
        const int64_t b = r + c->validity_bit_offset;
        return (c->validity[b >> 3] >> (b & 7)) & 1;
    };
    const bool direct = !is_view && host_ptr_pinned(c->data[0]) && host_ptr_pinned(c->offsets_or_views);
    // byte position of every group boundary in the (virtual) contiguous column.  Offsets layout: the caller's offsets, checked here
    // because the host itself indexes with them (chunk cuts, copy ranges); everything between two boundaries is checked on the
    // device (k_validate) against the copied slice.  Views: the running sum of the valid rows' lengths, every view checked.
    std::vector<int64_t> gpos;
    if (is_view) {
        if (c->n_buffers < 0 || (c->n_buffers > 0 && (!c->data || !c->data_len))) throw ArgFail{ROGTK_ERR_INVALID_ARG, "view layout: null buffer table"};
        gpos.assign((size_t)n_groups + 1, 0);
        std::atomic<int> bad{0};
        host_parallel(n_groups, nt, 4096, [&](int64_t g0, int64_t g1) {
            for (int64_t g = g0; g < g1; g++) {
                int64_t sum = 0;
                for (int64_t r = group_offsets[g]; r < group_offsets[g + 1]; r++) {
                    if (!row_valid(r)) continue;
                    const View16 v = views[r];
                    if (v.len > 12u && (v.len >= 0x80000000u || v.b >= (uint32_t)c->n_buffers || !c->data[v.b] ||
                                        (int64_t)v.c + (int64_t)v.len > c->data_len[v.b]))
                        bad = 1;
                    else
                        sum += (int64_t)v.len;
                }
                gpos[(size_t)g + 1] = sum;
            }
        });
        if (bad) throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: a string view names a missing data buffer or bytes beyond its end;"};
        for (int64_t g = 0; g < n_groups; g++) gpos[(size_t)g + 1] += gpos[(size_t)g];
    }
    auto gbyte = [&](int64_t g) -> int64_t { return is_view ? gpos[(size_t)g] : off_at(group_offsets[g]); };
    std::vector<int64_t> cut{0};
    {
        int64_t start_byte = gbyte(0);
        int64_t prev_end = start_byte;
        const int64_t limit = is_view ? gpos[(size_t)n_groups] : c->data_len[0];
        if (start_byte < 0 || start_byte > limit)
            throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: string offsets exceed the data buffer;"};
        for (int64_t g = 0; g < n_groups; g++) {
            const int64_t end_byte = gbyte(g + 1);
            if (end_byte < prev_end) throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: string offsets not monotone;"};
            if (end_byte > limit) throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: string offsets exceed the data buffer;"};
            prev_end = end_byte;
            if (end_byte - start_byte >= (int64_t)ctx->chunk_bytes && g + 1 < n_groups) {
                cut.push_back(g + 1);
                start_byte = end_byte;
            }
        }
        cut.push_back(n_groups);
    }
    const int n_chunks = (int)cut.size() - 1;
    cudaStream_t s_in = ctx->copy_stream, s_comp = ctx->stream, s_out = ctx->out_stream;
    struct InSet {
        DevBuf data, offs, goff, valid;
        cudaEvent_t ready = nullptr, consumed = nullptr;
        Staged st;
    } sets[2];
    for (auto &x : sets) {
        CK(cudaEventCreateWithFlags(&x.ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&x.consumed, cudaEventDisableTiming));
    }
    cudaEvent_t comp_done;
    CK(cudaEventCreateWithFlags(&comp_done, cudaEventDisableTiming));
    struct Guard {
        InSet *s; cudaEvent_t e;
        ~Guard() { for (int i = 0; i < 2; i++) { cudaEventDestroy(s[i].ready); cudaEventDestroy(s[i].consumed); } cudaEventDestroy(e); }
    } guard{sets, comp_done};

    // ROGTK_TRACE=1: per-chunk timeline (H2D / count / D2H begin and end, ms from the first copy) on stderr
    const bool trace = getenv("ROGTK_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto tmark = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        CK(cudaEventRecord(e, s));
        tev.push_back(e);
    };
    std::vector<double> host_ms;
    const auto host_t0 = std::chrono::steady_clock::now();
    auto hmark = [&]() {
        if (trace) host_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
    };

    // ---- bounce buffers + staging thread (pageable / view input) -------------------------------------------------------------------
    struct Bounce {
        uint8_t *data = nullptr, *offs = nullptr, *goff = nullptr, *valid = nullptr;
        size_t data_cap = 0, offs_cap = 0, goff_cap = 0, valid_cap = 0;
        cudaEvent_t h2d_done = nullptr;
    } bounce[2];
    struct StageSync {
        std::mutex m;
        std::condition_variable cv;
        std::vector<char> filled, issued;
        bool abort = false;
        std::exception_ptr err;
    } ss;
    ss.filled.assign((size_t)n_chunks, 0);
    ss.issued.assign((size_t)n_chunks, 0);
    auto bounce_fit = [&](uint8_t *&p, size_t &cap, size_t need) {
        if (need <= cap) return;
        if (p) pinned_put(ctx, p);
        p = (uint8_t *)pinned_get(ctx, need + need / 8 + 4096);
        cap = need + need / 8 + 4096;
    };
    auto fill_chunk = [&](int ci) {   // staging thread: chunk ci -> bounce[ci & 1]
        Bounce &bb = bounce[ci & 1];
        const int64_t g0 = cut[ci], g1 = cut[ci + 1];
        const int64_t r0 = group_offsets[g0], r1 = group_offsets[g1];
        const int64_t b0 = gbyte(g0), b1 = gbyte(g1);
        bounce_fit(bb.data, bb.data_cap, (size_t)(b1 - b0) + 32);
        bounce_fit(bb.offs, bb.offs_cap, (size_t)(r1 - r0 + 1) * osz);
        bounce_fit(bb.goff, bb.goff_cap, (size_t)(g1 - g0 + 1) * 8);
        memcpy(bb.goff, group_offsets + g0, (size_t)(g1 - g0 + 1) * 8);
        if (c->validity) {
            const int64_t by0 = (r0 + c->validity_bit_offset) >> 3, by1 = (r1 + c->validity_bit_offset + 7) >> 3;
            bounce_fit(bb.valid, bb.valid_cap, (size_t)(by1 - by0) + 8);
            memcpy(bb.valid, c->validity + by0, (size_t)(by1 - by0));
        }
        if (!is_view) {
            memcpy(bb.offs, (const uint8_t *)c->offsets_or_views + (size_t)r0 * osz, (size_t)(r1 - r0 + 1) * osz);
            host_parallel(b1 - b0, nt, 4 << 20, [&](int64_t x0, int64_t x1) { memcpy(bb.data + x0, c->data[0] + b0 + x0, (size_t)(x1 - x0)); });
            return;
        }
        // views: offsets of the rows (absolute positions in the virtual column), then the bytes
        int64_t *offs = (int64_t *)bb.offs;
        const int64_t nr = r1 - r0;
        const int slices = (int)std::max<int64_t>(1, std::min<int64_t>(nt, nr / 8192));
        std::vector<int64_t> sbase((size_t)slices + 1, 0);
        auto vlen = [&](int64_t r) -> int64_t { return row_valid(r) ? (int64_t)views[r].len : 0; };
        host_parallel(slices, slices, 1, [&](int64_t s0, int64_t s1) {
            for (int64_t sl = s0; sl < s1; sl++) {
                int64_t sum = 0;
                for (int64_t r = r0 + nr * sl / slices; r < r0 + nr * (sl + 1) / slices; r++) sum += vlen(r);
                sbase[(size_t)sl + 1] = sum;
            }
        });
        for (int sl = 0; sl < slices; sl++) sbase[(size_t)sl + 1] += sbase[(size_t)sl];
        host_parallel(slices, slices, 1, [&](int64_t s0, int64_t s1) {
            for (int64_t sl = s0; sl < s1; sl++) {
                int64_t pos = sbase[(size_t)sl];
                for (int64_t r = r0 + nr * sl / slices; r < r0 + nr * (sl + 1) / slices; r++) {
                    offs[r - r0] = b0 + pos;
                    const int64_t len = vlen(r);
                    if (len) {
                        const View16 &v = views[r];
                        const uint8_t *src = v.len <= 12u ? reinterpret_cast<const uint8_t *>(&v) + 4 : c->data[v.b] + v.c;
                        memcpy(bb.data + pos, src, (size_t)len);
                        pos += len;
                    }
                }
            }
        });
        offs[nr] = b1;
    };
    std::thread stager;
    struct StagerGuard {   // an exception on the main thread must not unwind locals the staging thread still uses
        StageSync &ss;
        std::thread &t;
        rogtk_gpu_ctx *ctx;
        Bounce *b;
        ~StagerGuard()
        {
            {
                std::lock_guard<std::mutex> l(ss.m);
                ss.abort = true;
            }
            ss.cv.notify_all();
            if (t.joinable()) t.join();
            for (int i = 0; i < 2; i++) {
                if (b[i].h2d_done) {
                    cudaEventSynchronize(b[i].h2d_done);
                    cudaEventDestroy(b[i].h2d_done);
                }
                for (uint8_t *p : {b[i].data, b[i].offs, b[i].goff, b[i].valid})
                    if (p) pinned_put(ctx, p);
            }
        }
    } stager_guard{ss, stager, ctx, bounce};
    if (!direct) {
        for (auto &bb : bounce) CK(cudaEventCreateWithFlags(&bb.h2d_done, cudaEventDisableTiming));
        const int dev = ctx->device;
        stager = std::thread([&, dev]() {
            try {
                cudaSetDevice(dev);
                for (int ci = 0; ci < n_chunks; ci++) {
                    if (ci >= 2) {   // bounce[ci & 1] still feeds the copy of chunk ci - 2
                        std::unique_lock<std::mutex> l(ss.m);
                        ss.cv.wait(l, [&] { return ss.abort || ss.issued[(size_t)ci - 2]; });
                        if (ss.abort) return;
                        l.unlock();
                        if (cudaEventSynchronize(bounce[ci & 1].h2d_done) != cudaSuccess) throw CudaFail{cudaGetLastError(), "staging: cudaEventSynchronize"};
                    }
                    {
                        std::lock_guard<std::mutex> l(ss.m);
                        if (ss.abort) return;
                    }
                    fill_chunk(ci);
                    {
                        std::lock_guard<std::mutex> l(ss.m);
                        ss.filled[(size_t)ci] = 1;
                    }
                    ss.cv.notify_all();
                }
            } catch (...) {
                std::lock_guard<std::mutex> l(ss.m);
                ss.err = std::current_exception();
                ss.cv.notify_all();
            }
        });
    }

    auto stage_chunk = [&](int ci) {
        InSet &x = sets[ci & 1];
        const int64_t g0 = cut[ci], g1 = cut[ci + 1];
        const int64_t r0 = group_offsets[g0], r1 = group_offsets[g1];
        const int64_t b0 = gbyte(g0), b1 = gbyte(g1);
        const uint8_t *src_data = nullptr, *src_offs = nullptr, *src_goff = nullptr, *src_valid = nullptr;
        if (direct) {
            src_data = c->data[0] + b0;
            src_offs = (const uint8_t *)c->offsets_or_views + (size_t)r0 * osz;
            src_goff = (const uint8_t *)(group_offsets + g0);
        } else {
            std::unique_lock<std::mutex> l(ss.m);
            ss.cv.wait(l, [&] { return ss.err || ss.filled[(size_t)ci]; });
            if (ss.err) std::rethrow_exception(ss.err);
            l.unlock();
            Bounce &bb = bounce[ci & 1];
            src_data = bb.data; src_offs = bb.offs; src_goff = bb.goff; src_valid = bb.valid;
        }
        CK(cudaStreamWaitEvent(s_in, x.consumed, 0));   // the compute that last read this set has finished
        const size_t pad = (size_t)(b0 & 15);           // keep the 16-byte phase of the bytes so that aligned chunks stay inside
        x.data.alloc(pad + (size_t)(b1 - b0) + 32, s_in);
        x.offs.alloc((size_t)(r1 - r0 + 1) * osz, s_in);
        x.goff.alloc((size_t)(g1 - g0 + 1) * 8, s_in);
        hmark();
        tmark(s_in);
        if (b1 > b0) CK(cudaMemcpyAsync(x.data.as<uint8_t>() + pad, src_data, (size_t)(b1 - b0), cudaMemcpyHostToDevice, s_in));
        CK(cudaMemcpyAsync(x.offs.p, src_offs, (size_t)(r1 - r0 + 1) * osz, cudaMemcpyHostToDevice, s_in));
        CK(cudaMemcpyAsync(x.goff.p, src_goff, (size_t)(g1 - g0 + 1) * 8, cudaMemcpyHostToDevice, s_in));
        Staged &s = x.st;
        s = Staged();
        s.col.n_rows = r1 - r0;
        s.col.row_begin = r0;
        s.col.off64 = o64;
        // "virtual base" pointers: the kernels keep indexing with absolute rows / byte offsets
        s.col.offsets = reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(x.offs.p) - (uintptr_t)r0 * osz);
        s.col.data = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(x.data.p) + pad - (uintptr_t)b0);
        s.data_begin = x.data.as<uint8_t>();
        s.data_end = x.data.as<uint8_t>() + pad + (size_t)(b1 - b0);
        s.group_offsets = x.goff.as<int64_t>();
        s.col.validity = nullptr;
        s.col.validity_bit_offset = c->validity_bit_offset;
        if (c->validity) {
            const int64_t bit0 = r0 + c->validity_bit_offset, bit1 = r1 + c->validity_bit_offset;
            const int64_t by0 = bit0 >> 3, by1 = (bit1 + 7) >> 3;
            x.valid.alloc((size_t)(by1 - by0) + 8, s_in);
            if (by1 > by0) CK(cudaMemcpyAsync(x.valid.p, direct ? c->validity + by0 : src_valid, (size_t)(by1 - by0), cudaMemcpyHostToDevice, s_in));
            s.col.validity = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(x.valid.p) - (uintptr_t)by0);
        }
        CK(cudaEventRecord(x.ready, s_in));
        if (!direct) {
            CK(cudaEventRecord(bounce[ci & 1].h2d_done, s_in));
            {
                std::lock_guard<std::mutex> l(ss.m);
                ss.issued[(size_t)ci] = 1;
            }
            ss.cv.notify_all();
        }
        tmark(s_in);
    };

    HostArr h_off, h_lo, h_hi, h_exts, h_counts, h_stats, h_rowptr, h_cols;
    struct HostGuard {   // an error on the way out hands the page-locked blocks back to the context's cache
        rogtk_gpu_ctx *ctx;
        HostArr *arr[8];
        cudaStream_t drain;
        bool armed = true;
        ~HostGuard()
        {
            if (!armed) return;
            cudaStreamSynchronize(drain);   // no copy may still be writing into a block that goes back to the cache
            for (HostArr *h : arr)
                if (h->p) pinned_put(ctx, h->p);
        }
    } host_guard{ctx, {&h_off, &h_lo, &h_hi, &h_exts, &h_counts, &h_stats, &h_rowptr, &h_cols}, s_out};
    struct PostGuard {
        rogtk_gpu_ctx *ctx;
        ChunkPost *post;
        cudaStream_t drain;
        bool armed = true;
        ~PostGuard()
        {
            if (!armed || !post) return;
            cudaStreamSynchronize(drain);
            for (HostArr *h : {&post->h_soff, &post->h_sdata, &post->h_sflags})
                if (h->p) {
                    pinned_put(ctx, h->p);
                    h->p = nullptr;
                }
        }
    } post_guard{ctx, post, s_out};
    if (post) {
        host_reserve(ctx, post->h_soff, (size_t)(n_groups + 1) * 8, 0, s_out);
        host_reserve(ctx, post->h_sflags, (size_t)n_groups + 16, 0, s_out);
        post->n_bytes = 0;
    }
    host_reserve(ctx, h_off, (size_t)(n_groups + 1) * 8, 0, s_out);
    if (flags & ROGTK_EMIT_STATS) host_reserve(ctx, h_stats, (size_t)n_groups * 12 + 16, 0, s_out);
    int64_t base_k = 0, base_e = 0, n_obs = 0, n_spill = 0, n_multi = 0;
    float ms_total = 0, ms_count = 0, ms_hot = 0, ms_sort = 0, ms_csr = 0, ms_stats = 0;
    const int64_t total_bytes = std::max<int64_t>(1, gbyte(n_groups) - gbyte(0));
    std::vector<DeviceResult> inflight(2);

    stage_chunk(0);
    for (int ci = 0; ci < n_chunks; ci++) {
        if (ci + 1 < n_chunks) stage_chunk(ci + 1);
        InSet &x = sets[ci & 1];
        const int64_t g0 = cut[ci], g1 = cut[ci + 1], ng = g1 - g0;
        CK(cudaStreamWaitEvent(s_comp, x.ready, 0));
        DeviceResult &r = inflight[ci & 1];
        // the previous result held in this slot has been copied out: its buffers are released in s_out order below
        r = DeviceResult();
        tmark(s_comp);
        run_pipeline(ctx, x.st, ng, K, min_coverage, flags, r);
        // shift the chunk-local offsets to their global base (in place, on the compute stream)
        if (base_k) k_add_i64<<<blocks_for(ng + 1, 256), 256, 0, s_comp>>>(r.off.as<int64_t>(), ng + 1, base_k);
        if ((flags & ROGTK_EMIT_CSR) && base_e)
            k_add_i64<<<blocks_for(r.n_kmers + 1, 256), 256, 0, s_comp>>>(r.rowptr.as<int64_t>(), r.n_kmers + 1, base_e);
        const int64_t done_bytes = std::max<int64_t>(1, gbyte(g1) - gbyte(0));
        const double scale = (double)total_bytes / (double)done_bytes;
        if (post) {
            // Compression on the chunk's device-resident nodes; only the strings go back
            rogtk_graph_batch view{};
            view.n_groups = ng; view.n_kmers = r.n_kmers; view.k_eff = K; view.on_device = 1;
            view.kmer_offsets = r.off.as<int64_t>(); view.kmers_lo = r.lo.as<uint64_t>(); view.kmers_hi = K > 32 ? r.hi.as<uint64_t>() : nullptr;
            view.exts = r.exts.as<uint8_t>(); view.counts = r.counts.as<uint16_t>();
            if (base_k) k_add_i64<<<blocks_for(ng + 1, 256), 256, 0, s_comp>>>(r.off.as<int64_t>(), ng + 1, -base_k);   // chunk-local again
            rogtk_contig_batch cb{};
            BatchOwner cown;
            cown.on_device = true;
            DevBuf roff, rows, rflags;
            int64_t nb = 0;
            try {
                compress_unitigs_impl(ctx, &view, post->min_length, ROGTK_OUTPUT_ON_DEVICE, &cb, &cown);
                contig_rows_device(ctx, cb, post->min_length, post->only_largest, roff, rows, rflags, nb);
            } catch (...) {
                for (void *p : cown.ptrs) dev_free(p, s_comp);
                throw;
            }
            for (void *p : cown.ptrs) dev_free(p, s_comp);
            if (post->n_bytes) k_add_i64<<<blocks_for(ng + 1, 256), 256, 0, s_comp>>>(roff.as<int64_t>(), ng + 1, post->n_bytes);
            CK(cudaEventRecord(comp_done, s_comp));
            CK(cudaEventRecord(x.consumed, s_comp));
            tmark(s_comp);
            const size_t need_b = (size_t)(post->n_bytes + nb);
            const size_t est_b = (size_t)((double)need_b * scale * 1.05) + 4096;
            host_reserve(ctx, post->h_sdata, std::max(need_b, ci == 0 ? est_b : need_b) + 16, (size_t)post->n_bytes, s_out);
            CK(cudaStreamWaitEvent(s_out, comp_done, 0));
            tmark(s_out);
            CK(cudaMemcpyAsync(post->h_soff.p + (size_t)g0 * 8, roff.p, (size_t)(ng + 1) * 8, cudaMemcpyDeviceToHost, s_out));
            if (nb) CK(cudaMemcpyAsync(post->h_sdata.p + (size_t)post->n_bytes, rows.p, (size_t)nb, cudaMemcpyDeviceToHost, s_out));
            if (ng) CK(cudaMemcpyAsync(post->h_sflags.p + (size_t)g0, rflags.p, (size_t)ng, cudaMemcpyDeviceToHost, s_out));
            roff.st = rows.st = rflags.st = s_out;
            roff.release(); rows.release(); rflags.release();
            tmark(s_out);
            r.off.release(); r.lo.release(); r.hi.release(); r.exts.release(); r.counts.release();
            r.stats.release(); r.rowptr.release(); r.cols.release();
            CK(cudaEventSynchronize(ctx->ev[13]));
            float msu = 0;
            CK(cudaEventElapsedTime(&msu, ctx->ev[12], ctx->ev[13]));
            post->ms_unitig += msu;
            post->n_bytes += nb;
        } else {
        CK(cudaEventRecord(comp_done, s_comp));
        CK(cudaEventRecord(x.consumed, s_comp));
        tmark(s_comp);
        // capacity of the final host arrays: after the first chunk extrapolate from bytes processed
        const size_t est_k = (size_t)((double)(base_k + r.n_kmers) * scale * 1.05) + 4096;
        const size_t est_e = (size_t)((double)(base_e + r.n_edges) * scale * 1.05) + 4096;
        const size_t need_k = (size_t)(base_k + r.n_kmers);
        host_reserve(ctx, h_lo, std::max(need_k, ci == 0 ? est_k : need_k) * 8, (size_t)base_k * 8, s_out);
        if (K > 32) host_reserve(ctx, h_hi, std::max(need_k, ci == 0 ? est_k : need_k) * 8, (size_t)base_k * 8, s_out);
        host_reserve(ctx, h_exts, std::max(need_k, ci == 0 ? est_k : need_k), (size_t)base_k, s_out);
        host_reserve(ctx, h_counts, std::max(need_k, ci == 0 ? est_k : need_k) * 2, (size_t)base_k * 2, s_out);
        if (flags & ROGTK_EMIT_CSR) {
            host_reserve(ctx, h_rowptr, (std::max(need_k, ci == 0 ? est_k : need_k) + 1) * 8, (size_t)(base_k + 1) * 8, s_out);
            const size_t need_e = (size_t)(base_e + r.n_edges);
            host_reserve(ctx, h_cols, std::max(need_e, ci == 0 ? est_e : need_e) * 4 + 16, (size_t)base_e * 4, s_out);
        }
        // D2H on its own stream
        CK(cudaStreamWaitEvent(s_out, comp_done, 0));
        tmark(s_out);
        auto pull = [&](HostArr &h, size_t dst_off, DevBuf &d, size_t bytes) {
            if (bytes) CK(cudaMemcpyAsync(h.p + dst_off, d.p, bytes, cudaMemcpyDeviceToHost, s_out));
            d.st = s_out;   // released after the copy, in s_out order
        };
        pull(h_off, (size_t)g0 * 8, r.off, (size_t)(ng + 1) * 8);
        pull(h_lo, (size_t)base_k * 8, r.lo, (size_t)r.n_kmers * 8);
        if (K > 32) pull(h_hi, (size_t)base_k * 8, r.hi, (size_t)r.n_kmers * 8);
        pull(h_exts, (size_t)base_k, r.exts, (size_t)r.n_kmers);
        pull(h_counts, (size_t)base_k * 2, r.counts, (size_t)r.n_kmers * 2);
        if (flags & ROGTK_EMIT_STATS) pull(h_stats, (size_t)g0 * 12, r.stats, (size_t)ng * 12);
        if (flags & ROGTK_EMIT_CSR) {
            pull(h_rowptr, (size_t)base_k * 8, r.rowptr, (size_t)(r.n_kmers + 1) * 8);
            pull(h_cols, (size_t)base_e * 4, r.cols, (size_t)r.n_edges * 4);
        }
        tmark(s_out);
        r.off.release(); r.lo.release(); r.hi.release(); r.exts.release(); r.counts.release();
        r.stats.release(); r.rowptr.release(); r.cols.release();
        }
        // timings of this chunk (waits for its compute; the next chunk's input copy is already in flight)
        CK(cudaEventSynchronize(ctx->ev[3]));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
        ms_total += ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
        ms_count += ms;
        ms_hot += r.ms_hot;
        CK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
        ms_sort += ms;
        if (r.has_stats) {
            CK(cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]));
            ms_stats += ms;
        }
        if (r.has_csr) {
            CK(cudaEventElapsedTime(&ms, ctx->ev[10], ctx->ev[11]));
            ms_csr += ms;
        }
        base_k += r.n_kmers;
        base_e += r.n_edges;
        n_obs += r.n_obs;
        n_spill += r.n_spill;
        n_multi += r.n_multi;
    }
    CK(cudaStreamSynchronize(s_out));
    CK(cudaStreamSynchronize(s_in));
    if (trace) {
        // event order: chunk 0 H2D (2), then per chunk ci: [H2D of ci+1 (2) if any], count (2), D2H (2)
        auto at = [&](size_t i) { float ms = 0; cudaEventElapsedTime(&ms, tev[0], tev[i]); return ms; };
        size_t i = 2, hi = 1;
        std::vector<float> h0(n_chunks), h1(n_chunks);
        h0[0] = at(0); h1[0] = at(1);
        for (int ci = 0; ci < n_chunks; ci++) {
            if (ci + 1 < n_chunks) { h0[ci + 1] = at(i); h1[ci + 1] = at(i + 1); i += 2; hi++; }
            fprintf(stderr, "chunk %3d  h2d %8.2f-%8.2f  count %8.2f-%8.2f  d2h %8.2f-%8.2f  staged@host %8.2f\n", ci, h0[ci], h1[ci],
                    at(i), at(i + 1), at(i + 2), at(i + 3), host_ms[ci]);
            i += 4;
        }
        for (auto e : tev) cudaEventDestroy(e);
    }
    for (auto &x : sets) { x.data.release(); x.offs.release(); x.goff.release(); x.valid.release(); }
    out->n_kmers = base_k;
    out->n_edges = base_e;
    out->n_observations = n_obs;
    out->n_spilled_groups = n_spill;
    out->n_multipass_groups = (int32_t)n_multi;
    auto adopt = [&](HostArr &h) -> void * {
        if (h.p) own->ptrs.push_back(h.p);
        return h.p;
    };
    out->kmer_offsets = (int64_t *)adopt(h_off);
    out->kmers_lo = (uint64_t *)adopt(h_lo);
    out->kmers_hi = K > 32 ? (uint64_t *)adopt(h_hi) : nullptr;
    out->exts = (uint8_t *)adopt(h_exts);
    out->counts = (uint16_t *)adopt(h_counts);
    if (flags & ROGTK_EMIT_STATS) out->stats = (uint32_t *)adopt(h_stats);
    if (flags & ROGTK_EMIT_CSR) {
        out->csr_row_ptr = (int64_t *)adopt(h_rowptr);
        out->csr_col = (int32_t *)adopt(h_cols);
    }
    out->ms_total = ms_total;
    out->ms_count = ms_count;
    out->ms_hot_kernel = ms_hot;
    out->ms_sort = ms_sort;
    out->ms_csr = ms_csr;
    out->ms_stats = ms_stats;
    host_guard.armed = false;   // the blocks now belong to the batch
    post_guard.armed = false;
    CK(cudaGetLastError());
}

// assemble_sequences(method = compression) on the overlapped host path: true when it ran (large host-resident input), false when the
// caller should take the plain route.  `post` receives the strings in page-locked blocks the caller adopts.
bool assemble_compression_chunked(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups, int64_t k,
                                  uint32_t min_coverage, uint32_t flags, ChunkPost &post, rogtk_graph_batch *info)
{
    if ((flags & (ROGTK_INPUT_ON_DEVICE | ROGTK_OUTPUT_ON_DEVICE)) || n_groups <= 1 || !reads || !group_offsets) return false;
    int64_t total = 0;
    if (reads->layout == ROGTK_UTF8_VIEW) {
        for (int b = 0; b < reads->n_buffers && reads->data_len; b++) total += std::max<int64_t>(0, reads->data_len[b]);
        total = std::max<int64_t>(total, reads->n_rows * 12);
    } else if (reads->n_buffers == 1 && reads->data && reads->data_len) {
        total = reads->data_len[0];
    }
    if (total <= (int64_t)ctx->chunk_bytes + (int64_t)(ctx->chunk_bytes >> 1)) return false;
    const int K = (flags & ROGTK_EXACT_K) ? (int)k : k_eff_host(k);
    if (K < 1 || K > 64 || !reads->offsets_or_views || reads->n_rows < 0) return false;
    {   // the host indexes with the group offsets: same check as count_prune_impl
        int64_t prev = group_offsets[0];
        bool ok = prev >= 0;
        for (int64_t g = 1; g <= n_groups && ok; g++) {
            ok = group_offsets[g] >= prev;
            prev = group_offsets[g];
        }
        if (!ok || prev > reads->n_rows) throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: group_offsets not monotone or outside the rows;"};
    }
    CK(cudaSetDevice(ctx->device));
    memset(info, 0, sizeof(*info));
    BatchOwner own;   // the node arrays never reach the host on this route: nothing to adopt but the offsets block
    run_chunked(ctx, reads, group_offsets, n_groups, K, min_coverage, flags & ~(ROGTK_EMIT_CSR | ROGTK_EMIT_STATS), info, &own, &post);
    for (void *p : own.ptrs) pinned_put(ctx, p);
    info->k_eff = K;
    return true;
}

int count_prune_impl(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets, int64_t n_groups, int64_t k,
                     uint32_t min_coverage, uint32_t flags, rogtk_graph_batch *out)
{
    if (!reads || !out || n_groups < 0 || (n_groups > 0 && !group_offsets) || reads->n_rows < 0)
        throw ArgFail{ROGTK_ERR_INVALID_ARG, "null or negative argument"};
    if (n_groups >= (int64_t)1 << 31) throw ArgFail{ROGTK_ERR_UNSUPPORTED, "more than 2^31 groups per batch"};
    if (n_groups > 0 && !reads->offsets_or_views) throw ArgFail{ROGTK_ERR_INVALID_ARG, "null offsets / views buffer"};
    if (reads->layout != ROGTK_UTF8_VIEW && reads->n_rows > 0 && (!reads->data || !reads->data_len || reads->n_buffers != 1))
        throw ArgFail{ROGTK_ERR_INVALID_ARG, "offset layouts need exactly one data buffer"};
    memset(out, 0, sizeof(*out));
    const bool in_dev = flags & ROGTK_INPUT_ON_DEVICE, out_dev = flags & ROGTK_OUTPUT_ON_DEVICE;
    int K;
    if (flags & ROGTK_EXACT_K) {
        if (k < 1 || k > 64) throw ArgFail{ROGTK_ERR_INVALID_ARG, "EXACT_K needs 1 <= k <= 64"};
        K = (int)k;
    } else {
        K = k_eff_host(k);
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    BatchOwner *own = new BatchOwner();
    own->on_device = out_dev;
    out->_owner = own;
    out->n_groups = n_groups;
    out->k_eff = K;
    out->on_device = out_dev ? 1 : 0;
    if (K < 0) {
        fill_empty_batch(ctx, n_groups, out_dev, flags, out, own);
        return ROGTK_OK;
    }

    if (!in_dev && n_groups > 0) {   // host-visible group offsets: reject nonsense before the host itself indexes with them
        int64_t prev = group_offsets[0];
        bool ok = prev >= 0;
        for (int64_t g = 1; g <= n_groups && ok; g++) {
            ok = group_offsets[g] >= prev;
            prev = group_offsets[g];
        }
        if (!ok || prev > reads->n_rows) throw ArgFail{ROGTK_ERR_INVALID_ARG, "malformed input: group_offsets not monotone or outside the rows;"};
    }
    // host input + host output of a large column: chunked, copies overlapped with compute
    if (!in_dev && !out_dev && n_groups > 1) {
        int64_t total = 0;
        if (reads->layout == ROGTK_UTF8_VIEW) {
            for (int b = 0; b < reads->n_buffers && reads->data_len; b++) total += std::max<int64_t>(0, reads->data_len[b]);
            total = std::max<int64_t>(total, reads->n_rows * 12);   // (short rows live inside the views)
        } else if (reads->n_buffers == 1 && reads->data && reads->data_len) {
            total = reads->data_len[0];
        }
        if (total > (int64_t)ctx->chunk_bytes + (int64_t)(ctx->chunk_bytes >> 1)) {
            run_chunked(ctx, reads, group_offsets, n_groups, K, min_coverage, flags, out, own);
            return ROGTK_OK;
        }
    }

    Staged in;
    stage_input(ctx, reads, group_offsets, n_groups, in_dev, in);
    DeviceResult r;
    run_pipeline(ctx, in, n_groups, K, min_coverage, flags, r);

    // ---- hand over ------------------------------------------------------------------------------------------------------
    out->n_kmers = r.n_kmers;
    out->n_edges = r.n_edges;
    out->n_observations = r.n_obs;
    out->n_spilled_groups = r.n_spill;
    out->n_multipass_groups = (int32_t)r.n_multi;
    auto give = [&](DevBuf &b, size_t bytes) -> void * {
        if (!b.p) return nullptr;
        if (out_dev) {
            void *p = b.detach();
            own->ptrs.push_back(p);
            return p;
        }
        void *h = pinned_get(ctx, bytes);
        own->ptrs.push_back(h);
        if (bytes) CK(cudaMemcpyAsync(h, b.p, bytes, cudaMemcpyDeviceToHost, st));
        return h;
    };
    out->kmer_offsets = (int64_t *)give(r.off, (size_t)(n_groups + 1) * 8);
    out->kmers_lo = (uint64_t *)give(r.lo, (size_t)r.n_kmers * 8);
    out->kmers_hi = K > 32 ? (uint64_t *)give(r.hi, (size_t)r.n_kmers * 8) : nullptr;
    out->exts = (uint8_t *)give(r.exts, (size_t)r.n_kmers);
    out->counts = (uint16_t *)give(r.counts, (size_t)r.n_kmers * 2);
    if (flags & ROGTK_EMIT_STATS) out->stats = (uint32_t *)give(r.stats, (size_t)n_groups * 12);
    if (flags & ROGTK_EMIT_CSR) {
        out->csr_row_ptr = (int64_t *)give(r.rowptr, (size_t)(r.n_kmers + 1) * 8);
        out->csr_col = (int32_t *)give(r.cols, (size_t)r.n_edges * 4);
    }
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
    out->ms_total = ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
    out->ms_count = ms;
    out->ms_hot_kernel = r.ms_hot;
    CK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
    out->ms_sort = ms;
    if (r.has_stats) {
        CK(cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]));
        out->ms_stats = ms;
    }
    if (r.has_csr) {
        CK(cudaEventElapsedTime(&ms, ctx->ev[10], ctx->ev[11]));
        out->ms_csr = ms;
    }
    CK(cudaGetLastError());
    return ROGTK_OK;
}

}  // namespace

extern "C" void rogtk_gpu_free_batch(rogtk_gpu_ctx *ctx, rogtk_graph_batch *b)
{
    if (!ctx || !b || !b->_owner) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    BatchOwner *own = reinterpret_cast<BatchOwner *>(b->_owner);
    cudaSetDevice(ctx->device);
    for (void *p : own->ptrs) {
        if (own->on_device) dev_free(p, ctx->stream);
        else pinned_put(ctx, p);
    }
    delete own;
    memset(b, 0, sizeof(*b));
}

extern "C" int rogtk_gpu_count_prune_batch(rogtk_gpu_ctx *ctx, const rogtk_strcol *reads, const int64_t *group_offsets,
                                           int64_t n_groups, int64_t k, uint32_t min_coverage, uint32_t flags,
                                           rogtk_graph_batch *out)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    int rc = ROGTK_ERR_INTERNAL;
    try {
        rc = count_prune_impl(ctx, reads, group_offsets, n_groups, k, min_coverage, flags, out);
    } catch (...) {
        rc = fail_code(ctx);
    }
    if (rc != ROGTK_OK && out && out->_owner) {
        BatchOwner *own = reinterpret_cast<BatchOwner *>(out->_owner);
        for (void *p : own->ptrs) {
            if (own->on_device) dev_free(p, ctx->stream);
            else pinned_put(ctx, p);
        }
        delete own;
        memset(out, 0, sizeof(*out));
    }
    return rc;
}

// -----------------------------------------------------------------------------------------------------------------
// threshold re-use
// -----------------------------------------------------------------------------------------------------------------
extern "C" int rogtk_gpu_rethreshold(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *in, uint32_t min_coverage2, uint32_t flags,
                                     rogtk_graph_batch *out)
{
    if (!ctx || !in || !out || !in->on_device || in->n_groups < 0 || in->n_kmers < 0 || !in->kmer_offsets ||
        (in->n_kmers > 0 && (!in->kmers_lo || !in->exts || !in->counts)) || (in->k_eff > 32 && in->n_kmers > 0 && !in->kmers_hi))
        return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    memset(out, 0, sizeof(*out));
    BatchOwner *own = new BatchOwner();
    const bool out_dev = flags & ROGTK_OUTPUT_ON_DEVICE;
    own->on_device = out_dev;
    out->_owner = own;
    int rc = ROGTK_OK;
    try {
        CK(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        CK(cudaEventRecord(ctx->ev[0], st));
        const int64_t n = in->n_kmers, ng = in->n_groups;
        const int K = in->k_eff;
        const uint32_t thr = min_coverage2 > 1u ? min_coverage2 : 1u;
        const size_t nn = (size_t)std::max<int64_t>(n, 1);
        DevBuf keep(nn * 4, st), pos((nn + 1) * 8, st), scr(scan_scratch_bytes(n), st), n_off((size_t)(ng + 1) * 8, st);
        if (n > 0) k_thr_flags<<<blocks_for(n, 256), 256, 0, st>>>(in->counts, n, thr, keep.as<int32_t>());
        exclusive_scan<int32_t>(keep.as<int32_t>(), n, pos.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
        int64_t n_out = 0;
        CK(cudaMemcpyAsync(&n_out, pos.as<int64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
        k_thr_offsets<<<blocks_for(ng + 1, 256), 256, 0, st>>>(in->kmer_offsets, ng, pos.as<int64_t>(), n_off.as<int64_t>());
        CK(cudaStreamSynchronize(st));
        const size_t no = (size_t)std::max<int64_t>(n_out, 1);
        DevBuf o_lo(no * 8, st), o_hi(K > 32 ? no * 8 : 0, st), o_exts(no, st), o_counts(no * 2, st), d_stats, d_rowptr, d_cols;
        uint64_t *p_hi = K > 32 ? o_hi.as<uint64_t>() : nullptr;
        if (n > 0)
            k_thr_compact<<<blocks_for(n, 256), 256, 0, st>>>(n, keep.as<int32_t>(), pos.as<int64_t>(), in->kmers_lo, in->kmers_hi, in->exts,
                                                              in->counts, o_lo.as<uint64_t>(), p_hi, o_exts.as<uint8_t>(),
                                                              o_counts.as<uint16_t>());
        ctx->launches += 3;
        int64_t n_edges = 0;
        if (flags & ROGTK_EMIT_STATS) {
            d_stats.alloc((size_t)ng * 12 + 16, st);
            if (ng > 0) k_stats<<<ctx->sm_count * 8, 256, 0, st>>>(ng, n_off.as<int64_t>(), o_exts.as<uint8_t>(), d_stats.as<uint32_t>());
        }
        if (flags & ROGTK_EMIT_CSR) {
            emit_csr(ctx, ng, n_out, n_off.as<int64_t>(), o_lo.as<uint64_t>(), p_hi, o_exts.as<uint8_t>(), K, d_rowptr, d_cols, n_edges);
        }
        CK(cudaEventRecord(ctx->ev[3], st));
        out->n_groups = ng;
        out->n_kmers = n_out;
        out->n_edges = n_edges;
        out->n_observations = in->n_observations;
        out->k_eff = K;
        out->on_device = out_dev ? 1 : 0;
        auto give = [&](DevBuf &b, size_t bytes) -> void * {
            if (!b.p) return nullptr;
            if (out_dev) {
                void *p = b.detach();
                own->ptrs.push_back(p);
                return p;
            }
            void *h = pinned_get(ctx, bytes);
            own->ptrs.push_back(h);
            if (bytes) CK(cudaMemcpyAsync(h, b.p, bytes, cudaMemcpyDeviceToHost, st));
            return h;
        };
        out->kmer_offsets = (int64_t *)give(n_off, (size_t)(ng + 1) * 8);
        out->kmers_lo = (uint64_t *)give(o_lo, (size_t)n_out * 8);
        out->kmers_hi = K > 32 ? (uint64_t *)give(o_hi, (size_t)n_out * 8) : nullptr;
        out->exts = (uint8_t *)give(o_exts, (size_t)n_out);
        out->counts = (uint16_t *)give(o_counts, (size_t)n_out * 2);
        if (flags & ROGTK_EMIT_STATS) out->stats = (uint32_t *)give(d_stats, (size_t)ng * 12);
        if (flags & ROGTK_EMIT_CSR) {
            out->csr_row_ptr = (int64_t *)give(d_rowptr, (size_t)(n_out + 1) * 8);
            out->csr_col = (int32_t *)give(d_cols, (size_t)n_edges * 4);
        }
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
        out->ms_total = ms;
        CK(cudaGetLastError());
    } catch (...) {
        rc = fail_code(ctx);
    }
    if (rc != ROGTK_OK) {
        for (void *p : own->ptrs) {
            if (own->on_device) dev_free(p, ctx->stream);
            else pinned_put(ctx, p);
        }
        delete own;
        memset(out, 0, sizeof(*out));
    }
    return rc;
}

// -----------------------------------------------------------------------------------------------------------------
// global k-mer spectrum
// -----------------------------------------------------------------------------------------------------------------
extern "C" int rogtk_gpu_spectrum_owner(uint64_t kmer_lo, uint64_t kmer_hi, int k_eff, int n_ranks)
{
    if (k_eff < 1 || k_eff > 64 || n_ranks < 1 || n_ranks > 256) return -1;
    return spectrum_owner(kmer_lo, kmer_hi, 2 * k_eff, n_ranks);
}

extern "C" int rogtk_gpu_spectrum_split(rogtk_gpu_ctx *ctx, const uint64_t *d_lo, const uint64_t *d_hi, int64_t n, int k_eff,
                                        int n_ranks, int64_t *h_bounds)
{
    if (!ctx || !h_bounds || n < 0 || k_eff < 1 || k_eff > 64 || n_ranks < 1 || n_ranks > 256 || (n > 0 && !d_lo))
        return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    try {
        CK(cudaSetDevice(ctx->device));
        DevBuf b((size_t)(n_ranks + 1) * 8, ctx->stream);
        k_spectrum_bounds<<<blocks_for(n_ranks + 1, 128), 128, 0, ctx->stream>>>(d_lo, k_eff > 32 ? d_hi : nullptr, n, 2 * k_eff, n_ranks,
                                                                               b.as<int64_t>());
        ctx->launches++;
        CK(cudaMemcpyAsync(h_bounds, b.p, (size_t)(n_ranks + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        return fail_code(ctx);
    }
    return ROGTK_OK;
}

// What the merge knows when its input is the exchange's: R ascending runs, all inside ranges [base, base + n_ranges) of `total`
// equal ranges of the k-mer space (this rank's share).  Then the merge is one shared-memory table per range (kernels_part.cuh).
struct MergeRuns {
    std::vector<int64_t> run_off;   // [R + 1]
    uint32_t total = 0, base = 0;
    int n_ranges = 0;
};

static int spectrum_merge_impl(rogtk_gpu_ctx *ctx, const uint64_t *d_lo, const uint64_t *d_hi, const uint16_t *d_counts, const uint8_t *d_exts,
                               int64_t n, int k_eff, uint32_t min_coverage, uint32_t flags, rogtk_graph_batch *out, const MergeRuns *runs)
{
    if (!ctx || !out || n < 0 || k_eff < 1 || k_eff > 64 || (n > 0 && (!d_lo || !d_counts || !d_exts)) || (k_eff > 32 && n > 0 && !d_hi))
        return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    memset(out, 0, sizeof(*out));
    BatchOwner *own = new BatchOwner();
    const bool out_dev = flags & ROGTK_OUTPUT_ON_DEVICE;
    own->on_device = out_dev;
    out->_owner = own;
    int rc = ROGTK_OK;
    try {
        CK(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        CK(cudaEventRecord(ctx->ev[0], st));
        const size_t nn = (size_t)std::max<int64_t>(n, 1);
        int64_t n_out = 0;
        DevBuf o_lo, o_hi, o_exts, o_counts, o_off(16, st), o_stats(16, st);
        bool merged = false;
        if (runs && n > 0 && runs->n_ranges > 0 && !(flags & ROGTK_FORCE_SPILL)) {
            const int R = (int)runs->run_off.size() - 1;
            DevBuf d_ro((size_t)(R + 1) * 8, st);
            CK(cudaMemcpyAsync(d_ro.p, runs->run_off.data(), (size_t)(R + 1) * 8, cudaMemcpyHostToDevice, st));
            if (k_eff > 32)
                merged = merge_sorted_runs_w<2>(st, ctx->peek_slot, ctx->sm_count, d_lo, d_hi, d_counts, d_exts, d_ro.as<int64_t>(), R, n, k_eff, min_coverage,
                                                runs->total, runs->base, runs->n_ranges, o_lo, o_hi, o_exts, o_counts, n_out, &ctx->launches);
            else
                merged = merge_sorted_runs_w<1>(st, ctx->peek_slot, ctx->sm_count, d_lo, nullptr, d_counts, d_exts, d_ro.as<int64_t>(), R, n, k_eff, min_coverage,
                                                runs->total, runs->base, runs->n_ranges, o_lo, o_hi, o_exts, o_counts, n_out, &ctx->launches);
        }
        if (!merged) {
        DevBuf recs(nn * sizeof(SpecRec), st), keep(nn * 4, st), pos((nn + 1) * 8, st), scr(scan_scratch_bytes(n), st);
        if (n > 0) {
            k_spec_pack<<<blocks_for(n, 256), 256, 0, st>>>(d_lo, k_eff > 32 ? d_hi : nullptr, d_counts, d_exts, n, recs.as<SpecRec>());
            {
                DevBuf sorted(nn * sizeof(SpecRec), st);
                sort_spec_records(st, recs.as<SpecRec>(), sorted.as<SpecRec>(), n, k_eff);
                recs = std::move(sorted);
            }
            k_spec_reduce<<<blocks_for(n, 256), 256, 0, st>>>(recs.as<SpecRec>(), n, min_coverage, keep.as<int32_t>());
            exclusive_scan<int32_t>(keep.as<int32_t>(), n, pos.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
            CK(cudaMemcpyAsync(&n_out, pos.as<int64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            ctx->launches += 5;
        }
        const size_t no = (size_t)std::max<int64_t>(n_out, 1);
        o_lo.alloc(no * 8, st);
        if (k_eff > 32) o_hi.alloc(no * 8, st);
        o_exts.alloc(no, st);
        o_counts.alloc(no * 2, st);

        if (n > 0) {
            k_spec_emit<<<blocks_for(n, 256), 256, 0, st>>>(recs.as<SpecRec>(), n, keep.as<int32_t>(), pos.as<int64_t>(), o_lo.as<uint64_t>(),
                                                            k_eff > 32 ? o_hi.as<uint64_t>() : nullptr, o_exts.as<uint8_t>(),
                                                            o_counts.as<uint16_t>());
            ctx->launches++;
        }
        }
        const int64_t h_off[2] = {0, n_out};
        CK(cudaMemcpyAsync(o_off.p, h_off, 16, cudaMemcpyHostToDevice, st));
        if (flags & ROGTK_EMIT_STATS) {
            k_stats<<<1, 256, 0, st>>>(1, o_off.as<int64_t>(), o_exts.as<uint8_t>(), o_stats.as<uint32_t>());
            ctx->launches++;
        }
        CK(cudaEventRecord(ctx->ev[3], st));
        out->n_groups = 1;
        out->n_kmers = n_out;
        out->k_eff = k_eff;
        out->on_device = out_dev ? 1 : 0;
        auto give = [&](DevBuf &b, size_t bytes) -> void * {
            if (out_dev) {
                void *p = b.detach();
                own->ptrs.push_back(p);
                return p;
            }
            void *h = pinned_get(ctx, bytes);
            own->ptrs.push_back(h);
            if (bytes) CK(cudaMemcpyAsync(h, b.p, bytes, cudaMemcpyDeviceToHost, st));
            return h;
        };
        out->kmer_offsets = (int64_t *)give(o_off, 16);
        out->kmers_lo = (uint64_t *)give(o_lo, (size_t)n_out * 8);
        out->kmers_hi = k_eff > 32 ? (uint64_t *)give(o_hi, (size_t)n_out * 8) : nullptr;
        out->exts = (uint8_t *)give(o_exts, (size_t)n_out);
        out->counts = (uint16_t *)give(o_counts, (size_t)n_out * 2);
        if (flags & ROGTK_EMIT_STATS) out->stats = (uint32_t *)give(o_stats, 12);
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
        out->ms_total = ms;
        CK(cudaGetLastError());
    } catch (...) {
        rc = fail_code(ctx);
    }
    if (rc != ROGTK_OK) {
        for (void *p : own->ptrs) {
            if (own->on_device) dev_free(p, ctx->stream);
            else pinned_put(ctx, p);
        }
        delete own;
        memset(out, 0, sizeof(*out));
    }
    return rc;
}

extern "C" int rogtk_gpu_spectrum_merge(rogtk_gpu_ctx *ctx, const uint64_t *d_lo, const uint64_t *d_hi, const uint16_t *d_counts,
                                        const uint8_t *d_exts, int64_t n, int k_eff, uint32_t min_coverage, uint32_t flags,
                                        rogtk_graph_batch *out)
{
    return spectrum_merge_impl(ctx, d_lo, d_hi, d_counts, d_exts, n, k_eff, min_coverage, flags, out, nullptr);
}

// -----------------------------------------------------------------------------------------------------------------
// unitig compaction + contig selection
// -----------------------------------------------------------------------------------------------------------------
namespace {

constexpr int UG_SMALL_NMAX = 1536, UG_SMALL_NT = 256;   // 46 KB (8-byte keys): 4 CTAs per SM
constexpr int UG_BIG_NMAX = 6144, UG_BIG_NT = 1024;      // 184 KB: 1 CTA per SM

template <int W, int NMAX, int NT>
void unitig_fused_launch(rogtk_gpu_ctx *ctx, UnitigGroupArgs &a, bool emit, int64_t ng)
{
    const size_t smem = unitig_group_smem<W, NMAX>();
    // (per device, cheap: set on every call rather than tracked per context)
    CK(cudaFuncSetAttribute(k_unitig_group<W, NMAX, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_unitig_group<W, NMAX, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = std::max<int>(1, std::min<int>(2048 / NT, (int)((ctx->smem_optin + 1024) / (smem + 1024))));
    const unsigned grid = (unsigned)std::min<int64_t>(ng, (int64_t)ctx->sm_count * per_sm);
    if (emit) k_unitig_group<W, NMAX, NT, true><<<grid, NT, smem, ctx->stream>>>(a);
    else k_unitig_group<W, NMAX, NT, false><<<grid, NT, smem, ctx->stream>>>(a);
    ctx->launches++;
}

void unitig_fused(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *in, int64_t min_length, bool all_small, bool out_dev, BatchOwner *own,
                  rogtk_contig_batch *out)
{
    cudaStream_t st = ctx->stream;
    const int64_t ng = in->n_groups;
    const int K = in->k_eff;
    DevBuf g_nu((size_t)ng * 4, st), g_nb((size_t)ng * 8, st), uoff((size_t)(ng + 1) * 8, st), sbase((size_t)(ng + 1) * 8, st),
        scr(scan_scratch_bytes(ng), st);
    UnitigGroupArgs a{};
    a.n_groups = ng; a.offsets = in->kmer_offsets; a.lo = in->kmers_lo; a.hi = K > 32 ? in->kmers_hi : nullptr;
    a.exts = in->exts; a.counts = in->counts; a.K = K;
    a.g_nu = g_nu.as<int32_t>(); a.g_nb = g_nb.as<int64_t>();
    // groups of up to UG_SMALL_NMAX nodes (and the empty ones) in the small instantiation, the others in the large one
    auto launch = [&](bool emit) {
        a.n_min = 0;
        if (K > 32) unitig_fused_launch<2, UG_SMALL_NMAX, UG_SMALL_NT>(ctx, a, emit, ng);
        else unitig_fused_launch<1, UG_SMALL_NMAX, UG_SMALL_NT>(ctx, a, emit, ng);
        if (all_small) return;
        a.n_min = UG_SMALL_NMAX;
        if (K > 32) unitig_fused_launch<2, UG_BIG_NMAX, UG_BIG_NT>(ctx, a, emit, ng);
        else unitig_fused_launch<1, UG_BIG_NMAX, UG_BIG_NT>(ctx, a, emit, ng);
    };
    launch(false);
    exclusive_scan<int32_t>(g_nu.as<int32_t>(), ng, uoff.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
    exclusive_scan<int64_t>(g_nb.as<int64_t>(), ng, sbase.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
    int64_t nu = 0, nb = 0;
    Peek pk(ctx->peek_slot, st, &ctx->launches);
    pk.add(uoff.as<int64_t>() + ng, 8, &nu);
    pk.add(sbase.as<int64_t>() + ng, 8, &nb);
    pk.sync();
    const size_t nun = (size_t)std::max<int64_t>(nu, 1);
    DevBuf soff((nun + 1) * 8, st), seq((size_t)std::max<int64_t>(nb, 1) + 8, st), ulen(nun * 4, st), ucnt(nun * 2, st), uext(nun, st),
        ucyc(nun, st), largest((size_t)std::max<int64_t>(ng, 1) * 8, st), lflags((size_t)std::max<int64_t>(ng, 1), st);
    a.uoff = uoff.as<int64_t>(); a.sbase = sbase.as<int64_t>();
    a.seq_offsets = soff.as<int64_t>(); a.seq = seq.as<uint8_t>(); a.ulen = ulen.as<int32_t>(); a.ucnt = ucnt.as<uint16_t>();
    a.uext = uext.as<uint8_t>(); a.ucyc = ucyc.as<uint8_t>();
    launch(true);
    CK(cudaMemcpyAsync(soff.as<int64_t>() + nu, sbase.as<int64_t>() + ng, 8, cudaMemcpyDeviceToDevice, st));
    k_unitig_largest<<<blocks_for(ng, 256), 256, 0, st>>>(ng, uoff.as<int64_t>(), ulen.as<int32_t>(), ucyc.as<uint8_t>(), min_length,
                                                          largest.as<int64_t>(), lflags.as<uint8_t>());
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev[13], st));
    out->n_groups = ng;
    out->n_unitigs = nu;
    out->n_bases = nb;
    out->k_eff = K;
    out->on_device = out_dev ? 1 : 0;
    auto give = [&](DevBuf &b, size_t bytes) -> void * {
        if (out_dev) {
            void *p = b.detach();
            own->ptrs.push_back(p);
            return p;
        }
        void *h = pinned_get(ctx, bytes);
        own->ptrs.push_back(h);
        if (bytes) CK(cudaMemcpyAsync(h, b.p, bytes, cudaMemcpyDeviceToHost, st));
        return h;
    };
    out->unitig_offsets = (int64_t *)give(uoff, (size_t)(ng + 1) * 8);
    out->seq_offsets = (int64_t *)give(soff, (size_t)(nu + 1) * 8);
    out->seq = (uint8_t *)give(seq, (size_t)nb);
    out->counts = (uint16_t *)give(ucnt, (size_t)nu * 2);
    out->exts = (uint8_t *)give(uext, (size_t)nu);
    out->cyclic = (uint8_t *)give(ucyc, (size_t)nu);
    out->largest = (int64_t *)give(largest, (size_t)ng * 8);
    out->largest_flags = (uint8_t *)give(lflags, (size_t)ng);
}

}  // namespace

namespace {
// (no lock, no argument checks: callers hold ctx->mu.  Records ev[12] / ev[13] around the device work.)
void compress_unitigs_impl(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *in, int64_t min_length, uint32_t flags, rogtk_contig_batch *out,
                           BatchOwner *own)
{
    const bool out_dev = flags & ROGTK_OUTPUT_ON_DEVICE;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[12], st));
    const int64_t n = in->n_kmers, ng = in->n_groups;
    const int K = in->k_eff;
    const size_t nn = (size_t)std::max<int64_t>(n, 1);
    // every group small enough for one CTA's shared memory: fused per-group kernels (two passes), else node-parallel kernels
    unsigned long long max_n = 0;
    if (n > 0 && ng > 0 && !(flags & ROGTK_UNITIG_NODE_PARALLEL)) {
        DevBuf mx(8, st);
        CK(cudaMemsetAsync(mx.p, 0, 8, st));
        k_max_group<<<ctx->sm_count * 4, 256, 0, st>>>(in->kmer_offsets, ng, mx.as<unsigned long long>());
        Peek pk0(ctx->peek_slot, st, &ctx->launches);
        pk0.add(mx.p, 8, &max_n);
        pk0.sync();
        ctx->launches++;
    }
    const bool fused = max_n > 0 && max_n <= (unsigned long long)UG_BIG_NMAX;
    if (fused) {
        unitig_fused(ctx, in, min_length, max_n <= (unsigned long long)UG_SMALL_NMAX, out_dev, own, out);
    } else {
    DevBuf fx(nn + 4, st), nxt(nn * 4, st), fl(nn, st), hlen(nn * 4, st), hcnt(nn * 2, st), hext(nn, st), is_head(nn * 4, st),
        upos((nn + 1) * 8, st), scr(scan_scratch_bytes(n), st), uoff((size_t)(ng + 1) * 8, st);
    CK(cudaMemsetAsync(fl.p, 0, nn, st));
    CK(cudaMemsetAsync(fx.p, 0, nn + 4, st));
    UnitigArgs a{};
    a.n = n; a.n_groups = ng; a.offsets = in->kmer_offsets;
    a.lo = in->kmers_lo; a.hi = K > 32 ? in->kmers_hi : nullptr; a.exts = in->exts; a.counts = in->counts; a.K = K;
    a.fx = fx.as<uint8_t>(); a.nxt = nxt.as<int32_t>(); a.flags = fl.as<uint8_t>(); a.hlen = hlen.as<int32_t>();
    a.hcnt = hcnt.as<uint16_t>(); a.hext = hext.as<uint8_t>();
    if (n > 0) {
        const unsigned gb = blocks_for(n, 256);
        k_unitig_links<<<gb, 256, 0, st>>>(a);
        k_unitig_fixleft<<<gb, 256, 0, st>>>(a);
        k_unitig_next<<<gb, 256, 0, st>>>(a);
        k_unitig_heads<<<gb, 256, 0, st>>>(a);
        k_unitig_cycles<<<gb, 256, 0, st>>>(a);
        k_unitig_headflags<<<gb, 256, 0, st>>>(a.flags, n, is_head.as<int32_t>());
        ctx->launches += 6;
    }
    exclusive_scan<int32_t>(is_head.as<int32_t>(), n, upos.as<int64_t>(), scr.as<int64_t>(), st, &ctx->launches);
    k_thr_offsets<<<blocks_for(ng + 1, 256), 256, 0, st>>>(in->kmer_offsets, ng, upos.as<int64_t>(), uoff.as<int64_t>());
    int64_t nu = 0;
    Peek pk(ctx->peek_slot, st, &ctx->launches);
    pk.add(upos.as<int64_t>() + n, 8, &nu);
    pk.sync();
    const size_t nun = (size_t)std::max<int64_t>(nu, 1);
    DevBuf uhead(nun * 8, st), ulen(nun * 4, st), ucnt(nun * 2, st), uext(nun, st), ucyc(nun, st), soff((nun + 1) * 8, st),
        scr2(scan_scratch_bytes(nu), st), largest((size_t)std::max<int64_t>(ng, 1) * 8, st), lflags((size_t)std::max<int64_t>(ng, 1), st);
    if (n > 0)
        k_unitig_compact<<<blocks_for(n, 256), 256, 0, st>>>(a, upos.as<int64_t>(), uhead.as<int64_t>(), ulen.as<int32_t>(),
                                                             ucnt.as<uint16_t>(), uext.as<uint8_t>(), ucyc.as<uint8_t>());
    exclusive_scan<int32_t>(ulen.as<int32_t>(), nu, soff.as<int64_t>(), scr2.as<int64_t>(), st, &ctx->launches);
    int64_t nb = 0;
    pk.add(soff.as<int64_t>() + nu, 8, &nb);
    pk.sync();
    DevBuf seq((size_t)std::max<int64_t>(nb, 1) + 8, st);
    if (nu > 0)
        k_unitig_emit<<<blocks_for(nu, 128), 128, 0, st>>>(a, nu, uhead.as<int64_t>(), ulen.as<int32_t>(), soff.as<int64_t>(),
                                                           seq.as<uint8_t>());
    if (ng > 0)
        k_unitig_largest<<<blocks_for(ng, 256), 256, 0, st>>>(ng, uoff.as<int64_t>(), ulen.as<int32_t>(), ucyc.as<uint8_t>(),
                                                              min_length, largest.as<int64_t>(), lflags.as<uint8_t>());
    ctx->launches += 4;
    CK(cudaEventRecord(ctx->ev[13], st));
    out->n_groups = ng;
    out->n_unitigs = nu;
    out->n_bases = nb;
    out->k_eff = K;
    out->on_device = out_dev ? 1 : 0;
    auto give = [&](DevBuf &b, size_t bytes) -> void * {
        if (out_dev) {
            void *p = b.detach();
            own->ptrs.push_back(p);
            return p;
        }
        void *h = pinned_get(ctx, bytes);
        own->ptrs.push_back(h);
        if (bytes) CK(cudaMemcpyAsync(h, b.p, bytes, cudaMemcpyDeviceToHost, st));
        return h;
    };
    out->unitig_offsets = (int64_t *)give(uoff, (size_t)(ng + 1) * 8);
    out->seq_offsets = (int64_t *)give(soff, (size_t)(nu + 1) * 8);
    out->seq = (uint8_t *)give(seq, (size_t)nb);
    out->counts = (uint16_t *)give(ucnt, (size_t)nu * 2);
    out->exts = (uint8_t *)give(uext, (size_t)nu);
    out->cyclic = (uint8_t *)give(ucyc, (size_t)nu);
    out->largest = (int64_t *)give(largest, (size_t)ng * 8);
    out->largest_flags = (uint8_t *)give(lflags, (size_t)ng);
    }
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
    out->ms_total = ms;
    CK(cudaGetLastError());
}
}  // namespace

extern "C" int rogtk_gpu_compress_unitigs(rogtk_gpu_ctx *ctx, const rogtk_graph_batch *in, int64_t min_length, uint32_t flags,
                                          rogtk_contig_batch *out)
{
    if (!ctx || !in || !out || !in->on_device || in->n_groups < 0 || in->n_kmers < 0 || !in->kmer_offsets || in->k_eff > 64 ||
        (in->n_kmers > 0 && (!in->kmers_lo || !in->exts || !in->counts || in->k_eff < 1)) ||
        (in->k_eff > 32 && in->n_kmers > 0 && !in->kmers_hi))
        return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    memset(out, 0, sizeof(*out));
    BatchOwner *own = new BatchOwner();
    const bool out_dev = flags & ROGTK_OUTPUT_ON_DEVICE;
    own->on_device = out_dev;
    out->_owner = own;
    int rc = ROGTK_OK;
    try {
        compress_unitigs_impl(ctx, in, min_length, flags, out, own);
    } catch (...) {
        rc = fail_code(ctx);
    }
    if (rc != ROGTK_OK) {
        for (void *p : own->ptrs) {
            if (own->on_device) dev_free(p, ctx->stream);
            else pinned_put(ctx, p);
        }
        delete own;
        memset(out, 0, sizeof(*out));
    }
    return rc;
}

extern "C" void rogtk_gpu_free_contigs(rogtk_gpu_ctx *ctx, rogtk_contig_batch *b)
{
    if (!ctx || !b || !b->_owner) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    BatchOwner *own = reinterpret_cast<BatchOwner *>(b->_owner);
    cudaSetDevice(ctx->device);
    for (void *p : own->ptrs) {
        if (own->on_device) dev_free(p, ctx->stream);
        else pinned_put(ctx, p);
    }
    delete own;
    memset(b, 0, sizeof(*b));
}

// -----------------------------------------------------------------------------------------------------------------
// bench / test utilities
// -----------------------------------------------------------------------------------------------------------------
extern "C" int rogtk_gpu_synth_reads_from(rogtk_gpu_ctx *ctx, uint64_t seed, int64_t umi_base, int64_t n_umis, int32_t reads_per_umi,
                                          int32_t read_base, int32_t read_len, int32_t template_len, uint32_t err_per_64k, uint8_t *d_data,
                                          int64_t *d_offsets, int64_t *d_group_offsets, const int64_t *d_umi_ids)
{
    if (!ctx || n_umis < 0 || reads_per_umi < 1 || read_base < 0 || read_len < 1 || template_len < read_len) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    try {
        CK(cudaSetDevice(ctx->device));
        const int64_t total = n_umis * (int64_t)reads_per_umi * read_len;
        const int64_t threads = std::max<int64_t>((total + 3) / 4, n_umis * (int64_t)reads_per_umi + 1);
        k_synth<<<blocks_for(threads, 256), 256, 0, ctx->stream>>>(seed, umi_base, n_umis, reads_per_umi, read_len, template_len,
                                                                   err_per_64k, d_data, d_offsets, d_group_offsets, read_base, d_umi_ids);
        ctx->launches++;
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
    } catch (...) {
        return fail_code(ctx);
    }
    return ROGTK_OK;
}

extern "C" int rogtk_gpu_synth_reads(rogtk_gpu_ctx *ctx, uint64_t seed, int64_t umi_base, int64_t n_umis, int32_t reads_per_umi,
                                     int32_t read_len, int32_t template_len, uint32_t err_per_64k, uint8_t *d_data,
                                     int64_t *d_offsets, int64_t *d_group_offsets)
{
    return rogtk_gpu_synth_reads_from(ctx, seed, umi_base, n_umis, reads_per_umi, 0, read_len, template_len, err_per_64k, d_data, d_offsets,
                                      d_group_offsets, nullptr);
}

extern "C" int rogtk_gpu_synth_reads_var(rogtk_gpu_ctx *ctx, uint64_t seed, int64_t umi_base, int64_t n_umis, const int64_t *d_group_offsets,
                                         int64_t n_reads, int32_t read_len, int32_t template_len, uint32_t err_per_64k, uint8_t *d_data,
                                         int64_t *d_offsets)
{
    if (!ctx || n_umis < 0 || n_reads < 0 || read_len < 1 || template_len < read_len || !d_group_offsets) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    try {
        CK(cudaSetDevice(ctx->device));
        const int64_t total = n_reads * (int64_t)read_len;
        const int64_t threads = std::max<int64_t>((total + 3) / 4, n_reads + 1);
        k_synth_var<<<blocks_for(threads, 256), 256, 0, ctx->stream>>>(seed, umi_base, n_umis, d_group_offsets, read_len, template_len, err_per_64k,
                                                                       d_data, d_offsets);
        ctx->launches++;
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
    } catch (...) {
        return fail_code(ctx);
    }
    return ROGTK_OK;
}

extern "C" void *rogtk_gpu_device_alloc(rogtk_gpu_ctx *ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
        set_error(ctx, "cudaMalloc(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void rogtk_gpu_device_free(rogtk_gpu_ctx *ctx, void *p)
{
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFree(p);
}
extern "C" void *rogtk_gpu_pinned_alloc(rogtk_gpu_ctx *ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
        set_error(ctx, "cudaHostAlloc(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void rogtk_gpu_pinned_free(rogtk_gpu_ctx *ctx, void *p)
{
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFreeHost(p);
}
extern "C" int rogtk_gpu_memcpy_h2d(rogtk_gpu_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        set_error(ctx, "h2d copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return ROGTK_ERR_CUDA;
    }
    return ROGTK_OK;
}
extern "C" int rogtk_gpu_memcpy_d2h(rogtk_gpu_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        set_error(ctx, "d2h copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return ROGTK_ERR_CUDA;
    }
    return ROGTK_OK;
}
extern "C" int64_t rogtk_gpu_launch_count(rogtk_gpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int rogtk_gpu_debug_cache_stats(rogtk_gpu_ctx *ctx, int64_t *out7)
{
    if (!ctx || !out7 || !ctx->cache) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->cache->mu);
    out7[0] = ctx->cache->n_malloc;
    out7[1] = ctx->cache->n_reuse;
    out7[2] = (int64_t)ctx->cache->idle_bytes;
    out7[3] = (int64_t)ctx->cache->live_bytes;
    out7[4] = (int64_t)ctx->cache->peak_bytes;
    out7[5] = (int64_t)ctx->cache->idle.size();
    out7[6] = ctx->cache->n_released_for_others;
    return ROGTK_OK;
}

extern "C" int rogtk_gpu_debug_phase_cycles(rogtk_gpu_ctx *ctx, int enable, uint64_t *out16)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    ctx->phase_enabled = enable != 0;
    if (out16) memcpy(out16, ctx->phase_cycles, sizeof(ctx->phase_cycles));
    return ROGTK_OK;
}

extern "C" int rogtk_gpu_flush_l2(rogtk_gpu_ctx *ctx)
{
    if (!ctx) return ROGTK_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CacheScope cache_scope(ctx->cache);
    cudaSetDevice(ctx->device);
    if (!ctx->flush_buf) {
        ctx->flush_bytes = (size_t)256 << 20;   // > 126 MB L2
        if (cudaMalloc(&ctx->flush_buf, ctx->flush_bytes) != cudaSuccess) {
            cudaGetLastError();
            return ROGTK_ERR_OOM;
        }
    }
    static int tick = 0;
    cudaMemsetAsync(ctx->flush_buf, ++tick & 0xff, ctx->flush_bytes, ctx->stream);
    return cudaStreamSynchronize(ctx->stream) == cudaSuccess ? ROGTK_OK : ROGTK_ERR_CUDA;
}

#include "pathfind_device.cuh"
#include "assemble_host.cuh"
#include "multi_host.cuh"
#include "spectrum_nccl.cuh"
