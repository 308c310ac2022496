// api.cu -- extern "C" entry points of libtreedist_cuda.so (see include/treedist_cuda.h).
#include <cuda_runtime.h>

#include <chrono>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "kernels.h"
#include "td_internal.h"

namespace td {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_counter_add(uint64_t n) { return g_launches.fetch_add(n) + n; }

// ---- per-device block cache: cudaMalloc / cudaFree cost milliseconds, a matrix call must not pay them every time -----------
struct BlockPool {
    std::mutex mu;
    std::vector<std::pair<size_t, void *>> free_blocks;
    size_t cached = 0;
};
static BlockPool g_pools[64];
// bytes kept per device before blocks really go back to the driver: TREEDIST_CACHE_BYTES, default 4 GiB (td_trim_cache() empties it)
static size_t pool_limit()
{
    static const size_t v = [] {
        const char *e = getenv("TREEDIST_CACHE_BYTES");
        if (e && *e) { char *end = nullptr; const unsigned long long x = strtoull(e, &end, 10); if (end != e) return (size_t)x; }
        return (size_t)(4ull << 30);
    }();
    return v;
}

static cudaError_t pool_alloc(int device, size_t bytes, void **out, size_t *got)
{
    bytes = (std::max<size_t>(bytes, 1) + ((1u << 20) - 1)) & ~(size_t)((1u << 20) - 1);
    BlockPool &P = g_pools[device & 63];
    {
        std::lock_guard<std::mutex> g(P.mu);
        size_t best = (size_t)-1, best_i = 0;
        for (size_t i = 0; i < P.free_blocks.size(); i++) {
            const size_t sz = P.free_blocks[i].first;
            if (sz >= bytes && sz <= 2 * bytes + (64u << 20) && sz < best) { best = sz; best_i = i; }
        }
        if (best != (size_t)-1) {
            *out = P.free_blocks[best_i].second; *got = best; P.cached -= best;
            P.free_blocks.erase(P.free_blocks.begin() + best_i);
            return cudaSuccess;
        }
    }
    *got = bytes;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {            // give the cache back and retry once
        cudaGetLastError();
        std::lock_guard<std::mutex> g(P.mu);
        for (auto &b : P.free_blocks) cudaFree(b.second);
        P.free_blocks.clear(); P.cached = 0;
        e = cudaMalloc(out, bytes);
    }
    return e;
}

static void pool_trim(int device)
{
    BlockPool &P = g_pools[device & 63];
    std::lock_guard<std::mutex> g(P.mu);
    for (auto &b : P.free_blocks) cudaFree(b.second);
    P.free_blocks.clear(); P.cached = 0;
}

static void pool_free(int device, void *p, size_t bytes)
{
    if (!p) return;
    BlockPool &P = g_pools[device & 63];
    std::lock_guard<std::mutex> g(P.mu);
    if (P.cached + bytes > pool_limit()) { cudaFree(p); return; }
    P.free_blocks.emplace_back(bytes, p); P.cached += bytes;
}

// ---- pinned host memory: cudaHostAlloc costs tens of milliseconds per 100 MB (the pages are faulted in and locked), so blocks are
// cached for the life of the process, up to TREEDIST_PINNED_CACHE_BYTES (default 4 GiB)
struct PinnedPool {
    std::mutex mu;
    std::vector<std::pair<size_t, void *>> free_blocks;
    std::vector<std::pair<void *, size_t>> live;
    size_t cached = 0;
};
static PinnedPool g_pinned;
static size_t pinned_limit()
{
    static const size_t v = [] {
        const char *e = getenv("TREEDIST_PINNED_CACHE_BYTES");
        if (e && *e) { char *end = nullptr; const unsigned long long x = strtoull(e, &end, 10); if (end != e) return (size_t)x; }
        return (size_t)(4ull << 30);
    }();
    return v;
}

static cudaError_t pinned_alloc(size_t bytes, void **out)
{
    bytes = (std::max<size_t>(bytes, 1) + ((1u << 20) - 1)) & ~(size_t)((1u << 20) - 1);
    PinnedPool &P = g_pinned;
    {
        std::lock_guard<std::mutex> g(P.mu);
        size_t best = (size_t)-1, best_i = 0;
        for (size_t i = 0; i < P.free_blocks.size(); i++) {
            const size_t sz = P.free_blocks[i].first;
            if (sz >= bytes && sz <= 2 * bytes + (16u << 20) && sz < best) { best = sz; best_i = i; }
        }
        if (best != (size_t)-1) {
            *out = P.free_blocks[best_i].second; P.cached -= best;
            P.free_blocks.erase(P.free_blocks.begin() + best_i);
            P.live.emplace_back(*out, best);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {                          // give the cache back and retry once
        cudaGetLastError();
        std::lock_guard<std::mutex> g(P.mu);
        for (auto &b : P.free_blocks) cudaFreeHost(b.second);
        P.free_blocks.clear(); P.cached = 0;
        e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> g(P.mu); P.live.emplace_back(*out, bytes); }
    return e;
}

static bool pinned_free(void *p)
{
    if (!p) return true;
    PinnedPool &P = g_pinned;
    std::lock_guard<std::mutex> g(P.mu);
    for (size_t i = 0; i < P.live.size(); i++) {
        if (P.live[i].first != p) continue;
        const size_t bytes = P.live[i].second;
        P.live.erase(P.live.begin() + i);
        if (P.cached + bytes > pinned_limit()) cudaFreeHost(p);
        else { P.free_blocks.emplace_back(bytes, p); P.cached += bytes; }
        return true;
    }
    return false;
}

static void pinned_trim()
{
    PinnedPool &P = g_pinned;
    std::lock_guard<std::mutex> g(P.mu);
    for (auto &b : P.free_blocks) cudaFreeHost(b.second);
    P.free_blocks.clear(); P.cached = 0;
}

// ---- streams and events of one host-buffer call (td_pairwise, td_pairwise_square): kernels of band k on `compute`, the device-to-host
// copy of band k on `copy`, two bands in flight.  Cached per device; a call takes a bundle for its duration, so concurrent calls on one
// set (one host thread per GPU, or several per GPU) never share a stream.
struct BandStreams {
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
};
struct StreamPool { std::mutex mu; std::vector<BandStreams> idle; };
static StreamPool g_streams[64];

static cudaError_t acquire_streams(int device, BandStreams *bs)
{
    StreamPool &P = g_streams[device & 63];
    {
        std::lock_guard<std::mutex> g(P.mu);
        if (!P.idle.empty()) { *bs = P.idle.back(); P.idle.pop_back(); return cudaSuccess; }
    }
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&bs->compute, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&bs->copy, cudaStreamNonBlocking)) != cudaSuccess) return e;
    for (int b = 0; b < 2; b++) {
        if ((e = cudaEventCreateWithFlags(&bs->done[b], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&bs->copied[b], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static void release_streams(int device, const BandStreams &bs)
{
    if (!bs.compute) return;
    StreamPool &P = g_streams[device & 63];
    std::lock_guard<std::mutex> g(P.mu);
    P.idle.push_back(bs);
}

struct DevProps { int sm_count = 0; size_t smem_optin = 0; bool valid = false; };
static DevProps g_props[64];
static std::mutex g_props_mu;

constexpr int kCounters = 64;
constexpr int kWorklists = 8;
constexpr unsigned kWorklistCap = 1u << 20;       // pairs per exact fix-up list (beyond it the tile kernel recomputes in place, slowly)

struct DeviceSet {
    int device = -1, sm_count = 148;
    int64_t N = 0, Npad = 0;
    int n_taxa = 0, n_k = 0, W = 1, stride = 1, sh_stride = 2;
    double *leafT = nullptr, *leaflen = nullptr;
    SplitRec *shared = nullptr;
    int32_t *nsplits = nullptr, *nshared = nullptr;
    double *sum_len = nullptr, *norm = nullptr, *single_l1 = nullptr, *single_l2 = nullptr, *q1 = nullptr, *q2 = nullptr, *q2t = nullptr;
    int32_t *post_tree = nullptr, *post_end = nullptr, *shared_x = nullptr;
    double *post_len = nullptr;
    int64_t n_post = 0;
    uint32_t *ids = nullptr;
    double *lens = nullptr;
    uint64_t *masks = nullptr, *leafmask = nullptr;
    int32_t *rooted = nullptr;
    unsigned long long *work_counter = nullptr, *stats = nullptr;   // work_counter: ring of kCounters, one per launch in flight
    std::atomic<uint32_t> next_counter{0};
    int2 *wl_pairs = nullptr;          // kWorklists rings of kWorklistCap pairs for the persistent hash-join kernel
    unsigned *wl_count = nullptr;
    std::atomic<uint32_t> next_worklist{0};
    int8_t *X = nullptr;               // split-incidence matrix [rows_x][d_pad] for the int8 tcgen05 RF path (built on demand)
    int64_t d_pad = 0, rows_x = 0;
    alignas(64) unsigned char tmap[128], tmap_half[128];   // boxes of a whole tile / of the half a CTA multicasts to its cluster partner
    alignas(64) unsigned char tmap_leaf_a[128], tmap_leaf_b[128];   // TMA descriptors over leafT (postings pipeline)
    bool leaf_maps = false;
    // split-extended matrix [n_k + dict_shared][Npad] for the Euclidean metric of densely sharing sets (built on demand)
    bool full_lists = false;            // ids / lens / masks / leafmask / rooted uploaded (geodesic and restrict-then-join kernels)
    std::mutex full_mu;
    double *leafX = nullptr;
    int x_rows = 0;
    // hybrid postings pipeline: leafT plus the heavy_rows most widely held shared splits as rows (built on demand), q1 without them
    double *leafH = nullptr, *q1_light = nullptr;
    alignas(64) unsigned char tmap_h_a[128], tmap_h_b[128];
    std::mutex lh_mu;
    alignas(64) unsigned char tmap_x_a[128], tmap_x_b[128];
    std::mutex lx_mu;
    std::mutex x_mu;
    // event buffers of the two-kernel join (one launch at a time per set: later launches wait on ev_done)
    std::mutex ev_mu;
    unsigned char *ev_block = nullptr;
    size_t ev_block_bytes = 0, ev_block_got = 0;
    cudaEvent_t ev_done = nullptr;
    int *error_flag = nullptr;
    // leaf-set groups of a non-uniform set: group id per tree, per group the member list and a scratch matrix per metric (the group's own
    // condensed slice before it is scattered into the set's matrix; one launch at a time per set: later launches wait on grp_done)
    int32_t *group_of = nullptr;
    int32_t *cross_cols = nullptr;
    int64_t *cross_first = nullptr, *cross_prefix = nullptr;
    std::vector<int32_t *> grp_members;
    std::mutex grp_mu;
    void *grp_tmp[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t grp_tmp_bytes = 0, grp_tmp_got[4] = {0, 0, 0, 0};
    cudaEvent_t grp_done = nullptr;
    // per-warp scratch of the big-tree kernel (one launch at a time per set: later launches wait on big_done)
    std::mutex big_mu;
    unsigned char *big_scratch = nullptr;
    size_t big_bytes = 0, big_got = 0;
    cudaEvent_t big_done = nullptr;
    size_t max_smem_optin = 0;
    td_geo_stats geo_stats{};
    std::vector<std::pair<size_t, void *>> allocs;      // pool blocks owned by this set
};

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? TD_ERR_NO_DEVICE : TD_ERR_CUDA; \
        }                                                                                                \
    } while (0)

template <class T>
static int dev_alloc(DeviceSet *d, T **ptr, size_t count)
{
    void *p = nullptr; size_t got = 0;
    CUDA_TRY(pool_alloc(d->device, std::max<size_t>(count, 1) * sizeof(T), &p, &got));
    d->allocs.emplace_back(got, p); *ptr = reinterpret_cast<T *>(p);
    return TD_OK;
}

static void free_device(DeviceSet *d)
{
    if (!d) return;
    int cur = -1; cudaGetDevice(&cur);
    if (d->device >= 0) cudaSetDevice(d->device);
    // launches on any stream may still read or write the set's memory: nothing goes back to the pool (where another thread could be
    // handed it) before the device has drained
    if (d->device >= 0) cudaDeviceSynchronize();
    for (auto &b : d->allocs) pool_free(d->device, b.second, b.first);
    if (d->ev_done) cudaEventDestroy(d->ev_done);
    if (d->ev_block) pool_free(d->device, d->ev_block, d->ev_block_got);
    if (d->grp_done) cudaEventDestroy(d->grp_done);
    for (int m = 0; m < 4; m++) if (d->grp_tmp[m]) pool_free(d->device, d->grp_tmp[m], d->grp_tmp_got[m]);
    if (d->big_done) cudaEventDestroy(d->big_done);
    if (d->big_scratch) pool_free(d->device, d->big_scratch, d->big_got);
    if (cur >= 0) cudaSetDevice(cur);
    delete d;
}

static int64_t condensed_offset(int64_t N, int64_t row) { return row * N - row * (row + 1) / 2; }

static int upload(td_treeset *ts, int device, cudaStream_t st)
{
    HostSet &h = ts->host;
    const bool timing = getenv("TREEDIST_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (!timing) return; auto n = std::chrono::steady_clock::now(); fprintf(stderr, "[upload] %-14s %.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t_last).count()); t_last = n; };
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) { set_error("no CUDA device available (%s); libtreedist_cuda has no CPU fallback", cudaGetErrorString(ce)); return TD_ERR_NO_DEVICE; }
    if (device < 0 || device >= count) { set_error("device %d out of range (0..%d)", device, count - 1); return TD_ERR_ARG; }
    std::lock_guard<std::mutex> guard(ts->mu);
    if ((int)ts->devs.size() <= device) ts->devs.resize(device + 1, nullptr);
    if (ts->devs[device]) return TD_OK;                 // already resident on this GPU
    CUDA_TRY(cudaSetDevice(device));
    DeviceSet *d = new DeviceSet();
    struct Guard { DeviceSet *d; bool keep = false; ~Guard() { if (!keep) free_device(d); } } own{d};
    d->device = device;
    {
        std::lock_guard<std::mutex> g(g_props_mu);
        DevProps &dp = g_props[device & 63];
        if (!dp.valid) {
            int v = 0;
            CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); dp.sm_count = v;
            CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)); dp.smem_optin = (size_t)v;
            dp.valid = true;
        }
        d->sm_count = dp.sm_count; d->max_smem_optin = dp.smem_optin;
    }
    d->N = h.N; d->Npad = (h.N + kTile - 1) / kTile * kTile;
    d->n_taxa = h.n_taxa; d->W = h.W; d->stride = std::max(h.max_splits, 1); d->sh_stride = h.sh_stride;
    const int kc = pairwise_kc();
    d->n_k = (h.n_taxa + kc - 1) / kc * kc;
    const size_t N = h.N, Npad = d->Npad;
    int rc;
    // One arena for every array of the set (256-byte aligned pieces): a single pool block, a single memset.  The arrays that come from the
    // host lie first and in one run: they are staged (valid rows + padding rows) in one page-locked buffer by the encoder's worker pool
    // and cross PCIe as ONE copy (seventeen pageable copies cost 1.2 ms for 6.6 MB).
    {
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off = (off + std::max<size_t>(bytes, 8) + 255) & ~(size_t)255; return o; };
        const size_t n_post = h.post_tree.size();
        const size_t o_leaflen = take(N * std::max(h.n_taxa, 1) * 8);
        const size_t o_shared = take(Npad * d->sh_stride * sizeof(SplitRec)), o_shx = take(Npad * d->sh_stride * 4);
        const size_t o_nsplits = take(Npad * 4), o_nshared = take(Npad * 4);
        const size_t o_sum = take(Npad * 8), o_norm = take(Npad * 8), o_s1 = take(Npad * 8), o_s2 = take(Npad * 8), o_q1 = take(Npad * 8), o_q2 = take(Npad * 8), o_q2t = take(Npad * 8), o_q1l = take(Npad * 8);
        const size_t o_ptree = take(n_post * 4), o_pend = take(n_post * 4), o_plen = take(n_post * 8);
        const size_t staged = off;                          // end of the part that comes from the host
        const size_t o_leafT = take((size_t)d->n_k * Npad * 8);
        const size_t o_ids = take(N * d->stride * 4), o_lens = take(N * d->stride * 8), o_masks = take(N * d->stride * d->W * 8);
        const size_t o_leafmask = take(N * d->W * 8), o_rooted = take(N * 4);
        const size_t o_counter = take(kCounters * 8), o_stats = take(4 * 8), o_err = take(4), o_wlc = take(kWorklists * 4);
        const size_t o_wl = take((size_t)kWorklists * kWorklistCap * sizeof(int2));
        unsigned char *base = nullptr;
        lap("setup");
        if ((rc = dev_alloc(d, &base, off))) return rc;
        lap("device block");
        // everything behind the staged part except the work lists starts as zero
        cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
        if (timing) { for (auto &e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], st); }
        CUDA_TRY(cudaMemsetAsync(base + staged, 0, o_wl - staged, st));
        if (timing) cudaEventRecord(ev[1], st);
        d->leafT = (double *)(base + o_leafT); d->leaflen = (double *)(base + o_leaflen); d->shared = (SplitRec *)(base + o_shared);
        d->nsplits = (int32_t *)(base + o_nsplits); d->nshared = (int32_t *)(base + o_nshared);
        d->sum_len = (double *)(base + o_sum); d->norm = (double *)(base + o_norm); d->single_l1 = (double *)(base + o_s1);
        d->single_l2 = (double *)(base + o_s2); d->q1 = (double *)(base + o_q1); d->q2 = (double *)(base + o_q2); d->q2t = (double *)(base + o_q2t); d->q1_light = (double *)(base + o_q1l);
        d->post_tree = (int32_t *)(base + o_ptree); d->post_end = (int32_t *)(base + o_pend); d->post_len = (double *)(base + o_plen);
        d->n_post = (int64_t)n_post; d->shared_x = (int32_t *)(base + o_shx);
        d->ids = (uint32_t *)(base + o_ids); d->lens = (double *)(base + o_lens); d->masks = (uint64_t *)(base + o_masks);
        d->leafmask = (uint64_t *)(base + o_leafmask); d->rooted = (int32_t *)(base + o_rooted);
        d->work_counter = (unsigned long long *)(base + o_counter); d->stats = (unsigned long long *)(base + o_stats);
        d->error_flag = (int *)(base + o_err); d->wl_count = (unsigned *)(base + o_wlc); d->wl_pairs = (int2 *)(base + o_wl);

        void *stage = nullptr;
        if (cudaError_t e = pinned_alloc(staged, &stage); e != cudaSuccess) { set_error("cudaHostAlloc of %lld bytes failed: %s", (long long)staged, cudaGetErrorString(e)); return TD_ERR_NOMEM; }
        struct Unpin { void *p; ~Unpin() { pinned_free(p); } } unpin{stage};
        lap("pinned block");
        unsigned char *sb = (unsigned char *)stage;
        const size_t pad = Npad - N;
        // padded trees: shared-split rows of sentinels (id = 0xFFFFFFFF; the length of a sentinel is never read), postings index -1, zeros elsewhere
        std::vector<CopyJob> jobs = {
            {sb + o_leaflen, h.leaflen.data(), N * h.n_taxa * sizeof(double), 0, 0},
            {sb + o_shared, h.shared.data(), N * d->sh_stride * sizeof(SplitRec), pad * d->sh_stride * sizeof(SplitRec), 0xFF},
            {sb + o_shx, h.shared_x.data(), N * d->sh_stride * sizeof(int32_t), pad * d->sh_stride * sizeof(int32_t), 0xFF},
            {sb + o_nsplits, h.n_splits.data(), N * 4, pad * 4, 0}, {sb + o_nshared, h.n_shared.data(), N * 4, pad * 4, 0},
            {sb + o_sum, h.sum_len.data(), N * 8, pad * 8, 0}, {sb + o_norm, h.norm.data(), N * 8, pad * 8, 0},
            {sb + o_s1, h.single_l1.data(), N * 8, pad * 8, 0}, {sb + o_s2, h.single_l2.data(), N * 8, pad * 8, 0},
            {sb + o_q1, h.q1.data(), N * 8, pad * 8, 0}, {sb + o_q2, h.q2.data(), N * 8, pad * 8, 0},
            {sb + o_q2t, h.q2t.data(), N * 8, pad * 8, 0}, {sb + o_q1l, h.q1_light.data(), N * 8, pad * 8, 0},
            {sb + o_ptree, h.post_tree.data(), n_post * 4, 0, 0}, {sb + o_pend, h.post_end.data(), n_post * 4, 0, 0},
            {sb + o_plen, h.post_len.data(), n_post * 8, 0, 0}};
        parallel_copy(jobs);
        lap("stage");
        CUDA_TRY(cudaMemcpyAsync(base, stage, staged, cudaMemcpyHostToDevice, st));
        if (timing) cudaEventRecord(ev[2], st);
        if (h.n_taxa > 0) {
            CUDA_TRY(launch_transpose_leaf(d->leaflen, d->leafT, h.N, h.n_taxa, d->Npad, st));
            launch_counter_add(1);
        }
        if (h.n_taxa > 0 && make_leaf_tensor_maps(d->leafT, (int64_t)d->Npad, d->n_k, d->tmap_leaf_a, d->tmap_leaf_b) == cudaSuccess) d->leaf_maps = true;
        else cudaGetLastError();
        if (timing) cudaEventRecord(ev[3], st);
        lap("enqueue");
        CUDA_TRY(cudaStreamSynchronize(st));                // (the staging buffer goes back to the pool after this)
        lap("drain");
        if (timing) { float a = 0, b = 0, c = 0; cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&c, ev[2], ev[3]);
            fprintf(stderr, "[upload] device: memset %.3f ms (%zu bytes), copy %.3f ms (%zu bytes), transpose %.3f ms\n", a, o_wl - staged, b, staged, c); for (auto &e : ev) cudaEventDestroy(e); }
    }
    // the full split lists (ids, lengths, masks), leaf masks and rootedness are only read by the geodesic and the restrict-then-join
    // kernels: they follow on first use (upload_full_lists), a matrix of rf / wrf / euc on one leaf set never pays for them
    own.keep = true;
    ts->devs[device] = d;
    return TD_OK;
}

// second part of the upload, on first use by the geodesic / restrict-then-join kernels (space is already part of the set's arena)
static int upload_full_lists(td_treeset *ts, DeviceSet *d, cudaStream_t st)
{
    std::lock_guard<std::mutex> guard(d->full_mu);
    if (d->full_lists) return TD_OK;
    const HostSet &h = ts->host;
    const size_t N = (size_t)h.N;
    CUDA_TRY(cudaMemcpyAsync(d->ids, h.ids.data(), N * d->stride * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d->lens, h.lens.data(), N * d->stride * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d->masks, h.masks.data(), N * d->stride * d->W * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d->leafmask, h.leafmask.data(), N * d->W * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d->rooted, h.tree_rooted.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));                  // launches on other streams may read the arrays from now on
    d->full_lists = true;
    return TD_OK;
}

static DeviceSet *resident(td_treeset *ts, int device)
{
    std::lock_guard<std::mutex> guard(ts->mu);
    return (device >= 0 && device < (int)ts->devs.size()) ? ts->devs[device] : nullptr;
}

// the GPU a set of output pointers lives on (they must all be device memory of one GPU)
static int device_of(void *const d_out[4], uint32_t mask, int *device)
{
    *device = -1;
    for (int m = 0; m < 4; m++) {
        if (!(mask >> m & 1u)) continue;
        if (!d_out[m]) { set_error("output pointer for metric %d is null", m); return TD_ERR_ARG; }
        cudaPointerAttributes attr;
        CUDA_TRY(cudaPointerGetAttributes(&attr, d_out[m]));
        if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) { set_error("output pointer for metric %d is not device memory", m); return TD_ERR_ARG; }
        if (*device >= 0 && attr.device != *device) { set_error("output pointers live on different GPUs"); return TD_ERR_ARG; }
        *device = attr.device;
    }
    return TD_OK;
}

static int check_ready(td_treeset *ts, int device, DeviceSet **d)
{
    if (!ts) { set_error("null treeset"); return TD_ERR_ARG; }
    *d = resident(ts, device);
    if (!*d) { set_error("treeset is not resident on device %d: call td_upload first", device); return TD_ERR_NOT_UPLOADED; }
    return TD_OK;
}

static int pick_tile(const DeviceSet *d, bool weighted)
{
    const size_t budget = std::min<size_t>(d->max_smem_optin, 200 * 1024);
    if (pairwise_smem_bytes(64, d->sh_stride, weighted) <= budget / 2) return 64;     // two CTAs per SM
    if (pairwise_smem_bytes(64, d->sh_stride, weighted) <= budget) return 64;
    if (pairwise_smem_bytes(32, d->sh_stride, weighted) <= budget) return 32;
    return 0;
}

constexpr size_t kEventBytesMax = 6ull << 30;       // above this the fused (single kernel) hash join is used instead

static size_t split_event_bytes(const DeviceSet *d, int64_t capacity, int rec)
{
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t groups = (size_t)d->Npad / pairsplit_row_group(), col_tiles = ((size_t)d->N + 31) / 32;
    return 256 + up(groups * sizeof(unsigned)) + up(groups * col_tiles * sizeof(uint2)) + up((size_t)capacity * rec);
}

static bool rf_incidence_fits(const HostSet &h);
static int run_rf_incidence(td_treeset *ts, DeviceSet *d, int normalise, int64_t row_begin, int64_t row_end, void *d_out, int out_is_u16,
                            cudaStream_t st);

// the hybrid form of the postings pipeline applies: heavy shared splits as matrix rows, the rest as events, rf from the incidence kernel
static bool hybrid_applies(const HostSet &h, const DeviceSet *d, uint32_t m3, bool rect)
{
    return !rect && (m3 & 6u) && h.heavy_rows > 0 && d->leaf_maps && rf_incidence_fits(h) && !getenv("TREEDIST_NO_HYBRID");
}

// postings pipeline: claim the set's event buffers (growing them if needed), order this launch after the previous one that used them
// a second row block handled by the same launch set (multi-GPU folded blocks)
struct RowBlock { int64_t row_begin, row_end; void *const *d_out; };

static int run_pairsplit(td_treeset *ts, DeviceSet *d, uint32_t m3, bool rect, PairParams &p, int64_t capacity, void *const d_out[4], cudaStream_t st,
                         bool rf_apart = false, const RowBlock *second = nullptr)
{
    const long long tiles = pairsplit_tiles(rect, p);
    if (tiles <= 0 && !second) return TD_OK;
    std::lock_guard<std::mutex> g(d->ev_mu);
    const size_t need = split_event_bytes(d, capacity, pairsplit_record_bytes(m3));
    if (!d->ev_done) CUDA_TRY(cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming));
    if (need > d->ev_block_bytes) {
        CUDA_TRY(cudaEventSynchronize(d->ev_done));
        if (d->ev_block) pool_free(d->device, d->ev_block, d->ev_block_got);
        d->ev_block = nullptr; d->ev_block_bytes = 0;
        void *blk = nullptr; size_t got = 0;
        CUDA_TRY(pool_alloc(d->device, need, &blk, &got));
        d->ev_block = (unsigned char *)blk; d->ev_block_bytes = need; d->ev_block_got = got;
    }
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t groups = (size_t)d->Npad / pairsplit_row_group(), col_tiles = ((size_t)d->N + 31) / 32;
    unsigned char *b = d->ev_block;
    p.ev_cursor = (uint32_t *)b; b += 256;
    unsigned *ready = (unsigned *)b; b += up(groups * sizeof(unsigned));          // (cleared together with the cursor)
    p.ready = ready; p.ready_target = 0;                     // (launch_pairsplit decides whether the launch set uses them)
    p.seg = (uint2 *)b; b += up(groups * col_tiles * sizeof(uint2));
    p.ev = b;
    p.ev_capacity = (uint32_t)capacity;
    p.col_tiles_total = (int)col_tiles; p.rg0 = p.tile_row0 * (64 / pairsplit_row_group());
    p.tmap_a = d->leaf_maps ? d->tmap_leaf_a : nullptr; p.tmap_b = d->leaf_maps ? d->tmap_leaf_b : nullptr;
    p.post_tree = d->post_tree; p.post_end = d->post_end; p.post_len = d->post_len; p.n_post = d->n_post; p.q2t = d->q2t; p.shared_x = d->shared_x;
    const HostSet &h = ts->host;
    const bool hybrid = hybrid_applies(h, d, m3, rect);
    uint32_t mk = m3;
    if (hybrid) {
        {
            std::lock_guard<std::mutex> guard(d->lh_mu);
            if (!d->leafH) {
                const int rows = d->n_k + h.heavy_rows;
                double *X = nullptr; int rc;
                if ((rc = dev_alloc(d, &X, (size_t)rows * d->Npad))) return rc;
                CUDA_TRY(cudaMemcpyAsync(X, d->leafT, (size_t)d->n_k * d->Npad * sizeof(double), cudaMemcpyDeviceToDevice, st));
                CUDA_TRY(cudaMemsetAsync(X + (size_t)d->n_k * d->Npad, 0, (size_t)h.heavy_rows * d->Npad * sizeof(double), st));
                CUDA_TRY(launch_build_split_rows(d->shared, d->sh_stride, h.N, d->Npad, X + (size_t)d->n_k * d->Npad, h.heavy_rows, st));
                launch_counter_add(1);
                if (make_leaf_tensor_maps(X, (int64_t)d->Npad, rows, d->tmap_h_a, d->tmap_h_b) != cudaSuccess) { set_error("cuTensorMapEncodeTiled failed"); return TD_ERR_CUDA; }
                CUDA_TRY(cudaStreamSynchronize(st));          // other streams may use the matrix from now on
                d->leafH = X;
            }
        }
        p.heavy_rows = h.heavy_rows; p.k_rows = d->n_k + h.heavy_rows; p.tmap_a = d->tmap_h_a; p.tmap_b = d->tmap_h_b;
        p.q1 = d->q1_light;                                   // wrf: the heavy splits are part of the L1 sum over the rows
        mk = m3 & 6u;                                         // rf comes from the incidence contraction (the events only count light splits)
    }
    if (rf_apart) mk = m3 & 6u;                               // more common splits than a byte counts: rf from the incidence contraction as well
    int launches = 0;
    if (mk) {
    CUDA_TRY(cudaStreamWaitEvent(st, d->ev_done, 0));
    CUDA_TRY(cudaMemsetAsync(p.ev_cursor, 0, 256 + up(groups * sizeof(unsigned)), st));
    unsigned long long *trace = nullptr;
    const char *trace_path = getenv("TREEDIST_TRACE");      // debug aid, only honoured by builds with -DPS_TRACE
    if (trace_path) { CUDA_TRY(cudaMalloc(&trace, (size_t)tiles * 64)); CUDA_TRY(cudaMemsetAsync(trace, 0, (size_t)tiles * 64, st)); p.trace = trace; }
    PairParams pb{};
    if (second) {                                             // block B: same buffers and tables, its own rows, outputs and tile geometry
        pb = p;
        pb.row_begin = second->row_begin; pb.row_end = second->row_end; pb.p_base = condensed_offset(d->N, second->row_begin);
        for (int m = 0; m < 3; m++) pb.out[m] = (double *)second->d_out[m];
        pairsplit_tiles(rect, pb);
        pb.rg0 = pb.tile_row0 * (64 / pairsplit_row_group());
    }
    CUDA_TRY(launch_pairsplit(mk, rect, p, d->sm_count, st, &launches, second ? &pb : nullptr));
    if (trace) {
        std::vector<unsigned long long> host((size_t)tiles * 8);
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(host.data(), trace, host.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(trace);
        if (FILE *f = fopen(trace_path, "wb")) { fwrite(host.data(), 8, host.size(), f); fclose(f); }
    }
    CUDA_TRY(cudaEventRecord(d->ev_done, st));
    }
    launch_counter_add(launches);
    if ((hybrid || rf_apart) && (m3 & 1u)) {
        if (!d_out[TD_RF]) { set_error("output pointer for metric 0 is null"); return TD_ERR_ARG; }
        int rc = run_rf_incidence(ts, d, p.normalise, p.row_begin, p.row_end, d_out[TD_RF], 0, st);
        if (rc) return rc;
        if (second) {
            if (!second->d_out[TD_RF]) { set_error("output pointer for metric 0 is null"); return TD_ERR_ARG; }
            if ((rc = run_rf_incidence(ts, d, p.normalise, second->row_begin, second->row_end, second->d_out[TD_RF], 0, st))) return rc;
        }
    }
    return TD_OK;
}

// Exact RF through the int8 tcgen05 split-incidence contraction (kernels_incidence.cu); builds the incidence matrix on first use.
static bool rf_incidence_fits(const HostSet &h)
{
    const int bk = incidence_bk(), tile = incidence_tile();
    const int64_t d_pad = std::max<int64_t>((h.dict_shared + bk - 1) / bk * bk, bk), rows = (h.N + tile - 1) / tile * tile;
    return h.uniform && h.max_splits <= 32767 && (double)d_pad * (double)rows <= 64e9;
}

static int run_rf_incidence(td_treeset *ts, DeviceSet *d, int normalise, int64_t row_begin, int64_t row_end, void *d_out, int out_is_u16,
                            cudaStream_t st)
{
    const HostSet &h = ts->host;
    int rc;
    CUDA_TRY(cudaSetDevice(d->device));
    {
        std::lock_guard<std::mutex> guard(d->x_mu);
        if (!d->X) {
            const int bk = incidence_bk(), tile = incidence_tile();
            const int64_t d_pad = std::max<int64_t>((h.dict_shared + bk - 1) / bk * bk, bk);
            const int64_t rows = (h.N + tile - 1) / tile * tile;
            if ((double)d_pad * (double)rows > 64e9) { set_error("incidence matrix of %lld x %lld bytes is too large: the set shares too few splits for this path (use td_pairwise_device)", (long long)rows, (long long)d_pad); return TD_ERR_UNSUPPORTED; }
            int8_t *X = nullptr;
            if ((rc = dev_alloc(d, &X, (size_t)rows * d_pad))) return rc;
            CUDA_TRY(cudaMemsetAsync(X, 0, (size_t)rows * d_pad, st));
            CUDA_TRY(launch_build_incidence(d->shared, d->sh_stride, h.N, d_pad, X, st));
            launch_counter_add(1);
            cudaError_t e = make_incidence_tensor_map(X, rows, d_pad, d->tmap, 0);
            if (e == cudaSuccess) e = make_incidence_tensor_map(X, rows, d_pad, d->tmap_half, incidence_half());
            if (e != cudaSuccess) { set_error("cuTensorMapEncodeTiled failed (%s)", cudaGetErrorString(e)); return TD_ERR_CUDA; }
            CUDA_TRY(cudaStreamSynchronize(st));              // other streams may use the matrix from now on (as for leafH / leafX)
            d->X = X; d->d_pad = d_pad; d->rows_x = rows;
        }
    }
    CUDA_TRY(launch_rf_incidence(d->tmap, d->tmap_half, d->nsplits, h.N, d->d_pad, row_begin, row_end, condensed_offset(h.N, row_begin), normalise,
                                 out_is_u16, d_out, st));
    launch_counter_add(1);
    return TD_OK;
}

// Euclidean metric (and RF through the incidence contraction) of a DENSELY sharing set: the shared splits become extra rows of the
// matrix behind the tile kernel's tensor maps, so the common-split products are part of its fp64 tensor-core dot product and no events
// are needed; wrf is the L1 distance of the same rows.
constexpr double kSplitRowsBytesMax = 8e9;

static bool split_rows_fit(const HostSet &h, const DeviceSet *d)
{
    const double rows = (double)d->n_k + (double)h.dict_shared;
    return h.uniform && d->leaf_maps && h.dict_shared > 0 && rows * (double)d->Npad * 8.0 <= kSplitRowsBytesMax
           && rf_incidence_fits(h);
}

static int run_split_rows(td_treeset *ts, DeviceSet *d, uint32_t m3, PairParams &p, void *const d_out[4], cudaStream_t st)
{
    const HostSet &h = ts->host;
    int rc;
    {
        std::lock_guard<std::mutex> guard(d->lx_mu);
        if (!d->leafX) {
            const int rows = d->n_k + (int)h.dict_shared;
            double *X = nullptr;
            if ((rc = dev_alloc(d, &X, (size_t)rows * d->Npad))) return rc;
            CUDA_TRY(cudaMemcpyAsync(X, d->leafT, (size_t)d->n_k * d->Npad * sizeof(double), cudaMemcpyDeviceToDevice, st));
            CUDA_TRY(cudaMemsetAsync(X + (size_t)d->n_k * d->Npad, 0, (size_t)h.dict_shared * d->Npad * sizeof(double), st));
            CUDA_TRY(launch_build_split_rows(d->shared, d->sh_stride, h.N, d->Npad, X + (size_t)d->n_k * d->Npad, (int)h.dict_shared, st));
            launch_counter_add(1);
            if (make_leaf_tensor_maps(X, (int64_t)d->Npad, rows, d->tmap_x_a, d->tmap_x_b) != cudaSuccess) { set_error("cuTensorMapEncodeTiled failed"); return TD_ERR_CUDA; }
            CUDA_TRY(cudaStreamSynchronize(st));              // other streams may use the matrix from now on
            d->leafX = X; d->x_rows = rows;
        }
    }
    if (pairsplit_tiles(false, p) > 0) {
        p.seg = nullptr; p.ev = nullptr; p.ev_capacity = 0; p.trace = nullptr;
        p.tmap_a = d->tmap_x_a; p.tmap_b = d->tmap_x_b; p.k_rows = d->x_rows; p.q2t = d->q2t;
        p.col_tiles_total = (int)((d->N + 31) / 32);
        // wrf: the L1 distance of the extended rows is the whole weighted-RF sum over shared splits and leaves (purely additive);
        // the splits a tree shares with nobody enter through q1 := single_l1
        p.q1 = d->single_l1;
        const uint32_t mw = m3 & 6u;
        CUDA_TRY(launch_pairsplit_tiles(mw, false, p, st));
        CUDA_TRY(launch_exact_fixup(mw, false, p, d->sm_count, st));
        launch_counter_add(2);
    }
    if (m3 & 1u) {
        if (!d_out[TD_RF]) { set_error("output pointer for metric 0 is null"); return TD_ERR_ARG; }
        if ((rc = run_rf_incidence(ts, d, p.normalise, p.row_begin, p.row_end, d_out[TD_RF], 0, st))) return rc;
    }
    return TD_OK;
}

// ---- leaf-set groups of a non-uniform set ---------------------------------------------------------------------------------------------------
constexpr int64_t kMaxCrossList = 64ll << 20;       // entries of td_treeset::cross_cols (256 MB); beyond it same-group pairs are skipped chunk by chunk
constexpr int kMinGroup = 48;                       // fewer trees with one leaf set are not worth kernels of their own

static int ensure_groups(td_treeset *ts)
{
    std::lock_guard<std::mutex> guard(ts->groups_mu);
    if (ts->groups_built) return TD_OK;
    const HostSet &h = ts->host;
    ts->group_of.assign((size_t)h.N, -1);
    if (!h.uniform && !getenv("TREEDIST_NO_GROUPS")) {
        // bucket by (rootedness, leaf mask)
        std::vector<int32_t> order((size_t)h.N);
        for (int64_t i = 0; i < h.N; i++) order[(size_t)i] = (int32_t)i;
        auto key_less = [&](int32_t a, int32_t b) {
            if (h.tree_rooted[a] != h.tree_rooted[b]) return h.tree_rooted[a] < h.tree_rooted[b];
            const int c = memcmp(&h.leafmask[(size_t)a * h.W], &h.leafmask[(size_t)b * h.W], (size_t)h.W * 8);
            return c != 0 ? c < 0 : a < b;
        };
        std::sort(order.begin(), order.end(), key_less);
        for (size_t lo = 0; lo < order.size();) {
            size_t hi = lo + 1;
            while (hi < order.size() && h.tree_rooted[order[hi]] == h.tree_rooted[order[lo]]
                   && !memcmp(&h.leafmask[(size_t)order[hi] * h.W], &h.leafmask[(size_t)order[lo] * h.W], (size_t)h.W * 8)) hi++;
            if ((int)(hi - lo) >= kMinGroup) {
                td_treeset::Group g;
                g.members.assign(order.begin() + lo, order.begin() + hi);          // ascending (ties of the sort are broken by index)
                g.sub = new (std::nothrow) td_treeset();
                if (!g.sub) { set_error("out of memory"); return TD_ERR_NOMEM; }
                int rc;
                try { rc = encode_group(h, g.members, 0, g.sub->host); }
                catch (const std::bad_alloc &) { set_error("out of memory while encoding a leaf-set group"); rc = TD_ERR_NOMEM; }
                if (rc) { delete g.sub; return rc; }
                g.sub->groups_built = true;                                          // a group is uniform: no groups of its own
                for (int32_t t : g.members) ts->group_of[(size_t)t] = (int32_t)ts->groups.size();
                ts->groups.push_back(std::move(g));
            }
            lo = hi;
        }
        // the pairs across groups as lists (td_internal.h): per group the sorted trees outside it; a tree in no group wants every later tree
        if (!ts->groups.empty() && !getenv("TREEDIST_NO_CROSS_LISTS")) {
            int64_t total = 0;
            for (auto &g : ts->groups) total += h.N - (int64_t)g.members.size();
            if (total <= kMaxCrossList) {
                std::vector<int64_t> group_begin(ts->groups.size() + 1, 0);
                for (size_t g = 0; g < ts->groups.size(); g++) group_begin[g + 1] = group_begin[g] + (h.N - (int64_t)ts->groups[g].members.size());
                ts->cross_cols.resize((size_t)total);
                std::vector<int64_t> fill(group_begin.begin(), group_begin.end() - 1);
                for (int64_t t = 0; t < h.N; t++)
                    for (size_t g = 0; g < ts->groups.size(); g++) if (ts->group_of[(size_t)t] != (int32_t)g) ts->cross_cols[(size_t)fill[g]++] = (int32_t)t;
                ts->cross_first.assign((size_t)h.N, -1); ts->cross_prefix.assign((size_t)h.N + 1, 0);
                for (int64_t i = 0; i < h.N; i++) {
                    const int32_t g = ts->group_of[(size_t)i];
                    int64_t count = h.N - 1 - i;
                    if (g >= 0) {
                        const int32_t *b = ts->cross_cols.data() + group_begin[(size_t)g], *e = ts->cross_cols.data() + group_begin[(size_t)g + 1];
                        const int32_t *pos = std::upper_bound(b, e, (int32_t)i);
                        ts->cross_first[(size_t)i] = pos - ts->cross_cols.data(); count = e - pos;
                    }
                    ts->cross_prefix[(size_t)i + 1] = ts->cross_prefix[(size_t)i] + count;
                }
            }
        }
    }
    ts->groups_built = true;
    return TD_OK;
}

// Trees with more than 128 internal edges (or more than 3 mask words): the big-tree kernel, whose per-split state lives in a per-warp
// slab of global memory.  g carries the pair range and the outputs (g.outs / g.metric_mask).
constexpr size_t kBigScratchBudget = 6ull << 30;

static int run_bigtree(td_treeset *ts, DeviceSet *d, GeoParams &g, int general, cudaStream_t st)
{
    const HostSet &h = ts->host;
    if (h.max_splits > bigtree_max_splits() || h.W > 64) {
        set_error("trees with %d internal edges over %d taxa (%d mask words) exceed the big-tree kernel (at most %d internal edges, 4096 taxa)",
                  h.max_splits, h.n_taxa, h.W, bigtree_max_splits());
        return TD_ERR_UNSUPPORTED;
    }
    const size_t warp_bytes = bigtree_warp_bytes(h.max_splits, h.W, h.n_taxa);
    const int blocks = bigtree_blocks(g.p_end - g.p_begin, d->sm_count, warp_bytes, kBigScratchBudget);
    const size_t need = warp_bytes * 4 * (size_t)blocks;
    std::lock_guard<std::mutex> guard(d->big_mu);
    if (!d->big_done) CUDA_TRY(cudaEventCreateWithFlags(&d->big_done, cudaEventDisableTiming));
    if (need > d->big_bytes) {
        CUDA_TRY(cudaEventSynchronize(d->big_done));
        if (d->big_scratch) pool_free(d->device, d->big_scratch, d->big_got);
        d->big_scratch = nullptr; d->big_bytes = 0;
        void *blk = nullptr; size_t got = 0;
        cudaError_t e = pool_alloc(d->device, need, &blk, &got);
        if (e != cudaSuccess) { set_error("cudaMalloc of %lld bytes of big-tree scratch failed: %s", (long long)need, cudaGetErrorString(e)); return TD_ERR_NOMEM; }
        d->big_scratch = (unsigned char *)blk; d->big_bytes = need; d->big_got = got;
    }
    CUDA_TRY(cudaStreamWaitEvent(st, d->big_done, 0));
    CUDA_TRY(cudaMemsetAsync(g.work_counter, 0, sizeof(unsigned long long), st));
    CUDA_TRY(launch_bigtree(g, general, h.max_splits, d->big_scratch, blocks, st));
    CUDA_TRY(cudaEventRecord(d->big_done, st));
    launch_counter_add(1);
    return TD_OK;
}

// Warp-per-pair GTP kernels (geodesic / general): plan the launch, provide the per-warp flow matrices in global memory when the plan asks
// for them (the set's big-tree scratch buffer, one such launch at a time: later ones wait on big_done), launch.
static int run_warp_kernel(td_treeset *ts, DeviceSet *d, GeoParams &g, int general, cudaStream_t st)
{
    const HostSet &h = ts->host;
    const int cap = h.max_splits <= 32 ? 32 : (h.max_splits <= 64 ? 64 : 128);
    GeoPlan plan;
    CUDA_TRY(general ? plan_general(g, cap, d->sm_count, plan) : plan_geodesic(g, cap, d->sm_count, plan));
    CUDA_TRY(cudaMemsetAsync(g.work_counter, 0, sizeof(unsigned long long), st));
    if (plan.scratch_bytes == 0) {
        g.flow_scratch = nullptr;
        CUDA_TRY(general ? launch_general(g, cap, plan, st) : launch_geodesic(g, cap, plan, st));
    } else {
        std::lock_guard<std::mutex> guard(d->big_mu);
        if (!d->big_done) CUDA_TRY(cudaEventCreateWithFlags(&d->big_done, cudaEventDisableTiming));
        if (plan.scratch_bytes > d->big_bytes) {
            CUDA_TRY(cudaEventSynchronize(d->big_done));
            if (d->big_scratch) pool_free(d->device, d->big_scratch, d->big_got);
            d->big_scratch = nullptr; d->big_bytes = 0;
            void *blk = nullptr; size_t got = 0;
            cudaError_t e = pool_alloc(d->device, plan.scratch_bytes, &blk, &got);
            if (e != cudaSuccess) { set_error("cudaMalloc of %lld bytes of flow scratch failed: %s", (long long)plan.scratch_bytes, cudaGetErrorString(e)); return TD_ERR_NOMEM; }
            d->big_scratch = (unsigned char *)blk; d->big_bytes = plan.scratch_bytes; d->big_got = got;
        }
        CUDA_TRY(cudaStreamWaitEvent(st, d->big_done, 0));
        g.flow_scratch = (double *)d->big_scratch;
        CUDA_TRY(general ? launch_general(g, cap, plan, st) : launch_geodesic(g, cap, plan, st));
        CUDA_TRY(cudaEventRecord(d->big_done, st));
    }
    launch_counter_add(1);
    return TD_OK;
}

static int upload(td_treeset *ts, int device, cudaStream_t st);
static int run_pairwise(td_treeset *ts, DeviceSet *d, uint32_t mask, int normalise, int min_overlap, double fail_value, bool rect,
                        int64_t split, int64_t row_begin, int64_t row_end, void *const d_out[4], cudaStream_t st, const RowBlock *second = nullptr);

// Same-group pairs of a non-uniform set: every group runs the uniform kernels on its own sub-set (local rows that fall into
// [row_begin, row_end)) into a scratch matrix, which is then scattered into the set's condensed matrix.
static int run_groups(td_treeset *ts, DeviceSet *d, uint32_t mask, int normalise, int64_t row_begin, int64_t row_end, void *const d_out[4], cudaStream_t st)
{
    const HostSet &h = ts->host;
    std::lock_guard<std::mutex> guard(d->grp_mu);
    if (!d->grp_done) CUDA_TRY(cudaEventCreateWithFlags(&d->grp_done, cudaEventDisableTiming));
    if (d->grp_members.empty()) {
        d->grp_members.assign(ts->groups.size(), nullptr);
        for (size_t k = 0; k < ts->groups.size(); k++) {
            int rc = dev_alloc(d, &d->grp_members[k], ts->groups[k].members.size()); if (rc) return rc;
            CUDA_TRY(cudaMemcpyAsync(d->grp_members[k], ts->groups[k].members.data(), ts->groups[k].members.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    const int64_t p_base = condensed_offset(h.N, row_begin);
    for (size_t k = 0; k < ts->groups.size(); k++) {
        td_treeset::Group &g = ts->groups[k];
        const int64_t n_local = (int64_t)g.members.size();
        const int64_t a0 = std::lower_bound(g.members.begin(), g.members.end(), (int32_t)row_begin) - g.members.begin();
        const int64_t a1 = std::lower_bound(g.members.begin(), g.members.end(), (int32_t)std::min<int64_t>(row_end, INT32_MAX)) - g.members.begin();
        const int64_t local_pairs = condensed_offset(n_local, a1) - condensed_offset(n_local, a0);
        if (local_pairs <= 0) continue;
        int rc;
        if ((rc = upload(g.sub, d->device, st))) return rc;
        DeviceSet *sd = resident(g.sub, d->device);
        const size_t need = (size_t)local_pairs * sizeof(double);
        CUDA_TRY(cudaStreamWaitEvent(st, d->grp_done, 0));                         // the scratch matrices are free again
        if (need > d->grp_tmp_bytes) {
            CUDA_TRY(cudaEventSynchronize(d->grp_done));
            for (int m = 0; m < 4; m++) if (d->grp_tmp[m]) { pool_free(d->device, d->grp_tmp[m], d->grp_tmp_got[m]); d->grp_tmp[m] = nullptr; }
            d->grp_tmp_bytes = need;
        }
        void *tmp[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int m = 0; m < 4; m++) if (mask >> m & 1u) {
            if (!d->grp_tmp[m]) {                                                  // (every scratch matrix holds grp_tmp_bytes)
                cudaError_t e = pool_alloc(d->device, d->grp_tmp_bytes, &d->grp_tmp[m], &d->grp_tmp_got[m]);
                if (e != cudaSuccess) { set_error("cudaMalloc of %lld bytes of group scratch failed: %s", (long long)d->grp_tmp_bytes, cudaGetErrorString(e)); return TD_ERR_NOMEM; }
            }
            tmp[m] = d->grp_tmp[m];
        }
        if ((rc = run_pairwise(g.sub, sd, mask, normalise, 4, 0.0, false, 0, a0, a1, tmp, st))) return rc;
        for (int m = 0; m < 4; m++) if (mask >> m & 1u) {
            CUDA_TRY(launch_scatter_group((const double *)tmp[m], (double *)d_out[m], d->grp_members[k], n_local, a0, a1, h.N, p_base, st));
            launch_counter_add(1);
        }
        CUDA_TRY(cudaEventRecord(d->grp_done, st));
    }
    return TD_OK;
}

// second: another row block of the same matrix computed by the same call (two folded blocks of a multi-GPU rank): the postings pipeline
// then shares one events kernel and one fix-up kernel between the blocks; every other path simply runs the blocks one after the other
static int run_pairwise(td_treeset *ts, DeviceSet *d, uint32_t mask, int normalise, int min_overlap, double fail_value, bool rect,
                        int64_t split, int64_t row_begin, int64_t row_end, void *const d_out[4], cudaStream_t st, const RowBlock *second)
{
    const HostSet &h = ts->host;
    const bool big = h.max_splits > 128 || h.W > 3;
    if (second && (!h.uniform || rect || second->row_end <= second->row_begin)) {
        int rc = run_pairwise(ts, d, mask, normalise, min_overlap, fail_value, rect, split, row_begin, row_end, d_out, st, nullptr);
        if (rc || second->row_end <= second->row_begin) return rc;
        return run_pairwise(ts, d, mask, normalise, min_overlap, fail_value, rect, split, second->row_begin, second->row_end, second->d_out, st, nullptr);
    }
    bool second_m3_done = false;                      // rf / wrf / euc of the second block went into the merged launch set
    if (!h.uniform) {
        // pairs may differ in leaf set: restrict-then-join kernel, one warp per pair, all requested metrics at once
        CUDA_TRY(cudaSetDevice(d->device));
        { int rc = upload_full_lists(ts, d, st); if (rc) return rc; }
        // trees that share a leaf set with enough others form groups served by the uniform kernels (triangular launches only)
        { int rc = ensure_groups(ts); if (rc) return rc; }
        const bool grouped = !rect && !ts->groups.empty();
        if (grouped) {
            std::lock_guard<std::mutex> guard(d->grp_mu);
            if (!d->group_of) {
                int rc = dev_alloc(d, &d->group_of, (size_t)h.N); if (rc) return rc;
                CUDA_TRY(cudaMemcpyAsync(d->group_of, ts->group_of.data(), (size_t)h.N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
                if (!ts->cross_prefix.empty()) {
                    rc = dev_alloc(d, &d->cross_cols, std::max<size_t>(ts->cross_cols.size(), 1)); if (rc) return rc;
                    rc = dev_alloc(d, &d->cross_first, (size_t)h.N); if (rc) return rc;
                    rc = dev_alloc(d, &d->cross_prefix, (size_t)h.N + 1); if (rc) return rc;
                    CUDA_TRY(cudaMemcpyAsync(d->cross_cols, ts->cross_cols.data(), ts->cross_cols.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
                    CUDA_TRY(cudaMemcpyAsync(d->cross_first, ts->cross_first.data(), (size_t)h.N * sizeof(int64_t), cudaMemcpyHostToDevice, st));
                    CUDA_TRY(cudaMemcpyAsync(d->cross_prefix, ts->cross_prefix.data(), ((size_t)h.N + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
                }
                CUDA_TRY(cudaStreamSynchronize(st));
            }
        }
        GeoParams g{};
        g.lens = d->lens; g.masks = d->masks; g.nsplits = d->nsplits; g.leaflen = d->leaflen; g.norm = d->norm;
        g.leafmask = d->leafmask; g.rooted = d->rooted; g.group_of = grouped ? d->group_of : nullptr;
        g.N = d->N; g.stride = d->stride; g.n_taxa = d->n_taxa; g.W = d->W; g.normalise = normalise; g.flow_rows = std::max(h.max_splits, 1);
        g.metric_mask = mask; g.min_overlap = min_overlap; g.fail_value = fail_value;
        for (int m = 0; m < 4; m++) {
            g.outs[m] = (double *)d_out[m];
            if ((mask >> m & 1u) && !g.outs[m]) { set_error("output pointer for metric %d is null", m); return TD_ERR_ARG; }
        }
        g.work_counter = d->work_counter + (d->next_counter.fetch_add(1) % kCounters); g.stats = d->stats; g.error_flag = d->error_flag;
        if (rect) { g.rect = 1; g.rect_split = split; g.rect_cols = d->N - split; g.p_begin = 0; g.p_end = split * g.rect_cols; g.p_base = 0; }
        else { g.rect = 0; g.p_begin = condensed_offset(d->N, row_begin); g.p_end = condensed_offset(d->N, row_end); g.p_base = g.p_begin; }
        const bool any_pairs = g.p_end > g.p_begin;
        if (grouped && d->cross_prefix) {                   // the wanted pairs of these rows, by list
            g.group_of = nullptr; g.cross_prefix = d->cross_prefix; g.cross_first = d->cross_first; g.cross_cols = d->cross_cols;
            g.p_begin = ts->cross_prefix[(size_t)row_begin]; g.p_end = ts->cross_prefix[(size_t)row_end];
        }
        if (g.p_end > g.p_begin) {
            if (big) { int rc = run_bigtree(ts, d, g, 1, st); if (rc) return rc; }
            else { int rc = run_warp_kernel(ts, d, g, 1, st); if (rc) return rc; }
        }
        if (any_pairs && grouped) return run_groups(ts, d, mask, normalise, row_begin, row_end, d_out, st);
        return TD_OK;
    }
    CUDA_TRY(cudaSetDevice(d->device));
    const uint32_t m3 = mask & 7u;
    if (m3) {
        const bool weighted = (m3 & 6u) != 0;
        // kernel choice: sparse sharing -> tile hash join (work ~ entries + matches); dense sharing -> merge-join
        const int slots = pairjoin_table_slots(h.max_shared);
        const char *force = getenv("TREEDIST_PAIR_KERNEL");
        if (force && (!*force || !strcmp(force, "auto"))) force = nullptr;
        const int row_slots = pairjoin_row_table_slots(h.max_tile_entries);
        const bool join_fits = h.max_shared <= 255 && pairjoin_persistent_smem_bytes(m3, row_slots, d->sh_stride) <= std::min<size_t>(d->max_smem_optin, 200 * 1024)
                               && pairjoin_smem_bytes(m3, slots) <= std::min<size_t>(d->max_smem_optin, 200 * 1024);
        // events the postings pipeline can generate at most: all (pair, common split) incidences, minus those of the heavy rows in its hybrid form
        const int64_t ev_capacity = std::max<int64_t>(h.total_common - (hybrid_applies(h, d, m3, rect) ? h.heavy_events : 0), 1);
        // the tile kernel counts a pair's common splits in one byte per cell: with more than 255 shared splits per tree RF has to come from
        // the incidence contraction (int32 accumulators), as in the hybrid form; without RF the counters are not read at all
        const bool rf_apart = (m3 & 1u) && h.max_shared > 255 && !hybrid_applies(h, d, m3, rect);
        const bool rf_apart_ok = !rf_apart || (!rect && rf_incidence_fits(h) && h.dict_shared > 0);
        const bool split_fits = rf_apart_ok && (d->leaf_maps || !weighted) && ev_capacity < (1ll << 31) && split_event_bytes(d, ev_capacity, 32) <= kEventBytesMax
                                && pairsplit_events_smem_bytes(m3, d->sh_stride, (int)((d->N + 31) / 32)) <= std::min<size_t>(d->max_smem_optin, 200 * 1024);
        // Kernel choice by a cost model in picoseconds per pair, calibrated on one B200 (tools/kernel_choice.py, tools/rf_dense_sharing.py):
        //   postings pipeline      11 + 13.4 x (expected common splits per pair): work follows the events; in its hybrid form the
        //                          splits held by the most trees are rows of the matrix (0.064 each) and leave the events
        //   merge kernel           1.83 x (S_i + S_j): every pair walks both shared lists
        //   split-extended tiles   0.064 x rows of the extended matrix (x 2.4 when it has to do the L1 sums of wrf), wrf + euc
        //                          together excluded (the merge kernel does both in one pass in that time); rf rides on the
        //                          incidence contraction, which alone costs next to nothing
        const double c_split = hybrid_applies(h, d, m3, rect)
                                   ? 11.0 + 13.4 * (h.mean_common - h.heavy_common) + 0.064 * h.heavy_rows * ((m3 & 2u) ? 2.4 : 1.0)
                                   : 11.0 + 13.4 * h.mean_common;
        const double c_merge = 1.83 * 2.0 * (double)h.total_splits / (double)std::max<int64_t>(h.N, 1);
        const bool ext_ok = !rect && split_rows_fit(h, d) && ((m3 & 6u) != 6u);
        const double c_ext = ext_ok ? 0.064 * ((double)d->n_k + (double)h.dict_shared) * ((m3 & 2u) ? 2.4 : 1.0) : 1e30;
        const bool inc_ok = m3 == 1u && !rect && rf_incidence_fits(h) && h.dict_shared > 0;
        enum { K_SPLIT, K_JOIN, K_MERGE, K_EXT, K_INC } kernel;
        if (inc_ok && c_split > 20.0) kernel = K_INC;                               // rf alone unless the set shares almost nothing
        else if (split_fits && c_split <= std::min(c_merge, c_ext)) kernel = K_SPLIT;
        else if (!split_fits && join_fits && 29.0 + 62.0 * h.mean_common <= std::min(c_merge, c_ext)) kernel = K_JOIN;   // fused hash join (event buffer too large)
        else kernel = (m3 & 6u) && c_ext < c_merge ? K_EXT : K_MERGE;
        if (force) {                                                                // a forced kernel that cannot run falls back to the merge kernel
            kernel = K_MERGE;
            if (!strcmp(force, "merge")) kernel = K_MERGE;
            else if (!strcmp(force, "split") && (split_fits || join_fits)) kernel = split_fits ? K_SPLIT : K_JOIN;
            else if ((!strcmp(force, "join") || !strcmp(force, "join_tile")) && join_fits) kernel = K_JOIN;
            else if (!strcmp(force, "incidence")) kernel = inc_ok ? K_INC : K_MERGE;
            else if (!strcmp(force, "extended")) kernel = (!rect && split_rows_fit(h, d) && (m3 & 6u)) ? K_EXT : (inc_ok ? K_INC : K_MERGE);
        }
        if (getenv("TREEDIST_DEBUG")) fprintf(stderr, "[treedist] mask %u: kernel %d (0 split, 1 join, 2 merge, 3 extended, 4 incidence); cost model split %.1f merge %.1f ext %.1f ps/pair; join_fits %d split_fits %d; heavy rows %d covering %.2f of %.2f common splits per pair, hybrid %d\n", m3, (int)kernel, c_split, c_merge, c_ext, (int)join_fits, (int)split_fits, h.heavy_rows, h.heavy_common, h.mean_common, (int)hybrid_applies(h, d, m3, rect));
        const bool use_split = kernel == K_SPLIT, use_join = kernel == K_SPLIT || kernel == K_JOIN, use_ext = kernel == K_EXT;
        if (kernel == K_INC) {
            if (!d_out[TD_RF]) { set_error("output pointer for metric 0 is null"); return TD_ERR_ARG; }
            int rc = run_rf_incidence(ts, d, normalise, row_begin, row_end, d_out[TD_RF], 0, st);
            if (rc) return rc;
            goto geodesic;
        }
        const int tile = (use_join || use_ext) ? 64 : pick_tile(d, weighted);
        if (!tile) {
            // lists too long for the shared-memory merge kernel and no other form applies: the big-tree kernel (one warp per pair) serves
            // rf / wrf / euc as well - slower, but no tree size is refused
            CUDA_TRY(cudaSetDevice(d->device));
            { int rc = upload_full_lists(ts, d, st); if (rc) return rc; }
            GeoParams g{};
            g.ids = d->ids; g.lens = d->lens; g.masks = d->masks; g.nsplits = d->nsplits; g.leaflen = d->leaflen; g.norm = d->norm;
            g.leafmask = d->leafmask; g.rooted = d->rooted;
            g.N = d->N; g.stride = d->stride; g.n_taxa = d->n_taxa; g.W = d->W; g.normalise = normalise; g.flow_rows = std::max(h.max_splits, 1);
            g.metric_mask = m3; g.min_overlap = min_overlap; g.fail_value = fail_value;
            for (int m = 0; m < 3; m++) { g.outs[m] = (double *)d_out[m]; if ((m3 >> m & 1u) && !g.outs[m]) { set_error("output pointer for metric %d is null", m); return TD_ERR_ARG; } }
            g.work_counter = d->work_counter + (d->next_counter.fetch_add(1) % kCounters); g.stats = d->stats; g.error_flag = d->error_flag;
            if (rect) { g.rect = 1; g.rect_split = split; g.rect_cols = d->N - split; g.p_begin = 0; g.p_end = split * g.rect_cols; g.p_base = 0; }
            else { g.rect = 0; g.p_begin = condensed_offset(d->N, row_begin); g.p_end = condensed_offset(d->N, row_end); g.p_base = g.p_begin; }
            if (g.p_end > g.p_begin) { int rc = run_bigtree(ts, d, g, 0, st); if (rc) return rc; }
            goto geodesic;
        }
        PairParams p{};
        p.leafT = d->leafT; p.ldT = d->Npad; p.shared = d->shared; p.sh_stride = d->sh_stride; p.nsplits = d->nsplits; p.nshared = d->nshared;
        p.sum_len = d->sum_len; p.norm = d->norm; p.single_l1 = d->single_l1; p.single_l2 = d->single_l2;
        p.q1 = d->q1; p.q2 = d->q2; p.table_slots = slots; p.row_table_slots = row_slots;
        p.N = d->N; p.n_k = d->n_k; p.n_taxa = d->n_taxa; p.k_rows = d->n_taxa; p.row_begin = row_begin; p.row_end = row_end; p.normalise = normalise;
        p.out[0] = (double *)d_out[TD_RF]; p.out[1] = (double *)d_out[TD_WRF]; p.out[2] = (double *)d_out[TD_EUC];
        for (int m = 0; m < 3; m++) if ((m3 >> m & 1u) && !p.out[m]) { set_error("output pointer for metric %d is null", m); return TD_ERR_ARG; }
        int64_t tile_cols;
        if (rect) {
            // column tiles start at the tile boundary at or below `split` so that every B tile is tile aligned
            // (16-byte cp.async sources, no read past the padded set); columns below `split` are masked on store
            p.col_offset = split; p.n_cols = d->N - split; p.col_end = d->N; p.p_base = 0;
            p.col_tile0 = (int)(split / tile);
            tile_cols = (d->N + tile - 1) / tile - p.col_tile0;
        }
        else { p.col_offset = 0; p.n_cols = d->N; p.col_end = d->N; p.col_tile0 = 0; p.p_base = condensed_offset(d->N, row_begin); tile_cols = (d->N + tile - 1) / tile; }
        p.tile_row0 = (int)(row_begin / tile);
        const int64_t tile_rows = (row_end + tile - 1) / tile - p.tile_row0;
        if (tile_rows > 0 && tile_cols > 0) {
            if (use_join || use_ext) {
                const uint32_t w = d->next_worklist.fetch_add(1) % kWorklists;
                p.wl_pairs = d->wl_pairs + (size_t)w * kWorklistCap; p.wl_count = d->wl_count + w; p.wl_capacity = kWorklistCap;
                CUDA_TRY(cudaMemsetAsync(p.wl_count, 0, sizeof(unsigned), st));
            }
            if (use_ext) { int rc = run_split_rows(ts, d, m3, p, d_out, st); if (rc) return rc; }
            else if (use_split) {
                if (second) for (int m = 0; m < 3; m++) if ((m3 >> m & 1u) && !second->d_out[m]) { set_error("output pointer for metric %d is null", m); return TD_ERR_ARG; }
                int rc = run_pairsplit(ts, d, m3, rect, p, ev_capacity, d_out, st, rf_apart, second); if (rc) return rc;
                second_m3_done = second != nullptr;
            }
            else {
                if (use_join) CUDA_TRY(launch_pairjoin(m3, rect, p, d->sm_count, !(force && !strcmp(force, "join_tile")), st));
                else CUDA_TRY(launch_pairwise(m3, rect, p, tile_rows, tile_cols, tile, st));
                launch_counter_add(use_join && weighted ? 2 : 1);
            }
        }
    }
geodesic:
    if (mask & TD_MASK(TD_GEO)) {
        if (!d_out[TD_GEO]) { set_error("output pointer for the geodesic metric is null"); return TD_ERR_ARG; }
        { int rc = upload_full_lists(ts, d, st); if (rc) return rc; }
        GeoParams g{};
        g.ids = d->ids; g.lens = d->lens; g.masks = d->masks; g.nsplits = d->nsplits; g.leaflen = d->leaflen; g.norm = d->norm;
        g.N = d->N; g.stride = d->stride; g.n_taxa = d->n_taxa; g.W = d->W; g.normalise = normalise; g.flow_rows = std::max(h.max_splits, 1);
        g.out = (double *)d_out[TD_GEO]; g.work_counter = d->work_counter + (d->next_counter.fetch_add(1) % kCounters); g.stats = d->stats; g.error_flag = d->error_flag;
        if (rect) { g.rect = 1; g.rect_split = split; g.rect_cols = d->N - split; g.p_begin = 0; g.p_end = split * g.rect_cols; g.p_base = 0; }
        else { g.rect = 0; g.p_begin = condensed_offset(d->N, row_begin); g.p_end = condensed_offset(d->N, row_end); g.p_base = g.p_begin; }
        if (g.p_end > g.p_begin && big) {
            g.outs[3] = g.out; g.metric_mask = 8u;
            int rc = run_bigtree(ts, d, g, 0, st); if (rc) return rc;
        } else if (g.p_end > g.p_begin) {
            int rc = run_warp_kernel(ts, d, g, 0, st); if (rc) return rc;
        }
    }
    if (second) {                                     // what the merged launch set did not cover for the second block
        const uint32_t rest = second_m3_done ? (mask & TD_MASK(TD_GEO)) : mask;
        if (rest) return run_pairwise(ts, d, rest, normalise, min_overlap, fail_value, rect, split, second->row_begin, second->row_end, second->d_out, st, nullptr);
    }
    return TD_OK;
}

}  // namespace td

using namespace td;

td_treeset::~td_treeset()
{
    for (Group &g : groups) delete g.sub;
    for (DeviceSet *d : devs) free_device(d);
}

namespace {
// the entry points select the set's GPU; the caller's current device is put back on the way out
struct DeviceScope {
    int saved = -1;
    DeviceScope() { if (cudaGetDevice(&saved) != cudaSuccess) { saved = -1; cudaGetLastError(); } }
    ~DeviceScope() { if (saved >= 0) cudaSetDevice(saved); }
};

// reads and clears the sticky error flag of a resident set and of its leaf-set groups (the caller has synchronised the launches it asks about)
int take_error_flag(td_treeset *ts, DeviceSet *d)
{
    int flag = 0;
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaMemcpy(&flag, d->error_flag, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) CUDA_TRY(cudaMemset(d->error_flag, 0, sizeof flag));
    {
        std::lock_guard<std::mutex> guard(ts->groups_mu);
        for (td_treeset::Group &g : ts->groups) {
            DeviceSet *sd = resident(g.sub, d->device);
            if (!sd) continue;
            int f = 0;
            CUDA_TRY(cudaMemcpy(&f, sd->error_flag, sizeof f, cudaMemcpyDeviceToHost));
            if (f) CUDA_TRY(cudaMemset(sd->error_flag, 0, sizeof f));
            flag |= f;
        }
    }
    if (flag & 2) { set_error("cannot compare a rooted with an unrooted tree (undefined in the reference as well)"); return TD_ERR_MIXED_ROOTING; }
    if (flag & 1) { set_error("geodesic kernel hit its iteration cap on at least one pair (result is NaN there)"); return TD_ERR_CUDA; }
    return TD_OK;
}
}  // namespace

extern "C" {

const char *td_last_error(void) { return last_error(); }
const char *td_version(void) { return "treedist_cuda 0.1 (sm_100a)"; }
uint64_t td_launch_count(void) { return g_launches.load(); }

int td_encode_newick(const char *const *newick, int64_t n_trees, int n_threads, td_treeset **out)
{
    if (!out) { set_error("null output pointer"); return TD_ERR_ARG; }
    *out = nullptr;
    td_treeset *ts = new (std::nothrow) td_treeset();
    if (!ts) { set_error("out of memory"); return TD_ERR_NOMEM; }
    int rc;
    const auto t0 = std::chrono::steady_clock::now();
    try { rc = encode_newick(newick, n_trees, n_threads, ts->host); }
    catch (const std::bad_alloc &) { set_error("out of memory while encoding"); rc = TD_ERR_NOMEM; }
    if (getenv("TREEDIST_TIMING")) fprintf(stderr, "[encode] td_encode_newick %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    if (rc) { delete ts; return rc; }
    *out = ts;
    return TD_OK;
}

int td_encode_kendall_colijn(const char *const *newick, int64_t n_trees, double lambda, int n_threads, td_treeset **out)
{
    if (!out) { set_error("null output pointer"); return TD_ERR_ARG; }
    *out = nullptr;
    if (!(lambda >= 0.0 && lambda <= 1.0)) { set_error("lambda must lie in [0, 1]"); return TD_ERR_ARG; }
    td_treeset *ts = new (std::nothrow) td_treeset();
    if (!ts) { set_error("out of memory"); return TD_ERR_NOMEM; }
    int rc;
    try { rc = encode_kendall_colijn(newick, n_trees, lambda, n_threads, ts->host); }
    catch (const std::bad_alloc &) { set_error("out of memory while encoding"); rc = TD_ERR_NOMEM; }
    if (rc) { delete ts; return rc; }
    *out = ts;
    return TD_OK;
}

int td_encode_newick_joined(const char *buffer, int64_t total_bytes, int64_t n_trees, int n_threads, td_treeset **out)
{
    if (!buffer || total_bytes <= 0 || n_trees < 1) { set_error("empty buffer"); return TD_ERR_ARG; }
    if (buffer[total_bytes - 1] != 0) { set_error("the joined buffer must end with a NUL byte"); return TD_ERR_ARG; }
    std::vector<const char *> ptrs; ptrs.reserve((size_t)n_trees);
    const char *p = buffer, *end = buffer + total_bytes;
    while (p < end && (int64_t)ptrs.size() < n_trees) { ptrs.push_back(p); p = (const char *)memchr(p, 0, (size_t)(end - p)) + 1; }
    if ((int64_t)ptrs.size() != n_trees || p != end) { set_error("the buffer does not hold exactly %lld NUL-terminated strings (embedded NUL?)", (long long)n_trees); return TD_ERR_ARG; }
    return td_encode_newick(ptrs.data(), n_trees, n_threads, out);
}

int td_treeset_serialize(const td_treeset *ts, void *dst, int64_t cap_bytes, int64_t *need_bytes)
{
    if (!ts) { set_error("null treeset"); return TD_ERR_ARG; }
    try {
        const size_t need = serialize_set(ts->host, nullptr);
        if (need_bytes) *need_bytes = (int64_t)need;
        if (!dst) return TD_OK;
        if (cap_bytes < (int64_t)need) { set_error("buffer of %lld bytes is too small for the serialised set (%lld bytes)", (long long)cap_bytes, (long long)need); return TD_ERR_ARG; }
        std::vector<unsigned char> buf; buf.reserve(need);
        serialize_set(ts->host, &buf);
        memcpy(dst, buf.data(), buf.size());
    } catch (const std::bad_alloc &) { set_error("out of memory"); return TD_ERR_NOMEM; }
    return TD_OK;
}

int td_treeset_deserialize(const void *data, int64_t bytes, td_treeset **out)
{
    if (!out) { set_error("null output pointer"); return TD_ERR_ARG; }
    *out = nullptr;
    if (!data || bytes <= 0) { set_error("empty buffer"); return TD_ERR_ARG; }
    td_treeset *ts = new (std::nothrow) td_treeset();
    if (!ts) { set_error("out of memory"); return TD_ERR_NOMEM; }
    int rc;
    try { rc = deserialize_set(data, (size_t)bytes, ts->host); }
    catch (const std::bad_alloc &) { set_error("out of memory"); rc = TD_ERR_NOMEM; }
    if (rc) { delete ts; return rc; }
    *out = ts;
    return TD_OK;
}

void td_treeset_free(td_treeset *ts)
{
    if (!ts) return;
    DeviceScope scope;
    delete ts;
}

int td_treeset_info(const td_treeset *ts, td_info *info)
{
    if (!ts || !info) { set_error("null argument"); return TD_ERR_ARG; }
    const HostSet &h = ts->host;
    memset(info, 0, sizeof *info);
    info->n_trees = h.N; info->n_taxa = h.n_taxa; info->words = h.W; info->max_splits = h.max_splits; info->max_shared = h.max_shared;
    info->uniform = h.uniform; info->rooted = h.rooted; info->dict_size = h.dict_size; info->dict_shared = h.dict_shared;
    info->total_splits = h.total_splits; info->device = -1; info->n_postings = (int32_t)h.post_tree.size();
    info->heavy_rows = h.heavy_rows; info->mean_common = h.mean_common; info->heavy_common = h.heavy_common;
    info->n_groups = ts->groups_built ? (int32_t)ts->groups.size() : -1;
    for (DeviceSet *d : ts->devs) if (d) { info->device = d->device; break; }
    return TD_OK;
}

int td_treeset_get(const td_treeset *ts, int what, void *dst, int64_t cap_bytes, int64_t *need_bytes)
{
    if (!ts) { set_error("null treeset"); return TD_ERR_ARG; }
    const HostSet &h = ts->host;
    const void *src = nullptr; int64_t n = 0; std::string joined;
    switch (what) {
        case TD_GET_LABELS: for (size_t i = 0; i < h.labels.size(); i++) { if (i) joined.push_back('\n'); joined += h.labels[i]; }
                            src = joined.data(); n = (int64_t)joined.size(); break;
        case TD_GET_NSPLITS: src = h.n_splits.data(); n = h.n_splits.size() * sizeof(int32_t); break;
        case TD_GET_MASKS: src = h.masks.data(); n = h.masks.size() * sizeof(uint64_t); break;
        case TD_GET_LENGTHS: src = h.lens.data(); n = h.lens.size() * sizeof(double); break;
        case TD_GET_LEAFLEN: src = h.leaflen.data(); n = (int64_t)h.N * h.n_taxa * sizeof(double); break;
        case TD_GET_IDS: src = h.ids.data(); n = h.ids.size() * sizeof(uint32_t); break;
        case TD_GET_LEAFMASK: src = h.leafmask.data(); n = h.leafmask.size() * sizeof(uint64_t); break;
        case TD_GET_ROOTED: src = h.tree_rooted.data(); n = h.tree_rooted.size() * sizeof(int32_t); break;
        default: set_error("unknown TD_GET_* selector %d", what); return TD_ERR_ARG;
    }
    if (need_bytes) *need_bytes = n;
    if (dst && cap_bytes > 0) memcpy(dst, src, (size_t)std::min<int64_t>(n, cap_bytes));
    return TD_OK;
}

int td_newick_labels(const char *newick, char *dst, int64_t cap_bytes, int64_t *need_bytes, int *rooted)
{
    if (!newick) { set_error("null newick"); return TD_ERR_ARG; }
    std::vector<std::string> labels; int r = 0;
    int rc = newick_labels(newick, labels, r); if (rc) return rc;
    std::string joined;
    for (size_t i = 0; i < labels.size(); i++) { if (i) joined.push_back('\n'); joined += labels[i]; }
    if (need_bytes) *need_bytes = (int64_t)joined.size() + 1;
    if (rooted) *rooted = r;
    if (dst && cap_bytes > 0) { size_t n = std::min<size_t>(joined.size(), (size_t)cap_bytes - 1); memcpy(dst, joined.data(), n); dst[n] = 0; }
    return TD_OK;
}

int td_newick_prune(const char *newick, const char *const *keep, int64_t n_keep, char *dst, int64_t cap_bytes, int64_t *need_bytes)
{
    if (!newick || (!keep && n_keep > 0)) { set_error("null argument"); return TD_ERR_ARG; }
    std::vector<std::string> k; for (int64_t i = 0; i < n_keep; i++) k.emplace_back(keep[i]);
    std::string out; int rc = newick_prune(newick, k, out); if (rc) return rc;
    if (need_bytes) *need_bytes = (int64_t)out.size() + 1;
    if (dst && cap_bytes > 0) { size_t n = std::min<size_t>(out.size(), (size_t)cap_bytes - 1); memcpy(dst, out.data(), n); dst[n] = 0; }
    return TD_OK;
}

int td_device_count(int *count)
{
    int c = 0; cudaError_t e = cudaGetDeviceCount(&c);
    if (count) *count = (e == cudaSuccess) ? c : 0;
    if (e != cudaSuccess) { set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return TD_ERR_NO_DEVICE; }
    return TD_OK;
}

int td_upload(td_treeset *ts, int device, void *stream)
{
    if (!ts) { set_error("null treeset"); return TD_ERR_ARG; }
    DeviceScope scope;
    return upload(ts, device, (cudaStream_t)stream);
}

int64_t td_condensed_offset(int64_t n_trees, int64_t row) { return condensed_offset(n_trees, row); }

int td_pairwise_device(td_treeset *ts, uint32_t metric_mask, int normalise, int min_overlap, double overlap_fail_value,
                       int64_t row_begin, int64_t row_end, void *const d_out[4], void *stream)
{
    if (!ts) { set_error("null treeset"); return TD_ERR_ARG; }
    DeviceScope scope;
    const int64_t N = ts->host.N;
    if (!(metric_mask & 15u) || (metric_mask & ~15u)) { set_error("metric_mask 0x%x: expected a combination of TD_MASK(TD_RF..TD_GEO)", metric_mask); return TD_ERR_ARG; }
    if (row_begin < 0 || row_end > N || row_begin > row_end) { set_error("row range [%lld,%lld) outside [0,%lld]", (long long)row_begin, (long long)row_end, (long long)N); return TD_ERR_ARG; }
    if (!d_out) { set_error("null output table"); return TD_ERR_ARG; }
    int device, rc; DeviceSet *d;
    if ((rc = device_of(d_out, metric_mask, &device))) return rc;
    if ((rc = check_ready(ts, device, &d))) return rc;
    return run_pairwise(ts, d, metric_mask, normalise, min_overlap, overlap_fail_value, false, 0, row_begin, row_end, d_out, (cudaStream_t)stream);
}

int td_pairwise_device2(td_treeset *ts, uint32_t metric_mask, int normalise, int min_overlap, double overlap_fail_value,
                        const int64_t row_begin[2], const int64_t row_end[2], void *const d_out_a[4], void *const d_out_b[4], void *stream)
{
    if (!ts || !row_begin || !row_end || !d_out_a || !d_out_b) { set_error("null argument"); return TD_ERR_ARG; }
    DeviceScope scope;
    const int64_t N = ts->host.N;
    if (!(metric_mask & 15u) || (metric_mask & ~15u)) { set_error("metric_mask 0x%x: expected a combination of TD_MASK(TD_RF..TD_GEO)", metric_mask); return TD_ERR_ARG; }
    for (int b = 0; b < 2; b++)
        if (row_begin[b] < 0 || row_end[b] > N || row_begin[b] > row_end[b]) { set_error("row range [%lld,%lld) outside [0,%lld]", (long long)row_begin[b], (long long)row_end[b], (long long)N); return TD_ERR_ARG; }
    if (row_end[0] > row_begin[1] && row_end[1] > row_begin[0]) { set_error("the two row blocks overlap"); return TD_ERR_ARG; }
    int device, dev_b, rc; DeviceSet *d;
    if ((rc = device_of(d_out_a, metric_mask, &device))) return rc;
    if ((rc = device_of(d_out_b, metric_mask, &dev_b))) return rc;
    if (device != dev_b) { set_error("output pointers live on different GPUs"); return TD_ERR_ARG; }
    if ((rc = check_ready(ts, device, &d))) return rc;
    if (row_end[0] <= row_begin[0])
        return run_pairwise(ts, d, metric_mask, normalise, min_overlap, overlap_fail_value, false, 0, row_begin[1], row_end[1], d_out_b, (cudaStream_t)stream);
    const RowBlock second{row_begin[1], row_end[1], d_out_b};
    return run_pairwise(ts, d, metric_mask, normalise, min_overlap, overlap_fail_value, false, 0, row_begin[0], row_end[0], d_out_a, (cudaStream_t)stream, &second);
}

int td_pairwise_rect_device(td_treeset *ts, int64_t split, uint32_t metric_mask, int normalise, void *const d_out[4], void *stream)
{
    if (!ts) { set_error("null treeset"); return TD_ERR_ARG; }
    DeviceScope scope;
    const int64_t N = ts->host.N;
    if (split <= 0 || split >= N) { set_error("split %lld must be inside (0,%lld)", (long long)split, (long long)N); return TD_ERR_ARG; }
    if (!(metric_mask & 15u) || (metric_mask & ~15u) || !d_out) { set_error("bad metric_mask / output table"); return TD_ERR_ARG; }
    int device, rc; DeviceSet *d;
    if ((rc = device_of(d_out, metric_mask, &device))) return rc;
    if ((rc = check_ready(ts, device, &d))) return rc;
    return run_pairwise(ts, d, metric_mask, normalise, 4, 0.0, true, split, 0, split, d_out, (cudaStream_t)stream);
}

// ---- host-buffer calls: row bands, two in flight, kernels of band k under the device-to-host copy of band k - 1 -------------------------
namespace {

// bytes of one band per metric (TREEDIST_BAND_BYTES, default 32 MiB): small enough that the copy of the first band starts early and
// two bands stay L2 / HBM friendly, large enough that a band is thousands of tiles
size_t band_bytes_target()
{
    const char *e = getenv("TREEDIST_BAND_BYTES");
    if (e && *e) { char *end = nullptr; const unsigned long long x = strtoull(e, &end, 10); if (end != e && x >= (1u << 16)) return (size_t)x; }
    return (size_t)(32u << 20);
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

// pageable destination: the band went to a pinned staging block first, the host moves it on (a few threads: one core does ~10 GB/s)
void host_copy(void *dst, const void *src, size_t bytes)
{
    constexpr size_t kChunk = 8u << 20;
    const int parts = (int)std::min<size_t>(4, bytes / kChunk);
    if (parts <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = (bytes / parts + 63) & ~(size_t)63;
    for (int t = 1; t < parts; t++) {
        const size_t o = (size_t)t * per, n = t == parts - 1 ? bytes - o : per;
        th.emplace_back([=] { memcpy((char *)dst + o, (const char *)src + o, n); });
    }
    memcpy(dst, src, per);
    for (auto &x : th) x.join();
}

// rows [row_begin, row_end) cut into bands of about `target` pairs at multiples of 64 rows (whole tile rows: no tile is computed twice)
std::vector<int64_t> plan_bands(int64_t N, int64_t row_begin, int64_t row_end, int64_t target)
{
    std::vector<int64_t> cuts{row_begin};
    const int64_t total = condensed_offset(N, row_end) - condensed_offset(N, row_begin);
    if (total > target + target / 2) {
        int64_t r = row_begin;
        while (r < row_end) {
            const int64_t base = condensed_offset(N, r);
            int64_t e = (r / 64 + 1) * 64;
            while (e < row_end && condensed_offset(N, std::min(e, row_end)) - base < target) e += 64;
            e = std::min(e, row_end);
            if (condensed_offset(N, row_end) - condensed_offset(N, e) < target / 4) e = row_end;      // no tiny last band
            cuts.push_back(e); r = e;
        }
    } else cuts.push_back(row_end);
    return cuts;
}

struct HostCall {                       // everything a host-buffer call owns; released on every exit path
    int device = -1;
    BandStreams bs;
    void *dbuf[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    size_t dsize[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    void *stage[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    void *extra = nullptr; size_t extra_size = 0;
    ~HostCall()
    {
        if (bs.compute) { cudaStreamSynchronize(bs.compute); cudaStreamSynchronize(bs.copy); }
        for (int b = 0; b < 2; b++) for (int m = 0; m < 4; m++) { if (dbuf[b][m]) pool_free(device, dbuf[b][m], dsize[b][m]); pinned_free(stage[b][m]); }
        if (extra) pool_free(device, extra, extra_size);
        release_streams(device, bs);
    }
};

}  // namespace

// rows [row_begin, row_end) of the requested matrices into HOST buffers.  u16_rf: metric_mask is RF alone and out[TD_RF] is uint16
// (the int8 tcgen05 incidence contraction with its packed epilogue, BASELINE config 4); otherwise all outputs are double.
static int host_pairwise(td_treeset *ts, uint32_t metric_mask, int normalise, int min_overlap, double overlap_fail_value,
                         int64_t row_begin, int64_t row_end, void *const out[4], bool u16_rf, int device)
{
    if (!ts || !out) { set_error("null argument"); return TD_ERR_ARG; }
    DeviceScope scope;
    int rc;
    const size_t esz = u16_rf ? sizeof(uint16_t) : sizeof(double);
    if (!(metric_mask & 15u) || (metric_mask & ~15u)) { set_error("metric_mask 0x%x: expected a combination of TD_MASK(TD_RF..TD_GEO)", metric_mask); return TD_ERR_ARG; }
    if ((rc = upload(ts, device, 0))) return rc;        // no-op when already resident
    DeviceSet *dset = resident(ts, device);
    const int64_t N = ts->host.N;
    if (row_begin < 0 || row_end > N || row_begin > row_end) { set_error("row range outside [0,N]"); return TD_ERR_ARG; }
    const int64_t pairs = condensed_offset(N, row_end) - condensed_offset(N, row_begin);
    if (pairs <= 0) return TD_OK;
    for (int m = 0; m < 4; m++) if ((metric_mask >> m & 1u) && !out[m]) { set_error("output pointer for metric %d is null", m); return TD_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(device));

    const std::vector<int64_t> cuts = plan_bands(N, row_begin, row_end, (int64_t)(band_bytes_target() / esz));
    const int n_bands = (int)cuts.size() - 1;
    int64_t max_band = 0;
    for (int k = 0; k < n_bands; k++) max_band = std::max(max_band, condensed_offset(N, cuts[k + 1]) - condensed_offset(N, cuts[k]));
    HostCall hc; hc.device = device;
    bool pageable[4] = {false, false, false, false}, any_pageable = false;
    for (int m = 0; m < 4; m++) if (metric_mask >> m & 1u) {
        pageable[m] = !is_pinned(out[m]); any_pageable |= pageable[m];
        for (int b = 0; b < std::min(n_bands, 2); b++) {
            cudaError_t e = pool_alloc(device, (size_t)max_band * esz, &hc.dbuf[b][m], &hc.dsize[b][m]);
            if (e != cudaSuccess) { set_error("cudaMalloc of %lld bytes failed: %s", (long long)(max_band * esz), cudaGetErrorString(e)); return TD_ERR_NOMEM; }
            if (pageable[m]) {
                e = pinned_alloc((size_t)max_band * esz, &hc.stage[b][m]);
                if (e != cudaSuccess) { set_error("cudaHostAlloc of %lld bytes failed: %s", (long long)(max_band * esz), cudaGetErrorString(e)); return TD_ERR_NOMEM; }
            }
        }
    }
    CUDA_TRY(acquire_streams(device, &hc.bs));
    const BandStreams &bs = hc.bs;
    const int64_t p_first = condensed_offset(N, row_begin);
    auto drain = [&](int k) -> int {                       // host side of band k: wait for its copy, move pageable outputs on
        CUDA_TRY(cudaEventSynchronize(bs.copied[k & 1]));
        const int64_t p0 = condensed_offset(N, cuts[k]) - p_first, n = condensed_offset(N, cuts[k + 1]) - condensed_offset(N, cuts[k]);
        for (int m = 0; m < 4; m++) if ((metric_mask >> m & 1u) && pageable[m]) host_copy((char *)out[m] + (size_t)p0 * esz, hc.stage[k & 1][m], (size_t)n * esz);
        return TD_OK;
    };
    for (int k = 0; k < n_bands; k++) {
        const int b = k & 1;
        const int64_t p0 = condensed_offset(N, cuts[k]) - p_first, n = condensed_offset(N, cuts[k + 1]) - condensed_offset(N, cuts[k]);
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(bs.compute, bs.copied[b], 0));        // the band buffers are free again
        if (u16_rf) rc = run_rf_incidence(ts, dset, 0, cuts[k], cuts[k + 1], hc.dbuf[b][TD_RF], 1, bs.compute);
        else rc = run_pairwise(ts, dset, metric_mask, normalise, min_overlap, overlap_fail_value, false, 0, cuts[k], cuts[k + 1], hc.dbuf[b], bs.compute);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(bs.done[b], bs.compute));
        CUDA_TRY(cudaStreamWaitEvent(bs.copy, bs.done[b], 0));
        for (int m = 0; m < 4; m++) if (metric_mask >> m & 1u)
            CUDA_TRY(cudaMemcpyAsync(pageable[m] ? hc.stage[b][m] : (void *)((char *)out[m] + (size_t)p0 * esz), hc.dbuf[b][m], (size_t)n * esz, cudaMemcpyDeviceToHost, bs.copy));
        CUDA_TRY(cudaEventRecord(bs.copied[b], bs.copy));
        if (k >= 1 && any_pageable && (rc = drain(k - 1))) return rc;                   // under the kernels / copy of band k
    }
    if (any_pageable) { if ((rc = drain(n_bands - 1))) return rc; }
    {
        cudaError_t e = cudaStreamSynchronize(bs.copy);
        if (e == cudaSuccess) e = cudaStreamSynchronize(bs.compute);
        if (e != cudaSuccess) { set_error("kernel execution failed: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
    }
    return take_error_flag(ts, dset);
}

int td_pairwise(td_treeset *ts, uint32_t metric_mask, int normalise, int min_overlap, double overlap_fail_value,
                int64_t row_begin, int64_t row_end, double *const out[4], int device)
{
    if (!out) { set_error("null argument"); return TD_ERR_ARG; }
    void *const o[4] = {out[0], out[1], out[2], out[3]};
    return host_pairwise(ts, metric_mask, normalise, min_overlap, overlap_fail_value, row_begin, row_end, o, false, device);
}

int td_rf_incidence(td_treeset *ts, int64_t row_begin, int64_t row_end, uint16_t *out, int device)
{
    if (!ts || !out) { set_error("null argument"); return TD_ERR_ARG; }
    const HostSet &h = ts->host;
    if (!h.uniform) { set_error("the incidence path needs one shared leaf set"); return TD_ERR_UNSUPPORTED; }
    if (h.max_splits > 32767) { set_error("too many splits per tree for the u16 epilogue"); return TD_ERR_UNSUPPORTED; }
    void *const o[4] = {out, nullptr, nullptr, nullptr};
    return host_pairwise(ts, TD_MASK(TD_RF), 0, 4, 0.0, row_begin, row_end, o, true, device);
}

int td_pairwise_square(td_treeset *ts, int metric, int normalise, int min_overlap, double overlap_fail_value, double *out, int device)
{
    if (!ts || !out) { set_error("null argument"); return TD_ERR_ARG; }
    if (metric < 0 || metric > 3) { set_error("unknown metric %d", metric); return TD_ERR_ARG; }
    DeviceScope scope;
    int rc;
    if ((rc = upload(ts, device, 0))) return rc;
    DeviceSet *dset = resident(ts, device);
    const int64_t N = ts->host.N, pairs = N * (N - 1) / 2;
    if (N == 1) { out[0] = 0.0; return TD_OK; }
    CUDA_TRY(cudaSetDevice(device));
    HostCall hc; hc.device = device;
    {
        cudaError_t e = pool_alloc(device, (size_t)pairs * sizeof(double), &hc.extra, &hc.extra_size);
        if (e != cudaSuccess) { set_error("cudaMalloc of %lld bytes failed: %s", (long long)pairs * 8, cudaGetErrorString(e)); return TD_ERR_NOMEM; }
    }
    // bands of whole rows of the square matrix
    const int64_t rows_per_band = std::max<int64_t>(32, (int64_t)(band_bytes_target() / sizeof(double)) / N / 32 * 32);
    const int n_bands = (int)((N + rows_per_band - 1) / rows_per_band);
    const bool pageable = !is_pinned(out);
    for (int b = 0; b < std::min(n_bands, 2); b++) {
        cudaError_t e = pool_alloc(device, (size_t)rows_per_band * N * sizeof(double), &hc.dbuf[b][0], &hc.dsize[b][0]);
        if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return TD_ERR_NOMEM; }
        if (pageable && (e = pinned_alloc((size_t)rows_per_band * N * sizeof(double), &hc.stage[b][0])) != cudaSuccess) { set_error("cudaHostAlloc failed: %s", cudaGetErrorString(e)); return TD_ERR_NOMEM; }
    }
    CUDA_TRY(acquire_streams(device, &hc.bs));
    const BandStreams &bs = hc.bs;
    void *d_out[4] = {nullptr, nullptr, nullptr, nullptr};
    d_out[metric] = hc.extra;
    if ((rc = run_pairwise(ts, dset, 1u << metric, normalise, min_overlap, overlap_fail_value, false, 0, 0, N, d_out, bs.compute))) return rc;
    auto drain = [&](int k) -> int {
        CUDA_TRY(cudaEventSynchronize(bs.copied[k & 1]));
        const int64_t r0 = (int64_t)k * rows_per_band, r1 = std::min(N, r0 + rows_per_band);
        if (pageable) host_copy(out + r0 * N, hc.stage[k & 1][0], (size_t)(r1 - r0) * N * sizeof(double));
        return TD_OK;
    };
    for (int k = 0; k < n_bands; k++) {
        const int b = k & 1;
        const int64_t r0 = (int64_t)k * rows_per_band, r1 = std::min(N, r0 + rows_per_band);
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(bs.compute, bs.copied[b], 0));
        CUDA_TRY(launch_square_rows((const double *)hc.extra, (double *)hc.dbuf[b][0], N, r0, r1, bs.compute));
        launch_counter_add(1);
        CUDA_TRY(cudaEventRecord(bs.done[b], bs.compute));
        CUDA_TRY(cudaStreamWaitEvent(bs.copy, bs.done[b], 0));
        CUDA_TRY(cudaMemcpyAsync(pageable ? hc.stage[b][0] : (void *)(out + r0 * N), hc.dbuf[b][0], (size_t)(r1 - r0) * N * sizeof(double), cudaMemcpyDeviceToHost, bs.copy));
        CUDA_TRY(cudaEventRecord(bs.copied[b], bs.copy));
        if (k >= 1 && pageable && (rc = drain(k - 1))) return rc;
    }
    if (pageable && (rc = drain(n_bands - 1))) return rc;
    {
        cudaError_t e = cudaStreamSynchronize(bs.copy);
        if (e == cudaSuccess) e = cudaStreamSynchronize(bs.compute);
        if (e != cudaSuccess) { set_error("kernel execution failed: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
    }
    return take_error_flag(ts, dset);
}

// ---- classical MDS of a device-resident matrix --------------------------------------------------------------------------------
namespace {
// cond: condensed matrix on `device` (rows 0..n); coords: n x dims on the device.  Synchronous.
// what to do with the square matrix once it is on the device
struct EmbedJob {
    int kind = 0;                     // 0: classical MDS; 1: spectral embedding (Spectral.spectral_embedding_); 2: affinity matrix only
    int prune_k = 0, prune_and = 0, scale_k = 0;
    double *affinity_host = nullptr;  // kind 2: n x n result (host)
};

int cmds_on_device(int device, const double *cond, const double *square_host, int64_t n, int dims, double *d_coords, double *eigenvalues, int max_iter,
                   double tol, int *iterations, double *residual, cudaStream_t st, const EmbedJob &job = EmbedJob())
{
    if (job.kind != 2 && (n < 2 || dims < 1 || dims > cmds_block() - 4 || dims >= n)) { set_error("embedding: need 1 <= dimensions <= %d and dimensions < n_trees", cmds_block() - 4); return TD_ERR_ARG; }
    if (job.kind != 0 && (job.prune_k < 0 || job.prune_k >= n || job.scale_k < 0 || job.scale_k >= n)) { set_error("spectral: pruning / scale neighbour counts must lie in [0, n_trees)"); return TD_ERR_ARG; }
    if (max_iter <= 0) max_iter = 5000;
    if (!(tol > 0.0)) tol = 1e-12;
    struct Own { int device; void *sq = nullptr, *work = nullptr, *pin = nullptr; size_t sq_got = 0, work_got = 0;
                 ~Own() { if (sq) pool_free(device, sq, sq_got); if (work) pool_free(device, work, work_got); pinned_free(pin); } } own{device};
    cudaError_t e = pool_alloc(device, (size_t)n * n * sizeof(double), &own.sq, &own.sq_got);
    if (e == cudaSuccess) e = pool_alloc(device, cmds_workspace_doubles(n) * sizeof(double), &own.work, &own.work_got);
    if (e != cudaSuccess) { set_error("cudaMalloc of the %lld x %lld matrix failed: %s", (long long)n, (long long)n, cudaGetErrorString(e)); return TD_ERR_NOMEM; }
    if ((e = pinned_alloc(3 * 16 * 16 * sizeof(double) + 64, &own.pin)) != cudaSuccess) { set_error("cudaHostAlloc failed: %s", cudaGetErrorString(e)); return TD_ERR_NOMEM; }
    if (cond) { CUDA_TRY(launch_square_rows(cond, (double *)own.sq, n, 0, n, st)); launch_counter_add(1); }
    else CUDA_TRY(cudaMemcpyAsync(own.sq, square_host, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (job.kind == 0)
        CUDA_TRY(run_cmds((double *)own.sq, n, dims, (double *)own.work, (double *)own.pin, d_coords, eigenvalues, max_iter, tol, iterations, residual, st));
    else if (job.kind == 1)
        CUDA_TRY(run_spectral((double *)own.sq, n, dims, job.prune_k, job.prune_and, job.scale_k, (double *)own.work, (double *)own.pin, d_coords, eigenvalues,
                              max_iter, tol, iterations, residual, st));
    else {
        // Spectral.decompose: the affinity matrix with a unit diagonal (treeCl/clustering.py:171-228), back to the host
        CUDA_TRY(run_affinity((double *)own.sq, n, job.prune_k, job.prune_and, job.scale_k, 1.0, (double *)own.work, st));
        CUDA_TRY(cudaMemcpyAsync(job.affinity_host, own.sq, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return TD_OK;
}
}  // namespace

int td_cmds_device(const double *d_condensed, int64_t n_trees, int dimensions, double *d_coords, double *eigenvalues, int max_iter, double tol,
                   int *iterations, double *residual, void *stream)
{
    if (!d_condensed || !d_coords || !eigenvalues) { set_error("null argument"); return TD_ERR_ARG; }
    DeviceScope scope;
    cudaPointerAttributes attr;
    CUDA_TRY(cudaPointerGetAttributes(&attr, d_condensed));
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) { set_error("the condensed matrix is not device memory"); return TD_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(attr.device));
    return cmds_on_device(attr.device, d_condensed, nullptr, n_trees, dimensions, d_coords, eigenvalues, max_iter, tol, iterations, residual, (cudaStream_t)stream);
}

int td_cmds(td_treeset *ts, int metric, int normalise, int min_overlap, double overlap_fail_value, int dimensions, double *coords, double *eigenvalues,
            int max_iter, double tol, int *iterations, double *residual, int device)
{
    if (!ts || !coords || !eigenvalues) { set_error("null argument"); return TD_ERR_ARG; }
    if (metric < 0 || metric > 3) { set_error("unknown metric %d", metric); return TD_ERR_ARG; }
    DeviceScope scope;
    int rc;
    if ((rc = upload(ts, device, 0))) return rc;
    DeviceSet *dset = resident(ts, device);
    const int64_t N = ts->host.N, pairs = N * (N - 1) / 2;
    CUDA_TRY(cudaSetDevice(device));
    HostCall hc; hc.device = device;
    cudaError_t e = pool_alloc(device, (size_t)std::max<int64_t>(pairs, 1) * sizeof(double), &hc.extra, &hc.extra_size);
    if (e == cudaSuccess) e = pool_alloc(device, (size_t)N * dimensions * sizeof(double) + 64, &hc.dbuf[0][0], &hc.dsize[0][0]);
    if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return TD_ERR_NOMEM; }
    CUDA_TRY(acquire_streams(device, &hc.bs));
    void *d_out[4] = {nullptr, nullptr, nullptr, nullptr};
    d_out[metric] = hc.extra;
    if ((rc = run_pairwise(ts, dset, 1u << metric, normalise, min_overlap, overlap_fail_value, false, 0, 0, N, d_out, hc.bs.compute))) return rc;
    if ((rc = cmds_on_device(device, (const double *)hc.extra, nullptr, N, dimensions, (double *)hc.dbuf[0][0], eigenvalues, max_iter, tol, iterations, residual, hc.bs.compute))) return rc;
    CUDA_TRY(cudaMemcpyAsync(coords, hc.dbuf[0][0], (size_t)N * dimensions * sizeof(double), cudaMemcpyDeviceToHost, hc.bs.compute));
    CUDA_TRY(cudaStreamSynchronize(hc.bs.compute));
    return take_error_flag(ts, dset);
}

// shared body of the host-matrix entry points (td_cmds_host, td_spectral_embedding_host, td_affinity_host)
static int embed_host(const double *square, int64_t n_trees, int dimensions, double *coords, double *eigenvalues, int max_iter, double tol, int *iterations,
                      double *residual, int device, const EmbedJob &job)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); set_error("no CUDA device available; libtreedist_cuda has no CPU fallback"); return TD_ERR_NO_DEVICE; }
    if (device < 0 || device >= count) { set_error("device %d out of range", device); return TD_ERR_ARG; }
    if (n_trees < 1) { set_error("empty matrix"); return TD_ERR_ARG; }
    DeviceScope scope;
    CUDA_TRY(cudaSetDevice(device));
    HostCall hc; hc.device = device;
    cudaError_t e = pool_alloc(device, (size_t)n_trees * std::max(dimensions, 1) * sizeof(double) + 64, &hc.dbuf[0][0], &hc.dsize[0][0]);
    if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return TD_ERR_NOMEM; }
    CUDA_TRY(acquire_streams(device, &hc.bs));
    int rc;
    if ((rc = cmds_on_device(device, nullptr, square, n_trees, dimensions, (double *)hc.dbuf[0][0], eigenvalues, max_iter, tol, iterations, residual, hc.bs.compute, job))) return rc;
    if (job.kind != 2) {
        CUDA_TRY(cudaMemcpyAsync(coords, hc.dbuf[0][0], (size_t)n_trees * dimensions * sizeof(double), cudaMemcpyDeviceToHost, hc.bs.compute));
        CUDA_TRY(cudaStreamSynchronize(hc.bs.compute));
    }
    return TD_OK;
}

int td_cmds_host(const double *square, int64_t n_trees, int dimensions, double *coords, double *eigenvalues, int max_iter, double tol, int *iterations,
                 double *residual, int device)
{
    if (!square || !coords || !eigenvalues) { set_error("null argument"); return TD_ERR_ARG; }
    return embed_host(square, n_trees, dimensions, coords, eigenvalues, max_iter, tol, iterations, residual, device, EmbedJob());
}

int td_spectral_embedding_host(const double *square, int64_t n_trees, int dimensions, int prune_k, int prune_and, int scale_k, double *coords,
                               double *eigenvalues, int max_iter, double tol, int *iterations, double *residual, int device)
{
    if (!square || !coords || !eigenvalues) { set_error("null argument"); return TD_ERR_ARG; }
    EmbedJob job; job.kind = 1; job.prune_k = prune_k; job.prune_and = prune_and; job.scale_k = scale_k;
    return embed_host(square, n_trees, dimensions, coords, eigenvalues, max_iter, tol, iterations, residual, device, job);
}

int td_affinity_host(const double *square, int64_t n_trees, int prune_k, int prune_and, int scale_k, double *affinity, int device)
{
    if (!square || !affinity) { set_error("null argument"); return TD_ERR_ARG; }
    EmbedJob job; job.kind = 2; job.prune_k = prune_k; job.prune_and = prune_and; job.scale_k = scale_k; job.affinity_host = affinity;
    double dummy = 0.0;
    return embed_host(square, n_trees, 1, &dummy, &dummy, 0, 0.0, nullptr, nullptr, device, job);
}

int td_spectral_embedding(td_treeset *ts, int metric, int normalise, int dimensions, int prune_k, int prune_and, int scale_k, double *coords,
                          double *eigenvalues, int max_iter, double tol, int *iterations, double *residual, int device)
{
    if (!ts || !coords || !eigenvalues) { set_error("null argument"); return TD_ERR_ARG; }
    if (metric < 0 || metric > 3) { set_error("unknown metric %d", metric); return TD_ERR_ARG; }
    DeviceScope scope;
    int rc;
    if ((rc = upload(ts, device, 0))) return rc;
    DeviceSet *dset = resident(ts, device);
    const int64_t N = ts->host.N, pairs = N * (N - 1) / 2;
    CUDA_TRY(cudaSetDevice(device));
    HostCall hc; hc.device = device;
    cudaError_t e = pool_alloc(device, (size_t)std::max<int64_t>(pairs, 1) * sizeof(double), &hc.extra, &hc.extra_size);
    if (e == cudaSuccess) e = pool_alloc(device, (size_t)N * std::max(dimensions, 1) * sizeof(double) + 64, &hc.dbuf[0][0], &hc.dsize[0][0]);
    if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return TD_ERR_NOMEM; }
    CUDA_TRY(acquire_streams(device, &hc.bs));
    void *d_out[4] = {nullptr, nullptr, nullptr, nullptr};
    d_out[metric] = hc.extra;
    if ((rc = run_pairwise(ts, dset, 1u << metric, normalise, 4, 0.0, false, 0, 0, N, d_out, hc.bs.compute))) return rc;
    EmbedJob job; job.kind = 1; job.prune_k = prune_k; job.prune_and = prune_and; job.scale_k = scale_k;
    if ((rc = cmds_on_device(device, (const double *)hc.extra, nullptr, N, dimensions, (double *)hc.dbuf[0][0], eigenvalues, max_iter, tol, iterations, residual, hc.bs.compute, job))) return rc;
    CUDA_TRY(cudaMemcpyAsync(coords, hc.dbuf[0][0], (size_t)N * dimensions * sizeof(double), cudaMemcpyDeviceToHost, hc.bs.compute));
    CUDA_TRY(cudaStreamSynchronize(hc.bs.compute));
    return take_error_flag(ts, dset);
}

int td_alloc_pinned(int64_t bytes, void **out)
{
    if (!out || bytes < 0) { set_error("bad argument"); return TD_ERR_ARG; }
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); set_error("no CUDA device available: pinned host memory needs one"); return TD_ERR_NO_DEVICE; }
    cudaError_t e = pinned_alloc((size_t)bytes, out);
    if (e != cudaSuccess) { set_error("cudaHostAlloc of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e)); return TD_ERR_NOMEM; }
    return TD_OK;
}

int td_free_pinned(void *p)
{
    if (!pinned_free(p)) { set_error("pointer was not allocated by td_alloc_pinned"); return TD_ERR_ARG; }
    return TD_OK;
}

int td_check_errors(td_treeset *ts, int device)
{
    if (!ts) { set_error("null treeset"); return TD_ERR_ARG; }
    DeviceScope scope;
    DeviceSet *d; int rc;
    if ((rc = check_ready(ts, device, &d))) return rc;
    return take_error_flag(ts, d);
}

int td_trim_cache(int device)
{
    encoder_trim();
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return TD_OK; }      // no GPU: nothing is cached
    DeviceScope scope;
    for (int dev = 0; dev < count && dev < 64; dev++) {
        if (device >= 0 && dev != device) continue;
        if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); continue; }
        pool_trim(dev);
    }
    if (device < 0) pinned_trim();
    return TD_OK;
}

int td_pairwise_rect(td_treeset *query, td_treeset *reference, uint32_t metric_mask, int normalise, double *const out[4], int device)
{
    if (!query || !reference || !out) { set_error("null argument"); return TD_ERR_ARG; }
    if (!(metric_mask & 15u) || (metric_mask & ~15u)) { set_error("metric_mask 0x%x: expected a combination of TD_MASK(TD_RF..TD_GEO)", metric_mask); return TD_ERR_ARG; }
    for (int m = 0; m < 4; m++) if ((metric_mask >> m & 1u) && !out[m]) { set_error("output pointer for metric %d is null", m); return TD_ERR_ARG; }
    DeviceScope scope;
    // one set with a joint split dictionary: query trees first, then the reference trees
    td_treeset *both = new (std::nothrow) td_treeset();
    if (!both) { set_error("out of memory"); return TD_ERR_NOMEM; }
    struct Free { td_treeset *t; ~Free() { delete t; } } own{both};
    int rc;
    try { rc = concat_sets(query->host, reference->host, 0, both->host); }
    catch (const std::bad_alloc &) { set_error("out of memory"); rc = TD_ERR_NOMEM; }
    if (rc) return rc;
    if ((rc = upload(both, device, 0))) return rc;
    DeviceSet *dset = resident(both, device);
    const int64_t nq = query->host.N, nr = reference->host.N, cells = nq * nr;
    CUDA_TRY(cudaSetDevice(device));
    HostCall hc; hc.device = device;
    void *d_out[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int m = 0; m < 4; m++) if (metric_mask >> m & 1u) {
        cudaError_t e = pool_alloc(device, (size_t)cells * sizeof(double), &hc.dbuf[0][m], &hc.dsize[0][m]);
        if (e != cudaSuccess) { set_error("cudaMalloc of %lld bytes failed: %s", (long long)cells * 8, cudaGetErrorString(e)); return TD_ERR_NOMEM; }
        d_out[m] = hc.dbuf[0][m];
    }
    CUDA_TRY(acquire_streams(device, &hc.bs));
    if ((rc = run_pairwise(both, dset, metric_mask, normalise, 4, 0.0, true, nq, 0, nq, d_out, hc.bs.compute))) return rc;
    for (int m = 0; m < 4; m++) if (metric_mask >> m & 1u)
        CUDA_TRY(cudaMemcpyAsync(out[m], d_out[m], (size_t)cells * sizeof(double), cudaMemcpyDeviceToHost, hc.bs.compute));
    CUDA_TRY(cudaStreamSynchronize(hc.bs.compute));
    return take_error_flag(both, dset);
}

int td_rf_incidence_device(td_treeset *ts, int normalise, int64_t row_begin, int64_t row_end, void *d_out, int out_is_u16, void *stream)
{
    if (!ts || !d_out) { set_error("null argument"); return TD_ERR_ARG; }
    DeviceScope scope;
    const HostSet &h = ts->host;
    if (!h.uniform) { set_error("the incidence path needs one shared leaf set"); return TD_ERR_UNSUPPORTED; }
    if (row_begin < 0 || row_end > h.N || row_begin > row_end) { set_error("row range outside [0,N]"); return TD_ERR_ARG; }
    if (out_is_u16 && normalise) { set_error("normalised RF needs the double output"); return TD_ERR_ARG; }
    if (h.max_splits > 32767) { set_error("too many splits per tree for the u16 / int32 epilogue"); return TD_ERR_UNSUPPORTED; }
    cudaPointerAttributes attr;
    CUDA_TRY(cudaPointerGetAttributes(&attr, d_out));
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) { set_error("output pointer is not device memory"); return TD_ERR_ARG; }
    DeviceSet *d; int rc;
    if ((rc = check_ready(ts, attr.device, &d))) return rc;
    return run_rf_incidence(ts, d, normalise, row_begin, row_end, d_out, out_is_u16, (cudaStream_t)stream);
}

int td_geo_get_stats(td_treeset *ts, td_geo_stats *out)
{
    if (!ts || !out) { set_error("null argument"); return TD_ERR_ARG; }
    memset(out, 0, sizeof *out);
    std::vector<DeviceSet *> devs;
    { std::lock_guard<std::mutex> guard(ts->mu); devs = ts->devs; }
    for (DeviceSet *d : devs) {
        if (!d) continue;
        unsigned long long s[4] = {0, 0, 0, 0};
        CUDA_TRY(cudaSetDevice(d->device));
        CUDA_TRY(cudaMemcpy(s, d->stats, sizeof s, cudaMemcpyDeviceToHost));
        out->pairs += s[0]; out->maxflow_solves += s[1]; out->ratios += s[2]; out->augmentations += s[3];
        CUDA_TRY(cudaMemset(d->stats, 0, sizeof s));
    }
    return TD_OK;
}

}  // extern "C"
