// host_encoder.cpp -- Newick -> canonical split bitmasks + branch lengths over a shared taxon index.
//
// Replaces, once per TREE instead of once per PAIR, what the reference does inside every task:
//   Tree(newick)  -> dendropy parse                       treeCl/tree.py:745-762
//   Tree.rooted   -> root has exactly two children        treeCl/tree.py:858-862
//   Tree.phylotree-> tree_distance.PhyloTree(newick, rooted): one bitset per internal edge + length,
//                    one length per leaf edge             treeCl/tree.py:830-845 (arithmetic in PyPI tree_distance)
// and builds the global split dictionary the merge / incidence kernels work on.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <pthread.h>
#include <string_view>
#include <thread>
#include <unordered_map>

#include "td_internal.h"

namespace td {

static thread_local char g_error[512];
void set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_error, sizeof g_error, fmt, ap); va_end(ap);
}
const char *last_error() { return g_error; }

// ------------------------------------------------------------------------------------------------------------
// Newick
// ------------------------------------------------------------------------------------------------------------
struct FlatTree {
    std::vector<int32_t> parent, nchild;
    std::vector<double> len;
    std::vector<std::string_view> label;     // leaves only; views into the input text or into `owned`
    std::deque<std::string> owned;           // unescaped quoted labels
    void clear() { parent.clear(); nchild.clear(); len.clear(); label.clear(); owned.clear(); }
    int add(int p)
    {
        parent.push_back(p); nchild.push_back(0); len.push_back(0.0); label.emplace_back();
        if (p >= 0) nchild[p]++;
        return (int)parent.size() - 1;
    }
    int size() const { return (int)parent.size(); }
};

static inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }
static inline bool is_delim(char c) { return c == '(' || c == ')' || c == ':' || c == ',' || c == ';' || c == '['; }

static const char *skip_blank(const char *s, const char *e)
{
    for (;;) {
        while (s < e && is_ws(*s)) s++;
        if (s < e && *s == '[') { while (s < e && *s != ']') s++; if (s < e) s++; continue; }
        return s;
    }
}

// reads a (possibly quoted) label starting at s; returns the end or nullptr on an unterminated quote
static const char *scan_label(const char *s, const char *e, FlatTree *t, std::string_view &out)
{
    if (s < e && *s == '\'') {
        const char *b = ++s; bool esc = false;
        while (true) {
            if (s >= e) return nullptr;
            if (*s == '\'') { if (s + 1 < e && s[1] == '\'') { esc = true; s += 2; continue; } break; }
            s++;
        }
        if (!esc) out = std::string_view(b, s - b);
        else if (t) {
            std::string u; for (const char *p = b; p < s; p++) { u.push_back(*p); if (*p == '\'') p++; }
            t->owned.push_back(std::move(u)); out = t->owned.back();
        }
        return s + 1;
    }
    const char *b = s;
    while (s < e && !is_delim(*s) && !is_ws(*s)) s++;
    out = std::string_view(b, s - b);
    return s;
}

// Node 0 is the root; "(": the current node gets its first child; ",": next sibling; ")": back to the parent.
static int parse_newick(const char *text, FlatTree &t)
{
    t.clear();
    const char *s = text, *e = text + strlen(text);
    int cur = t.add(-1), depth = 0; bool want_subtree = true, have_len = false;
    s = skip_blank(s, e);
    if (s >= e || *s != '(') { set_error("newick: expected '(' at start"); return TD_ERR_PARSE; }
    while (s < e) {
        s = skip_blank(s, e);
        if (s >= e) break;
        char c = *s;
        if (want_subtree) {
            if (c == '(') { depth++; cur = t.add(cur); have_len = false; s++; continue; }
            std::string_view lab; const char *n = scan_label(s, e, &t, lab);
            if (!n) { set_error("newick: unterminated quoted label"); return TD_ERR_PARSE; }
            if (lab.empty()) { set_error("newick: empty leaf label near '%.16s'", s); return TD_ERR_PARSE; }
            t.label[cur] = lab; s = n; want_subtree = false; continue;
        }
        if (c == ':') {
            s = skip_blank(s + 1, e);
            const char *b = s; if (b < e && *b == '+') b++;
            double v = 0; auto r = std::from_chars(b, e, v);
            if (r.ec != std::errc() || r.ptr == b) { set_error("newick: bad branch length near '%.16s'", s); return TD_ERR_PARSE; }
            if (!std::isfinite(v)) { set_error("newick: non-finite branch length near '%.16s'", s); return TD_ERR_PARSE; }
            if (have_len) { set_error("newick: second branch length on one node near '%.16s'", s); return TD_ERR_PARSE; }
            t.len[cur] = v; have_len = true; s = r.ptr; continue;
        }
        if (c == ',') {
            if (depth == 0) { set_error("newick: ',' outside parentheses"); return TD_ERR_PARSE; }
            cur = t.add(t.parent[cur]); have_len = false; s++; want_subtree = true; continue;
        }
        if (c == ')') {
            if (depth == 0) { set_error("newick: unbalanced ')'"); return TD_ERR_PARSE; }
            depth--; cur = t.parent[cur]; have_len = false; s++; continue;
        }
        if (c == ';') break;
        // internal node label / support value: dropped (treeCl/tree.py:823-825)
        std::string_view lab; const char *n = scan_label(s, e, nullptr, lab);
        if (!n || n == s) { set_error("newick: unexpected '%c'", c); return TD_ERR_PARSE; }
        s = n;
    }
    if (depth != 0 || cur != 0) { set_error("newick: unbalanced '('"); return TD_ERR_PARSE; }
    int leaves = 0;
    for (int i = 0; i < t.size(); i++) if (t.nchild[i] == 0) { if (t.label[i].empty()) { set_error("newick: unlabeled leaf"); return TD_ERR_PARSE; } leaves++; }
    if (leaves < 2) { set_error("newick: fewer than two leaves"); return TD_ERR_PARSE; }
    return TD_OK;
}

// effective root: skip unifurcations at the top (their stem edges carry no split)
static int effective_root(const FlatTree &t, const std::vector<int> *first_child = nullptr)
{
    int r = 0;
    while (t.nchild[r] == 1) {
        int c = -1;
        for (int i = r + 1; i < t.size(); i++) if (t.parent[i] == r) { c = i; break; }
        if (c < 0 || t.nchild[c] == 0) break;
        r = c;
    }
    (void)first_child;
    return r;
}

int newick_labels(const char *newick, std::vector<std::string> &labels, int &rooted)
{
    FlatTree t; int rc = parse_newick(newick, t); if (rc) return rc;
    labels.clear();
    for (int i = 0; i < t.size(); i++) if (t.nchild[i] == 0) labels.emplace_back(t.label[i]);
    std::sort(labels.begin(), labels.end());
    rooted = t.nchild[effective_root(t)] == 2;
    return TD_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Tree.prune_to_subset(subset).newick  (treeCl/tree.py:989-1001): drop the other leaves, suppress unifurcations
// summing the merged edge lengths, keep an unrooted tree's base trifurcating.
// ------------------------------------------------------------------------------------------------------------
struct PNode { std::string label; double len = 0; std::vector<PNode> kids; };

static bool prune_rec(const FlatTree &t, const std::vector<std::vector<int>> &ch, int v,
                      const std::unordered_map<std::string_view, int> &keep, PNode &out)
{
    if (t.nchild[v] == 0) {
        if (!keep.count(t.label[v])) return false;
        out.label = std::string(t.label[v]); out.len = t.len[v]; return true;
    }
    std::vector<PNode> kids;
    for (int c : ch[v]) { PNode k; if (prune_rec(t, ch, c, keep, k)) kids.push_back(std::move(k)); }
    if (kids.empty()) return false;
    if (kids.size() == 1) { out = std::move(kids[0]); out.len += t.len[v]; return true; }
    out.kids = std::move(kids); out.len = t.len[v]; return true;
}

static void put_double(std::string &s, double v)
{
    char buf[40]; auto r = std::to_chars(buf, buf + sizeof buf, v); s.append(buf, r.ptr);
}

static bool needs_quote(const std::string &l)
{
    for (char c : l) if (is_delim(c) || is_ws(c) || c == '\'' || c == ']') return true;
    return false;
}

static void write_rec(const PNode &n, bool is_root, std::string &s)
{
    if (n.kids.empty()) {
        if (needs_quote(n.label)) { s.push_back('\''); for (char c : n.label) { s.push_back(c); if (c == '\'') s.push_back('\''); } s.push_back('\''); }
        else s += n.label;
    } else {
        s.push_back('(');
        for (size_t i = 0; i < n.kids.size(); i++) { if (i) s.push_back(','); write_rec(n.kids[i], false, s); }
        s.push_back(')');
    }
    if (!is_root) { s.push_back(':'); put_double(s, n.len); }
}

int newick_prune(const char *newick, const std::vector<std::string> &keep_list, std::string &out)
{
    FlatTree t; int rc = parse_newick(newick, t); if (rc) return rc;
    std::unordered_map<std::string_view, int> keep;
    for (auto &k : keep_list) keep.emplace(std::string_view(k), 1);
    std::vector<std::vector<int>> ch(t.size());
    for (int i = 1; i < t.size(); i++) ch[t.parent[i]].push_back(i);
    bool was_rooted = t.nchild[effective_root(t)] == 2;
    PNode root;
    if (!prune_rec(t, ch, 0, keep, root)) { set_error("prune: no leaf of the tree is in the subset"); return TD_ERR_ARG; }
    root.len = 0;
    // unrooted trees keep a trifurcating base: merge the two root edges (lengths summed)
    if (!was_rooted && root.kids.size() == 2) {
        int del = -1;
        if (root.kids[1].kids.size() >= 2) del = 1; else if (root.kids[0].kids.size() >= 2) del = 0;
        if (del >= 0) {
            PNode d = std::move(root.kids[del]); root.kids.erase(root.kids.begin() + del);
            root.kids[0].len += d.len;
            for (auto &k : d.kids) root.kids.push_back(std::move(k));
        }
    }
    out.clear(); write_rec(root, true, out); out.push_back(';');
    return TD_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Split encoding of one tree over the shared taxon table.
// ------------------------------------------------------------------------------------------------------------
struct TreeEnc {
    int rooted = 0, n_leaves = 0;
    std::vector<uint64_t> leafmask;          // W
    std::vector<uint64_t> masks;             // S*W, sorted by mask, unique
    std::vector<double> lens;                // S
    std::vector<double> leaflen;             // n_taxa
    double sum_len = 0, sum_sq = 0;
};

struct TaxonTable {
    std::vector<std::string> labels;                         // sorted
    // label -> index: open addressing over a power-of-two table (load <= 1/4), FNV-1a of the label bytes.  (7 % of the per-tree encode time
    // against std::unordered_map: no modulo, no node dereference.)
    std::vector<int32_t> slots;
    uint32_t mask = 0;
    static uint32_t hash(std::string_view s) { uint32_t h = 2166136261u; for (unsigned char c : s) { h ^= c; h *= 16777619u; } return h ^ (h >> 15); }
    void build(std::vector<std::string> &&l)
    {
        labels = std::move(l); std::sort(labels.begin(), labels.end());
        labels.erase(std::unique(labels.begin(), labels.end()), labels.end());
        size_t cap = 16; while (cap < labels.size() * 4) cap <<= 1;
        slots.assign(cap, -1); mask = (uint32_t)(cap - 1);
        for (size_t i = 0; i < labels.size(); i++) { uint32_t e = hash(labels[i]) & mask; while (slots[e] >= 0) e = (e + 1) & mask; slots[e] = (int32_t)i; }
    }
    int find(std::string_view s) const                       // -1: not in the table
    {
        if (slots.empty()) return -1;
        for (uint32_t e = hash(s) & mask;; e = (e + 1) & mask) {
            const int32_t i = slots[e];
            if (i < 0) return -1;
            if (labels[(size_t)i] == s) return i;
        }
    }
};

static inline int mask_cmp(const uint64_t *a, const uint64_t *b, int W)
{
    for (int w = W - 1; w >= 0; w--) { if (a[w] != b[w]) return a[w] < b[w] ? -1 : 1; }
    return 0;
}
static inline int mask_popc(const uint64_t *m, int W) { int c = 0; for (int w = 0; w < W; w++) c += __builtin_popcountll(m[w]); return c; }
static inline int mask_low(const uint64_t *m, int W) { for (int w = 0; w < W; w++) if (m[w]) return 64 * w + __builtin_ctzll(m[w]); return -1; }

// returns TD_OK, TD_ERR_PARSE, or -1 when a label is missing from the table (caller rebuilds the table)
static int encode_tree(const FlatTree &t, const TaxonTable &tab, TreeEnc &e, std::vector<uint64_t> &clade,
                       std::vector<int> &order)
{
    const int n_taxa = (int)tab.labels.size(), W = std::max(1, (n_taxa + 63) / 64), nn = t.size();
    e.leafmask.assign(W, 0); e.leaflen.assign(n_taxa, 0.0); e.masks.clear(); e.lens.clear();
    clade.assign((size_t)nn * W, 0);
    for (int i = nn - 1; i >= 0; i--) {       // children have larger indices than their parent
        uint64_t *m = &clade[(size_t)i * W];
        if (t.nchild[i] == 0) {
            const int k = tab.find(t.label[i]);
            if (k < 0) return -1;
            if (e.leafmask[k >> 6] >> (k & 63) & 1) { set_error("newick: duplicate leaf label '%.*s'", (int)t.label[i].size(), t.label[i].data()); return TD_ERR_PARSE; }
            m[k >> 6] |= 1ull << (k & 63); e.leafmask[k >> 6] |= 1ull << (k & 63);
        }
        if (t.parent[i] >= 0) { uint64_t *p = &clade[(size_t)t.parent[i] * W]; for (int w = 0; w < W; w++) p[w] |= m[w]; }
    }
    const int root = effective_root(t);
    e.rooted = t.nchild[root] == 2;
    e.n_leaves = mask_popc(e.leafmask.data(), W);
    const int ref = mask_low(e.leafmask.data(), W);
    order.clear();
    for (int i = 1; i < nn; i++) {
        if (i == root) continue;
        uint64_t *m = &clade[(size_t)i * W];
        const int c = mask_popc(m, W), cc = e.n_leaves - c;
        if (c == 0 || cc == 0) continue;                                   // stem above the effective root
        if (c == 1) { e.leaflen[mask_low(m, W)] += t.len[i]; continue; }   // leaf edge (or a unifurcation above one)
        if (!e.rooted) {
            if (cc == 1) { for (int w = 0; w < W; w++) m[w] = e.leafmask[w] & ~m[w]; e.leaflen[mask_low(m, W)] += t.len[i]; continue; }
            if (m[ref >> 6] >> (ref & 63) & 1) for (int w = 0; w < W; w++) m[w] = e.leafmask[w] & ~m[w];
        }
        order.push_back(i);
    }
    if (W == 1) {                                   // one word per mask: sort (mask, node) keys directly, no indirect comparisons
        std::pair<uint64_t, int> keys[64]; std::vector<std::pair<uint64_t, int>> more;
        std::pair<uint64_t, int> *k = keys;
        if (order.size() > 64) { more.resize(order.size()); k = more.data(); }
        for (size_t j = 0; j < order.size(); j++) k[j] = {clade[(size_t)order[j]], order[j]};
        std::sort(k, k + order.size());
        for (size_t j = 0; j < order.size(); j++) order[j] = k[j].second;
    } else
        std::sort(order.begin(), order.end(), [&](int a, int b) { return mask_cmp(&clade[(size_t)a * W], &clade[(size_t)b * W], W) < 0; });
    e.masks.reserve(order.size() * W); e.lens.reserve(order.size());
    for (int i : order) {
        const uint64_t *m = &clade[(size_t)i * W];
        size_t S = e.lens.size();
        if (S && mask_cmp(&e.masks[(S - 1) * W], m, W) == 0) { e.lens[S - 1] += t.len[i]; continue; }   // merged edges: lengths add
        e.masks.insert(e.masks.end(), m, m + W); e.lens.push_back(t.len[i]);
    }
    e.sum_len = 0; e.sum_sq = 0;
    for (double l : e.lens) { e.sum_len += l; e.sum_sq += l * l; }
    for (double l : e.leaflen) { e.sum_len += l; e.sum_sq += l * l; }
    return TD_OK;
}

static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
static inline uint64_t mask_hash(const uint64_t *m, int W)
{
    uint64_t h = 0x9e3779b97f4a7c15ULL;
    for (int w = 0; w < W; w++) h = mix64(h ^ m[w]) + 0x9e3779b97f4a7c15ULL * (w + 1);
    return h;
}

// Worker pool shared by all encode calls of the process: spawning 16 threads for each of the ~10 parallel sections of one encode
// cost several milliseconds.  Workers sleep on a condition variable between sections; a section that finds the pool busy (another
// encode running concurrently) falls back to plain threads.
class WorkerPool {
public:
    explicit WorkerPool(int workers)
    {
        for (int t = 0; t < workers; t++) std::thread([this, t] { loop(t + 1); }).detach();
        n_workers_ = workers;
    }
    int workers() const { return n_workers_; }
    bool try_acquire() { return busy_.try_lock(); }
    void release() { busy_.unlock(); }
    // run job(tid) on the caller (tid 0) and on workers 1 .. n_threads-1; returns when all are done
    void run(int n_threads, const std::function<void(int)> &job)
    {
        const int helpers = std::min(n_threads - 1, n_workers_);
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = &job; helpers_ = helpers; pending_ = helpers; generation_++;
        }
        cv_.notify_all();
        job(0);
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int tid)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)> *job = nullptr;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
                if (tid <= helpers_) job = job_;
            }
            if (job) {
                (*job)(tid);
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex mu_, busy_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *job_ = nullptr;
    int helpers_ = 0, pending_ = 0, n_workers_ = 0;
    uint64_t generation_ = 0;
};

static std::atomic<WorkerPool *> g_pool{nullptr};
static std::mutex g_pool_mu;

static WorkerPool *worker_pool()
{
    WorkerPool *p = g_pool.load(std::memory_order_acquire);
    if (p) return p;
    std::lock_guard<std::mutex> g(g_pool_mu);
    p = g_pool.load(std::memory_order_relaxed);
    if (!p) {
        // a forked child has none of the parent's threads: it starts over with a pool of its own (the parent's is simply abandoned)
        static bool registered = false;
        if (!registered) { pthread_atfork(nullptr, nullptr, [] { g_pool.store(nullptr); new (&g_pool_mu) std::mutex(); }); registered = true; }
        p = new WorkerPool((int)std::max(1u, std::thread::hardware_concurrency()) - 1);   // never destroyed: workers are detached
        g_pool.store(p, std::memory_order_release);
    }
    return p;
}

template <class F>
static void parallel_for(int64_t n, int n_threads, F &&fn)
{
    if (n_threads <= 1 || n < 2) { for (int64_t i = 0; i < n; i++) fn(i, 0); return; }
    std::atomic<int64_t> next{0};
    const int64_t chunk = std::max<int64_t>(1, n / (n_threads * 16));
    // an exception inside a worker (bad_alloc) must not escape its thread (std::terminate): it is parked and rethrown on the caller
    std::atomic<bool> failed{false};
    auto body = [&](int t) {
        try {
            for (;;) { int64_t b = next.fetch_add(chunk); if (b >= n || failed.load(std::memory_order_relaxed)) break; int64_t e = std::min(n, b + chunk); for (int64_t i = b; i < e; i++) fn(i, t); }
        } catch (...) { failed.store(true); }
    };
    WorkerPool *pool = worker_pool();
    if (pool->workers() + 1 >= n_threads && pool->try_acquire()) {
        const std::function<void(int)> job = body;
        pool->run(n_threads, job);
        pool->release();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(body, t);
        for (auto &x : th) x.join();
    }
    if (failed.load()) throw std::bad_alloc();
}

static int finish_encode(std::vector<TreeEnc> &enc, const std::vector<std::string> &labels, int n_threads, HostSet &hs,
                         const std::function<void(const char *)> &lap);

// The per-tree scratch of the last encode (four small vectors per tree) is kept for the next one when it is small: an application that
// encodes one set after another (bootstrap replicates, simulations, the bench's end-to-end loop) then neither allocates nor frees 20,000
// vectors per call.  At most kSpareBytes are kept; td_trim_cache() drops them.
constexpr size_t kSpareBytes = 64u << 20;
static std::mutex g_spare_mu;
static std::vector<TreeEnc> g_spare;

void encoder_trim()
{
    std::vector<TreeEnc> gone;
    { std::lock_guard<std::mutex> g(g_spare_mu); gone.swap(g_spare); }
}

int encode_newick(const char *const *newick, int64_t N, int n_threads, HostSet &hs)
{
    if (N < 1 || !newick) { set_error("td_encode_newick: need at least one tree"); return TD_ERR_ARG; }
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, N / 8));

    const bool timing = getenv("TREEDIST_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (!timing) return; auto n = std::chrono::steady_clock::now(); fprintf(stderr, "[encode] %-12s %.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t_last).count()); t_last = n; };
    std::vector<TreeEnc> enc;
    { std::lock_guard<std::mutex> g(g_spare_mu); enc.swap(g_spare); }
    enc.resize(N);
    TaxonTable tab;
    std::atomic<int> status{TD_OK};
    std::mutex err_mu; std::string err_msg;
    auto fail = [&](int rc, int64_t i) {
        int exp = TD_OK;
        if (status.compare_exchange_strong(exp, rc)) { std::lock_guard<std::mutex> g(err_mu); err_msg = "tree " + std::to_string(i) + ": " + last_error(); }
    };
    // seed the taxon table from tree 0; trees with labels outside it force one rebuild over the union
    {
        std::vector<std::string> l0; int r0;
        int rc = newick_labels(newick[0], l0, r0);
        if (rc) { std::string m = std::string("tree 0: ") + last_error(); set_error("%s", m.c_str()); return rc; }
        tab.build(std::move(l0));
    }
    lap("setup");
    std::atomic<bool> missing{false};
    struct Scratch { FlatTree t; std::vector<uint64_t> clade; std::vector<int> order; };
    std::vector<Scratch> scratch(n_threads);
    auto run_pass = [&] {
        parallel_for(N, n_threads, [&](int64_t i, int tid) {
            if (status.load(std::memory_order_relaxed) != TD_OK) return;
            Scratch &sc = scratch[tid];
            int rc = parse_newick(newick[i], sc.t);
            if (rc) { fail(rc, i); return; }
            rc = encode_tree(sc.t, tab, enc[i], sc.clade, sc.order);
            if (rc == -1) missing = true; else if (rc) fail(rc, i);
        });
    };
    run_pass();
    if (status == TD_OK && missing) {
        std::vector<std::vector<std::string>> per(n_threads);
        parallel_for(N, n_threads, [&](int64_t i, int tid) {
            Scratch &sc = scratch[tid];
            if (parse_newick(newick[i], sc.t)) return;
            for (int v = 0; v < sc.t.size(); v++) if (sc.t.nchild[v] == 0 && tab.find(sc.t.label[v]) < 0) per[tid].emplace_back(sc.t.label[v]);
        });
        std::vector<std::string> all = tab.labels;
        for (auto &p : per) all.insert(all.end(), p.begin(), p.end());
        tab.build(std::move(all));
        missing = false;
        run_pass();
    }
    if (status != TD_OK) { set_error("%s", err_msg.c_str()); return status; }
    lap("parse+encode");
    return finish_encode(enc, tab.labels, n_threads, hs, lap);
}

// Copies for the upload staging buffer (api.cu): pieces of at most 256 KB spread over the worker pool.
void parallel_copy(const std::vector<CopyJob> &jobs)
{
    struct Piece { unsigned char *dst; const unsigned char *src; size_t bytes; int fill; };
    std::vector<Piece> pieces;
    constexpr size_t kPiece = 256 << 10;
    for (const CopyJob &j : jobs) {
        for (size_t o = 0; o < j.bytes; o += kPiece) pieces.push_back({(unsigned char *)j.dst + o, (const unsigned char *)j.src + o, std::min(kPiece, j.bytes - o), -1});
        if (j.pad_bytes) pieces.push_back({(unsigned char *)j.dst + j.bytes, nullptr, j.pad_bytes, j.pad_value});
    }
    const int n_threads = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(pieces.size() / 2, 1));
    // Non-temporal stores: the buffer is read next by the GPU's copy engine, not by a core.  Written with ordinary stores by 16 threads,
    // its lines sit dirty in 16 caches and the DMA ran at 5 GB/s (6.7 MB: 1.26 ms); streamed to memory it runs at PCIe speed.
    parallel_for((int64_t)pieces.size(), n_threads, [&](int64_t i, int) {
        const Piece &pc = pieces[(size_t)i];
        if (!pc.src) { memset(pc.dst, pc.fill, pc.bytes); return; }
#if defined(__SSE2__)
        unsigned char *d = pc.dst; const unsigned char *s = pc.src; size_t n = pc.bytes;
        const size_t head = std::min(n, (size_t)((16 - ((uintptr_t)d & 15)) & 15));
        memcpy(d, s, head); d += head; s += head; n -= head;
        for (; n >= 64; n -= 64, d += 64, s += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)s), b = _mm_loadu_si128((const __m128i *)(s + 16));
            const __m128i c = _mm_loadu_si128((const __m128i *)(s + 32)), e = _mm_loadu_si128((const __m128i *)(s + 48));
            _mm_stream_si128((__m128i *)d, a); _mm_stream_si128((__m128i *)(d + 16), b);
            _mm_stream_si128((__m128i *)(d + 32), c); _mm_stream_si128((__m128i *)(d + 48), e);
        }
        memcpy(d, s, n);
        _mm_sfence();
#else
        memcpy(pc.dst, pc.src, pc.bytes);
#endif
    });
}

// Everything after the per-tree encoding: SoA packing, the split dictionary with its postings, the per-tree lists and aggregates.
static int finish_encode(std::vector<TreeEnc> &enc, const std::vector<std::string> &labels, int n_threads, HostSet &hs,
                         const std::function<void(const char *)> &lap)
{
    const int64_t N = (int64_t)enc.size();
    // ---- pack ------------------------------------------------------------------------------------------
    const int n_taxa = (int)labels.size(), W = std::max(1, (n_taxa + 63) / 64);
    hs = HostSet();
    hs.N = N; hs.n_taxa = n_taxa; hs.W = W; hs.labels = labels;
    hs.n_splits.resize(N); hs.n_leaves.resize(N); hs.tree_rooted.resize(N); hs.n_shared.assign(N, 0);
    hs.leafmask.resize((size_t)N * W); hs.sum_len.resize(N); hs.norm.resize(N); hs.sum_sq.resize(N);
    hs.single_l1.assign(N, 0.0); hs.single_l2.assign(N, 0.0);
    hs.q1.assign(N, 0.0); hs.q2.assign(N, 0.0);
    int smax = 0; bool uniform = true; int64_t total = 0;
    for (int64_t i = 0; i < N; i++) {
        const TreeEnc &e = enc[i];
        hs.n_splits[i] = (int)e.lens.size(); hs.n_leaves[i] = e.n_leaves; hs.tree_rooted[i] = e.rooted;
        memcpy(&hs.leafmask[(size_t)i * W], e.leafmask.data(), W * 8);
        hs.sum_len[i] = e.sum_len; hs.norm[i] = std::sqrt(e.sum_sq); hs.sum_sq[i] = e.sum_sq;
        smax = std::max(smax, hs.n_splits[i]); total += hs.n_splits[i];
        if (e.rooted != enc[0].rooted || memcmp(e.leafmask.data(), enc[0].leafmask.data(), W * 8)) uniform = false;
    }
    hs.max_splits = smax; hs.uniform = uniform; hs.rooted = enc[0].rooted; hs.total_splits = total;
    const size_t stride = (size_t)std::max(smax, 1);
    // (uninitialised: every row is written in full, padding included, by the parallel `lists` pass below)
    hs.masks.resize((size_t)N * stride * W); hs.lens.resize((size_t)N * stride);
    hs.ids.resize((size_t)N * stride);
    hs.leaflen.resize((size_t)N * std::max(n_taxa, 1));
    if (n_taxa == 0) std::fill(hs.leaflen.begin(), hs.leaflen.end(), 0.0);
    parallel_for(N, n_threads, [&](int64_t i, int) {
        memcpy(&hs.leaflen[(size_t)i * n_taxa], enc[i].leaflen.data(), (size_t)n_taxa * 8);
    });

    lap("pack");
    // ---- global split dictionary (exact: ties on the 64-bit hash are resolved by comparing the masks) --------
    struct Ent { uint64_t h; uint32_t tree, slot; };
    // per (tree, slot): dictionary id, multiplicity, compact shared index, position in the postings.  Only slots < n_splits are ever
    // written or read, so the arrays are left uninitialised (a serial fill of four of them and of the two entry tables cost ~1.5 ms)
    const size_t n_slots = (size_t)N * stride;
    // (one 16-byte record per slot: the dictionary pass writes them in hash order, i.e. at random - one cache line per entry, not four)
    struct SlotInfo { uint32_t id, mult, sid, x; };
    std::unique_ptr<SlotInfo[]> slot_of(new SlotInfo[n_slots]);
    if (total > 0) {
        std::vector<int64_t> off(N + 1, 0);
        for (int64_t i = 0; i < N; i++) off[i + 1] = off[i] + hs.n_splits[i];
        // hash every split and scatter the entries into 256 hash buckets: the trees are cut into chunks, every chunk counts its
        // entries per bucket, a prefix over (bucket, chunk) gives every chunk its own slice of every bucket, and the chunks scatter
        // in parallel (entries of a bucket stay in tree order)
        constexpr int NB = 256;
        std::unique_ptr<Ent[]> ents(new Ent[total]);
        const int64_t n_chunks = std::min<int64_t>(N, (int64_t)n_threads * 4);
        std::vector<int64_t> cpos((size_t)n_chunks * NB, 0);
        parallel_for(n_chunks, n_threads, [&](int64_t c, int) {
            int64_t *cnt = &cpos[(size_t)c * NB];
            for (int64_t i = N * c / n_chunks; i < N * (c + 1) / n_chunks; i++)
                for (int s = 0; s < hs.n_splits[i]; s++) {
                    const Ent en{mask_hash(&enc[i].masks[(size_t)s * W], W), (uint32_t)i, (uint32_t)s};
                    ents[off[i] + s] = en; cnt[en.h >> 56]++;
                }
        });
        std::vector<int64_t> bstart(NB + 1, 0);
        {
            int64_t run = 0;
            for (int b = 0; b < NB; b++) {
                bstart[b] = run;
                for (int64_t c = 0; c < n_chunks; c++) { const int64_t n = cpos[(size_t)c * NB + b]; cpos[(size_t)c * NB + b] = run; run += n; }
            }
            bstart[NB] = run;
        }
        std::unique_ptr<Ent[]> sorted(new Ent[total]);
        parallel_for(n_chunks, n_threads, [&](int64_t c, int) {
            int64_t *pos = &cpos[(size_t)c * NB];
            for (int64_t e = off[N * c / n_chunks]; e < off[N * (c + 1) / n_chunks]; e++) sorted[pos[ents[e].h >> 56]++] = ents[e];
        });
        ents.reset();
        lap("  dict: hash+scatter");
        auto mask_of = [&](const Ent &en) { return &enc[en.tree].masks[(size_t)en.slot * W]; };
        std::vector<int64_t> uniq(NB + 1, 0), shared(NB + 1, 0), shent(NB + 1, 0);
        std::vector<double> copairs(NB, 0.0);
        std::vector<std::vector<uint32_t>> bmult(NB);          // multiplicities of the shared runs of every bucket, in run order
        parallel_for(NB, n_threads, [&](int64_t b, int) {
            Ent *lo = sorted.get() + bstart[b], *hi = sorted.get() + bstart[b + 1];
            std::sort(lo, hi, [&](const Ent &x, const Ent &y) {
                if (x.h != y.h) return x.h < y.h;
                int c = mask_cmp(mask_of(x), mask_of(y), W);
                if (c) return c < 0;
                return x.tree < y.tree;
            });
            int64_t u = 0, sh = 0, she = 0; double cp = 0;
            for (auto it = lo; it != hi;) {
                auto run = it + 1;
                while (run != hi && run->h == it->h && mask_cmp(mask_of(*run), mask_of(*it), W) == 0) run++;
                u++; if (run - it >= 2) { sh++; she += run - it; bmult[b].push_back((uint32_t)(run - it)); }
                cp += 0.5 * (double)(run - it) * (double)(run - it - 1);
                it = run;
            }
            uniq[b + 1] = u; shared[b + 1] = sh; shent[b + 1] = she; copairs[b] = cp;
        });
        lap("  dict: sort+count");
        double cp_total = 0;
        for (int b = 0; b < NB; b++) { uniq[b + 1] += uniq[b]; shared[b + 1] += shared[b]; shent[b + 1] += shent[b]; cp_total += copairs[b]; }
        hs.dict_shared = shared[NB];
        hs.total_common = (int64_t)(cp_total + 0.5);
        hs.mean_common = N > 1 ? cp_total / (0.5 * (double)N * (double)(N - 1)) : 0.0;
        hs.dict_size = uniq[NB];
        // Compact index of the shared splits = rank by DESCENDING multiplicity: the splits held by many trees (most of the events) come
        // first, so "the heavy rows" of the hybrid tile kernel are simply the indices below hs.heavy_rows.
        std::vector<uint32_t> sid_rank((size_t)hs.dict_shared), mult_sorted((size_t)hs.dict_shared);
        {
            std::vector<uint32_t> mult_old((size_t)hs.dict_shared);
            for (int b = 0; b < NB; b++) std::copy(bmult[b].begin(), bmult[b].end(), mult_old.begin() + shared[b]);
            std::vector<uint32_t> order((size_t)hs.dict_shared);
            for (size_t k = 0; k < order.size(); k++) order[k] = (uint32_t)k;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return mult_old[a] > mult_old[c]; });
            for (size_t r = 0; r < order.size(); r++) { sid_rank[order[r]] = (uint32_t)r; mult_sorted[r] = mult_old[order[r]]; }
        }
        // Heavy rows by the cost model of the postings pipeline (ps per pair): a row of the extended matrix costs 0.064, an expected
        // common split per pair left to the events costs 13.4.  Rows are taken (in steps of 4, at most 4096) while they pay.
        {
            const double P = 0.5 * (double)N * (double)(N - 1);
            double best = 0.0, gain = 0.0; int best_r = 0;
            for (size_t r = 0; r < mult_sorted.size() && r < 4096; r++) {
                gain += 13.4 * 0.5 * (double)mult_sorted[r] * (double)(mult_sorted[r] - 1) / std::max(P, 1.0) - 0.064;
                if (((r + 1) % 4 == 0 || r + 1 == mult_sorted.size()) && gain > best) { best = gain; best_r = (int)(r + 1); }
            }
            hs.heavy_rows = best > 2.0 ? best_r : 0;           // not for a gain below 2 ps per pair (15 % of an almost event-free step)
            double he = 0.0;
            for (int r = 0; r < hs.heavy_rows; r++) he += 0.5 * (double)mult_sorted[r] * (double)(mult_sorted[r] - 1);
            hs.heavy_common = he / std::max(P, 1.0); hs.heavy_events = (int64_t)(he + 0.5);
        }
        const int64_t n_post = shent[NB];
        if (n_post >= (1ll << 31)) { set_error("too many shared-split entries (%lld)", (long long)n_post); return TD_ERR_UNSUPPORTED; }
        // ids, multiplicities, the compact index of the shared splits (column of the incidence matrix) and their postings: a run of
        // equal splits in the sorted bucket IS the posting list of that split (trees ascending)
        hs.post_tree.resize(n_post); hs.post_end.resize(n_post); hs.post_len.resize(n_post);
        parallel_for(NB, n_threads, [&](int64_t b, int) {
            Ent *lo = sorted.get() + bstart[b], *hi = sorted.get() + bstart[b + 1];
            int64_t id = uniq[b], sid = shared[b], x = shent[b];
            for (auto it = lo; it != hi;) {
                auto run = it + 1;
                while (run != hi && run->h == it->h && mask_cmp(mask_of(*run), mask_of(*it), W) == 0) run++;
                const bool is_shared = run - it >= 2;
                const int64_t x_end = x + (run - it);
                for (auto k = it; k != run; k++) {
                    const size_t at = (size_t)k->tree * stride + k->slot;
                    SlotInfo &si = slot_of[at];
                    si.id = (uint32_t)id; si.mult = (uint32_t)(run - it);
                    if (is_shared) {
                        si.sid = sid_rank[(size_t)sid]; si.x = (uint32_t)x;
                        hs.post_tree[x] = (int32_t)k->tree; hs.post_end[x] = (int32_t)x_end; hs.post_len[x] = enc[k->tree].lens[k->slot];
                        x++;
                    }
                }
                if (is_shared) sid++;
                id++; it = run;
            }
        });
    }
    lap("dictionary");
    // ---- per-tree lists sorted by dictionary id; shared-split lists ------------------------------------------
    std::vector<int32_t> nsh(N, 0);
    parallel_for(N, n_threads, [&](int64_t i, int) {
        int sh = 0;
        for (int s = 0; s < hs.n_splits[i]; s++) sh += slot_of[(size_t)i * stride + s].mult >= 2;
        nsh[i] = sh;
    });
    int shmax = 0; for (int64_t i = 0; i < N; i++) shmax = std::max(shmax, nsh[i]);
    hs.n_shared.assign(nsh.begin(), nsh.end()); hs.max_shared = shmax;
    for (int64_t b = 0; b < N; b += kTile) { int tot = 0; for (int64_t i = b; i < std::min<int64_t>(N, b + kTile); i++) tot += nsh[i]; hs.max_tile_entries = std::max(hs.max_tile_entries, tot); }
    hs.sh_stride = ((shmax + 1) + 1) & ~1;                   // >= one sentinel, even -> 32-byte rows
    hs.shared.resize((size_t)N * hs.sh_stride);            // (rows written in full below: entries, then sentinels)
    hs.shared_x.resize((size_t)N * hs.sh_stride);
    hs.q2t.assign(N, 0.0); hs.q1_light.assign(N, 0.0);
    parallel_for(N, n_threads, [&](int64_t i, int) {
        const int S = hs.n_splits[i];
        std::vector<int> perm(S);
        for (int s = 0; s < S; s++) perm[s] = s;
        std::sort(perm.begin(), perm.end(), [&](int a, int b) { return slot_of[(size_t)i * stride + a].id < slot_of[(size_t)i * stride + b].id; });
        int sh = 0; double s1 = 0, s2 = 0, t1 = 0, t2 = 0, h1 = 0;
        for (int k = 0; k < S; k++) {
            const int s = perm[k];
            const size_t at = (size_t)i * stride + s;
            const double len = enc[i].lens[s];
            memcpy(&hs.masks[((size_t)i * stride + k) * W], &enc[i].masks[(size_t)s * W], W * 8);
            hs.lens[(size_t)i * stride + k] = len;
            const SlotInfo &si = slot_of[at];
            hs.ids[(size_t)i * stride + k] = si.id;
            t1 += std::fabs(len); t2 += len * len;
            if (!(si.mult >= 2 && (int)si.sid < hs.heavy_rows)) h1 += std::fabs(len);      // everything but the heavy splits
            if (si.mult >= 2) { hs.shared_x[(size_t)i * hs.sh_stride + sh] = (int32_t)si.x; hs.shared[(size_t)i * hs.sh_stride + sh++] = SplitRec{si.id, si.sid, len}; }
            else { s1 += std::fabs(len); s2 += len * len; }
        }
        for (size_t k = (size_t)S; k < stride; k++) {                                  // padding of the full lists
            for (int w = 0; w < W; w++) hs.masks[((size_t)i * stride + k) * W + w] = 0;
            hs.lens[(size_t)i * stride + k] = 0.0; hs.ids[(size_t)i * stride + k] = kSentinelId;
        }
        for (int k = sh; k < hs.sh_stride; k++) { hs.shared[(size_t)i * hs.sh_stride + k] = SplitRec{kSentinelId, 0, 0.0}; hs.shared_x[(size_t)i * hs.sh_stride + k] = -1; }
        hs.single_l1[i] = s1; hs.single_l2[i] = s2; hs.q1[i] = t1; hs.q2[i] = t2; hs.q1_light[i] = h1;
        double sl = 0.0;
        for (int k = 0; k < n_taxa; k++) { const double l = hs.leaflen[(size_t)i * n_taxa + k]; sl += l * l; }
        hs.q2t[i] = t2 + sl;
    });
    lap("lists");
    // the per-tree vectors (four per tree): kept for the next encode when small (see g_spare), else released by the workers - left to the
    // caller's single thread, 20,000 frees cost ~1 ms
    const size_t scratch_bytes = (size_t)N * (sizeof(TreeEnc) + (size_t)W * 8 + (size_t)n_taxa * 8) + (size_t)total * ((size_t)W * 8 + 8);
    bool kept = false;
    if (scratch_bytes <= kSpareBytes) {
        std::lock_guard<std::mutex> g(g_spare_mu);
        if (g_spare.empty()) { g_spare.swap(enc); kept = true; }
    }
    if (!kept) parallel_for(N, n_threads, [&](int64_t i, int) { TreeEnc gone; std::swap(gone, enc[(size_t)i]); });
    lap("release");
    return TD_OK;
}

// ---- Kendall-Colijn vectors (treeCl/utils/kendallcolijn.py:38-79) ------------------------------------------------------
// v = (1 - lambda) m + lambda M: for every pair of leaves (sorted label order, itertools.combinations order) the number of edges and
// the path length from the root to their MRCA, then per leaf the constant 1 and its edge length.  The set is packed as "trees" without
// internal edges whose leaf-edge lengths are the vector entries, so that the Euclidean kernel (td_pairwise with TD_EUC) returns exactly
// KendallColijn.get_distance for trees on one leaf set.
int encode_kendall_colijn(const char *const *newick, int64_t N, double lambda, int n_threads, HostSet &hs)
{
    if (N < 1 || !newick) { set_error("td_encode_kendall_colijn: need at least one tree"); return TD_ERR_ARG; }
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, N / 8));
    TaxonTable tab;
    {
        std::vector<std::string> l0; int r0;
        int rc = newick_labels(newick[0], l0, r0);
        if (rc) { std::string m = std::string("tree 0: ") + last_error(); set_error("%s", m.c_str()); return rc; }
        tab.build(std::move(l0));
    }
    const int n = (int)tab.labels.size();
    const int64_t D = (int64_t)n * (n - 1) / 2 + n;
    if (D >= (1 << 24)) { set_error("Kendall-Colijn vectors of %lld entries are not supported", (long long)D); return TD_ERR_UNSUPPORTED; }
    const int W = (int)((D + 63) / 64);
    std::vector<TreeEnc> enc(N);
    std::atomic<int> status{TD_OK};
    std::mutex err_mu; std::string err_msg;
    auto fail = [&](int rc, int64_t i, const std::string &what) {
        int exp = TD_OK;
        if (status.compare_exchange_strong(exp, rc)) { std::lock_guard<std::mutex> g(err_mu); err_msg = "tree " + std::to_string(i) + ": " + what; }
    };
    struct Scratch { FlatTree t; std::vector<int> edges, taxon, first, next, tail; std::vector<double> dist; std::vector<char> seen; };
    std::vector<Scratch> scratch(n_threads);
    parallel_for(N, n_threads, [&](int64_t i, int tid) {
        if (status.load(std::memory_order_relaxed) != TD_OK) return;
        Scratch &sc = scratch[tid];
        if (int rc = parse_newick(newick[i], sc.t)) { fail(rc, i, last_error()); return; }
        const FlatTree &t = sc.t;
        const int nn = t.size();
        // depths (parents precede their children), leaves -> taxon index
        sc.edges.assign(nn, 0); sc.dist.assign(nn, 0.0); sc.taxon.assign(nn, -1);
        int n_leaves = 0;
        sc.seen.assign(n, 0);
        for (int v = 1; v < nn; v++) { sc.edges[v] = sc.edges[t.parent[v]] + 1; sc.dist[v] = sc.dist[t.parent[v]] + t.len[v]; }
        TreeEnc &e = enc[i];
        e.rooted = 0; e.n_leaves = (int)D; e.masks.clear(); e.lens.clear();
        e.leafmask.assign(W, 0);
        for (int64_t k = 0; k < D; k++) e.leafmask[k >> 6] |= 1ull << (k & 63);
        e.leaflen.assign((size_t)D, 0.0);
        for (int v = 0; v < nn; v++)
            if (t.nchild[v] == 0) {
                const int k = tab.find(t.label[v]);
                if (k < 0 || sc.seen[k]) { fail(TD_ERR_UNSUPPORTED, i, "Kendall-Colijn sets need one common leaf set (prune first)"); return; }
                sc.taxon[v] = k; sc.seen[k] = 1; n_leaves++;
                e.leaflen[(size_t)(D - n + k)] = (1.0 - lambda) + lambda * t.len[v];
            }
        if (n_leaves != n) { fail(TD_ERR_UNSUPPORTED, i, "Kendall-Colijn sets need one common leaf set (prune first)"); return; }
        // leaf lists per node as linked lists; node v (children have larger indices) hands its list to its parent after pairing it
        // with the leaves the parent has collected so far: every cross pair has the parent as MRCA
        sc.first.assign(nn, -1); sc.tail.assign(nn, -1); sc.next.assign(nn, -1);
        for (int v = nn - 1; v >= 1; v--) {
            if (t.nchild[v] == 0) { sc.first[v] = sc.tail[v] = v; }
            const int p = t.parent[v];
            const double val = (1.0 - lambda) * (double)sc.edges[p] + lambda * sc.dist[p];
            for (int x = sc.first[p]; x >= 0; x = sc.next[x])
                for (int y = sc.first[v]; y >= 0; y = sc.next[y]) {
                    int a = sc.taxon[x], b = sc.taxon[y];
                    if (a > b) std::swap(a, b);
                    e.leaflen[(size_t)((int64_t)a * n - (int64_t)a * (a + 1) / 2 + (b - a - 1))] = val;
                }
            if (sc.first[v] >= 0) {
                if (sc.first[p] < 0) sc.first[p] = sc.first[v]; else sc.next[sc.tail[p]] = sc.first[v];
                sc.tail[p] = sc.tail[v];
            }
        }
        double sl = 0.0, sq = 0.0;
        for (double x : e.leaflen) { sl += x; sq += x * x; }
        e.sum_len = sl; e.sum_sq = sq;
    });
    if (status != TD_OK) { set_error("%s", err_msg.c_str()); return status; }
    // pseudo-labels of the vector entries: "a|b" for the pairs, "a" for the leaves
    std::vector<std::string> names; names.reserve((size_t)D);
    for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) names.push_back(tab.labels[a] + "|" + tab.labels[b]);
    for (int a = 0; a < n; a++) names.push_back(tab.labels[a]);
    return finish_encode(enc, names, n_threads, hs, [](const char *) {});
}

}  // namespace td

// ------------------------------------------------------------------------------------------------------------
// Leaf-set groups.  The per-tree encodings of the members are rebuilt from the parent's packed arrays (a non-uniform set keeps every tree's
// masks sorted by mask, which is what finish_encode expects) and go through the same dictionary / postings build as a fresh set.
// ------------------------------------------------------------------------------------------------------------
namespace td {

// per-tree encoding of tree t of a packed set, splits sorted by mask (a uniform set keeps them sorted by dictionary id)
static void tree_of_set(const HostSet &h, size_t t, TreeEnc &e)
{
    const int W = h.W, n_taxa = h.n_taxa, S = h.n_splits[t];
    const size_t stride = (size_t)std::max(h.max_splits, 1);
    e.rooted = h.tree_rooted[t]; e.n_leaves = h.n_leaves[t];
    e.leafmask.assign(h.leafmask.begin() + t * W, h.leafmask.begin() + (t + 1) * W);
    std::vector<int> order(S);
    for (int s = 0; s < S; s++) order[s] = s;
    const uint64_t *m = &h.masks[t * stride * W];
    std::sort(order.begin(), order.end(), [&](int a, int b) { return mask_cmp(m + (size_t)a * W, m + (size_t)b * W, W) < 0; });
    e.masks.resize((size_t)S * W); e.lens.resize(S);
    for (int s = 0; s < S; s++) { memcpy(&e.masks[(size_t)s * W], m + (size_t)order[s] * W, (size_t)W * 8); e.lens[s] = h.lens[t * stride + order[s]]; }
    e.leaflen.assign(h.leaflen.begin() + t * n_taxa, h.leaflen.begin() + (t + 1) * n_taxa);
    e.sum_len = h.sum_len[t]; e.sum_sq = h.sum_sq[t];
}

// query set followed by reference set as ONE set (joint split dictionary): what the rectangular kernels need.  Both sets must have been
// encoded over the same taxon table (same labels); otherwise encode the two lists of trees with one call.
int concat_sets(const HostSet &a, const HostSet &b, int n_threads, HostSet &out)
{
    if (a.labels != b.labels) {
        set_error("the two sets were encoded over different taxon tables (%d and %d labels): encode their trees with one td_encode_newick call", a.n_taxa, b.n_taxa);
        return TD_ERR_UNSUPPORTED;
    }
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<TreeEnc> enc((size_t)(a.N + b.N));
    parallel_for(a.N + b.N, n_threads, [&](int64_t i, int) {
        if (i < a.N) tree_of_set(a, (size_t)i, enc[(size_t)i]); else tree_of_set(b, (size_t)(i - a.N), enc[(size_t)i]);
    });
    return finish_encode(enc, a.labels, n_threads, out, [](const char *) {});
}

int encode_group(const HostSet &parent, const std::vector<int32_t> &members, int n_threads, HostSet &sub)
{
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int W = parent.W, n_taxa = parent.n_taxa;
    const size_t stride = (size_t)std::max(parent.max_splits, 1);
    std::vector<TreeEnc> enc(members.size());
    for (size_t k = 0; k < members.size(); k++) {
        const size_t t = (size_t)members[k];
        TreeEnc &e = enc[k];
        const int S = parent.n_splits[t];
        e.rooted = parent.tree_rooted[t]; e.n_leaves = parent.n_leaves[t];
        e.leafmask.assign(parent.leafmask.begin() + t * W, parent.leafmask.begin() + (t + 1) * W);
        e.masks.assign(parent.masks.begin() + t * stride * W, parent.masks.begin() + (t * stride + S) * W);
        e.lens.assign(parent.lens.begin() + t * stride, parent.lens.begin() + t * stride + S);
        e.leaflen.assign(parent.leaflen.begin() + t * n_taxa, parent.leaflen.begin() + (t + 1) * n_taxa);
        e.sum_len = parent.sum_len[t]; e.sum_sq = parent.sum_sq[t];
    }
    return finish_encode(enc, parent.labels, n_threads, sub, [](const char *) {});
}

}  // namespace td

// ------------------------------------------------------------------------------------------------------------
// Serialised form of an encoded set (td_treeset_serialize / td_treeset_deserialize): encode once, ship the bytes to the other ranks
// (one broadcast) instead of parsing and encoding the same trees in every process.  Little-endian, this build's layout; a version
// word guards against mixing builds.
// ------------------------------------------------------------------------------------------------------------
namespace td {

namespace {
constexpr uint64_t kSerialMagic = 0x3154455345455254ull;      // "TREESET1"
constexpr uint32_t kSerialVersion = 3;

struct Writer {
    std::vector<unsigned char> *out; size_t size = 0;         // out == nullptr: only measure
    void raw(const void *p, size_t n) { if (out) { const unsigned char *b = (const unsigned char *)p; out->insert(out->end(), b, b + n); } size += n; }
    template <class T> void pod(const T &v) { raw(&v, sizeof v); }
    template <class T, class A> void vec(const std::vector<T, A> &v) { const uint64_t n = v.size(); pod(n); raw(v.data(), n * sizeof(T)); const uint64_t pad = (8 - (n * sizeof(T)) % 8) % 8, z = 0; raw(&z, pad); }
};
struct Reader {
    const unsigned char *p, *end; bool ok = true;
    void raw(void *dst, size_t n) { if (!ok || (size_t)(end - p) < n) { ok = false; return; } memcpy(dst, p, n); p += n; }
    template <class T> void pod(T &v) { raw(&v, sizeof v); }
    template <class T, class A> void vec(std::vector<T, A> &v)
    {
        uint64_t n = 0; pod(n);
        if (!ok || n > (uint64_t)(end - p) / sizeof(T)) { ok = false; return; }
        v.resize(n); raw(v.data(), n * sizeof(T));
        const size_t pad = (8 - (n * sizeof(T)) % 8) % 8;
        if ((size_t)(end - p) < pad) { ok = false; return; }
        p += pad;
    }
};

template <class IO, class HS>
void serial_fields(IO &io, HS &h)
{
    io.pod(h.N); io.pod(h.n_taxa); io.pod(h.W); io.pod(h.max_splits); io.pod(h.max_shared); io.pod(h.max_tile_entries);
    io.pod(h.uniform); io.pod(h.rooted); io.pod(h.dict_size); io.pod(h.dict_shared); io.pod(h.total_splits);
    io.pod(h.heavy_rows); io.pod(h.heavy_events); io.pod(h.heavy_common); io.pod(h.total_common); io.pod(h.mean_common); io.pod(h.sh_stride);
    io.vec(h.n_splits); io.vec(h.n_leaves); io.vec(h.tree_rooted); io.vec(h.n_shared); io.vec(h.leafmask);
    io.vec(h.sum_len); io.vec(h.norm); io.vec(h.sum_sq); io.vec(h.single_l1); io.vec(h.single_l2); io.vec(h.q1); io.vec(h.q2); io.vec(h.q1_light); io.vec(h.q2t);
    io.vec(h.masks); io.vec(h.lens); io.vec(h.ids); io.vec(h.leaflen);
    io.vec(h.shared); io.vec(h.shared_x); io.vec(h.post_tree); io.vec(h.post_end); io.vec(h.post_len);
}
}  // namespace

size_t serialize_set(const HostSet &hs, std::vector<unsigned char> *out)
{
    Writer w{out};
    w.pod(kSerialMagic); w.pod(kSerialVersion); const uint32_t rec = (uint32_t)sizeof(SplitRec); w.pod(rec);
    std::string joined;
    for (size_t i = 0; i < hs.labels.size(); i++) { if (i) joined.push_back('\n'); joined += hs.labels[i]; }
    const std::vector<char> lab(joined.begin(), joined.end());
    w.vec(lab);
    serial_fields(w, hs);
    return w.size;
}

int deserialize_set(const void *data, size_t bytes, HostSet &hs)
{
    Reader r{(const unsigned char *)data, (const unsigned char *)data + bytes};
    uint64_t magic = 0; uint32_t version = 0, rec = 0;
    r.pod(magic); r.pod(version); r.pod(rec);
    if (!r.ok || magic != kSerialMagic || version != kSerialVersion || rec != sizeof(SplitRec)) { set_error("not a serialised tree set of this library version"); return TD_ERR_ARG; }
    std::vector<char> lab; r.vec(lab);
    serial_fields(r, hs);
    if (!r.ok || r.p != r.end) { set_error("serialised tree set is truncated or has trailing bytes"); return TD_ERR_ARG; }
    hs.labels.clear();
    for (size_t b = 0; b < lab.size();) { size_t e = b; while (e < lab.size() && lab[e] != '\n') e++; hs.labels.emplace_back(lab.data() + b, e - b); b = e + 1; }
    // consistency of the sizes the kernels index with (a corrupted buffer must not become an out-of-bounds read on the device)
    const size_t N = (size_t)hs.N, S = (size_t)std::max(hs.max_splits, 1);
    auto per_tree = [&](size_t n) { return n == N; };
    const bool sane = hs.N >= 1 && hs.n_taxa >= 0 && hs.W >= 1 && hs.W >= (hs.n_taxa + 63) / 64 && hs.sh_stride >= 2 && hs.max_splits >= 0
        && (int)hs.labels.size() == hs.n_taxa
        && per_tree(hs.n_splits.size()) && per_tree(hs.n_shared.size()) && per_tree(hs.tree_rooted.size()) && hs.leafmask.size() >= N * (size_t)hs.W
        && per_tree(hs.sum_len.size()) && per_tree(hs.norm.size()) && per_tree(hs.sum_sq.size()) && per_tree(hs.n_leaves.size()) && per_tree(hs.single_l1.size()) && per_tree(hs.single_l2.size())
        && per_tree(hs.q1.size()) && per_tree(hs.q2.size()) && per_tree(hs.q1_light.size()) && per_tree(hs.q2t.size())
        && hs.leaflen.size() >= N * (size_t)hs.n_taxa && hs.lens.size() >= N * S && hs.masks.size() >= N * S * (size_t)hs.W && hs.ids.size() >= N * S
        && hs.shared.size() == N * (size_t)hs.sh_stride && hs.shared_x.size() == hs.shared.size()
        && hs.post_end.size() == hs.post_tree.size() && hs.post_len.size() == hs.post_tree.size();
    if (!sane) { set_error("serialised tree set is inconsistent"); return TD_ERR_ARG; }
    return TD_OK;
}

}  // namespace td
