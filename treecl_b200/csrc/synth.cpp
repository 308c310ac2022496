// synth.cpp -- fast generator of the synthetic tree sets of SURVEY.md 8(d) (bench / test tooling, not on the product path).
//
//   kind 0  uniform     treeCl.tree.RandomTree.new(n, rooted=False) (treeCl/tree.py:1234-1276): 3-leaf star, every further leaf attached
//                       to a uniformly random existing edge -> binary, trifurcating root, n - 3 internal edges; lengths i.i.d. Exp(1)
//                       (randomise_branch_lengths defaults, treeCl/tree.py:1003-1025)
//   kind 1  structured  Simulator.generate_class_trees (treeCl/simulator.py:127-190): K class trees = master tree + class_nni random
//                       NNIs; gene tree k = class tree k % K + Poisson(lam) further NNIs, lengths redrawn
// Same models as treecl_b200/synth.py (whose NumPy streams it does not reproduce); 100,000 x 200-taxon trees take well under a
// second here instead of minutes.  Lengths are printed with 17 significant digits: every consumer parses identical doubles.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/treedist_synth.h"
#include "td_internal.h"

namespace {

struct Rng {                              // splitmix64-seeded xoshiro256**
    uint64_t s[4];
    static uint64_t mix(uint64_t &x) { uint64_t z = (x += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    Rng(uint64_t seed, uint64_t stream) { uint64_t x = seed * 0xD1342543DE82EF95ull + stream; for (auto &v : s) v = mix(x); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next()
    {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }          // [0, 1)
    int below(int n) { return (int)(((unsigned __int128)next() * (uint64_t)n) >> 64); }
    double exponential() { return -std::log(1.0 - uniform()); }
    int poisson(double lam) { const double L = std::exp(-lam); int k = 0; double p = 1.0; do { k++; p *= uniform(); } while (p > L); return k - 1; }
};

struct Topo {                             // node 0 = root (a trifurcation); leaves carry a taxon index
    std::vector<int> parent, taxon;
    std::vector<std::vector<int>> children;
    int add(int p, int tx)
    {
        parent.push_back(p); taxon.push_back(tx); children.emplace_back();
        const int k = (int)parent.size() - 1;
        if (p >= 0) children[p].push_back(k);
        return k;
    }
};

Topo random_topology(int n, Rng &rng)
{
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    for (int i = n - 1; i > 0; i--) std::swap(order[i], order[rng.below(i + 1)]);
    Topo t; t.add(-1, -1);
    for (int k = 0; k < 3 && k < n; k++) t.add(0, order[k]);
    for (int k = 3; k < n; k++) {
        const int v = 1 + rng.below((int)t.parent.size() - 1);            // the edge above node v, uniform over the existing edges
        const int p = t.parent[v], mid = t.add(-1, -1);
        t.parent[mid] = p;
        *std::find(t.children[p].begin(), t.children[p].end(), v) = mid;
        t.children[mid].push_back(v); t.parent[v] = mid;
        t.add(mid, order[k]);
    }
    return t;
}

void nni(Topo &t, Rng &rng)                // one random nearest-neighbour interchange around a random internal edge
{
    std::vector<int> internal;
    for (int v = 1; v < (int)t.parent.size(); v++) if (!t.children[v].empty()) internal.push_back(v);
    if (internal.empty()) return;
    const int v = internal[rng.below((int)internal.size())], p = t.parent[v];
    std::vector<int> sibs;
    for (int c : t.children[p]) if (c != v) sibs.push_back(c);
    const int a = sibs[rng.below((int)sibs.size())], b = t.children[v][rng.below((int)t.children[v].size())];
    *std::find(t.children[p].begin(), t.children[p].end(), a) = b;
    *std::find(t.children[v].begin(), t.children[v].end(), b) = a;
    t.parent[a] = v; t.parent[b] = p;
}

void write_newick(const Topo &t, int width, Rng &rng, std::string &out)
{
    out.clear();
    char buf[64];
    std::vector<std::pair<int, int>> stack{{0, 0}};
    while (!stack.empty()) {
        auto [v, i] = stack.back(); stack.pop_back();
        const auto &ch = t.children[v];
        if (ch.empty()) {
            out.append(buf, (size_t)snprintf(buf, sizeof buf, "t%0*d:%.17g", width, t.taxon[v], rng.exponential()));
            continue;
        }
        if (i == 0) out.push_back('(');
        if (i < (int)ch.size()) {
            if (i > 0) out.push_back(',');
            stack.emplace_back(v, i + 1); stack.emplace_back(ch[i], 0);
        } else {
            out.push_back(')');
            if (v != 0) out.append(buf, (size_t)snprintf(buf, sizeof buf, ":%.17g", rng.exponential()));
        }
    }
    out.push_back(';');
}

}  // namespace

extern "C" {

int td_synth_trees(int kind, int64_t n_trees, int n_taxa, uint64_t seed, int n_classes, int class_nni, double lam, int n_threads,
                   char **out_buffer, int64_t *out_bytes)
{
    using namespace td;
    if (!out_buffer || !out_bytes) { set_error("null output pointer"); return TD_ERR_ARG; }
    *out_buffer = nullptr; *out_bytes = 0;
    if (n_trees < 1 || n_taxa < 3 || (kind != 0 && kind != 1)) { set_error("td_synth_trees: need n_trees >= 1, n_taxa >= 3, kind 0 or 1"); return TD_ERR_ARG; }
    if (kind == 1 && (n_classes < 1 || class_nni < 0 || !(lam >= 0.0) || lam > 500.0)) { set_error("td_synth_trees: bad class parameters"); return TD_ERR_ARG; }
    int width = 3; for (int x = n_taxa - 1; x >= 1000; x /= 10) width++;
    std::vector<Topo> classes;
    if (kind == 1) {
        Rng r0(seed, 0xC1A55ull << 32);
        const Topo master = random_topology(n_taxa, r0);
        for (int c = 0; c < n_classes; c++) { Topo t = master; for (int k = 0; k < class_nni; k++) nni(t, r0); classes.push_back(std::move(t)); }
    }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (n_trees + 63) / 64));
    std::vector<std::string> text((size_t)n_trees);
    std::atomic<int64_t> next{0};
    std::atomic<bool> failed{false};
    auto work = [&] {
        try {
            for (;;) {
                const int64_t b = next.fetch_add(64);
                if (b >= n_trees) break;
                for (int64_t k = b; k < std::min(n_trees, b + 64); k++) {
                    Rng rng(seed, (uint64_t)k + 1);
                    if (kind == 0) { const Topo t = random_topology(n_taxa, rng); write_newick(t, width, rng, text[(size_t)k]); }
                    else {
                        Topo t = classes[(size_t)(k % n_classes)];
                        for (int q = rng.poisson(lam); q > 0; q--) nni(t, rng);
                        write_newick(t, width, rng, text[(size_t)k]);
                    }
                }
            }
        } catch (...) { failed = true; }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; i++) th.emplace_back(work);
    work();
    for (auto &x : th) x.join();
    if (failed) { set_error("out of memory while generating trees"); return TD_ERR_NOMEM; }
    size_t total = 0;
    for (auto &s : text) total += s.size() + 1;
    char *buf = (char *)malloc(total);
    if (!buf) { set_error("out of memory"); return TD_ERR_NOMEM; }
    char *p = buf;
    for (auto &s : text) { memcpy(p, s.c_str(), s.size() + 1); p += s.size() + 1; }
    *out_buffer = buf; *out_bytes = (int64_t)total;
    return TD_OK;
}

void td_synth_free(char *buffer) { free(buffer); }

}  // extern "C"
