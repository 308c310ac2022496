// td_internal.h -- internal structures shared by the host encoder, the C-ABI layer and the kernels.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <memory>
#include <new>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/treedist_cuda.h"

namespace td {

constexpr uint32_t kSentinelId = 0xFFFFFFFFu;
constexpr int kTile = 64;          // trees per tile edge; N is padded to a multiple of this on the device

// one entry of a per-tree sorted split list as the merge kernels see it (16 B -> one LDS.128 / LDG.128)
struct SplitRec {
    uint32_t id;
    uint32_t pad;      // compact index among the shared splits (column of the incidence matrix)
    double len;
};
static_assert(sizeof(SplitRec) == 16, "SplitRec must be 16 bytes");

// std::vector whose resize() leaves trivially constructible elements uninitialised: the big arrays of a set are written in full by the
// encoder's workers (first touch in parallel) instead of being zero-filled by one thread first (7 MB, ~1 ms of the bench set's encode)
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    default_init_allocator() = default;
    template <class U> default_init_allocator(const default_init_allocator<U> &) noexcept {}
    template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
template <class T> using raw_vector = std::vector<T, default_init_allocator<T>>;

struct HostSet {
    int64_t N = 0;
    int32_t n_taxa = 0, W = 1;
    int32_t max_splits = 0, max_shared = 0;
    int32_t max_tile_entries = 0;              // largest sum of shared-list lengths over a 64-tree tile
    bool uniform = false;
    int32_t rooted = 0;
    int64_t dict_size = 0, dict_shared = 0, total_splits = 0;
    std::vector<std::string> labels;           // sorted taxon table

    // per tree (N entries)
    std::vector<int32_t> n_splits, n_leaves, tree_rooted, n_shared;
    std::vector<uint64_t> leafmask;            // [N][W]
    std::vector<double> sum_len, norm;         // all edges incl. leaves: sum l, sqrt(sum l^2)
    std::vector<double> sum_sq;                // sum l^2 (kept beside its root so that a leaf-set group re-encodes to the same norm bits)
    std::vector<double> single_l1, single_l2;  // sums over splits held by this tree only: sum |l|, sum l^2
    std::vector<double> q1, q2;                // sums over ALL internal splits: sum |l|, sum l^2
    std::vector<double> q1_light;              // q1 minus the lengths of the tree's HEAVY shared splits (compact index < heavy_rows)
    int32_t heavy_rows = 0;                    // shared splits (by descending multiplicity) that the hybrid tile kernel takes as matrix rows
    int64_t heavy_events = 0;                  // sum over the heavy splits of m (m - 1) / 2: events the hybrid form does not generate
    double heavy_common = 0.0;                 // expected common splits per pair covered by the heavy rows
    std::vector<double> q2t;                   // q2 + sum of squared leaf-edge lengths (dot-product form of the Euclidean kernel)
    int64_t total_common = 0;                  // sum over all pairs of their number of common splits (exact)
    double mean_common = 0.0;                  // expected number of common splits of a random pair of the set

    // full lists, rows sorted by split id (uniform sets) or by mask (otherwise); stride max_splits
    raw_vector<uint64_t> masks;                // [N][max_splits][W]
    raw_vector<double> lens;                   // [N][max_splits]
    raw_vector<uint32_t> ids;                  // [N][max_splits], kSentinelId padded (uniform sets only)
    raw_vector<double> leaflen;                // [N][n_taxa]

    // shared-split lists (splits that occur in >= 2 trees), sentinel padded, stride sh_stride
    int32_t sh_stride = 0;
    raw_vector<SplitRec> shared;               // [N][sh_stride]
    raw_vector<int32_t> shared_x;              // [N][sh_stride] position of the entry in the postings, -1 = none

    // postings of the shared splits: entries grouped by split, trees ascending inside a group.  post_end[x] is the end of x's
    // group, so "all later holders of the split of entry x" is the range (x, post_end[x]).
    std::vector<int32_t> post_tree, post_end;  // [n_post]
    std::vector<double> post_len;              // [n_post]
};

struct DeviceSet;   // defined in api.cu

void set_error(const char *fmt, ...);
const char *last_error();

// host_encoder.cpp
int encode_newick(const char *const *newick, int64_t n_trees, int n_threads, HostSet &hs);
void encoder_trim();                       // drops the per-tree scratch kept between encodes
// dst[0, bytes) = src[0, bytes), then dst[bytes, bytes + pad_bytes) = pad_value; all jobs spread over the encoder's worker pool
struct CopyJob { void *dst; const void *src; size_t bytes, pad_bytes; int pad_value; };
void parallel_copy(const std::vector<CopyJob> &jobs);
int encode_kendall_colijn(const char *const *newick, int64_t n_trees, double lambda, int n_threads, HostSet &hs);
int newick_labels(const char *newick, std::vector<std::string> &labels, int &rooted);
int newick_prune(const char *newick, const std::vector<std::string> &keep, std::string &out);
// leaf-set group of a non-uniform set: the member trees (ascending indices into `parent`) re-encoded as a set of their own - uniform, with
// its own split dictionary - for the fast kernels
int concat_sets(const HostSet &a, const HostSet &b, int n_threads, HostSet &out);
int encode_group(const HostSet &parent, const std::vector<int32_t> &members, int n_threads, HostSet &sub);
size_t serialize_set(const HostSet &hs, std::vector<unsigned char> *out);      // out == nullptr: size only
int deserialize_set(const void *data, size_t bytes, HostSet &hs);

}  // namespace td

struct td_treeset {
    td::HostSet host;
    std::mutex mu;                             // guards `devs`
    std::vector<td::DeviceSet *> devs;         // indexed by CUDA device ordinal; the set is replicated per GPU
    // Leaf-set groups of a NON-uniform set (built on first use): trees with the same leaf set and rootedness, when there are enough of
    // them, form a uniform sub-set served by the fast kernels; only pairs across groups take the warp-per-pair kernels
    // (treeCl/treedist.py:63-68 prunes only the pairs whose leaf sets differ).
    struct Group { std::vector<int32_t> members; td_treeset *sub = nullptr; };
    std::mutex groups_mu;
    bool groups_built = false;
    std::vector<Group> groups;
    std::vector<int32_t> group_of;             // [N] index into `groups`, -1: in no group
    // The pairs ACROSS groups, enumerated without visiting the others: for row i the wanted columns j > i are cross_cols[cross_first[i] ..]
    // (cross_first[i] < 0: every j > i, the row of a tree in no group), cross_prefix[i] = wanted pairs of the rows before i.  Empty when
    // the lists would be too large (the kernels then skip same-group pairs chunk by chunk).
    std::vector<int32_t> cross_cols;
    std::vector<int64_t> cross_first, cross_prefix;
    ~td_treeset();
};
