// kernels.h -- launch interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <set>
#include <utility>

#include <cstdint>

#include "td_internal.h"

namespace td {

// Opt a kernel in to the device's full dynamic shared memory, once per (kernel, device).  Every launcher calls this instead of setting the
// attribute to its own launch's size: the attribute belongs to the function, not to the launch, so two host threads launching the same
// kernel for sets of different sizes could shrink it under each other's feet ("invalid argument" at launch).
inline cudaError_t allow_full_smem(const void *kernel)
{
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> guard(mu);
    if (done.count({kernel, dev})) return cudaSuccess;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, kernel)) != cudaSuccess) return e;
    int optin = 0;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes)) != cudaSuccess) return e;
    done.insert({kernel, dev});
    return cudaSuccess;
}


struct PairParams {
    const double *leafT;      // [n_k][ldT] taxon-major leaf-edge lengths (zero padded)
    int64_t ldT;
    const SplitRec *shared;   // [Npad][sh_stride] sorted shared-split lists, sentinel padded
    int sh_stride;
    const int32_t *nsplits;   // [Npad] internal edges per tree
    const int32_t *nshared;   // [Npad] entries of the tree's shared-split list
    const double *sum_len, *norm, *single_l1, *single_l2;   // [Npad]
    const double *q1, *q2;    // [Npad] totals over ALL internal edges: sum |l|, sum l^2 (hash-join kernel)
    int table_slots, tiles_x; // hash-join kernel: power-of-two table size; tile columns of the grid
    int row_table_slots;      // persistent hash-join kernel: table of the 64 row trees
    long long n_tiles;        // persistent hash-join kernel: length of the tile list
    int2 *wl_pairs;           // persistent hash-join kernel: pairs to recompute exactly (work list)
    unsigned *wl_count; unsigned wl_capacity;
    // postings pipeline (kernels_pairsplit.cu): the common splits of all pairs ("events") are enumerated from the postings of
    // the shared splits and binned by 64 x 32 tile; events of tile t live in [tile_off[t], tile_off[t + 1])
    const int32_t *post_tree, *post_end;   // [n_post]
    const double *post_len;                // [n_post]
    int64_t n_post;
    const double *q2t;        // [Npad] sum of squares of ALL edge lengths (internal + leaf)
    const void *tmap_a, *tmap_b;   // host copies of the CUtensorMaps over leafT (make_leaf_tensor_maps)
    void *ev;                 // event records (layout depends on the metric mask, see kernels_pairsplit.cu)
    const int32_t *shared_x;  // [Npad][sh_stride] position of every shared-list entry in the postings (-1: none)
    uint2 *seg;               // [Npad / 16][col_tiles_total] (offset, count) of the events of (row group, column tile)
    uint32_t *ev_cursor;      // next free record of the event buffer
    // per tile row (four row groups): CTAs of the events kernel that have finished their share of it (zeroed before the launch).  Set: a
    // tile waits for its tile row to reach ready_target instead of for the whole events grid; nullptr: grid-wide dependency
    unsigned *ready; unsigned ready_target;
    int col_tiles_total, rg0; // ceil(N / 32); first row group of this launch
    uint32_t ev_capacity;
    unsigned long long *trace;   // debug builds (-DPS_TRACE): 8 clock stamps per tile
    int64_t N;                // trees
    int n_k;                  // taxa padded to the k-chunk
    int n_taxa;               // rows of leafT beyond n_taxa are zero: the k-loops may stop there
    int heavy_rows;           // postings pipeline: shared splits with compact index < heavy_rows are rows of the matrix, not events
    int k_rows;               // tile kernel of the postings pipeline: rows of the matrix behind the tensor maps (n_taxa, or n_k + shared splits)
    int64_t row_begin, row_end, p_base;
    int64_t col_offset, n_cols, col_end;   // rectangular variant: columns are trees [col_offset, col_offset + n_cols)
    int tile_row0, col_tile0;
    int tri_rows;             // postings pipeline: tile rows of this launch
    int normalise;
    double *out[3];           // rf, wrf, euc (device)
    // second row block of the same launch set (multi-GPU folded blocks: a rank's long-row and short-row block share ONE events kernel and
    // ONE exact fix-up kernel).  Only those two kernels look at these fields; row_end2 <= row_begin2: no second block
    int64_t row_begin2, row_end2, p_base2;
    int rg0b, groups_a;       // events kernel: row groups [0, groups_a) of the grid belong to block A (from rg0), the rest to block B (from rg0b)
    double *out2[3];          // fix-up kernel: pairs with i >= row_begin2 write here, relative to p_base2
};

size_t pairwise_smem_bytes(int tile, int sh_stride, bool weighted);
int pairwise_kc();
cudaError_t launch_pairwise(uint32_t mask, bool rect, PairParams p, int64_t tile_rows, int64_t tile_cols, int tile,
                            cudaStream_t st);
int pairjoin_table_slots(int max_shared);
int pairjoin_row_table_slots(int max_tile_entries);
size_t pairjoin_smem_bytes(uint32_t mask, int table_slots);
size_t pairjoin_persistent_smem_bytes(uint32_t mask, int table_slots, int sh_stride);
cudaError_t launch_pairjoin(uint32_t mask, bool rect, PairParams p, int n_sms, bool persistent, cudaStream_t st);
cudaError_t launch_exact_fixup(uint32_t mask, bool rect, const PairParams &p, int n_sms, cudaStream_t st);
size_t pairsplit_dense_smem_bytes(uint32_t mask);
int pairsplit_record_bytes(uint32_t mask);
size_t pairsplit_events_smem_bytes(uint32_t mask, int sh_stride, int col_tiles_total);
int pairsplit_row_group();
cudaError_t make_leaf_tensor_maps(const double *leafT, int64_t ldT, int n_k, void *tmap_a, void *tmap_b);
cudaError_t launch_build_split_rows(const SplitRec *shared, int sh_stride, int64_t N, int64_t ld, double *rows, int max_rows, cudaStream_t st);
long long pairsplit_tiles(bool rect, PairParams &p);
cudaError_t launch_pairsplit(uint32_t mask, bool rect, const PairParams &p, int n_sms, cudaStream_t st, int *launches, const PairParams *second = nullptr);
cudaError_t launch_pairsplit_tiles(uint32_t mask, bool rect, const PairParams &p, cudaStream_t st);
cudaError_t launch_scatter_group(const double *tmp, double *out, const int32_t *members, int64_t n_local, int64_t a0, int64_t a1, int64_t N,
                                 int64_t p_base, cudaStream_t st);
cudaError_t launch_square_rows(const double *cond, double *sq, int64_t N, int64_t r0, int64_t r1, cudaStream_t st);
cudaError_t launch_transpose_leaf(const double *src, double *dst, int64_t N, int n_taxa, int64_t ldT, cudaStream_t st);

struct GeoParams {
    const uint32_t *ids;      // [N][stride] sorted split ids, sentinel padded
    const double *lens;       // [N][stride]
    const uint64_t *masks;    // [N][stride][W]
    const int32_t *nsplits;   // [N]
    const double *leaflen;    // [N][n_taxa] tree-major
    const double *norm;       // [N]
    const uint64_t *leafmask; // [N][W] taxa present in each tree (general kernel)
    const int32_t *rooted;    // [N]
    const int32_t *group_of;  // [N] leaf-set group served by uniform kernels (-1: none), or nullptr: the warp-per-pair kernels skip same-group pairs
    // pairs across leaf-set groups as lists (td_treeset::cross_*): when cross_prefix is set, [p_begin, p_end) counts WANTED pairs
    const int64_t *cross_prefix, *cross_first;
    const int32_t *cross_cols;
    int64_t N;
    int stride, n_taxa, W;
    int flow_rows;                       // max internal edges of any tree of the set: the flow matrix is flow_rows x (flow_rows|1)
    int64_t p_begin, p_end, p_base;      // condensed pair range (triangular) ...
    int64_t rect_split, rect_cols;       // ... or rectangular: rows [0,split) x cols [split, split+cols); p = a*cols + b
    int rect;
    int normalise;
    double *out;              // geodesic kernel (uniform sets)
    double *outs[4];          // general kernel: one output per metric of metric_mask
    uint32_t metric_mask;
    int min_overlap;
    double fail_value;
    unsigned long long *work_counter;    // dynamic pair scheduler
    double *flow_scratch;                // per-warp flow matrices in global memory (GeoPlan::scratch_bytes), or nullptr: in shared memory
    unsigned long long *stats;           // [4]: pairs, maxflow solves, final ratios, augmentations
    int *error_flag;                     // bit 0: an iteration cap was hit; bit 1: rooted vs unrooted pair
};

// Launch shape of the warp-per-pair GTP kernels.  The fp64 flow matrix of a warp (flow_rows x (flow_rows|1)) lives in shared memory when
// that does not cost residency, else in a per-warp slice of global memory (scratch_bytes > 0: the caller provides GeoParams::flow_scratch):
// since the BFS reads bit rows held in registers, the matrix is touched only by the greedy start and along augmenting paths.
struct GeoPlan { int warps = 0, blocks = 0; size_t smem = 0, scratch_bytes = 0; };
// cap: 32, 64 or 128 (max internal edges per tree the warp-per-pair GTP kernel holds in registers/shared memory)
cudaError_t plan_geodesic(const GeoParams &p, int cap, int n_sms, GeoPlan &plan);
cudaError_t launch_geodesic(const GeoParams &p, int cap, const GeoPlan &plan, cudaStream_t st);
// pairs with different leaf sets (non-uniform sets): restrict-then-join, all four metrics
cudaError_t plan_general(const GeoParams &p, int cap, int n_sms, GeoPlan &plan);
cudaError_t launch_general(const GeoParams &p, int cap, const GeoPlan &plan, cudaStream_t st);

// trees with more than 128 internal edges (kernels_bigtrees.cu): all four metrics, same or different leaf sets, per-warp global scratch
size_t bigtree_warp_bytes(int max_splits, int W, int n_taxa);
int bigtree_max_splits();
int bigtree_blocks(int64_t pairs, int n_sms, size_t warp_bytes, size_t scratch_budget);
cudaError_t launch_bigtree(const GeoParams &g, int general, int max_splits, unsigned char *scratch, int blocks, cudaStream_t st);

// classical MDS on the device-resident matrix (kernels_embed.cu)
size_t cmds_workspace_doubles(int64_t n);
int cmds_block();
cudaError_t run_affinity(double *sq, int64_t n, int prune_k, int prune_and, int scale_k, double diag, double *vec, cudaStream_t st);
cudaError_t run_spectral(double *sq, int64_t n, int dims, int prune_k, int prune_and, int scale_k, double *work, double *h_pinned, double *d_coords,
                         double *eigenvalues, int max_iter, double tol, int *iterations, double *residual, cudaStream_t st);
cudaError_t run_cmds(double *sq, int64_t n, int dims, double *work, double *h_pinned, double *d_coords, double *eigenvalues, int max_iter, double tol,
                     int *iterations, double *residual, cudaStream_t st);

// int8 tcgen05 split-incidence RF (kernels_incidence.cu)
size_t incidence_smem_bytes();
int incidence_tile();
int incidence_bk();
cudaError_t launch_build_incidence(const SplitRec *shared, int sh_stride, int64_t N, int64_t ldx, int8_t *X, cudaStream_t st);
int incidence_half();
cudaError_t make_incidence_tensor_map(const int8_t *X, int64_t rows, int64_t ldx, void *tmap_out, int box_rows);     // box_rows 0: a whole tile
cudaError_t launch_rf_incidence(const void *tmap, const void *tmap_half, const int32_t *nsplits, int64_t N, int64_t d_pad, int64_t row_begin, int64_t row_end,
                                int64_t p_base, int normalise, int out_u16, void *out, cudaStream_t st);

uint64_t launch_counter_add(uint64_t n);

}  // namespace td
