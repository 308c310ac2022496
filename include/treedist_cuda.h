/*
 * treedist_cuda.h -- C ABI of libtreedist_cuda.so, the B200-native engine for treeCl's all-pairs inter-tree
 * distance matrix (Robinson-Foulds, weighted RF, branch-length Euclidean, BHV geodesic).
 *
 * What this boundary replaces in the reference (/root/reference, treeCl 0.1.41).  The reference has no C ABI
 * for this path; its native seam is the Cython module `tree_distance` (PyPI, un-vendored), reached per PAIR:
 *     PhyloTree(bytes newick, bool rooted)                       treeCl/tree.py:830-845
 *     getRobinsonFouldsDistance / getWeightedRobinsonFouldsDistance /
 *     getEuclideanDistance / getGeodesicDistance (t1, t2, normalise) -> float
 *                                                                treeCl/treedist.py:7-8,70,111-118
 * driven by the per-pair task + JobHandler loop                  treeCl/tasks.py:41-75,648-678,
 *                                                                treeCl/parutils.py:113-254,
 *                                                                treeCl/collection.py:421-456.
 * This library is bound at the level of the WHOLE matrix instead: trees are encoded once (td_encode_newick
 * replaces N*(N-1) PhyloTree constructions), and one call fills the scipy-condensed vector that
 * collection.py:456 hands to squareform()/DistanceMatrix.from_array.
 *
 * Conventions: every function returns a td_status (0 = ok); td_last_error() gives a thread-local message.
 * All pointers are plain host or device pointers; no torch / C++ types cross this boundary.  There is NO CPU
 * fallback: compute entry points return TD_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef TREEDIST_CUDA_H
#define TREEDIST_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct td_treeset td_treeset;

typedef enum {
    TD_RF = 0,   /* getRobinsonFouldsDistance         (treeCl/treedist.py:113) */
    TD_WRF = 1,  /* getWeightedRobinsonFouldsDistance (treeCl/treedist.py:114) */
    TD_EUC = 2,  /* getEuclideanDistance              (treeCl/treedist.py:111) */
    TD_GEO = 3   /* getGeodesicDistance               (treeCl/treedist.py:112) */
} td_metric;

#define TD_MASK(metric) (1u << (metric))

typedef enum {
    TD_OK = 0,
    TD_ERR_PARSE = 1,          /* malformed Newick           -> Python ValueError, as tree_distance raises */
    TD_ERR_MIXED_ROOTING = 2,  /* rooted vs unrooted pair    -> undefined in the reference; refused here  */
    TD_ERR_ARG = 3,
    TD_ERR_NOMEM = 4,
    TD_ERR_CUDA = 5,
    TD_ERR_NO_DEVICE = 6,      /* no usable GPU: there is deliberately no CPU fallback                      */
    TD_ERR_UNSUPPORTED = 7,    /* e.g. geodesic on trees with more internal edges than the kernel supports  */
    TD_ERR_NOT_UPLOADED = 8
} td_status;

typedef struct {
    int64_t n_trees;
    int32_t n_taxa;          /* size of the shared (sorted) taxon index = union of all leaf labels           */
    int32_t words;           /* W = ceil(n_taxa / 64) 64-bit words per split mask                            */
    int32_t max_splits;      /* max internal edges of any tree                                               */
    int32_t max_shared;      /* max per-tree count of splits that also occur in some other tree              */
    int32_t uniform;         /* 1: all trees share one leaf set and one rootedness -> fast kernels           */
    int32_t rooted;          /* valid when uniform                                                           */
    int64_t dict_size;       /* D: distinct splits over the whole set (uniform sets only, else 0)            */
    int64_t dict_shared;     /* distinct splits held by >= 2 trees                                           */
    int64_t total_splits;
    int32_t device;          /* device the set is resident on, -1 if not uploaded                            */
    int32_t n_postings;      /* entries of the shared-split postings (= sum over trees of their shared-list lengths)      */
    int32_t heavy_rows;      /* shared splits (by descending multiplicity) the hybrid tile kernel takes as matrix rows     */
    int32_t n_groups;        /* non-uniform sets: leaf-set groups served by the uniform kernels (-1 until the first matrix call)   */
    double mean_common;      /* expected number of common splits of a random pair of the set                              */
    double heavy_common;     /* ... of which the heavy rows cover                                                          */
} td_info;

/* ---- host side: split encoder ------------------------------------------------------------------------ */

/* Parse n_trees Newick strings and encode every tree's bipartitions as canonical taxon bitmasks + fp64 branch
 * lengths over a shared taxon index (replaces PhyloTree(newick, rooted), treeCl/tree.py:830-845; rooted <=> the
 * root has exactly two children, treeCl/tree.py:858-862).  n_threads <= 0: use all host cores. */
int td_encode_newick(const char *const *newick, int64_t n_trees, int n_threads, td_treeset **out);

/* Same, for n_trees NUL-terminated Newick strings stored back to back in one buffer of total_bytes bytes (saves the caller
 * building a pointer table: the Python binding joins its strings once). */
int td_encode_newick_joined(const char *buffer, int64_t total_bytes, int64_t n_trees, int n_threads, td_treeset **out);

/* Kendall-Colijn metric (treeCl/utils/kendallcolijn.py).  Encodes every tree as its vector
 *     v = (1 - lambda) m + lambda M        KendallColijn._get_vectors / get_vector, kendallcolijn.py:53-79
 * (m, M: edges / path length from the root to the MRCA of every leaf pair in sorted label order, then 1 / edge length per leaf) and
 * returns a set whose Euclidean matrix - td_pairwise(ts, TD_MASK(TD_EUC), ...) - is KendallColijn.get_distance (kendallcolijn.py:81-95)
 * for all pairs.  All trees must have one leaf set (the reference prunes differing pairs to their intersection first: do that with
 * td_newick_prune and encode the pruned pair); other sets are refused with TD_ERR_UNSUPPORTED.  td_info.n_taxa of the result is the
 * vector length n(n-1)/2 + n.  Only TD_EUC (and TD_WRF = the L1 distance of the vectors) are meaningful on such a set. */
int td_encode_kendall_colijn(const char *const *newick, int64_t n_trees, double lambda, int n_threads, td_treeset **out);

/* Encode once, run everywhere (SURVEY 8(b) td_treeset_from_soa, 8(e)): the encoded set as one flat byte string.  Rank 0 encodes and
 * serialises, the bytes travel by one broadcast (torch.distributed / NCCL / MPI - the library does not care), every other rank
 * deserialises instead of parsing and encoding the same Newick strings again.  td_treeset_serialize(ts, NULL, 0, &need) sizes the
 * buffer.  The format is this build's memory layout behind a version word: not a storage format. */
int td_treeset_serialize(const td_treeset *ts, void *dst, int64_t cap_bytes, int64_t *need_bytes);
int td_treeset_deserialize(const void *data, int64_t bytes, td_treeset **out);

void td_treeset_free(td_treeset *ts);
int td_treeset_info(const td_treeset *ts, td_info *info);

/* Introspection (tests, INTEGRATION.md examples).  `what`: see TD_GET_*.  Copies at most cap_bytes, returns the
 * full size through *need_bytes. */
enum {
    TD_GET_LABELS = 0,       /* '\n'-joined sorted taxon labels (char)                                       */
    TD_GET_NSPLITS = 1,      /* int32[n_trees]                                                              */
    TD_GET_MASKS = 2,        /* uint64[n_trees][max_splits][W], rows sorted by split id, zero padded         */
    TD_GET_LENGTHS = 3,      /* double[n_trees][max_splits]                                                 */
    TD_GET_LEAFLEN = 4,      /* double[n_trees][n_taxa]                                                     */
    TD_GET_IDS = 5,          /* uint32[n_trees][max_splits] dictionary ids (uniform sets), 0xFFFFFFFF padded */
    TD_GET_LEAFMASK = 6,     /* uint64[n_trees][W]                                                          */
    TD_GET_ROOTED = 7        /* int32[n_trees]                                                              */
};
int td_treeset_get(const td_treeset *ts, int what, void *dst, int64_t cap_bytes, int64_t *need_bytes);

/* Host-side helpers that mirror treeCl.Tree for this path (no GPU needed):
 *   td_newick_labels : '\n'-joined sorted leaf labels + rootedness   (Tree.labels / Tree.rooted, tree.py:792-796,858-862)
 *   td_newick_prune  : Tree.prune_to_subset(subset).newick            (tree.py:989-1001; merged edge lengths are summed) */
int td_newick_labels(const char *newick, char *dst, int64_t cap_bytes, int64_t *need_bytes, int *rooted);
int td_newick_prune(const char *newick, const char *const *keep, int64_t n_keep, char *dst, int64_t cap_bytes,
                    int64_t *need_bytes);

/* ---- device side --------------------------------------------------------------------------------------- */

int td_device_count(int *count);

/* Make the encoded set resident in the HBM of `device` (SoA: taxon-major leaf lengths, per-tree sorted split
 * ids + lengths, split masks, per-tree aggregates).  `stream` is a cudaStream_t (NULL = legacy default). */
int td_upload(td_treeset *ts, int device, void *stream);

/* Number of condensed entries belonging to rows [row_begin, row_end) of the upper triangle. */
int64_t td_condensed_offset(int64_t n_trees, int64_t row);

/* Kernels only.  For every metric bit set in metric_mask (TD_MASK(TD_RF) | ...), fill
 *     d_out[metric][p - td_condensed_offset(N,row_begin)]   for all pairs p=(i,j), row_begin <= i < row_end, i < j,
 * in scipy-condensed order (itertools.combinations order, treeCl/tasks.py:651-653).  d_out[m] are DEVICE pointers
 * (double; entries for metrics not requested are ignored).  Asynchronous on `stream`.  Multi-GPU: every rank
 * uploads the same set and calls this with its own row range; no collective is involved.
 * Kernel choice is internal (postings pipeline / merge kernel / int8 incidence contraction for RF alone on densely sharing sets /
 * restrict-then-join for differing leaf sets); all give the same values.  Pairs whose leaf sets differ follow treeCl/treedist.py:63-68: fewer than min_overlap common leaves ->
 * overlap_fail_value, else both trees are restricted to the intersection first. */
int td_pairwise_device(td_treeset *ts, uint32_t metric_mask, int normalise, int min_overlap,
                       double overlap_fail_value, int64_t row_begin, int64_t row_end, void *const d_out[4],
                       void *stream);

/* Two disjoint row blocks in one call (a multi-GPU rank's share under the folded assignment: a block of long rows from the top of the
 * triangle and a block of short rows from the bottom, engine.folded_row_ranges).  Same results as two td_pairwise_device calls; the
 * postings pipeline shares one events kernel and one fix-up kernel between the blocks.  d_out_a / d_out_b: the blocks' output tables. */
int td_pairwise_device2(td_treeset *ts, uint32_t metric_mask, int normalise, int min_overlap, double overlap_fail_value,
                        const int64_t row_begin[2], const int64_t row_end[2], void *const d_out_a[4], void *const d_out_b[4], void *stream);

/* The kernels cannot return a status: a pair that compares a rooted with an unrooted tree (undefined in the reference) or a
 * geodesic pair that hits the iteration cap leaves NaN in the output and raises a sticky flag on the resident set.  td_pairwise
 * reads it itself; after the asynchronous entry points (td_pairwise_device, td_pairwise_rect_device) call td_check_errors once the
 * stream has been synchronised: it returns TD_ERR_MIXED_ROOTING / TD_ERR_CUDA and clears the flag (one flag per set and GPU, shared by
 * all launches in flight on it). */
int td_check_errors(td_treeset *ts, int device);

/* The library caches freed device blocks per GPU (cudaMalloc / cudaFree cost milliseconds): at most TREEDIST_CACHE_BYTES
 * (environment, default 4 GiB) per device.  td_trim_cache gives the cache of `device` (-1: all) back to the driver. */
int td_trim_cache(int device);

/* Same, end to end with HOST output buffers: uploads the set if needed, runs the kernels, and streams the
 * condensed result(s) into out[metric] (host pointers, pinned or pageable).  Synchronous. */
int td_pairwise(td_treeset *ts, uint32_t metric_mask, int normalise, int min_overlap, double overlap_fail_value,
                int64_t row_begin, int64_t row_end, double *const out[4], int device);

/* Same for ONE metric, but the result is the N x N SQUARE symmetric matrix with a zero diagonal, row-major in out[N * N]: what
 * scipy.spatial.distance.squareform makes of the condensed vector at treeCl/collection.py:456, produced on the device so that
 * get_inter_tree_distances can wrap the caller's buffer in a DistanceMatrix without another pass over it. */
int td_pairwise_square(td_treeset *ts, int metric, int normalise, int min_overlap, double overlap_fail_value, double *out, int device);

/* Output buffers in page-locked host memory are filled by direct device-to-host copies at PCIe rate, in row bands that overlap the
 * kernels; pageable buffers work too but go through a pinned staging block and a host copy.  td_alloc_pinned / td_free_pinned hand
 * out page-locked blocks from a cache (page-locking costs tens of milliseconds per 100 MB; at most TREEDIST_PINNED_CACHE_BYTES, default
 * 4 GiB, stay cached).  Band size: TREEDIST_BAND_BYTES per metric (default 32 MiB). */
int td_alloc_pinned(int64_t bytes, void **out);
int td_free_pinned(void *p);

/* Rectangular variant (treeCl/bootstrap.py:174-218): all |A| x |B| distances, row-major
 * out[a * nB + b].  Both sets must have been encoded TOGETHER (one td_encode_newick call) - `split` is the index
 * of the first tree of B inside `ts`. */
int td_pairwise_rect_device(td_treeset *ts, int64_t split, uint32_t metric_mask, int normalise, void *const d_out[4],
                            void *stream);

/* The same for two SEPARATELY encoded sets and host buffers: out[metric][q * n_reference + r] = distance(query tree q, reference tree r),
 * the block treeCl/bootstrap.py:174-218 fills with a doubly nested Python loop over _fast_geo (run_optimise_bootstrap_coords,
 * run_out_of_sample_mds, run_analytical_fit).  Both sets must share one taxon table (same leaf labels); the library joins their split
 * dictionaries, uploads the joint set, runs the rectangular kernels and copies the block back.  Synchronous. */
int td_pairwise_rect(td_treeset *query, td_treeset *reference, uint32_t metric_mask, int normalise, double *const out[4], int device);

/* Exact RF through the int8 tcgen05 split-incidence contraction C = X * X^T (uniform sets whose shared-split
 * dictionary is small enough); RF = S_i + S_j - 2 C_ij.  d_out: uint16 or double condensed, device pointer. */
int td_rf_incidence_device(td_treeset *ts, int normalise, int64_t row_begin, int64_t row_end, void *d_out,
                           int out_is_u16, void *stream);

/* Same with a HOST output buffer (uint16 condensed, rows [row_begin, row_end)): row bands, the kernels of band k under the
 * device-to-host copy of band k - 1 (BASELINE config 4: "tiles streamed to pinned host").  Pinned `out` (td_alloc_pinned): direct copies. */
int td_rf_incidence(td_treeset *ts, int64_t row_begin, int64_t row_end, uint16_t *out, int device);

/* ---- consumers of the matrix on the device ------------------------------------------------------- */

/* Classical multidimensional scaling, DistanceMatrix.embedding(dimensions, 'cmds') of the reference:
 *     double_centre                treeCl/distance_matrix.py:58-74     B = -1/2 J D^2 J
 *     eigen                        treeCl/distance_matrix.py:285-298   eigenvalues descending (np.linalg.eigh)
 *     _embedding_classical_mds     treeCl/distance_matrix.py:301-315   coords = evecs[:, :dims] * sqrt(|vals[:dims]|)
 * but only the `dimensions` leading eigenpairs are computed (blocked subspace iteration with Rayleigh-Ritz on the fp64 tensor cores),
 * and in td_cmds the N x N matrix never leaves the GPU: only the N x dimensions embedding does.  Eigenvectors are defined up to sign, as
 * with eigh.  1 <= dimensions <= 12.  The iteration stops when the residuals || B v - lambda v || of the wanted eigenpairs fall below `tol` x the
 * largest |eigenvalue| (<= 0: 1e-12), or after max_iter (<= 0: 5000) products; *iterations / *residual (that relative residual) report how it ended - a
 * matrix whose leading eigenvalues are (nearly) degenerate, e.g. independent random trees, converges slowly and has no well-defined
 * embedding in the reference either.
 *   td_cmds_device : condensed matrix already on the device (output of td_pairwise_device), coordinates stay on the device
 *   td_cmds        : distances of `metric` computed on the device, embedded there; coords[n_trees * dimensions] (row-major), host
 *   td_cmds_host   : square symmetric matrix in host memory (what DistanceMatrix.to_array() holds) */
int td_cmds_device(const double *d_condensed, int64_t n_trees, int dimensions, double *d_coords, double *eigenvalues, int max_iter, double tol,
                   int *iterations, double *residual, void *stream);
int td_cmds(td_treeset *ts, int metric, int normalise, int min_overlap, double overlap_fail_value, int dimensions, double *coords,
            double *eigenvalues, int max_iter, double tol, int *iterations, double *residual, int device);
int td_cmds_host(const double *square, int64_t n_trees, int dimensions, double *coords, double *eigenvalues, int max_iter, double tol,
                 int *iterations, double *residual, int device);

/* Spectral clustering's front end (treeCl/clustering.py:125-303, class Spectral) for the pruning options NONE / MANUAL and the scale options
 * MEDIAN / MANUAL (the ESTIMATE options search a connected k-nearest-neighbour graph on the host and are not rebuilt):
 *     kdists / kmask / kscale      treeCl/distance_matrix.py:153-186   k-th nearest distance per row, neighbour mask (OR / AND), local scale
 *     affinity                     treeCl/distance_matrix.py:32-55     exp(-d^2 / (s_i s_j)), masked entries 0
 *     laplace + eigen              treeCl/distance_matrix.py:189-207,285-298
 * scale_k = 0: s = row medians (LOCAL_SCALE_MEDIAN), else the scale_k-th nearest distance (LOCAL_SCALE_MANUAL); prune_k = 0: no pruning,
 * else kmask(prune_k) combined with its transpose by OR (prune_and = 0) or AND.
 *   td_affinity_host             Spectral(...).affinity: the N x N affinity matrix with a unit diagonal (Spectral.decompose)
 *   td_spectral_embedding_host   Spectral(...).spectral_embedding_(dimensions) (algo ZELNIKMANOR): affinity with a zero diagonal ->
 *                                D^-1/2 A D^-1/2 -> leading eigenvectors -> rows scaled to unit length; coords[n_trees * dimensions]
 *   td_spectral_embedding        the same with the distances of `metric` computed on the device: the matrix never leaves the GPU
 * When the leading eigenvalues are (nearly) equal - well separated clusters - the eigenvectors are defined up to a rotation inside their
 * subspace (in the reference too); the unit-length rows then agree up to that rotation. */
int td_affinity_host(const double *square, int64_t n_trees, int prune_k, int prune_and, int scale_k, double *affinity, int device);
int td_spectral_embedding_host(const double *square, int64_t n_trees, int dimensions, int prune_k, int prune_and, int scale_k, double *coords,
                               double *eigenvalues, int max_iter, double tol, int *iterations, double *residual, int device);
int td_spectral_embedding(td_treeset *ts, int metric, int normalise, int dimensions, int prune_k, int prune_and, int scale_k, double *coords,
                          double *eigenvalues, int max_iter, double tol, int *iterations, double *residual, int device);

/* Counters of the last geodesic launch on this set (mean GTP iterations etc.), for DESIGN.md / bench.py. */
typedef struct {
    uint64_t pairs, maxflow_solves, ratios, augmentations, launches;
} td_geo_stats;
int td_geo_get_stats(td_treeset *ts, td_geo_stats *out);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t td_launch_count(void);

const char *td_last_error(void);
const char *td_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TREEDIST_CUDA_H */
