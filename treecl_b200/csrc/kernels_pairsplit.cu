// kernels_pairsplit.cu -- rf / wrf / euc on one leaf set: postings pipeline + fp64 tensor-core tile kernel (sm_100a).
//
// Same contract as kernels_pairwise.cu (getRobinsonFoulds / getWeightedRobinsonFoulds / getEuclideanDistance of
// treeCl/treedist.py:111-114 for all pairs of a row range), split into a sparse and a dense part:
//
//   1. events (pair_events_kernel).  A pair (i, j) has a common split d  <=>  i and j both appear in the POSTINGS of d (the encoder
//      groups the shared-list entries by split, trees ascending).  Every later holder j of a split of tree i is one EVENT of pair
//      (i, j) carrying the corrections   cnt += 1,   E2 += -2 la lb,   E1 += |la - lb| - |la| - |lb|.
//      Work is proportional to the number of events, no searching at all.  One CTA per group of 16 row trees bins its events by
//      64 x 32 tile with a shared-memory counting sort and claims one contiguous range of the event buffer.
//   2. tiles (pair_dense_kernel), one CTA per tile: the tile's events are accumulated into per-cell slots in shared memory (64-bit
//      fixed point relative to the cell's magnitude: integer adds commute, so the result does not depend on the event order), the
//      leaf edges go through the fp64 tensor cores in the dot-product form
//          euc^2 = (Q2_i + Q2_j) - 2 sum_k a_k b_k + E2           (Q2 = sum of squares of ALL edge lengths of a tree)
//      (DMMA m8n8k4: the same 37 TFLOP/s as DFMA but 1/8 of the issue slots, and one FMA per term instead of DADD + DFMA),
//      wrf keeps the direct form sum_k |a_k - b_k| on the fp64 pipe, and the epilogue leaves through shared memory so that every
//      warp writes whole 256-byte row segments of the condensed matrix.
//   Hybrid form: the matrix behind the tile kernel's TMA descriptors may carry, below the taxon rows, one row per HEAVY shared split
//   (the splits held by the most trees, PairParams::heavy_rows of them, chosen by the encoder's cost model; build_split_rows_kernel).
//   Their products are then part of the dot product (their |la - lb| of the L1 sum), the events kernel skips them, and rf comes from
//   the incidence contraction.  With ALL shared splits as rows no events are needed at all (the split-extended form of api.cu).
// (Q + E) subtracts; cells whose correction nearly cancels Q are queued for exact_fixup_kernel, which recomputes them additively in a
// fixed order: results are bit-reproducible, independent of the row sharding, and identical trees give exactly 0.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "exact_pair.cuh"

namespace cg = cooperative_groups;

namespace td {

namespace {

constexpr int TR = 64, TC = 32;       // tile: rows x columns
constexpr int NT = 256;
constexpr int KCH = 16;               // taxa per staged chunk
#ifndef PS_NSTAGE
#define PS_NSTAGE 4
#endif
// CTAs per SM and rows of the epilogue handled together, per metric mask.  Four CTAs per SM (64 registers) with two rows at a time beat
// three CTAs (80 registers) with eight rows - 0.1403 vs 0.1479 ms per step of the bench set, A/B in one call - wherever the kernel fits
// 64 registers without spilling, i.e. for every mask except wrf + euc together (20 bytes of spills there: 5.42 vs 5.34 ms on config 3).
// -DPS_MINB / -DPS_OB force one setting for all masks (experiments).
#ifdef PS_MINB
constexpr int min_blocks(uint32_t) { return PS_MINB; }
#else
constexpr int min_blocks(uint32_t mask) { return (mask & 6u) == 6u ? 3 : 4; }
#endif
#ifdef PS_OB
constexpr int rows_at_once(uint32_t) { return PS_OB; }
#else
constexpr int rows_at_once(uint32_t mask) { return (mask & 6u) == 6u ? 8 : 2; }
#endif
constexpr int NSTAGE = PS_NSTAGE;     // chunks in flight
constexpr int SA = 68, SB = 36;       // row strides (doubles) of the staged leaf chunks: = 4 mod 16 -> conflict-free DMMA fragments
constexpr int ES = 40;                // row stride (doubles) of the per-cell slots: conflict-free 16-byte accesses of the C fragments
// (Q + E) subtracts.  A result below 0.2 % of Q = Q_i + Q_j (error amplification 500: relative error <= 500 * (n_taxa + splits) * 2^-53,
// i.e. <= 7e-11 even for the 1,275-entry Kendall-Colijn vectors) is recomputed exactly by the fix-up kernel.  For euc that is a pair
// closer than 4.5 % of the root of the summed squared norms: near-duplicates, not the bulk of a real data set.
constexpr double kRisk1 = 0.998;      // wrf: exact recompute when -E1 > kRisk1 * (Q1_i + Q1_j)
constexpr double kRisk2 = 0.998;      // euc: exact recompute when -E2 > kRisk2 * (Q2_i + Q2_j)

// ---- mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one TMA request moves a whole box (KCH taxon rows x 68 or 36 trees) of the taxon-major leaf matrix; out-of-range parts arrive as 0
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t spins = 0;; spins++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();          // watchdog: never hang the GPU on a lost transaction
    }
}

// bytes of the per-cell slots (E1, E2, counts) and of the region they share with the ring of leaf chunks
__host__ __device__ constexpr size_t kCellBytes(uint32_t mask)
{
    return ((mask & 2u) ? TR * ES * sizeof(double) : 0) + ((mask & 4u) ? TR * ES * sizeof(double) : 0) + TR * TC;
}
__host__ __device__ constexpr size_t kRegionBytes(uint32_t mask)
{
    const size_t ring = (mask & 6u) ? (size_t)NSTAGE * KCH * (SA + SB) * sizeof(double) : 0, cells = (kCellBytes(mask) + 127) & ~(size_t)127;
    return ring > cells ? ring : cells;
}

// Order-independent accumulation of the corrections of one cell: 64-bit fixed point relative to the cell's own magnitude
// Q = Q_i + Q_j.  Every correction and every partial sum is bounded by Q (2 la lb <= la^2 + lb^2;  | |la-lb| - |la| - |lb| | <= |la| + |lb|),
// so with the scale 2^(61 - exponent(Q)) nothing overflows, the resolution is Q * 2^-61 (finer than fp64 rounding of Q + E), and integer
// adds commute: any number of corrections per cell sums to the same bits in any order.
struct FixScale { double to_fix, from_fix; };
__device__ __forceinline__ FixScale fix_scale(double q)
{
    // biased exponent of to_fix = 2^(61 - (e - 1023)) is 2107 - e; capped at 2045 so that from_fix = 1 / to_fix (biased exponent
    // 2046 - that) is a normal number too.  The cap only applies to Q < 2^-961 (zero included), where corrections times 2^1022 still fit.
    const int e = (__double2hiint(q) >> 20) & 0x7FF;                  // biased exponent of q (q >= 0, finite)
    const int be = min(2107 - e, 2045);
    FixScale f;
    f.to_fix = __hiloint2double(be << 20, 0);
    f.from_fix = __hiloint2double((2046 - be) << 20, 0);
    return f;
}

// Programmatic dependent launch: pair_dense_kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization right behind
// pair_events_kernel, whose CTAs all raise the trigger first thing - so tile CTAs start on SMs the events kernel no longer fills and run
// their TMA + fp64 tensor-core phase under its tail; griddepcontrol.wait (a no-op in a normal launch) then blocks until the events
// kernel has completed and its segment table and records are visible.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// D(8x8) += A(8x4, row) * B(4x8, col), fp64.  Lane l = 4 g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------------------------------------
// events from the postings.  FILL = false: count per tile;  FILL = true: write the events (tile_cnt was cleared by the scan)
// ------------------------------------------------------------------------------------------------------------------------------
// Event records (array of structures, one aligned store per event):
//   WMODE 0 (rf only)      4 B  { u32 cell }
//   WMODE 1 (wrf) / 2 (euc) 16 B { f64 t1 or t2, u32 cell, u32 0 }
//   WMODE 3 (wrf + euc)    32 B { f64 t2, f64 t1, u32 cell, 12 B pad }
// cell = row * 32 + column inside the 64 x 32 tile.
__host__ __device__ constexpr int wmode_of(uint32_t mask) { return (int)((mask >> 1) & 3u); }
__host__ __device__ constexpr int rec_bytes(int wmode) { return wmode == 0 ? 4 : (wmode == 3 ? 32 : 16); }

constexpr int RG = 16;                // trees per row group: one CTA of pair_events_kernel; a 64-row tile reads 64 / RG segments

// One CLUSTER of P CTAs (P = 1, 2, 4 or 8, chosen at launch so that the grid is at least two waves on any GPU share) per ROW GROUP of RG
// consecutive trees: all events (i, j) with i in the group, binned by column tile j / 32.  CTA `part` of the cluster owns the shared-list
// positions pos = part (mod P) of the group's trees.
//   entries : the group's shared-list entries, each with its position x in the postings; entry (d, i) owns the events
//             (i, post_tree[b]) for b in (x, post_end[x])
//   count   : a warp walks EVW entries at a time, lane b takes the b-th later holder of each: no searching, coalesced reads of a
//             posting list, EVW independent reads in flight per lane; a shared-memory histogram over the column tiles counts them
//   place   : the CTAs read each other's histograms through distributed shared memory: bin totals, block scan, ONE global atomicAdd
//             (CTA 0) claims the group's contiguous range of the event buffer; CTA q starts bin c behind the events of CTAs < q
//   fill    : the same walk again, slots from shared-memory atomics, one aligned record store per event
// and the (offset, count) of every (row group, column tile) segment goes to the segment table for pair_dense_kernel.
// Nothing but the range claim touches a global atomic, and there is no separate count / scan / fill launch.
constexpr int EVT = 1024;             // threads of pair_events_kernel
#ifndef PS_EVW
#define PS_EVW 2
#endif
constexpr int EVW = PS_EVW;           // entries a warp walks at a time

template <int WMODE, bool RECT>
__global__ void __launch_bounds__(EVT) pair_events_kernel(const __grid_constant__ PairParams p)
{
    extern __shared__ __align__(16) unsigned char ev_smem[];
    const int n_ent = RG * p.sh_stride, ct = p.col_tiles_total;
    double *ent_len = reinterpret_cast<double *>(ev_smem);                                 // [n_ent]     (weighted only, else unused)
    uint32_t *ent_n = reinterpret_cast<uint32_t *>(ev_smem + (WMODE ? (size_t)n_ent * 8 : 0));     // [n_ent + 1] events of the entry
    int *ent_x = reinterpret_cast<int *>(ent_n + n_ent + 1);                               // [n_ent] position in the postings
    uint32_t *hist = reinterpret_cast<uint32_t *>(ent_x + n_ent);                          // [ct] this CTA's counts, then its cursors
    uint32_t *bin_total = hist + ct, *bin_before = bin_total + ct;                         // [ct] events of the whole cluster / of the CTAs before this one
    __shared__ uint32_t warp_sum[EVT / 32];
    __shared__ uint32_t s_base;

    cg::cluster_group cluster = cg::this_cluster();
    const unsigned n_parts = cluster.num_blocks(), part = cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g_idx = (int)(blockIdx.x / n_parts);
    const int rg = g_idx < p.groups_a ? g_idx + p.rg0 : g_idx - p.groups_a + p.rg0b, tree0 = rg * RG;       // (second row block: see PairParams)
    griddep_launch_dependents();                           // the tile kernel may start filling free SMs (it waits before it reads our output)

    // block-wide exclusive scan of one value per thread; returns the exclusive prefix, the total through `total`
    auto block_scan = [&](uint32_t v, uint32_t &total) -> uint32_t {
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        __syncthreads();                                   // warp_sum may still be read from a previous call
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        uint32_t ws = warp_sum[lane], wi = ws;             // EVT / 32 == 32 warp sums: one per lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += y; }
        total = __shfl_sync(0xffffffffu, wi, 31);
        const uint32_t before = __shfl_sync(0xffffffffu, wi - ws, warp);
        return before + incl - v;
    };

    // ---- this CTA's entries (q = m * RG + local tree, shared-list position pos = part + m * n_parts) and their sizes ------------------------
    for (int c = tid; c < ct; c += EVT) hist[c] = 0u;
    const int n_own = ((p.sh_stride - (int)part + (int)n_parts - 1) / (int)n_parts) * RG;
    for (int q = tid; q < n_own; q += EVT) {
        const int i = tree0 + (q & (RG - 1)), pos = (q / RG) * (int)n_parts + (int)part;
        int x = -1;
        if ((i >= p.row_begin && i < p.row_end) || (i >= p.row_begin2 && i < p.row_end2)) x = p.shared_x[(size_t)i * p.sh_stride + pos];
        if (x >= 0 && p.heavy_rows > 0 && p.shared[(size_t)i * p.sh_stride + pos].pad < (uint32_t)p.heavy_rows) x = -1;    // a row of the matrix, no events
        uint32_t n = 0;
        if (x >= 0) { n = (uint32_t)(p.post_end[x] - x - 1); if constexpr (WMODE) ent_len[q] = p.post_len[x]; }
        ent_x[q] = x; ent_n[q] = n;
    }
    __syncthreads();

    // ---- count: a warp walks EVW entries at a time, lane b takes the b-th later holder of each (no search; EVW independent reads) ----------
    for (int e0 = warp * EVW; e0 < n_own; e0 += (EVT / 32) * EVW) {
        int x[EVW]; uint32_t n[EVW], n_max = 0;
#pragma unroll
        for (int u = 0; u < EVW; u++) {
            const int e = e0 + u;
            x[u] = -1; n[u] = 0;
            if (e < n_own) { x[u] = ent_x[e]; n[u] = ent_n[e]; }
            n_max = max(n_max, n[u]);
        }
        for (uint32_t b = lane; b < n_max; b += 32) {
            int j[EVW];
#pragma unroll
            for (int u = 0; u < EVW; u++) j[u] = b < n[u] ? p.post_tree[x[u] + 1 + (int)b] : -1;
#pragma unroll
            for (int u = 0; u < EVW; u++)
                if (j[u] >= (RECT ? (int)p.col_offset : 0)) atomicAdd(&hist[j[u] >> 5], 1u);
        }
    }
    __syncthreads();
    // ---- place: bin totals over the cluster, scan, claim the range, publish the segment table --------------------------------------------------
    cluster.sync();                                        // every CTA's histogram is complete
    {
        const int cper = (ct + EVT - 1) / EVT, c_lo = min(tid * cper, ct), c_hi = min(c_lo + cper, ct);
        uint32_t sum = 0;
        for (int c = c_lo; c < c_hi; c++) {
            uint32_t tot = 0, before = 0;
            for (unsigned q = 0; q < n_parts; q++) {
                const uint32_t h = q == part ? hist[c] : *cluster.map_shared_rank(&hist[c], q);
                if (q < part) before += h;
                tot += h;
            }
            bin_total[c] = tot; bin_before[c] = before; sum += tot;
        }
        uint32_t total;
        uint32_t at = block_scan(sum, total);
        if (part == 0 && tid == 0) {
            const uint32_t base = total ? atomicAdd(p.ev_cursor, total) : 0u;
            for (unsigned q = 0; q < n_parts; q++) *cluster.map_shared_rank(&s_base, q) = base;
        }
        cluster.sync();                                    // s_base has arrived everywhere, nobody reads a remote histogram any more
        at += s_base;
        uint2 *seg = p.seg + (size_t)rg * ct;
        for (int c = c_lo; c < c_hi; c++) {
            const uint32_t n = bin_total[c];
            if (part == 0) seg[c] = make_uint2(at, n);
            hist[c] = at + bin_before[c]; at += n;
        }
    }
    __syncthreads();
    // ---- fill ----------------------------------------------------------------------------------------------------------------------------
    auto emit = [&](int e, int j, double lb) {              // one event of entry e with the later holder j (length lb)
        uint32_t o = 0xFFFFFFFFu;
        if (j >= (RECT ? (int)p.col_offset : 0)) o = atomicAdd(&hist[j >> 5], 1u);   // else: no event / the other set's own side
        if (o < p.ev_capacity) {
            const uint32_t cell = (uint32_t)((((tree0 + (e & (RG - 1))) & 63) << 5) | (j & 31));
            if constexpr (WMODE == 0) reinterpret_cast<uint32_t *>(p.ev)[o] = cell;
            else {
                const double la = ent_len[e];
                const double t2 = -2.0 * la * lb, t1 = fabs(la - lb) - fabs(la) - fabs(lb);
                int4 *dst = reinterpret_cast<int4 *>(p.ev) + (size_t)o * (rec_bytes(WMODE) / 16);
                if constexpr (WMODE == 3) {
                    dst[0] = make_int4(__double2loint(t2), __double2hiint(t2), __double2loint(t1), __double2hiint(t1));
                    dst[1] = make_int4((int)cell, 0, 0, 0);
                } else {
                    const double tv = WMODE == 2 ? t2 : t1;
                    dst[0] = make_int4(__double2loint(tv), __double2hiint(tv), (int)cell, 0);
                }
            }
        }
    };
    for (int e0 = warp * EVW; e0 < n_own; e0 += (EVT / 32) * EVW) {
        int x[EVW]; uint32_t n[EVW], n_max = 0;
#pragma unroll
        for (int u = 0; u < EVW; u++) {
            const int e = e0 + u;
            x[u] = -1; n[u] = 0;
            if (e < n_own) { x[u] = ent_x[e]; n[u] = ent_n[e]; }
            n_max = max(n_max, n[u]);
        }
        for (uint32_t b = lane; b < n_max; b += 32) {
            int j[EVW]; double lb[EVW];
#pragma unroll
            for (int u = 0; u < EVW; u++) {
                j[u] = -1; lb[u] = 0.0;
                if (b < n[u]) { j[u] = p.post_tree[x[u] + 1 + (int)b]; if constexpr (WMODE) lb[u] = p.post_len[x[u] + 1 + (int)b]; }
            }
#pragma unroll
            for (int u = 0; u < EVW; u++) emit(e0 + u, j[u], lb[u]);
        }
    }
    // this CTA's share of the row group is written (segment table rows and records): tell the tiles that wait for the group
    if (p.ready) {
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(p.ready + rg / (TR / RG), 1u);
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// tiles
// ------------------------------------------------------------------------------------------------------------------------------
// rare path, kept out of line: queue the pair for the exact fix-up kernel (same stream, right after this one); when the work
// list is full the exact sums are computed here instead
__device__ __noinline__ void queue_exact(const PairParams &p, int64_t i, int64_t j, double &t1, double &t2)
{
    const unsigned slot = atomicAdd(p.wl_count, 1u);
    if (slot < p.wl_capacity) p.wl_pairs[slot] = make_int2((int)i, (int)j);
    else exact_pair(p, i, j, t1, t2);
}

#ifdef PS_TRACE
#define TRACE(k) do { if (tid == 0 && p.trace) p.trace[t * 8 + (k)] = clock64(); } while (0)
#else
#define TRACE(k) do {} while (0)
#endif

// One CTA per 64 x 32 tile.  Every global read of the tile is requested in the first instructions (segment bounds, per-tree scalars
// into registers, the first NSTAGE leaf chunks as TMA bulk copies) and consumed as late as possible.
template <uint32_t MASK, bool RECT, bool NORM>
__global__ void __launch_bounds__(NT, min_blocks(MASK)) pair_dense_kernel(const __grid_constant__ PairParams p, const __grid_constant__ CUtensorMap map_a,
                                                                 const __grid_constant__ CUtensorMap map_b)
{
    constexpr bool kRF = MASK & 1u, kWRF = (MASK >> 1) & 1u, kEUC = (MASK >> 2) & 1u;
    constexpr bool kWeighted = kWRF || kEUC;
    constexpr int WMODE = wmode_of(MASK), REC = rec_bytes(WMODE);
    constexpr int OB = rows_at_once(MASK);                 // rows of the epilogue handled together

    // ---- which tile.  Rectangular: blockIdx = (column tile, tile row).  Triangular: tile row u holds L0 - 2u column tiles; rows u
    // and R-1-u are folded into one grid row of constant length 2 L0 - 2 (R-1), so no CTA has to invert the triangular numbering.
    int u = blockIdx.y, x = blockIdx.x;
    const int L0 = p.tiles_x - 2 * p.tile_row0;
    if (!RECT) {
        const int len = L0 - 2 * u;
        if (x >= len) {
            x -= len; u = p.tri_rows - 1 - u;
            if (u == (int)blockIdx.y || x >= L0 - 2 * u) return;      // middle row of an odd fold / past the end: nothing here
        }
    }
    const int tile_r = u + p.tile_row0, tile_c = RECT ? x + p.col_tile0 : 2 * tile_r + x;
    const long long t = RECT ? (long long)u * p.tiles_x + x : (long long)u * L0 - (long long)u * (u - 1) + x;
    const int i0 = tile_r * TR, j0 = tile_c * TC;
    (void)t;                                               // only the -DPS_TRACE build uses the linear tile index

    // Shared memory.  The ring of leaf chunks and the per-cell slots share one region: the slots are only needed once the leaf sums are
    // in registers (the ring holds NSTAGE * KCH = 64 taxa, so up to 64 taxa there is no refill and no barrier inside the k-loop).
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *cur = smem_raw;
    double *E1 = nullptr, *E2 = nullptr, *sA = nullptr, *sB = nullptr, *sQ1 = nullptr, *sQ2 = nullptr;
    {
        unsigned char *c2 = cur;                                                                           // cells
        if constexpr (kWRF) { E1 = reinterpret_cast<double *>(c2); c2 += TR * ES * sizeof(double); }
        if constexpr (kEUC) { E2 = reinterpret_cast<double *>(c2); c2 += TR * ES * sizeof(double); }
    }
    unsigned char *cnt = cur + kCellBytes(MASK) - TR * TC;                                                 // [TR*TC] common splits
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>(cnt);
    if constexpr (kWeighted) {                                                                             // ring: TMA destinations, 128-byte aligned
        sA = reinterpret_cast<double *>(cur);
        sB = reinterpret_cast<double *>(cur + NSTAGE * KCH * SA * sizeof(double));
    }
    cur += kRegionBytes(MASK);
    uint64_t *full = reinterpret_cast<uint64_t *>(cur); cur += NSTAGE * 8;                                 // one mbarrier per stage
    if constexpr (kWRF) { sQ1 = reinterpret_cast<double *>(cur); cur += (TR + TC) * sizeof(double); }     // rows, then columns
    if constexpr (kEUC) { sQ2 = reinterpret_cast<double *>(cur); cur += (TR + TC) * sizeof(double); }
    int *sS = reinterpret_cast<int *>(cur); cur += (TR + TC) * sizeof(int);                                // internal edges per tree

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;            // DMMA fragment coordinates
    const int wr = warp >> 1, wc = warp & 1;           // warp tile: rows 16 wr .., columns 16 wc ..
    TRACE(0);
    // with per-group ready counters nothing in a later tile launch of the same launch set depends on this one: let it in early
    if (kWeighted && p.ready) griddep_launch_dependents();

    // ---- leaf-length chunks: two TMA requests per chunk (rows of A, rows of B), issued by one thread ------------------------------------
    const int k_used = kWeighted ? (p.k_rows + 3) & ~3 : 0;                   // rows of the matrix behind the tensor maps (beyond: zeros)
    const int n_chunks = (k_used + KCH - 1) / KCH;
    auto issue_chunk = [&](int chunk, int stage) {                            // one thread
        mbar_expect_tx(&full[stage], (uint32_t)(KCH * (SA + SB) * sizeof(double)));
        tma_load_2d(sA + stage * KCH * SA, &map_a, i0, chunk * KCH, &full[stage]);
        tma_load_2d(sB + stage * KCH * SB, &map_b, j0, chunk * KCH, &full[stage]);
    };
    if constexpr (kWeighted) {
        if (tid == 0) {
            for (int st = 0; st < NSTAGE; st++) mbar_init(&full[st], 1);
            mbar_init_fence();
            for (int c = 0; c < NSTAGE && c < n_chunks; c++) issue_chunk(c, c);
        }
    }
    // this tile's events: one segment per row group.  The table is the events kernel's output: nothing of it is touched before
    // griddep_wait() (see above); requested after the first leaf chunk, consumed after the second
    uint2 sg[TR / RG];
#pragma unroll
    for (int q = 0; q < TR / RG; q++) sg[q] = make_uint2(0u, 0u);
    auto load_segments = [&]() {
        if (!p.seg) return;                                // no table: no events
        if (kWeighted && p.ready) {
            // only the four row groups of this tile row have to be complete, not the whole events grid: every thread acquires the tile
            // row's counter itself (its own later reads of the table and the records are then ordered behind the events kernel's writes)
            const unsigned *f = p.ready + tile_r;
            while (ld_acquire_gpu(f) < p.ready_target) __nanosleep(100);
        } else
        griddep_wait();
        const uint2 *segp = p.seg + (size_t)(tile_r * (TR / RG)) * p.col_tiles_total + tile_c;
#pragma unroll
        for (int q = 0; q < TR / RG; q++) sg[q] = segp[(size_t)q * p.col_tiles_total];
    };
    // ---- per-tree scalars: requested now (threads 0..95), parked in registers, put into shared memory just before the epilogue -------
    int sc_s = 0; double sc_q1 = 0.0, sc_q2 = 0.0;
    if (tid < TR + TC) {
        const int tree = tid < TR ? i0 + tid : j0 + (tid - TR);
        sc_s = p.nsplits[tree];
        if constexpr (kWRF) sc_q1 = p.q1[tree];
        if constexpr (kEUC) sc_q2 = p.q2t[tree];
    }
    // the other threads must not poll the stage barriers before thread 0 has initialised them (the shared memory still holds whatever the
    // previous CTA - of any kernel - left there); the loads requested above stay in flight across the barrier
    if constexpr (kWeighted) __syncthreads();
    TRACE(1);
    // ---- this tile's events: one aligned load per record.  The first EVR records of every thread are requested now and applied after
    // the leaf sums (their latency hides behind the tensor-core work).
    constexpr int EVR = REC == 32 ? 1 : 2;
    static_assert(TR / RG == 4, "four row groups per tile");
    uint32_t n0 = 0, n1 = 0, n2 = 0, n_ev = 0;             // running sizes of the four segments (set when the table entries are first used)
    const unsigned char *rec = reinterpret_cast<const unsigned char *>(p.ev);
    int4 pre[EVR][2];
    auto load_rec = [&](uint32_t k, int4 &v, int4 &w) {
        const size_t o = k < n0 ? sg[0].x + k : (k < n1 ? sg[1].x + (k - n0) : (k < n2 ? sg[2].x + (k - n1) : sg[3].x + (k - n2)));
        if constexpr (REC == 4) v.x = (int)reinterpret_cast<const uint32_t *>(rec)[o];
        else if constexpr (REC == 16) v = reinterpret_cast<const int4 *>(rec)[o];
        else { v = reinterpret_cast<const int4 *>(rec)[2 * o]; w = reinterpret_cast<const int4 *>(rec)[2 * o + 1]; }
    };
    auto apply_rec = [&](const int4 &v, const int4 &w) {
        const uint32_t cell = REC == 4 ? (uint32_t)v.x : (REC == 16 ? (uint32_t)v.z : (uint32_t)w.x);
        atomicAdd(&cnt32[cell >> 2], 1u << (8 * (cell & 3)));
        const int row = cell >> 5, col = cell & 31, slot = row * ES + col;
        if constexpr (REC == 16) {
            const double tv = __hiloint2double(v.y, v.x);
            if constexpr (kEUC) atomicAdd(reinterpret_cast<unsigned long long *>(E2 + slot), (unsigned long long)__double2ll_rn(tv * fix_scale(sQ2[row] + sQ2[TR + col]).to_fix));
            else atomicAdd(reinterpret_cast<unsigned long long *>(E1 + slot), (unsigned long long)__double2ll_rn(tv * fix_scale(sQ1[row] + sQ1[TR + col]).to_fix));
        }
        if constexpr (REC == 32) {
            atomicAdd(reinterpret_cast<unsigned long long *>(E2 + slot), (unsigned long long)__double2ll_rn(__hiloint2double(v.y, v.x) * fix_scale(sQ2[row] + sQ2[TR + col]).to_fix));
            atomicAdd(reinterpret_cast<unsigned long long *>(E1 + slot), (unsigned long long)__double2ll_rn(__hiloint2double(v.w, v.z) * fix_scale(sQ1[row] + sQ1[TR + col]).to_fix));
        }
    };
    // requested after the first leaf chunk: by then the segment table entries have arrived, and the records arrive under the other chunks
    auto prefetch_events = [&]() {
        n0 = sg[0].y; n1 = n0 + sg[1].y; n2 = n1 + sg[2].y; n_ev = n2 + sg[3].y;
#pragma unroll
        for (int e = 0; e < EVR; e++) { pre[e][0] = make_int4(0, 0, 0, 0); pre[e][1] = pre[e][0]; if (tid + e * NT < n_ev) load_rec(tid + e * NT, pre[e][0], pre[e][1]); }
    };
    if (!kWeighted || n_chunks == 0) { load_segments(); prefetch_events(); }
    auto apply_events = [&]() {
#pragma unroll
        for (int e = 0; e < EVR; e++) if (tid + e * NT < n_ev) apply_rec(pre[e][0], pre[e][1]);
        for (uint32_t o = EVR * NT + tid; o < n_ev; o += NT) { int4 v = make_int4(0, 0, 0, 0), w = v; load_rec(o, v, w); apply_rec(v, w); }
    };
    TRACE(2);

    // ---- leaf edges: dot products on the fp64 tensor cores (euc), |a - b| sums on the fp64 pipe (wrf) ------------------------------
    double dot[2][2][2] = {}, l1[2][2][2] = {};        // [row block][column block][element] of the warp's 16 x 16 tile
    if constexpr (kWeighted) {
        for (int chunk = 0; chunk < n_chunks; chunk++) {
            const int stage = chunk % NSTAGE;
            mbar_wait(&full[stage], (chunk / NSTAGE) & 1);
            const double *cA = sA + stage * KCH * SA + wr * 16 + g, *cB = sB + stage * KCH * SB + wc * 16;
            const int k_here = min(KCH, k_used - chunk * KCH);
            auto step4 = [&](int k0) {
                if constexpr (kEUC) {
                    const double a0 = cA[(k0 + t4) * SA], a1 = cA[(k0 + t4) * SA + 8];
                    const double b0 = cB[(k0 + t4) * SB + g], b1 = cB[(k0 + t4) * SB + g + 8];
                    dmma884(dot[0][0][0], dot[0][0][1], a0, b0);
                    dmma884(dot[0][1][0], dot[0][1][1], a0, b1);
                    dmma884(dot[1][0][0], dot[1][0][1], a1, b0);
                    dmma884(dot[1][1][0], dot[1][1][1], a1, b1);
                }
                if constexpr (kWRF) {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const double a[2] = {cA[(k0 + kk) * SA], cA[(k0 + kk) * SA + 8]};
                        const double2 b0 = *reinterpret_cast<const double2 *>(cB + (k0 + kk) * SB + 2 * t4);
                        const double2 b1 = *reinterpret_cast<const double2 *>(cB + (k0 + kk) * SB + 8 + 2 * t4);
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            l1[r][0][0] += fabs(a[r] - b0.x); l1[r][0][1] += fabs(a[r] - b0.y);
                            l1[r][1][0] += fabs(a[r] - b1.x); l1[r][1][1] += fabs(a[r] - b1.y);
                        }
                    }
                }
            };
            if (k_here == KCH) {
#pragma unroll
                for (int k0 = 0; k0 < KCH; k0 += 4) step4(k0);
            } else {
                for (int k0 = 0; k0 < k_here; k0 += 4) step4(k0);
            }
            if (chunk == 0) load_segments();
            if (chunk == (n_chunks > 1 ? 1 : 0)) prefetch_events();
            if (chunk + NSTAGE < n_chunks) {               // refill this stage once every warp is done with it
                __syncthreads();
                if (tid == 0) issue_chunk(chunk + NSTAGE, stage);
            }
        }
    }
    TRACE(3);
    __syncthreads();                                       // the ring is drained: its memory now holds the per-cell slots
    // every thread clears the cells it owns (DMMA fragment layout)
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int row = wr * 16 + r * 8 + g, col = wc * 16 + c * 8 + 2 * t4;
            if constexpr (kEUC) *reinterpret_cast<double2 *>(E2 + row * ES + col) = make_double2(0.0, 0.0);     // (all-zero bits: also 0 as int64)
            if constexpr (kWRF) *reinterpret_cast<double2 *>(E1 + row * ES + col) = make_double2(0.0, 0.0);
            *reinterpret_cast<uint16_t *>(cnt + row * TC + col) = 0;
        }
    if (tid < TR + TC) {
        sS[tid] = sc_s;
        if constexpr (kWRF) sQ1[tid] = sc_q1;
        if constexpr (kEUC) sQ2[tid] = sc_q2;
    }
    __syncthreads();
    apply_events();                                        // fixed-point integer adds: the order does not matter
    __syncthreads();                                       // all event atomics have landed
    if constexpr (kWeighted) {
        // the cells this thread owns: fixed point back to fp64, leaf sums folded in
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int row = wr * 16 + r * 8 + g, col = wc * 16 + c * 8 + 2 * t4, slot = row * ES + col;
                if constexpr (kEUC) {
                    double2 *q = reinterpret_cast<double2 *>(E2 + slot);
                    const longlong2 a = *reinterpret_cast<const longlong2 *>(q);
                    const double qi = sQ2[row];
                    double2 v;
                    v.x = fma(-2.0, dot[r][c][0], (double)a.x * fix_scale(qi + sQ2[TR + col]).from_fix);
                    v.y = fma(-2.0, dot[r][c][1], (double)a.y * fix_scale(qi + sQ2[TR + col + 1]).from_fix);
                    *q = v;
                }
                if constexpr (kWRF) {
                    double2 *q = reinterpret_cast<double2 *>(E1 + slot);
                    const longlong2 a = *reinterpret_cast<const longlong2 *>(q);
                    const double qi = sQ1[row];
                    double2 v;
                    v.x = (double)a.x * fix_scale(qi + sQ1[TR + col]).from_fix + l1[r][c][0];
                    v.y = (double)a.y * fix_scale(qi + sQ1[TR + col + 1]).from_fix + l1[r][c][1];
                    *q = v;
                }
            }
        __syncthreads();
    }
    TRACE(4);

    // ---- out: warp w writes rows 8w .. 8w+7, lane = column: whole 256-byte row segments.  Branch-free over the 8 rows so that the
    // square roots (long dependent chains) overlap; stores are predicated.
    const int j = j0 + lane;
    const bool col_ok = j < p.col_end && (!RECT || j >= p.col_offset);
    const int s_j = sS[TR + lane];
    double q1j = 0.0, q2j = 0.0, sumj = 0.0, normj = 0.0;
    if constexpr (kWRF) q1j = sQ1[TR + lane];
    if constexpr (kEUC) q2j = sQ2[TR + lane];
    if constexpr (NORM) {
        if constexpr (kWRF) sumj = p.sum_len[j];
        if constexpr (kEUC) normj = p.norm[j];
    }
    const int iw = i0 + warp * 8;
    // condensed index of (iw, j); row iw + r is at idx0 + r (N - iw - 2) - r (r - 1) / 2   (rectangular: idx0 + r n_cols)
    const int64_t idx0 = (RECT ? (int64_t)iw * p.n_cols - p.col_offset : ((int64_t)iw * p.N - (int64_t)iw * (iw + 1) / 2 - iw - 1 - p.p_base)) + j;
#pragma unroll 1
    for (int r0 = 0; r0 < 8; r0 += OB) {                   // OB rows at a time (their square roots overlap)
        uint32_t ok = 0, risky = 0;
        double v1[OB], v2[OB];
        int common[OB];
#pragma unroll
        for (int rr = 0; rr < OB; rr++) {
            const int row = warp * 8 + r0 + rr, i = iw + r0 + rr;
            ok |= (uint32_t)(col_ok && i >= p.row_begin && i < p.row_end && (RECT || j > i)) << rr;
            common[rr] = cnt[row * TC + lane];
            bool rk = false;
            v1[rr] = 0.0; v2[rr] = 0.0;
            if constexpr (kWRF) {
                const double q = sQ1[row] + q1j, e = E1[row * ES + lane];
                rk |= -e > kRisk1 * q;                     // e already holds the (non-negative) leaf sum: only makes the test milder
                v1[rr] = q + e;
            }
            if constexpr (kEUC) {
                const double q = sQ2[row] + q2j, e = E2[row * ES + lane];
                rk |= -e > kRisk2 * q;
                const double sum = q + e;
                v2[rr] = sum < 0.0 ? 0.0 : sum;             // (cheaper than fmax: no NaN handling needed here)
            }
            risky |= (uint32_t)rk << rr;
        }
        if constexpr (kWeighted) {
            risky &= ok;
            while (risky) {                                // rare: normally overwritten by the fix-up kernel
                const int rr = __ffs(risky) - 1; risky &= risky - 1;
                double t1 = 0.0, t2 = 0.0;
#pragma unroll
                for (int q = 0; q < OB; q++) if (q == rr) { t1 = v1[q]; t2 = v2[q]; }
                queue_exact(p, iw + r0 + rr, j, t1, t2);
#pragma unroll
                for (int q = 0; q < OB; q++) if (q == rr) { v1[q] = t1; v2[q] = t2; }
            }
        }
#pragma unroll
        for (int rr = 0; rr < OB; rr++) {
            const int r = r0 + rr, row = warp * 8 + r, i = iw + r;
            const int64_t idx = idx0 + (RECT ? (int64_t)r * p.n_cols : (int64_t)r * (p.N - iw - 2) - r * (r - 1) / 2);
            const bool st = ok >> rr & 1u;
            if constexpr (kRF) {
                const int den = sS[row] + s_j;
                double v = (double)(den - 2 * common[rr]);
                if constexpr (NORM) v = den ? v / (double)den : 0.0;
                if (st) p.out[0][idx] = v;
            }
            if constexpr (kWRF) {
                double v = v1[rr];
                if constexpr (NORM) { const double den = (st ? p.sum_len[i] : 1.0) + sumj; v = den != 0.0 ? v / den : 0.0; }
                if (st) p.out[1][idx] = v;
            }
            if constexpr (kEUC) {
                double v = sqrt(v2[rr]);
                if constexpr (NORM) { const double den = (st ? p.norm[i] : 1.0) + normj; v = den != 0.0 ? v / den : 0.0; }
                if (st) p.out[2][idx] = v;
            }
        }
    }
    TRACE(5);
#ifdef PS_TRACE
    if (tid == 0 && p.trace) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.trace[t * 8 + 6] = sm; p.trace[t * 8 + 7] = n_ev; }
#endif
}

// pdl: the launch directly follows pair_events_kernel on `st` and may start before it has finished (programmatic dependent launch)
template <uint32_t MASK, bool RECT>
cudaError_t launch_dense(const PairParams &p, dim3 grid, size_t smem, cudaStream_t st, bool pdl)
{
    auto k = p.normalise ? pair_dense_kernel<MASK, RECT, true> : pair_dense_kernel<MASK, RECT, false>;
    cudaError_t e = allow_full_smem((const void *)k);
    if (e != cudaSuccess) return e;
    CUtensorMap ma, mb;
    if (p.tmap_a && p.tmap_b) { memcpy(&ma, p.tmap_a, sizeof ma); memcpy(&mb, p.tmap_b, sizeof mb); }
    else { if (MASK & 6u) return cudaErrorInvalidValue; memset(&ma, 0, sizeof ma); memset(&mb, 0, sizeof mb); }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, p, ma, mb);
}

template <bool RECT>
cudaError_t launch_dense_mask(uint32_t mask, const PairParams &p, dim3 grid, size_t smem, cudaStream_t st, bool pdl)
{
    switch (mask & 7u) {
        case 1: return launch_dense<1, RECT>(p, grid, smem, st, pdl);
        case 2: return launch_dense<2, RECT>(p, grid, smem, st, pdl);
        case 3: return launch_dense<3, RECT>(p, grid, smem, st, pdl);
        case 4: return launch_dense<4, RECT>(p, grid, smem, st, pdl);
        case 5: return launch_dense<5, RECT>(p, grid, smem, st, pdl);
        case 6: return launch_dense<6, RECT>(p, grid, smem, st, pdl);
        case 7: return launch_dense<7, RECT>(p, grid, smem, st, pdl);
    }
    return cudaErrorInvalidValue;
}

template <bool RECT>
cudaError_t launch_events(int wmode, const PairParams &p, unsigned groups, unsigned parts, size_t smem, cudaStream_t st)
{
    auto run = [&](auto k) {
        cudaError_t e = allow_full_smem((const void *)k);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(groups * parts); cfg.blockDim = dim3(EVT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = parts; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k, p);
    };
    switch (wmode) {
        case 0: return run(pair_events_kernel<0, RECT>);
        case 1: return run(pair_events_kernel<1, RECT>);
        case 2: return run(pair_events_kernel<2, RECT>);
        default: return run(pair_events_kernel<3, RECT>);
    }
}

}  // namespace

// 2-D tensor maps over the taxon-major leaf matrix leafT [n_k][ldT] (fp64): boxes of KCH taxon rows x 68 (row trees) or 36 (column
// trees) values.  The 4 surplus values per row only pad the shared-memory row stride to 4 mod 16 doubles (conflict-free fragments).
cudaError_t make_leaf_tensor_maps(const double *leafT, int64_t ldT, int n_k, void *tmap_a, void *tmap_b)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    const cuuint64_t dims[2] = {(cuuint64_t)ldT, (cuuint64_t)n_k};
    const cuuint64_t strides[1] = {(cuuint64_t)ldT * sizeof(double)};      // bytes between taxon rows
    const cuuint32_t estr[2] = {1, 1};
    for (int which = 0; which < 2; which++) {
        const cuuint32_t box[2] = {(cuuint32_t)(which ? SB : SA), (cuuint32_t)KCH};
        CUresult r = encode(reinterpret_cast<CUtensorMap *>(which ? tmap_b : tmap_a), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(leafT), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

// Split-extended matrix for densely sharing sets: rows [0, n_k) are leafT, row n_k + sid holds the lengths of shared split sid (0 where a
// tree does not have it).  With it the common-split products of the Euclidean metric are part of the tile kernel's dot product and no
// events are needed:  euc^2 = Q2t_i + Q2t_j - 2 <row_i, row_j>.
__global__ void build_split_rows_kernel(const SplitRec *__restrict__ shared, int sh_stride, int64_t N, int64_t ld, double *__restrict__ rows, int max_rows)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * sh_stride) return;
    const SplitRec r = shared[e];
    if (r.id != kSentinelId && r.pad < (uint32_t)max_rows) rows[(size_t)r.pad * ld + e / sh_stride] = r.len;
}

cudaError_t launch_build_split_rows(const SplitRec *shared, int sh_stride, int64_t N, int64_t ld, double *rows, int max_rows, cudaStream_t st)
{
    const int64_t n = N * sh_stride;
    if (n <= 0) return cudaSuccess;
    build_split_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(shared, sh_stride, N, ld, rows, max_rows);
    return cudaGetLastError();
}

size_t pairsplit_dense_smem_bytes(uint32_t mask)
{
    size_t s = 128 + kRegionBytes(mask) + NSTAGE * 8 + (TR + TC) * sizeof(int);
    if (mask & 2u) s += (TR + TC) * sizeof(double);
    if (mask & 4u) s += (TR + TC) * sizeof(double);
    return s;
}

int pairsplit_record_bytes(uint32_t mask) { return rec_bytes(wmode_of(mask));
}

// number of 64 x 32 tiles of this launch (the length of p.tile_cnt); fills the tile geometry of p
long long pairsplit_tiles(bool rect, PairParams &p)
{
    const int64_t col_tiles_total = (p.N + TC - 1) / TC;
    p.tile_row0 = (int)(p.row_begin / TR);
    const int64_t tile_rows = (p.row_end + TR - 1) / TR - p.tile_row0;
    long long grid;
    if (rect) {
        p.col_tile0 = (int)(p.col_offset / TC);
        p.tiles_x = (int)(col_tiles_total - p.col_tile0);
        p.tri_rows = (int)tile_rows;
        grid = tile_rows * p.tiles_x;
    } else {
        p.col_tile0 = 0;
        p.tiles_x = (int)col_tiles_total;
        const int64_t L0 = col_tiles_total - 2 * p.tile_row0;
        // rows whose first column tile lies beyond the last one contribute nothing (only when N is not a multiple of 64)
        int64_t rows = tile_rows;
        while (rows > 0 && L0 - 2 * (rows - 1) <= 0) rows--;
        grid = rows * L0 - rows * (rows - 1);
        p.tri_rows = (int)rows;
    }
    p.n_tiles = grid > 0 ? grid : 0;
    return p.n_tiles;
}

size_t pairsplit_events_smem_bytes(uint32_t mask, int sh_stride, int col_tiles_total)
{
    const size_t n_ent = (size_t)RG * sh_stride;
    return (wmode_of(mask) ? n_ent * 8 : 0) + (n_ent + 1) * 4 + n_ent * 4 + (size_t)col_tiles_total * 12 + 16;
}
int pairsplit_row_group() { return RG; }

// the tile kernel alone (p carries the tile geometry; p.seg == nullptr: no events)
static cudaError_t launch_tiles(uint32_t mask, bool rect, const PairParams &p, cudaStream_t st, bool pdl)
{
    if (p.n_tiles <= 0) return cudaSuccess;
    const size_t smem_d = pairsplit_dense_smem_bytes(mask);
    // rectangular: (column tiles, tile rows); triangular: tile rows u and R-1-u folded into one grid row
    const dim3 grid = rect ? dim3((unsigned)p.tiles_x, (unsigned)p.tri_rows)
                           : dim3((unsigned)(2 * (p.tiles_x - 2 * p.tile_row0) - 2 * (p.tri_rows - 1)), (unsigned)((p.tri_rows + 1) / 2));
    return rect ? launch_dense_mask<true>(mask, p, grid, smem_d, st, pdl) : launch_dense_mask<false>(mask, p, grid, smem_d, st, pdl);
}

cudaError_t launch_pairsplit_tiles(uint32_t mask, bool rect, const PairParams &p, cudaStream_t st) { return launch_tiles(mask, rect, p, st, false); }

// p must carry the tile geometry (pairsplit_tiles), the postings, the segment table and the event buffer; *ev_cursor must be zero
static bool no_pdl_env() { static const bool v = getenv("TREEDIST_NO_PDL") != nullptr; return v; }

cudaError_t launch_pairsplit(uint32_t mask, bool rect, const PairParams &p_in, int n_sms, cudaStream_t st, int *launches, const PairParams *second)
{
    *launches = 0;
    if (p_in.n_tiles <= 0 && (!second || second->n_tiles <= 0)) return cudaSuccess;
    const bool weighted = (mask & 6u) != 0;
    // TREEDIST_SYNC_DEBUG: synchronise after every launch and say which kernel failed (debug aid; kills the overlap)
    static const bool sync_debug = getenv("TREEDIST_SYNC_DEBUG") != nullptr;
    auto checked = [&](cudaError_t e, const char *what) {
        if (e == cudaSuccess && sync_debug) e = cudaStreamSynchronize(st);
        if (sync_debug) fprintf(stderr, "[treedist] %s (mask %u): %s\n", what, mask, cudaGetErrorString(e));
        return e;
    };
    // p: block A plus, for the events and fix-up kernels, the row range / outputs of block B
    PairParams p = p_in;
    long long groups = (long long)p.tri_rows * (TR / RG);
    p.groups_a = (int)groups; p.row_begin2 = p.row_end2 = 0; p.rg0b = 0; p.p_base2 = 0;
    if (second) {
        p.row_begin2 = second->row_begin; p.row_end2 = second->row_end; p.p_base2 = second->p_base; p.rg0b = second->rg0;
        for (int m = 0; m < 3; m++) p.out2[m] = second->out[m];
        groups += (long long)second->tri_rows * (TR / RG);
    }
    cudaError_t e = cudaSuccess;
    {
        // the row groups of all tile rows of this launch (p.rg0 = first group of the first tile row): the kernel writes their rows of
        // the segment table completely, rows outside the row range(s) as empty segments
        const size_t smem_e = pairsplit_events_smem_bytes(mask, p.sh_stride, p.col_tiles_total);
        // CTAs per row group (a thread-block cluster).  One when the row groups alone fill the GPU (measured on the bench set: 316 groups,
        // 1 CTA each 0.156 ms per step, clusters of 2 = 632 CTAs = 1.07 waves 0.160 ms); a launch over a share of the rows (multi-GPU
        // row blocks: 76 - 106 groups) splits its groups until one wave of 2 CTAs per SM is full
        unsigned parts = 1;
        if (const char *f = getenv("TREEDIST_EVENT_PARTS")) parts = (unsigned)atoi(f);
        else {
            while (parts < 8 && groups * parts * 2 <= 2ll * n_sms) parts *= 2;
            // One block that is a share of the rows (the middle rank of a multi-GPU fold; the other ranks have a long-row and a short-row
            // block and hide the events kernel behind the short block's tiles, see the ready counters below): the events kernel has to be
            // short itself.  It is bound by per-SM throughput - the SMs that host two of e.g. 156 CTAs set its time - so the row groups
            // are split into 1, 2 or 4 CTAs, whichever spreads most evenly over the SMs (83 -> 68 us; that rank's share 0.729 -> 0.709 ms).
            if (!second && !rect && parts < 4 && (p.row_end - p.row_begin) * 2 < p.N) {
                double best = 1e30;
                for (unsigned c = 1; c <= 4; c *= 2) {
                    const double per_sm = (double)(groups * c) / (double)n_sms, waste = std::ceil(per_sm) / per_sm * (1.0 + 0.02 * (c - 1));
                    if (waste < best) { best = waste; parts = c; }
                }
            }
        }
        if (parts != 1 && parts != 2 && parts != 4 && parts != 8) parts = 1;
        // Per-tile-row ready counters instead of the grid-wide dependency, for launch sets with two row blocks (multi-GPU folded blocks):
        // the tiles of the short-row block, launched first, run while the long rows' events are still being written.  Config 3 as 8
        // ranks, per rank on one GPU: 0.727 -> 0.69-0.72 ms (4 ranks 1.413 -> 1.390, 2 ranks 2.760 -> 2.725); a single block gains
        // nothing (0.1396 vs 0.1393 ms on the bench set) and keeps the plain dependency.
        {
            static const char *force = getenv("TREEDIST_READY_FLAGS");           // "0" / "1": experiments
            bool use = second && weighted && !rect && !no_pdl_env() && !sync_debug;
            if (force && *force) use = *force != '0' && weighted && !no_pdl_env() && !sync_debug;
            if (!use) p.ready = nullptr;
        }
        p.ready_target = parts * (TR / RG);
        e = checked(rect ? launch_events<true>(wmode_of(mask), p, (unsigned)groups, parts, smem_e, st) : launch_events<false>(wmode_of(mask), p, (unsigned)groups, parts, smem_e, st), "pair_events_kernel");
        ++*launches;
        if (e != cudaSuccess) return e;
    }
    const bool pdl = !no_pdl_env() && !sync_debug;
    PairParams sb{};
    if (second) { sb = *second; sb.ready = p.ready; sb.ready_target = p.ready_target; }
    // With ready counters the SHORT-row block goes first: its row groups finish early in the events kernel and its tiles run while the
    // long rows' events are still being written; block A's tiles follow under its tail.
    const bool b_first = second && weighted && p.ready;
    if (b_first) {
        e = checked(launch_tiles(mask, rect, sb, st, pdl), "pair_dense_kernel (second block, first)");
        ++*launches;
        if (e != cudaSuccess) return e;
    }
    e = checked(launch_tiles(mask, rect, p, st, pdl), "pair_dense_kernel");
    ++*launches;
    if (e != cudaSuccess) return e;
    if (second && !b_first) {
        // block B's tiles: a launch of its own geometry reading the same segment table and records; it may start under block A's tail
        e = checked(launch_tiles(mask, rect, sb, st, pdl), "pair_dense_kernel (second block)");
        ++*launches;
        if (e != cudaSuccess) return e;
    }
    if (!weighted) return e;
    ++*launches;
    return checked(launch_exact_fixup(mask, rect, p, n_sms, st), "exact_fixup_kernel");
}

}  // namespace td
