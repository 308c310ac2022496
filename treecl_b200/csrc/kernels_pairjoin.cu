// kernels_pairjoin.cu -- RF / weighted RF / Euclidean for SPARSELY sharing tree sets: tile-level hash join (sm_100a).
//
// Same contract as kernels_pairwise.cu (getRobinsonFoulds / getWeightedRobinsonFoulds / getEuclideanDistance of
// treeCl/treedist.py:111-114 for a whole tile of pairs), different way of finding the common splits.
//
// The merge kernel walks both trees' shared-split lists for every pair: pairs x (s_i + s_j) dependent shared-memory
// loads, although in a set of independent random trees a pair has ~0.2 splits in common.  Here the work is
// proportional to the tile's ENTRIES plus its MATCHES.  A tile is 64 row trees x 32 column trees, 256 threads:
//   build : the column trees' entries (id -> column, length) go into an open-addressing hash table in shared memory
//   probe : ONE thread per row tree walks that tree's entries in list order and looks each id up; every hit is a
//           (row, column) match that it adds to per-cell slots in shared memory which only it writes:
//           cnt += 1,  E2 += -2*la*lb,  E1 += |la-lb| - |la| - |lb|.  No atomics, fixed summation order: results are
//           bit-reproducible and do not depend on how rows are sharded over GPUs.
//   leaf  : dense fp64 "distance GEMM" over the taxon-major leaf lengths (4x2 register tile per thread, k-chunks
//           double buffered with cp.async), accumulators initialised with E
//   out   : rf  = S_i + S_j - 2 cnt
//           euc = sqrt((Q2_i + Q2_j) + E2 + leaf2),  wrf = (Q1_i + Q1_j) + E1 + leaf1     (Q = per-tree totals)
// (Q + E) subtracts, so a pair whose |E| exceeds kRisk * (Q_i + Q_j) - nearly identical trees - is recomputed exactly
// by a merge over the two lists (all terms non-negative): identical trees still give exactly 0.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"
#include "exact_pair.cuh"

namespace td {

namespace {

constexpr int KC = 8;
constexpr int TR = 64, TC = 32;       // tile: rows x columns
constexpr int NT = 256;
constexpr double kRisk = 0.9;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t hash_id(uint32_t id, uint32_t mask) { return (id * 2654435761u >> 7) & mask; }

template <uint32_t MASK, bool RECT>
__global__ void __launch_bounds__(NT, 3) pairjoin_kernel(const PairParams p)
{
    constexpr bool kRF = MASK & 1u, kWRF = (MASK >> 1) & 1u, kEUC = (MASK >> 2) & 1u;
    constexpr bool kWeighted = kWRF || kEUC;

    // ---- which tile ---------------------------------------------------------------------------------------------
    int tile_r, tile_c;           // in units of 64 rows / 32 columns
    if (RECT) {
        tile_r = (int)(blockIdx.x / p.tiles_x) + p.tile_row0;
        tile_c = (int)(blockIdx.x % p.tiles_x) + p.col_tile0;
    } else {
        // tile row u (from p.tile_row0) needs column tiles 2*tile_r .. tiles_x-1: L0 - 2u of them, L0 = tiles_x - 2*tile_row0
        const long long t = blockIdx.x, L0 = p.tiles_x - 2 * p.tile_row0;
        const double b = (double)L0 + 1.0;
        long long u = (long long)floor((b - sqrt(b * b - 4.0 * (double)t)) * 0.5);
        if (u < 0) u = 0;
        while (u * L0 - u * (u - 1) > t) u--;
        while ((u + 1) * L0 - (u + 1) * u <= t) u++;
        tile_r = (int)u + p.tile_row0;
        tile_c = 2 * tile_r + (int)(t - (u * L0 - u * (u - 1)));
    }
    const int64_t i0 = (int64_t)tile_r * TR, j0 = (int64_t)tile_c * TC;

    // ---- shared memory ------------------------------------------------------------------------------------------
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t slots = p.table_slots, smask = slots - 1;
    unsigned char *cur = smem_raw;
    double *tab_len = nullptr, *E1 = nullptr, *E2 = nullptr, *leaf = nullptr;
    if constexpr (kWeighted) { tab_len = reinterpret_cast<double *>(cur); cur += (size_t)slots * sizeof(double); }
    if constexpr (kWRF) { E1 = reinterpret_cast<double *>(cur); cur += TR * TC * sizeof(double); }
    if constexpr (kEUC) { E2 = reinterpret_cast<double *>(cur); cur += TR * TC * sizeof(double); }
    if constexpr (kWeighted) { leaf = reinterpret_cast<double *>(cur); cur += 2 * KC * (TR + TC) * sizeof(double); }
    uint32_t *tab_id = reinterpret_cast<uint32_t *>(cur); cur += (size_t)slots * sizeof(uint32_t);
    unsigned char *tab_col = cur; cur += slots;
    unsigned char *cnt = cur;                                       // [TR*TC] common-split counts (<= 255)

    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;

    // ---- leaf-length k-chunks start flowing while the join runs --------------------------------------------------------------
    const int n_chunks = kWeighted ? p.n_k / KC : 0;
    auto load_chunk = [&](int chunk, int stage) {
        // per stage: A [KC][TR] then B [KC][TC]
        double *sA = leaf + (size_t)stage * KC * (TR + TC), *sB = sA + KC * TR;
        constexpr int PA = KC * TR / 2, PB = KC * TC / 2;
        for (int e = tid; e < PA + PB; e += NT) {
            if (e < PA) {
                const int kk = e / (TR / 2), c2 = e % (TR / 2);
                cp_async16(sA + kk * TR + 2 * c2, p.leafT + (size_t)(chunk * KC + kk) * p.ldT + i0 + 2 * c2);
            } else {
                const int f = e - PA, kk = f / (TC / 2), c2 = f % (TC / 2);
                cp_async16(sB + kk * TC + 2 * c2, p.leafT + (size_t)(chunk * KC + kk) * p.ldT + j0 + 2 * c2);
            }
        }
    };
    if constexpr (kWeighted) {
        if (n_chunks > 0) load_chunk(0, 0);
        cp_async_commit();
        if (n_chunks > 1) load_chunk(1, 1);
        cp_async_commit();
    }

    // ---- init table and per-cell slots ---------------------------------------------------------------------------------------------
    {
        uint4 *t4 = reinterpret_cast<uint4 *>(tab_id);
        for (uint32_t e = tid; e < slots / 4; e += NT) t4[e] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
        if constexpr (kWRF) { double2 *z = reinterpret_cast<double2 *>(E1); for (int e = tid; e < TR * TC / 2; e += NT) z[e] = make_double2(0.0, 0.0); }
        if constexpr (kEUC) { double2 *z = reinterpret_cast<double2 *>(E2); for (int e = tid; e < TR * TC / 2; e += NT) z[e] = make_double2(0.0, 0.0); }
        uint32_t *z = reinterpret_cast<uint32_t *>(cnt);
        for (int e = tid; e < TR * TC / 4; e += NT) z[e] = 0u;
    }
    __syncthreads();

    // ---- build: column trees -> hash table (multi-map: equal ids from different trees sit in successive slots) -------------------
    {
        const SplitRec *gB = p.shared + j0 * p.sh_stride;
        const int n_rec = TC * p.sh_stride;
        for (int e = tid; e < n_rec; e += NT) {
            const SplitRec b = gB[e];
            if (b.id == kSentinelId) continue;
            uint32_t h = hash_id(b.id, smask);
            while (atomicCAS(&tab_id[h], kEmpty, b.id) != kEmpty) h = (h + 1) & smask;
            tab_col[h] = (unsigned char)(e / p.sh_stride);
            if constexpr (kWeighted) tab_len[h] = b.len;
        }
    }
    __syncthreads();

    // ---- probe: thread `row` walks its tree's entries in list order; it alone writes the slots of its row ---------------------------
    if (tid < TR) {
        const int row = tid;
        const int64_t i = i0 + row;
        const int n = p.nshared[i];
        const SplitRec *A = p.shared + i * p.sh_stride;
        const int first_col = RECT ? 0 : (int)(i - j0) + 1;        // columns below this are on or under the diagonal
        for (int base = 0; base < n; base += 4) {
            SplitRec a[4];
#pragma unroll
            for (int k = 0; k < 4; k++) a[k] = (base + k < n) ? A[base + k] : SplitRec{kSentinelId, 0, 0.0};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (a[k].id == kSentinelId) continue;
                uint32_t h = hash_id(a[k].id, smask);
                for (;;) {
                    const uint32_t sid = tab_id[h];
                    if (sid == kEmpty) break;
                    if (sid == a[k].id) {
                        const int col = tab_col[h];
                        if (col >= first_col) {
                            const int cell = row * TC + col;
                            cnt[cell] = (unsigned char)(cnt[cell] + 1);
                            if constexpr (kWeighted) {
                                const double lb = tab_len[h];
                                if constexpr (kEUC) E2[cell] += -2.0 * a[k].len * lb;
                                if constexpr (kWRF) E1[cell] += fabs(a[k].len - lb) - fabs(a[k].len) - fabs(lb);
                            }
                        }
                    }
                    h = (h + 1) & smask;
                }
            }
        }
    }
    __syncthreads();

    // ---- my pairs: rows 4ty..4ty+3, columns tx and tx+16.  acc starts as the (negative) correction E --------------------------------
    double acc1[4][2], acc2[4][2];
    int common[4][2];
    uint32_t risky = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int64_t i = i0 + 4 * ty + r;
        double q1i = 0, q2i = 0;
        if constexpr (kWRF) q1i = p.q1[i];
        if constexpr (kEUC) q2i = p.q2[i];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int cell = (4 * ty + r) * TC + tx + 16 * c;
            const int64_t j = j0 + tx + 16 * c;
            common[r][c] = cnt[cell];
            acc1[r][c] = 0.0; acc2[r][c] = 0.0;
            bool rk = false;
            if constexpr (kWRF) { acc1[r][c] = E1[cell]; rk |= (-acc1[r][c] > kRisk * (q1i + p.q1[j])); }
            if constexpr (kEUC) { acc2[r][c] = E2[cell]; rk |= (-acc2[r][c] > kRisk * (q2i + p.q2[j])); }
            risky |= (rk ? 1u : 0u) << (2 * r + c);
        }
    }

    // ---- leaf edges -----------------------------------------------------------------------------------------------------------------
    if constexpr (kWeighted) {
        for (int chunk = 0; chunk < n_chunks; chunk++) {
            const int stage = chunk & 1;
            cp_async_wait<1>();
            __syncthreads();
            const double *sA = leaf + (size_t)stage * KC * (TR + TC), *sB = sA + KC * TR;
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const double2 a01 = *reinterpret_cast<const double2 *>(sA + kk * TR + 4 * ty);
                const double2 a23 = *reinterpret_cast<const double2 *>(sA + kk * TR + 4 * ty + 2);
                const double av[4] = {a01.x, a01.y, a23.x, a23.y};
                const double bv[2] = {sB[kk * TC + tx], sB[kk * TC + tx + 16]};
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const double d = av[r] - bv[c];
                        if constexpr (kWRF) acc1[r][c] += fabs(d);
                        if constexpr (kEUC) acc2[r][c] = fma(d, d, acc2[r][c]);
                    }
            }
            __syncthreads();
            if (chunk + 2 < n_chunks) load_chunk(chunk + 2, stage);
            cp_async_commit();
        }
    }

    // ---- epilogue -------------------------------------------------------------------------------------------------------------------
    int sj[2]; double q1j[2], q2j[2], sumj[2], normj[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int64_t j = j0 + tx + 16 * c;
        sj[c] = p.nsplits[j];
        if constexpr (kWRF) { q1j[c] = p.q1[j]; sumj[c] = p.sum_len[j]; }
        if constexpr (kEUC) { q2j[c] = p.q2[j]; normj[c] = p.norm[j]; }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int64_t i = i0 + 4 * ty + r;
        if (i < p.row_begin || i >= p.row_end) continue;
        const int si = p.nsplits[i];
        double q1i = 0, q2i = 0, sumi = 0, normi = 0;
        if constexpr (kWRF) { q1i = p.q1[i]; sumi = p.sum_len[i]; }
        if constexpr (kEUC) { q2i = p.q2[i]; normi = p.norm[i]; }
        const int64_t row_off = RECT ? i * p.n_cols - p.col_offset : (i * p.N - i * (i + 1) / 2 - i - 1 - p.p_base);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int64_t j = j0 + tx + 16 * c;
            if (j >= p.col_end || (RECT ? j < p.col_offset : j <= i)) continue;
            const int64_t idx = row_off + j;
            if constexpr (kRF) {
                const int den = si + sj[c];
                double v = (double)(den - 2 * common[r][c]);
                if (p.normalise) v = den ? v / (double)den : 0.0;
                p.out[0][idx] = v;
            }
            if constexpr (kWeighted) {
                // acc = E + leaf; the per-tree totals complete the sum: (Q_i + Q_j) + E + leaf
                double t1 = 0.0, t2 = 0.0;
                if constexpr (kWRF) t1 = (q1i + q1j[c]) + acc1[r][c];
                if constexpr (kEUC) t2 = (q2i + q2j[c]) + acc2[r][c];
                if (risky >> (2 * r + c) & 1u) exact_pair(p, i, j, t1, t2);
                if constexpr (kWRF) {
                    double v = t1;
                    if (p.normalise) { const double den = sumi + sumj[c]; v = den != 0.0 ? v / den : 0.0; }
                    p.out[1][idx] = v;
                }
                if constexpr (kEUC) {
                    double v = sqrt(t2);
                    if (p.normalise) { const double den = normi + normj[c]; v = den != 0.0 ? v / den : 0.0; }
                    p.out[2][idx] = v;
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Persistent, warp-specialised version.  384 threads: warps 0-7 are CONSUMERS (leaf GEMM + epilogue of tile t), warps
// 8-11 are PRODUCERS (hash join of tile t+1: table init, build, probe into the per-cell slots).  Here the probe runs
// over ALL entries of the 64 row trees in parallel and accumulates with shared-memory atomics.  That keeps results
// reproducible because (a) the counts are integers, (b) a cell that received one or two corrections holds their exact
// commutative sum, and (c) cells with three or more corrections - or with a correction that nearly cancels Q_i + Q_j -
// are queued on a work list and recomputed by exact_fixup_kernel with the additive merge (fixed order).
// Each CTA walks a contiguous range of the tile list; the two groups hand the per-cell slots over with named barriers:
//   FULL  (producers arrive, consumers wait): slots of the next tile are complete
//   EMPTY (consumers arrive, producers wait): consumers have copied their cells to registers and cleared them
// so the join of the next tile runs under the fp64 work of the current one and no warp waits at a CTA-wide barrier.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NCONS = 256, NPROD = 128, NTP = NCONS + NPROD;
constexpr int BAR_FULL = 1, BAR_EMPTY = 2, BAR_PROD = 3, BAR_CONS = 4;
constexpr int kEventsPerThread = 8;                   // private queue slots of a producer thread
constexpr uint32_t kEventCap = NPROD * kEventsPerThread;

// ---- mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t spins = 0;; spins++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
        __nanosleep(32);                           // do not burn issue slots that the consumer warps need
        if (spins > (1u << 24)) __trap();          // watchdog: never hang the GPU on a lost transaction
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// rare path, kept out of line: queue the pair for exact_fixup_kernel (same stream, right after the main kernel)
__device__ __noinline__ void queue_exact(const PairParams &p, int64_t i, int64_t j, double &t1, double &t2)
{
    const unsigned slot = atomicAdd(p.wl_count, 1u);
    if (slot < p.wl_capacity) p.wl_pairs[slot] = make_int2((int)i, (int)j);
    else exact_pair(p, i, j, t1, t2);
}

template <bool RECT>
__device__ __forceinline__ void decode_tile(const PairParams &p, long long t, int &tile_r, int &tile_c)
{
    if (RECT) {
        tile_r = (int)(t / p.tiles_x) + p.tile_row0;
        tile_c = (int)(t % p.tiles_x) + p.col_tile0;
    } else {
        const long long L0 = p.tiles_x - 2 * p.tile_row0;
        const double b = (double)L0 + 1.0;
        long long u = (long long)floor((b - sqrt(b * b - 4.0 * (double)t)) * 0.5);
        if (u < 0) u = 0;
        while (u * L0 - u * (u - 1) > t) u--;
        while ((u + 1) * L0 - (u + 1) * u <= t) u++;
        tile_r = (int)u + p.tile_row0;
        tile_c = 2 * tile_r + (int)(t - (u * L0 - u * (u - 1)));
    }
}

template <uint32_t MASK, bool RECT, bool NORM>
__global__ void __launch_bounds__(NTP, 2) pairjoin_persistent(const PairParams p)
{
    constexpr bool kRF = MASK & 1u, kWRF = (MASK >> 1) & 1u, kEUC = (MASK >> 2) & 1u;
    constexpr bool kWeighted = kWRF || kEUC;
    constexpr int KCP = (kWRF && kEUC) ? 4 : KC;      // two fp64 slot arrays: smaller k-chunks keep two CTAs per SM

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t slots = p.row_table_slots, smask = slots - 1;
    unsigned char *cur = smem_raw;
    double *E1 = nullptr, *E2 = nullptr, *leaf = nullptr;
    if constexpr (kWRF) { E1 = reinterpret_cast<double *>(cur); cur += TR * TC * sizeof(double); }
    if constexpr (kEUC) { E2 = reinterpret_cast<double *>(cur); cur += TR * TC * sizeof(double); }
    if constexpr (kWeighted) { leaf = reinterpret_cast<double *>(cur); cur += 2 * KCP * (TR + TC) * sizeof(double); }
    uint32_t *tab_id = reinterpret_cast<uint32_t *>(cur); cur += (size_t)slots * sizeof(uint32_t);      // split id or kEmpty
    uint32_t *tab_lo = reinterpret_cast<uint32_t *>(cur); cur += (size_t)slots * sizeof(uint32_t);      // row trees 0-31 holding it
    uint32_t *tab_hi = reinterpret_cast<uint32_t *>(cur); cur += (size_t)slots * sizeof(uint32_t);      // row trees 32-63
    unsigned char *cnt = cur; cur += TR * TC;
    SplitRec *stageA = reinterpret_cast<SplitRec *>(cur); cur += (size_t)TR * p.sh_stride * sizeof(SplitRec);   // row trees' lists
    SplitRec *stageB = reinterpret_cast<SplitRec *>(cur); cur += (size_t)TC * p.sh_stride * sizeof(SplitRec);   // column trees' lists
    uint32_t *events = reinterpret_cast<uint32_t *>(cur); cur += kEventCap * sizeof(uint32_t);          // row | col<<8 | pos<<16
    uint64_t *mbar = reinterpret_cast<uint64_t *>(cur); cur += 8;

    const int tid = threadIdx.x;
    const long long t_begin = p.n_tiles * (long long)blockIdx.x / gridDim.x;
    const long long t_end = p.n_tiles * ((long long)blockIdx.x + 1) / gridDim.x;

    // per-cell slots start cleared; afterwards every consumer clears the cells it has just read
    {
        if constexpr (kWRF) { double2 *z = reinterpret_cast<double2 *>(E1); for (int e = tid; e < TR * TC / 2; e += NTP) z[e] = make_double2(0.0, 0.0); }
        if constexpr (kEUC) { double2 *z = reinterpret_cast<double2 *>(E2); for (int e = tid; e < TR * TC / 2; e += NTP) z[e] = make_double2(0.0, 0.0); }
        uint32_t *z = reinterpret_cast<uint32_t *>(cnt);
        for (int e = tid; e < TR * TC / 4; e += NTP) z[e] = 0u;
    }
    if (tid == NCONS) mbar_init(mbar, 1);
    __syncthreads();
    if (t_begin >= t_end) return;

    if (tid >= NCONS) {
        // ================================================= PRODUCERS ====================================================
        // The hash table is built from the 64 ROW trees (id -> 64-bit mask of the rows that hold the split; one slot per
        // split: atomicCAS claims / finds it, atomicOr sets the row bit) and lives as long as the CTA stays in the same tile
        // row - with contiguous tile ranges that is ~20 tiles.  Per tile only the 32 COLUMN trees' lists are staged by one
        // TMA bulk copy (mbarrier completion) and probed: every entry finds its slot in ~1 step; matches go to the probing
        // thread's PRIVATE queue slots (no atomics) and are resolved by that same thread right after the probe: partner
        // length by binary search in the row tree's staged list, shared-memory atomics on the cell.  Every lane advances
        // through its own entries independently inside one warp-uniform loop, so a long chain in one lane does not stall
        // the others.
        const int ptid = tid - NCONS;
        uint32_t *cnt32 = reinterpret_cast<uint32_t *>(cnt);
        uint32_t parity = 0;
        int staged_row = -1;
        const int stride = p.sh_stride;
        const uint32_t bytesA = (uint32_t)(TR * stride * sizeof(SplitRec)), bytesB = (uint32_t)(TC * stride * sizeof(SplitRec));
        auto resolve = [&](uint32_t ev) {
            const int row = ev & 0xFF, col = (ev >> 8) & 0xFF, pos = ev >> 16;
            const int cell = row * TC + col;
            atomicAdd(&cnt32[cell >> 2], 1u << (8 * (cell & 3)));
            if constexpr (kWeighted) {
                const SplitRec b = stageB[(size_t)col * stride + pos];
                const SplitRec *A = stageA + (size_t)row * stride;
                int lo = 0, hi = stride - 1;                         // sentinels (0xFFFFFFFF) keep the padded list sorted
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (A[mid].id < b.id) lo = mid + 1; else hi = mid; }
                const double la = A[lo].len;
                if constexpr (kEUC) atomicAdd(&E2[cell], -2.0 * la * b.len);
                if constexpr (kWRF) atomicAdd(&E1[cell], fabs(la - b.len) - fabs(la) - fabs(b.len));
            }
        };
        for (long long t = t_begin; t < t_end; t++) {
            int tile_r, tile_c;
            decode_tile<RECT>(p, t, tile_r, tile_c);
            const int64_t i0 = (int64_t)tile_r * TR, j0 = (int64_t)tile_c * TC;
            const bool new_row = tile_r != staged_row;
            staged_row = tile_r;
            bar_sync(BAR_PROD, NPROD);           // every producer is done with the previous staging buffers (and table)
            if (ptid == 0) {
                fence_proxy_async();             // generic-proxy reads of the staging buffers precede the async-proxy writes
                mbar_expect_tx(mbar, bytesB + (new_row ? bytesA : 0u));
                bulk_g2s(stageB, p.shared + j0 * stride, bytesB, mbar);
                if (new_row) bulk_g2s(stageA, p.shared + i0 * stride, bytesA, mbar);
            }
            if (new_row) {
                uint4 *t4 = reinterpret_cast<uint4 *>(tab_id);           // ids, then the two mask arrays, are contiguous
                for (uint32_t e = ptid; e < slots / 4; e += NPROD) {
                    t4[e] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
                    t4[slots / 4 + e] = make_uint4(0, 0, 0, 0); t4[slots / 2 + e] = make_uint4(0, 0, 0, 0);
                }
                bar_sync(BAR_PROD, NPROD);
            }
            mbar_wait(mbar, parity); parity ^= 1u;
            if (new_row) {
                // ---- build: thread -> row (ptid & 63), entries (ptid >> 6) + 2k of that row tree ------------------------------
                const int row = ptid & (TR - 1);
                const uint32_t *A = reinterpret_cast<const uint32_t *>(stageA + (size_t)row * stride);     // id of entry k at A[4k]
                int pos = ptid >> 6;
                uint32_t id = pos < stride ? A[4 * pos] : kSentinelId;
                uint32_t h = hash_id(id, smask);
                while (__any_sync(0xffffffffu, id != kSentinelId)) {
                    if (id != kSentinelId) {
                        const uint32_t old = atomicCAS(&tab_id[h], kEmpty, id);
                        if (old == kEmpty || old == id) {
                            atomicOr(row < 32 ? &tab_lo[h] : &tab_hi[h], 1u << (row & 31));
                            pos += NPROD / TR;
                            id = pos < stride ? A[4 * pos] : kSentinelId;
                            h = hash_id(id, smask);
                        } else h = (h + 1) & smask;
                    }
                }
                bar_sync(BAR_PROD, NPROD);
            }
            // ---- probe: thread -> column (ptid & 31), entries (ptid >> 5) + 4k of that column tree --------------------------
            int my_events = 0;
            uint32_t *my_queue = events + ptid * kEventsPerThread;
            const int col = ptid & (TC - 1);
            const long long lim = RECT ? 64 : (long long)(j0 + col - i0);        // rows above the diagonal: i0 + row < j0 + col
            const unsigned long long rowvalid = lim >= 64 ? ~0ull : (lim <= 0 ? 0ull : ((1ull << lim) - 1ull));
            const uint32_t *B = reinterpret_cast<const uint32_t *>(stageB + (size_t)col * stride);
            int ovf_pos = -1; unsigned long long ovf_m = 0ull;           // where this thread's private queue filled up (rare)
            {
                int pos = ptid >> 5;
                uint32_t id = pos < stride ? B[4 * pos] : kSentinelId;
                uint32_t h = hash_id(id, smask);
                while (__any_sync(0xffffffffu, id != kSentinelId)) {
                    if (id != kSentinelId) {
                        const uint32_t sid = tab_id[h];
                        if (sid != kEmpty && sid != id) h = (h + 1) & smask;
                        else {
                            unsigned long long m = 0ull;
                            if (sid == id) m = (((unsigned long long)tab_hi[h] << 32) | tab_lo[h]) & rowvalid;
                            while (m && my_events < kEventsPerThread) {
                                const int row = __ffsll((long long)m) - 1; m &= m - 1;
                                my_queue[my_events++] = (uint32_t)row | ((uint32_t)col << 8) | ((uint32_t)pos << 16);
                            }
                            if (m) { ovf_pos = pos; ovf_m = m; id = kSentinelId; }      // continue after the cells are free
                            else {
                                pos += NPROD / TC;
                                id = pos < stride ? B[4 * pos] : kSentinelId;
                                h = hash_id(id, smask);
                            }
                        }
                    }
                }
            }
            // the probe above only reads the table and the staged lists, so it ran while the consumers were still using
            // the cells of the previous tile; the cell updates below need them released
            bar_sync(BAR_EMPTY, NTP);
            for (int e = 0; e < my_events; e++) resolve(my_queue[e]);
            if (ovf_pos >= 0) {
                // private queue overflowed (dense sharing): finish this thread's entries, resolving matches directly
                int pos = ovf_pos; unsigned long long m = ovf_m;
                for (;;) {
                    while (m) {
                        const int row = __ffsll((long long)m) - 1; m &= m - 1;
                        resolve((uint32_t)row | ((uint32_t)col << 8) | ((uint32_t)pos << 16));
                    }
                    pos += NPROD / TC;
                    if (pos >= stride) break;
                    const uint32_t id = B[4 * pos];
                    if (id == kSentinelId) break;
                    uint32_t h = hash_id(id, smask);
                    for (;;) { const uint32_t sid = tab_id[h]; if (sid == kEmpty || sid == id) { if (sid == id) m = (((unsigned long long)tab_hi[h] << 32) | tab_lo[h]) & rowvalid; break; } h = (h + 1) & smask; }
                }
            }
            __threadfence_block();
            bar_arrive(BAR_FULL, NTP);
        }
        return;
    }

    // ===================================================== CONSUMERS ========================================================
    const int tx = tid % 16, ty = tid / 16;
    bar_arrive(BAR_EMPTY, NTP);                 // initial credit: the slots are free
    for (long long t = t_begin; t < t_end; t++) {
        int tile_r, tile_c;
        decode_tile<RECT>(p, t, tile_r, tile_c);
        const int64_t i0 = (int64_t)tile_r * TR, j0 = (int64_t)tile_c * TC;

        const int n_chunks = kWeighted ? p.n_k / KCP : 0;
        auto load_chunk = [&](int chunk, int stage) {
            double *sA = leaf + (size_t)stage * KCP * (TR + TC), *sB = sA + KCP * TR;
            constexpr int PA = KCP * TR / 2, PB = KCP * TC / 2;
            for (int e = tid; e < PA + PB; e += NCONS) {
                if (e < PA) {
                    const int kk = e / (TR / 2), c2 = e % (TR / 2);
                    cp_async16(sA + kk * TR + 2 * c2, p.leafT + (size_t)(chunk * KCP + kk) * p.ldT + i0 + 2 * c2);
                } else {
                    const int f = e - PA, kk = f / (TC / 2), c2 = f % (TC / 2);
                    cp_async16(sB + kk * TC + 2 * c2, p.leafT + (size_t)(chunk * KCP + kk) * p.ldT + j0 + 2 * c2);
                }
            }
        };
        if constexpr (kWeighted) {
            if (n_chunks > 0) load_chunk(0, 0);
            cp_async_commit();
            if (n_chunks > 1) load_chunk(1, 1);
            cp_async_commit();
        }

        // per-tree scalars of my rows / columns (global, L1/L2 hits) while the producers finish the join
        int si[4], sj[2]; double q1i[4], q2i[4], q1j[2], q2j[2];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int64_t i = i0 + 4 * ty + r;
            si[r] = p.nsplits[i];
            if constexpr (kWRF) q1i[r] = p.q1[i];
            if constexpr (kEUC) q2i[r] = p.q2[i];
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int64_t j = j0 + tx + 16 * c;
            sj[c] = p.nsplits[j];
            if constexpr (kWRF) q1j[c] = p.q1[j];
            if constexpr (kEUC) q2j[c] = p.q2[j];
        }

        bar_sync(BAR_FULL, NTP);                 // slots of this tile are complete
        double acc1[4][2], acc2[4][2];
        int common[4][2];
        uint32_t risky = 0;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int cell = (4 * ty + r) * TC + tx + 16 * c;
                common[r][c] = cnt[cell]; cnt[cell] = 0;
                acc1[r][c] = 0.0; acc2[r][c] = 0.0;
                bool rk = false;
                if constexpr (kWRF) { acc1[r][c] = E1[cell]; E1[cell] = 0.0; rk |= (-acc1[r][c] > kRisk * (q1i[r] + q1j[c])); }
                if constexpr (kEUC) { acc2[r][c] = E2[cell]; E2[cell] = 0.0; rk |= (-acc2[r][c] > kRisk * (q2i[r] + q2j[c])); }
                if constexpr (kWeighted) rk |= common[r][c] >= 3;     // sum order of >= 3 atomic adds is not fixed
                risky |= (rk ? 1u : 0u) << (2 * r + c);
            }
        if (t + 1 < t_end) { __threadfence_block(); bar_arrive(BAR_EMPTY, NTP); }    // producers may fill the next tile

        if constexpr (kWeighted) {
            for (int chunk = 0; chunk < n_chunks; chunk++) {
                const int stage = chunk & 1;
                cp_async_wait<1>();
                bar_sync(BAR_CONS, NCONS);
                const double *sA = leaf + (size_t)stage * KCP * (TR + TC), *sB = sA + KCP * TR;
                auto step = [&](int kk) {
                    const double2 a01 = *reinterpret_cast<const double2 *>(sA + kk * TR + 4 * ty);
                    const double2 a23 = *reinterpret_cast<const double2 *>(sA + kk * TR + 4 * ty + 2);
                    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
                    const double bv[2] = {sB[kk * TC + tx], sB[kk * TC + tx + 16]};
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const double d = av[r] - bv[c];
                            if constexpr (kWRF) acc1[r][c] += fabs(d);
                            if constexpr (kEUC) acc2[r][c] = fma(d, d, acc2[r][c]);
                        }
                };
                const int k_left = p.n_taxa - chunk * KCP;           // the padded rows of the last chunk are zero: skip them
                if (k_left >= KCP) {
#pragma unroll
                    for (int kk = 0; kk < KCP; kk++) step(kk);
                } else {
                    for (int kk = 0; kk < k_left; kk++) step(kk);
                }
                bar_sync(BAR_CONS, NCONS);
                if (chunk + 2 < n_chunks) load_chunk(chunk + 2, stage);
                cp_async_commit();
            }
        }

        // epilogue
        double sumj[2] = {0.0, 0.0}, normj[2] = {0.0, 0.0};
        if constexpr (NORM) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int64_t j = j0 + tx + 16 * c;
                if constexpr (kWRF) sumj[c] = p.sum_len[j];
                if constexpr (kEUC) normj[c] = p.norm[j];
            }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int64_t i = i0 + 4 * ty + r;
            if (i < p.row_begin || i >= p.row_end) continue;
            double sumi = 0, normi = 0;
            if constexpr (NORM) {
                if constexpr (kWRF) sumi = p.sum_len[i];
                if constexpr (kEUC) normi = p.norm[i];
            }
            const int64_t row_off = RECT ? i * p.n_cols - p.col_offset : (i * p.N - i * (i + 1) / 2 - i - 1 - p.p_base);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int64_t j = j0 + tx + 16 * c;
                if (j >= p.col_end || (RECT ? j < p.col_offset : j <= i)) continue;
                const int64_t idx = row_off + j;
                if constexpr (kRF) {
                    const int den = si[r] + sj[c];
                    double v = (double)(den - 2 * common[r][c]);
                    if constexpr (NORM) v = den ? v / (double)den : 0.0;
                    p.out[0][idx] = v;
                }
                if constexpr (kWeighted) {
                    double t1 = 0.0, t2 = 0.0;
                    if constexpr (kWRF) t1 = (q1i[r] + q1j[c]) + acc1[r][c];
                    if constexpr (kEUC) t2 = (q2i[r] + q2j[c]) + acc2[r][c];
                    if (risky >> (2 * r + c) & 1u) queue_exact(p, i, j, t1, t2);
                    if constexpr (kWRF) {
                        double v = t1;
                        if constexpr (NORM) { const double den = sumi + sumj[c]; v = den != 0.0 ? v / den : 0.0; }
                        p.out[1][idx] = v;
                    }
                    if constexpr (kEUC) {
                        double v = sqrt(t2);
                        if constexpr (NORM) { const double den = normi + normj[c]; v = den != 0.0 ? v / den : 0.0; }
                        p.out[2][idx] = v;
                    }
                }
            }
        }
    }
}


// Recompute the queued pairs additively (no Q + E subtraction) and overwrite their wrf / euc entries.  One thread per pair running
// exact_pair() - the very function a tile kernel falls back to when the work list is full, so a pair gets the same bits either way.
template <bool RECT>
__global__ void __launch_bounds__(256) exact_fixup_kernel(const PairParams p, const uint32_t mask)
{
    const unsigned n = min(*p.wl_count, p.wl_capacity);
    for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
        const int2 ij = p.wl_pairs[w];
        const int64_t i = ij.x, j = ij.y;
        double s1, s2;
        exact_pair_inline(p, i, j, s1, s2);                // (inlined: a call would give this usually empty kernel a stack frame)
        // (a launch set of two row blocks shares this kernel: rows of the second block write to its own outputs)
        const bool second = !RECT && p.row_end2 > p.row_begin2 && i >= p.row_begin2 && i < p.row_end2;
        const int64_t idx = (RECT ? i * p.n_cols - p.col_offset : (i * p.N - i * (i + 1) / 2 - i - 1 - (second ? p.p_base2 : p.p_base))) + j;
        double *const *out = second ? p.out2 : p.out;
        if (mask & 2u) {
            double v = s1;
            if (p.normalise) { const double den = p.sum_len[i] + p.sum_len[j]; v = den != 0.0 ? v / den : 0.0; }
            out[1][idx] = v;
        }
        if (mask & 4u) {
            double v = sqrt(s2);
            if (p.normalise) { const double den = p.norm[i] + p.norm[j]; v = den != 0.0 ? v / den : 0.0; }
            out[2][idx] = v;
        }
    }
}

template <uint32_t MASK, bool RECT>
cudaError_t launch_one_persistent(const PairParams &p, unsigned grid, size_t smem, cudaStream_t st)
{
    auto k = p.normalise ? pairjoin_persistent<MASK, RECT, true> : pairjoin_persistent<MASK, RECT, false>;
    cudaError_t e = allow_full_smem((const void *)k);
    if (e != cudaSuccess) return e;
    k<<<grid, NTP, smem, st>>>(p);
    return cudaGetLastError();
}

template <bool RECT>
cudaError_t launch_mask_persistent(uint32_t mask, const PairParams &p, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (mask & 7u) {
        case 1: return launch_one_persistent<1, RECT>(p, grid, smem, st);
        case 2: return launch_one_persistent<2, RECT>(p, grid, smem, st);
        case 3: return launch_one_persistent<3, RECT>(p, grid, smem, st);
        case 4: return launch_one_persistent<4, RECT>(p, grid, smem, st);
        case 5: return launch_one_persistent<5, RECT>(p, grid, smem, st);
        case 6: return launch_one_persistent<6, RECT>(p, grid, smem, st);
        case 7: return launch_one_persistent<7, RECT>(p, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

template <uint32_t MASK, bool RECT>
cudaError_t launch_one(const PairParams &p, unsigned grid, size_t smem, cudaStream_t st)
{
    auto k = pairjoin_kernel<MASK, RECT>;
    cudaError_t e = allow_full_smem((const void *)k);
    if (e != cudaSuccess) return e;
    k<<<grid, NT, smem, st>>>(p);
    return cudaGetLastError();
}

template <bool RECT>
cudaError_t launch_mask(uint32_t mask, const PairParams &p, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (mask & 7u) {
        case 1: return launch_one<1, RECT>(p, grid, smem, st);
        case 2: return launch_one<2, RECT>(p, grid, smem, st);
        case 3: return launch_one<3, RECT>(p, grid, smem, st);
        case 4: return launch_one<4, RECT>(p, grid, smem, st);
        case 5: return launch_one<5, RECT>(p, grid, smem, st);
        case 6: return launch_one<6, RECT>(p, grid, smem, st);
        case 7: return launch_one<7, RECT>(p, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

int pairjoin_table_slots(int max_shared)
{
    int slots = 512;
    while (slots < 2 * TC * max_shared) slots *= 2;
    return slots;
}

// persistent kernel: the table holds the splits of the 64 ROW trees of a tile; `max_tile_entries` is the largest number of
// shared-list entries of any 64-tree tile (an upper bound of its distinct splits); load factor <= 0.75
int pairjoin_row_table_slots(int max_tile_entries)
{
    int slots = 512;
    while (3 * slots < 4 * max_tile_entries) slots *= 2;
    return slots;
}

size_t pairjoin_persistent_smem_bytes(uint32_t mask, int table_slots, int sh_stride)
{
    size_t s = (size_t)table_slots * 12 + TR * TC + (size_t)(TR + TC) * sh_stride * sizeof(SplitRec) + kEventCap * 4 + 16;
    if (mask & 6u) s += 2 * ((mask & 6u) == 6u ? 4 : KC) * (TR + TC) * sizeof(double);
    if (mask & 2u) s += TR * TC * sizeof(double);
    if (mask & 4u) s += TR * TC * sizeof(double);
    return s;
}

size_t pairjoin_smem_bytes(uint32_t mask, int table_slots)
{
    size_t s = (size_t)table_slots * 5 + TR * TC;
    if (mask & 6u) s += (size_t)table_slots * sizeof(double) + 2 * KC * (TR + TC) * sizeof(double);
    if (mask & 2u) s += TR * TC * sizeof(double);
    if (mask & 4u) s += TR * TC * sizeof(double);
    return s;
}

cudaError_t launch_exact_fixup(uint32_t mask, bool rect, const PairParams &p, int n_sms, cudaStream_t st)
{
    // two CTAs per SM (grid-stride over the work list): the list is empty for ordinary data, the launch has to cost next to nothing
    if (rect) exact_fixup_kernel<true><<<2 * n_sms, 256, 0, st>>>(p, mask);
    else exact_fixup_kernel<false><<<2 * n_sms, 256, 0, st>>>(p, mask);
    return cudaGetLastError();
}

// Triangular: rows [row_begin,row_end) -> 64-row tiles from p.tile_row0 (set here); rect: rows [0,row_end), columns from `split`.
cudaError_t launch_pairjoin(uint32_t mask, bool rect, PairParams p, int n_sms, bool persistent, cudaStream_t st)
{
    const size_t smem = pairjoin_smem_bytes(mask, p.table_slots);
    const int64_t col_tiles_total = (p.N + TC - 1) / TC;
    p.tile_row0 = (int)(p.row_begin / TR);
    const int64_t tile_rows = (p.row_end + TR - 1) / TR - p.tile_row0;
    int64_t grid;
    if (rect) {
        p.col_tile0 = (int)(p.col_offset / TC);
        p.tiles_x = (int)(col_tiles_total - p.col_tile0);
        grid = tile_rows * p.tiles_x;
    } else {
        p.col_tile0 = 0;
        p.tiles_x = (int)col_tiles_total;
        const int64_t L0 = col_tiles_total - 2 * p.tile_row0;
        // rows whose first column tile lies beyond the last one contribute nothing (only when N is not a multiple of 64)
        int64_t rows = tile_rows;
        while (rows > 0 && L0 - 2 * (rows - 1) <= 0) rows--;
        grid = rows * L0 - rows * (rows - 1);
    }
    if (grid <= 0) return cudaSuccess;
    if (persistent) {
        const size_t smem = pairjoin_persistent_smem_bytes(mask, p.row_table_slots, p.sh_stride);
        p.n_tiles = grid;
        int64_t ctas = 2 * (int64_t)n_sms;                    // two 384-thread CTAs per SM
        if (ctas > grid) ctas = grid;
        cudaError_t e = rect ? launch_mask_persistent<true>(mask, p, (unsigned)ctas, smem, st)
                             : launch_mask_persistent<false>(mask, p, (unsigned)ctas, smem, st);
        if (e != cudaSuccess || !(mask & 6u)) return e;
        return launch_exact_fixup(mask, rect, p, n_sms, st);
    }
    return rect ? launch_mask<true>(mask, p, (unsigned)grid, smem, st) : launch_mask<false>(mask, p, (unsigned)grid, smem, st);
}

}  // namespace td
