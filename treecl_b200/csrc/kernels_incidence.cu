// kernels_incidence.cu -- exact Robinson-Foulds through an int8 tcgen05 contraction of the split-incidence matrix (sm_100a).
//
// X in {0,1}^{N x D}: X[i][s] = 1 iff tree i holds shared split s (D = splits of the dictionary held by >= 2 trees;
// splits held by one tree can never match).  C = X * X^T counts the common splits of every pair exactly (int32
// accumulation of 0/1 products), and RF(i,j) = S_i + S_j - 2 C_ij  (getRobinsonFouldsDistance, treeCl/treedist.py:113).
// This is the dense-sharing alternative to the merge / hash-join kernels: worthwhile when D is small (clustered gene
// trees), pointless for independent random trees (D ~ N * S).  It is the only place tensor cores are used.
//
// One CTA per 256 x 256 tile of the upper triangle (one CTA per SM: the whole TMEM), 320 threads; two CTAs with horizontally adjacent
// tiles form a cluster and share the A (row) slab: each loads one 128-row half of it and TMA-multicasts that half into both CTAs'
// shared memory, so a CTA pulls 128 + 256 instead of 256 + 256 rows of K through L2 per k-step.  The 128 x 128 tiles of the first
// version were L2-bound (ncu, profiles/r2_incidence_summary.txt: L2 72 %, tensor pipe 29 %): every tile pulls 2 x 128 rows of K
// through L2; a 256 x 256 tile pulls 2 x 256 rows for four times the MMAs, i.e. half the L2 bytes per operation.
//   warp 0     TMA producer : cp.async.bulk.tensor.2d (SWIZZLE_128B) of the A (256 rows) and B (256 columns) k-slabs, 128 B of K each,
//                             into a 3-stage shared-memory ring (64 KB per stage); completion on `full` mbarriers
//   warp 1     MMA issuer   : one elected thread issues tcgen05.mma.cta_group::1.kind::i8 (M=128, N=256, K=32), two per k-step (the
//                             two 128-row halves of A), accumulating in TMEM (128 lanes x 2 x 256 columns of int32); tcgen05.commit
//                             frees the ring slot (`empty` mbarriers) and finally signals `tmem_full`
//   warps 2-17 epilogue     : tcgen05.ld 32x32b.x32 (one TMEM lane = one row per thread; four warps per lane quarter: upper / lower half
//                             of the tile x left / right 128 columns), RF transform, transposition through shared memory, condensed store
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace td {

namespace {

constexpr int TILE = 256;            // rows and columns of an output tile: two UMMA M = 128 halves x UMMA N = 256
constexpr int HALF = 128;            // UMMA M
constexpr int BK = 128;              // bytes (= int8 elements) of K per stage = one 128-byte swizzle row
constexpr int UMMA_K = 32;           // K per tcgen05.mma for 8-bit operands
constexpr int STAGES = 3;
constexpr int NTHREADS = 576;        // TMA warp, MMA warp, 16 epilogue warps
constexpr uint32_t STAGE_BYTES = 2 * TILE * BK;      // A + B slab
constexpr uint32_t TMEM_COLS = 512;  // 2 halves x 256 int32 columns

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t spins = 0;; spins++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 22)) __trap();          // watchdog: a lost TMA / MMA completion must not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows 128 B apart, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start [0,14), LBO [16,30) = 1, SBO [32,46) = 1024 >> 4, version [46,48) = 1, layout [61,64) = 2)
__device__ __forceinline__ uint64_t umma_desc(const void *smem)
{
    const uint64_t addr = (smem_u32(smem) & 0x3FFFFu) >> 4;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(HALF >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same arrival on the barrier at this shared-memory offset in every CTA of `mask` (both CTAs of the pair wait for both consumers
// before a ring slot - which both producers write into - is refilled)
__device__ __forceinline__ void umma_commit_multicast(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint16_t mask)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// ---- cta_group::2 forms: the two CTAs of a cluster (one TPC) run ONE MMA of M = 256: each CTA supplies its own 128 rows of A and HALF of
// the N = 256 rows of B (the tensor cores read the other half from the partner's shared memory), accumulators of a CTA's rows in its own
// TMEM.  Only the even CTA (rank 0, the leader) issues MMAs and commits; the loads of both CTAs report to the leader's barrier.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the even CTA of the pair
constexpr uint32_t kIdesc2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)((2 * HALF) >> 4) << 24);   // M = 256
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t accumulate)
{
    const uint32_t z = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_c), "l"(da), "l"(db), "r"(kIdesc2), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// clusters (tile pairs) before tile row u when the first row has L0 tiles: sum over v < u of ceil((L0 - v) / 2)
__host__ __device__ __forceinline__ long long pair_prefix(long long L0, long long u)
{
    auto S = [](long long m) { const long long q = m / 2; return q * (q + 1) + (m & 1) * (q + 1); };      // sum_{k=1..m} ceil(k / 2)
    return S(L0) - S(L0 - u);
}

struct IncParams {
    const int32_t *nsplits;
    int64_t N;
    int64_t row_begin, row_end, p_base;
    int tile_row0, tiles;        // first tile row, tile columns in total
    int k_blocks;                // Dpad / BK
    int normalise, out_u16;
    int multicast;               // 1: the two CTAs of a cluster each load half of the shared A slab and multicast it; 0: every CTA loads its own
    void *out;
};

// TWO_SM = false: every CTA a 256 x 256 tile of its own (cta_group::1), the two CTAs of a cluster share the A slab by TMA multicast.
// TWO_SM = true : the cluster computes a 512 x 256 tile with cta_group::2 MMAs - CTA `rank` owns rows [256 rank, 256 rank + 256) of it and
//                 loads those 256 rows of A plus 128 of the 256 rows of B per k-step: 48 KB instead of 64 KB through its shared-memory
//                 port for the same 256 x 256 x 128 products (the port, ~64 B / clk / SM, is what bounds the 1-SM form; ncu summary in
//                 profiles/), with a 4-stage ring.
template <bool TWO_SM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
rf_incidence_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_half, const IncParams p)
{
    constexpr uint32_t kStageBytes = TWO_SM ? (uint32_t)((TILE + HALF) * BK) : STAGE_BYTES;       // A 256 rows + B 128 / 256 rows
    constexpr int kStages = TWO_SM ? 4 : STAGES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // [stages][A 32 KB | B 16 / 32 KB]; the 128-byte swizzle wants 1024-byte aligned slabs (the launch reserves the slack)
    unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    static_assert((uint32_t)kStages * kStageBytes == (uint32_t)STAGES * STAGE_BYTES, "both forms use the same 192 KB ring");
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + STAGES * STAGE_BYTES);
    uint64_t *empty = full + 4;
    uint64_t *tmem_full = empty + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    int32_t *s_cols = reinterpret_cast<int32_t *>(tmem_slot + 2);        // S_j of the tile's 256 columns

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // upper-triangular tile list, row-major from tile row p.tile_row0.  1-SM form: two adjacent tiles of a row per cluster (a row with an
    // odd number of tiles ends with a pair whose second tile lies beyond the matrix: it loads zeros and stores nothing).  2-SM form: a
    // cluster is a 512-row x 256-column tile, i.e. tile rows 2R and 2R + 1 of one tile column c >= 2R (its lower half lies below the
    // diagonal when c == 2R: computed, never stored).
    const uint32_t rank = cluster_ctarank();
    int tile_r, tile_c;
    if constexpr (TWO_SM) {
        const long long c = blockIdx.x >> 1, T = p.tiles - p.tile_row0;             // p.tile_row0 is even here
        long long lo = 0, hi = (T + 1) / 2;                                          // largest R with R T - R (R - 1) <= c
        while (lo + 1 < hi) { const long long mid = (lo + hi) >> 1; if (mid * T - mid * (mid - 1) <= c) lo = mid; else hi = mid; }
        tile_r = p.tile_row0 + 2 * (int)lo + (int)rank;
        tile_c = p.tile_row0 + 2 * (int)lo + (int)(c - (lo * T - lo * (lo - 1)));
    } else {
        const long long c = blockIdx.x >> 1, L0 = p.tiles - p.tile_row0;
        long long lo = 0, hi = L0;                                       // largest u with pair_prefix(u) <= c
        while (lo + 1 < hi) { const long long mid = (lo + hi) >> 1; if (pair_prefix(L0, mid) <= c) lo = mid; else hi = mid; }
        tile_r = (int)lo + p.tile_row0;
        tile_c = tile_r + 2 * (int)(c - pair_prefix(L0, lo)) + (int)rank;
    }
    const int64_t i0 = (int64_t)tile_r * TILE, j0 = (int64_t)tile_c * TILE;

    if (threadIdx.x == 0) {
        // empty: 1-SM multicast form: both CTAs' consumers release a slot; 2-SM form: the leader's commit, multicast to both CTAs
        for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], (!TWO_SM && p.multicast) ? 2 : 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                                     // TMEM: 512 columns x 128 lanes of int32 (all of it)
        if constexpr (TWO_SM) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + TILE) { const int c = threadIdx.x - 64; s_cols[c] = (j0 + c < p.N) ? p.nsplits[j0 + c] : 0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                                  // the partner's barriers exist before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < p.k_blocks; kb++) {
                const int s = kb % kStages; const uint32_t round = kb / kStages;
                mbar_wait(&empty[s], (round & 1u) ^ 1u);                 // slot free (first pass returns immediately)
                unsigned char *a = ring + (size_t)s * kStageBytes, *b = a + TILE * BK;
                if constexpr (TWO_SM) {
                    // both CTAs' loads report to the LEADER's barrier, which expects the bytes of both (its own producer arms it)
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * kStageBytes);
                    tma_load_2d_2sm(a, &tmap, kb * BK, (int)i0, &full[s]);                                       // my 256 rows of A
                    tma_load_2d_2sm(b, &tmap_half, kb * BK, (int)j0 + (int)rank * HALF, &full[s]);               // my half of the 256 rows of B
                } else {
                    mbar_expect_tx(&full[s], kStageBytes);
                    // my half of the shared A slab goes to both CTAs (each barrier receives both halves and its own B slab)
                    if (p.multicast) tma_load_2d_multicast(a + rank * HALF * BK, &tmap_half, kb * BK, (int)i0 + (int)rank * HALF, &full[s], (uint16_t)3);
                    else tma_load_2d(a, &tmap, kb * BK, (int)i0, &full[s]);
                    tma_load_2d(b, &tmap, kb * BK, (int)j0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && (!TWO_SM || rank == 0)) {
            for (int kb = 0; kb < p.k_blocks; kb++) {
                const int s = kb % kStages; const uint32_t round = kb / kStages;
                mbar_wait(&full[s], round & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned char *a = ring + (size_t)s * kStageBytes, *b = a + TILE * BK;
                const uint64_t da = umma_desc(a), db = umma_desc(b);
                const uint64_t da2 = umma_desc(a + HALF * BK);           // rows 128..255 of the A slab
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; k++) {                  // +32 bytes of K inside the swizzled row: start address += 2
                    const uint64_t ko = (uint64_t)(k * (UMMA_K >> 4));
                    if constexpr (TWO_SM) {
                        umma_i8_2sm(tmem_base, da + ko, db + ko, (kb | k) ? 1u : 0u);
                        umma_i8_2sm(tmem_base + TILE, da2 + ko, db + ko, (kb | k) ? 1u : 0u);
                    } else {
                        umma_i8(tmem_base, da + ko, db + ko, (kb | k) ? 1u : 0u);
                        umma_i8(tmem_base + TILE, da2 + ko, db + ko, (kb | k) ? 1u : 0u);
                    }
                }
                // the slot is free (in both CTAs) once the MMAs above have read it
                if constexpr (TWO_SM) umma_commit_2sm(&empty[s], (uint16_t)3);
                else if (p.multicast) umma_commit_multicast(&empty[s], (uint16_t)3);
                else umma_commit(&empty[s]);
            }
            if constexpr (TWO_SM) umma_commit_2sm(tmem_full, (uint16_t)3);          // accumulators complete, in both CTAs
            else umma_commit(tmem_full);
        }
    } else {
        // ---- epilogue: this warp may touch TMEM lanes 32*(warp % 4) .. +31.  Sixteen warps: four per lane quarter, one for each
        // (half of the tile = 256 TMEM columns) x (128 of its 256 columns): a warp drains four 32-column chunks, and the TMEM-load ->
        // transposition -> store chains of sixteen warps overlap
        const int q = warp & 3, part = (warp - 2) >> 2, half = part >> 1, col_lo = (part & 1) * (TILE / 2);
        const int row = HALF * half + 32 * q + lane;
        const int64_t i = i0 + row;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool row_ok = i >= p.row_begin && i < p.row_end && i < p.N;
        const int si = row_ok ? p.nsplits[i] : 0;
        const int64_t row_off = i * p.N - i * (i + 1) / 2 - i - 1 - p.p_base;
#pragma unroll 1
        for (int c0 = col_lo; c0 < col_lo + TILE / 2; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(TILE * half + c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int64_t jb = j0 + c0;
            if (!p.normalise) {
                // Transposed store.  After tcgen05.ld a lane holds 32 columns of ONE row: storing from that layout makes every store
                // instruction touch 32 different rows (32 sectors for 64 or 128 useful bytes).  The 32 x 32 block goes through shared
                // memory (the ring is idle by now) so that a lane holds one COLUMN and a store instruction writes one contiguous row
                // segment: 64 B (u16) or 256 B (f64) in 2-3 or 8 sectors.
                int *blk = reinterpret_cast<int *>(ring) + (warp - 2) * (32 * 33);
#pragma unroll
                for (int k = 0; k < 32; k++) blk[lane * 33 + k] = si + s_cols[c0 + k] - 2 * (int)v[k];
                __syncwarp();
                const int64_t j = jb + lane;
                const int64_t r0 = i0 + HALF * half + 32 * q;                       // first row of the block
                int64_t off = r0 * p.N - r0 * (r0 + 1) / 2 - r0 - 1 - p.p_base + j;  // condensed index of (r0, j); next row: + N - r - 2
#pragma unroll 8
                for (int rr = 0; rr < 32; rr++) {
                    const int64_t r = r0 + rr;
                    const int rf = blk[rr * 33 + lane];
                    if (r >= p.row_begin && r < p.row_end && j > r && j < p.N) {
                        if (p.out_u16) reinterpret_cast<uint16_t *>(p.out)[off] = (uint16_t)rf;
                        else reinterpret_cast<double *>(p.out)[off] = (double)rf;
                    }
                    off += p.N - r - 2;
                }
                __syncwarp();
            } else if (row_ok) {
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    const int64_t j = jb + k;
                    if (j <= i || j >= p.N) continue;
                    const int den = si + s_cols[c0 + k];
                    const int rf = den - 2 * (int)v[k];
                    double d = (double)rf;
                    d = den ? d / (double)den : 0.0;
                    reinterpret_cast<double *>(p.out)[row_off + j] = d;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                                  // no CTA leaves while its partner may still write to it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if constexpr (TWO_SM) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// X[i][sid] = 1 for every shared split of every tree (X is zeroed first); sid travels in SplitRec::pad
__global__ void build_incidence_kernel(const SplitRec *__restrict__ shared, int sh_stride, int64_t N, int64_t ldx, int8_t *__restrict__ X)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * sh_stride) return;
    const SplitRec r = shared[e];
    if (r.id != kSentinelId) X[(e / sh_stride) * ldx + r.pad] = 1;
}

}  // namespace

size_t incidence_smem_bytes() { return (size_t)STAGES * STAGE_BYTES + (2 * 4 + 1) * 8 + 16 + TILE * 4 + 1024; }
int incidence_tile() { return TILE; }
int incidence_bk() { return BK; }

cudaError_t launch_build_incidence(const SplitRec *shared, int sh_stride, int64_t N, int64_t ldx, int8_t *X, cudaStream_t st)
{
    const int64_t n = N * sh_stride;
    if (n <= 0) return cudaSuccess;
    build_incidence_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(shared, sh_stride, N, ldx, X);
    return cudaGetLastError();
}

// 2-D tensor map over X [rows][ldx] int8, box 128 (K bytes) x 256 (rows), 128-byte swizzle.  Returns cudaErrorNotSupported if
// the driver entry point is unavailable.
cudaError_t make_incidence_tensor_map(const int8_t *X, int64_t rows, int64_t ldx, void *tmap_out, int box_rows)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    const cuuint64_t dims[2] = {(cuuint64_t)ldx, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx};                       // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(box_rows > 0 ? box_rows : TILE)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(reinterpret_cast<CUtensorMap *>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(X), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

int incidence_half() { return HALF; }

cudaError_t launch_rf_incidence(const void *tmap, const void *tmap_half, const int32_t *nsplits, int64_t N, int64_t d_pad, int64_t row_begin, int64_t row_end,
                                int64_t p_base, int normalise, int out_u16, void *out, cudaStream_t st)
{
    IncParams p{};
    p.nsplits = nsplits; p.N = N; p.row_begin = row_begin; p.row_end = row_end; p.p_base = p_base;
    p.tile_row0 = (int)(row_begin / TILE); p.tiles = (int)((N + TILE - 1) / TILE);
    p.k_blocks = (int)(d_pad / BK); p.normalise = normalise; p.out_u16 = out_u16; p.out = out;
    { const char *e = getenv("TREEDIST_INC_MULTICAST"); p.multicast = e ? atoi(e) != 0 : 1; }
    if (getenv("TREEDIST_INC_DEBUG_NOSTORE")) p.row_end = p.row_begin;      // measurement aid: the whole kernel minus its stores (every row masked)
    if (p.k_blocks <= 0) return cudaSuccess;
    const size_t smem = incidence_smem_bytes();
    CUtensorMap map, map_half; memcpy(&map, tmap, sizeof map); memcpy(&map_half, tmap_half, sizeof map_half);
    // TREEDIST_INC_2SM=0: the 1-SM form (every CTA a 256 x 256 tile); default: 512 x 256 tiles by cta_group::2 pairs
    const char *two = getenv("TREEDIST_INC_2SM");
    if (!two || atoi(two) != 0) {
        p.tile_row0 &= ~1;                                                // pair tiles start at even tile rows (rows before row_begin are masked)
        const int64_t T = p.tiles - p.tile_row0, pair_rows = ((row_end + TILE - 1) / TILE - p.tile_row0 + 1) / 2;
        const int64_t clusters = pair_rows * T - pair_rows * (pair_rows - 1);
        if (clusters <= 0) return cudaSuccess;
        cudaError_t e = allow_full_smem((const void *)rf_incidence_kernel<true>);
        if (e != cudaSuccess) return e;
        rf_incidence_kernel<true><<<(unsigned)(2 * clusters), NTHREADS, smem, st>>>(map, map_half, p);
        return cudaGetLastError();
    }
    const int64_t tile_rows = (row_end + TILE - 1) / TILE - p.tile_row0;
    const int64_t L0 = p.tiles - p.tile_row0;
    const int64_t clusters = pair_prefix(L0, tile_rows);
    if (clusters <= 0) return cudaSuccess;
    cudaError_t e = allow_full_smem((const void *)rf_incidence_kernel<false>);
    if (e != cudaSuccess) return e;
    rf_incidence_kernel<false><<<(unsigned)(2 * clusters), NTHREADS, smem, st>>>(map, map_half, p);
    return cudaGetLastError();
}

}  // namespace td
