// kernels_pairsplit.cu -- rf / wrf / euc for sparsely sharing sets: postings pipeline + fp64 tensor-core tile kernel (sm_100a).
//
// Same contract as kernels_pairwise.cu (getRobinsonFoulds / getWeightedRobinsonFoulds / getEuclideanDistance of
// treeCl/treedist.py:111-114 for all pairs of a row range), split into a sparse and a dense part:
//
//   1. events.  A pair (i, j) has a common split d  <=>  i and j both appear in the POSTINGS of d (the encoder groups the
//      shared-list entries by split, trees ascending).  pair_postings_kernel walks the postings with one warp per entry x = (d, i):
//      every later entry (d, j) of the same group is one EVENT of pair (i, j) carrying the corrections
//          cnt += 1,   E2 += -2 la lb,   E1 += |la - lb| - |la| - |lb|.
//      Work is proportional to the number of events (sum over pairs of their common splits), no searching at all.  Events are
//      binned by 64 x 32 tile with a counting sort: count pass (atomicAdd per tile), exclusive scan, fill pass.
//   2. tiles.  pair_dense_kernel, one CTA per tile: the tile's events are applied to per-cell slots in shared memory, the leaf
//      edges go through the fp64 tensor cores in the dot-product form
//          euc^2 = (Q2_i + Q2_j) - 2 sum_k a_k b_k + E2           (Q2 = sum of squares of ALL edge lengths of a tree)
//      (DMMA m8n8k4: the same 37 TFLOP/s as DFMA but 1/8 of the issue slots, and one FMA per term instead of DADD + DFMA),
//      wrf keeps the direct form sum_k |a_k - b_k| on the fp64 pipe, and the epilogue leaves through shared memory so that every
//      warp writes whole 256-byte row segments of the condensed matrix.
// (Q + E) subtracts; cells whose correction nearly cancels Q, or that received >= 3 events (the order of >= 3 shared-memory
// atomic adds is not fixed), are queued for exact_fixup_kernel, which recomputes them additively in a fixed order: results are
// bit-reproducible, independent of the row sharding, and identical trees give exactly 0.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace td {

namespace {

constexpr int TR = 64, TC = 32;       // tile: rows x columns
constexpr int NT = 256;
constexpr int KCH = 16;               // taxa per staged chunk
constexpr int SA = 68, SB = 36;       // row strides (doubles) of the staged leaf chunks: = 4 mod 16 -> conflict-free DMMA fragments
constexpr int ES = 40;                // row stride (doubles) of the per-cell slots: conflict-free 16-byte accesses of the C fragments
constexpr double kRisk1 = 0.9;        // wrf: exact recompute when -E1 > kRisk1 * (Q1_i + Q1_j)
constexpr double kRisk2 = 0.98;       // euc: exact recompute when the result is < 2 % of Q2_i + Q2_j (error amplification <= 50)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// D(8x8) += A(8x4, row) * B(4x8, col), fp64.  Lane l = 4 g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <bool RECT>
__device__ __forceinline__ void decode_tile(const PairParams &p, long long t, int &tile_r, int &tile_c)
{
    if (RECT) {
        tile_r = (int)(t / p.tiles_x) + p.tile_row0;
        tile_c = (int)(t % p.tiles_x) + p.col_tile0;
    } else {
        // tile row u (from p.tile_row0) holds column tiles 2*tile_r .. tiles_x-1: L0 - 2u of them, L0 = tiles_x - 2*tile_row0
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

// ------------------------------------------------------------------------------------------------------------------------------
// events from the postings.  FILL = false: count per tile;  FILL = true: write the events (tile_cnt was cleared by the scan)
// ------------------------------------------------------------------------------------------------------------------------------
template <bool WEIGHTED, bool RECT, bool FILL>
__global__ void __launch_bounds__(256) pair_postings_kernel(const PairParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long L0 = p.tiles_x - 2 * p.tile_row0;
    for (long long x = warp; x < p.n_post; x += n_warps) {
        const int i = p.post_tree[x];
        if (i < p.row_begin || i >= p.row_end) continue;
        const int end = p.post_end[x];
        const int tile_r = i >> 6;
        const long long u = tile_r - p.tile_row0;
        const long long base_t = RECT ? u * p.tiles_x - p.col_tile0 : u * L0 - u * (u - 1) - 2 * tile_r;
        double la = 0.0;
        if constexpr (WEIGHTED && FILL) la = p.post_len[x];
        for (long long b = x + 1 + lane; b < end; b += 32) {
            const int j = p.post_tree[b];
            uint32_t *c = p.tile_cnt + (base_t + (j >> 5));
            if (RECT && j < p.col_offset) {}                       // the other set's own side: not a pair of this launch
            else if constexpr (!FILL) atomicAdd(c, 1u);
            else {
                const uint32_t o = p.tile_off[base_t + (j >> 5)] + atomicAdd(c, 1u);
                if (o < p.ev_capacity) {
                    p.ev_cell[o] = (uint16_t)(((i & 63) << 5) | (j & 31));
                    if constexpr (WEIGHTED) {
                        const double lb = p.post_len[b];
                        p.ev_t2[o] = -2.0 * la * lb;
                        p.ev_t1[o] = fabs(la - lb) - fabs(la) - fabs(lb);
                    }
                }
            }
        }
    }
}

// exclusive scan of tile_cnt[0..n) into tile_off[0..n], tile_cnt cleared for the fill pass.  One block; every thread owns a
// contiguous piece.
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t *cnt, uint32_t *off, long long n)
{
    __shared__ uint32_t warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long per = (n + 1023) / 1024, lo = tid * per < n ? tid * per : n, hi = lo + per < n ? lo + per : n;
    uint32_t s = 0;
    for (long long e = lo; e < hi; e++) s += cnt[e];
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sum[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += v; }
        warp_sum[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = warp_sum[warp] + incl - s;
    for (long long e = lo; e < hi; e++) { const uint32_t c = cnt[e]; off[e] = run; run += c; cnt[e] = 0; }
    if (tid == 1023) off[n] = run;
}

// ------------------------------------------------------------------------------------------------------------------------------
// tiles
// ------------------------------------------------------------------------------------------------------------------------------
// exact, purely additive recomputation of one pair (merge of the two shared-split lists + strided walk over the leaf matrix)
__device__ __noinline__ void exact_pair(const PairParams &p, int64_t i, int64_t j, double &s1, double &s2)
{
    s1 = 0.0; s2 = 0.0;
    const SplitRec *A = p.shared + i * p.sh_stride, *B = p.shared + j * p.sh_stride;
    SplitRec a = *A, b = *B;
    while ((a.id & b.id) != kSentinelId) {
        const bool adv_a = a.id <= b.id, adv_b = b.id <= a.id;
        const double la = adv_a ? a.len : 0.0, lb = adv_b ? b.len : 0.0;
        const double d = la - lb;
        s1 += fabs(d); s2 = fma(d, d, s2);
        if (adv_a) a = *++A;
        if (adv_b) b = *++B;
    }
    s1 += p.single_l1[i] + p.single_l1[j];
    s2 += p.single_l2[i] + p.single_l2[j];
    for (int k = 0; k < p.n_k; k++) {
        const double d = p.leafT[(size_t)k * p.ldT + i] - p.leafT[(size_t)k * p.ldT + j];
        s1 += fabs(d); s2 = fma(d, d, s2);
    }
}

// rare path, kept out of line: queue the pair for the exact fix-up kernel (same stream, right after this one); when the work
// list is full the exact sums are computed here instead
__device__ __noinline__ void queue_exact(const PairParams &p, int64_t i, int64_t j, double &t1, double &t2)
{
    const unsigned slot = atomicAdd(p.wl_count, 1u);
    if (slot < p.wl_capacity) p.wl_pairs[slot] = make_int2((int)i, (int)j);
    else exact_pair(p, i, j, t1, t2);
}

template <uint32_t MASK, bool RECT, bool NORM>
__global__ void __launch_bounds__(NT, 3) pair_dense_kernel(const __grid_constant__ PairParams p)
{
    constexpr bool kRF = MASK & 1u, kWRF = (MASK >> 1) & 1u, kEUC = (MASK >> 2) & 1u;
    constexpr bool kWeighted = kWRF || kEUC;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *cur = smem_raw;
    double *E1 = nullptr, *E2 = nullptr, *sA = nullptr, *sB = nullptr, *sQ1 = nullptr, *sQ2 = nullptr;
    if constexpr (kWRF) { E1 = reinterpret_cast<double *>(cur); cur += TR * ES * sizeof(double); }
    if constexpr (kEUC) { E2 = reinterpret_cast<double *>(cur); cur += TR * ES * sizeof(double); }
    if constexpr (kWeighted) {
        sA = reinterpret_cast<double *>(cur); cur += 2 * KCH * SA * sizeof(double);
        sB = reinterpret_cast<double *>(cur); cur += 2 * KCH * SB * sizeof(double);
    }
    if constexpr (kWRF) { sQ1 = reinterpret_cast<double *>(cur); cur += (TR + TC) * sizeof(double); }     // rows, then columns
    if constexpr (kEUC) { sQ2 = reinterpret_cast<double *>(cur); cur += (TR + TC) * sizeof(double); }
    int *sS = reinterpret_cast<int *>(cur); cur += (TR + TC) * sizeof(int);                                // internal edges per tree
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>(cur);
    const unsigned char *cnt = cur;                                                                        // [TR*TC] common splits

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;            // DMMA fragment coordinates
    const int wr = warp >> 1, wc = warp & 1;           // warp tile: rows 16 wr .., columns 16 wc ..
    const long long t = blockIdx.x;
    const uint32_t ev0 = p.tile_off[t], ev1 = p.tile_off[t + 1];
    int tile_r, tile_c;
    decode_tile<RECT>(p, t, tile_r, tile_c);
    const int64_t i0 = (int64_t)tile_r * TR, j0 = (int64_t)tile_c * TC;

    // ---- leaf-length chunks start flowing --------------------------------------------------------------------------------------
    const int k_used = kWeighted ? (p.n_taxa + 3) & ~3 : 0;                  // <= n_k; rows >= n_taxa of leafT are zero
    const int n_chunks = (k_used + KCH - 1) / KCH;
    auto load_chunk = [&](int chunk, int stage) {
        // 16-byte pieces: KCH rows x (32 of A + 16 of B)
        for (int e = tid; e < KCH * 48; e += NT) {
            const int kk = e / 48, c = e % 48, k = chunk * KCH + kk;
            if (k < k_used) {
                const double *src = p.leafT + (size_t)k * p.ldT;
                if (c < 32) cp_async16(sA + (size_t)(stage * KCH + kk) * SA + 2 * c, src + i0 + 2 * c);
                else cp_async16(sB + (size_t)(stage * KCH + kk) * SB + 2 * (c - 32), src + j0 + 2 * (c - 32));
            }
        }
    };
    if constexpr (kWeighted) {
        if (n_chunks > 0) load_chunk(0, 0);
        cp_async_commit();
        if (n_chunks > 1) load_chunk(1, 1);
        cp_async_commit();
    }
    // ---- cells start cleared, per-tree scalars go to shared memory -----------------------------------------------------------------
    {
        if constexpr (kWRF) { double2 *z = reinterpret_cast<double2 *>(E1); for (int e = tid; e < TR * ES / 2; e += NT) z[e] = make_double2(0.0, 0.0); }
        if constexpr (kEUC) { double2 *z = reinterpret_cast<double2 *>(E2); for (int e = tid; e < TR * ES / 2; e += NT) z[e] = make_double2(0.0, 0.0); }
        for (int e = tid; e < TR * TC / 4; e += NT) cnt32[e] = 0u;
        if (tid < TR + TC) {
            const int64_t tree = tid < TR ? i0 + tid : j0 + (tid - TR);
            sS[tid] = p.nsplits[tree];
            if constexpr (kWRF) sQ1[tid] = p.q1[tree];
            if constexpr (kEUC) sQ2[tid] = p.q2t[tree];
        }
    }
    __syncthreads();
    // ---- this tile's events ------------------------------------------------------------------------------------------------------
    for (uint32_t o = ev0 + tid; o < ev1; o += NT) {
        const int cell = p.ev_cell[o];
        atomicAdd(&cnt32[cell >> 2], 1u << (8 * (cell & 3)));
        const int slot = (cell >> 5) * ES + (cell & 31);
        if constexpr (kEUC) atomicAdd(&E2[slot], p.ev_t2[o]);
        if constexpr (kWRF) atomicAdd(&E1[slot], p.ev_t1[o]);
    }

    // ---- leaf edges: dot products on the fp64 tensor cores (euc), |a - b| sums on the fp64 pipe (wrf) ------------------------------
    double dot[2][2][2] = {}, l1[2][2][2] = {};        // [row block][column block][element] of the warp's 16 x 16 tile
    if constexpr (kWeighted) {
        for (int chunk = 0; chunk < n_chunks; chunk++) {
            const int stage = chunk & 1;
            cp_async_wait<1>();
            __syncthreads();
            const double *cA = sA + (size_t)stage * KCH * SA + wr * 16 + g, *cB = sB + (size_t)stage * KCH * SB + wc * 16;
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
            __syncthreads();
            if (chunk + 2 < n_chunks) load_chunk(chunk + 2, stage);
            cp_async_commit();
        }
        if (n_chunks == 0) __syncthreads();
        // fold the leaf sums into the cells this thread owns (all events have landed: every thread passed a barrier after its atomics)
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int slot = (wr * 16 + r * 8 + g) * ES + wc * 16 + c * 8 + 2 * t4;
                if constexpr (kEUC) {
                    double2 *q = reinterpret_cast<double2 *>(E2 + slot);
                    double2 v = *q; v.x = fma(-2.0, dot[r][c][0], v.x); v.y = fma(-2.0, dot[r][c][1], v.y); *q = v;
                }
                if constexpr (kWRF) {
                    double2 *q = reinterpret_cast<double2 *>(E1 + slot);
                    double2 v = *q; v.x += l1[r][c][0]; v.y += l1[r][c][1]; *q = v;
                }
            }
    }
    __syncthreads();

    // ---- out: warp w writes rows 8w .. 8w+7, lane = column: whole 256-byte row segments ----------------------------------------------
    const int64_t j = j0 + lane;
    const int s_j = sS[TR + lane];
    double q1j = 0.0, q2j = 0.0, sumj = 0.0, normj = 0.0;
    if constexpr (kWRF) q1j = sQ1[TR + lane];
    if constexpr (kEUC) q2j = sQ2[TR + lane];
    if constexpr (NORM) {
        if constexpr (kWRF) sumj = p.sum_len[j];
        if constexpr (kEUC) normj = p.norm[j];
    }
#pragma unroll 2
    for (int rr = 0; rr < 8; rr++) {
        const int row = warp * 8 + rr;
        const int64_t i = i0 + row;
        if (i < p.row_begin || i >= p.row_end) continue;                                   // warp-uniform
        if (j >= p.col_end || (RECT ? j < p.col_offset : j <= i)) {}
        else {
        const int64_t idx = (RECT ? i * p.n_cols - p.col_offset : (i * p.N - i * (i + 1) / 2 - i - 1 - p.p_base)) + j;
        const int common = cnt[row * TC + lane];
        if constexpr (kRF) {
            const int den = sS[row] + s_j;
            double v = (double)(den - 2 * common);
            if constexpr (NORM) v = den ? v / (double)den : 0.0;
            p.out[0][idx] = v;
        }
        if constexpr (kWeighted) {
            double t1 = 0.0, t2 = 0.0;
            bool risky = common >= 3;                      // the order of >= 3 atomic adds is not fixed
            if constexpr (kWRF) {
                const double q = sQ1[row] + q1j, e = E1[row * ES + lane];
                risky |= -e > kRisk1 * q;                  // e already holds the (non-negative) leaf sum: only makes the test milder
                t1 = q + e;
            }
            if constexpr (kEUC) {
                const double q = sQ2[row] + q2j, e = E2[row * ES + lane];
                risky |= -e > kRisk2 * q;
                t2 = fmax(q + e, 0.0);
            }
            if (risky) queue_exact(p, i, j, t1, t2);       // normally overwritten by the fix-up kernel
            if constexpr (kWRF) {
                double v = t1;
                if constexpr (NORM) { const double den = p.sum_len[i] + sumj; v = den != 0.0 ? v / den : 0.0; }
                p.out[1][idx] = v;
            }
            if constexpr (kEUC) {
                double v = sqrt(t2);
                if constexpr (NORM) { const double den = p.norm[i] + normj; v = den != 0.0 ? v / den : 0.0; }
                p.out[2][idx] = v;
            }
        }
        }
    }
}

template <uint32_t MASK, bool RECT>
cudaError_t launch_dense(const PairParams &p, unsigned grid, size_t smem, cudaStream_t st)
{
    auto k = p.normalise ? pair_dense_kernel<MASK, RECT, true> : pair_dense_kernel<MASK, RECT, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, NT, smem, st>>>(p);
    return cudaGetLastError();
}

template <bool RECT>
cudaError_t launch_dense_mask(uint32_t mask, const PairParams &p, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (mask & 7u) {
        case 1: return launch_dense<1, RECT>(p, grid, smem, st);
        case 2: return launch_dense<2, RECT>(p, grid, smem, st);
        case 3: return launch_dense<3, RECT>(p, grid, smem, st);
        case 4: return launch_dense<4, RECT>(p, grid, smem, st);
        case 5: return launch_dense<5, RECT>(p, grid, smem, st);
        case 6: return launch_dense<6, RECT>(p, grid, smem, st);
        case 7: return launch_dense<7, RECT>(p, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

template <bool WEIGHTED, bool RECT>
cudaError_t launch_postings(const PairParams &p, bool fill, unsigned grid, cudaStream_t st)
{
    if (fill) pair_postings_kernel<WEIGHTED, RECT, true><<<grid, 256, 0, st>>>(p);
    else pair_postings_kernel<false, RECT, false><<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

size_t pairsplit_dense_smem_bytes(uint32_t mask)
{
    size_t s = TR * TC + (TR + TC) * sizeof(int);
    if (mask & 6u) s += 2 * KCH * (SA + SB) * sizeof(double);
    if (mask & 2u) s += (TR * ES + TR + TC) * sizeof(double);
    if (mask & 4u) s += (TR * ES + TR + TC) * sizeof(double);
    return s;
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
    p.n_tiles = grid > 0 ? grid : 0;
    return p.n_tiles;
}

// p must carry the tile geometry (pairsplit_tiles), the postings and the event buffers; tile_cnt must be zero
cudaError_t launch_pairsplit(uint32_t mask, bool rect, const PairParams &p, int n_sms, cudaStream_t st, int *launches)
{
    *launches = 0;
    if (p.n_tiles <= 0) return cudaSuccess;
    const bool weighted = (mask & 6u) != 0;
    cudaError_t e = cudaSuccess;
    if (p.n_post > 0) {
        long long blocks = (p.n_post + 7) / 8;                      // one warp per postings entry, capped at a few waves
        if (blocks > 16ll * n_sms) blocks = 16ll * n_sms;
        for (int fill = 0; fill < 2 && e == cudaSuccess; fill++) {
            if (weighted) e = rect ? launch_postings<true, true>(p, fill, (unsigned)blocks, st) : launch_postings<true, false>(p, fill, (unsigned)blocks, st);
            else e = rect ? launch_postings<false, true>(p, fill, (unsigned)blocks, st) : launch_postings<false, false>(p, fill, (unsigned)blocks, st);
            ++*launches;
            if (e == cudaSuccess && !fill) { tile_scan_kernel<<<1, 1024, 0, st>>>(p.tile_cnt, p.tile_off, p.n_tiles); e = cudaGetLastError(); ++*launches; }
        }
        if (e != cudaSuccess) return e;
    } else {
        e = cudaMemsetAsync(p.tile_off, 0, (size_t)(p.n_tiles + 1) * 4, st);
        if (e != cudaSuccess) return e;
    }
    const size_t smem_d = pairsplit_dense_smem_bytes(mask);
    e = rect ? launch_dense_mask<true>(mask, p, (unsigned)p.n_tiles, smem_d, st) : launch_dense_mask<false>(mask, p, (unsigned)p.n_tiles, smem_d, st);
    ++*launches;
    if (e != cudaSuccess || !weighted) return e;
    ++*launches;
    return launch_exact_fixup(mask, rect, p, n_sms, st);
}

}  // namespace td
