// kernels_pairwise.cu -- RF / weighted RF / branch-length Euclidean over a tile of tree pairs (sm_100a).
//
// Replaces the per-pair calls getRobinsonFouldsDistance / getWeightedRobinsonFouldsDistance /
// getEuclideanDistance(t1.phylotree, t2.phylotree, normalise) of treeCl/treedist.py:70,111-114 for every pair of
// one 64x64 (or 32x32) tile of the upper triangle at once.
//
// Per pair (i,j), same leaf set:
//   internal edges : merge-join of the two trees' sorted shared-split lists (id, length) held in shared memory;
//                    splits that no other tree of the set holds are pre-summed per tree by the encoder
//                    (single_l1 / single_l2) because they can never match.
//                    common  -> |li-lj| , (li-lj)^2 ; one-sided -> |l| , l^2 ; count of common ids -> RF.
//   leaf edges     : dense fp64 "distance GEMM" over the taxon-major leaf-length matrix, 4x4 register tile per
//                    thread, k-chunks double buffered through shared memory with cp.async.
// All sums are over non-negative terms (no ||a||^2+||b||^2-2ab expansion), so identical trees give exactly 0.
#include <cuda_runtime.h>

#include <algorithm>

#include <cstdint>

#include "kernels.h"

namespace td {

namespace {

constexpr int KC = 8;   // taxa per k-chunk of the leaf-length GEMM

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <uint32_t MASK, int TILE, bool RECT>
__global__ void __launch_bounds__((TILE / 4) * (TILE / 4), (TILE == 64 ? 2 : 4))
pairwise_kernel(const PairParams p)
{
    constexpr bool kRF = MASK & 1u, kWRF = (MASK >> 1) & 1u, kEUC = (MASK >> 2) & 1u;
    constexpr bool kWeighted = kWRF || kEUC;
    constexpr int TD4 = TILE / 4, NT = TD4 * TD4;

    const int tile_c = blockIdx.x, tile_r = blockIdx.y + p.tile_row0;
    if (!RECT && tile_c < tile_r) return;
    const int64_t i0 = (int64_t)tile_r * TILE;
    const int64_t j0 = (int64_t)(tile_c + p.col_tile0) * TILE;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    SplitRec *recA = reinterpret_cast<SplitRec *>(smem_raw);
    SplitRec *recB = recA + (size_t)TILE * p.sh_stride;
    double *leaf = reinterpret_cast<double *>(recB + (size_t)TILE * p.sh_stride);   // [2 stages][A|B][KC][TILE]

    const int tid = threadIdx.x, tx = tid % TD4, ty = tid / TD4;

    // ---- stage the two tiles' split lists (contiguous in HBM: trees of a tile are adjacent rows) -------------
    {
        const int n16 = TILE * p.sh_stride;     // 16-byte records per tile
        const SplitRec *gA = p.shared + i0 * p.sh_stride, *gB = p.shared + j0 * p.sh_stride;
        for (int e = tid; e < n16; e += NT) { cp_async16(recA + e, gA + e); cp_async16(recB + e, gB + e); }
        cp_async_commit();
    }
    // ---- leaf-length k-chunks ----------------------------------------------------------------------------------
    const int n_chunks = kWeighted ? p.n_k / KC : 0;
    auto load_chunk = [&](int chunk, int stage) {
        // [KC][TILE] doubles per operand = KC*TILE/2 16-byte pieces
        constexpr int PIECES = KC * TILE / 2;
        double *sA = leaf + (size_t)stage * 2 * KC * TILE, *sB = sA + KC * TILE;
        for (int e = tid; e < PIECES; e += NT) {
            const int kk = e / (TILE / 2), c2 = e % (TILE / 2);
            const double *row = p.leafT + (size_t)(chunk * KC + kk) * p.ldT;
            cp_async16(sA + kk * TILE + 2 * c2, row + i0 + 2 * c2);
            cp_async16(sB + kk * TILE + 2 * c2, row + j0 + 2 * c2);
        }
    };
    if (kWeighted) {
        if (n_chunks > 0) load_chunk(0, 0);
        cp_async_commit();
        if (n_chunks > 1) load_chunk(1, 1);
        cp_async_commit();
        cp_async_wait<2>();
    } else {
        cp_async_wait<0>();
    }
    __syncthreads();

    // ---- merge-join of the shared-split lists: thread owns pairs (4ty+r, 4tx+c) -------------------------------
    double acc1[4][4], acc2[4][4];
    int cnt[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const SplitRec *A = recA + (size_t)(4 * ty + r) * p.sh_stride;
            const SplitRec *B = recB + (size_t)(4 * tx + c) * p.sh_stride;
            int n_common = 0;
            if constexpr (kWeighted) {
                double s1 = 0.0, s2 = 0.0;
                SplitRec a = A[0], b = B[0];
                while ((a.id & b.id) != kSentinelId) {
                    const bool adv_a = a.id <= b.id, adv_b = b.id <= a.id;
                    const double la = adv_a ? a.len : 0.0, lb = adv_b ? b.len : 0.0;
                    const double d = la - lb;                 // common: la-lb, one-sided: +-l
                    if constexpr (kWRF) s1 += fabs(d);
                    if constexpr (kEUC) s2 = fma(d, d, s2);
                    n_common += (adv_a && adv_b);
                    if (adv_a) a = *++A;
                    if (adv_b) b = *++B;
                }
                acc1[r][c] = s1; acc2[r][c] = s2;
            } else {
                uint32_t a = A[0].id, b = B[0].id;
                while ((a & b) != kSentinelId) {
                    const bool adv_a = a <= b, adv_b = b <= a;
                    n_common += (adv_a && adv_b);
                    if (adv_a) a = (++A)->id;
                    if (adv_b) b = (++B)->id;
                }
            }
            cnt[r][c] = n_common;
        }
    }

    // ---- leaf edges: sum_k |a_k - b_k| and (a_k - b_k)^2 ---------------------------------------------------------
    if constexpr (kWeighted) {
        for (int chunk = 0; chunk < n_chunks; chunk++) {
            const int stage = chunk & 1;
            cp_async_wait<1>();
            __syncthreads();
            const double *sA = leaf + (size_t)stage * 2 * KC * TILE, *sB = sA + KC * TILE;
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const double2 a01 = *reinterpret_cast<const double2 *>(sA + kk * TILE + 4 * ty);
                const double2 a23 = *reinterpret_cast<const double2 *>(sA + kk * TILE + 4 * ty + 2);
                const double2 b01 = *reinterpret_cast<const double2 *>(sB + kk * TILE + 4 * tx);
                const double2 b23 = *reinterpret_cast<const double2 *>(sB + kk * TILE + 4 * tx + 2);
                const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
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

    // ---- epilogue: per-tree aggregates, normalisation, condensed (or row-major) store ----------------------------
    int sj[4]; double l1j[4], l2j[4], sumj[4], normj[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int64_t j = j0 + 4 * tx + c;
        sj[c] = p.nsplits[j];
        if constexpr (kWRF) { l1j[c] = p.single_l1[j]; sumj[c] = p.sum_len[j]; }
        if constexpr (kEUC) { l2j[c] = p.single_l2[j]; normj[c] = p.norm[j]; }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int64_t i = i0 + 4 * ty + r;
        if (i < p.row_begin || i >= p.row_end) continue;
        const int si = p.nsplits[i];
        double l1i = 0, l2i = 0, sumi = 0, normi = 0;
        if constexpr (kWRF) { l1i = p.single_l1[i]; sumi = p.sum_len[i]; }
        if constexpr (kEUC) { l2i = p.single_l2[i]; normi = p.norm[i]; }
        const int64_t row_off = RECT ? i * p.n_cols - p.col_offset : (i * p.N - i * (i + 1) / 2 - i - 1 - p.p_base);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int64_t j = j0 + 4 * tx + c;
            if (j >= p.col_end || (RECT ? j < p.col_offset : j <= i)) continue;
            const int64_t idx = row_off + j;
            if constexpr (kRF) {
                const int den = si + sj[c];
                double v = (double)(den - 2 * cnt[r][c]);
                if (p.normalise) v = den ? v / (double)den : 0.0;
                p.out[0][idx] = v;
            }
            if constexpr (kWRF) {
                double v = acc1[r][c] + (l1i + l1j[c]);
                if (p.normalise) { const double den = sumi + sumj[c]; v = den != 0.0 ? v / den : 0.0; }
                p.out[1][idx] = v;
            }
            if constexpr (kEUC) {
                double v = sqrt(acc2[r][c] + (l2i + l2j[c]));
                if (p.normalise) { const double den = normi + normj[c]; v = den != 0.0 ? v / den : 0.0; }
                p.out[2][idx] = v;
            }
        }
    }
}

__global__ void transpose_leaf_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t N, int n_taxa,
                                      int64_t ldT)
{
    // src [N][n_taxa] tree-major  ->  dst [k][ldT] taxon-major (rows beyond n_taxa / cols beyond N stay zero)
    __shared__ double tile[32][33];
    const int64_t t0 = (int64_t)blockIdx.x * 32; const int k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t t = t0 + r; const int k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (t < N && k < n_taxa) ? src[t * n_taxa + k] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r; const int64_t t = t0 + threadIdx.x;
        if (k < n_taxa && t < N) dst[(size_t)k * ldT + t] = tile[threadIdx.x][r];
    }
}

template <uint32_t MASK, int TILE, bool RECT>
cudaError_t launch_one(const PairParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    auto k = pairwise_kernel<MASK, TILE, RECT>;
    cudaError_t e = allow_full_smem((const void *)k);
    if (e != cudaSuccess) return e;
    k<<<grid, (TILE / 4) * (TILE / 4), smem, st>>>(p);
    return cudaGetLastError();
}

template <int TILE, bool RECT>
cudaError_t launch_mask(uint32_t mask, const PairParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    switch (mask & 7u) {
        case 1: return launch_one<1, TILE, RECT>(p, grid, smem, st);
        case 2: return launch_one<2, TILE, RECT>(p, grid, smem, st);
        case 3: return launch_one<3, TILE, RECT>(p, grid, smem, st);
        case 4: return launch_one<4, TILE, RECT>(p, grid, smem, st);
        case 5: return launch_one<5, TILE, RECT>(p, grid, smem, st);
        case 6: return launch_one<6, TILE, RECT>(p, grid, smem, st);
        case 7: return launch_one<7, TILE, RECT>(p, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

size_t pairwise_smem_bytes(int tile, int sh_stride, bool weighted)
{
    return (size_t)2 * tile * sh_stride * sizeof(SplitRec) + (weighted ? (size_t)2 * 2 * KC * tile * sizeof(double) : 0);
}

int pairwise_kc() { return KC; }

cudaError_t launch_pairwise(uint32_t mask, bool rect, PairParams p, int64_t tile_rows, int64_t tile_cols, int tile,
                            cudaStream_t st)
{
    const size_t smem = pairwise_smem_bytes(tile, p.sh_stride, (mask & 6u) != 0);
    dim3 grid((unsigned)tile_cols, (unsigned)tile_rows);
    if (tile == 64) return rect ? launch_mask<64, true>(mask, p, grid, smem, st) : launch_mask<64, false>(mask, p, grid, smem, st);
    return rect ? launch_mask<32, true>(mask, p, grid, smem, st) : launch_mask<32, false>(mask, p, grid, smem, st);
}

cudaError_t launch_transpose_leaf(const double *src, double *dst, int64_t N, int n_taxa, int64_t ldT, cudaStream_t st)
{
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((n_taxa + 31) / 32)), block(32, 8);
    transpose_leaf_kernel<<<grid, block, 0, st>>>(src, dst, N, n_taxa, ldT);
    return cudaGetLastError();
}

// Rows [r0, r1) of the SQUARE symmetric matrix (zero diagonal) from the condensed vector: scipy.spatial.distance.squareform on the
// device (treeCl/collection.py:456), so that the drop-in call can copy the N x N matrix straight into the caller's memory.  32 x 32
// tiles; tiles below the diagonal read the condensed vector along its contiguous direction and turn through shared memory.
__global__ void square_rows_kernel(const double *__restrict__ cond, double *__restrict__ sq, int64_t N, int64_t r0, int64_t r1)
{
    __shared__ double tile[32][33];
    const int64_t i0 = r0 + (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x;
    if (j0 + 31 < i0) {
        // strictly below the diagonal: value (i, j) = cond[(j, i)], contiguous along i for a fixed j
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int64_t j = j0 + r, i = i0 + tx;
            tile[r][tx] = (i < r1 && j < N) ? cond[j * N - j * (j + 1) / 2 + (i - j - 1)] : 0.0;
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int64_t i = i0 + r, j = j0 + tx;
            if (i < r1 && j < N) sq[(i - r0) * N + j] = tile[tx][r];
        }
        return;
    }
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t i = i0 + r, j = j0 + tx;
        if (i >= r1 || j >= N) continue;
        double v = 0.0;
        if (j > i) v = cond[i * N - i * (i + 1) / 2 + (j - i - 1)];
        else if (j < i) v = cond[j * N - j * (j + 1) / 2 + (i - j - 1)];
        sq[(i - r0) * N + j] = v;
    }
}

// A leaf-set group's own condensed slice (local rows [a0, a1) of its n_local trees) into the condensed matrix of the whole set:
// local pair (a, b) is the pair (members[a], members[b]) of the set.
__global__ void scatter_group_kernel(const double *__restrict__ tmp, double *__restrict__ out, const int32_t *__restrict__ members, int64_t n_local,
                                     int64_t a0, int64_t a1, int64_t N, int64_t p_base)
{
    const int64_t a = a0 + blockIdx.y;
    if (a >= a1) return;
    const int64_t local_base = (a * n_local - a * (a + 1) / 2) - (a0 * n_local - a0 * (a0 + 1) / 2);
    const int64_t gi = members[a], row_base = gi * N - gi * (gi + 1) / 2 - gi - 1 - p_base;
    // four independent elements per thread and pass (one block per 1,024 columns: blocks of one element per thread were launch-bound)
    const int64_t step = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t b0 = a + 1 + (int64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; b0 < n_local; b0 += step) {
        double v[4]; int32_t m[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const int64_t b = b0 + k * blockDim.x; if (b < n_local) { v[k] = tmp[local_base + (b - a - 1)]; m[k] = members[b]; } }
#pragma unroll
        for (int k = 0; k < 4; k++) { const int64_t b = b0 + k * blockDim.x; if (b < n_local) out[row_base + m[k]] = v[k]; }
    }
}

cudaError_t launch_scatter_group(const double *tmp, double *out, const int32_t *members, int64_t n_local, int64_t a0, int64_t a1, int64_t N,
                                 int64_t p_base, cudaStream_t st)
{
    if (a1 <= a0) return cudaSuccess;
    for (int64_t r = a0; r < a1; r += 65535) {                    // (grid.y limit)
        const int64_t r1 = std::min(a1, r + 65535);
        const int64_t tmp_off = (r * n_local - r * (r + 1) / 2) - (a0 * n_local - a0 * (a0 + 1) / 2);
        const dim3 grid((unsigned)std::min<int64_t>(16, (n_local + 1023) / 1024), (unsigned)(r1 - r));
        scatter_group_kernel<<<grid, 256, 0, st>>>(tmp + tmp_off, out, members, n_local, r, r1, N, p_base);
    }
    return cudaGetLastError();
}

cudaError_t launch_square_rows(const double *cond, double *sq, int64_t N, int64_t r0, int64_t r1, cudaStream_t st)
{
    if (r1 <= r0) return cudaSuccess;
    const dim3 grid((unsigned)((N + 31) / 32), (unsigned)((r1 - r0 + 31) / 32)), block(32, 8);
    square_rows_kernel<<<grid, block, 0, st>>>(cond, sq, N, r0, r1);
    return cudaGetLastError();
}

}  // namespace td
