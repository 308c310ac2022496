// kernels_embed.cu -- classical multidimensional scaling of a distance matrix that never leaves the device.
//
// Reference (per call, NumPy on the host, after the N x N matrix has gone device -> host -> DataFrame -> ndarray):
//     double_centre(matrix)                   treeCl/distance_matrix.py:58-74    B = -1/2 J D^2 J,  J = I - 11'/n
//     eigen(dbc)                              treeCl/distance_matrix.py:285-298  np.linalg.eigh, eigenvalues descending
//     _embedding_classical_mds(m, dims)       treeCl/distance_matrix.py:301-315  coords = evecs[:, :dims] * sqrt(|vals[:dims]|)
// Here: the condensed matrix produced by the distance kernels is expanded and double-centred in HBM, and only the few leading eigenpairs
// are computed, by blocked subspace iteration with Rayleigh-Ritz acceleration: Y = B X on the fp64 tensor cores (mma.sync m8n8k4, SASS
// DMMA; the product is HBM-bound: B is read once per iteration), Gram matrices by a fixed-order two-stage reduction, the b x b algebra
// (Jacobi rotation, Cholesky) on the host.  Only the N x dims embedding travels to the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "kernels.h"

namespace td {

namespace {

constexpr int EB = 16;                 // block width of the subspace iteration (dims + oversampling, padded)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// row sums of the squared distances (square matrix, row-major, ld = n)
__global__ void rowsum_sq_kernel(const double *__restrict__ sq, int64_t n, double *__restrict__ rowsum)
{
    const int64_t i = blockIdx.x;
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) { const double d = sq[i * n + j]; s = fma(d, d, s); }
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) { if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w]; __syncthreads(); }
    if (threadIdx.x == 0) rowsum[i] = red[0];
}

// grand sum in a fixed order (one CTA)
__global__ void grand_sum_kernel(const double *__restrict__ rowsum, int64_t n, double *__restrict__ out)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += rowsum[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) { if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w]; __syncthreads(); }
    if (threadIdx.x == 0) out[0] = red[0];
}

// in place: d_ij -> -1/2 (d_ij^2 - rowmean_i - rowmean_j + grandmean)      (double_centre, treeCl/distance_matrix.py:64-73)
__global__ void double_centre_kernel(double *__restrict__ sq, int64_t n, const double *__restrict__ rowsum, const double *__restrict__ grand)
{
    const int64_t i = blockIdx.x;
    const double inv = 1.0 / (double)n, gm = grand[0] * inv * inv, rm = rowsum[i] * inv;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double d = sq[i * n + j];
        sq[i * n + j] = (d * d - (rm + rowsum[j] * inv - gm)) / -2.0;
    }
}

// Y = B X:  B [n][n] row-major symmetric, X [n_pad][EB] row-major, Y [n_pad][EB].  One warp per 8 rows of Y: A fragments straight from
// global memory (a warp load covers 8 rows x 32 contiguous bytes, every sector fully used; 8 k-steps unrolled = 256 contiguous bytes per
// row in flight), the X fragments from L1 / L2 (the X panel is 128 bytes per row, shared by all warps).
__global__ void __launch_bounds__(128) bx_kernel(const double *__restrict__ B, int64_t n, const double *__restrict__ X, double *__restrict__ Y)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8;
    if (row0 >= n) return;
    const int64_t row = min(row0 + g, n - 1);                  // (rows past the end repeat the last one; their results are not stored)
    const double *a_row = B + row * n;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int64_t n4 = n & ~(int64_t)31;
    for (int64_t k0 = 0; k0 < n4; k0 += 32) {
        double a[8], b0[8], b1[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a[u] = a_row[k0 + 4 * u + t];
            b0[u] = X[(k0 + 4 * u + t) * EB + g];
            b1[u] = X[(k0 + 4 * u + t) * EB + 8 + g];
        }
#pragma unroll
        for (int u = 0; u < 8; u++) { dmma884(acc[0][0], acc[0][1], a[u], b0[u]); dmma884(acc[1][0], acc[1][1], a[u], b1[u]); }
    }
    for (int64_t k0 = n4; k0 < n; k0 += 4) {
        const int64_t k = k0 + t;
        const double a = k < n ? a_row[k] : 0.0;
        const double b0 = k < n ? X[k * EB + g] : 0.0, b1 = k < n ? X[k * EB + 8 + g] : 0.0;
        dmma884(acc[0][0], acc[0][1], a, b0); dmma884(acc[1][0], acc[1][1], a, b1);
    }
    if (row0 + g < n) {
        double *y = Y + (row0 + g) * EB;
        *reinterpret_cast<double2 *>(y + 2 * t) = make_double2(acc[0][0], acc[0][1]);
        *reinterpret_cast<double2 *>(y + 8 + 2 * t) = make_double2(acc[1][0], acc[1][1]);
    }
}

// partial Gram matrices of a slab of rows:  H = X' Y,  G = Y' Y  (EB x EB each), fixed summation order
constexpr int GRAM_ROWS = 256;
__global__ void __launch_bounds__(256) gram_partial_kernel(const double *__restrict__ X, const double *__restrict__ Y, int64_t n, double *__restrict__ part)
{
    __shared__ double sx[32][EB + 1], sy[32][EB + 1];
    const int p = threadIdx.x / EB, q = threadIdx.x % EB;          // thread (p, q) owns H[p][q] and G[p][q]
    const int64_t r0 = (int64_t)blockIdx.x * GRAM_ROWS, r1 = min(n, r0 + GRAM_ROWS);
    double h = 0.0, gsum = 0.0;
    for (int64_t rb = r0; rb < r1; rb += 32) {
        for (int e = threadIdx.x; e < 32 * EB; e += 256) {
            const int64_t r = rb + e / EB;
            sx[e / EB][e % EB] = r < r1 ? X[r * EB + e % EB] : 0.0;
            sy[e / EB][e % EB] = r < r1 ? Y[r * EB + e % EB] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; r++) { h = fma(sx[r][p], sy[r][q], h); gsum = fma(sy[r][p], sy[r][q], gsum); }
        __syncthreads();
    }
    part[((size_t)blockIdx.x * 2 + 0) * EB * EB + threadIdx.x] = h;
    part[((size_t)blockIdx.x * 2 + 1) * EB * EB + threadIdx.x] = gsum;
}

__global__ void __launch_bounds__(256) gram_final_kernel(const double *__restrict__ part, int n_part, double *__restrict__ out)
{
    for (int which = 0; which < 2; which++) {
        double s = 0.0;
        for (int b = 0; b < n_part; b++) s += part[((size_t)b * 2 + which) * EB * EB + threadIdx.x];
        out[which * EB * EB + threadIdx.x] = s;
    }
}

// X = Y T   (T: EB x EB, row-major, in constant-size global memory)
__global__ void __launch_bounds__(256) transform_kernel(const double *__restrict__ Y, const double *__restrict__ T, int64_t n, double *__restrict__ X)
{
    __shared__ double sT[EB][EB];
    if (threadIdx.x < EB * EB) sT[threadIdx.x / EB][threadIdx.x % EB] = T[threadIdx.x];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * (256 / EB) + threadIdx.x / EB;
    const int q = threadIdx.x % EB;
    if (r >= n) return;
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < EB; p++) s = fma(Y[r * EB + p], sT[p][q], s);
    X[r * EB + q] = s;
}

// W = Y - X H   (H: EB x EB): the block residual of the Rayleigh-Ritz step, formed explicitly - its Gram matrix S'(W'W)S gives the
// residual norms of the Ritz pairs without the cancellation of (Y'Y) - H^2
__global__ void __launch_bounds__(256) block_residual_kernel(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Hm, int64_t n,
                                                           double *__restrict__ Wm)
{
    __shared__ double sH[EB][EB];
    if (threadIdx.x < EB * EB) sH[threadIdx.x / EB][threadIdx.x % EB] = Hm[threadIdx.x];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * (256 / EB) + threadIdx.x / EB;
    const int q = threadIdx.x % EB;
    if (r >= n) return;
    double s = Y[r * EB + q];
#pragma unroll
    for (int p = 0; p < EB; p++) s = fma(-X[r * EB + p], sH[p][q], s);
    Wm[r * EB + q] = s;
}

// deterministic start block: a counter-based hash -> uniform(-1, 1)
__global__ void init_block_kernel(double *__restrict__ X, int64_t n)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * EB) return;
    uint64_t z = (uint64_t)e * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    X[e] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// coords[i][j] = V[i][j] * sqrt(|lambda_j|),  j < dims
__global__ void coords_kernel(const double *__restrict__ V, const double *__restrict__ scale, int64_t n, int dims, double *__restrict__ coords)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * dims) return;
    const int64_t i = e / dims; const int j = (int)(e % dims);
    coords[e] = V[i * EB + j] * scale[j];
}

// rows of V (first `dims` columns) scaled to unit length; zero rows stay (normalise_rows, treeCl/distance_matrix.py:141-150)
__global__ void unit_rows_kernel(const double *__restrict__ V, int64_t n, int dims, double *__restrict__ coords)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int j = 0; j < dims; j++) s = fma(V[i * EB + j], V[i * EB + j], s);
    const double len = s > 0.0 ? sqrt(s) : 1.0;
    for (int j = 0; j < dims; j++) coords[i * dims + j] = V[i * EB + j] / len;
}

__global__ void add_diagonal_kernel(double *__restrict__ sq, int64_t n, double v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sq[i * n + i] += v;
}

// ---- order statistics of the rows of the distance matrix: kdists (the k-th nearest distance, treeCl/distance_matrix.py:153-164: the
// diagonal zero is rank 0) and the row medians of Spectral.decompose (treeCl/clustering.py:205-207).  One CTA per row; an 8-pass
// most-significant-byte radix select on the order-preserving 64-bit image of the doubles: no sort, the row is read 8 times from L2.
__device__ __forceinline__ unsigned long long order_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}
// out[i] = mean of the rank_a-th and rank_b-th smallest entries of row i (rank_a == rank_b: one order statistic)
__global__ void __launch_bounds__(256) row_select_kernel(const double *__restrict__ sq, int64_t n, int64_t rank_a, int64_t rank_b, double *__restrict__ out)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ long long s_rank;
    const double *row = sq + (int64_t)blockIdx.x * n;
    double acc = 0.0;
    for (int which = 0; which < 2; which++) {
        const int64_t want = which ? rank_b : rank_a;
        if (which && rank_b == rank_a) break;
        if (threadIdx.x == 0) { s_prefix = 0ull; s_rank = want; }
        __syncthreads();
        for (int pass = 0; pass < 8; pass++) {
            const int shift = 56 - 8 * pass;
            hist[threadIdx.x] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            for (int64_t j = threadIdx.x; j < n; j += 256) {
                const unsigned long long k = order_key(row[j]);
                if (pass == 0 || (k >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                long long r = s_rank; int b = 0;
                while (b < 255 && r >= (long long)hist[b]) { r -= hist[b]; b++; }
                s_rank = r; s_prefix = prefix | ((unsigned long long)b << shift);
            }
            __syncthreads();
        }
        acc += key_value(s_prefix);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = rank_a == rank_b ? acc : 0.5 * acc;
}

// affinity(matrix, mask, scale) of treeCl/distance_matrix.py:32-55 with the mask and the scale the Spectral manager builds
// (treeCl/clustering.py:171-228): scale_ij = s_i s_j (s = row medians or k-th nearest distances), mask_ij = d_ij <= kd_i, combined with its
// transpose by OR / AND (kmask, :167-180), NaN (0 / 0) -> exp(-1), masked entries 0, diagonal `diag` (decompose sets 1, spectral_embedding_ 0).
// In place on the square matrix.
__global__ void affinity_kernel(double *__restrict__ sq, int64_t n, const double *__restrict__ s, const double *__restrict__ kd, int logic_and, double diag)
{
    const int64_t i = blockIdx.x;
    const double si = s[i], kdi = kd ? kd[i] : 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double d = sq[i * n + j];
        double v = -(d * d) / (si * s[j]);
        if (v != v) v = -1.0;
        v = exp(v);
        if (kd) {
            // the matrix is symmetric: d_ji = d_ij
            const bool mij = d <= kdi, mji = d <= kd[j];
            if (!(logic_and ? (mij && mji) : (mij || mji))) v = 0.0;
        }
        if (i == j) v = diag;
        sq[i * n + j] = v;
    }
}

// laplace(affinity) of treeCl/distance_matrix.py:189-207, default form: L = D^-1/2 A D^-1/2 with D = row sums without the diagonal
// (degrees <= 1e-10 count as 1).  Two kernels: degrees, then the scaling in place.
__global__ void degree_kernel(const double *__restrict__ a, int64_t n, double *__restrict__ deg)
{
    const int64_t i = blockIdx.x;
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) if (j != i) s += a[i * n + j];
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) { if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w]; __syncthreads(); }
    if (threadIdx.x == 0) { const double d = red[0]; deg[i] = sqrt(d <= 1e-10 ? 1.0 : d); }
}
__global__ void laplace_kernel(double *__restrict__ a, int64_t n, const double *__restrict__ sqrt_deg)
{
    const int64_t i = blockIdx.x;
    const double di = sqrt_deg[i];
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) a[i * n + j] = a[i * n + j] / sqrt_deg[j] / di;
}

// ---- host: EB x EB symmetric algebra -------------------------------------------------------------------------------------------------
// cyclic Jacobi: A = V diag(w) V'; eigenvalues sorted descending
void jacobi_eigen(int b, std::vector<double> A, std::vector<double> &w, std::vector<double> &V)
{
    V.assign((size_t)b * b, 0.0);
    for (int i = 0; i < b; i++) V[(size_t)i * b + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < b; p++) { diag += A[(size_t)p * b + p] * A[(size_t)p * b + p]; for (int q = p + 1; q < b; q++) off += A[(size_t)p * b + q] * A[(size_t)p * b + q]; }
        if (off <= 1e-60 + 1e-34 * diag) break;
        for (int p = 0; p < b; p++)
            for (int q = p + 1; q < b; q++) {
                const double apq = A[(size_t)p * b + q];
                if (apq == 0.0) continue;
                const double theta = (A[(size_t)q * b + q] - A[(size_t)p * b + p]) / (2.0 * apq);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
                for (int k = 0; k < b; k++) {
                    const double akp = A[(size_t)k * b + p], akq = A[(size_t)k * b + q];
                    A[(size_t)k * b + p] = c * akp - s * akq; A[(size_t)k * b + q] = s * akp + c * akq;
                }
                for (int k = 0; k < b; k++) {
                    const double apk = A[(size_t)p * b + k], aqk = A[(size_t)q * b + k];
                    A[(size_t)p * b + k] = c * apk - s * aqk; A[(size_t)q * b + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < b; k++) {
                    const double vkp = V[(size_t)k * b + p], vkq = V[(size_t)k * b + q];
                    V[(size_t)k * b + p] = c * vkp - s * vkq; V[(size_t)k * b + q] = s * vkp + c * vkq;
                }
            }
    }
    std::vector<int> order(b);
    for (int i = 0; i < b; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return A[(size_t)x * b + x] > A[(size_t)y * b + y]; });
    w.resize(b);
    std::vector<double> Vs((size_t)b * b);
    for (int j = 0; j < b; j++) { w[j] = A[(size_t)order[j] * b + order[j]]; for (int k = 0; k < b; k++) Vs[(size_t)k * b + j] = V[(size_t)k * b + order[j]]; }
    V.swap(Vs);
}

}  // namespace

// workspace (doubles) the caller provides besides the n x n matrix
size_t cmds_workspace_doubles(int64_t n)
{
    const size_t n_part = (size_t)((n + GRAM_ROWS - 1) / GRAM_ROWS);
    return (((size_t)n + 1) & ~(size_t)1) + 8 + 3 * (size_t)n * EB + n_part * 2 * EB * EB + 2 * EB * EB + EB * EB + EB;
}
int cmds_block() { return EB; }

// Leading eigenpairs of the symmetric matrix `sq` (n x n on the device), largest algebraic first, by blocked subspace iteration with
// Rayleigh-Ritz.  `shift` has already been added to the diagonal by the caller (so that the wanted eigenvalues are the largest in
// magnitude) and is taken off the reported eigenvalues.  mode 0: coords = V sqrt(|lambda|) (classical MDS); mode 1: rows of V scaled to
// unit length (normalise_rows, treeCl/distance_matrix.py:141-150).  work: cmds_workspace_doubles(n) doubles, its first n + 8 free for
// the caller; h_pinned: 3 * EB * EB doubles of page-locked host memory.  Synchronises `st` once per iteration.
static cudaError_t run_eigs(double *sq, int64_t n, int dims, double shift, int mode, double *work, double *h_pinned, double *d_coords,
                            double *eigenvalues, int max_iter, double tol, int *iterations, double *residual, cudaStream_t st)
{
    const int b = EB;
    const int n_part = (int)((n + GRAM_ROWS - 1) / GRAM_ROWS);
    // (X, Y: 16-byte aligned rows for the double2 stores)
    double *rowsum = work, *grand = rowsum + ((n + 1) & ~(int64_t)1), *X = grand + 8, *Y = X + (size_t)n * EB, *Wm = Y + (size_t)n * EB, *part = Wm + (size_t)n * EB;
    double *gram = part + (size_t)n_part * 2 * EB * EB, *T = gram + 2 * EB * EB, *scale = T + EB * EB;
    cudaError_t e;
    (void)rowsum;
    init_block_kernel<<<(unsigned)((n * EB + 255) / 256), 256, 0, st>>>(X, n);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    launch_counter_add(1);

    std::vector<double> H((size_t)b * b), G((size_t)b * b), w, S, prev(b, 0.0), Tm((size_t)b * b);
    const unsigned bx_grid = (unsigned)((n + 31) / 32);
    bool have_basis = false;                       // X orthonormal?
    int it = 0;
    double res = 1.0;
    for (it = 0; it < max_iter + 1; it++) {
        if (!have_basis) {
            // orthonormalise the start block: G = X'X by the same kernels (Y := X)
            gram_partial_kernel<<<n_part, 256, 0, st>>>(X, X, n, part);
            gram_final_kernel<<<1, 256, 0, st>>>(part, n_part, gram);
            launch_counter_add(2);
        } else {
            bx_kernel<<<bx_grid, 128, 0, st>>>(sq, n, X, Y);
            gram_partial_kernel<<<n_part, 256, 0, st>>>(X, Y, n, part);
            gram_final_kernel<<<1, 256, 0, st>>>(part, n_part, gram);
            launch_counter_add(3);
        }
        if ((e = cudaMemcpyAsync(h_pinned, gram, 2 * EB * EB * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        for (int k = 0; k < b * b; k++) { H[k] = h_pinned[k]; G[k] = h_pinned[EB * EB + k]; }
        // Ritz rotation S (identity for the start block), then T = S R^-1 with R'R = S' G S (Cholesky): X_new = Y T is orthonormal
        std::vector<double> M((size_t)b * b, 0.0), GS((size_t)b * b, 0.0);
        if (have_basis) {
            for (int p = 0; p < b; p++) for (int q = p + 1; q < b; q++) { const double m = 0.5 * (H[(size_t)p * b + q] + H[(size_t)q * b + p]); H[(size_t)p * b + q] = H[(size_t)q * b + p] = m; }
            jacobi_eigen(b, H, w, S);
        } else {
            S.assign((size_t)b * b, 0.0);
            for (int i = 0; i < b; i++) S[(size_t)i * b + i] = 1.0;
        }
        for (int i = 0; i < b; i++) for (int j = 0; j < b; j++) { double s = 0; for (int k = 0; k < b; k++) s += G[(size_t)i * b + k] * S[(size_t)k * b + j]; GS[(size_t)i * b + j] = s; }
        for (int i = 0; i < b; i++) for (int j = 0; j < b; j++) { double s = 0; for (int k = 0; k < b; k++) s += S[(size_t)k * b + i] * GS[(size_t)k * b + j]; M[(size_t)i * b + j] = s; }
        if (have_basis) {
            // Stop on the residuals || B v_j - theta_j v_j || of the wanted Ritz pairs (eigenvector error ~ residual / gap), not on the
            // change of the eigenvalues from one product to the next, which is small whenever convergence is slow.  B v - theta v =
            // (Y - X H) s, formed explicitly on the device ((Y'Y) - H^2 would cancel to sqrt(eps)); looked at once the eigenvalues have
            // settled to 1e-6, every other product.
            double change = 0.0, wmax = 0.0;
            for (int j = 0; j < b; j++) wmax = std::max(wmax, std::fabs(w[j]));
            for (int j = 0; j < dims; j++) change = std::max(change, std::fabs(w[j] - prev[j]));
            prev = w;
            if (it == max_iter) break;
            if (change <= 1e-6 * wmax && (it & 1) == 0) {
                memcpy(h_pinned + 2 * EB * EB, H.data(), sizeof(double) * b * b);
                if ((e = cudaMemcpyAsync(T, h_pinned + 2 * EB * EB, EB * EB * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
                block_residual_kernel<<<(unsigned)((n + 256 / EB - 1) / (256 / EB)), 256, 0, st>>>(X, Y, T, n, Wm);
                gram_partial_kernel<<<n_part, 256, 0, st>>>(Wm, Wm, n, part);
                gram_final_kernel<<<1, 256, 0, st>>>(part, n_part, gram);
                launch_counter_add(3);
                if ((e = cudaMemcpyAsync(h_pinned, gram, EB * EB * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
                if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
                double rmax = 0.0;
                for (int j = 0; j < dims; j++) {
                    double q = 0.0;
                    for (int a = 0; a < b; a++) { double t = 0.0; for (int c = 0; c < b; c++) t += h_pinned[a * b + c] * S[(size_t)c * b + j]; q += S[(size_t)a * b + j] * t; }
                    rmax = std::max(rmax, std::sqrt(std::max(q, 0.0)));
                }
                res = wmax > 0.0 ? rmax / wmax : 0.0;
                if (res <= tol) break;
            }
        }
        // columns of Y S differ in length by the ratios of the Ritz values: scale M to a unit diagonal first (c_j), then Cholesky
        // M' = R'R (upper R); a vanishing pivot (rank-deficient block: fewer non-zero eigenvalues than columns) keeps its column as it is
        std::vector<double> R((size_t)b * b, 0.0), Rinv((size_t)b * b, 0.0), cs(b, 1.0);
        for (int i = 0; i < b; i++) cs[i] = M[(size_t)i * b + i] > 0.0 ? 1.0 / std::sqrt(M[(size_t)i * b + i]) : 1.0;
        for (int i = 0; i < b; i++) for (int j = 0; j < b; j++) M[(size_t)i * b + j] *= cs[i] * cs[j];
        for (int j = 0; j < b; j++) {
            double d = M[(size_t)j * b + j];
            for (int k = 0; k < j; k++) d -= R[(size_t)k * b + j] * R[(size_t)k * b + j];
            d = d > 1e-24 ? std::sqrt(d) : 1.0;                                          // (degenerate direction: left un-normalised)
            R[(size_t)j * b + j] = d;
            for (int i = j + 1; i < b; i++) {
                double s = M[(size_t)j * b + i];
                for (int k = 0; k < j; k++) s -= R[(size_t)k * b + j] * R[(size_t)k * b + i];
                R[(size_t)j * b + i] = s / d;
            }
        }
        for (int j = 0; j < b; j++) {                      // R^-1 (upper triangular) by back substitution
            Rinv[(size_t)j * b + j] = 1.0 / R[(size_t)j * b + j];
            for (int i = j - 1; i >= 0; i--) {
                double s = 0; for (int k = i + 1; k <= j; k++) s += R[(size_t)i * b + k] * Rinv[(size_t)k * b + j];
                Rinv[(size_t)i * b + j] = -s / R[(size_t)i * b + i];
            }
        }
        for (int i = 0; i < b; i++) for (int j = 0; j < b; j++) { double s = 0; for (int k = 0; k < b; k++) s += S[(size_t)i * b + k] * cs[k] * Rinv[(size_t)k * b + j]; Tm[(size_t)i * b + j] = s; }
        memcpy(h_pinned + 2 * EB * EB, Tm.data(), sizeof(double) * b * b);
        if ((e = cudaMemcpyAsync(T, h_pinned + 2 * EB * EB, EB * EB * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        // start block: X := X T (through Y as scratch); afterwards X := Y T
        if (!have_basis) {
            if ((e = cudaMemcpyAsync(Y, X, (size_t)n * EB * sizeof(double), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
        }
        transform_kernel<<<(unsigned)((n + 256 / EB - 1) / (256 / EB)), 256, 0, st>>>(Y, T, n, X);
        launch_counter_add(1);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;       // h_pinned is reused next iteration
        if (!have_basis) { have_basis = true; it--; }                      // (the orthonormalisation pass is not an iteration)
    }
    // Ritz vectors of the last product: V = X S (X is the orthonormal block that produced H); eigenvalues w
    memcpy(h_pinned + 2 * EB * EB, S.data(), sizeof(double) * b * b);
    if ((e = cudaMemcpyAsync(T, h_pinned + 2 * EB * EB, EB * EB * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    transform_kernel<<<(unsigned)((n + 256 / EB - 1) / (256 / EB)), 256, 0, st>>>(X, T, n, Y);
    for (int j = 0; j < dims; j++) w[j] -= shift;
    for (int j = 0; j < EB; j++) h_pinned[j] = j < dims ? (mode == 0 ? std::sqrt(std::fabs(w[j])) : 1.0) : 0.0;
    if ((e = cudaMemcpyAsync(scale, h_pinned, EB * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if (mode == 0) coords_kernel<<<(unsigned)((n * dims + 255) / 256), 256, 0, st>>>(Y, scale, n, dims, d_coords);
    else unit_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Y, n, dims, d_coords);
    launch_counter_add(2);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    for (int j = 0; j < dims; j++) eigenvalues[j] = w[j];
    if (iterations) *iterations = it;
    if (residual) *residual = res;
    return cudaGetLastError();
}

// classical MDS: sq = square distance matrix, overwritten by the double-centred matrix
cudaError_t run_cmds(double *sq, int64_t n, int dims, double *work, double *h_pinned, double *d_coords, double *eigenvalues, int max_iter, double tol,
                     int *iterations, double *residual, cudaStream_t st)
{
    double *rowsum = work, *grand = rowsum + ((n + 1) & ~(int64_t)1);
    rowsum_sq_kernel<<<(unsigned)n, 256, 0, st>>>(sq, n, rowsum);
    grand_sum_kernel<<<1, 256, 0, st>>>(rowsum, n, grand);
    double_centre_kernel<<<(unsigned)n, 256, 0, st>>>(sq, n, rowsum, grand);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    launch_counter_add(3);
    return run_eigs(sq, n, dims, 0.0, 0, work, h_pinned, d_coords, eigenvalues, max_iter, tol, iterations, residual, st);
}

// affinity matrix of the Spectral manager in place of the square distance matrix `sq` (scale_k 0: row medians, else kscale(k);
// prune_k 0: no pruning, else kmask(k) with OR / AND); diagonal = `diag`.  vec: 2 n doubles of device scratch.
cudaError_t run_affinity(double *sq, int64_t n, int prune_k, int prune_and, int scale_k, double diag, double *vec, cudaStream_t st)
{
    double *s = vec, *kd = vec + n;
    if (scale_k > 0) row_select_kernel<<<(unsigned)n, 256, 0, st>>>(sq, n, scale_k, scale_k, s);
    else row_select_kernel<<<(unsigned)n, 256, 0, st>>>(sq, n, (n - 1) / 2, n / 2, s);                  // np.median(matrix, axis=1)
    if (prune_k > 0) row_select_kernel<<<(unsigned)n, 256, 0, st>>>(sq, n, prune_k, prune_k, kd);
    affinity_kernel<<<(unsigned)n, 256, 0, st>>>(sq, n, s, prune_k > 0 ? kd : nullptr, prune_and, diag);
    launch_counter_add(prune_k > 0 ? 3 : 2);
    return cudaGetLastError();
}

// Spectral(...).spectral_embedding_(dims) of treeCl/clustering.py:288-303: affinity with a zero diagonal -> laplace -> eigen -> the first
// `dims` eigenvectors with unit-length rows.  sq: square distance matrix, overwritten.
cudaError_t run_spectral(double *sq, int64_t n, int dims, int prune_k, int prune_and, int scale_k, double *work, double *h_pinned, double *d_coords,
                         double *eigenvalues, int max_iter, double tol, int *iterations, double *residual, cudaStream_t st)
{
    double *X = work + ((n + 1) & ~(int64_t)1) + 8;          // (the eigen-solver's blocks are free until it starts: 2 n doubles of scratch)
    cudaError_t e = run_affinity(sq, n, prune_k, prune_and, scale_k, 0.0, X, st);
    if (e != cudaSuccess) return e;
    degree_kernel<<<(unsigned)n, 256, 0, st>>>(sq, n, X);
    laplace_kernel<<<(unsigned)n, 256, 0, st>>>(sq, n, X);
    // the spectrum of D^-1/2 A D^-1/2 lies in [-1, 1]: + I makes the largest algebraic eigenvalues the largest in magnitude
    add_diagonal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sq, n, 1.0);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    launch_counter_add(3);
    return run_eigs(sq, n, dims, 1.0, 1, work, h_pinned, d_coords, eigenvalues, max_iter, tol, iterations, residual, st);
}

}  // namespace td
