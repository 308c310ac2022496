// kernels_incidence.cu -- exact Robinson-Foulds through an int8 tcgen05 contraction of the split-incidence matrix (sm_100a).
//
// X in {0,1}^{N x D}: X[i][s] = 1 iff tree i holds shared split s (D = splits of the dictionary held by >= 2 trees;
// splits held by one tree can never match).  C = X * X^T counts the common splits of every pair exactly (int32
// accumulation of 0/1 products), and RF(i,j) = S_i + S_j - 2 C_ij  (getRobinsonFouldsDistance, treeCl/treedist.py:113).
// This is the dense-sharing alternative to the merge / hash-join kernels: worthwhile when D is small (clustered gene
// trees), pointless for independent random trees (D ~ N * S).  It is the only place tensor cores are used.
//
// One CTA per 128 x 128 tile of the upper triangle, 192 threads:
//   warp 0     TMA producer : cp.async.bulk.tensor.2d (SWIZZLE_128B) of the A (rows) and B (columns) k-slabs, 128 B of K each,
//                             into a 3-stage shared-memory ring; completion on `full` mbarriers
//   warp 1     MMA issuer   : one elected thread issues tcgen05.mma.cta_group::1.kind::i8 (M=128, N=128, K=32) on shared-memory
//                             descriptors, accumulating in TMEM (128 lanes x 128 columns of int32); tcgen05.commit frees the
//                             ring slot (`empty` mbarriers) and finally signals `tmem_full`
//   warps 2-5  epilogue     : tcgen05.ld 32x32b.x32 (one TMEM lane = one row per thread), RF transform, condensed store
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "kernels.h"

namespace td {

namespace {

constexpr int TILE = 128;            // rows and columns of an output tile = UMMA M and N
constexpr int BK = 128;              // bytes (= int8 elements) of K per stage = one 128-byte swizzle row
constexpr int UMMA_K = 32;           // K per tcgen05.mma for 8-bit operands
constexpr int STAGES = 3;
constexpr int NTHREADS = 192;
constexpr uint32_t STAGE_BYTES = 2 * TILE * BK;      // A + B slab

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
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct IncParams {
    const int32_t *nsplits;
    int64_t N;
    int64_t row_begin, row_end, p_base;
    int tile_row0, tiles;        // first tile row, tile columns in total
    int k_blocks;                // Dpad / BK
    int normalise, out_u16;
    void *out;
};

__global__ void __launch_bounds__(NTHREADS, 2)
rf_incidence_kernel(const __grid_constant__ CUtensorMap tmap, const IncParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // [STAGES][A 16 KB | B 16 KB]; the 128-byte swizzle wants 1024-byte aligned slabs (the launch reserves the slack)
    unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    int32_t *s_cols = reinterpret_cast<int32_t *>(tmem_slot + 2);        // S_j of the tile's 128 columns

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // upper-triangular tile list, row-major from tile row p.tile_row0
    int tile_r, tile_c;
    {
        const long long t = blockIdx.x, L0 = p.tiles - p.tile_row0;
        const double b = 2.0 * (double)L0 + 1.0;
        long long u = (long long)floor((b - sqrt(b * b - 8.0 * (double)t)) * 0.5);
        if (u < 0) u = 0;
        while (u * L0 - u * (u - 1) / 2 > t) u--;
        while ((u + 1) * L0 - (u + 1) * u / 2 <= t) u++;
        tile_r = (int)u + p.tile_row0;
        tile_c = tile_r + (int)(t - (u * L0 - u * (u - 1) / 2));
    }
    const int64_t i0 = (int64_t)tile_r * TILE, j0 = (int64_t)tile_c * TILE;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                                     // TMEM: 128 columns x 128 lanes of int32
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TILE) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 64) { const int c = threadIdx.x - 64; s_cols[c] = (j0 + c < p.N) ? p.nsplits[j0 + c] : 0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < p.k_blocks; kb++) {
                const int s = kb % STAGES; const uint32_t round = kb / STAGES;
                mbar_wait(&empty[s], (round & 1u) ^ 1u);                 // slot free (first pass returns immediately)
                mbar_expect_tx(&full[s], STAGE_BYTES);
                unsigned char *a = ring + (size_t)s * STAGE_BYTES, *b = a + TILE * BK;
                tma_load_2d(a, &tmap, kb * BK, (int)i0, &full[s]);
                tma_load_2d(b, &tmap, kb * BK, (int)j0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < p.k_blocks; kb++) {
                const int s = kb % STAGES; const uint32_t round = kb / STAGES;
                mbar_wait(&full[s], round & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned char *a = ring + (size_t)s * STAGE_BYTES, *b = a + TILE * BK;
                const uint64_t da = umma_desc(a), db = umma_desc(b);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; k++)                    // +32 bytes of K inside the swizzled row: start address += 2
                    umma_i8(tmem_base, da + (uint64_t)(k * (UMMA_K >> 4)), db + (uint64_t)(k * (UMMA_K >> 4)), (kb | k) ? 1u : 0u);
                umma_commit(&empty[s]);                                  // arrives when the MMAs above have read the slot
            }
            umma_commit(tmem_full);                                      // accumulator complete
        }
    } else {
        // ---- epilogue: this warp may touch TMEM lanes 32*(warp % 4) .. +31 -----------------------------------------------
        const int q = warp & 3;
        const int row = 32 * q + lane;
        const int64_t i = i0 + row;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool row_ok = i >= p.row_begin && i < p.row_end && i < p.N;
        const int si = row_ok ? p.nsplits[i] : 0;
        const int64_t row_off = i * p.N - i * (i + 1) / 2 - i - 1 - p.p_base;
#pragma unroll 1
        for (int c0 = 0; c0 < TILE; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
                const int64_t jb = j0 + c0;
                if (p.out_u16 && jb > i && jb + 31 < p.N) {
                    // whole run valid: pack two results per 32-bit store (the condensed offset may be odd)
                    uint16_t *o = reinterpret_cast<uint16_t *>(p.out) + (row_off + jb);
                    uint32_t r[32];
#pragma unroll
                    for (int k = 0; k < 32; k++) r[k] = (uint32_t)(si + s_cols[c0 + k] - 2 * (int)v[k]) & 0xFFFFu;
                    if ((reinterpret_cast<uintptr_t>(o) & 3u) == 0) {
#pragma unroll
                        for (int k = 0; k < 32; k += 2) reinterpret_cast<uint32_t *>(o)[k >> 1] = r[k] | (r[k + 1] << 16);
                    } else {
                        o[0] = (uint16_t)r[0];
#pragma unroll
                        for (int k = 1; k < 31; k += 2) reinterpret_cast<uint32_t *>(o + 1)[k >> 1] = r[k] | (r[k + 1] << 16);
                        o[31] = (uint16_t)r[31];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; k++) {
                        const int64_t j = jb + k;
                        if (j <= i || j >= p.N) continue;
                        const int den = si + s_cols[c0 + k];
                        const int rf = den - 2 * (int)v[k];
                        if (p.out_u16) reinterpret_cast<uint16_t *>(p.out)[row_off + j] = (uint16_t)rf;
                        else {
                            double d = (double)rf;
                            if (p.normalise) d = den ? d / (double)den : 0.0;
                            reinterpret_cast<double *>(p.out)[row_off + j] = d;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TILE) : "memory");
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

size_t incidence_smem_bytes() { return (size_t)STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + TILE * 4 + 1024; }
int incidence_tile() { return TILE; }
int incidence_bk() { return BK; }

cudaError_t launch_build_incidence(const SplitRec *shared, int sh_stride, int64_t N, int64_t ldx, int8_t *X, cudaStream_t st)
{
    const int64_t n = N * sh_stride;
    if (n <= 0) return cudaSuccess;
    build_incidence_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(shared, sh_stride, N, ldx, X);
    return cudaGetLastError();
}

// 2-D tensor map over X [rows][ldx] int8, box 128 (K bytes) x 128 (rows), 128-byte swizzle.  Returns cudaErrorNotSupported if
// the driver entry point is unavailable.
cudaError_t make_incidence_tensor_map(const int8_t *X, int64_t rows, int64_t ldx, void *tmap_out)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    const cuuint64_t dims[2] = {(cuuint64_t)ldx, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx};                       // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)TILE};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(reinterpret_cast<CUtensorMap *>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(X), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_rf_incidence(const void *tmap, const int32_t *nsplits, int64_t N, int64_t d_pad, int64_t row_begin, int64_t row_end,
                                int64_t p_base, int normalise, int out_u16, void *out, cudaStream_t st)
{
    IncParams p{};
    p.nsplits = nsplits; p.N = N; p.row_begin = row_begin; p.row_end = row_end; p.p_base = p_base;
    p.tile_row0 = (int)(row_begin / TILE); p.tiles = (int)((N + TILE - 1) / TILE);
    p.k_blocks = (int)(d_pad / BK); p.normalise = normalise; p.out_u16 = out_u16; p.out = out;
    const int64_t tile_rows = (row_end + TILE - 1) / TILE - p.tile_row0;
    const int64_t L0 = p.tiles - p.tile_row0;
    const int64_t grid = tile_rows * L0 - tile_rows * (tile_rows - 1) / 2;
    if (grid <= 0 || p.k_blocks <= 0) return cudaSuccess;
    const size_t smem = incidence_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(rf_incidence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    CUtensorMap map; memcpy(&map, tmap, sizeof map);
    rf_incidence_kernel<<<(unsigned)grid, NTHREADS, smem, st>>>(map, p);
    return cudaGetLastError();
}

}  // namespace td
