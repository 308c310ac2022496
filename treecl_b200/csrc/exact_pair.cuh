// exact_pair.cuh -- exact, purely additive recomputation of wrf / euc for one pair (nearly identical trees, where the dot-product form
// of the tile kernels cancels).  ONE summation order for every caller: the partial sums of the 32 lanes of a warp (leaf edges strided
// over the taxon-major matrix, then the shared-split lists by binary search) combined by the xor butterfly.  exact_fixup_kernel runs it
// with a warp per pair; exact_pair() is the same arithmetic done by a single thread (work list overflow, join_tile kernel), so a pair
// gives the same bits whichever way it is served - results stay byte-identical under any row sharding.
#pragma once
#include "kernels.h"

namespace td {

__device__ __forceinline__ void exact_pair_lane(const PairParams &p, int64_t i, int64_t j, int lane, double &s1, double &s2)
{
    s1 = 0.0; s2 = 0.0;
    for (int k = lane; k < p.n_k; k += 32) {
        const double d = p.leafT[(size_t)k * p.ldT + i] - p.leafT[(size_t)k * p.ldT + j];
        s1 += fabs(d); s2 = fma(d, d, s2);
    }
    const SplitRec *A = p.shared + i * p.sh_stride, *B = p.shared + j * p.sh_stride;
    const int hi0 = p.sh_stride - 1;
    for (int e = lane; e < p.sh_stride; e += 32) {
        const SplitRec a = A[e];
        if (a.id != kSentinelId) {
            int lo = 0, hi = hi0;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (B[mid].id < a.id) lo = mid + 1; else hi = mid; }
            const SplitRec b = B[lo];
            const double d = (b.id == a.id) ? a.len - b.len : a.len;       // common, or held by i only
            s1 += fabs(d); s2 = fma(d, d, s2);
        }
        const SplitRec b = B[e];
        if (b.id != kSentinelId) {
            int lo = 0, hi = hi0;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (A[mid].id < b.id) lo = mid + 1; else hi = mid; }
            if (A[lo].id != b.id) { s1 += fabs(b.len); s2 = fma(b.len, b.len, s2); }          // held by j only
        }
    }
}

// warp form: every lane calls this with its own lane index; all lanes return the full sums
__device__ __forceinline__ void exact_pair_warp(const PairParams &p, int64_t i, int64_t j, int lane, double &s1, double &s2)
{
    exact_pair_lane(p, i, j, lane, s1, s2);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, d); s2 += __shfl_xor_sync(0xffffffffu, s2, d); }
    s1 += p.single_l1[i] + p.single_l1[j];
    s2 += p.single_l2[i] + p.single_l2[j];
}

// single-thread form, same bits (x + y == y + x exactly, so after every butterfly step lanes l and l ^ d hold the same value)
static __device__ __noinline__ void exact_pair(const PairParams &p, int64_t i, int64_t j, double &s1, double &s2)
{
    double a1[32], a2[32];
    for (int l = 0; l < 32; l++) exact_pair_lane(p, i, j, l, a1[l], a2[l]);
    for (int d = 16; d > 0; d >>= 1)
        for (int l = 0; l < d; l++) { a1[l] += a1[l + d]; a2[l] += a2[l + d]; }
    s1 = a1[0] + (p.single_l1[i] + p.single_l1[j]);
    s2 = a2[0] + (p.single_l2[i] + p.single_l2[j]);
}

}  // namespace td
