// exact_pair.cuh -- exact, purely additive recomputation of wrf / euc for one pair (nearly identical trees, where the dot-product form
// of the tile kernels cancels).  ONE summation order for every caller - a sequential merge of the two shared-split lists, then the
// per-tree sums of the unshared splits, then the leaf edges - so a pair gives the same bits whether exact_fixup_kernel serves it from
// the work list (one thread per pair) or a tile kernel computes it in place because the work list is full: results stay byte-identical
// under any row sharding.  Keep this function free of arrays: a stack frame in this rarely taken path cost pair_dense_kernel 30 us per
// launch (measured: 0.141 -> 0.173 ms per step of the bench set with a 464-byte frame).
#pragma once
#include "kernels.h"

namespace td {

static __device__ __forceinline__ void exact_pair_inline(const PairParams &p, int64_t i, int64_t j, double &s1, double &s2)
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

// out-of-line form for the tile kernels (rare path; keeps their register allocation free of it)
static __device__ __noinline__ void exact_pair(const PairParams &p, int64_t i, int64_t j, double &s1, double &s2) { exact_pair_inline(p, i, j, s1, s2); }

}  // namespace td
