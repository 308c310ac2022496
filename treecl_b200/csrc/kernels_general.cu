// kernels_general.cu -- all four metrics for pairs of trees with DIFFERENT leaf sets: restrict-then-join, one warp per pair.
//
// Restates on the split level what the reference does on the tree level for such a pair (treeCl/treedist.py:63-68):
//     if t1 ^ t2:  if len(t1 & t2) < min_overlap: return overlap_fail_value
//                  else: t1, t2 = prune both to the intersection          (Tree.prune_to_subset, treeCl/tree.py:989-1001)
// Pruning a tree to a leaf subset M and re-encoding it (dendropy retain_taxa + suppress_unifurcations +
// collapse_basal_bifurcation, then PhyloTree) is, in split terms: intersect every split mask with M; drop splits with an
// empty side; a split with a one-leaf side becomes part of that leaf's edge (lengths add); re-canonicalise against the
// lowest taxon of M; equal masks merge and their lengths ADD (pinned by the reference's partial-overlap golden,
// tests/test.py:280-286).  After that the pair is an ordinary same-leaf-set pair; the GTP part is geo_core.cuh.
// Used for every pair of a non-uniform tree set (slow path: a few thousand pairs/s per SM, but on the GPU).
#include <cuda_runtime.h>

#include <cstdint>

#include "geo_core.cuh"
#include "kernels.h"

namespace td {

namespace {

using namespace geo;

template <int H, int W>
struct GeneralSmem {
    WarpSmem<H, W> core;                     // flow, lenB, maskB, idA/idB (unused here), stack
    uint64_t maskA[32 * H * W];
    double lenA[32 * H];
    int table[64 * H];                       // open-addressing table of slot indices keyed by restricted mask (load <= 1/2)
    int partnerA[32 * H];                    // per slot of tree A: slot of the equal split of tree B, or -1
    double leafA[64 * W], leafB[64 * W];     // restricted leaf-edge lengths (taxa outside the intersection stay 0)
};

template <int W> __device__ __forceinline__ bool mask_eq(const uint64_t *a, const uint64_t (&b)[W])
{
    bool e = true;
#pragma unroll
    for (int w = 0; w < W; w++) e &= (a[w] == b[w]);
    return e;
}

template <int W> __device__ __forceinline__ uint32_t mask_hash(const uint64_t (&m)[W])
{
    uint64_t x = m[0];
#pragma unroll
    for (int w = 1; w < W; w++) x ^= m[w] * (0x9E3779B97F4A7C15ull + 2 * w);
    x *= 0xD6E8FEB86659FD93ull;
    return (uint32_t)(x >> 40);
}

// restrict my split(s) of one tree to the common leaf set M; leaf folds go to `leaf`, internal splits (deduplicated, lengths
// summed in ascending slot order) stay lane-owned: mine[h] (valid representative), m[h], len[h].  Returns the valid set.
template <int H, int W>
__device__ __forceinline__ Bits<H> restrict_tree(const GeoParams &p, int64_t t, const uint64_t (&M)[W], int nM, int ref, int rooted,
                                                 uint64_t *smask, double *slen, double *leaf, int *table, int lane,
                                                 uint64_t (&m)[H][W], double (&len)[H])
{
    constexpr int T = 64 * H;
    const int S = p.nsplits[t];
    for (int k = lane; k < T; k += 32) table[k] = -1;
    Bits<H> valid;
#pragma unroll
    for (int h = 0; h < H; h++) {
        const int s = lane + 32 * h;
        bool internal = false; int leaf_k = -1; len[h] = 0.0;
        if (s < S) {
            len[h] = p.lens[t * p.stride + s];
            int c = 0;
#pragma unroll
            for (int w = 0; w < W; w++) { m[h][w] = p.masks[(t * p.stride + s) * W + w] & M[w]; c += __popcll(m[h][w]); }
            const int cc = nM - c;
            if (c == 0 || cc == 0) { /* edge inside a pruned region, or the stem above the new root */ }
            else if (c == 1 || (!rooted && cc == 1)) {
                if (c != 1) {
#pragma unroll
                    for (int w = 0; w < W; w++) m[h][w] = M[w] & ~m[h][w];
                }
#pragma unroll
                for (int w = 0; w < W; w++) if (m[h][w]) leaf_k = 64 * w + __ffsll((long long)m[h][w]) - 1;
            } else {
                internal = true;
                if (!rooted && ((m[h][ref >> 6] >> (ref & 63)) & 1ull)) {
#pragma unroll
                    for (int w = 0; w < W; w++) m[h][w] = M[w] & ~m[h][w];
                }
            }
        }
        // leaf folds, in slot order (one writer at a time -> deterministic sums)
        uint32_t folds = __ballot_sync(FULL, leaf_k >= 0);
        while (folds) {
            const int src = __ffs(folds) - 1; folds &= folds - 1;
            const int k = __shfl_sync(FULL, leaf_k, src); const double l = __shfl_sync(FULL, len[h], src);
            if (lane == 0) leaf[k] += l;
        }
        __syncwarp();
#pragma unroll
        for (int w = 0; w < W; w++) smask[s * W + w] = internal ? m[h][w] : 0ull;
        slen[s] = len[h];
        valid.w[h] = __ballot_sync(FULL, internal);
    }
    __syncwarp();
    // merge equal masks: the lowest slot is the representative and takes the sum, in ascending slot order.  Every valid slot enters the
    // hash table (on an equal mask the entry keeps the lowest slot); the table is left holding the representatives for the caller.
    // (Round 2 history, 5,000 x 50 with 1 % odd trees, rf+euc: a walk over the valid-set bits per lane 5.6 ms, a linear scan 1.9 ms.)
    // (the probe loops are warp-uniform - one step per pass for every lane still looking - so that the warp stays converged: with a
    // per-lane `for (;;)` the lanes left their loops at different times and ran the rest of the pair in up to five separate groups)
    int entry[H];
#pragma unroll
    for (int h = 0; h < H; h++) {
        const int s = lane + 32 * h;
        uint32_t e = mask_hash<W>(m[h]) & (T - 1);
        bool looking = valid.mine(h, lane);
        while (__any_sync(FULL, looking)) {
            if (looking) {
                const int prev = atomicCAS(&table[e], -1, s);
                if (prev == -1) looking = false;
                else if (mask_eq<W>(&smask[prev * W], m[h])) { atomicMin(&table[e], s); looking = false; }
                else e = (e + 1) & (T - 1);
            }
        }
        entry[h] = (int)e;
    }
    __syncwarp();
    Bits<H> reps;
    int my_rep[H];
#pragma unroll
    for (int h = 0; h < H; h++) {
        my_rep[h] = valid.mine(h, lane) ? table[entry[h]] : -1;
        reps.w[h] = __ballot_sync(FULL, my_rep[h] == lane + 32 * h);
    }
#pragma unroll
    for (int h2 = 0; h2 < H; h2++) {                    // merged slots hand their length to the representative, in ascending slot order
        uint32_t dups = valid.w[h2] & ~reps.w[h2];
        while (dups) {
            const int src = __ffs(dups) - 1; dups &= dups - 1;
            const int r = __shfl_sync(FULL, my_rep[h2], src); const double l = __shfl_sync(FULL, len[h2], src);
#pragma unroll
            for (int h = 0; h < H; h++) if (lane + 32 * h == r) len[h] += l;
        }
    }
#pragma unroll
    for (int h = 0; h < H; h++) if (!reps.mine(h, lane)) len[h] = 0.0;
    __syncwarp();
    // only representatives keep their mask
#pragma unroll
    for (int h = 0; h < H; h++) if (!reps.mine(h, lane)) {
#pragma unroll
        for (int w = 0; w < W; w++) smask[(lane + 32 * h) * W + w] = 0ull;
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < H; h++) slen[lane + 32 * h] = len[h];
    __syncwarp();
    return reps;
}

#ifndef GEN_MINB
#define GEN_MINB 8
#endif
// GEO: the launch wants the geodesic.  Without it the kernel needs neither the GTP code nor its registers (128 -> the 64 of the cap
// below: twice the warps on an SM for a latency-bound kernel).
template <int H, int W, bool GEO>
__global__ void __launch_bounds__(GEO_WARPS * 32, GEO ? 1 : GEN_MINB) general_kernel(const GeoParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int FS = p.flow_rows | 1;
    const size_t flow_doubles = GEO ? (size_t)p.flow_rows * FS : 0;
    const size_t warp_bytes = sizeof(GeneralSmem<H, W>) + (p.flow_scratch ? 0 : flow_doubles * sizeof(double));
    unsigned char *wbase = smem_raw + (size_t)(threadIdx.x >> 5) * warp_bytes;
    GeneralSmem<H, W> &sm = *reinterpret_cast<GeneralSmem<H, W> *>(wbase);
    double *flow = p.flow_scratch ? p.flow_scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * flow_doubles
                                  : reinterpret_cast<double *>(wbase + sizeof(GeneralSmem<H, W>));
    const int lane = threadIdx.x & 31;
    Counters st;
    bool failed = false;
    const double kNaN = __longlong_as_double(0x7ff8000000000000ll);

    PairFeed feed;
    for (;;) {
        int64_t pidx, ti, tj;
        if (!next_pair(p, feed, lane, pidx, ti, tj)) break;
        st.pairs++;
        const int64_t o = pidx - p.p_base;
        auto write_all = [&](double v) {
            if (lane == 0) for (int m = 0; m < 4; m++) if ((p.metric_mask >> m & 1u) && p.outs[m]) p.outs[m][o] = v;
        };

        // ---- common leaf set -----------------------------------------------------------------------------------------
        uint64_t M[W]; bool same = true; int nM = 0, ref = -1;
#pragma unroll
        for (int w = 0; w < W; w++) {
            const uint64_t a = p.leafmask[ti * W + w], b = p.leafmask[tj * W + w];
            same &= (a == b); M[w] = a & b; nM += __popcll(M[w]);
        }
#pragma unroll
        for (int w = W - 1; w >= 0; w--) if (M[w]) ref = 64 * w + __ffsll((long long)M[w]) - 1;
        const int rooted = p.rooted[ti];
        if (rooted != p.rooted[tj]) { write_all(kNaN); if (lane == 0) atomicOr(p.error_flag, 2); continue; }
        if (!same && nM < p.min_overlap) { write_all(p.fail_value); continue; }       // treeCl/treedist.py:64-65

        // ---- restricted leaf lengths, then both trees' restricted split lists ---------------------------------------------------
        __syncwarp();
        for (int k = lane; k < 64 * W; k += 32) {
            const bool in = k < p.n_taxa && ((M[k >> 6] >> (k & 63)) & 1ull);
            sm.leafA[k] = in ? p.leaflen[ti * p.n_taxa + k] : 0.0;
            sm.leafB[k] = in ? p.leaflen[tj * p.n_taxa + k] : 0.0;
        }
        __syncwarp();
        uint64_t mA[H][W], mB[H][W]; double lenA[H], lenB[H];
        // tree B first: the table is left holding the representatives of tree A
        const Bits<H> validB = restrict_tree<H, W>(p, tj, M, nM, ref, rooted, sm.core.maskB, sm.core.lenB, sm.leafB, sm.table, lane, mB, lenB);
        const Bits<H> validA = restrict_tree<H, W>(p, ti, M, nM, ref, rooted, sm.maskA, sm.lenA, sm.leafA, sm.table, lane, mA, lenA);

        // ---- leaf terms and per-tree normalisers ---------------------------------------------------------------------------------
        double l1 = 0, l2 = 0, sumA = 0, sumB = 0, sqA = 0, sqB = 0;
        for (int k = lane; k < 64 * W; k += 32) {
            const double a = sm.leafA[k], b = sm.leafB[k], d = a - b;
            l1 += fabs(d); l2 = fma(d, d, l2); sumA += a; sumB += b; sqA = fma(a, a, sqA); sqB = fma(b, b, sqB);
        }
        // ---- common splits: equal restricted masks; the lanes of tree B look their masks up in tree A's table ------------------------
        Bits<H> setA, setB;
        double c1 = 0, c2 = 0, o1 = 0, o2 = 0;      // common: |la-lb|, (la-lb)^2 ; one-sided: |l|, l^2
        double la2[H], lb2[H];
#pragma unroll
        for (int h = 0; h < H; h++) sm.partnerA[lane + 32 * h] = -1;
        __syncwarp();
#pragma unroll
        for (int h = 0; h < H; h++) {
            bool ncB = false;
            lb2[h] = lenB[h] * lenB[h];
            if (validB.mine(h, lane)) { sumB += lenB[h]; sqB += lb2[h]; }
            int partner = -1;
            uint32_t e = mask_hash<W>(mB[h]) & (64 * H - 1);
            bool looking = validB.mine(h, lane);
            while (__any_sync(FULL, looking)) {                 // warp-uniform probe loop (see restrict_tree)
                if (looking) {
                    const int a = sm.table[e];
                    if (a < 0) looking = false;
                    else if (mask_eq<W>(&sm.maskA[a * W], mB[h])) { partner = a; looking = false; }
                    else e = (e + 1) & (64 * H - 1);
                }
            }
            if (validB.mine(h, lane)) {
                if (partner >= 0) sm.partnerA[partner] = lane + 32 * h;
                else { ncB = true; o1 += fabs(lenB[h]); o2 += lb2[h]; }
            }
            setB.w[h] = __ballot_sync(FULL, ncB);
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < H; h++) {
            bool ncA = false;
            la2[h] = lenA[h] * lenA[h];
            if (validA.mine(h, lane)) {
                sumA += lenA[h]; sqA += la2[h];
                const int partner = sm.partnerA[lane + 32 * h];
                if (partner >= 0) { const double d = lenA[h] - sm.core.lenB[partner]; c1 += fabs(d); c2 = fma(d, d, c2); }
                else { ncA = true; o1 += fabs(lenA[h]); o2 += la2[h]; }
            }
            setA.w[h] = __ballot_sync(FULL, ncA);
        }
        const int rf = setA.count() + setB.count(), rf_den = validA.count() + validB.count();
        // only the sums the requested metrics read (ten warp reductions were a tenth of the kernel's instructions)
        if (p.metric_mask & 2u) { l1 = warp_sum(l1); c1 = warp_sum(c1); o1 = warp_sum(o1); if (p.normalise) { sumA = warp_sum(sumA); sumB = warp_sum(sumB); } }
        if (p.metric_mask & 12u) { l2 = warp_sum(l2); c2 = warp_sum(c2); o2 = warp_sum(o2); if (p.normalise) { sqA = warp_sum(sqA); sqB = warp_sum(sqB); } }

        double geo2 = 0.0;
        if constexpr (GEO) {
            double part = 0.0;
            const double total = gtp_noncommon<H, W>(sm.core, flow, FS, mA, setA, setB, la2, lb2, lane, part, st, failed);
            geo2 = warp_sum(part) + total + c2 + l2;
        }
        if (lane == 0) {
            const double nden = sqrt(sqA) + sqrt(sqB);
            if ((p.metric_mask & 1u) && p.outs[0]) { double v = (double)rf; if (p.normalise) v = rf_den ? v / (double)rf_den : 0.0; p.outs[0][o] = v; }
            if ((p.metric_mask & 2u) && p.outs[1]) { double v = (c1 + o1) + l1; if (p.normalise) { const double den = sumA + sumB; v = den != 0.0 ? v / den : 0.0; } p.outs[1][o] = v; }
            if ((p.metric_mask & 4u) && p.outs[2]) { double v = sqrt((c2 + o2) + l2); if (p.normalise) v = nden != 0.0 ? v / nden : 0.0; p.outs[2][o] = v; }
            if ((p.metric_mask & 8u) && p.outs[3]) {
                double v = sqrt(geo2); if (p.normalise) v = nden != 0.0 ? v / nden : 0.0;
                if (failed) v = kNaN;
                p.outs[3][o] = v;
            }
        }
        if (failed) { if (lane == 0) atomicOr(p.error_flag, 1); failed = false; }
    }
    if (lane == 0 && p.stats) {
        atomicAdd(&p.stats[0], st.pairs); atomicAdd(&p.stats[1], st.flows);
        atomicAdd(&p.stats[2], st.ratios); atomicAdd(&p.stats[3], st.aug);
    }
}

template <int H, int W>
cudaError_t plan_gen(const GeoParams &p, int n_sms, GeoPlan &plan)
{
    if (p.metric_mask & 8u)
        return plan_warp_kernel(general_kernel<H, W, true>, sizeof(GeneralSmem<H, W>), (size_t)p.flow_rows * (p.flow_rows | 1) * sizeof(double), p.p_end - p.p_begin, n_sms, plan);
    return plan_warp_kernel(general_kernel<H, W, false>, sizeof(GeneralSmem<H, W>), 0, p.p_end - p.p_begin, n_sms, plan);
}

template <int H, int W, bool GEO>
cudaError_t launch_gen_as(const GeoParams &p, const GeoPlan &plan, cudaStream_t st)
{
    auto k = general_kernel<H, W, GEO>;
    if ((plan.scratch_bytes != 0) != (p.flow_scratch != nullptr)) return cudaErrorInvalidValue;
    cudaError_t e = allow_full_smem((const void *)k);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)plan.blocks, plan.warps * 32, plan.smem, st>>>(p);
    return cudaGetLastError();
}

template <int H, int W>
cudaError_t launch_gen(const GeoParams &p, const GeoPlan &plan, cudaStream_t st)
{
    return (p.metric_mask & 8u) ? launch_gen_as<H, W, true>(p, plan, st) : launch_gen_as<H, W, false>(p, plan, st);
}

}  // namespace

#define GEN_DISPATCH(fn, ...)                                         \
    if (cap <= 32 && p.W == 1) return fn<1, 1>(__VA_ARGS__);          \
    if (cap <= 64 && p.W == 1) return fn<2, 1>(__VA_ARGS__);          \
    if (cap <= 64 && p.W == 2) return fn<2, 2>(__VA_ARGS__);          \
    if (cap <= 128 && p.W == 2) return fn<4, 2>(__VA_ARGS__);         \
    if (cap <= 128 && p.W == 3) return fn<4, 3>(__VA_ARGS__);         \
    return cudaErrorInvalidValue;

cudaError_t plan_general(const GeoParams &p, int cap, int n_sms, GeoPlan &plan) { GEN_DISPATCH(plan_gen, p, n_sms, plan) }
cudaError_t launch_general(const GeoParams &p, int cap, const GeoPlan &plan, cudaStream_t st) { GEN_DISPATCH(launch_gen, p, plan, st) }

}  // namespace td
