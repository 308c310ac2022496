// kernels_bigtrees.cu -- all four metrics for trees too large for the register-resident kernels (more than 128 internal edges per
// tree, i.e. more than ~130 taxa): one warp per pair, every per-split quantity in a per-warp slab of GLOBAL memory (L1 / L2 cached),
// only the warp-uniform split sets in shared memory.  No limit on the tree size but memory.
//
// Same semantics as kernels_geodesic.cu (same leaf set: getGeodesicDistance, treeCl/treedist.py:112) and kernels_general.cu (different
// leaf sets: restrict both trees to the common leaves first, treeCl/treedist.py:63-68, Tree.prune_to_subset treeCl/tree.py:989-1001),
// which stay the fast paths for ordinary tree sizes; this kernel removes the size limit the reference does not have either.
//   restrict   : mask & M, drop empty sides, fold one-leaf sides into the leaf edge, re-canonicalise, merge equal masks (lengths add)
//   common     : equal restricted masks (uniform sets: equal dictionary ids, binary search)
//   GTP        : incompatibility bit matrix, extended common-edge rule, ratio stack, min-weight vertex cover by max-flow (greedy
//                saturation + BFS augmenting paths), exactly the steps of geo_core.cuh with run-time sizes
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "geo_core.cuh"
#include "kernels.h"

namespace td {

namespace {

using geo::FULL;
using geo::FLOW_TOL;
using geo::SPLIT_TOL;
using geo::warp_sum;
using geo::tri_row_start;
using geo::PairFeed;
using geo::next_pair;

constexpr int BIG_WARPS = 4;
constexpr int kSetWords = 64;          // split sets of up to 2048 splits
constexpr int kSets = 11;

struct Slab {                          // per-warp scratch, carved from one allocation (sizes in BigParams)
    uint64_t *mA, *mB;                 // [S][W] restricted canonical masks (0: not an internal split of the restricted tree)
    double *lenA, *lenB, *la2, *lb2;   // [S]
    double *ra, *rb, *wa, *wb;         // [S] residual capacities and vertex weights of the current ratio
    int *parA, *parB;                  // [S] BFS parents
    uint32_t *inc, *incT;              // [S][HW] incompatibility rows: split a of A over B, split b of B over A
    uint32_t *fpos;                    // [S][HW] row a of "flow[a][b] > FLOW_TOL", kept in step with every write of the flow matrix (geo_core.cuh)
    uint32_t *stack;                   // [S + 2][2 HW] ratios still to be examined
    double *flow;                      // [S][FS]
    double *leafA, *leafB;             // [n_taxa_pad] restricted leaf-edge lengths
};

__device__ __forceinline__ bool set_any(const uint32_t *s, int hw) { uint32_t x = 0; for (int h = 0; h < hw; h++) x |= s[h]; return x != 0; }
__device__ __forceinline__ int set_count(const uint32_t *s, int hw) { int c = 0; for (int h = 0; h < hw; h++) c += __popc(s[h]); return c; }
__device__ __forceinline__ bool set_test(const uint32_t *s, int i) { return (s[i >> 5] >> (i & 31)) & 1u; }

}  // namespace

struct BigParams {
    GeoParams g;
    unsigned char *scratch;
    size_t warp_bytes;
    int S, HW, FS, W, taxa_pad;        // S: split slots per tree (multiple of 32); HW = S / 32; FS: row stride of the flow matrix
    int general;                       // 1: leaf sets may differ (restrict first); 0: one leaf set, common splits by dictionary id
};

namespace {

__device__ __forceinline__ Slab carve(const BigParams &bp, int warp_global)
{
    unsigned char *b = bp.scratch + (size_t)warp_global * bp.warp_bytes;
    Slab s;
    auto take = [&](size_t bytes) { unsigned char *p = b; b += (bytes + 15) & ~(size_t)15; return p; };
    const size_t S = bp.S;
    s.mA = (uint64_t *)take(S * bp.W * 8); s.mB = (uint64_t *)take(S * bp.W * 8);
    s.lenA = (double *)take(S * 8); s.lenB = (double *)take(S * 8); s.la2 = (double *)take(S * 8); s.lb2 = (double *)take(S * 8);
    s.ra = (double *)take(S * 8); s.rb = (double *)take(S * 8); s.wa = (double *)take(S * 8); s.wb = (double *)take(S * 8);
    s.parA = (int *)take(S * 4); s.parB = (int *)take(S * 4);
    s.inc = (uint32_t *)take(S * bp.HW * 4); s.incT = (uint32_t *)take(S * bp.HW * 4); s.fpos = (uint32_t *)take(S * bp.HW * 4);
    s.stack = (uint32_t *)take((S + 2) * 2 * bp.HW * 4);
    s.flow = (double *)take(S * (size_t)bp.FS * 8);
    s.leafA = (double *)take((size_t)bp.taxa_pad * 8); s.leafB = (double *)take((size_t)bp.taxa_pad * 8);
    return s;
}

// one tree restricted to the common leaf set M (general) or taken as it is (uniform): masks / lengths of its internal splits into
// (m, len), one-leaf sides folded into `leaf`, the set of valid (representative) splits into `valid`
__device__ void load_tree(const BigParams &bp, int64_t t, const uint64_t *M, int nM, int ref, int rooted, uint64_t *m, double *len,
                          double *leaf, uint32_t *valid, int lane)
{
    const GeoParams &p = bp.g;
    const int W = bp.W, S = p.nsplits[t];
    for (int h = 0; h < bp.HW; h++) {
        const int s = lane + 32 * h;
        bool internal = false; int leaf_k = -1; double l = 0.0;
        if (s < S) {
            l = p.lens[t * p.stride + s];
            const uint64_t *src = p.masks + (t * p.stride + s) * W;
            if (!bp.general) { internal = true; for (int w = 0; w < W; w++) m[(size_t)s * W + w] = src[w]; }
            else {
                int c = 0;
                for (int w = 0; w < W; w++) { const uint64_t x = src[w] & M[w]; m[(size_t)s * W + w] = x; c += __popcll(x); }
                const int cc = nM - c;
                if (c == 0 || cc == 0) { /* edge inside a pruned region, or the stem above the new root */ }
                else if (c == 1 || (!rooted && cc == 1)) {
                    for (int w = 0; w < W; w++) { const uint64_t x = c != 1 ? (M[w] & ~m[(size_t)s * W + w]) : m[(size_t)s * W + w]; if (x) leaf_k = 64 * w + __ffsll((long long)x) - 1; }
                } else {
                    internal = true;
                    if (!rooted && ((m[(size_t)s * W + (ref >> 6)] >> (ref & 63)) & 1ull))
                        for (int w = 0; w < W; w++) m[(size_t)s * W + w] = M[w] & ~m[(size_t)s * W + w];
                }
            }
        }
        // leaf folds, in slot order (one writer at a time -> deterministic sums)
        uint32_t folds = __ballot_sync(FULL, leaf_k >= 0);
        while (folds) {
            const int src = __ffs(folds) - 1; folds &= folds - 1;
            const int k = __shfl_sync(FULL, leaf_k, src); const double lf = __shfl_sync(FULL, l, src);
            if (lane == 0) leaf[k] += lf;
        }
        if (!internal) for (int w = 0; w < W; w++) m[(size_t)s * W + w] = 0ull;
        len[s] = internal ? l : 0.0;
        const uint32_t v = __ballot_sync(FULL, internal);
        if (lane == 0) valid[h] = v;
    }
    __syncwarp();
    if (!bp.general) return;
    // merge equal masks: the lowest slot is the representative and takes the sum (ascending slot order)
    for (int h = 0; h < bp.HW; h++) {
        const int s = lane + 32 * h;
        bool rep = set_test(valid, s);
        double sum = len[s];
        if (rep) {
            for (int t2 = 0; t2 < bp.S; t2++) {
                if (t2 == s || !set_test(valid, t2)) continue;
                bool eq = true;
                for (int w = 0; w < W; w++) eq &= m[(size_t)t2 * W + w] == m[(size_t)s * W + w];
                if (!eq) continue;
                if (t2 < s) { rep = false; break; }
                sum += len[t2];
            }
        }
        // in place: a representative only adds the lengths of LATER equal slots (never representatives, so never overwritten with a sum)
        const uint32_t r = __ballot_sync(FULL, rep);
        if (rep) len[s] = sum;
        if (lane == 0) valid[bp.HW + h] = r;               // representatives, staged behind the valid set
    }
    __syncwarp();
    for (int h = lane; h < bp.HW; h += 32) valid[h] = valid[bp.HW + h];
    __syncwarp();
    // lengths of merged-away slots must not count: zero them (their masks stay, but they are no longer in `valid`)
    for (int s = lane; s < bp.S; s += 32) if (!set_test(valid, s)) len[s] = 0.0;
    __syncwarp();
}

__global__ void __launch_bounds__(BIG_WARPS * 32) bigtree_kernel(const BigParams bp)
{
    const GeoParams &p = bp.g;
    __shared__ uint32_t sets_all[BIG_WARPS][kSets + 1][kSetWords];
    __shared__ uint64_t M_all[BIG_WARPS][kSetWords];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int HW = bp.HW, W = bp.W, FS = bp.FS;
    uint32_t (*sets)[kSetWords] = sets_all[wl];
    uint32_t *validA = sets[0], *validB = sets[1], *setA = sets[2], *setB = sets[3], *RA = sets[4], *RB = sets[5], *visA = sets[6], *visB = sets[7],
             *frontA = sets[8], *newA = sets[9], *newB = sets[10];
    // validA / validB need 2 HW words while load_tree stages the representatives: sets[0] and sets[1] are followed by enough room only if
    // 2 HW <= kSetWords, which the launcher guarantees (S <= 1024)
    uint64_t *M = M_all[wl];
    const Slab sl = carve(bp, blockIdx.x * BIG_WARPS + wl);
    const double kNaN = __longlong_as_double(0x7ff8000000000000ll);
    unsigned long long n_pairs = 0, n_flows = 0, n_ratios = 0, n_aug = 0;

    PairFeed feed;
    for (;;) {
        int64_t pidx, ti, tj;
        if (!next_pair(p, feed, lane, pidx, ti, tj)) break;
        n_pairs++;
        const int64_t o = pidx - p.p_base;
        auto write_all = [&](double v) {
            if (lane == 0) for (int m = 0; m < 4; m++) if ((p.metric_mask >> m & 1u) && p.outs[m]) p.outs[m][o] = v;
        };
        __syncwarp();

        // ---- common leaf set ---------------------------------------------------------------------------------------------------------
        int nM = 0, ref = -1, rooted = 0; bool same = true;
        if (bp.general) {
            for (int w = 0; w < W; w++) {
                const uint64_t a = p.leafmask[ti * W + w], b = p.leafmask[tj * W + w];
                same &= (a == b); nM += __popcll(a & b);
                if (lane == 0) M[w] = a & b;
                if (ref < 0 && (a & b)) ref = 64 * w + __ffsll((long long)(a & b)) - 1;
            }
            rooted = p.rooted[ti];
            if (rooted != p.rooted[tj]) { write_all(kNaN); if (lane == 0) atomicOr(p.error_flag, 2); continue; }
            if (!same && nM < p.min_overlap) { write_all(p.fail_value); continue; }       // treeCl/treedist.py:64-65
        }
        __syncwarp();
        for (int k = lane; k < bp.taxa_pad; k += 32) {
            const bool in = k < p.n_taxa && (!bp.general || ((M[k >> 6] >> (k & 63)) & 1ull));
            sl.leafA[k] = in ? p.leaflen[ti * p.n_taxa + k] : 0.0;
            sl.leafB[k] = in ? p.leaflen[tj * p.n_taxa + k] : 0.0;
        }
        __syncwarp();
        load_tree(bp, ti, M, nM, ref, rooted, sl.mA, sl.lenA, sl.leafA, validA, lane);
        load_tree(bp, tj, M, nM, ref, rooted, sl.mB, sl.lenB, sl.leafB, validB, lane);

        // ---- leaf terms and per-tree normalisers -----------------------------------------------------------------------------------------
        double l1 = 0, l2 = 0, sumA = 0, sumB = 0, sqA = 0, sqB = 0;
        for (int k = lane; k < bp.taxa_pad; k += 32) {
            const double a = sl.leafA[k], b = sl.leafB[k], d = a - b;
            l1 += fabs(d); l2 = fma(d, d, l2); sumA += a; sumB += b; sqA = fma(a, a, sqA); sqB = fma(b, b, sqB);
        }
        // ---- common splits -------------------------------------------------------------------------------------------------------------------
        double c1 = 0, c2 = 0, o1 = 0, o2 = 0;      // common: |la-lb|, (la-lb)^2 ; one-sided: |l|, l^2
        const int Si = p.nsplits[ti], Sj = p.nsplits[tj];
        for (int h = 0; h < HW; h++) {
            const int s = lane + 32 * h;
            bool ncA = false, ncB = false;
            const double la = sl.lenA[s], lb = sl.lenB[s];
            sl.la2[s] = la * la; sl.lb2[s] = lb * lb;
            if (set_test(validA, s)) {
                sumA += la; sqA = fma(la, la, sqA);
                int partner = -1;
                if (!bp.general) partner = geo::find_id(p.ids + tj * p.stride, Sj, p.ids[ti * p.stride + s]);
                else for (int b = 0; b < bp.S && partner < 0; b++) {
                    if (!set_test(validB, b)) continue;
                    bool eq = true;
                    for (int w = 0; w < W; w++) eq &= sl.mB[(size_t)b * W + w] == sl.mA[(size_t)s * W + w];
                    if (eq) partner = b;
                }
                if (partner >= 0) { const double d = la - sl.lenB[partner]; c1 += fabs(d); c2 = fma(d, d, c2); }
                else { ncA = true; o1 += fabs(la); o2 = fma(la, la, o2); }
            }
            if (set_test(validB, s)) {
                sumB += lb; sqB = fma(lb, lb, sqB);
                bool found = false;
                if (!bp.general) found = geo::find_id(p.ids + ti * p.stride, Si, p.ids[tj * p.stride + s]) >= 0;
                else for (int a = 0; a < bp.S && !found; a++) {
                    if (!set_test(validA, a)) continue;
                    bool eq = true;
                    for (int w = 0; w < W; w++) eq &= sl.mA[(size_t)a * W + w] == sl.mB[(size_t)s * W + w];
                    found = eq;
                }
                if (!found) { ncB = true; o1 += fabs(lb); o2 = fma(lb, lb, o2); }
            }
            const uint32_t ba = __ballot_sync(FULL, ncA), bb = __ballot_sync(FULL, ncB);
            if (lane == 0) { setA[h] = ba; setB[h] = bb; }
        }
        __syncwarp();
        const int rf = set_count(setA, HW) + set_count(setB, HW), rf_den = set_count(validA, HW) + set_count(validB, HW);
        l1 = warp_sum(l1); l2 = warp_sum(l2); c1 = warp_sum(c1); c2 = warp_sum(c2); o1 = warp_sum(o1); o2 = warp_sum(o2);
        sumA = warp_sum(sumA); sumB = warp_sum(sumB); sqA = warp_sum(sqA); sqB = warp_sum(sqB);

        double geo2 = 0.0; bool failed = false;
        if (p.metric_mask & 8u) {
            double part = 0.0, total = 0.0;
            // ---- incompatibility bit matrix --------------------------------------------------------------------------------------------------
            for (int s = lane; s < bp.S; s += 32) for (int h = 0; h < HW; h++) { sl.inc[(size_t)s * HW + h] = 0u; sl.incT[(size_t)s * HW + h] = 0u; sl.fpos[(size_t)s * HW + h] = 0u; }
            __syncwarp();
            for (int hb = 0; hb < HW; hb++) {
                uint32_t wordB = setB[hb];
                while (wordB) {
                    const int b = 32 * hb + __ffs(wordB) - 1; wordB &= wordB - 1;
                    for (int h = 0; h < HW; h++) {
                        const int a = lane + 32 * h;
                        bool incompatible = false;
                        if (set_test(setA, a)) {
                            bool disjoint = true, asub = true, bsub = true;
                            for (int w = 0; w < W; w++) {
                                const uint64_t ma = sl.mA[(size_t)a * W + w], mb = sl.mB[(size_t)b * W + w], x = ma & mb;
                                disjoint &= (x == 0); asub &= (x == ma); bsub &= (x == mb);
                            }
                            incompatible = !(disjoint || asub || bsub);
                        }
                        if (incompatible) sl.inc[(size_t)a * HW + hb] |= 1u << (b & 31);
                        const uint32_t col = __ballot_sync(FULL, incompatible);
                        if (lane == 0) sl.incT[(size_t)b * HW + h] = col;
                    }
                }
            }
            __syncwarp();
            // ---- extended common-edge rule: a split compatible with every split of the other tree contributes its squared length -------------
            for (int h = 0; h < HW; h++) {
                const int s = lane + 32 * h;
                bool extA = false, extB = false;
                if (set_test(setA, s)) { uint32_t x = 0; for (int k = 0; k < HW; k++) x |= sl.inc[(size_t)s * HW + k]; extA = x == 0; }
                if (set_test(setB, s)) { uint32_t x = 0; for (int k = 0; k < HW; k++) x |= sl.incT[(size_t)s * HW + k]; extB = x == 0; }
                if (extA) part += sl.la2[s];
                if (extB) part += sl.lb2[s];
                const uint32_t ea = __ballot_sync(FULL, extA), eb = __ballot_sync(FULL, extB);
                __syncwarp();
                if (lane == 0) { setA[h] &= ~ea; setB[h] &= ~eb; }
                __syncwarp();
            }
            // ---- GTP ratio stack ------------------------------------------------------------------------------------------------------------------
            int sp = 0;
            if (set_any(setA, HW) && set_any(setB, HW)) {
                for (int h = lane; h < HW; h += 32) { sl.stack[h] = setA[h]; sl.stack[HW + h] = setB[h]; }
                sp = 1;
            }
            __syncwarp();
            int guard = 0;
            while (sp > 0 && !failed) {
                sp--;
                for (int h = lane; h < HW; h += 32) { RA[h] = sl.stack[(size_t)sp * 2 * HW + h]; RB[h] = sl.stack[(size_t)sp * 2 * HW + HW + h]; }
                __syncwarp();
                double sa = 0.0, sb = 0.0;
                for (int s = lane; s < bp.S; s += 32) { if (set_test(RA, s)) sa += sl.la2[s]; if (set_test(RB, s)) sb += sl.lb2[s]; }
                sa = warp_sum(sa); sb = warp_sum(sb);
                bool split = false;
                if (set_count(RA, HW) + set_count(RB, HW) > 2 && sa > 0.0 && sb > 0.0) {
                    n_flows++;
                    for (int s = lane; s < bp.S; s += 32) {
                        const bool ina = set_test(RA, s), inb = set_test(RB, s);
                        const double wa = ina ? sl.la2[s] / sa : 0.0, wb = inb ? sl.lb2[s] / sb : 0.0;
                        sl.wa[s] = wa; sl.wb[s] = wb; sl.ra[s] = wa; sl.rb[s] = wb;
                        // no edge carries flow.  A word without a neighbour in RB may keep bits of an earlier max-flow OF THIS PAIR: they sit on edges
                        // whose B end is outside RB, and every read masks with a subset of RB (rows are cleared in full when a new pair starts).  The flow matrix itself is never cleared
                        if (ina) for (int h = 0; h < HW; h++) if (sl.inc[(size_t)s * HW + h] & RB[h]) sl.fpos[(size_t)s * HW + h] = 0u;
                    }
                    __syncwarp();
                    // greedy saturation: push every source edge into its neighbours' sink edges, in split order
                    for (int ha = 0; ha < HW; ha++) {
                        uint32_t wordA = RA[ha];
                        while (wordA) {
                            const int a = 32 * ha + __ffs(wordA) - 1; wordA &= wordA - 1;
                            double x = sl.ra[a];
                            for (int h = 0; h < HW; h++) {
                                const uint32_t nbm = sl.inc[(size_t)a * HW + h] & RB[h];
                                if (!(x > FLOW_TOL) || nbm == 0) continue;       // warp-uniform
                                const int b = lane + 32 * h;
                                const double avail = ((nbm >> lane) & 1u) ? sl.rb[b] : 0.0;
                                double incl = avail;
                                for (int d = 1; d < 32; d <<= 1) { const double t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
                                const double tot = __shfl_sync(FULL, incl, 31);
                                const double take = fmin(avail, fmax(0.0, x - (incl - avail)));
                                if (take > 0.0) { sl.rb[b] -= take; sl.flow[(size_t)a * FS + b] = take; }
                                const uint32_t carries = __ballot_sync(FULL, take > FLOW_TOL);
                                if (lane == 0) sl.fpos[(size_t)a * HW + h] = carries;
                                x = fmax(0.0, x - tot);
                            }
                            __syncwarp();
                            if (lane == 0) sl.ra[a] = x;
                            __syncwarp();
                        }
                    }
                    // augmenting paths by BFS over warp-uniform frontier bitmasks
                    for (;;) {
                        if (++guard > 200000) { failed = true; break; }
                        for (int s = lane; s < bp.S; s += 32) { sl.parA[s] = -1; sl.parB[s] = -1; }
                        for (int h = 0; h < HW; h++) {
                            const uint32_t f = __ballot_sync(FULL, sl.ra[lane + 32 * h] > FLOW_TOL && set_test(RA, lane + 32 * h));
                            if (lane == 0) { frontA[h] = f; visA[h] = f; visB[h] = 0u; }
                        }
                        __syncwarp();
                        int sink = -1;
                        while (set_any(frontA, HW)) {
                            bool any_new = false; int first_sink = -1;
                            for (int h = 0; h < HW; h++) {
                                const int b = lane + 32 * h;
                                bool nb = false;
                                if (set_test(RB, b) && !set_test(visB, b)) {
                                    for (int k = 0; k < HW; k++) {
                                        const uint32_t reach = sl.incT[(size_t)b * HW + k] & frontA[k];
                                        if (reach) { nb = true; sl.parB[b] = 32 * k + __ffs(reach) - 1; break; }
                                    }
                                }
                                const uint32_t nbw = __ballot_sync(FULL, nb);
                                const uint32_t sinks = __ballot_sync(FULL, nb && sl.rb[b] > FLOW_TOL);
                                if (lane == 0) newB[h] = nbw;
                                any_new |= nbw != 0;
                                if (first_sink < 0 && sinks) first_sink = 32 * h + __ffs(sinks) - 1;
                            }
                            __syncwarp();
                            if (!any_new) break;
                            for (int h = lane; h < HW; h += 32) visB[h] |= newB[h];
                            __syncwarp();
                            if (first_sink >= 0) { sink = first_sink; break; }
                            // B -> A through edges that carry flow
                            for (int h = 0; h < HW; h++) {
                                const int a = lane + 32 * h;
                                bool found = false;
                                if (set_test(RA, a) && !set_test(visA, a)) {
                                    for (int k = 0; k < HW && !found; k++) {
                                        const uint32_t cand = sl.fpos[(size_t)a * HW + k] & newB[k];
                                        if (cand) { sl.parA[a] = 32 * k + __ffs(cand) - 1; found = true; }
                                    }
                                }
                                const uint32_t na = __ballot_sync(FULL, found);
                                if (lane == 0) newA[h] = na;
                            }
                            __syncwarp();
                            for (int h = lane; h < HW; h += 32) { visA[h] |= newA[h]; frontA[h] = newA[h]; }
                            __syncwarp();
                        }
                        if (sink < 0) break;            // no augmenting path: (visA, visB) is the source side of a min cut
                        n_aug++;
                        __syncwarp();
                        if (lane == 0) {                // the path is short; one lane walks it (bottleneck, then update)
                            double delta = sl.rb[sink];
                            int b = sink;
                            for (int hop = 0; hop < 2 * bp.S + 2; hop++) {
                                const int a = sl.parB[b], pb = sl.parA[a];
                                if (pb < 0) { delta = fmin(delta, sl.ra[a]); break; }
                                delta = fmin(delta, sl.flow[(size_t)a * FS + pb]); b = pb;
                            }
                            sl.rb[sink] -= delta;
                            b = sink;
                            for (int hop = 0; hop < 2 * bp.S + 2; hop++) {
                                const int a = sl.parB[b], pb = sl.parA[a];
                                uint32_t &wb_ = sl.fpos[(size_t)a * HW + (b >> 5)];
                                const double fwd = ((wb_ >> (b & 31)) & 1u ? sl.flow[(size_t)a * FS + b] : 0.0) + delta;
                                sl.flow[(size_t)a * FS + b] = fwd;
                                if (fwd > FLOW_TOL) wb_ |= 1u << (b & 31); else wb_ &= ~(1u << (b & 31));
                                if (pb < 0) { sl.ra[a] -= delta; break; }
                                const double back = sl.flow[(size_t)a * FS + pb] - delta;      // parA[a] = pb was taken from fpos: the cell is live
                                sl.flow[(size_t)a * FS + pb] = back;
                                if (!(back > FLOW_TOL)) sl.fpos[(size_t)a * HW + (pb >> 5)] &= ~(1u << (pb & 31));
                                b = pb;
                            }
                        }
                        __syncwarp();
                    }
                    if (failed) break;
                    // min-weight vertex cover = (A not reachable) + (B reachable); CA / CB reuse newA / newB
                    for (int h = lane; h < HW; h += 32) { newA[h] = RA[h] & ~visA[h]; newB[h] = visB[h] & RB[h]; }
                    __syncwarp();
                    double cw = 0.0;
                    for (int s = lane; s < bp.S; s += 32) { if (set_test(newA, s)) cw += sl.wa[s]; if (set_test(newB, s)) cw += sl.wb[s]; }
                    cw = warp_sum(cw);
                    bool eqA = true, eqB = true;
                    for (int h = 0; h < HW; h++) { eqA &= newA[h] == RA[h]; eqB &= newB[h] == RB[h]; }
                    split = cw < 1.0 - SPLIT_TOL && set_any(newA, HW) && !eqA && set_any(newB, HW) && !eqB;
                }
                if (split) {
                    // (C&A, B\C) precedes (A\C, C&B) in the ratio sequence; the sum does not depend on the order
                    for (int h = lane; h < HW; h += 32) {
                        sl.stack[(size_t)sp * 2 * HW + h] = RA[h] & ~newA[h]; sl.stack[(size_t)sp * 2 * HW + HW + h] = newB[h];
                        sl.stack[(size_t)(sp + 1) * 2 * HW + h] = newA[h]; sl.stack[(size_t)(sp + 1) * 2 * HW + HW + h] = RB[h] & ~newB[h];
                    }
                    sp += 2;
                    __syncwarp();
                } else {
                    const double s = sqrt(sa) + sqrt(sb);
                    total = fma(s, s, total); n_ratios++;
                }
                __syncwarp();
            }
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
            if (failed) atomicOr(p.error_flag, 1);
        }
    }
    if (lane == 0 && p.stats) {
        atomicAdd(&p.stats[0], n_pairs); atomicAdd(&p.stats[1], n_flows);
        atomicAdd(&p.stats[2], n_ratios); atomicAdd(&p.stats[3], n_aug);
    }
}

}  // namespace

// per-warp scratch bytes for trees of up to max_splits internal edges over W mask words
size_t bigtree_warp_bytes(int max_splits, int W, int n_taxa)
{
    const size_t S = (size_t)((std::max(max_splits, 1) + 31) / 32 * 32), HW = S / 32, FS = S | 1, tp = (size_t)((n_taxa + 63) / 64 * 64);
    auto up = [](size_t b) { return (b + 15) & ~(size_t)15; };
    return 2 * up(S * W * 8) + 8 * up(S * 8) + 2 * up(S * 4) + 3 * up(S * HW * 4) + up((S + 2) * 2 * HW * 4) + up(S * FS * 8) + 2 * up(tp * 8) + 64;
}

int bigtree_max_splits() { return 32 * (kSetWords / 2); }

// blocks of BIG_WARPS warps this launch wants (the caller sizes the scratch: blocks * BIG_WARPS * bigtree_warp_bytes)
int bigtree_blocks(int64_t pairs, int n_sms, size_t warp_bytes, size_t scratch_budget)
{
    int64_t blocks = (int64_t)n_sms * 4;
    const int64_t need = (pairs + BIG_WARPS - 1) / BIG_WARPS;
    if (blocks > need) blocks = need;
    const int64_t fit = (int64_t)(scratch_budget / (warp_bytes * BIG_WARPS));
    if (blocks > fit) blocks = fit;
    return (int)std::max<int64_t>(blocks, 1);
}

cudaError_t launch_bigtree(const GeoParams &g, int general, int max_splits, unsigned char *scratch, int blocks, cudaStream_t st)
{
    BigParams bp{};
    bp.g = g; bp.scratch = scratch; bp.general = general;
    bp.S = (std::max(max_splits, 1) + 31) / 32 * 32; bp.HW = bp.S / 32; bp.FS = bp.S | 1; bp.W = g.W; bp.taxa_pad = (g.n_taxa + 63) / 64 * 64;
    if (2 * bp.HW > kSetWords || g.W > kSetWords) return cudaErrorInvalidValue;
    bp.warp_bytes = bigtree_warp_bytes(max_splits, g.W, g.n_taxa);
    bigtree_kernel<<<(unsigned)blocks, BIG_WARPS * 32, 0, st>>>(bp);
    return cudaGetLastError();
}

}  // namespace td
