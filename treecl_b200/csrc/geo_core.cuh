// geo_core.cuh -- warp-level GTP machinery shared by the geodesic kernel (uniform sets) and the general kernel
// (pairs with different leaf sets).  See kernels_geodesic.cu for the algorithm description.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace td {
namespace geo {

constexpr unsigned FULL = 0xffffffffu;
constexpr int GEO_WARPS = 4;
// Min resident 4-warp blocks per SM (= register cap) of geodesic_kernel<H, W>.  With the flow matrix in global memory registers limit
// residency, and the kernel is latency-bound: measured in one call each (gpurun_out/r2_geo_ab4.log, r2_geo_ab5.log), 2,000 x 50 (H = 2):
// 192 ms at 96 registers, 164 at 80 (6 blocks), 139 at 64 (8), 133 at 48 (10), 128 at 40 (12); 1,000 x 100 (H = 4): 362 ms uncapped, 230 at
// 5 blocks, 227 at 6, 232 at 7, 240 at 8; 2,000 x 30 (H = 1, 70 registers uncapped): 35.8 ms at 8 blocks, 37.8 at 10, 36.9 at 12.
#ifndef GEO_MINB_H1
#define GEO_MINB_H1 8
#endif
#ifndef GEO_MINB_H2
#define GEO_MINB_H2 12
#endif
#ifndef GEO_MINB_H4
#define GEO_MINB_H4 6
#endif
constexpr int geo_min_blocks(int H) { return H == 1 ? GEO_MINB_H1 : (H == 2 ? GEO_MINB_H2 : GEO_MINB_H4); }
constexpr double FLOW_TOL = 1e-13;
constexpr double SPLIT_TOL = 1e-12;

template <int H>
struct Bits {
    uint32_t w[H];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int h = 0; h < H; h++) w[h] = 0; }
    __device__ __forceinline__ bool any() const { uint32_t x = 0;
#pragma unroll
        for (int h = 0; h < H; h++) x |= w[h]; return x != 0; }
    __device__ __forceinline__ int count() const { int c = 0;
#pragma unroll
        for (int h = 0; h < H; h++) c += __popc(w[h]); return c; }
    // i is warp-uniform; every word is touched with a computed mask (no indexing by i, so w[] stays in registers)
    __device__ __forceinline__ bool test(int i) const { uint32_t x = 0;
#pragma unroll
        for (int h = 0; h < H; h++) x |= w[h] & (0u - (uint32_t)((i >> 5) == h)); return (x >> (i & 31)) & 1u; }
    __device__ __forceinline__ bool mine(int h, int lane) const { return (w[h] >> lane) & 1u; }   // h compile-time
    __device__ __forceinline__ int first() const {
        int r = -1;
#pragma unroll
        for (int h = H - 1; h >= 0; h--) if (w[h]) r = 32 * h + __ffs(w[h]) - 1;
        return r; }
    // lowest set bit, cleared (the set must not be empty)
    __device__ __forceinline__ int pop_first() {
        int r = -1; bool done = false;
#pragma unroll
        for (int h = 0; h < H; h++) if (!done && w[h]) { r = 32 * h + __ffs(w[h]) - 1; w[h] &= w[h] - 1; done = true; }
        return r; }
    __device__ __forceinline__ void reset(int i) {
#pragma unroll
        for (int h = 0; h < H; h++) w[h] &= ~((uint32_t)((i >> 5) == h) << (i & 31)); }
    __device__ __forceinline__ void set(int i) {
#pragma unroll
        for (int h = 0; h < H; h++) w[h] |= (uint32_t)((i >> 5) == h) << (i & 31); }
    __device__ __forceinline__ bool equals(const Bits &o) const { bool e = true;
#pragma unroll
        for (int h = 0; h < H; h++) e &= (w[h] == o.w[h]); return e; }
};
template <int H> __device__ __forceinline__ Bits<H> operator&(const Bits<H> &a, const Bits<H> &b)
{
    Bits<H> r;
#pragma unroll
    for (int h = 0; h < H; h++) r.w[h] = a.w[h] & b.w[h];
    return r;
}
template <int H> __device__ __forceinline__ Bits<H> andnot(const Bits<H> &a, const Bits<H> &b)
{
    Bits<H> r;
#pragma unroll
    for (int h = 0; h < H; h++) r.w[h] = a.w[h] & ~b.w[h];
    return r;
}

// value owned by lane (idx & 31), slot (idx >> 5); idx is warp-uniform
template <int H>
__device__ __forceinline__ double fetch_d(const double (&v)[H], int idx)
{
    double r = __shfl_sync(FULL, v[0], idx & 31);
#pragma unroll
    for (int h = 1; h < H; h++) { const double rh = __shfl_sync(FULL, v[h], idx & 31); if ((idx >> 5) == h) r = rh; }
    return r;
}
template <int H>
__device__ __forceinline__ int fetch_i(const int (&v)[H], int idx)
{
    int r = __shfl_sync(FULL, v[0], idx & 31);
#pragma unroll
    for (int h = 1; h < H; h++) { const int rh = __shfl_sync(FULL, v[h], idx & 31); if ((idx >> 5) == h) r = rh; }
    return r;
}
template <int H>
__device__ __forceinline__ Bits<H> fetch_bits(const Bits<H> (&v)[H], int idx)
{
    Bits<H> r;
#pragma unroll
    for (int k = 0; k < H; k++) {
        uint32_t x = __shfl_sync(FULL, v[0].w[k], idx & 31);
#pragma unroll
        for (int h = 1; h < H; h++) { const uint32_t xh = __shfl_sync(FULL, v[h].w[k], idx & 31); if ((idx >> 5) == h) x = xh; }
        r.w[k] = x;
    }
    return r;
}

template <int H>
__device__ __forceinline__ void owner_store(double (&v)[H], int idx, int lane, double val)
{
#pragma unroll
    for (int h = 0; h < H; h++) if ((idx >> 5) == h && lane == (idx & 31)) v[h] = val;
}
template <int H>
__device__ __forceinline__ void owner_sub(double (&v)[H], int idx, int lane, double val)
{
#pragma unroll
    for (int h = 0; h < H; h++) if ((idx >> 5) == h && lane == (idx & 31)) v[h] -= val;
}

__device__ __forceinline__ double warp_sum(double v)
{
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

__device__ __forceinline__ int find_id(const uint32_t *arr, int n, uint32_t key)
{
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (arr[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < n && arr[lo] == key) ? lo : -1;
}

__device__ __forceinline__ int64_t tri_row_start(int64_t i, int64_t N) { return i * N - i * (i + 1) / 2; }

// Work distribution of the warp-per-pair kernels that may SKIP pairs (kernels_general.cu, kernels_bigtrees.cu): a warp takes a chunk of 32
// consecutive pair indices from the launch's work counter, every lane decodes its own pair and tests whether it is wanted - pairs of two
// trees of the same leaf-set group (GeoParams::group_of) belong to that group's uniform kernels - and the warp then walks the wanted
// pairs of the chunk one by one.  That costs one chunk per 32 skipped pairs (a 5,000-tree set with 50 odd trees: 390 k chunks for 250 k
// wanted pairs), so the library normally hands over the wanted pairs as lists instead (GeoParams::cross_*): the chunk then holds 32
// WANTED pairs, found by a binary search over the per-row prefix.
struct PairFeed {
    uint32_t pending = 0;
    int64_t base = 0, my_ti = 0, my_tj = 0;
};

__device__ __forceinline__ void decode_pair(const GeoParams &p, int64_t pidx, int64_t &ti, int64_t &tj)
{
    if (p.rect) { ti = pidx / p.rect_cols; tj = p.rect_split + pidx % p.rect_cols; return; }
    const double Nd = (double)p.N;
    int64_t i = (int64_t)floor(((2.0 * Nd - 1.0) - sqrt((2.0 * Nd - 1.0) * (2.0 * Nd - 1.0) - 8.0 * (double)pidx)) * 0.5);
    if (i < 0) i = 0; if (i > p.N - 2) i = p.N - 2;
    while (i + 1 <= p.N - 2 && tri_row_start(i + 1, p.N) <= pidx) i++;
    while (i > 0 && tri_row_start(i, p.N) > pidx) i--;
    ti = i; tj = pidx - tri_row_start(i, p.N) + i + 1;
}

// q-th wanted pair of the cross-group lists: the row is the last one whose prefix is <= q
__device__ __forceinline__ void decode_cross(const GeoParams &p, int64_t q, int64_t &ti, int64_t &tj)
{
    int64_t lo = 0, hi = p.N;                                   // cross_prefix has N + 1 entries; prefix[lo] <= q < prefix[hi]
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (p.cross_prefix[mid] <= q) lo = mid; else hi = mid; }
    const int64_t off = q - p.cross_prefix[lo], first = p.cross_first[lo];
    ti = lo; tj = first < 0 ? lo + 1 + off : (int64_t)p.cross_cols[first + off];
}

// false: no work left.  All lanes of the warp call it together and receive the same pair; pidx is its condensed (or rectangular) index.
__device__ __forceinline__ bool next_pair(const GeoParams &p, PairFeed &f, int lane, int64_t &pidx, int64_t &ti, int64_t &tj)
{
    while (f.pending == 0) {
        unsigned long long wk = 0;
        if (lane == 0) wk = atomicAdd(p.work_counter, 1ull);
        wk = __shfl_sync(FULL, wk, 0);
        f.base = p.p_begin + (int64_t)wk * 32;
        if (f.base >= p.p_end) return false;
        const int64_t mine = f.base + lane;
        bool want = mine < p.p_end;
        if (want) {
            if (p.cross_prefix) decode_cross(p, mine, f.my_ti, f.my_tj);
            else {
                decode_pair(p, mine, f.my_ti, f.my_tj);
                if (p.group_of) { const int ga = p.group_of[f.my_ti]; want = ga < 0 || ga != p.group_of[f.my_tj]; }
            }
        }
        f.pending = __ballot_sync(FULL, want);
    }
    const int src = __ffs(f.pending) - 1;
    f.pending &= f.pending - 1;
    ti = __shfl_sync(FULL, f.my_ti, src); tj = __shfl_sync(FULL, f.my_tj, src);
    pidx = p.cross_prefix ? tri_row_start(ti, p.N) + tj - ti - 1 : f.base + src;
    return true;
}

// Warps per CTA of a warp-per-pair kernel whose warps each own `warp_bytes` of dynamic shared memory: the block size (4, 2 or 1 warps)
// that puts the most warps on an SM.  These kernels are latency-bound (ncu: 8 warps/SM at 50 taxa with 4-warp blocks, issue slots
// 35 % busy), and the per-warp flow matrix is what limits residency, so the granularity of a block matters.
template <class Kernel>
inline cudaError_t pick_block_warps(Kernel k, size_t warp_bytes, int &warps, int &per_sm)
{
    warps = 0; per_sm = 0;
    if (warp_bytes > 225 * 1024) return cudaSuccess;          // does not fit: 0 warps
    int first = GEO_WARPS;
    while (first > 1 && warp_bytes * first > 200 * 1024) first >>= 1;
    cudaError_t e = allow_full_smem((const void *)k);
    if (e != cudaSuccess) return e;
    (void)cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    for (int w = first; w >= 1; w >>= 1) {
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, w * 32, warp_bytes * w);
        if (e != cudaSuccess) return e;
        if (n * w > per_sm * warps) { warps = w; per_sm = n; }
    }
    return cudaSuccess;
}

// Shared or global flow matrix (kernels.h GeoPlan): global when it keeps at least 1.3 x the warps resident.  TREEDIST_GEO_FLOW=smem|global
// forces one form (A/B runs).
template <class Kernel>
inline cudaError_t plan_warp_kernel(Kernel k, size_t fixed_bytes, size_t flow_bytes, int64_t pairs, int n_sms, GeoPlan &plan)
{
    int ws = 0, ns = 0, wg = 0, ng = 0;
    cudaError_t e = pick_block_warps(k, fixed_bytes, wg, ng);
    if (e != cudaSuccess) return e;
    e = pick_block_warps(k, fixed_bytes + flow_bytes, ws, ns);
    if (e != cudaSuccess) return e;
    if (wg * ng == 0) return cudaErrorInvalidConfiguration;
    bool global = ws * ns == 0 || wg * ng * 10 >= ws * ns * 13;
    if (const char *f = getenv("TREEDIST_GEO_FLOW")) { if (!strcmp(f, "global")) global = true; else if (!strcmp(f, "smem") && ws * ns > 0) global = false; }
    const int warps = global ? wg : ws, per_sm = global ? ng : ns;
    plan.warps = warps; plan.smem = (fixed_bytes + (global ? 0 : flow_bytes)) * warps;
    int64_t blocks = (int64_t)n_sms * per_sm;
    const int64_t need = (pairs + warps - 1) / warps;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    plan.blocks = (int)blocks;
    plan.scratch_bytes = global ? (size_t)blocks * warps * flow_bytes : 0;
    return cudaSuccess;
}

template <int H, int W>
struct WarpSmem {
    static constexpr int CAP = 32 * H;
    double lenB[CAP];
    uint64_t maskB[CAP * W];
    uint32_t idA[CAP], idB[CAP];
    uint32_t stack[CAP + 2][2 * H];
};


struct Counters { unsigned long long pairs = 0, flows = 0, ratios = 0, aug = 0; };

// Steps 3-5 for one pair, executed by a full warp.
//   flow    : this warp's fp64 flow matrix, rows a (splits of A) x FS columns (FS odd >= splits of B)
//   mA[h]   : mask of my split (lane + 32h) of tree A (canonical side);  sm.maskB[s] : masks of tree B's splits
//   setA/B  : non-common splits of each tree;  la2/lb2 : squared lengths of my splits
// Adds the extended-common-edge terms to `part` (per-lane partial) and returns sum_k (||A_k|| + ||B_k||)^2 (warp-uniform).
template <int H, int W>
__device__ __forceinline__ double gtp_noncommon(WarpSmem<H, W> &sm, double *flow, const int FS, const uint64_t (&mA)[H][W], Bits<H> setA, Bits<H> setB,
                                                const double (&la2)[H], const double (&lb2)[H], const int lane,
                                                double &part, Counters &st, bool &failed)
{
    constexpr int CAP = 32 * H;
    // ---- 3. incompatibility bit matrix: inc[h] row of my split a over B; incT[h] row of my split b over A ----
    Bits<H> inc[H], incT[H];
#pragma unroll
    for (int h = 0; h < H; h++) { inc[h].clear(); incT[h].clear(); }
    {
        Bits<H> it = setB;
        while (it.any()) {
            const int b = it.pop_first();
            uint64_t mb[W];
#pragma unroll
            for (int w = 0; w < W; w++) mb[w] = sm.maskB[b * W + w];
#pragma unroll
            for (int h = 0; h < H; h++) {
                bool disjoint = true, asub = true, bsub = true;
#pragma unroll
                for (int w = 0; w < W; w++) {
                    const uint64_t x = mA[h][w] & mb[w];
                    disjoint &= (x == 0); asub &= (x == mA[h][w]); bsub &= (x == mb[w]);
                }
                const bool incompatible = setA.mine(h, lane) && !(disjoint || asub || bsub);
                if (incompatible) inc[h].set(b);
                const uint32_t col = __ballot_sync(FULL, incompatible);     // column b over A, half h
#pragma unroll
                for (int hb = 0; hb < H; hb++) if ((b >> 5) == hb && lane == (b & 31)) incT[hb].w[h] = col;
            }
        }
    }
    // ---- 4. extended common-edge rule ------------------------------------------------------------------------
#pragma unroll
    for (int h = 0; h < H; h++) {
        const bool extA = setA.mine(h, lane) && !inc[h].any();
        const bool extB = setB.mine(h, lane) && !incT[h].any();
        if (extA) part += la2[h];
        if (extB) part += lb2[h];
        setA.w[h] &= ~__ballot_sync(FULL, extA); setB.w[h] &= ~__ballot_sync(FULL, extB);
    }

    // ---- 5. GTP ----------------------------------------------------------------------------------------------
    double total = 0.0;
    int sp = 0;
    if (setA.any() && setB.any()) {
        if (lane == 0) {
#pragma unroll
            for (int h = 0; h < H; h++) { sm.stack[0][h] = setA.w[h]; sm.stack[0][H + h] = setB.w[h]; } }
        sp = 1;
    }
    __syncwarp();
    int guard = 0;
    while (sp > 0) {
        sp--;
        Bits<H> RA, RB;
#pragma unroll
        for (int h = 0; h < H; h++) { RA.w[h] = sm.stack[sp][h]; RB.w[h] = sm.stack[sp][H + h]; }
        __syncwarp();
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int h = 0; h < H; h++) { if (RA.mine(h, lane)) sa += la2[h]; if (RB.mine(h, lane)) sb += lb2[h]; }
        sa = warp_sum(sa); sb = warp_sum(sb);
        bool split = false;
        Bits<H> CA, CB; CA.clear(); CB.clear();
        if (RA.count() + RB.count() > 2 && sa > 0.0 && sb > 0.0) {
            st.flows++;
            // vertex weights -> residual capacities of the source / sink edges
            double ra[H], rb[H], wa[H], wb[H];
            Bits<H> fpos[H];                    // my row a of "flow[a][b] > FLOW_TOL", kept in step with every write of the flow matrix
#pragma unroll
            for (int h = 0; h < H; h++) {
                fpos[h].clear();
                wa[h] = RA.mine(h, lane) ? la2[h] / sa : 0.0;
                wb[h] = RB.mine(h, lane) ? lb2[h] / sb : 0.0;
                ra[h] = wa[h]; rb[h] = wb[h];
            }
            __syncwarp();
            // greedy saturation: push every source edge into its neighbours' sink edges, in split order
            {
                Bits<H> it = RA;
                while (it.any()) {
                    const int a = it.pop_first();
                    double x = fetch_d<H>(ra, a);
                    const Bits<H> nbm = fetch_bits<H>(inc, a) & RB;
#pragma unroll
                    for (int h = 0; h < H; h++) {
                        if (!(x > FLOW_TOL) || nbm.w[h] == 0) continue;       // warp-uniform
                        // the neighbours that still have sink capacity take x one after the other, in split order, until it is used up
                        // (a source is mostly absorbed by its first one or two; the former warp scan over all 32 lanes: config 5 35.7 ms, this form 31.5)
                        uint32_t live = __ballot_sync(FULL, ((nbm.w[h] >> lane) & 1u) && rb[h] > 0.0);
                        bool carry = false;
                        while (live && x > 0.0) {
                            const int src = __ffs(live) - 1; live &= live - 1;
                            const double take = fmin(__shfl_sync(FULL, rb[h], src), x);
                            if (lane == src) { rb[h] -= take; flow[a * FS + lane + 32 * h] = take; carry = take > FLOW_TOL; }
                            x -= take;
                        }
                        const uint32_t carries = __ballot_sync(FULL, carry);
#pragma unroll
                        for (int ha = 0; ha < H; ha++) if ((a >> 5) == ha && lane == (a & 31)) fpos[ha].w[h] = carries;
                    }
                    owner_store<H>(ra, a, lane, x);
                }
            }
            __syncwarp();
            // augmenting paths by BFS over warp-uniform frontier bitmasks
            Bits<H> visA, visB;
            for (;;) {
                if (++guard > 20000) { failed = true; break; }
                int parA[H], parB[H];
                Bits<H> frontA;
#pragma unroll
                for (int h = 0; h < H; h++) { parA[h] = -1; parB[h] = -1; frontA.w[h] = __ballot_sync(FULL, ra[h] > FLOW_TOL); }
                visA = frontA; visB.clear();
                int sink = -1;
                while (frontA.any()) {
                    Bits<H> newB; bool mine_sink[H];
#pragma unroll
                    for (int h = 0; h < H; h++) {
                        const Bits<H> reach = incT[h] & frontA;
                        const bool nb = RB.mine(h, lane) && !visB.mine(h, lane) && reach.any();
                        if (nb) parB[h] = reach.first();
                        newB.w[h] = __ballot_sync(FULL, nb);
                        mine_sink[h] = nb && rb[h] > FLOW_TOL;
                    }
                    if (!newB.any()) break;
#pragma unroll
                    for (int h = 0; h < H; h++) visB.w[h] |= newB.w[h];
                    Bits<H> sinks;
#pragma unroll
                    for (int h = 0; h < H; h++) sinks.w[h] = __ballot_sync(FULL, mine_sink[h]);
                    if (sinks.any()) { sink = sinks.first(); break; }
                    // B -> A through edges that carry flow (fpos: no walk over the candidates, no shared-memory reads)
                    Bits<H> newA;
#pragma unroll
                    for (int h = 0; h < H; h++) {
                        const Bits<H> cand = fpos[h] & newB;
                        const bool found = RA.mine(h, lane) && !visA.mine(h, lane) && cand.any();
                        if (found) parA[h] = cand.first();
                        newA.w[h] = __ballot_sync(FULL, found);
                    }
#pragma unroll
                    for (int h = 0; h < H; h++) visA.w[h] |= newA.w[h];
                    frontA = newA;
                }
                if (sink < 0) break;            // no augmenting path: (visA, visB) is the source side of a min cut
                st.aug++;
                // bottleneck along sink <- ... <- source
                double delta = fetch_d<H>(rb, sink);
                int b = sink;
                for (int hop = 0; hop < 2 * CAP + 2; hop++) {
                    const int a = fetch_i<H>(parB, b); const int pb = fetch_i<H>(parA, a);
                    if (pb < 0) { delta = fmin(delta, fetch_d<H>(ra, a)); break; }
                    delta = fmin(delta, flow[a * FS + pb]); b = pb;
                }
                owner_sub<H>(rb, sink, lane, delta);
                b = sink;
                for (int hop = 0; hop < 2 * CAP + 2; hop++) {
                    const int a = fetch_i<H>(parB, b); const int pb = fetch_i<H>(parA, a);
                    // the owner lane of row a updates the flow matrix and its bit row; a cell whose bit is clear counts as 0 (it may
                    // hold a value of an earlier max-flow: the matrix is never cleared)
                    const bool mine_a = lane == (a & 31);
#pragma unroll
                    for (int ha = 0; ha < H; ha++) if ((a >> 5) == ha && mine_a) {
                        const double fwd = (fpos[ha].test(b) ? flow[a * FS + b] : 0.0) + delta;
                        flow[a * FS + b] = fwd;
                        if (fwd > FLOW_TOL) fpos[ha].set(b); else fpos[ha].reset(b);
                        if (pb >= 0) {
                            const double back = flow[a * FS + pb] - delta;          // parA[a] = pb was taken from fpos: the cell is live
                            flow[a * FS + pb] = back;
                            if (!(back > FLOW_TOL)) fpos[ha].reset(pb);
                        }
                    }
                    if (pb < 0) { owner_sub<H>(ra, a, lane, delta); break; }
                    b = pb;
                }
                __syncwarp();
            }
            if (failed) break;
            // min-weight vertex cover = (A not reachable) + (B reachable)
            CA = andnot(RA, visA); CB = visB & RB;
            double cw = 0.0;
#pragma unroll
            for (int h = 0; h < H; h++) { if (CA.mine(h, lane)) cw += wa[h]; if (CB.mine(h, lane)) cw += wb[h]; }
            cw = warp_sum(cw);
            split = cw < 1.0 - SPLIT_TOL && CA.any() && !CA.equals(RA) && CB.any() && !CB.equals(RB);
        }
        if (split) {
            // (C&A, B\C) precedes (A\C, C&B) in the ratio sequence; the sum does not depend on the order
            if (lane == 0) {
#pragma unroll
                for (int h = 0; h < H; h++) {
                    sm.stack[sp][h] = RA.w[h] & ~CA.w[h]; sm.stack[sp][H + h] = CB.w[h];
                    sm.stack[sp + 1][h] = CA.w[h]; sm.stack[sp + 1][H + h] = RB.w[h] & ~CB.w[h];
                }
            }
            sp += 2;
            __syncwarp();
        } else {
            const double s = sqrt(sa) + sqrt(sb);
            total = fma(s, s, total); st.ratios++;
        }
    }
    return total;
}

}  // namespace geo
}  // namespace td
