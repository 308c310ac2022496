// kernels_geodesic.cu -- BHV geodesic distance by GTP (Owen & Provan 2011), one warp per tree pair (sm_100a).
//
// Replaces getGeodesicDistance(t1.phylotree, t2.phylotree, normalise) of treeCl/treedist.py:70,112 (arithmetic in
// the un-vendored PyPI package tree_distance, a C++ port of M. Owen's GTP) for every pair of the condensed matrix.
//
// Per pair, entirely inside one warp (registers + shared memory, no global scratch):
//   1. leaf term            sum_k (a_k - b_k)^2, lanes strided over taxa
//   2. common splits        binary search of each tree's sorted split ids in the other's  -> sum (li - lj)^2
//   3. incompatibility      bit matrix inc[a] (over the other tree's non-common splits), one row per lane
//   4. extended common rule splits compatible with every split of the other tree contribute l^2
//   5. GTP                  explicit stack of ratios (A-subset, B-subset bitmasks); for each ratio a min-weight
//                           vertex cover by max-flow: greedy saturation + BFS augmenting paths with the frontier
//                           sets as warp-uniform bitmasks and the fp64 flow matrix in shared memory
//   distance^2 = sum_k (||A_k|| + ||B_k||)^2 + common + extended common + leaf
// The lane that owns split s of tree i (s = lane + 32 h) keeps its id / length / mask / residual capacity in
// registers; H = CAP / 32 splits per lane per tree.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace td {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int GEO_WARPS = 4;
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
    // i is warp-uniform; the word is picked by selects so that w[] stays in registers
    __device__ __forceinline__ bool test(int i) const { uint32_t x = w[0];
#pragma unroll
        for (int h = 1; h < H; h++) if ((i >> 5) == h) x = w[h]; return (x >> (i & 31)) & 1u; }
    __device__ __forceinline__ bool mine(int h, int lane) const { return (w[h] >> lane) & 1u; }   // h compile-time
    __device__ __forceinline__ int first() const {
        int r = -1;
#pragma unroll
        for (int h = H - 1; h >= 0; h--) if (w[h]) r = 32 * h + __ffs(w[h]) - 1;
        return r; }
    __device__ __forceinline__ void reset(int i) {
#pragma unroll
        for (int h = 0; h < H; h++) if ((i >> 5) == h) w[h] &= ~(1u << (i & 31)); }
    __device__ __forceinline__ void set(int i) {
#pragma unroll
        for (int h = 0; h < H; h++) if ((i >> 5) == h) w[h] |= 1u << (i & 31); }
    __device__ __forceinline__ bool equals(const Bits &o) const { bool e = true;
#pragma unroll
        for (int h = 0; h < H; h++) e &= (w[h] == o.w[h]); return e; }
};
template <int H> __device__ __forceinline__ Bits<H> operator&(const Bits<H> &a, const Bits<H> &b) { Bits<H> r; for (int h = 0; h < H; h++) r.w[h] = a.w[h] & b.w[h]; return r; }
template <int H> __device__ __forceinline__ Bits<H> andnot(const Bits<H> &a, const Bits<H> &b) { Bits<H> r; for (int h = 0; h < H; h++) r.w[h] = a.w[h] & ~b.w[h]; return r; }

// value owned by lane (idx & 31), slot (idx >> 5); idx is warp-uniform
template <int H>
__device__ __forceinline__ double fetch_d(const double (&v)[H], int idx)
{
    double r = __shfl_sync(FULL, v[0], idx & 31);
    if (H > 1) { const double r1 = __shfl_sync(FULL, v[H - 1], idx & 31); if (idx >> 5) r = r1; }
    return r;
}
template <int H>
__device__ __forceinline__ int fetch_i(const int (&v)[H], int idx)
{
    int r = __shfl_sync(FULL, v[0], idx & 31);
    if (H > 1) { const int r1 = __shfl_sync(FULL, v[H - 1], idx & 31); if (idx >> 5) r = r1; }
    return r;
}
template <int H>
__device__ __forceinline__ Bits<H> fetch_bits(const Bits<H> (&v)[H], int idx)
{
    Bits<H> r;
    for (int k = 0; k < H; k++) {
        uint32_t x = __shfl_sync(FULL, v[0].w[k], idx & 31);
        if (H > 1) { const uint32_t x1 = __shfl_sync(FULL, v[H - 1].w[k], idx & 31); if (idx >> 5) x = x1; }
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

template <int H, int W>
struct WarpSmem {
    static constexpr int CAP = 32 * H, FS = CAP + 1;
    double flow[CAP * FS];
    double lenB[CAP];
    uint64_t maskB[CAP * W];
    uint32_t idA[CAP], idB[CAP];
    uint32_t stack[CAP + 2][2 * H];
};

template <int H, int W>
__global__ void __launch_bounds__(GEO_WARPS * 32) geodesic_kernel(const GeoParams p)
{
    constexpr int CAP = 32 * H, FS = CAP + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSmem<H, W> &sm = reinterpret_cast<WarpSmem<H, W> *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    unsigned long long st_pairs = 0, st_flows = 0, st_ratios = 0, st_aug = 0;
    bool failed = false;

    for (;;) {
        unsigned long long wk = 0;
        if (lane == 0) wk = atomicAdd(p.work_counter, 1ull);
        wk = __shfl_sync(FULL, wk, 0);
        const int64_t pidx = p.p_begin + (int64_t)wk;
        if (pidx >= p.p_end) break;
        int64_t ti, tj;
        if (p.rect) { ti = pidx / p.rect_cols; tj = p.rect_split + pidx % p.rect_cols; }
        else {
            const double Nd = (double)p.N;
            int64_t i = (int64_t)floor(((2.0 * Nd - 1.0) - sqrt((2.0 * Nd - 1.0) * (2.0 * Nd - 1.0) - 8.0 * (double)pidx)) * 0.5);
            if (i < 0) i = 0; if (i > p.N - 2) i = p.N - 2;
            while (i + 1 <= p.N - 2 && tri_row_start(i + 1, p.N) <= pidx) i++;
            while (i > 0 && tri_row_start(i, p.N) > pidx) i--;
            ti = i; tj = pidx - tri_row_start(i, p.N) + i + 1;
        }
        st_pairs++;

        // ---- 1. leaf term + load both trees' split lists ---------------------------------------------------------
        double part = 0.0;     // per-lane partial of common + extended-common + leaf terms
        {
            const double *la = p.leaflen + ti * p.n_taxa, *lb = p.leaflen + tj * p.n_taxa;
            for (int k = lane; k < p.n_taxa; k += 32) { const double d = la[k] - lb[k]; part = fma(d, d, part); }
        }
        const int Si = p.nsplits[ti], Sj = p.nsplits[tj];
        uint32_t idA[H], idB[H]; double la2[H], lb2[H], lenA[H]; uint64_t mA[H][W];
        __syncwarp();
#pragma unroll
        for (int h = 0; h < H; h++) {
            const int s = lane + 32 * h;
            const bool va = s < Si, vb = s < Sj;
            idA[h] = va ? p.ids[ti * p.stride + s] : kSentinelId;
            idB[h] = vb ? p.ids[tj * p.stride + s] : kSentinelId;
            lenA[h] = va ? p.lens[ti * p.stride + s] : 0.0;
            const double lB = vb ? p.lens[tj * p.stride + s] : 0.0;
            la2[h] = lenA[h] * lenA[h]; lb2[h] = lB * lB;
#pragma unroll
            for (int w = 0; w < W; w++) {
                mA[h][w] = va ? p.masks[(ti * p.stride + s) * W + w] : 0ull;
                sm.maskB[s * W + w] = vb ? p.masks[(tj * p.stride + s) * W + w] : 0ull;
            }
            sm.idA[s] = idA[h]; sm.idB[s] = idB[h]; sm.lenB[s] = lB;
        }
        __syncwarp();

        // ---- 2. common splits ------------------------------------------------------------------------------------
        Bits<H> setA, setB;
#pragma unroll
        for (int h = 0; h < H; h++) {
            const int s = lane + 32 * h;
            bool ncA = false, ncB = false;
            if (s < Si) { const int pos = find_id(sm.idB, Sj, idA[h]); if (pos >= 0) { const double d = lenA[h] - sm.lenB[pos]; part = fma(d, d, part); } else ncA = true; }
            if (s < Sj) ncB = find_id(sm.idA, Si, idB[h]) < 0;
            setA.w[h] = __ballot_sync(FULL, ncA); setB.w[h] = __ballot_sync(FULL, ncB);
        }

        // ---- 3. incompatibility bit matrix: inc[h] row of my split a over B; incT[h] row of my split b over A ----
        Bits<H> inc[H], incT[H];
#pragma unroll
        for (int h = 0; h < H; h++) { inc[h].clear(); incT[h].clear(); }
        {
            Bits<H> it = setB;
            while (it.any()) {
                const int b = it.first(); it.reset(b);
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
            if (lane == 0) { for (int h = 0; h < H; h++) { sm.stack[0][h] = setA.w[h]; sm.stack[0][H + h] = setB.w[h]; } }
            sp = 1;
        }
        __syncwarp();
        int guard = 0;
        while (sp > 0) {
            sp--;
            Bits<H> RA, RB;
            for (int h = 0; h < H; h++) { RA.w[h] = sm.stack[sp][h]; RB.w[h] = sm.stack[sp][H + h]; }
            __syncwarp();
            double sa = 0.0, sb = 0.0;
#pragma unroll
            for (int h = 0; h < H; h++) { if (RA.mine(h, lane)) sa += la2[h]; if (RB.mine(h, lane)) sb += lb2[h]; }
            sa = warp_sum(sa); sb = warp_sum(sb);
            bool split = false;
            Bits<H> CA, CB; CA.clear(); CB.clear();
            if (RA.count() + RB.count() > 2 && sa > 0.0 && sb > 0.0) {
                st_flows++;
                // vertex weights -> residual capacities of the source / sink edges
                double ra[H], rb[H], wa[H], wb[H];
#pragma unroll
                for (int h = 0; h < H; h++) {
                    wa[h] = RA.mine(h, lane) ? la2[h] / sa : 0.0;
                    wb[h] = RB.mine(h, lane) ? lb2[h] / sb : 0.0;
                    ra[h] = wa[h]; rb[h] = wb[h];
                    if (RA.mine(h, lane)) {                             // zero my row of the flow matrix (only edges of the graph)
                        Bits<H> nb = inc[h] & RB;
                        while (nb.any()) { const int b = nb.first(); nb.reset(b); sm.flow[(lane + 32 * h) * FS + b] = 0.0; }
                    }
                }
                __syncwarp();
                // greedy saturation: push every source edge into its neighbours' sink edges, in split order
                {
                    Bits<H> it = RA;
                    while (it.any()) {
                        const int a = it.first(); it.reset(a);
                        double x = fetch_d<H>(ra, a);
                        const Bits<H> nbm = fetch_bits<H>(inc, a) & RB;
#pragma unroll
                        for (int h = 0; h < H; h++) {
                            if (!(x > FLOW_TOL) || nbm.w[h] == 0) continue;       // warp-uniform
                            const double avail = ((nbm.w[h] >> lane) & 1u) ? rb[h] : 0.0;
                            double incl = avail;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) { const double t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
                            const double tot = __shfl_sync(FULL, incl, 31);
                            const double take = fmin(avail, fmax(0.0, x - (incl - avail)));
                            if (take > 0.0) { rb[h] -= take; sm.flow[a * FS + lane + 32 * h] = take; }
                            x = fmax(0.0, x - tot);
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
                        // B -> A through edges that carry flow
                        Bits<H> newA;
#pragma unroll
                        for (int h = 0; h < H; h++) {
                            const int s = lane + 32 * h;
                            bool found = false;
                            if (RA.mine(h, lane) && !visA.mine(h, lane)) {
                                Bits<H> cand = inc[h] & newB;
                                while (cand.any()) { const int b = cand.first(); cand.reset(b); if (sm.flow[s * FS + b] > FLOW_TOL) { parA[h] = b; found = true; break; } }
                            }
                            newA.w[h] = __ballot_sync(FULL, found);
                        }
#pragma unroll
                        for (int h = 0; h < H; h++) visA.w[h] |= newA.w[h];
                        frontA = newA;
                    }
                    if (sink < 0) break;            // no augmenting path: (visA, visB) is the source side of a min cut
                    st_aug++;
                    // bottleneck along sink <- ... <- source
                    double delta = fetch_d<H>(rb, sink);
                    int b = sink;
                    for (int hop = 0; hop < 2 * CAP + 2; hop++) {
                        const int a = fetch_i<H>(parB, b); const int pb = fetch_i<H>(parA, a);
                        if (pb < 0) { delta = fmin(delta, fetch_d<H>(ra, a)); break; }
                        delta = fmin(delta, sm.flow[a * FS + pb]); b = pb;
                    }
                    owner_sub<H>(rb, sink, lane, delta);
                    b = sink;
                    for (int hop = 0; hop < 2 * CAP + 2; hop++) {
                        const int a = fetch_i<H>(parB, b); const int pb = fetch_i<H>(parA, a);
                        if (lane == 0) sm.flow[a * FS + b] += delta;
                        if (pb < 0) { owner_sub<H>(ra, a, lane, delta); break; }
                        if (lane == 0) sm.flow[a * FS + pb] -= delta;
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
                    for (int h = 0; h < H; h++) {
                        sm.stack[sp][h] = RA.w[h] & ~CA.w[h]; sm.stack[sp][H + h] = CB.w[h];
                        sm.stack[sp + 1][h] = CA.w[h]; sm.stack[sp + 1][H + h] = RB.w[h] & ~CB.w[h];
                    }
                }
                sp += 2;
                __syncwarp();
            } else {
                const double s = sqrt(sa) + sqrt(sb);
                total = fma(s, s, total); st_ratios++;
            }
        }
        part = warp_sum(part);
        if (lane == 0) {
            double d = sqrt(part + total);
            if (p.normalise) { const double den = p.norm[ti] + p.norm[tj]; d = den != 0.0 ? d / den : 0.0; }
            if (failed) d = __longlong_as_double(0x7ff8000000000000ll);
            p.out[pidx - p.p_base] = d;
        }
        if (failed) { if (lane == 0) atomicExch(p.error_flag, 1); failed = false; }
    }
    if (lane == 0 && p.stats) {
        atomicAdd(&p.stats[0], st_pairs); atomicAdd(&p.stats[1], st_flows);
        atomicAdd(&p.stats[2], st_ratios); atomicAdd(&p.stats[3], st_aug);
    }
}

template <int H, int W>
cudaError_t launch_geo(const GeoParams &p, int n_sms, cudaStream_t st)
{
    auto k = geodesic_kernel<H, W>;
    const size_t smem = sizeof(WarpSmem<H, W>) * GEO_WARPS;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, GEO_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int64_t pairs = p.p_end - p.p_begin;
    int64_t blocks = (int64_t)n_sms * per_sm;
    const int64_t need = (pairs + GEO_WARPS - 1) / GEO_WARPS;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    k<<<(unsigned)blocks, GEO_WARPS * 32, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

size_t geodesic_smem_bytes(int cap, int W, int warps)
{
    if (cap <= 32) return sizeof(WarpSmem<1, 1>) * warps;
    return (W == 1 ? sizeof(WarpSmem<2, 1>) : sizeof(WarpSmem<2, 2>)) * warps;
}

cudaError_t launch_geodesic(const GeoParams &p, int cap, int n_sms, cudaStream_t st)
{
    if (cap <= 32 && p.W == 1) return launch_geo<1, 1>(p, n_sms, st);
    if (cap <= 64 && p.W == 1) return launch_geo<2, 1>(p, n_sms, st);
    if (cap <= 64 && p.W == 2) return launch_geo<2, 2>(p, n_sms, st);
    return cudaErrorInvalidValue;
}

}  // namespace td
