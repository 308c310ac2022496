// kernels_geodesic.cu -- BHV geodesic distance by GTP (Owen & Provan 2011), one warp per tree pair (sm_100a).
//
// Replaces getGeodesicDistance(t1.phylotree, t2.phylotree, normalise) of treeCl/treedist.py:70,112 (arithmetic in
// the un-vendored PyPI package tree_distance, a C++ port of M. Owen's GTP) for every pair of the condensed matrix.
//
// Per pair, entirely inside one warp (registers + shared memory; the flow matrix in a per-warp slice of global memory when shared
// memory would limit residency, GeoPlan):
//   1. leaf term            sum_k (a_k - b_k)^2, lanes strided over taxa
//   2. common splits        binary search of each tree's sorted split ids in the other's  -> sum (li - lj)^2
//   3. incompatibility      bit matrix inc[a] (over the other tree's non-common splits), one row per lane
//   4. extended common rule splits compatible with every split of the other tree contribute l^2
//   5. GTP                  explicit stack of ratios (A-subset, B-subset bitmasks); for each ratio a min-weight
//                           vertex cover by max-flow: greedy saturation + BFS augmenting paths with the frontier
//                           sets as warp-uniform bitmasks; every lane keeps the bit row of its split's flow-carrying edges,
//                           so the fp64 flow matrix is touched only along the paths (geo_core.cuh)
//   distance^2 = sum_k (||A_k|| + ||B_k||)^2 + common + extended common + leaf
// The lane that owns split s of tree i (s = lane + 32 h) keeps its id / length / mask / residual capacity in
// registers; H = CAP / 32 splits per lane per tree.
#include <cuda_runtime.h>

#include <cstdint>

#include "geo_core.cuh"
#include "kernels.h"

namespace td {

namespace {

using namespace geo;

template <int H, int W>
__global__ void __launch_bounds__(GEO_WARPS * 32, geo_min_blocks(H)) geodesic_kernel(const GeoParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // per warp: fixed part, then the fp64 flow matrix sized for THIS set: flow_rows x FS doubles (FS odd: bank spread)
    const int FS = p.flow_rows | 1;
    const size_t flow_doubles = (size_t)p.flow_rows * FS;
    const size_t warp_bytes = sizeof(WarpSmem<H, W>) + (p.flow_scratch ? 0 : flow_doubles * sizeof(double));
    unsigned char *wbase = smem_raw + (size_t)(threadIdx.x >> 5) * warp_bytes;
    WarpSmem<H, W> &sm = *reinterpret_cast<WarpSmem<H, W> *>(wbase);
    double *flow = p.flow_scratch ? p.flow_scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * flow_doubles
                                  : reinterpret_cast<double *>(wbase + sizeof(WarpSmem<H, W>));
    const int lane = threadIdx.x & 31;
    Counters st;
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
        st.pairs++;

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

        // ---- 3-5. incompatibilities, extended common edges, GTP ratio sequence (geo_core.cuh) ---------------------
        const double total = gtp_noncommon<H, W>(sm, flow, FS, mA, setA, setB, la2, lb2, lane, part, st, failed);
        part = warp_sum(part);
        if (lane == 0) {
            double d = sqrt(part + total);
            if (p.normalise) { const double den = p.norm[ti] + p.norm[tj]; d = den != 0.0 ? d / den : 0.0; }
            if (failed) d = __longlong_as_double(0x7ff8000000000000ll);
            p.out[pidx - p.p_base] = d;
        }
        if (failed) { if (lane == 0) atomicOr(p.error_flag, 1); failed = false; }
    }
    if (lane == 0 && p.stats) {
        atomicAdd(&p.stats[0], st.pairs); atomicAdd(&p.stats[1], st.flows);
        atomicAdd(&p.stats[2], st.ratios); atomicAdd(&p.stats[3], st.aug);
    }
}

template <int H, int W>
cudaError_t plan_geo(const GeoParams &p, int n_sms, GeoPlan &plan)
{
    return plan_warp_kernel(geodesic_kernel<H, W>, sizeof(WarpSmem<H, W>), (size_t)p.flow_rows * (p.flow_rows | 1) * sizeof(double), p.p_end - p.p_begin, n_sms, plan);
}

template <int H, int W>
cudaError_t launch_geo(const GeoParams &p, const GeoPlan &plan, cudaStream_t st)
{
    auto k = geodesic_kernel<H, W>;
    if ((plan.scratch_bytes != 0) != (p.flow_scratch != nullptr)) return cudaErrorInvalidValue;
    cudaError_t e = allow_full_smem((const void *)k);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)plan.blocks, plan.warps * 32, plan.smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

#define GEO_DISPATCH(fn, ...)                                         \
    if (cap <= 32 && p.W == 1) return fn<1, 1>(__VA_ARGS__);          \
    if (cap <= 64 && p.W == 1) return fn<2, 1>(__VA_ARGS__);          \
    if (cap <= 64 && p.W == 2) return fn<2, 2>(__VA_ARGS__);          \
    if (cap <= 128 && p.W == 2) return fn<4, 2>(__VA_ARGS__);         \
    if (cap <= 128 && p.W == 3) return fn<4, 3>(__VA_ARGS__);         \
    return cudaErrorInvalidValue;

cudaError_t plan_geodesic(const GeoParams &p, int cap, int n_sms, GeoPlan &plan) { GEO_DISPATCH(plan_geo, p, n_sms, plan) }
cudaError_t launch_geodesic(const GeoParams &p, int cap, const GeoPlan &plan, cudaStream_t st) { GEO_DISPATCH(launch_geo, p, plan, st) }

}  // namespace td
