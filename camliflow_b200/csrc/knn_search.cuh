// Warp-cooperative exact k-NN search shared by every kernel that needs neighbours
// (index output, three-NN interpolation, point-correlation lookup).
//
// ONE WARP owns a query: the 32 lanes evaluate 32 candidates per step and the sorted result
// list is distributed over the lanes' registers (entry e lives in lane e%32, slot e/32), so an
// insertion is one ballot + one shuffle-shift.  Candidates are consumed in ascending input
// order and every insertion applies the reference's rule literally (skip if d > worst; place
// after all entries <= d, scanning down from j = min(idx, k-1); the last entry is dropped:
// k_nearest_neighbor_kernel.cu:80-90), so the result is bit-identical to the reference's
// one-thread-per-query insertion sort, including on exact distance ties.
#pragma once
#include "common.cuh"

constexpr float KNN_INIT_DIST = 1e9f;   // k_nearest_neighbor_kernel.cu:72

// Element strides of a [B, points, D] view: channel-last [B,N,D] and channel-first [B,D,N]
// tensors are both read in place.
struct KnnView { long long sb, sp, sd; };

// Functor giving candidate `i`'s squared distance to the query held in registers.
template <int D>
struct KnnPlainPoints {
    const float* __restrict__ base;   // batch-offset input pointer
    long long sp, sd;
    __device__ __forceinline__ float dist(int i, float ux, float uy, float uz) const {
        const float* p = base + i * sp;
        if (D == 3) return camli_sqdist3(ux - __ldg(p), uy - __ldg(p + sd), uz - __ldg(p + 2 * sd));
        return camli_sqdist2(ux - __ldg(p), uy - __ldg(p + sd));
    }
};

// Candidates = base points displaced by a per-point vector (xyz1 + flow of backwarp_3d,
// models/utils.py:156): the sum is rounded to fp32 exactly like the reference's tensor add.
struct KnnDisplacedPoints {
    const float* __restrict__ base;   // [.., m] channel-first xyz
    const float* __restrict__ disp;   // same layout
    long long sp, sd;
    __device__ __forceinline__ float dist(int i, float ux, float uy, float uz) const {
        const float* p = base + i * sp;
        const float* f = disp + i * sp;
        const float x = __fadd_rn(__ldg(p), __ldg(f));
        const float y = __fadd_rn(__ldg(p + sd), __ldg(f + sd));
        const float z = __fadd_rn(__ldg(p + 2 * sd), __ldg(f + 2 * sd));
        return camli_sqdist3(ux - x, uy - y, uz - z);
    }
};

// Result list of one warp: entry e = lane (slot 0) and e = 32 + lane (slot 1).
struct KnnList {
    float d0, d1;
    int i0, i1;
};

// SLOTS = 1 handles k <= 32, SLOTS = 2 handles k <= 64.  All 32 lanes must call.
template <int SLOTS, typename Points>
__device__ __forceinline__ KnnList knn_warp_search(const Points& pts, int m, int k, float ux, float uy, float uz) {
    const int lane = threadIdx.x & 31;
    float hd0 = KNN_INIT_DIST, hd1 = KNN_INIT_DIST;
    int hi0 = 0, hi1 = 0;
    float worst = KNN_INIT_DIST;          // distance held by entry k-1
    const int last = k - 1;

    float d_next = 0.f;                   // software prefetch of the next 32 candidates
    if (lane < m) d_next = pts.dist(lane, ux, uy, uz);

    for (int base = 0; base < m; base += 32) {
        const float d = d_next;
        const int nxt = base + 32 + lane;
        if (nxt < m) d_next = pts.dist(nxt, ux, uy, uz);

        const bool cand = (base + lane < m) && !(d > worst);
        unsigned pending = __ballot_sync(CAMLI_FULL_MASK, cand);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const float cd = __shfl_sync(CAMLI_FULL_MASK, d, src);
            if (cd > worst) continue;                       // list tightened since the ballot
            const int ci = base + src;
            const int limit = min(ci, last);                // reference: j = min(idx, k-1)
            // p = number of entries e < limit with dist <= cd  (list is sorted)
            int p = __popc(__ballot_sync(CAMLI_FULL_MASK, lane < limit && hd0 <= cd));
            if (SLOTS == 2)
                p += __popc(__ballot_sync(CAMLI_FULL_MASK, lane + 32 < limit && hd1 <= cd));
            // entries p < e <= limit take entry e-1, entry p takes the candidate
            const float up_d0 = __shfl_up_sync(CAMLI_FULL_MASK, hd0, 1);
            const int up_i0 = __shfl_up_sync(CAMLI_FULL_MASK, hi0, 1);
            if (SLOTS == 2) {
                float up_d1 = __shfl_up_sync(CAMLI_FULL_MASK, hd1, 1);
                int up_i1 = __shfl_up_sync(CAMLI_FULL_MASK, hi1, 1);
                const float wrap_d = __shfl_sync(CAMLI_FULL_MASK, hd0, 31);
                const int wrap_i = __shfl_sync(CAMLI_FULL_MASK, hi0, 31);
                if (lane == 0) { up_d1 = wrap_d; up_i1 = wrap_i; }
                const int e1 = lane + 32;
                if (e1 > p && e1 <= limit) { hd1 = up_d1; hi1 = up_i1; }
                else if (e1 == p) { hd1 = cd; hi1 = ci; }
            }
            if (lane > p && lane <= limit) { hd0 = up_d0; hi0 = up_i0; }
            else if (lane == p) { hd0 = cd; hi0 = ci; }
            worst = (SLOTS == 2 && last >= 32) ? __shfl_sync(CAMLI_FULL_MASK, hd1, last - 32)
                                               : __shfl_sync(CAMLI_FULL_MASK, hd0, last);
        }
    }
    KnnList r;
    r.d0 = hd0; r.d1 = hd1; r.i0 = hi0; r.i1 = hi1;
    return r;
}
