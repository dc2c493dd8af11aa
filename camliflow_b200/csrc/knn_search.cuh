// Warp-cooperative exact k-NN search shared by every kernel that needs neighbours
// (index output, three-NN interpolation / back-warping, point-correlation lookup).
//
// ONE WARP owns a query; the candidate cloud is staged tile by tile in shared memory (SoA,
// coalesced global reads, one copy per CTA shared by its 8 queries) and scanned 128 candidates
// per step (4 conflict-free LDS per coordinate and lane, all independent).  The sorted result
// list is distributed over the lanes' registers (entry e lives in lane e%32, slot e/32), so an
// insertion is one ballot + one shuffle-shift.  Candidates are consumed in ascending input order
// and every insertion applies the reference's rule literally (skip if d > worst; place after all
// entries <= d, scanning down from j = min(idx, k-1); the last entry is dropped:
// k_nearest_neighbor_kernel.cu:80-90), so the result is bit-identical to the reference's
// one-thread-per-query insertion sort, including on exact distance ties.
#pragma once
#include "common.cuh"

constexpr float KNN_INIT_DIST = 1e9f;   // k_nearest_neighbor_kernel.cu:72
constexpr int KNN_TILE = 2048;          // candidates staged per tile (24 KB as 3 x float[2048])
constexpr int KNN_WARPS = 8;            // queries (warps) per CTA

// Element strides of a [B, points, D] view: channel-last [B,N,D] and channel-first [B,D,N]
// tensors are both read in place.
struct KnnView { long long sb, sp, sd; };

// Candidate sources: point i of the cloud -> (x, y, z).
template <int D>
struct KnnPlainPoints {
    const float* __restrict__ base;   // batch-offset pointer
    long long sp, sd;
    __device__ __forceinline__ void load(int i, float& x, float& y, float& z) const {
        const float* p = base + i * sp;
        x = __ldg(p); y = __ldg(p + sd); z = (D == 3) ? __ldg(p + 2 * sd) : 0.f;
    }
};

// Candidates = base points displaced by a per-point vector (xyz1 + flow of backwarp_3d,
// models/utils.py:156): the sum is rounded to fp32 exactly like the reference's tensor add.
struct KnnDisplacedPoints {
    const float* __restrict__ base;   // [3, m] channel-first
    const float* __restrict__ disp;   // same layout
    long long sd;
    __device__ __forceinline__ void load(int i, float& x, float& y, float& z) const {
        x = __fadd_rn(__ldg(base + i), __ldg(disp + i));
        y = __fadd_rn(__ldg(base + sd + i), __ldg(disp + sd + i));
        z = __fadd_rn(__ldg(base + 2 * sd + i), __ldg(disp + 2 * sd + i));
    }
};

struct KnnTile {
    float x[KNN_TILE], y[KNN_TILE], z[KNN_TILE];
};

// Sorted result list of one warp: entry e = lane (slot 0) and e = 32 + lane (slot 1).
struct KnnList {
    float d0, d1;
    int i0, i1;
    float worst;                          // distance held by entry k-1
    __device__ __forceinline__ void init() {
        d0 = d1 = worst = KNN_INIT_DIST;
        i0 = i1 = 0;
    }
};

// All threads of the CTA: stage candidates [tile0, tile0+cnt) into shared memory.
template <typename Points>
__device__ __forceinline__ void knn_stage_tile(KnnTile& t, const Points& pts, int tile0, int cnt) {
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) pts.load(tile0 + i, t.x[i], t.y[i], t.z[i]);
}

// One block of 32 candidates (lane's candidate has global index base + lane) in ascending order.
// SLOTS = 1 handles k <= 32, SLOTS = 2 handles k <= 64.  All 32 lanes must call.
template <int SLOTS>
__device__ __forceinline__ void knn_insert_block(KnnList& L, float d, bool valid, int base, int k) {
    const int lane = threadIdx.x & 31;
    const int last = k - 1;
    unsigned pending = __ballot_sync(CAMLI_FULL_MASK, valid && !(d > L.worst));
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const float cd = __shfl_sync(CAMLI_FULL_MASK, d, src);
        if (cd > L.worst) continue;                     // list tightened since the ballot
        const int ci = base + src;
        const int limit = min(ci, last);                // reference: j = min(idx, k-1)
        // p = number of entries e < limit with dist <= cd  (list is sorted)
        int p = __popc(__ballot_sync(CAMLI_FULL_MASK, lane < limit && L.d0 <= cd));
        if (SLOTS == 2) p += __popc(__ballot_sync(CAMLI_FULL_MASK, lane + 32 < limit && L.d1 <= cd));
        // entries p < e <= limit take entry e-1, entry p takes the candidate
        const float up_d0 = __shfl_up_sync(CAMLI_FULL_MASK, L.d0, 1);
        const int up_i0 = __shfl_up_sync(CAMLI_FULL_MASK, L.i0, 1);
        if (SLOTS == 2) {
            float up_d1 = __shfl_up_sync(CAMLI_FULL_MASK, L.d1, 1);
            int up_i1 = __shfl_up_sync(CAMLI_FULL_MASK, L.i1, 1);
            const float wrap_d = __shfl_sync(CAMLI_FULL_MASK, L.d0, 31);
            const int wrap_i = __shfl_sync(CAMLI_FULL_MASK, L.i0, 31);
            if (lane == 0) { up_d1 = wrap_d; up_i1 = wrap_i; }
            const int e1 = lane + 32;
            if (e1 > p && e1 <= limit) { L.d1 = up_d1; L.i1 = up_i1; }
            else if (e1 == p) { L.d1 = cd; L.i1 = ci; }
        }
        if (lane > p && lane <= limit) { L.d0 = up_d0; L.i0 = up_i0; }
        else if (lane == p) { L.d0 = cd; L.i0 = ci; }
        L.worst = (SLOTS == 2 && last >= 32) ? __shfl_sync(CAMLI_FULL_MASK, L.d1, last - 32)
                                             : __shfl_sync(CAMLI_FULL_MASK, L.d0, last);
    }
}

// One warp: per-lane minimum distance over a staged tile (lane l sees candidates l, l+32, ...).
template <int D>
__device__ __forceinline__ float knn_tile_lane_min(const KnnTile& t, int cnt, float ux, float uy, float uz) {
    const int lane = threadIdx.x & 31;
    float m = INFINITY;
    for (int i = lane; i < cnt; i += 32) {
        const float d = (D == 3) ? camli_sqdist3(ux - t.x[i], uy - t.y[i], uz - t.z[i]) : camli_sqdist2(ux - t.x[i], uy - t.y[i]);
        m = fminf(m, d);
    }
    return m;
}

// The k-th smallest (1-based, k <= 32) of the 32 lane values, returned to every lane.
__device__ __forceinline__ float knn_kth_smallest_of_lanes(float v, int k) {
    const int lane = threadIdx.x & 31;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float o = __shfl_sync(CAMLI_FULL_MASK, v, j);
        rank += (o < v || (o == v && j < lane)) ? 1 : 0;
    }
    const unsigned hit = __ballot_sync(CAMLI_FULL_MASK, rank == k - 1);    // ranks are a permutation: exactly one lane
    return __shfl_sync(CAMLI_FULL_MASK, v, __ffs(hit) - 1);
}

// One warp: scan a staged tile for the query (ux,uy,uz); candidates farther than `tau` are not offered to the list.
template <int D, int SLOTS>
__device__ __forceinline__ void knn_scan_tile(KnnList& L, const KnnTile& t, int tile0, int cnt, int k,
                                              float ux, float uy, float uz, float tau = INFINITY) {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < cnt; base += 128) {
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = min(base + j * 32 + lane, cnt - 1);
            d[j] = (D == 3) ? camli_sqdist3(ux - t.x[i], uy - t.y[i], uz - t.z[i])
                            : camli_sqdist2(ux - t.x[i], uy - t.y[i]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = base + j * 32;
            if (b < cnt) knn_insert_block<SLOTS>(L, d[j], b + lane < cnt && !(d[j] > tau), tile0 + b, k);
        }
    }
}

// Whole search for a CTA of KNN_WARPS query-warps over a cloud of m candidates.  Every thread of
// the CTA must call (it contains __syncthreads); `active` = this warp has a real query.
//
// Threshold pre-pass.  The insertion sort costs ~k(1 + ln(m/k)) serial warp-collective insertions per query,
// almost all of them for candidates that are evicted again.  A first pass over the candidates gives every lane
// the minimum distance among ITS candidates; these 32 minima belong to 32 distinct candidates, so the k-th
// smallest of them, tau, bounds the final k-th distance from above.  The second pass offers only candidates
// with d <= tau to the (unchanged, reference-order) insertion routine.  Candidates with d > d_k never influence
// the final list (the result equals the replay of the points with d <= d_k alone, which holds for
// any candidate set containing them), so the output is bit-identical -- including ties at d_k, all of which
// pass d <= tau -- while the number of insertions drops ~4x (k = 16, m = 2048) to ~20x (k <= 3).
template <int D, int SLOTS, typename Points>
__device__ __forceinline__ void knn_cta_search(KnnList& L, KnnTile& tile, const Points& pts, int m, int k,
                                               bool active, float ux, float uy, float uz) {
    L.init();
    const bool prefilter = k <= 32 && m >= 128;          // CTA-uniform
    if (m <= KNN_TILE) {                                   // one tile: staged once, both passes read it
        knn_stage_tile(tile, pts, 0, m);
        __syncthreads();
        if (active) {
            const float tau = prefilter ? knn_kth_smallest_of_lanes(knn_tile_lane_min<D>(tile, m, ux, uy, uz), k) : INFINITY;
            knn_scan_tile<D, SLOTS>(L, tile, 0, m, k, ux, uy, uz, tau);
        }
        return;
    }
    float tau = INFINITY;
    if (prefilter) {
        float lane_min = INFINITY;
        for (int tile0 = 0; tile0 < m; tile0 += KNN_TILE) {
            const int cnt = min(KNN_TILE, m - tile0);
            if (tile0) __syncthreads();
            knn_stage_tile(tile, pts, tile0, cnt);
            __syncthreads();
            if (active) lane_min = fminf(lane_min, knn_tile_lane_min<D>(tile, cnt, ux, uy, uz));
        }
        if (active) tau = knn_kth_smallest_of_lanes(lane_min, k);
        __syncthreads();
    }
    for (int tile0 = 0; tile0 < m; tile0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, m - tile0);
        if (tile0) __syncthreads();
        knn_stage_tile(tile, pts, tile0, cnt);
        __syncthreads();
        if (active) knn_scan_tile<D, SLOTS>(L, tile, tile0, cnt, k, ux, uy, uz, tau);
    }
}
