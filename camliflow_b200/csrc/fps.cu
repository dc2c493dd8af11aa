// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel<1024> (reference
// models/csrc/furthest_point_sampling/furthest_point_sampling_kernel.cu:34-84).
//
// The S rounds are strictly dependent, so the whole cost is the latency of one
// round.  Design: one 1024-thread CTA per cloud; every thread keeps its points
// AND their running min-distances in registers for all S rounds (the reference
// re-reads xyz and round-trips `dists_temp` through global memory every round);
// the cloud is also staged once in shared memory as float4 so the current
// sample's coordinates are one broadcast LDS.128; the block argmax is two
// `redux.sync` levels with ONE __syncthreads per round (double-buffered slots)
// instead of the reference's 10-stage shared-memory tree with 6 barriers.
//
// Tie rule (bit-exact with the reference tree, which takes the right operand
// on `<=`): winner = max over points of (dist, bitrev10(i & 1023), -i).
#include <cooperative_groups.h>
#include <cub/block/block_radix_sort.cuh>
#include "common.cuh"

namespace {

constexpr int FPS_THREADS = 1024;
constexpr int FPS_WARPS = FPS_THREADS / 32;
constexpr int FPS_MAX_REG_POINTS = 8192;  // 8 points per thread in registers

// Picks the lane holding the lexicographic max of (key, tie); key compared as
// signed int (float bits of values >= 0, or of the negative "no point" marker).
__device__ __forceinline__ int fps_pick_lane(int key, unsigned tie) {
    const int kmax = __reduce_max_sync(CAMLI_FULL_MASK, key);
    const unsigned match = __ballot_sync(CAMLI_FULL_MASK, key == kmax);
    if (__popc(match) == 1) return __ffs(match) - 1;
    const unsigned t = (key == kmax) ? tie + 1u : 0u;
    const unsigned tmax = __reduce_max_sync(CAMLI_FULL_MASK, t);
    return __ffs(__ballot_sync(CAMLI_FULL_MASK, t == tmax)) - 1;
}

// Block-wide argmax of (best_d, brev, -best_i); returns the winning point index
// to every thread.  `slots` is double-buffered by round parity, which makes one
// barrier per round sufficient.
__device__ __forceinline__ int fps_block_argmax(float best_d, int best_i, unsigned brev,
                                                int4 (*slots)[FPS_WARPS], int parity,
                                                int warp, int lane) {
    const int key = __float_as_int(best_d);
    const int src = fps_pick_lane(key, brev);
    if (lane == src) slots[parity][warp] = make_int4(key, (int)brev, best_i, 0);
    __syncthreads();
    const int4 s = slots[parity][lane];
    const int src2 = fps_pick_lane(s.x, (unsigned)s.y);
    return __shfl_sync(CAMLI_FULL_MASK, s.z, src2);
}

template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_register_kernel(const float* __restrict__ xyz_all, int N, int S, int64_t* __restrict__ out_all) {
    extern __shared__ float4 s_pts[];          // [N]
    __shared__ int4 s_slots[2][FPS_WARPS];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* __restrict__ xyz = xyz_all + (size_t)blockIdx.x * N * 3;
    int64_t* __restrict__ out = out_all + (size_t)blockIdx.x * S;

    float px[PPT], py[PPT], pz[PPT], pd[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int i = t + j * FPS_THREADS;
        if (i < N) {
            px[j] = xyz[i * 3 + 0]; py[j] = xyz[i * 3 + 1]; pz[j] = xyz[i * 3 + 2];
            pd[j] = 1e10f;                       // furthest_point_sampling.cpp:12
            s_pts[i] = make_float4(px[j], py[j], pz[j], 0.f);
        } else {
            px[j] = py[j] = pz[j] = 0.f;
            pd[j] = -2.f;                        // can never beat the -1 "no point" marker
        }
    }
    const unsigned brev = __brev((unsigned)t) >> 22;   // 10-bit reversal of the owning thread id
    __syncthreads();

    int cur = 0;
    for (int s = 0; s < S; ++s) {
        if (t == 0) out[s] = (int64_t)cur;
        if (s == S - 1) break;
        const float4 c = s_pts[cur];
        float best_d = -1.f;
        int best_j = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = camli_sqdist3(px[j] - c.x, py[j] - c.y, pz[j] - c.z);
            const float nd = fminf(pd[j], d);
            pd[j] = nd;
            if (nd > best_d) { best_d = nd; best_j = j; }
        }
        const int best_i = (best_d < 0.f) ? 0 : t + best_j * FPS_THREADS;
        cur = fps_block_argmax(best_d, best_i, brev, s_slots, s & 1, warp, lane);
    }
}

// Any N: points stream from global memory (L1/L2 resident), running distances
// live in the caller's scratch buffer.  Same arithmetic, same tie rule.
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_streaming_kernel(const float* __restrict__ xyz_all, float* __restrict__ dist_all,
                     int N, int S, int64_t* __restrict__ out_all) {
    __shared__ int4 s_slots[2][FPS_WARPS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* __restrict__ xyz = xyz_all + (size_t)blockIdx.x * N * 3;
    float* __restrict__ dist = dist_all + (size_t)blockIdx.x * N;
    int64_t* __restrict__ out = out_all + (size_t)blockIdx.x * S;

    for (int i = t; i < N; i += FPS_THREADS) dist[i] = 1e10f;   // same thread re-reads it
    const unsigned brev = __brev((unsigned)t) >> 22;

    int cur = 0;
    for (int s = 0; s < S; ++s) {
        if (t == 0) out[s] = (int64_t)cur;
        if (s == S - 1) break;
        const float cx = __ldg(xyz + (size_t)cur * 3 + 0);
        const float cy = __ldg(xyz + (size_t)cur * 3 + 1);
        const float cz = __ldg(xyz + (size_t)cur * 3 + 2);
        float best_d = -1.f;
        int best_i = 0;
        for (int i = t; i < N; i += FPS_THREADS) {
            const float d = camli_sqdist3(__ldg(xyz + (size_t)i * 3 + 0) - cx,
                                          __ldg(xyz + (size_t)i * 3 + 1) - cy,
                                          __ldg(xyz + (size_t)i * 3 + 2) - cz);
            const float nd = fminf(dist[i], d);
            dist[i] = nd;
            if (nd > best_d) { best_d = nd; best_i = i; }
        }
        cur = fps_block_argmax(best_d, best_i, brev, s_slots, s & 1, warp, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// Pruned single-CTA path (2048 < N <= 8192, default).  A round of the kernels above touches every point
// although a new sample can only lower the running distance of points closer to it than that distance:
// late in the sampling that is a handful of points around the sample.  Here the cloud is first sorted along a
// Morton curve inside the kernel (cub block radix sort of 21-bit cell codes), so every warp owns a compact
// bucket of 32*PPT points with a bounding box and a cached (max distance, tie key) record.  Per round a warp
// compares the sample with its box: if even the box is no closer than the bucket's largest running distance,
// none of its distances can change and the cached record stands -- the warp goes straight to the barrier.
// The test is exact, not approximate: rounding is monotonic, so for a point inside the box every term of
// camli_sqdist3 is >= the same term of the box distance computed by the same function, hence d >= lb >= pd and
// fminf(pd, d) == pd bit for bit.  Typically 1..4 of the 32 warps do arithmetic in a round; the others only
// take part in the one barrier and the 32-record pick.  No cross-SM exchange at all: ~2x the round rate of the
// 8-CTA cluster on 1/8 of the SMs.
// Winner and tie rule are carried by the same explicit 64-bit key as in the cluster kernels below
// (distance bits, then fps_key_lo(original index)), so the order of the points inside the kernel is irrelevant.
__device__ __forceinline__ unsigned fps_key_lo(int i) {
    return ((__brev((unsigned)i & 1023u) >> 22) << 22) | (0x3FFFFFu - (unsigned)i);
}

__device__ __forceinline__ unsigned fps_spread7(unsigned v) {      // 7 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ float fps_warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(CAMLI_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float fps_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(CAMLI_FULL_MASK, v, o));
    return v;
}

template <int PPT>
struct FpsPruned {
    using Sort = cub::BlockRadixSort<unsigned, FPS_THREADS, PPT, unsigned>;
    static constexpr int kMaxN = FPS_THREADS * PPT;
    static constexpr size_t kXyzBytes = (size_t)3 * kMaxN * sizeof(float);
    static constexpr size_t kLoBytes = (size_t)kMaxN * sizeof(unsigned);
    static constexpr size_t kSmem = kXyzBytes + kLoBytes + sizeof(typename Sort::TempStorage) + 16;
};

template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_pruned_kernel(const float* __restrict__ xyz_all, int N, int S, int64_t* __restrict__ out_all) {
    using K = FpsPruned<PPT>;
    extern __shared__ __align__(16) uint8_t fpsp_smem[];
    float* sx = reinterpret_cast<float*>(fpsp_smem);              // the cloud by ORIGINAL index (the sample's coordinates)
    float* sy = sx + K::kMaxN;
    float* sz = sy + K::kMaxN;
    unsigned* s_lo = reinterpret_cast<unsigned*>(fpsp_smem + K::kXyzBytes);     // tie key of register point j of thread t: [j][t]
    typename K::Sort::TempStorage& sort_tmp =
        *reinterpret_cast<typename K::Sort::TempStorage*>(fpsp_smem + K::kXyzBytes + K::kLoBytes);
    __shared__ uint2 s_slots[2][FPS_WARPS];
    __shared__ float s_box[6][FPS_WARPS];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* __restrict__ xyz = xyz_all + (size_t)blockIdx.x * N * 3;
    int64_t* __restrict__ out = out_all + (size_t)blockIdx.x * S;

    // ---- the cloud -> shared memory, its bounding box
    const float INF = __int_as_float(0x7f800000);
    float lo_x = INF, lo_y = INF, lo_z = INF, hi_x = -INF, hi_y = -INF, hi_z = -INF;
    for (int i = t; i < N; i += FPS_THREADS) {
        const float x = __ldg(xyz + i * 3 + 0), y = __ldg(xyz + i * 3 + 1), z = __ldg(xyz + i * 3 + 2);
        sx[i] = x; sy[i] = y; sz[i] = z;
        lo_x = fminf(lo_x, x); hi_x = fmaxf(hi_x, x);
        lo_y = fminf(lo_y, y); hi_y = fmaxf(hi_y, y);
        lo_z = fminf(lo_z, z); hi_z = fmaxf(hi_z, z);
    }
    lo_x = fps_warp_min(lo_x); lo_y = fps_warp_min(lo_y); lo_z = fps_warp_min(lo_z);
    hi_x = fps_warp_max(hi_x); hi_y = fps_warp_max(hi_y); hi_z = fps_warp_max(hi_z);
    if (lane == 0) {
        s_box[0][warp] = lo_x; s_box[1][warp] = lo_y; s_box[2][warp] = lo_z;
        s_box[3][warp] = hi_x; s_box[4][warp] = hi_y; s_box[5][warp] = hi_z;
    }
    __syncthreads();
    lo_x = fps_warp_min(s_box[0][lane]); lo_y = fps_warp_min(s_box[1][lane]); lo_z = fps_warp_min(s_box[2][lane]);
    hi_x = fps_warp_max(s_box[3][lane]); hi_y = fps_warp_max(s_box[4][lane]); hi_z = fps_warp_max(s_box[5][lane]);
    const float extent = fmaxf(fmaxf(hi_x - lo_x, hi_y - lo_y), hi_z - lo_z);
    const float cell = (extent > 0.f && extent < INF) ? 127.999f / extent : 0.f;     // cubic cells; only the bucket quality depends on it

    // ---- Morton order: sorted position p = t * PPT + j lives in register j of thread t
    unsigned key[PPT], val[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int i = t + j * FPS_THREADS;
        val[j] = (unsigned)i;
        key[j] = 0x1FFFFFu;                                        // padding sorts last
        if (i < N) {
            const unsigned qx = (unsigned)min(127, max(0, (int)((sx[i] - lo_x) * cell)));
            const unsigned qy = (unsigned)min(127, max(0, (int)((sy[i] - lo_y) * cell)));
            const unsigned qz = (unsigned)min(127, max(0, (int)((sz[i] - lo_z) * cell)));
            key[j] = fps_spread7(qx) | (fps_spread7(qy) << 1) | (fps_spread7(qz) << 2);
        }
    }
    typename K::Sort(sort_tmp).Sort(key, val, 0, 21);

    float px[PPT], py[PPT], pz[PPT], pd[PPT];
    float b_lo_x = INF, b_lo_y = INF, b_lo_z = INF, b_hi_x = -INF, b_hi_y = -INF, b_hi_z = -INF;    // this warp's bucket
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int i = (int)val[j];
        if (i < N) {
            px[j] = sx[i]; py[j] = sy[i]; pz[j] = sz[i];
            pd[j] = 1e10f;                                        // furthest_point_sampling.cpp:12
            s_lo[j * FPS_THREADS + t] = fps_key_lo(i);
            b_lo_x = fminf(b_lo_x, px[j]); b_hi_x = fmaxf(b_hi_x, px[j]);
            b_lo_y = fminf(b_lo_y, py[j]); b_hi_y = fmaxf(b_hi_y, py[j]);
            b_lo_z = fminf(b_lo_z, pz[j]); b_hi_z = fmaxf(b_hi_z, pz[j]);
        } else {
            px[j] = py[j] = pz[j] = 0.f;
            pd[j] = -2.f;                                         // never beats a real point
            s_lo[j * FPS_THREADS + t] = 0u;
        }
    }
    b_lo_x = fps_warp_min(b_lo_x); b_lo_y = fps_warp_min(b_lo_y); b_lo_z = fps_warp_min(b_lo_z);
    b_hi_x = fps_warp_max(b_hi_x); b_hi_y = fps_warp_max(b_hi_y); b_hi_z = fps_warp_max(b_hi_z);

    // cached record of this warp's bucket: (max running distance, tie key); the initial value forces the first update.
    // A warp that only holds padding never competes.
    int w_hi = b_lo_x <= b_hi_x ? __float_as_int(1e10f) : __float_as_int(-3.f);
    unsigned w_lo = 0u;
    int cur = 0;
    for (int s = 0; s < S; ++s) {
        if (t == 0) out[s] = (int64_t)cur;
        if (s == S - 1) break;
        const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
        // distance from the sample to the bucket's box, by the same function and therefore <= every point's distance
        const float ex = fmaxf(fmaxf(b_lo_x - cx, cx - b_hi_x), 0.f);
        const float ey = fmaxf(fmaxf(b_lo_y - cy, cy - b_hi_y), 0.f);
        const float ez = fmaxf(fmaxf(b_lo_z - cz, cz - b_hi_z), 0.f);
        const float lb = camli_sqdist3(ex, ey, ez);
        if (!(lb >= __int_as_float(w_hi))) {                      // (warp-uniform) some running distance may drop
            float best_d = -3.f;
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
                const float d = camli_sqdist3(px[j] - cx, py[j] - cy, pz[j] - cz);
                const float nd = fminf(pd[j], d);
                pd[j] = nd;
                best_d = fmaxf(best_d, nd);
            }
            const int hi = __float_as_int(best_d);
            w_hi = __reduce_max_sync(CAMLI_FULL_MASK, hi);
            unsigned best_lo = 0u;
            if (hi == w_hi) {
#pragma unroll
                for (int j = 0; j < PPT; ++j)
                    if (pd[j] == best_d) best_lo = max(best_lo, s_lo[j * FPS_THREADS + t]);
            }
            w_lo = __reduce_max_sync(CAMLI_FULL_MASK, best_lo);
        }
        const int par = s & 1;
        if (lane == 0) s_slots[par][warp] = make_uint2((unsigned)w_hi, w_lo);
        __syncthreads();
        const uint2 r = s_slots[par][lane];
        const int gh = __reduce_max_sync(CAMLI_FULL_MASK, (int)r.x);
        const unsigned gl = __reduce_max_sync(CAMLI_FULL_MASK, (int)r.x == gh ? r.y : 0u);
        cur = (int)(0x3FFFFFu - (gl & 0x3FFFFFu));
    }
}

// ---------------------------------------------------------------------------------------------
// Cluster path (2048 < N <= 16384): the rounds of ONE cloud are spread over a thread-block cluster
// of 8 CTAs (8 SMs).  A single 1024-thread CTA is issue-bound at ~1 kcycle per round (N*10
// instructions through 4 schedulers); splitting the points 8 ways leaves ~100 cycles of math per
// round, plus one DSMEM exchange: every CTA reduces its slice to a (key, xyz) record, stores it
// into all 8 CTAs' shared memory (distributed shared memory), and after one cluster barrier every
// CTA picks the global winner locally.  Records are double-buffered by round parity so one
// barrier per round is enough.
//
// The tie rule is carried by an explicit 64-bit key: hi = float bits of the distance (>= 0, so
// integer order = float order), lo = bitrev10(i & 1023) << 22 | (0x3FFFFF - i); max over it =
// (largest distance, largest bit-reversed owner thread of the reference kernel, smallest i).
namespace cg = cooperative_groups;

constexpr int FPSC_CTAS = 8;
constexpr int FPSC_THREADS = 256;
constexpr int FPSC_WARPS = FPSC_THREADS / 32;
constexpr int FPSC_MAX_N = FPSC_CTAS * FPSC_THREADS * 8;   // 16384

struct __align__(16) FpsRecord { int hi; unsigned lo; float x, y, z; int pad0, pad1, pad2; };

template <int PPT>
__global__ void __cluster_dims__(FPSC_CTAS, 1, 1) __launch_bounds__(FPSC_THREADS, 1)
fps_cluster_kernel(const float* __restrict__ xyz_all, int N, int S, int64_t* __restrict__ out_all) {
    __shared__ float4 s_pts[PPT * FPSC_THREADS];                  // this CTA's slice
    __shared__ FpsRecord s_warp[FPSC_WARPS];
    __shared__ FpsRecord s_rec[2][FPSC_CTAS];                     // written by every CTA of the cluster

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / FPSC_CTAS;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* __restrict__ xyz = xyz_all + (size_t)cloud * N * 3;
    int64_t* __restrict__ out = out_all + (size_t)cloud * S;

    // point i = (j * CTAS + rank) * THREADS + t  (coalesced 256-point chunks, round-robin over CTAs)
    float px[PPT], py[PPT], pz[PPT], pd[PPT];
    unsigned plo[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int i = (j * FPSC_CTAS + rank) * FPSC_THREADS + t;
        if (i < N) {
            px[j] = xyz[i * 3 + 0]; py[j] = xyz[i * 3 + 1]; pz[j] = xyz[i * 3 + 2];
            pd[j] = 1e10f;                                        // furthest_point_sampling.cpp:12
            plo[j] = fps_key_lo(i);
        } else {
            px[j] = py[j] = pz[j] = 0.f;
            pd[j] = -2.f;                                         // never beats a real point
            plo[j] = 0u;
        }
        s_pts[j * FPSC_THREADS + t] = make_float4(px[j], py[j], pz[j], 0.f);
    }
    float cx = __ldg(xyz + 0), cy = __ldg(xyz + 1), cz = __ldg(xyz + 2);   // first sample: index 0
    int cur = 0;
    cluster.sync();                                               // all CTAs resident before any DSMEM store

    for (int s = 0; s < S; ++s) {
        if (rank == 0 && t == 0) out[s] = (int64_t)cur;
        if (s == S - 1) break;
        // ---- local slice: update running distances, thread-local best
        float best_d = -3.f;
        unsigned best_lo = 0u;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = camli_sqdist3(px[j] - cx, py[j] - cy, pz[j] - cz);
            const float nd = fminf(pd[j], d);
            pd[j] = nd;
            if (nd > best_d || (nd == best_d && plo[j] > best_lo)) { best_d = nd; best_lo = plo[j]; }
        }
        // ---- warp, then CTA
        int hi = __float_as_int(best_d);
        int hmax = __reduce_max_sync(CAMLI_FULL_MASK, hi);
        unsigned lmax = __reduce_max_sync(CAMLI_FULL_MASK, hi == hmax ? best_lo : 0u);
        if (hi == hmax && best_lo == lmax) { s_warp[warp].hi = hmax; s_warp[warp].lo = lmax; }
        __syncthreads();
        if (warp == 0) {
            const int whi = lane < FPSC_WARPS ? s_warp[lane].hi : (int)0x80000000;
            const unsigned wlo = lane < FPSC_WARPS ? s_warp[lane].lo : 0u;
            hmax = __reduce_max_sync(CAMLI_FULL_MASK, whi);
            lmax = __reduce_max_sync(CAMLI_FULL_MASK, whi == hmax ? wlo : 0u);
            if (lane < FPSC_CTAS) {
                FpsRecord r;
                r.hi = hmax; r.lo = lmax; r.pad0 = r.pad1 = r.pad2 = 0;
                r.x = r.y = r.z = 0.f;
                if (hmax >= 0) {                                  // this CTA owns at least one real point
                    const int i = (int)(0x3FFFFFu - (lmax & 0x3FFFFFu));
                    const int chunk = i / FPSC_THREADS;           // = j * CTAS + rank
                    const float4 p = s_pts[(chunk / FPSC_CTAS) * FPSC_THREADS + (i % FPSC_THREADS)];
                    r.x = p.x; r.y = p.y; r.z = p.z;
                }
                FpsRecord* dst = cluster.map_shared_rank(&s_rec[s & 1][rank], lane);
                *dst = r;
            }
        }
        cluster.sync();                                           // release/acquire: records visible cluster-wide
        // ---- every thread picks the winner among the 8 records (identical everywhere)
        int bh = (int)0x80000000;
        unsigned bl = 0u;
        int bi = 0;
#pragma unroll
        for (int c = 0; c < FPSC_CTAS; ++c) {
            const int h = s_rec[s & 1][c].hi;
            const unsigned l = s_rec[s & 1][c].lo;
            if (h > bh || (h == bh && l > bl)) { bh = h; bl = l; bi = c; }
        }
        cx = s_rec[s & 1][bi].x; cy = s_rec[s & 1][bi].y; cz = s_rec[s & 1][bi].z;
        cur = (int)(0x3FFFFFu - (bl & 0x3FFFFFu));
    }
    cluster.sync();                                               // no CTA exits while peers may still store to it
}

// ---------------------------------------------------------------------------------------------
// Async-exchange cluster path (default for 2048 < N <= 16384).  Same partition of one cloud over an
// 8-CTA cluster, but the per-round exchange costs one remote store instead of a shared-memory hop,
// a CTA barrier, a DSMEM store and a cluster barrier (~380 cycles + an L1 flush each round):
//   * every CTA keeps the WHOLE cloud in its shared memory (SoA, 12 B/point), so a record is just
//     the 64-bit key (distance bits, tie key) -- the winner's coordinates are a local LDS;
//   * every WARP reduces its points with two redux.sync and lanes 0..7 push the warp's key straight
//     into the 8 CTAs' record tables with `st.async ... mbarrier::complete_tx` -- the store itself
//     signals the destination's mbarrier, so there is no barrier instruction in the loop at all;
//   * every warp then waits on its own CTA's mbarrier (64 records x 8 B expected per round) and
//     picks the global winner from the 64 records (2 per lane + two redux.sync).
// Tables and mbarriers are double-buffered by round parity.  A CTA can run at most one round ahead
// of its peers (it needs all of round s's records to produce round s+1's), and a peer only sends
// round s+1 after it consumed round s-1's table, so the parity buffer is always free when written.
constexpr int FPSA_CTAS = 8;
constexpr int FPSA_THREADS = 256;
constexpr int FPSA_WARPS = FPSA_THREADS / 32;
constexpr int FPSA_RECORDS = FPSA_CTAS * FPSA_WARPS;        // 64 per round
constexpr int FPSA_MAX_N = FPSA_CTAS * FPSA_THREADS * 8;    // 16384 (196 KB of shared memory)
static_assert(FPSA_RECORDS == 64, "winner pick reads two records per lane");

__device__ __forceinline__ uint32_t fps_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Diagnostics: SM-clock stamps of rounds 100..103 as seen by thread 0 of CTA 0 (5 per round), when a buffer is attached.
__device__ long long* g_fps_timeline = nullptr;
#define FPS_STAMP(i) do { if (tl && s >= 100 && s < 104) tl[(s - 100) * 5 + (i)] = clock64(); } while (0)

template <int PPT>
__global__ void __cluster_dims__(FPSA_CTAS, 1, 1) __launch_bounds__(FPSA_THREADS, 1)
fps_cluster_async_kernel(const float* __restrict__ xyz_all, int N, int S, int64_t* __restrict__ out_all) {
    extern __shared__ float s_xyz[];                              // x[N] | y[N] | z[N]: the whole cloud
    __shared__ __align__(16) uint2 s_rec[2][FPSA_RECORDS];        // written by every warp of the cluster
    __shared__ __align__(8) uint64_t s_bar[2];

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / FPSA_CTAS;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* __restrict__ xyz = xyz_all + (size_t)cloud * N * 3;
    int64_t* __restrict__ out = out_all + (size_t)cloud * S;
    float* sx = s_xyz;
    float* sy = s_xyz + N;
    float* sz = s_xyz + 2 * N;

    for (int i = t; i < N; i += FPSA_THREADS) {
        sx[i] = __ldg(xyz + i * 3 + 0); sy[i] = __ldg(xyz + i * 3 + 1); sz[i] = __ldg(xyz + i * 3 + 2);
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(&s_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // point i = (j * CTAS + rank) * THREADS + t  (coalesced 256-point chunks, round-robin over CTAs)
    float px[PPT], py[PPT], pz[PPT], pd[PPT];
    unsigned plo[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int i = (j * FPSA_CTAS + rank) * FPSA_THREADS + t;
        if (i < N) {
            px[j] = sx[i]; py[j] = sy[i]; pz[j] = sz[i];
            pd[j] = 1e10f;                                        // furthest_point_sampling.cpp:12
            plo[j] = fps_key_lo(i);
        } else {
            px[j] = py[j] = pz[j] = 0.f;
            pd[j] = -2.f;                                         // never beats a real point
            plo[j] = 0u;
        }
    }
    // remote addresses of "my warp's" slot and of the mbarriers in CTA `lane` (lanes 0..7 send)
    uint32_t r_rec[2] = {0u, 0u}, r_bar[2] = {0u, 0u};
    if (lane < FPSA_CTAS) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                         : "=r"(r_rec[p]) : "r"(fps_smem_u32(&s_rec[p][rank * FPSA_WARPS + warp])), "r"(lane));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r_bar[p]) : "r"(fps_smem_u32(&s_bar[p])), "r"(lane));
        }
    }
    const uint32_t l_bar[2] = {fps_smem_u32(&s_bar[0]), fps_smem_u32(&s_bar[1])};
    long long* tl = (blockIdx.x == 0 && t == 0) ? g_fps_timeline : nullptr;
    cluster.sync();                                               // every CTA's barriers exist before any remote store

    int cur = 0;
    for (int s = 0; s < S; ++s) {
        if (rank == 0 && t == 0) out[s] = (int64_t)cur;
        if (s == S - 1) break;
        const int par = s & 1;
        FPS_STAMP(0);
        if (t == 0)                                               // arm this round's barrier (remote stores may already have landed)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         ::"r"(l_bar[par]), "r"((uint32_t)(FPSA_RECORDS * sizeof(uint2))) : "memory");
        const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
        // ---- this thread's points: update running distances, thread-local best
        float best_d = -3.f;
        unsigned best_lo = 0u;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = camli_sqdist3(px[j] - cx, py[j] - cy, pz[j] - cz);
            const float nd = fminf(pd[j], d);
            pd[j] = nd;
            if (nd > best_d || (nd == best_d && plo[j] > best_lo)) { best_d = nd; best_lo = plo[j]; }
        }
        FPS_STAMP(1);
        // ---- warp key, pushed to all 8 CTAs
        const int hi = __float_as_int(best_d);
        const int hmax = __reduce_max_sync(CAMLI_FULL_MASK, hi);
        const unsigned lmax = __reduce_max_sync(CAMLI_FULL_MASK, hi == hmax ? best_lo : 0u);
        if (lane < FPSA_CTAS)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                         ::"r"(r_rec[par]), "r"(hmax), "r"(lmax), "r"(r_bar[par]) : "memory");
        FPS_STAMP(2);
        // ---- wait for the 64 records of this round (bounded: a protocol bug traps instead of hanging)
        const uint32_t phase = (uint32_t)(s >> 1) & 1u;
        for (uint32_t spin = 0;; ++spin) {
            uint32_t done;
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(l_bar[par]), "r"(phase) : "memory");
            if (done) break;
            if (spin > (1u << 24)) __trap();
        }
        FPS_STAMP(3);
        // ---- global winner: max over the 64 (hi, lo) keys
        const uint2 r0 = s_rec[par][lane], r1 = s_rec[par][lane + 32];
        const bool second = (int)r1.x > (int)r0.x || (r1.x == r0.x && r1.y > r0.y);
        const int mh = second ? (int)r1.x : (int)r0.x;
        const unsigned ml = second ? r1.y : r0.y;
        const int gh = __reduce_max_sync(CAMLI_FULL_MASK, mh);
        const unsigned gl = __reduce_max_sync(CAMLI_FULL_MASK, mh == gh ? ml : 0u);
        cur = (int)(0x3FFFFFu - (gl & 0x3FFFFFu));
        FPS_STAMP(4);
    }
    cluster.sync();                                               // no CTA exits while peers may still store to it
}

template <int PPT>
int fps_launch_cluster_async(const float* xyz, int B, int N, int S, int64_t* out, cudaStream_t st) {
    const size_t smem = (size_t)N * 3 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fps_cluster_async_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fps_cluster_async_kernel<PPT><<<B * FPSA_CTAS, FPSA_THREADS, smem, st>>>(xyz, N, S, out);
    CAMLI_RETURN_LAUNCH_STATUS();
}

template <int PPT>
int fps_launch_pruned(const float* xyz, int B, int N, int S, int64_t* out, cudaStream_t st) {
    const size_t smem = FpsPruned<PPT>::kSmem;
    cudaError_t e = cudaFuncSetAttribute(fps_pruned_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fps_pruned_kernel<PPT><<<B, FPS_THREADS, smem, st>>>(xyz, N, S, out);
    CAMLI_RETURN_LAUNCH_STATUS();
}

template <int PPT>
int fps_launch_cluster(const float* xyz, int B, int N, int S, int64_t* out, cudaStream_t st) {
    fps_cluster_kernel<PPT><<<B * FPSC_CTAS, FPSC_THREADS, 0, st>>>(xyz, N, S, out);
    CAMLI_RETURN_LAUNCH_STATUS();
}

template <int PPT>
int fps_launch_register(const float* xyz, int B, int N, int S, int64_t* out, cudaStream_t st) {
    const size_t smem = (size_t)N * sizeof(float4);
    cudaError_t e = cudaFuncSetAttribute(fps_register_kernel<PPT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fps_register_kernel<PPT><<<B, FPS_THREADS, smem, st>>>(xyz, N, S, out);
    CAMLI_RETURN_LAUNCH_STATUS();
}

}  // namespace

// Tuning switch (tests exercise every path) for 2048 < N <= 16384: 0 = single-CTA register kernel,
// 1 = cluster kernel with a cluster barrier per round, 2 = cluster kernel with the st.async exchange (the default
// above 8192 points), 3 = Morton-bucketed single CTA with exact pruning (default; N <= 8192, else falls to 2).
static int camli_fps_use_cluster = 3;
extern "C" int camli_fps_set_cluster_path(int mode) {
    const int old = camli_fps_use_cluster;
    camli_fps_use_cluster = mode < 0 ? 0 : (mode > 3 ? 3 : mode);
    return old;
}

extern "C" int camli_fps_set_timeline(long long* device_buffer) {
    return (int)cudaMemcpyToSymbol(g_fps_timeline, &device_buffer, sizeof(device_buffer));
}

extern "C" int camli_furthest_point_sampling(const float* xyz, float* dists_tmp, int B, int N, int S,
                                             int64_t* out, void* stream) {
    if (B < 0 || N < 1 || S < 0) return CAMLI_EINVAL;
    if (B == 0 || S == 0) return CAMLI_OK;
    if (!xyz || !out) return CAMLI_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 1 * FPS_THREADS) return fps_launch_register<1>(xyz, B, N, S, out, st);
    if (N <= 2 * FPS_THREADS) return fps_launch_register<2>(xyz, B, N, S, out, st);
    if (camli_fps_use_cluster == 3) {
        if (N <= 4 * FPS_THREADS) return fps_launch_pruned<4>(xyz, B, N, S, out, st);
        if (N <= 8 * FPS_THREADS) return fps_launch_pruned<8>(xyz, B, N, S, out, st);
    }
    if (camli_fps_use_cluster >= 2) {
        const int per = FPSA_CTAS * FPSA_THREADS;
        if (N <= 2 * per) return fps_launch_cluster_async<2>(xyz, B, N, S, out, st);
        if (N <= 4 * per) return fps_launch_cluster_async<4>(xyz, B, N, S, out, st);
        if (N <= FPSA_MAX_N) return fps_launch_cluster_async<8>(xyz, B, N, S, out, st);
    }
    if (camli_fps_use_cluster == 1) {
        const int per = FPSC_CTAS * FPSC_THREADS;
        if (N <= 2 * per) return fps_launch_cluster<2>(xyz, B, N, S, out, st);
        if (N <= 4 * per) return fps_launch_cluster<4>(xyz, B, N, S, out, st);
        if (N <= FPSC_MAX_N) return fps_launch_cluster<8>(xyz, B, N, S, out, st);
    }
    if (N <= 4 * FPS_THREADS) return fps_launch_register<4>(xyz, B, N, S, out, st);
    if (N <= FPS_MAX_REG_POINTS) return fps_launch_register<8>(xyz, B, N, S, out, st);
    if (!dists_tmp) return CAMLI_EINVAL;
    fps_streaming_kernel<<<B, FPS_THREADS, 0, st>>>(xyz, dists_tmp, N, S, out);
    CAMLI_RETURN_LAUNCH_STATUS();
}
