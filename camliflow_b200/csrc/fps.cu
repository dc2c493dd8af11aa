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

extern "C" int camli_furthest_point_sampling(const float* xyz, float* dists_tmp, int B, int N, int S,
                                             int64_t* out, void* stream) {
    if (B < 0 || N < 1 || S < 0) return CAMLI_EINVAL;
    if (B == 0 || S == 0) return CAMLI_OK;
    if (!xyz || !out) return CAMLI_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 1 * FPS_THREADS) return fps_launch_register<1>(xyz, B, N, S, out, st);
    if (N <= 2 * FPS_THREADS) return fps_launch_register<2>(xyz, B, N, S, out, st);
    if (N <= 4 * FPS_THREADS) return fps_launch_register<4>(xyz, B, N, S, out, st);
    if (N <= FPS_MAX_REG_POINTS) return fps_launch_register<8>(xyz, B, N, S, out, st);
    if (!dists_tmp) return CAMLI_EINVAL;
    fps_streaming_kernel<<<B, FPS_THREADS, 0, st>>>(xyz, dists_tmp, N, S, out);
    CAMLI_RETURN_LAUNCH_STATUS();
}
