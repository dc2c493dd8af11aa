// PointConv set-abstraction grouping stage for sm_100a.
//
// Replaces the body of PointConv.forward up to the Linear (reference models/point_conv.py:56-66):
//   offsets = gather(xyz) - centre;  W = WeightNet(3->8->16, leaky 0.1)(offsets)      [B,16,S,k]
//   G = gather([xyz | features])                                                      [B,S,k,C+3]
//   out = (W^T viewed [B,S,16,k]) @ G  -> [B,S,16*(C+3)]
// The reference materialises offsets, W and G (channel-first 4-byte strided gathers) and runs a
// batched 16 x k x (C+3) matmul per centroid.  Here one warp owns a centroid: lane j evaluates
// the WeightNet of neighbour j in registers, lanes then sweep the channels of the neighbours'
// contiguous channel-last rows (coalesced) accumulating the 16 weighted sums per channel, and the
// [16,(C+3)] block is written in the layout the following Linear (a plain GEMM) consumes.
// Bytes: B*S*(k*((C+3)*4 + 8) + 16*(C+3)*4); flops 2*B*S*16*k*(C+3).
#include "common.cuh"

namespace {

constexpr int PC_WARPS = 4;
constexpr int PC_NW = 16;      // WeightNet outputs
constexpr int PC_H = 8;        // WeightNet hidden units
constexpr int PC_MAX_CHUNKS = 8;

template <int CHUNKS>
__global__ void __launch_bounds__(PC_WARPS * 32)
pointconv_group_kernel(int N, int S, int K, int k, int C,                                   // C = channels of `rows` (xyz included)
                       const float* __restrict__ rows, long long ldr,                       // [B,N,ldr]: xyz in columns 0..2
                       const float* __restrict__ centre, long long c_sb, long long c_sp, long long c_sd,
                       const int64_t* __restrict__ idx,
                       const float* __restrict__ W1, const float* __restrict__ b1,          // [8,3],[8]
                       const float* __restrict__ W2, const float* __restrict__ b2,          // [16,8],[16]
                       float slope, float* __restrict__ out) {                              // [B,S,16*C]
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * PC_WARPS + (threadIdx.x >> 5);
    if (s >= S) return;
    const int b = blockIdx.y;
    const float* rb = rows + (size_t)b * N * ldr;
    const float* cp = centre + b * c_sb + s * c_sp;
    const float cx = __ldg(cp), cy = __ldg(cp + c_sd), cz = __ldg(cp + 2 * c_sd);

    // lane j (< k): neighbour index and its 16 WeightNet outputs
    int my = 0;
    float w[PC_NW];
#pragma unroll
    for (int o = 0; o < PC_NW; ++o) w[o] = 0.f;
    if (lane < k) {
        my = (int)__ldg(idx + ((size_t)b * S + s) * K + lane);
        const float* p = rb + (size_t)my * ldr;
        const float dx = __ldg(p) - cx, dy = __ldg(p + 1) - cy, dz = __ldg(p + 2) - cz;
        float h[PC_H];
#pragma unroll
        for (int a = 0; a < PC_H; ++a)
            h[a] = camli_leaky(fmaf(__ldg(W1 + a * 3 + 2), dz, fmaf(__ldg(W1 + a * 3 + 1), dy,
                               fmaf(__ldg(W1 + a * 3), dx, __ldg(b1 + a)))), slope);
#pragma unroll
        for (int o = 0; o < PC_NW; ++o) {
            float acc = __ldg(b2 + o);
#pragma unroll
            for (int a = 0; a < PC_H; ++a) acc = fmaf(__ldg(W2 + o * PC_H + a), h[a], acc);
            w[o] = camli_leaky(acc, slope);
        }
    }

    float acc[CHUNKS][PC_NW];
#pragma unroll
    for (int t = 0; t < CHUNKS; ++t)
#pragma unroll
        for (int o = 0; o < PC_NW; ++o) acc[t][o] = 0.f;

    for (int j = 0; j < k; ++j) {
        const int ij = __shfl_sync(CAMLI_FULL_MASK, my, j);
        const float* g = rb + (size_t)ij * ldr;
        float gv[CHUNKS];
#pragma unroll
        for (int t = 0; t < CHUNKS; ++t) {
            const int c = t * 32 + lane;
            gv[t] = (c < C) ? __ldg(g + c) : 0.f;
        }
#pragma unroll
        for (int o = 0; o < PC_NW; ++o) {
            const float wj = __shfl_sync(CAMLI_FULL_MASK, w[o], j);
#pragma unroll
            for (int t = 0; t < CHUNKS; ++t) acc[t][o] = fmaf(wj, gv[t], acc[t][o]);
        }
    }
    float* ob = out + ((size_t)b * S + s) * PC_NW * C;
#pragma unroll
    for (int o = 0; o < PC_NW; ++o)
#pragma unroll
        for (int t = 0; t < CHUNKS; ++t) {
            const int c = t * 32 + lane;
            if (c < C) ob[(size_t)o * C + c] = acc[t][o];
        }
}

}  // namespace

extern "C" int camli_pointconv_group(int B, int N, int S, int K, int k, int C,
                                     const float* rows, int64_t ld_rows,
                                     const float* centre_xyz, int64_t c_sb, int64_t c_sp, int64_t c_sd,
                                     const int64_t* knn_idx, const float* W1, const float* b1, const float* W2,
                                     const float* b2, float negative_slope, float* out, void* stream) {
    if (B < 0 || N < 1 || S < 0 || k < 1 || K < k || C < 3 || ld_rows < C) return CAMLI_EINVAL;
    if (k > 32 || B > 65535 || C > 32 * PC_MAX_CHUNKS) return CAMLI_EUNSUPPORTED;
    if (B == 0 || S == 0) return CAMLI_OK;
    if (!rows || !centre_xyz || !knn_idx || !W1 || !b1 || !W2 || !b2 || !out) return CAMLI_EINVAL;
    dim3 grid(camli_div_up(S, PC_WARPS), B), block(PC_WARPS * 32);
    cudaStream_t st = (cudaStream_t)stream;
#define CAMLI_PC_LAUNCH(CH)                                                                                   \
    pointconv_group_kernel<CH><<<grid, block, 0, st>>>(N, S, K, k, C, rows, ld_rows, centre_xyz, c_sb, c_sp, \
                                                       c_sd, knn_idx, W1, b1, W2, b2, negative_slope, out)
    const int chunks = camli_div_up(C, 32);
    if (chunks <= 2) CAMLI_PC_LAUNCH(2);
    else if (chunks <= 4) CAMLI_PC_LAUNCH(4);
    else if (chunks <= 6) CAMLI_PC_LAUNCH(6);
    else CAMLI_PC_LAUNCH(8);
#undef CAMLI_PC_LAUNCH
    CAMLI_RETURN_LAUNCH_STATUS();
}
