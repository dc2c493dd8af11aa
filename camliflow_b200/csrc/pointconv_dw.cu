// PointConvDW grouping stage for sm_100a.
//
// Replaces the tail of PointConvDW.forward (reference models/point_conv.py:119-128):
//   knn_xyz = gather(xyz); offset = knn_xyz - centre; W = weight_net(offset)   [B,O,S,k]
//   out = max_k( gather(features) * W )                                          [B,O,S]
// which materialises three [B,O,S,k] tensors (33.5 MB each at O=128, k=32, S=2048) through
// 4-byte strided channel-first gathers, nine times per GRU iteration.
//
// Two kernels, both working on CHANNEL-LAST rows so that a neighbour's feature vector is one
// contiguous, coalesced read:
//  (1) pointconv_dw_weights: WeightNet 3->8->32->O (ReLU) of every neighbour offset, written as
//      rows [B,S,k,O].  The result depends only on the geometry and the layer's parameters, not
//      on the features, so inside the recurrent loop it is computed ONCE per layer and reused by
//      all iterations (the reference recomputes it every iteration).
//  (2) pointconv_dw_gather_max: out[s,o] = max_j feat[idx[s,j], o] * Wc[s,j,o].  Pure gather +
//      stream: algorithmic bytes B*S*k*(2*O*4 + 8) read + B*S*O*4 written; HBM-bound.
#include "common.cuh"

namespace {

constexpr int DW_H1 = 8, DW_H2 = 32;

// ---------------------------------------------------------------- (1) WeightNet
// One warp per centroid s.  Lane = hidden unit of layer 2; layer 3 keeps W3 transposed in
// shared memory ([32][O]) so lanes read consecutive output channels.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
dw_weights_kernel(int N, int S, int K, int k, int O,
                  const float* __restrict__ xyz, long long x_sb, long long x_sp, long long x_sd,
                  const float* __restrict__ centre, long long c_sb, long long c_sp, long long c_sd,
                  const int64_t* __restrict__ idx,
                  const float* __restrict__ W1, const float* __restrict__ b1,    // [8,3],[8]
                  const float* __restrict__ W2, const float* __restrict__ b2,    // [32,8],[32]
                  const float* __restrict__ W3, const float* __restrict__ b3,    // [O,32],[O]
                  float* __restrict__ out) {                                     // [B,S,k,O]
    extern __shared__ float s_w3t[];            // [32][O] then b3 [O]
    float* s_b3 = s_w3t + DW_H2 * O;
    for (int e = threadIdx.x; e < O * DW_H2; e += WARPS * 32) {
        const int o = e / DW_H2, a = e - o * DW_H2;
        s_w3t[a * O + o] = __ldg(W3 + e);
    }
    for (int o = threadIdx.x; o < O; o += WARPS * 32) s_b3[o] = __ldg(b3 + o);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (s >= S) return;
    const int b = blockIdx.y;

    float w1[DW_H1][3], bb1[DW_H1], w2[DW_H1];
#pragma unroll
    for (int a = 0; a < DW_H1; ++a) {
        w1[a][0] = __ldg(W1 + a * 3); w1[a][1] = __ldg(W1 + a * 3 + 1); w1[a][2] = __ldg(W1 + a * 3 + 2);
        bb1[a] = __ldg(b1 + a);
        w2[a] = __ldg(W2 + lane * DW_H1 + a);
    }
    const float bb2 = __ldg(b2 + lane);
    const float* cp = centre + b * c_sb + s * c_sp;
    const float cx = __ldg(cp), cy = __ldg(cp + c_sd), cz = __ldg(cp + 2 * c_sd);
    const int64_t* ip = idx + ((size_t)b * S + s) * K;
    const float* xb = xyz + b * x_sb;
    float* ob = out + ((size_t)b * S + s) * k * O;

    for (int j = 0; j < k; ++j) {
        const float* p = xb + __ldg(ip + j) * x_sp;
        const float dx = __ldg(p) - cx, dy = __ldg(p + x_sd) - cy, dz = __ldg(p + 2 * x_sd) - cz;
        float h2 = bb2;
#pragma unroll
        for (int a = 0; a < DW_H1; ++a) {
            const float h1 = fmaxf(fmaf(w1[a][2], dz, fmaf(w1[a][1], dy, fmaf(w1[a][0], dx, bb1[a]))), 0.f);
            h2 = fmaf(w2[a], h1, h2);
        }
        h2 = fmaxf(h2, 0.f);
        for (int o0 = 0; o0 < O; o0 += 32) {
            const int o = o0 + lane;
            float acc = (o < O) ? s_b3[o] : 0.f;
#pragma unroll
            for (int a = 0; a < DW_H2; ++a) {
                const float ha = __shfl_sync(CAMLI_FULL_MASK, h2, a);
                if (o < O) acc = fmaf(s_w3t[a * O + o], ha, acc);
            }
            if (o < O) ob[(size_t)j * O + o] = fmaxf(acc, 0.f);
        }
    }
}

// ---------------------------------------------------------------- (1b) WeightNet hidden layer only
// Layers 1-2 of the WeightNet (3 -> 8 -> 32, ReLU) per neighbour, written as rows [B,S,k,32]: the 32 -> O
// output layer is then ONE tensor-core GEMM over all B*S*k neighbours (camli_conv_gemm with its ReLU epilogue,
// K = 32 = a single k-block), which writes the [B,S,k,O] cache directly.  Lane = hidden unit.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
dw_hidden_kernel(int N, int S, int K, int k,
                 const float* __restrict__ xyz, long long x_sb, long long x_sp, long long x_sd,
                 const float* __restrict__ centre, long long c_sb, long long c_sp, long long c_sd,
                 const int64_t* __restrict__ idx,
                 const float* __restrict__ W1, const float* __restrict__ b1,    // [8,3],[8]
                 const float* __restrict__ W2, const float* __restrict__ b2,    // [32,8],[32]
                 float* __restrict__ out) {                                     // [B,S,k,32]
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (s >= S) return;
    const int b = blockIdx.y;
    float w1[DW_H1][3], bb1[DW_H1], w2[DW_H1];
#pragma unroll
    for (int a = 0; a < DW_H1; ++a) {
        w1[a][0] = __ldg(W1 + a * 3); w1[a][1] = __ldg(W1 + a * 3 + 1); w1[a][2] = __ldg(W1 + a * 3 + 2);
        bb1[a] = __ldg(b1 + a);
        w2[a] = __ldg(W2 + lane * DW_H1 + a);
    }
    const float bb2 = __ldg(b2 + lane);
    const float* cp = centre + b * c_sb + s * c_sp;
    const float cx = __ldg(cp), cy = __ldg(cp + c_sd), cz = __ldg(cp + 2 * c_sd);
    const int64_t* ip = idx + ((size_t)b * S + s) * K;
    const float* xb = xyz + b * x_sb;
    float* ob = out + ((size_t)b * S + s) * k * DW_H2;
    for (int j = 0; j < k; ++j) {
        const float* p = xb + __ldg(ip + j) * x_sp;
        const float dx = __ldg(p) - cx, dy = __ldg(p + x_sd) - cy, dz = __ldg(p + 2 * x_sd) - cz;
        float h2 = bb2;
#pragma unroll
        for (int a = 0; a < DW_H1; ++a) {
            const float h1 = fmaxf(fmaf(w1[a][2], dz, fmaf(w1[a][1], dy, fmaf(w1[a][0], dx, bb1[a]))), 0.f);
            h2 = fmaf(w2[a], h1, h2);
        }
        ob[j * DW_H2 + lane] = fmaxf(h2, 0.f);
    }
}

// ---------------------------------------------------------------- (2) gather * weight, max over k
// One warp per (centroid, 32-channel chunk).  The kernel is a pure stream (no reuse), so its speed is the
// number of bytes in flight: ALL 2k 128-byte loads of a warp (k feature rows, k weight rows) are issued
// before the first one is consumed (KT = k known at compile time, registers hold the k x 2 values); an
// in-order warp that consumed them batch by batch would expose one full memory latency per batch.
template <int WARPS, int KT>
__global__ void __launch_bounds__(WARPS * 32)
dw_gather_max_kernel(int N, int S, int K, int k, int O, int chunks,
                     const float* __restrict__ feat, long long ldf,      // rows [B,N,ldf]
                     const float* __restrict__ wc,                       // rows [B,S,k,O]
                     const int64_t* __restrict__ idx,                    // [B,S,K]
                     float* __restrict__ out, long long ldo) {           // rows [B,S,ldo]
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (wid >= (long long)S * chunks) return;
    const int s = (int)(wid / chunks), o = (int)(wid % chunks) * 32 + lane;
    const int b = blockIdx.y;
    const int64_t* ip = idx + ((size_t)b * S + s) * K;
    const int my = (lane < k) ? (int)__ldg(ip + lane) : 0;
    const float* fb = feat + (size_t)b * N * ldf;
    const float* wp = wc + ((size_t)b * S + s) * k * O;
    float best = -INFINITY;
    if (KT > 0) {
        float f[KT > 0 ? KT : 1], w[KT > 0 ? KT : 1];
        const bool live = o < O;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            const int ij = __shfl_sync(CAMLI_FULL_MASK, my, j);
            f[j] = live ? __ldg(fb + (size_t)ij * ldf + o) : 0.f;
            w[j] = live ? __ldcs(wp + (size_t)j * O + o) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < KT; ++j) best = fmaxf(best, f[j] * w[j]);
        if (live) out[((size_t)b * S + s) * ldo + o] = best;
        return;
    }
    // generic k: batches of 4 neighbours
    int j = 0;
    for (; j + 4 <= k; j += 4) {
        int ij[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) ij[u] = __shfl_sync(CAMLI_FULL_MASK, my, j + u);
        if (o < O) {
            float f[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                f[u] = __ldg(fb + (size_t)ij[u] * ldf + o);
                w[u] = __ldcs(wp + (size_t)(j + u) * O + o);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) best = fmaxf(best, f[u] * w[u]);
        }
    }
    for (; j < k; ++j) {
        const int ij = __shfl_sync(CAMLI_FULL_MASK, my, j);
        if (o < O) best = fmaxf(best, __ldg(fb + (size_t)ij * ldf + o) * __ldcs(wp + (size_t)j * O + o));
    }
    if (o < O) out[((size_t)b * S + s) * ldo + o] = best;
}

// 128-bit variant for O % 128 == 0 (the 128- and 256-channel layers that dominate the traffic): one warp owns a
// centroid and 128 consecutive channels, lane = 4 channels, so every request is a full 512-byte row segment (the
// k weight rows of a centroid are one contiguous 16 KB block) and the instruction count drops 4x.  k = KT * NCH
// neighbours are processed in NCH chunks of KT (2 * KT 128-bit loads in flight per lane: 128 registers at KT = 16),
// the running maximum carried across chunks.
template <int WARPS, int KT, int NCH>
__global__ void __launch_bounds__(WARPS * 32)
dw_gather_max_vec4_kernel(int N, int S, int K, int O, int chunks,          // chunks = O / 128
                          const float* __restrict__ feat, long long ldf, const float* __restrict__ wc,
                          const int64_t* __restrict__ idx, float* __restrict__ out, long long ldo) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (wid >= (long long)S * chunks) return;
    const int s = (int)(wid / chunks), o = (int)(wid % chunks) * 128 + lane * 4;
    const int b = blockIdx.y;
    const int64_t* ip = idx + ((size_t)b * S + s) * K;
    const int my = (lane < KT * NCH) ? (int)__ldg(ip + lane) : 0;          // k <= 32: one index per lane
    const float* fb = feat + (size_t)b * N * ldf + o;
    const float* wp = wc + ((size_t)b * S + s) * (KT * NCH) * O + o;
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float4 f[KT], w[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            const int ij = __shfl_sync(CAMLI_FULL_MASK, my, c * KT + j);
            f[j] = __ldg(reinterpret_cast<const float4*>(fb + (size_t)ij * ldf));
            w[j] = __ldcs(reinterpret_cast<const float4*>(wp + (size_t)(c * KT + j) * O));
        }
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            best.x = fmaxf(best.x, f[j].x * w[j].x); best.y = fmaxf(best.y, f[j].y * w[j].y);
            best.z = fmaxf(best.z, f[j].z * w[j].z); best.w = fmaxf(best.w, f[j].w * w[j].w);
        }
    }
    *reinterpret_cast<float4*>(out + ((size_t)b * S + s) * ldo + o) = best;
}

}  // namespace

extern "C" int camli_pointconv_dw_weights(int B, int N, int S, int K, int k, int O,
                                          const float* xyz, int64_t x_sb, int64_t x_sp, int64_t x_sd,
                                          const float* centre_xyz, int64_t c_sb, int64_t c_sp, int64_t c_sd,
                                          const int64_t* knn_idx,
                                          const float* W1, const float* b1, const float* W2, const float* b2,
                                          const float* W3, const float* b3, float* weights_out, void* stream) {
    if (B < 0 || N < 1 || S < 0 || k < 1 || K < k || O < 1) return CAMLI_EINVAL;
    if (k > 64 || B > 65535 || O > 1024) return CAMLI_EUNSUPPORTED;
    if (B == 0 || S == 0) return CAMLI_OK;
    if (!xyz || !centre_xyz || !knn_idx || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !weights_out) return CAMLI_EINVAL;
    constexpr int WARPS = 8;
    const size_t smem = (size_t)(DW_H2 + 1) * O * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(dw_weights_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(camli_div_up(S, WARPS), B);
    dw_weights_kernel<WARPS><<<grid, WARPS * 32, smem, (cudaStream_t)stream>>>(
        N, S, K, k, O, xyz, x_sb, x_sp, x_sd, centre_xyz, c_sb, c_sp, c_sd, knn_idx, W1, b1, W2, b2, W3, b3, weights_out);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_pointconv_dw_hidden(int B, int N, int S, int K, int k,
                                         const float* xyz, int64_t x_sb, int64_t x_sp, int64_t x_sd,
                                         const float* centre_xyz, int64_t c_sb, int64_t c_sp, int64_t c_sd,
                                         const int64_t* knn_idx, const float* W1, const float* b1, const float* W2,
                                         const float* b2, float* hidden_out, void* stream) {
    if (B < 0 || N < 1 || S < 0 || k < 1 || K < k) return CAMLI_EINVAL;
    if (k > 64 || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || S == 0) return CAMLI_OK;
    if (!xyz || !centre_xyz || !knn_idx || !W1 || !b1 || !W2 || !b2 || !hidden_out) return CAMLI_EINVAL;
    constexpr int WARPS = 8;
    dim3 grid(camli_div_up(S, WARPS), B);
    dw_hidden_kernel<WARPS><<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(
        N, S, K, k, xyz, x_sb, x_sp, x_sd, centre_xyz, c_sb, c_sp, c_sd, knn_idx, W1, b1, W2, b2, hidden_out);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_pointconv_dw_gather_max(int B, int N, int S, int K, int k, int O,
                                             const float* feat_rows, int64_t ld_feat, const float* weights,
                                             const int64_t* knn_idx, float* out_rows, int64_t ld_out, void* stream) {
    if (B < 0 || N < 1 || S < 0 || k < 1 || K < k || O < 1 || ld_feat < O || ld_out < O) return CAMLI_EINVAL;
    if (k > 32 || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || S == 0) return CAMLI_OK;
    if (!feat_rows || !weights || !knn_idx || !out_rows) return CAMLI_EINVAL;
    constexpr int WARPS = 8;
    const int chunks = camli_div_up(O, 32);
    dim3 grid((unsigned)camli_div_up_ll((long long)S * chunks, WARPS), B);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (O % 128 == 0) && (ld_feat % 4 == 0) && (ld_out % 4 == 0) && (k == 4 || k == 8 || k == 16 || k == 32) &&
                      ((reinterpret_cast<uintptr_t>(feat_rows) | reinterpret_cast<uintptr_t>(weights) |
                        reinterpret_cast<uintptr_t>(out_rows)) % 16 == 0);
    if (vec4) {          // (k = 32: two chunks of 16 neighbours, 128 registers of loads in flight each)
        const int ch = O / 128;
        constexpr int W4 = 2;          // 2-warp CTAs: S * ch / 2 = 1024+ CTAs, ~7 per SM -- an even spread over 148 SMs
        dim3 g4((unsigned)camli_div_up_ll((long long)S * ch, W4), B);
#define CAMLI_DW_LAUNCH4(KT, NCH)                                                                                       \
        dw_gather_max_vec4_kernel<W4, KT, NCH><<<g4, W4 * 32, 0, st>>>(N, S, K, O, ch, feat_rows, ld_feat, weights, \
                                                                       knn_idx, out_rows, ld_out)
        if (k == 4) CAMLI_DW_LAUNCH4(4, 1);
        else if (k == 8) CAMLI_DW_LAUNCH4(8, 1);
        else if (k == 16) CAMLI_DW_LAUNCH4(16, 1);
        else CAMLI_DW_LAUNCH4(16, 2);
#undef CAMLI_DW_LAUNCH4
        CAMLI_RETURN_LAUNCH_STATUS();
    }
#define CAMLI_DW_LAUNCH(KT)                                                                                    \
    dw_gather_max_kernel<WARPS, KT><<<grid, WARPS * 32, 0, st>>>(N, S, K, k, O, chunks, feat_rows, ld_feat, weights, \
                                                                 knn_idx, out_rows, ld_out)
    switch (k) {          // the neighbourhood sizes of the models get fully unrolled instances
        case 4: CAMLI_DW_LAUNCH(4); break;
        case 8: CAMLI_DW_LAUNCH(8); break;
        case 16: CAMLI_DW_LAUNCH(16); break;
        case 32: CAMLI_DW_LAUNCH(32); break;
        default: CAMLI_DW_LAUNCH(0); break;
    }
#undef CAMLI_DW_LAUNCH
    CAMLI_RETURN_LAUNCH_STATUS();
}
