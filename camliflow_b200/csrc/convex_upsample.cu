// Convex up-sampling of the 1/8-resolution flow for sm_100a, forward and backward.
//
// Replaces convex_upsample (reference models/utils.py:191-204): softmax over the 9 taps of a [B,1,9,8,8,H,W]
// view of the mask, F.unfold of 8*flow, a broadcast product, a sum over the taps and a permute + reshape --
// six launches that materialise two [B,2,9,8,8,H,W]-sized temporaries (19 MB of mask become ~100 MB of traffic;
// the training step does it for every refinement iteration, forward and backward).
//
//   up[b,c,8h+i,8w+j] = sum_k softmax_k(scale * mask[b, k*64 + i*8 + j, h, w]) * 8 * flow[b,c,h+dy_k,w+dx_k]
//
// with k = 3*(dy+1) + (dx+1) and zeros outside the map (F.unfold's padding); the factor 8 is RAFT's, CamLiPWC's finest
// level uses 4 (models/pwc_core.py:218-224).  `scale` is the reference's 0.25 on the mask (models/raft_core.py:197),
// folded in.  One warp owns one coarse pixel: a lane holds sub-pixels `lane` and `lane + 32` of the 8x8 block (x4: one
// of 16, half of the lanes), the mask row of a pixel (576 / 144 contiguous floats of the NHWC mask, as the mask head's
// convolution kernel writes it) is read with nine coalesced loads, the 18 flow taps are loaded by lanes 0..17 and
// broadcast by shuffles.  HBM-bound: B*H*W*(9 + 2)*S*S*4 bytes.
#include "common.cuh"

namespace {

constexpr int CU_WARPS = 8;
constexpr int CU_TAPS = 9;

template <int S>
struct Cu {
    static constexpr int kSub = S * S;                       // sub-pixels of a coarse pixel: 64 (RAFT, x8) or 16 (PWC, x4)
    static constexpr int kPerLane = (kSub + 31) / 32;        // 2 or 1
    static constexpr int kMask = CU_TAPS * kSub;             // mask channels: 576 or 144
};

struct CuPixel {
    int b, h, w;
    bool ok;
};

__device__ __forceinline__ CuPixel cu_pixel(int B, int H, int W) {
    const long long p = (long long)blockIdx.x * CU_WARPS + (threadIdx.x >> 5);
    CuPixel q;
    q.ok = p < (long long)B * H * W;
    const long long pp = q.ok ? p : 0;
    q.b = (int)(pp / ((long long)H * W));
    const int r = (int)(pp - (long long)q.b * H * W);
    q.h = r / W; q.w = r - q.h * W;
    return q;
}

// S * flow at the 9 taps of pixel (h, w), both channels: lane l < 18 loads tap (c = l / 9, k = l % 9)
template <int S>
__device__ __forceinline__ float cu_tap(const float* __restrict__ flow, const CuPixel& q, int H, int W, int lane) {
    float f = 0.f;
    if (lane < 2 * CU_TAPS) {
        const int c = lane / CU_TAPS, k = lane - c * CU_TAPS;
        const int y = q.h + k / 3 - 1, x = q.w + k % 3 - 1;
        if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W)
            f = __ldg(flow + (((size_t)q.b * 2 + c) * H + y) * W + x) * (float)S;
    }
    return f;
}

// softmax over the 9 taps of this lane's sub-pixels (torch.softmax: exp(x - max) / sum); lanes without a sub-pixel get zeros
template <int S>
__device__ __forceinline__ void cu_softmax(const float* __restrict__ m, float scale, int lane, float (&p)[Cu<S>::kPerLane][CU_TAPS]) {
#pragma unroll
    for (int u = 0; u < Cu<S>::kPerLane; ++u) {
        const int sub = lane + 32 * u;
        if (sub >= Cu<S>::kSub) {
#pragma unroll
            for (int k = 0; k < CU_TAPS; ++k) p[u][k] = 0.f;
            continue;
        }
        float mx = -3.402823466e+38f, sum = 0.f;
#pragma unroll
        for (int k = 0; k < CU_TAPS; ++k) {
            p[u][k] = __ldg(m + k * Cu<S>::kSub + sub) * scale;
            mx = fmaxf(mx, p[u][k]);
        }
#pragma unroll
        for (int k = 0; k < CU_TAPS; ++k) {
            p[u][k] = expf(p[u][k] - mx);
            sum += p[u][k];
        }
#pragma unroll
        for (int k = 0; k < CU_TAPS; ++k) p[u][k] = p[u][k] / sum;
    }
}

template <int S>
__global__ void __launch_bounds__(CU_WARPS * 32)
convex_upsample_kernel(int B, int H, int W, const float* __restrict__ flow, const float* __restrict__ mask_rows, float scale,
                       float* __restrict__ up) {
    constexpr int NPL = Cu<S>::kPerLane;
    const CuPixel q = cu_pixel(B, H, W);
    if (!q.ok) return;
    const int lane = threadIdx.x & 31;
    const float tap = cu_tap<S>(flow, q, H, W, lane);
    float p[NPL][CU_TAPS];
    cu_softmax<S>(mask_rows + (((size_t)q.b * H + q.h) * W + q.w) * Cu<S>::kMask, scale, lane, p);
    float acc[2][NPL];                                             // [channel][sub-pixel of this lane]
#pragma unroll
    for (int u = 0; u < NPL; ++u) acc[0][u] = acc[1][u] = 0.f;
#pragma unroll
    for (int k = 0; k < CU_TAPS; ++k) {
        const float f0 = __shfl_sync(CAMLI_FULL_MASK, tap, k), f1 = __shfl_sync(CAMLI_FULL_MASK, tap, CU_TAPS + k);
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            acc[0][u] = fmaf(p[u][k], f0, acc[0][u]);
            acc[1][u] = fmaf(p[u][k], f1, acc[1][u]);
        }
    }
    const size_t HS = (size_t)H * S, WS = (size_t)W * S;
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int sub = lane + 32 * u;
        if (sub >= Cu<S>::kSub) continue;
        const int i = sub / S, j = sub % S;
        float* o = up + (((size_t)q.b * 2) * HS + (size_t)q.h * S + i) * WS + (size_t)q.w * S + j;
        o[0] = acc[0][u];
        o[HS * WS] = acc[1][u];
    }
}

// Backward, pass 1 (same mapping): the softmax is recomputed; writes the mask gradient and, per pixel, the 18
// tap sums  v[c][k] = S * sum_{i,j} p_k(i,j) * g[b,c,S*h+i,S*w+j]  (the flow gradient before it is gathered).
template <int S>
__global__ void __launch_bounds__(CU_WARPS * 32)
convex_upsample_backward_kernel(int B, int H, int W, const float* __restrict__ flow, const float* __restrict__ mask_rows, float scale,
                                const float* __restrict__ g_up, float* __restrict__ g_mask_rows, float* __restrict__ tap_sums) {
    constexpr int NPL = Cu<S>::kPerLane;
    const CuPixel q = cu_pixel(B, H, W);
    if (!q.ok) return;
    const int lane = threadIdx.x & 31;
    const float tap = cu_tap<S>(flow, q, H, W, lane);
    float p[NPL][CU_TAPS];
    const size_t pix = ((size_t)q.b * H + q.h) * W + q.w;
    cu_softmax<S>(mask_rows + pix * Cu<S>::kMask, scale, lane, p);
    const size_t HS = (size_t)H * S, WS = (size_t)W * S;
    float g[2][NPL];
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int sub = lane + 32 * u;
        g[0][u] = g[1][u] = 0.f;
        if (sub < Cu<S>::kSub) {
            const int i = sub / S, j = sub % S;
            const float* gp = g_up + (((size_t)q.b * 2) * HS + (size_t)q.h * S + i) * WS + (size_t)q.w * S + j;
            g[0][u] = __ldg(gp);
            g[1][u] = __ldg(gp + HS * WS);
        }
    }
    // d p_k = sum_c g_c * f[c][k];  d logit_k = p_k * (d p_k - sum_k' p_k' d p_k')
    float dp[NPL][CU_TAPS], dot[NPL], mine = 0.f;
#pragma unroll
    for (int u = 0; u < NPL; ++u) dot[u] = 0.f;
#pragma unroll
    for (int k = 0; k < CU_TAPS; ++k) {
        const float f0 = __shfl_sync(CAMLI_FULL_MASK, tap, k), f1 = __shfl_sync(CAMLI_FULL_MASK, tap, CU_TAPS + k);
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            dp[u][k] = fmaf(g[1][u], f1, g[0][u] * f0);
            dot[u] = fmaf(p[u][k], dp[u][k], dot[u]);
            v0 = fmaf(p[u][k], g[0][u], v0);
            v1 = fmaf(p[u][k], g[1][u], v1);
        }
        // the 18 tap sums: lane l < 18 keeps the one of (c = l / 9, k = l % 9)
        v0 = camli_warp_sum(v0); v1 = camli_warp_sum(v1);
        if (lane == k) mine = v0;
        if (lane == CU_TAPS + k) mine = v1;
    }
    float* gm = g_mask_rows + pix * Cu<S>::kMask;
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int sub = lane + 32 * u;
        if (sub >= Cu<S>::kSub) continue;
#pragma unroll
        for (int k = 0; k < CU_TAPS; ++k) gm[k * Cu<S>::kSub + sub] = scale * p[u][k] * (dp[u][k] - dot[u]);
    }
    if (lane < 2 * CU_TAPS) tap_sums[pix * 2 * CU_TAPS + lane] = mine * (float)S;
}

// Backward, pass 2: g_flow[b,c,y,x] = sum_k v[b, y - dy_k, x - dx_k][c][k]  (a gather: deterministic, no atomics)
__global__ void __launch_bounds__(256)
convex_upsample_flow_grad_kernel(int B, int H, int W, const float* __restrict__ tap_sums, float* __restrict__ g_flow) {
    const long long n = (long long)B * 2 * H * W;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(e % W), y = (int)((e / W) % H), c = (int)((e / ((long long)W * H)) % 2), b = (int)(e / ((long long)W * H * 2));
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < CU_TAPS; ++k) {
            const int ys = y - (k / 3 - 1), xs = x - (k % 3 - 1);
            if ((unsigned)ys < (unsigned)H && (unsigned)xs < (unsigned)W)
                acc += __ldg(tap_sums + (((size_t)b * H + ys) * W + xs) * 2 * CU_TAPS + c * CU_TAPS + k);
        }
        g_flow[e] = acc;
    }
}

}  // namespace

extern "C" int camli_convex_upsample(int B, int H, int W, int factor, const float* flow, const float* mask_rows, float scale,
                                     float* up, void* stream) {
    if (B < 0 || H < 1 || W < 1) return CAMLI_EINVAL;
    if (factor != 4 && factor != 8) return CAMLI_EUNSUPPORTED;
    if ((long long)B * H * W * factor * factor > 2147483647LL) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!flow || !mask_rows || !up) return CAMLI_EINVAL;
    const unsigned grid = (unsigned)camli_div_up_ll((long long)B * H * W, CU_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
    if (factor == 8) convex_upsample_kernel<8><<<grid, CU_WARPS * 32, 0, st>>>(B, H, W, flow, mask_rows, scale, up);
    else             convex_upsample_kernel<4><<<grid, CU_WARPS * 32, 0, st>>>(B, H, W, flow, mask_rows, scale, up);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_convex_upsample_backward(int B, int H, int W, int factor, const float* flow, const float* mask_rows,
                                              float scale, const float* grad_up, float* grad_mask_rows, float* tap_scratch,
                                              float* grad_flow, void* stream) {
    if (B < 0 || H < 1 || W < 1) return CAMLI_EINVAL;
    if (factor != 4 && factor != 8) return CAMLI_EUNSUPPORTED;
    if ((long long)B * H * W * factor * factor > 2147483647LL) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!flow || !mask_rows || !grad_up || !grad_mask_rows || !tap_scratch || !grad_flow) return CAMLI_EINVAL;
    const long long pixels = (long long)B * H * W;
    const unsigned grid = (unsigned)camli_div_up_ll(pixels, CU_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
    if (factor == 8)
        convex_upsample_backward_kernel<8><<<grid, CU_WARPS * 32, 0, st>>>(B, H, W, flow, mask_rows, scale, grad_up, grad_mask_rows, tap_scratch);
    else
        convex_upsample_backward_kernel<4><<<grid, CU_WARPS * 32, 0, st>>>(B, H, W, flow, mask_rows, scale, grad_up, grad_mask_rows, tap_scratch);
    const long long blocks = camli_div_up_ll(pixels * 2, 256);
    convex_upsample_flow_grad_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(B, H, W, tap_scratch, grad_flow);
    CAMLI_RETURN_LAUNCH_STATUS();
}
