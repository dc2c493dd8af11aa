// CLFM camera<->LiDAR fusion kernels for sm_100a.
//
// Replaces the gather / ScoreNet part of FusionAwareInterp.forward (3D->2D, reference
// models/clfm.py:53-76) and grid_sample_wrapper (2D->3D, models/utils.py:262-269).
//
// 3D->2D: for every pixel of the feature grid, its nearest projected point nn (2-D k-NN, k=1,
// computed once per point set and reused by all fusion sites / iterations), the offset
// (du, dv, |d|), ScoreNet 3->16 (leaky 0.1) ->C (sigmoid), times the point's feature vector.
// The reference gathers a [B,C+2,HW,1] tensor channel-first (4-byte strided reads) and runs two
// 1x1 Conv2d; here a warp owns a pixel, lanes own channels, the point's row is one coalesced
// read and the result is written as NHWC rows [B,H*W,C].  Bytes: B*HW*(8 + 2*C*4).
//
// 2D->3D: bilinear sample (align_corners, zero padding) of an NHWC map at the projected points,
// lanes over channels: 4 coalesced row reads per point, rows [B,N,C] out.  Bytes: B*N*5*C*4.
#include "common.cuh"

namespace {

constexpr int CF_WARPS = 8;
constexpr int CF_H = 16;

__global__ void __launch_bounds__(CF_WARPS * 32)
clfm_interp_kernel(int HW, int W, int N, int C, const float* __restrict__ uv,          // [B,2,N]
                   const int64_t* __restrict__ nn_idx,                                  // [B,HW]
                   const float* __restrict__ feat3d, long long ldf,                     // rows [B,N,ldf]
                   const float* __restrict__ W1, const float* __restrict__ b1,          // [16,3],[16]
                   const float* __restrict__ W2, const float* __restrict__ b2,          // [C,16],[C]
                   float* __restrict__ out) {                                           // rows [B,HW,C]
    extern __shared__ float s_w2t[];           // [16][C], then b2 [C]
    float* s_b2 = s_w2t + CF_H * C;
    for (int e = threadIdx.x; e < C * CF_H; e += CF_WARPS * 32) {
        const int c = e / CF_H, a = e - c * CF_H;
        s_w2t[a * C + c] = __ldg(W2 + e);
    }
    for (int c = threadIdx.x; c < C; c += CF_WARPS * 32) s_b2[c] = __ldg(b2 + c);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * CF_WARPS + (threadIdx.x >> 5);
    if (p >= HW) return;
    const int b = blockIdx.y;
    const int nn = (int)__ldg(nn_idx + (size_t)b * HW + p);
    const float gx = (float)(p % W), gy = (float)(p / W);
    const float du = __ldg(uv + ((size_t)b * 2 + 0) * N + nn) - gx;
    const float dv = __ldg(uv + ((size_t)b * 2 + 1) * N + nn) - gy;
    const float dn = sqrtf(du * du + dv * dv);
    float h[CF_H];
#pragma unroll
    for (int a = 0; a < CF_H; ++a) {
        const float v = fmaf(__ldg(W1 + a * 3 + 2), dn, fmaf(__ldg(W1 + a * 3 + 1), dv, fmaf(__ldg(W1 + a * 3), du, __ldg(b1 + a))));
        h[a] = camli_leaky(v, 0.1f);
    }
    const float* frow = feat3d + ((size_t)b * N + nn) * ldf;
    float* orow = out + ((size_t)b * HW + p) * C;
    for (int c = lane; c < C; c += 32) {
        float acc = s_b2[c];
#pragma unroll
        for (int a = 0; a < CF_H; ++a) acc = fmaf(s_w2t[a * C + c], h[a], acc);
        const float score = 1.f / (1.f + expf(-acc));
        orow[c] = score * __ldg(frow + c);
    }
}

__global__ void __launch_bounds__(CF_WARPS * 32)
bilinear_sample_rows_kernel(int H, int W, int N, int C, const float* __restrict__ feat,   // NHWC [B,H,W,C]
                            const float* __restrict__ uv,                                  // [B,2,N]
                            float* __restrict__ out, long long ldo) {                      // rows [B,N,ldo]
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * CF_WARPS + (threadIdx.x >> 5);
    if (n >= N) return;
    const int b = blockIdx.y;
    float x = __ldg(uv + ((size_t)b * 2 + 0) * N + n), y = __ldg(uv + ((size_t)b * 2 + 1) * N + n);
    // grid_sample round trip of utils.py:262-267: normalise to [-1,1], then un-normalise (align_corners)
    // (torch divides a CUDA tensor by a host scalar as a multiplication by its fp32 reciprocal, and every
    // step of the reference is a separately rounded elementwise kernel: no FMA contraction here)
    x = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(2.0f * x, __frcp_rn((float)(W - 1))), 1.0f), 1.f), 0.5f), (float)(W - 1));
    y = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(2.0f * y, __frcp_rn((float)(H - 1))), 1.0f), 1.f), 0.5f), (float)(H - 1));
    x = fminf(fmaxf(x, -2.f), (float)W + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)H + 1.f);
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf;
    const float tx = x - xf, ty = y - yf;
    const float wnw = (1.f - tx) * (1.f - ty), wne = tx * (1.f - ty), wsw = (1.f - tx) * ty, wse = tx * ty;
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
    const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
    const float* fb = feat + (size_t)b * H * W * C;
    const float* pnw = fb + ((size_t)y0 * W + x0) * C;
    const float* pne = pnw + C;
    const float* psw = pnw + (size_t)W * C;
    const float* pse = psw + C;
    float* orow = out + ((size_t)b * N + n) * ldo;
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
        if (vy0 && vx0) acc = fmaf(__ldg(pnw + c), wnw, acc);
        if (vy0 && vx1) acc = fmaf(__ldg(pne + c), wne, acc);
        if (vy1 && vx0) acc = fmaf(__ldg(psw + c), wsw, acc);
        if (vy1 && vx1) acc = fmaf(__ldg(pse + c), wse, acc);
        orow[c] = acc;
    }
}

}  // namespace

extern "C" int camli_clfm_interp(int B, int H, int W, int N, int C, const float* uv, const int64_t* nn_idx,
                                 const float* feat3d_rows, int64_t ld_feat,
                                 const float* W1, const float* b1, const float* W2, const float* b2,
                                 float* out_rows, void* stream) {
    if (B < 0 || H < 1 || W < 1 || N < 1 || C < 1 || ld_feat < C) return CAMLI_EINVAL;
    if (B > 65535 || C > 2048) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!uv || !nn_idx || !feat3d_rows || !W1 || !b1 || !W2 || !b2 || !out_rows) return CAMLI_EINVAL;
    const size_t smem = (size_t)(CF_H + 1) * C * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(clfm_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(camli_div_up(H * W, CF_WARPS), B);
    clfm_interp_kernel<<<grid, CF_WARPS * 32, smem, (cudaStream_t)stream>>>(H * W, W, N, C, uv, nn_idx, feat3d_rows, ld_feat,
                                                                           W1, b1, W2, b2, out_rows);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_bilinear_sample_rows(int B, int H, int W, int N, int C, const float* feat_nhwc, const float* uv,
                                          float* out_rows, int64_t ld_out, void* stream) {
    if (B < 0 || H < 2 || W < 2 || N < 0 || C < 1 || ld_out < C) return CAMLI_EINVAL;
    if (B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || N == 0) return CAMLI_OK;
    if (!feat_nhwc || !uv || !out_rows) return CAMLI_EINVAL;
    dim3 grid(camli_div_up(N, CF_WARPS), B);
    bilinear_sample_rows_kernel<<<grid, CF_WARPS * 32, 0, (cudaStream_t)stream>>>(H, W, N, C, feat_nhwc, uv, out_rows, ld_out);
    CAMLI_RETURN_LAUNCH_STATUS();
}
