// Stride-1 "same" convolution with a handful of output channels (C_out <= 4) on channel-last activations.
//
// The flow heads end in such a layer (reference models/raft_core.py:176: Conv2d(256, 2, 3, padding=1);
// models/camliraft_l_core.py:110: Conv1d(64, 3, 1)).  On the tensor-core kernel a 2-channel output still costs
// a full 32-column accumulator tile per k-step; the layer is 38 MFLOP and purely a read of the input, so it
// runs on the CUDA cores: one warp per pixel, lanes sweep the input channels with 128-bit loads (the 9 taps
// of neighbouring pixels hit L1/L2), weights staged once per CTA in shared memory, a shuffle reduction per
// output channel, bias + activation fused.  fp32 FMA throughout (exactly the reference's arithmetic type).
#include "common.cuh"

namespace {

constexpr int CS_WARPS = 8;
constexpr int CS_MAX_OUT = 4;

__device__ __forceinline__ float cs_activate_base(float v, int act, float slope);

// act may carry CAMLI_ACT_FIX_NONFINITE: torch.nan_to_num of the activated value
__device__ __forceinline__ float cs_activate(float v, int act, float slope) {
    const float r = cs_activate_base(v, act & (CAMLI_ACT_FIX_NONFINITE - 1), slope);
    if (!(act & CAMLI_ACT_FIX_NONFINITE)) return r;
    if (isnan(r)) return 0.f;
    if (isinf(r)) return r > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
    return r;
}

__device__ __forceinline__ float cs_activate_base(float v, int act, float slope) {
    switch (act) {
        case CAMLI_ACT_RELU: return fmaxf(v, 0.f);
        case CAMLI_ACT_LEAKY: return v > 0.f ? v : v * slope;
        case CAMLI_ACT_TANH: return tanhf(v);
        case CAMLI_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

template <int NOUT>
__global__ void __launch_bounds__(CS_WARPS * 32)
conv_small_n_kernel(const float* __restrict__ x, int B, int H, int W, int Cin, long long ldx,
                    const float* __restrict__ w,          // [NOUT, kh*kw*Cin] (OHWI)
                    const float* __restrict__ bias, int kh, int kw, int act, float slope,
                    float* __restrict__ out, long long ldo) {
    extern __shared__ float4 s_w[];                       // [NOUT][kh*kw*Cin/4]
    const int c4n = Cin >> 2, taps = kh * kw;
    const int per_out = taps * c4n;
    for (int e = threadIdx.x; e < NOUT * per_out; e += CS_WARPS * 32) s_w[e] = __ldg(reinterpret_cast<const float4*>(w) + e);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long n_pix = (long long)B * H * W;
    const int pad_y = kh >> 1, pad_x = kw >> 1;
    for (long long p = (long long)blockIdx.x * CS_WARPS + (threadIdx.x >> 5); p < n_pix; p += (long long)gridDim.x * CS_WARPS) {
        const int xw = (int)(p % W), yh = (int)((p / W) % H);
        float acc[NOUT];
#pragma unroll
        for (int n = 0; n < NOUT; ++n) acc[n] = 0.f;
        for (int t = 0; t < taps; ++t) {
            const int yy = yh + t / kw - pad_y, xx = xw + t % kw - pad_x;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;          // zero padding (warp-uniform)
            const float4* px = reinterpret_cast<const float4*>(x + (p + (long long)(yy - yh) * W + (xx - xw)) * ldx);
            for (int c = lane; c < c4n; c += 32) {
                const float4 v = __ldg(px + c);
#pragma unroll
                for (int n = 0; n < NOUT; ++n) {
                    const float4 q = s_w[n * per_out + t * c4n + c];
                    acc[n] = fmaf(v.x, q.x, fmaf(v.y, q.y, fmaf(v.z, q.z, fmaf(v.w, q.w, acc[n]))));
                }
            }
        }
#pragma unroll
        for (int n = 0; n < NOUT; ++n) acc[n] = camli_warp_sum(acc[n]);
        if (lane == 0) {
#pragma unroll
            for (int n = 0; n < NOUT; ++n) out[p * ldo + n] = cs_activate(acc[n] + (bias ? __ldg(bias + n) : 0.f), act, slope);
        }
    }
}

}  // namespace

extern "C" int camli_conv_small_n(const float* x, int B, int H, int W, int Cin, int64_t ldx, const float* w, int Cout,
                                  int kh, int kw, const float* bias, int act, float slope, float* out, int64_t ldo,
                                  void* stream) {
    if (B < 0 || H < 1 || W < 1 || Cin < 1 || Cout < 1 || kh < 1 || kw < 1 || ldx < Cin || ldo < Cout) return CAMLI_EINVAL;
    if (act < CAMLI_ACT_NONE || (act & ~CAMLI_ACT_FIX_NONFINITE) > CAMLI_ACT_SIGMOID) return CAMLI_EINVAL;
    if (Cout > CS_MAX_OUT || (kh & 1) == 0 || (kw & 1) == 0 || (Cin & 3) || (ldx & 3)) return CAMLI_EUNSUPPORTED;
    const size_t smem = (size_t)Cout * kh * kw * Cin * sizeof(float);
    if (smem > 160 * 1024) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!x || !w || !out) return CAMLI_EINVAL;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) return CAMLI_EINVAL;
    const long long n_pix = (long long)B * H * W;
    const unsigned grid = (unsigned)(camli_div_up_ll(n_pix, CS_WARPS) < 148LL * 8 ? camli_div_up_ll(n_pix, CS_WARPS) : 148LL * 8);
    cudaStream_t st = (cudaStream_t)stream;
#define CAMLI_CS_LAUNCH(N)                                                                                              \
    do {                                                                                                                \
        cudaError_t e = cudaFuncSetAttribute(conv_small_n_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return (int)e;                                                                            \
        conv_small_n_kernel<N><<<grid, CS_WARPS * 32, smem, st>>>(x, B, H, W, Cin, ldx, w, bias, kh, kw, act, slope, out, ldo); \
    } while (0)
    switch (Cout) {
        case 1: CAMLI_CS_LAUNCH(1); break;
        case 2: CAMLI_CS_LAUNCH(2); break;
        case 3: CAMLI_CS_LAUNCH(3); break;
        default: CAMLI_CS_LAUNCH(4); break;
    }
#undef CAMLI_CS_LAUNCH
    CAMLI_RETURN_LAUNCH_STATUS();
}

// ---------------------------------------------------------------------------------------------
// Stride-1 "same" convolution with a handful of INPUT channels (C_in <= 4) and many outputs: the first layer
// of the flow path of the motion encoder (reference models/raft_core.py:154: Conv2d(2, 128, 7, padding=3)).
// 100 MFLOP and 65 KB of input: the tensor-core kernel would spend 49 taps x one 94 %-empty k-block on it, cuDNN
// takes 35 us with its bias / ReLU passes.  CUDA cores: a CTA owns an 8 x 32 pixel tile and 32 output channels;
// the (8 + kh - 1) x (32 + kw - 1) x C_in input halo and the 32 x taps x C_in weights sit in shared memory, every
// thread keeps its pixel's 32 accumulators in registers (weight reads are warp-wide broadcasts), bias +
// activation fused, one 128-byte row segment written per thread.
namespace {

constexpr int CI_TH = 8, CI_TW = 32, CI_NOUT = 32, CI_MAX_CIN = 4, CI_MAX_K = 7;

__global__ void __launch_bounds__(CI_TH * CI_TW)
conv_small_cin_kernel(const float* __restrict__ x, int H, int W, int Cin, long long ldx,
                      const float* __restrict__ w,          // [Cout, kh*kw*Cin] (OHWI)
                      const float* __restrict__ bias, int Cout, int kh, int kw, int act, float slope,
                      float* __restrict__ out, long long ldo) {
    extern __shared__ float s_mem[];
    const int taps_c = kh * kw * Cin;
    const int hh = CI_TH + kh - 1, hw = CI_TW + kw - 1;
    float* s_w = s_mem;                                   // [taps_c][32]: transposed, outputs contiguous
    float* s_in = s_mem + taps_c * CI_NOUT;               // [hh][hw][Cin]
    const int tiles_x = (W + CI_TW - 1) / CI_TW;
    const int tx0 = (blockIdx.x % tiles_x) * CI_TW, ty0 = (blockIdx.x / tiles_x) * CI_TH;
    const int n0 = blockIdx.y * CI_NOUT, b = blockIdx.z;
    const int t = threadIdx.x;
    for (int e = t; e < taps_c * CI_NOUT; e += CI_TH * CI_TW) {
        const int n = e / taps_c, j = e - n * taps_c;
        s_w[j * CI_NOUT + n] = (n0 + n < Cout) ? __ldg(w + (size_t)(n0 + n) * taps_c + j) : 0.f;
    }
    const int pad_y = kh >> 1, pad_x = kw >> 1;
    for (int e = t; e < hh * hw * Cin; e += CI_TH * CI_TW) {
        const int c = e % Cin, xx = (e / Cin) % hw + tx0 - pad_x, yy = e / (Cin * hw) + ty0 - pad_y;
        s_in[e] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + (((size_t)b * H + yy) * W + xx) * ldx + c) : 0.f;
    }
    __syncthreads();
    const int px = t % CI_TW, py = t / CI_TW;
    float acc[CI_NOUT];
#pragma unroll
    for (int n = 0; n < CI_NOUT; ++n) acc[n] = 0.f;
    for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx) {
            const float* ip = s_in + ((py + ky) * hw + px + kx) * Cin;
            const float4* wp = reinterpret_cast<const float4*>(s_w + (ky * kw + kx) * Cin * CI_NOUT);
            for (int c = 0; c < Cin; ++c) {
                const float v = ip[c];
#pragma unroll
                for (int n4 = 0; n4 < CI_NOUT / 4; ++n4) {
                    const float4 q = wp[c * (CI_NOUT / 4) + n4];
                    acc[n4 * 4 + 0] = fmaf(v, q.x, acc[n4 * 4 + 0]);
                    acc[n4 * 4 + 1] = fmaf(v, q.y, acc[n4 * 4 + 1]);
                    acc[n4 * 4 + 2] = fmaf(v, q.z, acc[n4 * 4 + 2]);
                    acc[n4 * 4 + 3] = fmaf(v, q.w, acc[n4 * 4 + 3]);
                }
            }
        }
    const int xo = tx0 + px, yo = ty0 + py;
    if (xo >= W || yo >= H) return;
    float* o = out + (((size_t)b * H + yo) * W + xo) * ldo + n0;
    const bool vec = ((reinterpret_cast<uintptr_t>(o) & 15) == 0) && n0 + CI_NOUT <= Cout;
#pragma unroll
    for (int n = 0; n < CI_NOUT; ++n)
        acc[n] = cs_activate(acc[n] + ((bias && n0 + n < Cout) ? __ldg(bias + n0 + n) : 0.f), act, slope);
    if (vec) {
#pragma unroll
        for (int n = 0; n < CI_NOUT; n += 4) *reinterpret_cast<float4*>(o + n) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
    } else {
#pragma unroll
        for (int n = 0; n < CI_NOUT; ++n)
            if (n0 + n < Cout) o[n] = acc[n];
    }
}

}  // namespace

extern "C" int camli_conv_small_cin(const float* x, int B, int H, int W, int Cin, int64_t ldx, const float* w, int Cout,
                                    int kh, int kw, const float* bias, int act, float slope, float* out, int64_t ldo,
                                    void* stream) {
    if (B < 0 || H < 1 || W < 1 || Cin < 1 || Cout < 1 || kh < 1 || kw < 1 || ldx < Cin || ldo < Cout) return CAMLI_EINVAL;
    if (act < CAMLI_ACT_NONE || (act & ~CAMLI_ACT_FIX_NONFINITE) > CAMLI_ACT_SIGMOID) return CAMLI_EINVAL;
    if (Cin > CI_MAX_CIN || (kh & 1) == 0 || (kw & 1) == 0 || kh > CI_MAX_K || kw > CI_MAX_K || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!x || !w || !out) return CAMLI_EINVAL;
    const size_t smem = ((size_t)kh * kw * Cin * CI_NOUT + (size_t)(CI_TH + kh - 1) * (CI_TW + kw - 1) * Cin) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(conv_small_cin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(camli_div_up(W, CI_TW) * camli_div_up(H, CI_TH), camli_div_up(Cout, CI_NOUT), B);
    conv_small_cin_kernel<<<grid, CI_TH * CI_TW, smem, (cudaStream_t)stream>>>(x, H, W, Cin, ldx, w, bias, Cout, kh, kw, act,
                                                                              slope, out, ldo);
    CAMLI_RETURN_LAUNCH_STATUS();
}
