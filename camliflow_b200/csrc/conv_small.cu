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

__device__ __forceinline__ float cs_activate(float v, int act, float slope) {
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
    if (act < CAMLI_ACT_NONE || act > CAMLI_ACT_SIGMOID) return CAMLI_EINVAL;
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
