// Fused pointwise / small-reduction kernels of the update block for sm_100a.
//
// (1) SK fusion tail (reference models/clfm.py:195-214, after the two `align` convolutions):
//     avg-pool of (a+b) -> Linear -> ReLU -> Linear -> Sigmoid -> softmax over the branch pair ->
//     a*w1 + b*w2.  The reference spends ~13 launches on it, twice per CLFM, four CLFMs per GRU
//     iteration; here it is three: partial sums, one tiny per-sample FC kernel, one blend pass.
//     Inputs are the PRE-activation outputs of the align layers; their leaky-ReLU is applied on load.
// (2) ConvGRU gates (reference models/raft_core.py:123-139): after the merged z|r convolution
//     z = sigmoid(.), r = sigmoid(.), and [r*h | x] is assembled for the q convolution in one pass;
//     after the q convolution h' = (1-z)*h + z*tanh(q) (+ nan_to_num on the last half).
// All tensors are channel-last rows [B, P, C] (NHWC maps or point rows).
#include "common.cuh"

namespace {

constexpr int SK_SPLITS = 32;

// partial[b, s, c] = sum over the s-th slice of positions of leaky(a) + leaky(b)
__global__ void __launch_bounds__(256)
sk_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int P, int C, float slope,
                  float* __restrict__ partial) {
    __shared__ float s_red[8][33];
    const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, s = blockIdx.y, bi = blockIdx.z;
    const int per = (P + SK_SPLITS - 1) / SK_SPLITS;
    const int p0 = s * per, p1 = min(P, p0 + per);
    float acc = 0.f;
    if (c < C) {
        const float* pa = a + (size_t)bi * P * C + c;
        const float* pb = b + (size_t)bi * P * C + c;
        for (int p = p0 + ry; p < p1; p += 8)
            acc += camli_leaky(__ldg(pa + (size_t)p * C), slope) + camli_leaky(__ldg(pb + (size_t)p * C), slope);
    }
    s_red[ry][lane] = acc;
    __syncthreads();
    if (ry == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_red[i][lane];
        partial[((size_t)bi * SK_SPLITS + s) * C + c] = t;
    }
}

// The two tiny FCs are latency-bound when one CTA walks the whole weight matrix (420 KB for the
// 324-channel correlation CLFM), so each gets a grid of its own: one warp per output row.
//
// mid[b, j] = relu( <w_mid[j, :], pooled[b, :]> ),  pooled = sum of the partial slices / P
__global__ void __launch_bounds__(256)
sk_mid_kernel(const float* __restrict__ partial, int P, int C, int Cm, const float* __restrict__ w_mid,   // [Cm, C]
              float* __restrict__ mid) {                                                                  // [B, Cm]
    extern __shared__ float s_pool[];                 // [C]
    const int bi = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_p = 1.f / (float)P;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
#pragma unroll 8
        for (int s = 0; s < SK_SPLITS; ++s) t += __ldg(partial + ((size_t)bi * SK_SPLITS + s) * C + c);
        s_pool[c] = t * inv_p;
    }
    __syncthreads();
    const int j = blockIdx.x * 8 + warp;
    if (j >= Cm) return;
    float t = 0.f;
    for (int c = lane; c < C; c += 32) t = fmaf(__ldg(w_mid + (size_t)j * C + c), s_pool[c], t);
    t = camli_warp_sum(t);
    if (lane == 0) mid[(size_t)bi * Cm + j] = fmaxf(t, 0.f);
}

// weights[b, c, :] = softmax( sigmoid(<w_out[2c, :], mid>), sigmoid(<w_out[2c+1, :], mid>) )
__global__ void __launch_bounds__(256)
sk_out_kernel(const float* __restrict__ mid, int C, int Cm, const float* __restrict__ w_out,   // [2C, Cm]
              float* __restrict__ weights) {                                                   // [B, C, 2]
    extern __shared__ float s_mid[];                  // [Cm]
    const int bi = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < Cm; j += blockDim.x) s_mid[j] = __ldg(mid + (size_t)bi * Cm + j);
    __syncthreads();
    const int c = blockIdx.x * 8 + warp;
    if (c >= C) return;
    float tu = 0.f, tv = 0.f;
    for (int j = lane; j < Cm; j += 32) {
        tu = fmaf(__ldg(w_out + (size_t)(2 * c) * Cm + j), s_mid[j], tu);
        tv = fmaf(__ldg(w_out + (size_t)(2 * c + 1) * Cm + j), s_mid[j], tv);
    }
    tu = camli_warp_sum(tu);
    tv = camli_warp_sum(tv);
    if (lane == 0) {
        const float u = 1.f / (1.f + expf(-tu)), v = 1.f / (1.f + expf(-tv));
        const float m = fmaxf(u, v);
        const float eu = expf(u - m), ev = expf(v - m);
        const float inv = 1.f / (eu + ev);
        weights[((size_t)bi * C + c) * 2] = eu * inv;
        weights[((size_t)bi * C + c) * 2 + 1] = ev * inv;
    }
}

// out = leaky(a) * w[b,c,0] + leaky(b) * w[b,c,1]; VEC = 4 when C % 4 == 0 (128-bit accesses)
template <int VEC>
__global__ void __launch_bounds__(256)
sk_blend_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ weights,
                size_t PC, int C, float slope, float* __restrict__ out, long long ldo) {   // out rows of pitch ldo (>= C); PC < 2^32
    const int bi = blockIdx.y;
    const float2* w = reinterpret_cast<const float2*>(weights) + (size_t)bi * C;
    const size_t nv = PC / VEC;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned e = (unsigned)(i * VEC), row = e / (unsigned)C;
        const int c = (int)(e - row * (unsigned)C);
        const size_t g = (size_t)bi * PC + e;
        float* o_ptr = out + ((size_t)bi * (PC / C) + row) * ldo + c;
        if (VEC == 4) {
            const float4 av = __ldcs(reinterpret_cast<const float4*>(a + g)), bv = __ldcs(reinterpret_cast<const float4*>(b + g));
            const float2 w0 = __ldg(w + c), w1 = __ldg(w + c + 1), w2 = __ldg(w + c + 2), w3 = __ldg(w + c + 3);
            float4 o;
            o.x = camli_leaky(av.x, slope) * w0.x + camli_leaky(bv.x, slope) * w0.y;
            o.y = camli_leaky(av.y, slope) * w1.x + camli_leaky(bv.y, slope) * w1.y;
            o.z = camli_leaky(av.z, slope) * w2.x + camli_leaky(bv.z, slope) * w2.y;
            o.w = camli_leaky(av.w, slope) * w3.x + camli_leaky(bv.w, slope) * w3.y;
            *reinterpret_cast<float4*>(o_ptr) = o;
        } else {
            const float2 w0 = __ldg(w + c);
            *o_ptr = camli_leaky(__ldg(a + g), slope) * w0.x + camli_leaky(__ldg(b + g), slope) * w0.y;
        }
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// zr [R, 2H] (z then r, pre-activation), h [R, H], x [R, X]  ->  z [R, H], rhx [R, H+X] = [r*h | x]
__global__ void __launch_bounds__(256)
gru_gate_kernel(const float* __restrict__ zr, const float* __restrict__ h, const float* __restrict__ x,
                size_t R, int H, int X, float* __restrict__ z, float* __restrict__ rhx) {
    const int W = H + X;
    const size_t total = R * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / W;
        const int c = (int)(i - r * W);
        if (c < H) {
            const float hv = __ldg(h + r * H + c);
            z[r * H + c] = sigmoidf_(__ldg(zr + r * 2 * H + c));
            rhx[i] = sigmoidf_(__ldg(zr + r * 2 * H + H + c)) * hv;
        } else {
            rhx[i] = __ldg(x + r * X + (c - H));
        }
    }
}

// h' = (1 - z) * h + z * tanh(q)   (optionally nan_to_num: nan -> 0, +-inf -> +-FLT_MAX)
__global__ void __launch_bounds__(256)
gru_update_kernel(const float* __restrict__ z, const float* __restrict__ h, const float* __restrict__ q, size_t n,
                  int fix_nonfinite, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float zv = __ldg(z + i), hv = __ldg(h + i);
        float v = (1.f - zv) * hv + zv * tanhf(__ldg(q + i));
        if (fix_nonfinite) {
            if (isnan(v)) v = 0.f;
            else if (isinf(v)) v = v > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
        }
        out[i] = v;
    }
}

inline unsigned grid_for(size_t n) {
    const size_t blocks = (n + 255) / 256;
    return (unsigned)(blocks < 148u * 16u ? (blocks ? blocks : 1) : 148u * 16u);
}

}  // namespace

extern "C" int camli_sk_fusion_tail(int B, int P, int C, int C_mid, const float* a_rows, const float* b_rows,
                                    float negative_slope, const float* w_mid, const float* w_out,
                                    float* partial_scratch, float* weights_scratch, float* out_rows, void* stream) {
    return camli_sk_fusion_tail_strided(B, P, C, C_mid, a_rows, b_rows, negative_slope, w_mid, w_out, partial_scratch,
                                        weights_scratch, out_rows, C, stream);
}

extern "C" int camli_sk_fusion_tail_strided(int B, int P, int C, int C_mid, const float* a_rows, const float* b_rows,
                                            float negative_slope, const float* w_mid, const float* w_out,
                                            float* partial_scratch, float* weights_scratch, float* out_rows, int64_t ld_out,
                                            void* stream) {
    if (B < 0 || P < 1 || C < 1 || C_mid < 1 || ld_out < C) return CAMLI_EINVAL;
    if ((long long)P * C > 4294967295LL) return CAMLI_EUNSUPPORTED;
    if (B > 65535 || (size_t)(C > C_mid ? C : C_mid) * sizeof(float) > 48 * 1024) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!a_rows || !b_rows || !w_mid || !w_out || !partial_scratch || !weights_scratch || !out_rows) return CAMLI_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    sk_partial_kernel<<<dim3(camli_div_up(C, 32), SK_SPLITS, B), 256, 0, st>>>(a_rows, b_rows, P, C, negative_slope,
                                                                            partial_scratch);
    // weights_scratch: [B, C, 2] blend weights followed by the [B, C_mid] hidden layer
    float* mid = weights_scratch + (size_t)B * C * 2;
    sk_mid_kernel<<<dim3(camli_div_up(C_mid, 8), B), 256, (size_t)C * sizeof(float), st>>>(partial_scratch, P, C, C_mid, w_mid, mid);
    sk_out_kernel<<<dim3(camli_div_up(C, 8), B), 256, (size_t)C_mid * sizeof(float), st>>>(mid, C, C_mid, w_out, weights_scratch);
    const size_t PC = (size_t)P * C;
    const bool vec = (C % 4 == 0) && (ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(a_rows) | reinterpret_cast<uintptr_t>(b_rows) |
                                                           reinterpret_cast<uintptr_t>(out_rows)) % 16 == 0);
    // batch b's rows start at out_rows + b * P * ld_out: the kernel's row index runs over all B * P rows
    if (vec) sk_blend_kernel<4><<<dim3(grid_for(PC / 4), B), 256, 0, st>>>(a_rows, b_rows, weights_scratch, PC, C, negative_slope, out_rows, ld_out);
    else     sk_blend_kernel<1><<<dim3(grid_for(PC), B), 256, 0, st>>>(a_rows, b_rows, weights_scratch, PC, C, negative_slope, out_rows, ld_out);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_gru_gate(int64_t rows, int H, int X, const float* zr, const float* h, const float* x, float* z,
                              float* rhx, void* stream) {
    if (rows < 0 || H < 1 || X < 0) return CAMLI_EINVAL;
    if (rows == 0) return CAMLI_OK;
    if (!zr || !h || (X > 0 && !x) || !z || !rhx) return CAMLI_EINVAL;
    gru_gate_kernel<<<grid_for((size_t)rows * (H + X)), 256, 0, (cudaStream_t)stream>>>(zr, h, x, (size_t)rows, H, X, z, rhx);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_gru_update(int64_t n, const float* z, const float* h, const float* q, int fix_nonfinite,
                                float* h_out, void* stream) {
    if (n < 0) return CAMLI_EINVAL;
    if (n == 0) return CAMLI_OK;
    if (!z || !h || !q || !h_out) return CAMLI_EINVAL;
    gru_update_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>(z, h, q, (size_t)n, fix_nonfinite, h_out);
    CAMLI_RETURN_LAUNCH_STATUS();
}
