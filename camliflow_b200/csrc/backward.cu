// Backward kernels of the fused hot-path operators (training path, BASELINE config 5) for sm_100a.
//
// The reference differentiates these stages through torch autograd over materialised intermediates
// (grid_sample / gather / max backward: models/raft_core.py:71-107, models/point_conv.py:119-128); the
// gradients are scatters whose structure the library knows, so they are written directly:
//   * RAFT correlation lookup: the gradient of volume slice V_l[p] comes only from pixel p's own 9x9 window,
//     i.e. one (2r+2)^2 footprint per (pixel, level), owned by exactly one CTA -> plain stores into a zeroed
//     volume, no atomics, bit-reproducible;
//   * PointConvDW gather-max: the arg-max neighbour of every (centroid, channel) is found again from the saved
//     inputs (cheaper than storing an index tensor in the forward); its feature row receives an atomic add,
//     its weight entry a plain store.
#include "common.cuh"

namespace {

constexpr int LB_R = 4;
constexpr int LB_WIN = 2 * LB_R + 1;      // 9
constexpr int LB_FP = LB_WIN + 1;         // 10
constexpr int LB_TP = 32;                 // pixels per CTA
constexpr int LB_THREADS = 256;
constexpr int LB_GSTRIDE = LB_WIN * LB_WIN + 2;   // 83 words per pixel
constexpr int LB_MAX_LEVELS = 8;

struct LookupGradLevels {
    float* vol[LB_MAX_LEVELS];            // [B, HW, h, w], zero-initialised by the caller
    int h[LB_MAX_LEVELS], w[LB_MAX_LEVELS];
};

__global__ void __launch_bounds__(LB_THREADS)
corr2d_lookup_backward_kernel(const __grid_constant__ LookupGradLevels lv, const float* __restrict__ coords,   // [B,2,HW]
                              const float* __restrict__ g,      // rows [B,HW,n_levels*81]
                              int HW, int n_levels) {
    __shared__ float s_g[LB_TP * LB_GSTRIDE];
    __shared__ float s_fx[LB_TP], s_fy[LB_TP];
    __shared__ int s_x0[LB_TP], s_y0[LB_TP];
    const int level = blockIdx.y, b = blockIdx.z;
    const int p0 = blockIdx.x * LB_TP;
    const int t = threadIdx.x;
    const int h = lv.h[level], w = lv.w[level];
    float* __restrict__ vol = lv.vol[level] + (size_t)b * HW * h * w;
    const int n_ch = n_levels * LB_WIN * LB_WIN;

    if (t < LB_TP) {                       // same centre arithmetic as the forward kernel (corr2d_lookup.cu)
        const int p = min(p0 + t, HW - 1);
        const float scale = 1.f / (float)(1 << level);
        const float x = __ldg(coords + ((size_t)b * 2 + 0) * HW + p) * scale;
        const float y = __ldg(coords + ((size_t)b * 2 + 1) * HW + p) * scale;
        const float xc = fminf(fmaxf(x, -16.f), (float)w + 16.f);
        const float yc = fminf(fmaxf(y, -16.f), (float)h + 16.f);
        const float xf = floorf(xc), yf = floorf(yc);
        s_fx[t] = xc - xf;
        s_fy[t] = yc - yf;
        s_x0[t] = (int)xf - LB_R;
        s_y0[t] = (int)yf - LB_R;
    }
    for (int e = t; e < LB_TP * LB_WIN * LB_WIN; e += LB_THREADS) {
        const int px = e / (LB_WIN * LB_WIN), c = e - px * (LB_WIN * LB_WIN);
        const int p = p0 + px;
        s_g[px * LB_GSTRIDE + c] = p < HW ? __ldg(g + ((size_t)b * HW + p) * n_ch + level * LB_WIN * LB_WIN + c) : 0.f;
    }
    __syncthreads();

    // footprint entry (ry, rx) collects the four window positions whose 2x2 stencil covers it; channel of
    // window position (i -> x offset, j -> y offset) is i*9 + j (the reference's meshgrid order)
    for (int e = t; e < LB_TP * LB_FP * LB_FP; e += LB_THREADS) {
        const int px = e / (LB_FP * LB_FP), a = e - px * (LB_FP * LB_FP);
        const int ry = a / LB_FP, rx = a - ry * LB_FP;
        const int p = p0 + px;
        const int yy = s_y0[px] + ry, xx = s_x0[px] + rx;
        if (p >= HW || yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
        const float fx = s_fx[px], fy = s_fy[px];
        const float* gp = s_g + px * LB_GSTRIDE;
        float acc = 0.f;
        if (rx < LB_WIN && ry < LB_WIN) acc = fmaf(gp[rx * LB_WIN + ry], (1.f - fx) * (1.f - fy), acc);
        if (rx > 0 && ry < LB_WIN) acc = fmaf(gp[(rx - 1) * LB_WIN + ry], fx * (1.f - fy), acc);
        if (rx < LB_WIN && ry > 0) acc = fmaf(gp[rx * LB_WIN + ry - 1], (1.f - fx) * fy, acc);
        if (rx > 0 && ry > 0) acc = fmaf(gp[(rx - 1) * LB_WIN + ry - 1], fx * fy, acc);
        vol[((size_t)p * h + yy) * w + xx] = acc;
    }
}

// One warp per (centroid, 32-channel chunk), as in the forward kernel.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
dw_gather_max_backward_kernel(int N, int S, int K, int k, int O, int chunks,
                              const float* __restrict__ feat,         // rows [B,N,O]
                              const float* __restrict__ wc,           // rows [B,S,k,O]
                              const int64_t* __restrict__ idx,        // [B,S,K]
                              const float* __restrict__ g,            // rows [B,S,O]
                              float* __restrict__ g_feat,             // rows [B,N,O], zero-initialised
                              float* __restrict__ g_w) {              // rows [B,S,k,O], zero-initialised, or null
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (wid >= (long long)S * chunks) return;
    const int s = (int)(wid / chunks), o = (int)(wid % chunks) * 32 + lane;
    const int b = blockIdx.y;
    const int64_t* ip = idx + ((size_t)b * S + s) * K;
    const int my = (lane < k) ? (int)__ldg(ip + lane) : 0;
    const float* fb = feat + (size_t)b * N * O;
    const float* wp = wc + ((size_t)b * S + s) * k * O;
    float best = -INFINITY, bf = 0.f, bw = 0.f;
    int bj = 0, bi = 0;
    for (int j = 0; j < k; ++j) {
        const int ij = __shfl_sync(CAMLI_FULL_MASK, my, j);
        if (o < O) {
            const float f = __ldg(fb + (size_t)ij * O + o), w = __ldg(wp + (size_t)j * O + o);
            const float v = f * w;
            if (v > best) { best = v; bf = f; bw = w; bj = j; bi = ij; }      // first maximum wins (torch.max)
        }
    }
    if (o >= O) return;
    const float go = __ldg(g + ((size_t)b * S + s) * O + o);
    atomicAdd(g_feat + ((size_t)b * N + bi) * O + o, go * bw);
    if (g_w) g_w[(((size_t)b * S + s) * k + bj) * O + o] = go * bf;
}

}  // namespace

extern "C" int camli_corr2d_lookup_backward(float* const* grad_volumes, const int* level_h, const int* level_w, int n_levels,
                                            const float* coords, const float* grad_out_rows, int B, int H, int W,
                                            int radius, void* stream) {
    if (B < 0 || H < 1 || W < 1 || n_levels < 1) return CAMLI_EINVAL;
    if (radius != LB_R || n_levels > LB_MAX_LEVELS || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!grad_volumes || !level_h || !level_w || !coords || !grad_out_rows) return CAMLI_EINVAL;
    LookupGradLevels lv;
    for (int l = 0; l < n_levels; ++l) {
        if (!grad_volumes[l] || level_h[l] < 1 || level_w[l] < 1) return CAMLI_EINVAL;
        lv.vol[l] = grad_volumes[l]; lv.h[l] = level_h[l]; lv.w[l] = level_w[l];
    }
    const int HW = H * W;
    dim3 grid(camli_div_up(HW, LB_TP), n_levels, B);
    corr2d_lookup_backward_kernel<<<grid, LB_THREADS, 0, (cudaStream_t)stream>>>(lv, coords, grad_out_rows, HW, n_levels);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_pointconv_dw_gather_max_backward(int B, int N, int S, int K, int k, int O, const float* feat_rows,
                                                      const float* weights, const int64_t* knn_idx, const float* grad_out_rows,
                                                      float* grad_feat_rows, float* grad_weights, void* stream) {
    if (B < 0 || N < 1 || S < 0 || k < 1 || K < k || O < 1) return CAMLI_EINVAL;
    if (k > 32 || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || S == 0) return CAMLI_OK;
    if (!feat_rows || !weights || !knn_idx || !grad_out_rows || !grad_feat_rows) return CAMLI_EINVAL;
    constexpr int WARPS = 8;
    const int chunks = camli_div_up(O, 32);
    dim3 grid((unsigned)camli_div_up_ll((long long)S * chunks, WARPS), B);
    dw_gather_max_backward_kernel<WARPS><<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(
        N, S, K, k, O, chunks, feat_rows, weights, knn_idx, grad_out_rows, grad_feat_rows, grad_weights);
    CAMLI_RETURN_LAUNCH_STATUS();
}
