// RAFT correlation lookup for sm_100a: all pyramid levels in one launch.
//
// Replaces Correlation2D.forward (reference models/raft_core.py:71-107): per level a
// [BHW,9,9,2] coordinate tensor + grid_sample + view, then cat + permute + contiguous.
//
// For every source pixel p and level l the (2r+1)^2 bilinear samples of the p-th volume
// slice V_l[p] (an H_l x W_l map) are taken at (x/2^l + a, y/2^l + b), a,b in [-r,r].
// All of them share one fractional offset, so they are exactly the 2x2-weighted sums over
// ONE (2r+2)^2 integer footprint: 100 loads instead of 324 per pixel and level.
// The op is HBM-bound: B*HW*L*(100+81)*4 bytes + coords.
//
// Mapping: a CTA owns TP=32 consecutive source pixels of one level.  Phase 1: all 256
// threads gather the 32 footprints (10-float row segments) into shared memory, zero outside
// the map (grid_sample's zero padding).  Phase 2: lane = pixel, warp = output channel
// (strided): each output is 4 conflict-free LDS (row stride 101 words) and 4 FMAs, and the
// store of one channel for 32 consecutive pixels is one coalesced 128-byte line of the
// NCHW output [B, L*81, H, W].
//
// Channel order follows the reference's quirk (raft_core.py:79-85): channel = l*81 + i*9 + j
// where i moves the x coordinate and j moves y.
#include "common.cuh"

namespace {

constexpr int LK_R = 4;
constexpr int LK_WIN = 2 * LK_R + 1;      // 9
constexpr int LK_FP = LK_WIN + 1;         // 10: footprint edge
constexpr int LK_TP = 32;                 // pixels per CTA
constexpr int LK_THREADS = 256;
constexpr int LK_STRIDE = LK_FP * LK_FP + 1;   // 101 words per pixel: conflict-free for lane = pixel
constexpr int LK_MAX_LEVELS = 8;

struct LookupLevels {
    const float* vol[LK_MAX_LEVELS];   // [B, HW, h, w]
    int h[LK_MAX_LEVELS], w[LK_MAX_LEVELS];
};

// (Forcing 8 CTAs per SM -- the whole 1020-CTA grid of a 68x120 map resident at once instead of 1.4 waves -- was
// measured and changes nothing: the kernel sits at ~2.5 TB/s of 64-byte-sector scattered DRAM reads either way.)
// Footprint load.  KEEP: tagged L2 evict-last.  The flow moves by a fraction of a pixel per refinement iteration, so
// iteration i+1 reads almost exactly the sectors iteration i read (~37 MB of the 354 MB pyramid); nothing else in the
// iteration is read twice, so keeping just these sectors resident in the 126 MB L2 turns every lookup after the first
// into an L2 read instead of 64-byte-granular scattered DRAM reads.
template <bool KEEP>
__device__ __forceinline__ float lk_load(const float* p, uint64_t policy) {
    if (!KEEP) return __ldg(p);
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy));
    return v;
}

template <bool NHWC, bool KEEP>
__global__ void __launch_bounds__(LK_THREADS)
corr2d_lookup_kernel(const __grid_constant__ LookupLevels lv, const float* __restrict__ coords,   // [B,2,HW]
                     float* __restrict__ out, int HW, int n_levels) {
    __shared__ float s_fp[LK_TP * LK_STRIDE];
    __shared__ float s_fx[LK_TP], s_fy[LK_TP];
    __shared__ int s_x0[LK_TP], s_y0[LK_TP];

    const int level = blockIdx.y, b = blockIdx.z;
    const int p0 = blockIdx.x * LK_TP;
    const int t = threadIdx.x;
    uint64_t policy = 0;
    if (KEEP) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    const int h = lv.h[level], w = lv.w[level];
    const float* __restrict__ vol = lv.vol[level] + (size_t)b * HW * h * w;

    if (t < LK_TP) {
        const int p = min(p0 + t, HW - 1);
        const float scale = 1.f / (float)(1 << level);          // coords / 2**i (exact)
        const float x = __ldg(coords + ((size_t)b * 2 + 0) * HW + p) * scale;
        const float y = __ldg(coords + ((size_t)b * 2 + 1) * HW + p) * scale;
        // clamp far-away centres so the int conversion is safe; anything beyond the map + window
        // samples zeros either way
        const float xc = fminf(fmaxf(x, -16.f), (float)w + 16.f);
        const float yc = fminf(fmaxf(y, -16.f), (float)h + 16.f);
        const float xf = floorf(xc), yf = floorf(yc);
        s_fx[t] = xc - xf;
        s_fy[t] = yc - yf;
        s_x0[t] = (int)xf - LK_R;
        s_y0[t] = (int)yf - LK_R;
    }
    __syncthreads();

    // ---- phase 1: footprints -> shared memory.  A warp owns 4 pixels; lane l handles footprint elements
    // l, l+32, l+64, l+96 of each, so the (row, column) of an element is a per-thread constant (no index
    // arithmetic in the loop) and a warp-wide load covers ~3 contiguous 40-byte row segments.  All 16 loads of
    // a thread are issued before the first shared-memory store (latency-bound otherwise).
    const int lane = t & 31, warp = t >> 5;
    constexpr int LK_PXW = LK_TP / (LK_THREADS / 32);       // 4 pixels per warp
    constexpr int LK_EPT = (LK_FP * LK_FP + 31) / 32;       // 4 elements per lane and pixel
    int off[LK_EPT], ry[LK_EPT], rx[LK_EPT];
#pragma unroll
    for (int i = 0; i < LK_EPT; ++i) {
        const int e = lane + 32 * i;
        ry[i] = e / LK_FP;
        rx[i] = e - ry[i] * LK_FP;
        off[i] = e < LK_FP * LK_FP ? e : -1;
    }
    float v[LK_PXW][LK_EPT];
#pragma unroll
    for (int q = 0; q < LK_PXW; ++q) {
        const int px = warp * LK_PXW + q, p = p0 + px;
        const int y0 = s_y0[px], x0 = s_x0[px];
        const float* __restrict__ slice = vol + (size_t)min(p, HW - 1) * h * w;
#pragma unroll
        for (int i = 0; i < LK_EPT; ++i) {
            const int yy = y0 + ry[i], xx = x0 + rx[i];
            v[q][i] = 0.f;
            if (off[i] >= 0 && p < HW && (unsigned)yy < (unsigned)h && (unsigned)xx < (unsigned)w)
                v[q][i] = lk_load<KEEP>(slice + yy * w + xx, policy);
        }
    }
#pragma unroll
    for (int q = 0; q < LK_PXW; ++q)
#pragma unroll
        for (int i = 0; i < LK_EPT; ++i)
            if (off[i] >= 0) s_fp[(warp * LK_PXW + q) * LK_STRIDE + off[i]] = v[q][i];

    const int n_ch = n_levels * LK_WIN * LK_WIN;
    if (NHWC) {
        // ---- phase 2 (NHWC out): the same warp turns its 4 footprints into 4 x 81 outputs; lane = window
        // position (c, c+32, c+64), whose (i -> x, j -> y) split is again a per-thread constant; a pixel's 81
        // channels of this level are contiguous in the output row
        __syncwarp();
        constexpr int LK_CPT = (LK_WIN * LK_WIN + 31) / 32;  // 3
        int qoff[LK_CPT];
#pragma unroll
        for (int j = 0; j < LK_CPT; ++j) {
            const int c = lane + 32 * j;
            const int ci = c / LK_WIN, cj = c - ci * LK_WIN;        // ci -> x offset, cj -> y offset
            qoff[j] = c < LK_WIN * LK_WIN ? cj * LK_FP + ci : -1;
        }
#pragma unroll
        for (int q = 0; q < LK_PXW; ++q) {
            const int px = warp * LK_PXW + q, p = p0 + px;
            if (p >= HW) break;
            const float fx = s_fx[px], fy = s_fy[px];
            const float w00 = (1.f - fx) * (1.f - fy), w10 = fx * (1.f - fy), w01 = (1.f - fx) * fy, w11 = fx * fy;
            const float* fp = s_fp + px * LK_STRIDE;
            float* __restrict__ o = out + ((size_t)b * HW + p) * n_ch + level * LK_WIN * LK_WIN;
#pragma unroll
            for (int j = 0; j < LK_CPT; ++j) {
                if (qoff[j] < 0) continue;
                const float* qp = fp + qoff[j];
                float r = qp[0] * w00;
                r = fmaf(qp[1], w10, r);
                r = fmaf(qp[LK_FP], w01, r);
                r = fmaf(qp[LK_FP + 1], w11, r);
                o[lane + 32 * j] = r;
            }
        }
        return;
    }
    __syncthreads();
    // ---- phase 2 (NCHW out): lane = pixel, warp strides over the 81 window positions
    const int p = p0 + lane;
    if (p >= HW) return;
    const float fx = s_fx[lane], fy = s_fy[lane];
    const float w00 = (1.f - fx) * (1.f - fy), w10 = fx * (1.f - fy), w01 = (1.f - fx) * fy, w11 = fx * fy;
    const float* fp = s_fp + lane * LK_STRIDE;
    float* __restrict__ o = out + ((size_t)b * n_ch + (size_t)level * LK_WIN * LK_WIN) * HW + p;
    for (int c = warp; c < LK_WIN * LK_WIN; c += LK_THREADS / 32) {
        const int i = c / LK_WIN, j = c - i * LK_WIN;          // i -> x offset, j -> y offset
        const float* q = fp + j * LK_FP + i;
        float v = q[0] * w00;
        v = fmaf(q[1], w10, v);
        v = fmaf(q[LK_FP], w01, v);
        v = fmaf(q[LK_FP + 1], w11, v);
        o[(size_t)c * HW] = v;
    }
}

// ---------------------------------------------------------------- pyramid pooling
// All coarser levels of the volume pyramid from ONE read of level 0 (the reference runs three
// avg_pool2d passes, each re-reading the previous level from HBM: raft_core.py:65-68).  A CTA
// owns one source pixel's h0 x w0 slice; level 1 is reduced straight from global memory, the
// deeper levels from shared memory.
constexpr int PP_THREADS = 256;

struct PoolLevels {
    float* out[LK_MAX_LEVELS];          // level l >= 1: [rows, h_l, w_l]
    int h[LK_MAX_LEVELS], w[LK_MAX_LEVELS];
};

__global__ void __launch_bounds__(PP_THREADS)
corr2d_pool_pyramid_kernel(const float* __restrict__ vol0, const __grid_constant__ PoolLevels lv, int n_levels) {
    extern __shared__ float s_buf[];    // two ping-pong maps of h1*w1 floats
    const size_t row = blockIdx.x;
    const int h0 = lv.h[0], w0 = lv.w[0];
    const float* __restrict__ src = vol0 + row * (size_t)h0 * w0;
    float* cur = s_buf;
    float* nxt = s_buf + lv.h[1] * lv.w[1];
    {
        const int h1 = lv.h[1], w1 = lv.w[1];
        float* __restrict__ dst = lv.out[1] + row * (size_t)h1 * w1;
        const bool vec = (w0 & 1) == 0;
        for (int e = threadIdx.x; e < h1 * w1; e += PP_THREADS) {
            const int y = e / w1, x = e - y * w1;
            const float* r0 = src + (size_t)(2 * y) * w0 + 2 * x;
            float a, b2, c, d;
            if (vec) {
                const float2 u = __ldcs(reinterpret_cast<const float2*>(r0));
                const float2 v = __ldcs(reinterpret_cast<const float2*>(r0 + w0));
                a = u.x; b2 = u.y; c = v.x; d = v.y;
            } else {
                a = __ldcs(r0); b2 = __ldcs(r0 + 1); c = __ldcs(r0 + w0); d = __ldcs(r0 + w0 + 1);
            }
            const float m = (a + b2 + c + d) * 0.25f;
            cur[e] = m;
            dst[e] = m;
        }
    }
    for (int l = 2; l < n_levels; ++l) {
        __syncthreads();
        const int hp = lv.h[l - 1], wp = lv.w[l - 1], hl = lv.h[l], wl = lv.w[l];
        (void)hp;
        float* __restrict__ dst = lv.out[l] + row * (size_t)hl * wl;
        for (int e = threadIdx.x; e < hl * wl; e += PP_THREADS) {
            const int y = e / wl, x = e - y * wl;
            const float* r0 = cur + (2 * y) * wp + 2 * x;
            const float m = (r0[0] + r0[1] + r0[wp] + r0[wp + 1]) * 0.25f;
            nxt[e] = m;
            dst[e] = m;
        }
        float* t2 = cur; cur = nxt; nxt = t2;
    }
}

}  // namespace

// 1: footprint loads carry an L2 evict-last policy (see lk_load); 0 (default): plain read-only loads.  Measured on the
// C2 graph (12 lookups, CUPTI): 13-17 us per lookup either way -- the hint does not turn the revisited sectors into L2
// hits across the ~100 MB of other traffic of an iteration, so it stays off.
static int camli_lookup_keep_l2 = 0;
extern "C" int camli_corr2d_lookup_set_l2_keep(int enabled) {
    const int old = camli_lookup_keep_l2;
    camli_lookup_keep_l2 = enabled ? 1 : 0;
    return old;
}

extern "C" int camli_corr2d_lookup(const float* const* volumes, const int* level_h, const int* level_w,
                                   int n_levels, const float* coords, float* out, int B, int H, int W,
                                   int radius, int out_nhwc, void* stream) {
    if (B < 0 || H < 1 || W < 1 || n_levels < 1) return CAMLI_EINVAL;
    if (radius != LK_R || n_levels > LK_MAX_LEVELS || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!volumes || !level_h || !level_w || !coords || !out) return CAMLI_EINVAL;
    LookupLevels lv;
    for (int l = 0; l < n_levels; ++l) {
        if (!volumes[l] || level_h[l] < 1 || level_w[l] < 1) return CAMLI_EINVAL;
        lv.vol[l] = volumes[l]; lv.h[l] = level_h[l]; lv.w[l] = level_w[l];
    }
    const int HW = H * W;
    dim3 grid(camli_div_up(HW, LK_TP), n_levels, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (camli_lookup_keep_l2) {
        if (out_nhwc) corr2d_lookup_kernel<true, true><<<grid, LK_THREADS, 0, st>>>(lv, coords, out, HW, n_levels);
        else          corr2d_lookup_kernel<false, true><<<grid, LK_THREADS, 0, st>>>(lv, coords, out, HW, n_levels);
    } else {
        if (out_nhwc) corr2d_lookup_kernel<true, false><<<grid, LK_THREADS, 0, st>>>(lv, coords, out, HW, n_levels);
        else          corr2d_lookup_kernel<false, false><<<grid, LK_THREADS, 0, st>>>(lv, coords, out, HW, n_levels);
    }
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_corr2d_pool_pyramid(const float* vol0, float* const* coarser_host, int n_levels, int64_t rows,
                                         int h0, int w0, void* stream) {
    if (rows < 0 || h0 < 1 || w0 < 1 || n_levels < 1) return CAMLI_EINVAL;
    if (n_levels > LK_MAX_LEVELS || rows > 2147483647LL) return CAMLI_EUNSUPPORTED;
    if (rows == 0 || n_levels == 1) return CAMLI_OK;
    if (!vol0 || !coarser_host) return CAMLI_EINVAL;
    PoolLevels lv;
    lv.h[0] = h0; lv.w[0] = w0; lv.out[0] = nullptr;
    for (int l = 1; l < n_levels; ++l) {
        lv.h[l] = lv.h[l - 1] / 2; lv.w[l] = lv.w[l - 1] / 2;
        if (lv.h[l] < 1 || lv.w[l] < 1) return CAMLI_EUNSUPPORTED;
        if (!coarser_host[l - 1]) return CAMLI_EINVAL;
        lv.out[l] = coarser_host[l - 1];
    }
    const size_t smem = 2 * (size_t)lv.h[1] * lv.w[1] * sizeof(float);
    if (smem > 200 * 1024) return CAMLI_EUNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(corr2d_pool_pyramid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    corr2d_pool_pyramid_kernel<<<(unsigned)rows, PP_THREADS, smem, (cudaStream_t)stream>>>(vol0, lv, n_levels);
    CAMLI_RETURN_LAUNCH_STATUS();
}
