// ResNet stem for sm_100a: 7x7 stride-2 convolution (3 -> 64, eval-mode BatchNorm folded into weight / bias) + ReLU +
// 3x3 stride-2 max-pool in ONE kernel.
//
// Replaces, on the image branch of CamLiRAFT, `conv1 -> bn1 -> relu -> maxpool` of the mmdet ResNet-50 the reference
// inherits its encoder from (models/raft_core.py:10-22,35-38; three 544x960 images per frame pair).  With three input
// channels the layer is no tensor-core shape (a 32-channel k-block would be 91 % padding), so it runs on the CUDA
// cores in fp32 FMA -- the reference's arithmetic type -- as a direct convolution: a CTA owns an 8 x 16 tile of POOLED
// pixels, i.e. a 17 x 33 tile of convolution outputs and a 39 x 71 x 3 input halo; halo (split by column parity, so the
// stride-2 reads of neighbouring outputs are conflict-free) and the whole 64 x 147 weight
// matrix sit in shared memory, every thread keeps 2 pixels x 16 channels of accumulators in registers, the activated
// convolution tile goes to shared memory 16 channels at a time and is pooled from there.  The 33 MB / image full-
// resolution activation never touches HBM.
#include "common.cuh"

namespace {

constexpr int ST_K = 7, ST_CIN = 3, ST_COUT = 64, ST_TAPS = ST_K * ST_K * ST_CIN;     // 147
constexpr int ST_TPY = 8, ST_TPX = 16;                    // pooled tile
constexpr int ST_CY = 2 * ST_TPY + 1, ST_CX = 2 * ST_TPX + 1;     // 17 x 33 convolution outputs
constexpr int ST_IY = 2 * (ST_CY - 1) + ST_K, ST_IX = 2 * (ST_CX - 1) + ST_K;         // 39 x 71 input pixels
constexpr int ST_THREADS = 288;                           // 281 output pairs of the 17 x 33 tile: one pass
constexpr int ST_CHUNK = 16;                              // output channels per pass
constexpr int ST_NPIX = ST_CY * ST_CX;                    // 561
constexpr int ST_IXH = (ST_IX + 1) / 2;                   // 36: halo columns per parity plane
// halo layout [channel][column parity][row][column / 2]: consecutive OUTPUT pixels (stride 2 in the input) read
// consecutive words -- no bank conflicts; tap kx picks plane kx & 1, column x + (kx >> 1)
constexpr int ST_IN_FLOATS = (ST_CIN * 2 * ST_IY * ST_IXH + 3) & ~3;  // keeps the weight / tile arrays 16-byte aligned
constexpr int ST_SMEM_FLOATS = ST_IN_FLOATS + ST_TAPS * ST_COUT + ST_NPIX * ST_CHUNK;

__global__ void __launch_bounds__(ST_THREADS)
stem_conv_pool_kernel(const float* __restrict__ x, int H, int W, long long ldx,          // [B,H,W,ldx>=3] channel-last
                      const float* __restrict__ w,                                       // [64][7][7][3] (OHWI)
                      const float* __restrict__ bias,                                    // [64]
                      int Hc, int Wc, int Hp, int Wp,
                      float* __restrict__ out, long long ldo) {                          // [B,Hp,Wp,ldo>=64]
    extern __shared__ float smem[];
    float* s_in = smem;                                           // [3][2][39][36]
    float* s_w = s_in + ST_IN_FLOATS;                             // [147][64]: tap-major, channels contiguous
    float* s_conv = s_w + ST_TAPS * ST_COUT;                      // [561][16]
    const int t = threadIdx.x, b = blockIdx.z;
    const int py0 = blockIdx.y * ST_TPY, px0 = blockIdx.x * ST_TPX;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;               // first convolution output of the tile
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;               // first input pixel of the halo
    const float* xb = x + (size_t)b * H * W * ldx;

    for (int e = t; e < ST_TAPS * ST_COUT; e += ST_THREADS) {
        const int o = e / ST_TAPS, j = e - o * ST_TAPS;
        s_w[j * ST_COUT + o] = __ldg(w + e);
    }
    for (int e = t; e < ST_IY * ST_IX * ST_CIN; e += ST_THREADS) {
        const int c = e % ST_CIN, lx = (e / ST_CIN) % ST_IX, ly = e / (ST_CIN * ST_IX);
        const int xx = lx + ix0, yy = ly + iy0;
        s_in[((c * 2 + (lx & 1)) * ST_IY + ly) * ST_IXH + (lx >> 1)] =
            (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xb + ((size_t)yy * W + xx) * ldx + c) : 0.f;
    }
    __syncthreads();

    for (int chunk = 0; chunk < ST_COUT / ST_CHUNK; ++chunk) {
        // ---- convolution + bias + ReLU of the tile for 16 channels: thread = outputs p and p + 281 of the 561 (lanes on
        // ---- consecutive outputs; the two share every weight read)
        constexpr int HALF = (ST_NPIX + 1) / 2;                    // 281
        for (int pp = t; pp < HALF; pp += ST_THREADS) {
            const int p0 = pp, p1 = min(pp + HALF, ST_NPIX - 1);
            const int y0 = p0 / ST_CX, x0 = p0 - y0 * ST_CX, y1 = p1 / ST_CX, x1 = p1 - y1 * ST_CX;
            float a0[ST_CHUNK], a1[ST_CHUNK];
#pragma unroll
            for (int o = 0; o < ST_CHUNK; ++o) a0[o] = a1[o] = 0.f;
            for (int ky = 0; ky < ST_K; ++ky) {
                const float* r0 = s_in + (2 * y0 + ky) * ST_IXH + x0;
                const float* r1 = s_in + (2 * y1 + ky) * ST_IXH + x1;
#pragma unroll
                for (int kx = 0; kx < ST_K; ++kx) {
#pragma unroll
                    for (int c = 0; c < ST_CIN; ++c) {
                        const int plane = ((c * 2 + (kx & 1)) * ST_IY) * ST_IXH + (kx >> 1);
                        const float v0 = r0[plane], v1 = r1[plane];
                        const float4* wp = reinterpret_cast<const float4*>(s_w + ((ky * ST_K + kx) * ST_CIN + c) * ST_COUT + chunk * ST_CHUNK);
#pragma unroll
                        for (int o4 = 0; o4 < ST_CHUNK / 4; ++o4) {
                            const float4 q = wp[o4];
                            a0[o4 * 4] = fmaf(v0, q.x, a0[o4 * 4]); a0[o4 * 4 + 1] = fmaf(v0, q.y, a0[o4 * 4 + 1]);
                            a0[o4 * 4 + 2] = fmaf(v0, q.z, a0[o4 * 4 + 2]); a0[o4 * 4 + 3] = fmaf(v0, q.w, a0[o4 * 4 + 3]);
                            a1[o4 * 4] = fmaf(v1, q.x, a1[o4 * 4]); a1[o4 * 4 + 1] = fmaf(v1, q.y, a1[o4 * 4 + 1]);
                            a1[o4 * 4 + 2] = fmaf(v1, q.z, a1[o4 * 4 + 2]); a1[o4 * 4 + 3] = fmaf(v1, q.w, a1[o4 * 4 + 3]);
                        }
                    }
                }
            }
            // outputs outside the convolution grid are the max-pool's padding: 0 is neutral after the ReLU (every
            // pooling window holds at least one real, non-negative output)
            const bool ok0 = (unsigned)(cy0 + y0) < (unsigned)Hc && (unsigned)(cx0 + x0) < (unsigned)Wc;
            const bool two = pp + HALF < ST_NPIX;
            const bool ok1 = two && (unsigned)(cy0 + y1) < (unsigned)Hc && (unsigned)(cx0 + x1) < (unsigned)Wc;
#pragma unroll
            for (int o4 = 0; o4 < ST_CHUNK / 4; ++o4) {
                const float4 bo = __ldg(reinterpret_cast<const float4*>(bias + chunk * ST_CHUNK) + o4);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(s_conv + p0 * ST_CHUNK + o4 * 4) = ok0 ?
                    make_float4(fmaxf(a0[o4 * 4] + bo.x, 0.f), fmaxf(a0[o4 * 4 + 1] + bo.y, 0.f), fmaxf(a0[o4 * 4 + 2] + bo.z, 0.f),
                                fmaxf(a0[o4 * 4 + 3] + bo.w, 0.f)) : z;
                if (two) *reinterpret_cast<float4*>(s_conv + p1 * ST_CHUNK + o4 * 4) = ok1 ?
                    make_float4(fmaxf(a1[o4 * 4] + bo.x, 0.f), fmaxf(a1[o4 * 4 + 1] + bo.y, 0.f), fmaxf(a1[o4 * 4 + 2] + bo.z, 0.f),
                                fmaxf(a1[o4 * 4 + 3] + bo.w, 0.f)) : z;
            }
        }
        __syncthreads();
        // ---- 3x3 stride-2 max-pool of the tile: 8 x 16 pooled pixels x 16 channels, 4 channels per thread and step
        for (int e = t; e < ST_TPY * ST_TPX * (ST_CHUNK / 4); e += ST_THREADS) {
            const int c4 = e % (ST_CHUNK / 4), q = e / (ST_CHUNK / 4);
            const int qx = q % ST_TPX, qy = q / ST_TPX;
            const int py = py0 + qy, px = px0 + qx;
            if (py >= Hp || px >= Wp) continue;
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 v = *reinterpret_cast<const float4*>(s_conv + ((2 * qy + dy) * ST_CX + 2 * qx + dx) * ST_CHUNK + c4 * 4);
                    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                }
            *reinterpret_cast<float4*>(out + (((size_t)b * Hp + py) * Wp + px) * ldo + chunk * ST_CHUNK + c4 * 4) = m;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int camli_stem_conv_pool(const float* x, int B, int H, int W, int64_t ldx, const float* w_ohwi, const float* bias,
                                    float* out, int64_t ldo, void* stream) {
    if (B < 0 || H < 1 || W < 1 || ldx < ST_CIN || ldo < ST_COUT) return CAMLI_EINVAL;
    if (B > 65535 || (ldo & 3)) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!x || !w_ohwi || !bias || !out) return CAMLI_EINVAL;
    if (reinterpret_cast<uintptr_t>(out) & 15) return CAMLI_EINVAL;
    const int Hc = (H - 1) / 2 + 1, Wc = (W - 1) / 2 + 1;          // 7x7, stride 2, padding 3
    const int Hp = (Hc - 1) / 2 + 1, Wp = (Wc - 1) / 2 + 1;        // 3x3, stride 2, padding 1
    const size_t smem = (size_t)ST_SMEM_FLOATS * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(stem_conv_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(camli_div_up(Wp, ST_TPX), camli_div_up(Hp, ST_TPY), B);
    stem_conv_pool_kernel<<<grid, ST_THREADS, smem, (cudaStream_t)stream>>>(x, H, W, ldx, w_ohwi, bias, Hc, Wc, Hp, Wp, out, ldo);
    CAMLI_RETURN_LAUNCH_STATUS();
}
