// Weight gradient of the implicit-GEMM convolution / linear layer (camli_conv_gemm) on the 5th-generation tensor
// cores: the training-side counterpart of conv_gemm.cu.  In the reference every dense layer's backward is a cuDNN
// wgrad / cuBLAS GEMM chosen by autograd (models/raft_core.py:110-197, models/mlp.py:41-128, train.py:143-171).
//
//   dW[n, tap, c] = sum_p  G[p, n] * X[p (+) tap, c]          G = dL/d(pre-activation) [B*H*W, Cout], X [B*H*W, Cin]
//
// As a GEMM the reduction runs over the PIXELS, the dimension both tensors are strided in.  A pre-pass
// (camli_transpose_split) therefore writes each operand once more as [C, B, H, W] -- pixel-contiguous, i.e. K-major for
// the tensor core -- already split into tf32 hi / lo parts (3xTF32: hi*hi + hi*lo + lo*hi, fp32-level accuracy); for G it
// also applies the activation derivative and emits the row-major gradient the data-gradient convolution reads.
//
// Kernel (persistent, one CTA per SM, TMA + tcgen05 + TMEM, same skeleton as allpairs_gemm.cu):
//   * tile = (tap, 128 output channels, 128 input channels, K split); a k-block is 32 consecutive pixels of one image row,
//     fetched by 5-D tensor maps (W, H, B, C, shift) with a (32, 1, 1, 128, 1) box.  The vertical tap offset is a coordinate
//     offset of the X box (rows outside the image are skipped: they contribute nothing); the horizontal one cannot be -- a TMA
//     box must start 16-byte aligned in the contiguous dimension, and a one-pixel shift is 4 bytes -- so the pre-pass writes
//     X^T once per horizontal tap offset, already shifted (zero outside the row); the ragged row end (W % 32) is TMA's
//     out-of-bounds fill;
//   * the pixel range is split over `splits` CTAs per tile so that every SM has work (a 128x128 tile of a 1x1 layer would
//     otherwise be ONE CTA walking all pixels); partial tiles are added to the zeroed dW with 128-bit `red.global.add`.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace {

using namespace camli_tc;

constexpr int WG_BM = 128, WG_BN = 128, WG_BK = 32;
constexpr int WG_STAGES = 3;
constexpr int WG_TILE_BYTES = WG_BM * WG_BK * 4;             // 16 KB
constexpr int WG_STAGE_BYTES = 4 * WG_TILE_BYTES;            // G_hi, G_lo, X_hi, X_lo
constexpr int WG_THREADS = 192;
constexpr int WG_TMEM_COLS = 2 * WG_BN;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;
constexpr uint32_t WG_IDESC = tf32_idesc(WG_BM, WG_BN), WG_IDESC_BF16 = bf16_idesc(WG_BM, WG_BN);

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

struct WgParams {
    int B, H, W, Cout, Cin, kh, kw, dil;   // H, W: the OUTPUT grid (rows of G); the input grid has Hin rows
    int Hin, stride;             // X row of output row y and vertical tap offset dy: y * stride + dy (must lie in [0, Hin))
    int wblocks;                 // k-blocks per image row = ceil(W / 32)
    int kblocks;                 // B * H * wblocks
    int splits, kb_per_split;
    int tiles_m, tiles_n;
    int passes;                  // 3: 3xTF32 (fp32-accurate), 1: hi x hi only (tf32 operands)
    int bf16;                    // operands are bf16 (one product, 64 pixels per 128-byte k-block row instead of 32)
    int bk_px;                   // pixels per k-block: 32 (fp32 operands) or 64 (bf16)
    float* dw;                   // [Cout, kh*kw*Cin], zeroed by the caller
    long long ldw;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                         const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                         const __grid_constant__ WgParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 4);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + WG_STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * WG_STAGES), bar_tempty = smem_u32(bars + 2 * WG_STAGES + 2);
    const uint32_t tiles_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)WG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int taps = P.kh * P.kw;
    const int total = taps * P.tiles_m * P.tiles_n * P.splits;

    // tile -> (tap, m tile, n tile, split); k-block range of the split
#define WG_DECODE(tile)                                                                        \
    int r_ = (tile);                                                                           \
    const int sp = r_ % P.splits; r_ /= P.splits;                                              \
    const int nt = r_ % P.tiles_n; r_ /= P.tiles_n;                                            \
    const int mt = r_ % P.tiles_m;                                                             \
    const int tap = r_ / P.tiles_m;                                                            \
    const int kb_lo = sp * P.kb_per_split, kb_hi = min(P.kblocks, kb_lo + P.kb_per_split);

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                WG_DECODE(tile)
                const int dy = (tap / P.kw - P.kh / 2) * P.dil, sx = tap % P.kw;      // sx: which pre-shifted copy of X^T
                // first k-block of the split -> (image, row, row block), then counted up
                int xb = kb_lo % P.wblocks, ry = kb_lo / P.wblocks;
                int y = ry % P.H, b = ry / P.H;
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    if ((unsigned)(y * P.stride + dy) < (unsigned)P.Hin) {            // (rows above / below the image: zero padding)
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t full = bar_full + 8 * stage;
                        const uint32_t dst = tiles_base + stage * WG_STAGE_BYTES;
                        mbar_expect_tx(full, P.passes == 1 ? 2 * WG_TILE_BYTES : WG_STAGE_BYTES);
                        tma_load_5d(dst + 0 * WG_TILE_BYTES, &map_g_hi, full, xb * P.bk_px, y, b, mt * WG_BM, 0);
                        tma_load_5d(dst + 2 * WG_TILE_BYTES, &map_x_hi, full, xb * P.bk_px, y * P.stride + dy, b, nt * WG_BN, sx);
                        if (P.passes != 1) {
                            tma_load_5d(dst + 1 * WG_TILE_BYTES, &map_g_lo, full, xb * P.bk_px, y, b, mt * WG_BM, 0);
                            tma_load_5d(dst + 3 * WG_TILE_BYTES, &map_x_lo, full, xb * P.bk_px, y * P.stride + dy, b, nt * WG_BN, sx);
                        }
                        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (++xb == P.wblocks) { xb = 0; if (++y == P.H) { y = 0; ++b; } }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
                WG_DECODE(tile)
                (void)mt; (void)nt;
                const int dy = (tap / P.kw - P.kh / 2) * P.dil;
                const int acc = it & 1;
                mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * WG_BN;
                int xb = kb_lo % P.wblocks, y = (kb_lo / P.wblocks) % P.H;
                uint32_t started = 0u;
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    if ((unsigned)(y * P.stride + dy) < (unsigned)P.Hin) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t src = tiles_base + stage * WG_STAGE_BYTES;
                        const uint64_t a_hi = make_kmajor_sw128_desc(src), a_lo = make_kmajor_sw128_desc(src + WG_TILE_BYTES);
                        const uint64_t b_hi = make_kmajor_sw128_desc(src + 2 * WG_TILE_BYTES),
                                       b_lo = make_kmajor_sw128_desc(src + 3 * WG_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < WG_BK / 8; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            if (P.bf16) mma_f16(tmem_d, a_hi + adv, b_hi + adv, WG_IDESC_BF16, started | (uint32_t)k);
                            else mma_tf32(tmem_d, a_hi + adv, b_hi + adv, WG_IDESC, started | (uint32_t)k);
                            if (P.passes != 1) {
                                mma_tf32(tmem_d, a_hi + adv, b_lo + adv, WG_IDESC, 1u);
                                mma_tf32(tmem_d, a_lo + adv, b_hi + adv, WG_IDESC, 1u);
                            }
                        }
                        started = 1u;
                        mma_commit(bar_empty + 8 * stage);
                        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (++xb == P.wblocks) { xb = 0; if (++y == P.H) y = 0; }
                }
                mma_commit(bar_tfull + 8 * acc);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5): partial tile += into dW =====================
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            WG_DECODE(tile)
            const int acc = it & 1;
            mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const int row = mt * WG_BM + q * 32 + lane;                               // output channel
            float* __restrict__ drow = P.dw + (size_t)row * P.ldw + (size_t)tap * P.Cin + nt * WG_BN;
            const bool vec_ok = (((size_t)row * P.ldw + (size_t)tap * P.Cin + nt * WG_BN) & 3) == 0;
            // nothing accumulated: a split past the end, or every row of its pixel range falls into the vertical padding
            bool empty = true;
            {
                const int dy = (tap / P.kw - P.kh / 2) * P.dil;
                int y = kb_lo < kb_hi ? (kb_lo / P.wblocks) % P.H : 0;
                for (int kb = kb_lo; kb < kb_hi && empty; kb += P.wblocks) {          // one probe per image row touched
                    if ((unsigned)(y * P.stride + dy) < (unsigned)P.Hin) empty = false;
                    if (++y == P.H) y = 0;
                }
                if (empty && kb_lo < kb_hi) {                                        // the last (partial) row of the range
                    const int yl = ((kb_hi - 1) / P.wblocks) % P.H;
                    if ((unsigned)(yl * P.stride + dy) < (unsigned)P.Hin) empty = false;
                }
            }
#pragma unroll 1
            for (int c = 0; c < WG_BN / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + acc * WG_BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
                if (row < P.Cout && !empty) {
                    const int col0 = nt * WG_BN + c * 32;
                    if (vec_ok && col0 + 32 <= P.Cin) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            atomicAdd(reinterpret_cast<float4*>(drow + c * 32 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < P.Cin) atomicAdd(drow + c * 32 + j, v[j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
    }
#undef WG_DECODE
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)WG_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Pre-pass: rows [P, C] (pitch ld) -> [C, P] tf32 hi / lo (32 x 32 shared-memory transpose).  With `y` the rows are a
// gradient: it is first multiplied by the derivative of the layer's activation at its OUTPUT y (relu / leaky: sign of y;
// tanh: 1 - y^2; sigmoid: y (1 - y)) and -- with `g_rows` -- also written back row-major (the operand of the
// data-gradient convolution); `colsum` [C] accumulates the bias gradient (atomics, zeroed by the caller).
// Tile: 32 pixels x 128 channels per CTA (256 threads).  Load: a thread reads one float4 (4 channels) of 4 pixels -- a warp
// covers whole 512-byte rows; store: a warp writes 32 consecutive pixels (128 bytes) of one channel, 16 channels per thread.
// (bf16 output: 64 pixels per tile, a lane writes two pixels as one bf16x2 -- 128-byte rows again.)
constexpr int TS_CH = 128;

template <int TS_PX>
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ x, long long ld, int P, int C, const float* __restrict__ y, long long ldy,
                       int act, float slope, float* __restrict__ hi_t, float* __restrict__ lo_t, long long ldt,
                       float* __restrict__ g_rows, float* __restrict__ colsum, int W, int n_shift, int shift_step,
                       int W_out, int xstride) {                                     // P counts OUTPUT pixels: rows of W_out
    __shared__ float s_t[TS_PX][TS_CH + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = blockIdx.x * TS_PX, c0 = blockIdx.y * TS_CH;
    const int cq = (tid & 31) * 4, pr = tid >> 5;                                    // load side: 4 channels, pixel rows pr + 8 i
    const bool vec = ((ld | ldy) & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                     (C & 3) == 0;
    for (int j = 0; j < n_shift; ++j) {
        const int shift = (j - n_shift / 2) * shift_step;                            // copy j holds the rows shifted by `shift` pixels
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < TS_PX / 8; ++i) {
            const int p = p0 + pr + 8 * i, c = c0 + cq;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (p < P && c < C) {
                const int row = p / W_out, xo = p - row * W_out;
                const int xs = xo * xstride + shift;                                 // stays inside its image row, or reads zero
                const bool inside = (unsigned)xs < (unsigned)W;
                const float* src = x + ((long long)row * W + xs) * ld + c;
                float o[4] = {1.f, 1.f, 1.f, 1.f};
                if (vec) {
                    if (inside) { const float4 t = __ldg(reinterpret_cast<const float4*>(src)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
                    if (y) { const float4 t = __ldg(reinterpret_cast<const float4*>(y + (long long)p * ldy + c)); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (c + e < C) {
                            if (inside) v[e] = __ldg(src + e);
                            if (y) o[e] = __ldg(y + (long long)p * ldy + c + e);
                        }
                    }
                }
                if (y) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (act == CAMLI_ACT_RELU) v[e] = o[e] > 0.f ? v[e] : 0.f;
                        else if (act == CAMLI_ACT_LEAKY) v[e] = o[e] > 0.f ? v[e] : v[e] * slope;
                        else if (act == CAMLI_ACT_TANH) v[e] = v[e] * (1.f - o[e] * o[e]);
                        else if (act == CAMLI_ACT_SIGMOID) v[e] = v[e] * o[e] * (1.f - o[e]);
                    }
                    if (g_rows) {
                        float* d = g_rows + (long long)p * C + c;
                        if (vec) *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
                        else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) if (c + e < C) d[e] = v[e];
                        }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) { part[e] += v[e]; s_t[pr + 8 * i][cq + e] = v[e]; }
        }
        __syncthreads();
        if (colsum) {                                                                // bias gradient: column sums of this tile
            for (int c = tid; c < TS_CH; c += 256) {
                if (c0 + c < C) {
                    float t = 0.f;
#pragma unroll 8
                    for (int r = 0; r < TS_PX; ++r) t += s_t[r][c];
                    atomicAdd(colsum + c0 + c, t);
                }
            }
        }
        (void)part;
        float* __restrict__ hj = hi_t + (long long)j * C * ldt;
        float* __restrict__ lj = lo_t ? lo_t + (long long)j * C * ldt : nullptr;
        if (TS_PX == 64) {
            // bf16 operands (hi_t is a bf16 tensor of the same shape; ldt and p0 are even): lane = pixel pair
            const int p = p0 + 2 * lane;
            __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(hi_t) + (long long)j * C * ldt;
#pragma unroll 4
            for (int i = 0; i < TS_CH / 8; ++i) {
                const int c = c0 + warp + 8 * i;
                if (c < C && p < P) {
                    const float a = s_t[(2 * lane) % TS_PX][warp + 8 * i], b2 = s_t[(2 * lane + 1) % TS_PX][warp + 8 * i];
                    if (p + 1 < P) *reinterpret_cast<__nv_bfloat162*>(hb + (long long)c * ldt + p) = __floats2bfloat162_rn(a, b2);
                    else hb[(long long)c * ldt + p] = __float2bfloat16_rn(a);
                }
            }
        } else {
            const int p = p0 + lane;
#pragma unroll 4
            for (int i = 0; i < TS_CH / 8; ++i) {
                const int c = c0 + warp + 8 * i;
                if (c < C && p < P) {
                    float h, l;
                    split_tf32(s_t[lane][warp + 8 * i], h, l);
                    hj[(long long)c * ldt + p] = h;
                    if (lj) lj[(long long)c * ldt + p] = l;
                }
            }
        }
        __syncthreads();
    }
}

int wg_encode_map(CUtensorMap* map, const void* base, int B, int H, int W, int C, int S, bool bf16) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)C, (cuuint64_t)S};
    const cuuint64_t es = bf16 ? 2 : 4;
    const cuuint64_t strides[4] = {(cuuint64_t)W * es, (cuuint64_t)W * H * es, (cuuint64_t)W * H * B * es, (cuuint64_t)W * H * B * C * es};
    const cuuint32_t box[5] = {(cuuint32_t)(bf16 ? 2 * WG_BK : WG_BK), 1, 1, WG_BM, 1};          // one 128-byte row of pixels
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace

extern "C" int camli_transpose_split(const float* rows, int64_t ld, int64_t P, int C, const float* y_rows, int64_t ldy,
                                     int act, float slope, int W, int n_shift, int shift_step, int xstride,
                                     float* hi_t, float* lo_t, float* g_rows, float* colsum, void* stream) {
    const int out_bf16 = (xstride & CAMLI_TRANSPOSE_BF16) ? 1 : 0;                   // flag bit: hi_t is bf16, lo_t unused
    xstride &= ~CAMLI_TRANSPOSE_BF16;
    if (P < 0 || C < 1 || ld < C || (y_rows && ldy < C) || W < 1 || n_shift < 1 || (n_shift & 1) == 0 || shift_step < 1 ||
        (xstride != 1 && xstride != 2))
        return CAMLI_EINVAL;
    if ((n_shift > 1 || xstride > 1) && (y_rows || g_rows || colsum || P % W)) return CAMLI_EINVAL;   // plain activations only
    const int W_out = (W - 1) / xstride + 1;
    const int64_t P_in = P;
    P = P / W * W_out + (xstride == 1 ? P % W : 0);                                 // output pixels (x subsampled by xstride)
    (void)P_in;
    if (act < CAMLI_ACT_NONE || act > CAMLI_ACT_SIGMOID) return CAMLI_EUNSUPPORTED;
    if (P == 0) return CAMLI_OK;
    if (!rows || !hi_t) return CAMLI_EINVAL;                                          // lo_t == NULL: hi parts only
    if (P > 2147483647LL - 64) return CAMLI_EUNSUPPORTED;
    if (out_bf16) {
        if (P & 1) return CAMLI_EUNSUPPORTED;                                        // bf16x2 stores: even row pitch
        const dim3 grid((unsigned)camli_div_up_ll(P, 64), (unsigned)camli_div_up(C, TS_CH));
        transpose_split_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(rows, ld, (int)P, C, y_rows, ldy, act, slope, hi_t, lo_t, P,
                                                                           g_rows, colsum, W, n_shift, shift_step, W_out, xstride);
        CAMLI_RETURN_LAUNCH_STATUS();
    }
    const dim3 grid((unsigned)camli_div_up_ll(P, 32), (unsigned)camli_div_up(C, TS_CH));
    transpose_split_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(rows, ld, (int)P, C, y_rows, ldy, act, slope, hi_t, lo_t, P,
                                                                       g_rows, colsum, W, n_shift, shift_step, W_out, xstride);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_conv_wgrad(const float* g_hi_t, const float* g_lo_t, const float* x_hi_t, const float* x_lo_t,
                                int B, int H, int W, int Cout, int Cin, int kh, int kw, int dilation, int stride, int Hin,
                                int passes, float* dw, void* stream) {
    const bool accumulate = (passes & CAMLI_WGRAD_ACCUMULATE) != 0;                 // dw += ... (the caller's gradient buffer)
    passes &= ~CAMLI_WGRAD_ACCUMULATE;
    const bool bf16 = passes == CAMLI_WGRAD_BF16;
    if (bf16) passes = 1;
    if (passes != 1 && passes != 3) return CAMLI_EINVAL;
    if ((stride != 1 && stride != 2) || Hin < 1 || (Hin - 1) / stride + 1 != H) return CAMLI_EINVAL;
    if (B < 0 || H < 1 || W < 1 || Cout < 1 || Cin < 1 || kh < 1 || kw < 1 || dilation < 1) return CAMLI_EINVAL;
    if ((kh & 1) == 0 || (kw & 1) == 0 || kh > 15 || kw > 15 || (W & (bf16 ? 7 : 3)) || (Cin & 3) || dilation > 64) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!g_hi_t || !x_hi_t || !dw) return CAMLI_EINVAL;
    if (passes == 1) { g_lo_t = g_hi_t; x_lo_t = x_hi_t; }                          // (never read)
    if (!g_lo_t || !x_lo_t) return CAMLI_EINVAL;
    if ((reinterpret_cast<uintptr_t>(g_hi_t) | reinterpret_cast<uintptr_t>(g_lo_t) | reinterpret_cast<uintptr_t>(x_hi_t) |
         reinterpret_cast<uintptr_t>(x_lo_t) | reinterpret_cast<uintptr_t>(dw)) & 15)
        return CAMLI_EINVAL;
    WgParams P;
    P.B = B; P.H = H; P.W = W; P.Cout = Cout; P.Cin = Cin; P.kh = kh; P.kw = kw; P.dil = dilation;
    P.Hin = Hin; P.stride = stride;
    P.bf16 = bf16 ? 1 : 0; P.bk_px = bf16 ? 2 * WG_BK : WG_BK;
    P.wblocks = camli_div_up(W, P.bk_px);
    const long long kblocks = (long long)B * H * P.wblocks;
    if (kblocks > 2147483647LL) return CAMLI_EUNSUPPORTED;
    P.kblocks = (int)kblocks;
    P.tiles_m = camli_div_up(Cout, WG_BM); P.tiles_n = camli_div_up(Cin, WG_BN);
    const int n_sms = sm_count();
    const long long base_tiles = (long long)kh * kw * P.tiles_m * P.tiles_n;
    // K splits: enough CTAs for every SM (two rounds when the pixel range is long), never fewer than 8 k-blocks per split
    long long splits = (2LL * n_sms + base_tiles - 1) / base_tiles;
    if (splits > kblocks / 8) splits = kblocks / 8;
    if (splits < 1) splits = 1;
    P.splits = (int)splits;
    P.kb_per_split = (int)((kblocks + splits - 1) / splits);
    P.passes = passes;
    P.dw = dw; P.ldw = (long long)kh * kw * Cin;
    const long long total = base_tiles * splits;
    if (total > 2147483647LL) return CAMLI_EUNSUPPORTED;

    CUtensorMap m_ghi, m_glo, m_xhi, m_xlo;
    int rc;
    if ((rc = wg_encode_map(&m_ghi, g_hi_t, B, H, W, Cout, 1, bf16))) return rc;
    if ((rc = wg_encode_map(&m_glo, g_lo_t, B, H, W, Cout, 1, bf16))) return rc;
    if ((rc = wg_encode_map(&m_xhi, x_hi_t, B, Hin, W, Cin, kw, bf16))) return rc;  // kw horizontally pre-shifted (and, for stride 2,
    if ((rc = wg_encode_map(&m_xlo, x_lo_t, B, Hin, W, Cin, kw, bf16))) return rc;  // horizontally subsampled) copies, all Hin rows
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    if (!accumulate) {
        e = cudaMemsetAsync(dw, 0, (size_t)Cout * P.ldw * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    const int grid = (int)(total < n_sms ? total : n_sms);
    conv_wgrad_tf32x3_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, st>>>(m_ghi, m_glo, m_xhi, m_xlo, P);
    CAMLI_RETURN_LAUNCH_STATUS();
}
