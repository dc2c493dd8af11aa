// Linear layers and stride-1 "same" convolutions on channel-last activations as ONE implicit GEMM on
// the 5th-generation tensor cores (TMA + tcgen05 + TMEM), fp32-accurate through 3xTF32, with the
// bias / residual / activation of the layer fused into the epilogue.
//
// Replaces, on the hot path, every `MLP1d` / `nn.Linear` / 1x1 `Conv1d|Conv2d` of the point branch
// (reference models/point_conv.py:29,62,106; models/mlp.py:41-128; models/camliraft_l_core.py:46,163)
// and the k x k convolutions of the update block and encoder (models/raft_core.py:110-197): in the
// reference each is a cuBLAS/cuDNN call followed by separate bias, norm and activation kernels.
//
//   out[p, n] = act( sum_{tap} sum_c X[p (+) tap, c] * W[n, tap, c] + bias[n] + residual[p, n] )
//
// p runs over the B*H*W pixels (rows) of a channel-last tensor; a linear layer is the H = 1, 1x1 case.
//
// A operand (activations): a 4-D tensor map (C, W, H, B) with a (32 ch, TW, TH, 1) box, TH*TW = 128:
// each k-block is the box at (c0, x0 + dx, y0 + dy, b) -- the tap shift is just a coordinate offset and
// the zero padding of the convolution is TMA's out-of-bounds fill.  No im2col buffer exists anywhere.
// B operand (weights): [Cout, taps*Cin] row-major (= the OHWI / channels_last weight layout), split once
// on the host side into tf32 hi / lo parts; 2-D map, (32, BN) box at column tap*Cin + c0.
//
// 3xTF32: x = hi + lo with hi = tf32(x), lo = tf32(x - hi); result = hi*hi + (lo*hi + hi*lo), the dropped lo*lo
// being ~2^-22.  Issued as TWO instructions per k-step: A_hi x [W_hi ; W_lo]^T into adjacent [main | corr]
// column ranges (N = 2 BN, the efficient wide shape) and A_lo x W_hi^T into the corr range.  The tensor core TRUNCATES its fp32 accumulator after every instruction (measured: the error
// of a long accumulation is a bias that grows linearly with the number of MMAs, ~0.5 ulp of the running sum
// each), so a plain in-TMEM accumulation over K = 2304 is ~50x less accurate than an fp32 SGEMM.  Two
// measures bring it back to SGEMM level: (1) the correction terms go to their OWN accumulator columns (2^-11
// of the main ones, so their truncations are harmless); (2) accumulation in TMEM runs in chunks of CG_CHUNK
// k-blocks -- each chunk starts from a zeroed accumulator, and the epilogue warps add the chunks in
// registers with round-to-nearest while the next chunk's MMAs run (two accumulator buffers ping-pong).  The activation split happens INSIDE the kernel: four converter
// warps rewrite each landed fp32 tile in place as `hi` and emit `lo` into a sibling buffer (conflict-free
// 16-byte chunks; the 128-byte swizzle is position-preserving), so activations make exactly one trip
// from HBM and no split tensors are ever materialised.
//
// Warp roles (one persistent CTA per SM, 448 threads):
//   warp 0     TMA producer: A (raw fp32), W_hi, W_lo of one k-block per pipeline stage;
//   warp 1     MMA issuer: 12 x tcgen05.mma kind::tf32 (M=128, N=BN, K=8) per k-block (4 main, 8 correction);
//              tcgen05.commit frees the stage / publishes a chunk or the tile's corrections;
//   warps 2-5  converters (hi/lo split of the A tile, generic -> async proxy fence, arrive);
//   warps 6-13 epilogue (two per TMEM lane quarter, half of the columns each): tcgen05.ld of every chunk (register sum),
//              + corrections; the finished 32-row x 32-column block of a warp is transposed through a private shared-
//              memory pad (TMEM hands every thread one ROW, but global memory wants a warp on one row: a thread-per-row
//              store touches 32 lines per instruction), then bias + residual + activation / ConvGRU arithmetic and the
//              store run with the lanes along the columns: 128-byte coalesced side loads and stores.
#include "common.cuh"
#include "tcgen05.cuh"

// Programmatic dependent launch of the kernel (its setup overlaps the previous kernel's tail); camli_conv_gemm_set_pdl.
// Off by default: measured on the C2 graph it gains nothing in latency (13.86 vs 13.78 ms) and costs 4 % throughput
// with three graphs in flight -- the early-resident CTAs (one per SM, ~200 KB of shared memory each) sit on SMs the
// other streams' kernels would have used.
static int camli_cg_pdl = 0;

namespace {

using namespace camli_tc;

constexpr int CG_BM = 128, CG_BK = 32;
constexpr int CG_THREADS = 448;                   // 14 warps: TMA, MMA, 4 converters, 8 epilogue
constexpr int CG_A_BYTES = CG_BM * CG_BK * 4;     // 16 KB
constexpr int CG_CHUNK = 2;                       // k-blocks (of 32 channels) accumulated in TMEM before a drain

template <int BN> struct CgCfg {
    static constexpr int kStageBytes = 2 * CG_A_BYTES + 2 * BN * CG_BK * 4;
    static constexpr int kStages = (BN >= 96) ? 3 : 4;
    static constexpr int kEpiWarps = BN >= 64 ? 8 : 4;
    static constexpr int kPadFloats = 16 * 36;    // one 16-row x 32-column transpose pad per epilogue warp (row pitch 36 floats)
    static constexpr int kSmem = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiWarps * kPadFloats * 4;
    static constexpr int kTmemCols = BN == 96 ? 512 : 4 * BN;      // two [main | correction] accumulator buffers (ping-pong per chunk); power of two
};

struct CgParams {
    int B, H, W, Cin, Cout;        // activation geometry, channels
    int kh, kw;                    // window (odd), padding k/2
    int stride;                    // 1 or 2: H, W above are the OUTPUT grid, the input grid is sampled every `stride` pixels
    int dil;                       // dilation: tap (i, j) reads the input at offset (i - kh/2, j - kw/2) * dil
    int th, tw;                    // pixel tile: th * tw = 128
    int tiles_y, tiles_x, tiles_n;
    const float* bias;             // [Cout] or null
    const float* residual;         // [rows, ldr] or null
    long long ldr;
    float* out;                    // [rows, ldo]
    long long ldo;
    int act;                       // CAMLI_ACT_*
    float slope;
    // fused ConvGRU epilogues (CAMLI_ACT_GRU_*): per-pixel side inputs and an optional second destination
    const float* aux1; long long ld1;
    const float* aux2; long long ld2;
    int split;                     // columns >= split: GATE multiplies by aux1[:, n - split]; with out2 they go to out2
    float* out2; long long ldo2;
    int passes;                    // 3: fp32-accurate 3xTF32 (default); 1: single tf32 product (10-bit mantissa operands, fp32 accumulate)
    long long* timeline;           // diagnostics: SM-clock stamps of CTA 0's pipeline events (null in production)
};

// stamp slot `i` with the SM clock (CTA 0 only, when a timeline buffer is attached)
#define CG_STAMP(i) do { if (P.timeline && blockIdx.x == 0) P.timeline[(i)] = clock64(); } while (0)

template <int ACT>
__device__ __forceinline__ float cg_activate(float v, float slope) {
    if (ACT == CAMLI_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == CAMLI_ACT_LEAKY) return v > 0.f ? v : v * slope;
    if (ACT == CAMLI_ACT_TANH) return tanhf(v);
    if (ACT == CAMLI_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

// Per-element epilogue arithmetic.  Plain layers: activation of v (= accumulator + bias + residual).  ConvGRU
// (reference models/raft_core.py:125-138), same arithmetic as gru_gate_kernel / gru_update_kernel:
//   GATE   : sigmoid(v), times a1 (the hidden state) for the reset-gate columns (gated)
//   UPDATE : (1 - z) * h + z * tanh(v) with z = a1, h = a2; the _FIX variant adds torch.nan_to_num
__device__ __forceinline__ float cg_nan_to_num(float r) {
    if (isnan(r)) return 0.f;
    if (isinf(r)) return r > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
    return r;
}

template <int ACT>
__device__ __forceinline__ float cg_finish(float v, float slope, float a1, float a2, bool gated) {
    if (ACT >= CAMLI_ACT_FIX_NONFINITE) {            // plain activation + torch.nan_to_num (torch's relu keeps a nan)
        if (ACT == (CAMLI_ACT_RELU | CAMLI_ACT_FIX_NONFINITE)) return isnan(v) ? 0.f : fminf(fmaxf(v, 0.f), 3.402823466e+38f);
        return cg_nan_to_num(cg_finish<ACT & (CAMLI_ACT_FIX_NONFINITE - 1)>(v, slope, a1, a2, gated));
    }
    if (ACT == CAMLI_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == CAMLI_ACT_LEAKY) return v > 0.f ? v : v * slope;
    if (ACT == CAMLI_ACT_TANH) return tanhf(v);
    if (ACT == CAMLI_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    if (ACT == CAMLI_ACT_GRU_GATE) {
        const float g = 1.f / (1.f + expf(-v));
        return gated ? g * a1 : g;
    }
    if (ACT == CAMLI_ACT_GRU_UPDATE || ACT == CAMLI_ACT_GRU_UPDATE_FIX) {
        float r = (1.f - a1) * a2 + a1 * tanhf(v);
        if (ACT == CAMLI_ACT_GRU_UPDATE_FIX) {
            if (isnan(r)) r = 0.f;
            else if (isinf(r)) r = r > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
        }
        return r;
    }
    return v;
}

// What one warp needs to finish a 16-row x 32-column block that sits in its pad ([16][36] floats: rows 16-byte
// aligned, conflict-free for the thread-per-row 128-bit writes and the 8-lanes-per-row 128-bit reads).  Read side:
// lane = (row within a group of 4) * 8 + (float4 of the 32 columns).
struct CgBlock {
    uint32_t pad;                     // shared-window address of the warp's pad
    int mypix;                        // pixel index of this LANE's row (lane = row of the 32-row block), -1 outside the image
    int row0;                         // first row (0 or 16) of the half that sits in the pad
    int n_ok;                         // valid columns among this lane's four (0..4)
    bool vec, gated;                  // vec: every pointer / pitch allows 128-bit accesses
    float4 bias;
    float slope;
    float* dst; long long ldd;        // already offset to this lane's first column
    const float* res; long long ldr;
    const float* a1; long long ld1;
    const float* a2; long long ld2;
};

__device__ __forceinline__ float4 cg_ld4(const float* p, bool vec, int n_ok) {
    if (vec) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_ok > 0) v.x = __ldg(p);
    if (n_ok > 1) v.y = __ldg(p + 1);
    if (n_ok > 2) v.z = __ldg(p + 2);
    if (n_ok > 3) v.w = __ldg(p + 3);
    return v;
}

// One activation's store loop for the 16 x 32 block in the pad.
template <int ACT>
__device__ __forceinline__ void cg_store_block(const CgBlock& k, int lane) {
    const int sub = lane >> 3, l8 = lane & 7;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it0 = 0; it0 < 4; it0 += 2) {                              // two rows per lane in flight (register budget)
        float4 v[2], rs[2], a1[2], a2[2];
        int pix[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rl = (it0 + u) * 4 + sub;                           // row within the 16-row half
            pix[u] = __shfl_sync(0xffffffffu, k.mypix, k.row0 + rl);
            const bool ok = pix[u] >= 0 && k.n_ok > 0;
            const size_t pp = ok ? (size_t)pix[u] : 0;
            v[u] = lds_v4(k.pad + (rl * 36 + l8 * 4) * 4);
            rs[u] = (k.res && ok) ? cg_ld4(k.res + pp * k.ldr, k.vec, k.n_ok) : z4;
            a1[u] = (ACT >= CAMLI_ACT_GRU_GATE && ACT < CAMLI_ACT_FIX_NONFINITE && k.a1 && ok) ? cg_ld4(k.a1 + pp * k.ld1, k.vec, k.n_ok) : z4;
            a2[u] = (ACT > CAMLI_ACT_GRU_GATE && ACT < CAMLI_ACT_FIX_NONFINITE && k.a2 && ok) ? cg_ld4(k.a2 + pp * k.ld2, k.vec, k.n_ok) : z4;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (pix[u] < 0 || k.n_ok <= 0) continue;
            float4 o;
            o.x = cg_finish<ACT>((v[u].x + k.bias.x) + rs[u].x, k.slope, a1[u].x, a2[u].x, k.gated);
            o.y = cg_finish<ACT>((v[u].y + k.bias.y) + rs[u].y, k.slope, a1[u].y, a2[u].y, k.gated);
            o.z = cg_finish<ACT>((v[u].z + k.bias.z) + rs[u].z, k.slope, a1[u].z, a2[u].z, k.gated);
            o.w = cg_finish<ACT>((v[u].w + k.bias.w) + rs[u].w, k.slope, a1[u].w, a2[u].w, k.gated);
            float* d = k.dst + (size_t)pix[u] * k.ldd;
            if (k.vec) {
                *reinterpret_cast<float4*>(d) = o;
            } else {
                d[0] = o.x;
                if (k.n_ok > 1) d[1] = o.y;
                if (k.n_ok > 2) d[2] = o.z;
                if (k.n_ok > 3) d[3] = o.w;
            }
        }
    }
}

// A warp's whole 32-row x CW-column share of a tile: per 32-column group and 16-row half, registers -> pad -> rows.
template <int ACT, int CW>
__device__ __forceinline__ void cg_store_tile(float (&sum)[CW], const CgParams& P, uint32_t pad, int mypix, int n0, int lane,
                                              int ncols) {                      // ncols <= CW: columns this warp really owns
#pragma unroll
    for (int c = 0; c < CW / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= P.Cout || c * 32 >= ncols) break;                           // (warp-uniform)
        const int col = col0 + (lane & 7) * 4;                                  // this lane's four columns (read side)
        // destination and side inputs of this 32-column group (split is a multiple of 32: a group never straddles it)
        const bool hi_part = ACT == CAMLI_ACT_GRU_GATE && col0 >= P.split;
        CgBlock k;
        k.pad = pad; k.mypix = mypix; k.gated = hi_part; k.slope = P.slope;
        k.n_ok = min(4, max(0, P.Cout - col));
        k.dst = (P.out2 && hi_part) ? P.out2 + (col - P.split) : P.out + col;
        k.ldd = (P.out2 && hi_part) ? P.ldo2 : P.ldo;
        k.res = P.residual ? P.residual + col : nullptr; k.ldr = P.ldr;
        k.a1 = nullptr; k.a2 = nullptr; k.ld1 = P.ld1; k.ld2 = P.ld2;
        if (hi_part) k.a1 = P.aux1 + (col - P.split);
        else if (ACT > CAMLI_ACT_GRU_GATE && ACT < CAMLI_ACT_FIX_NONFINITE) { k.a1 = P.aux1 + col; k.a2 = P.aux2 + col; }
        k.vec = k.n_ok == 4 && ((k.ldd | k.ldr | k.ld1 | k.ld2) & 3) == 0 &&
                ((reinterpret_cast<uintptr_t>(k.dst) | reinterpret_cast<uintptr_t>(k.res) | reinterpret_cast<uintptr_t>(k.a1) |
                  reinterpret_cast<uintptr_t>(k.a2)) & 15) == 0;
        k.bias = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.bias) k.bias = cg_ld4(P.bias + col, k.n_ok == 4 && (reinterpret_cast<uintptr_t>(P.bias + col) & 15) == 0, k.n_ok);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            __syncwarp();
            if ((lane >> 4) == h) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    sts_v4(pad + ((lane & 15) * 36 + j) * 4,
                           make_float4(sum[c * 32 + j], sum[c * 32 + j + 1], sum[c * 32 + j + 2], sum[c * 32 + j + 3]));
            }
            __syncwarp();
            k.row0 = h * 16;
            cg_store_block<ACT>(k, lane);
        }
    }
}

// tf32 parts of a landed fp32 value: the tensor core reads only the upper 19 bits of an fp32 operand, so `hi`
// is the value itself as it lies in shared memory (truncation, nothing to write); lo = x - trunc(x) (exact),
// rounded to tf32 by integer arithmetic (round half away from zero, as cvt.rna.tf32 does, at full ALU rate --
// the cvt instruction runs on the slow conversion pipe).
__device__ __forceinline__ float cg_lo_part(float x) {
    const float r = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
}

template <int BN>
__global__ void __launch_bounds__(CG_THREADS, 1)
conv_gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                        const __grid_constant__ CUtensorMap map_wlo, const __grid_constant__ CgParams P) {
    using Cfg = CgCfg<BN>;
    constexpr int STAGES = Cfg::kStages;
    constexpr int W_BYTES = BN * CG_BK * 4;
    // epilogue: two warps per TMEM lane quarter, each owning half of the accumulator columns (a 32-column
    // tile is not worth splitting: only the first group of four warps works then)
    constexpr int EPI_WARPS = BN >= 64 ? 8 : 4;
    // columns per epilogue warp (a 96-column tile splits 64 | 32: the register blocks are 32 columns wide)
    constexpr int CW = BN == 96 ? 64 : (BN >= 64 ? BN / 2 : BN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
    // full[s] (TMA landed), conv[s] (A split done), empty[s] (MMAs retired); afull/aempty[2]: the two
    // [main | correction] accumulator buffers (one hand-off per chunk)
    const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + STAGES), bar_empty = smem_u32(bars + 2 * STAGES);
    const uint32_t bar_afull = smem_u32(bars + 3 * STAGES), bar_aempty = smem_u32(bars + 3 * STAGES + 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
    const uint32_t tiles_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) CG_STAMP(0);
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 4);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_afull + 8 * a, 1); mbar_init(bar_aempty + 8 * a, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) touches no global data
    // and may overlap the tail of the previous kernel of the stream; from here on its results are needed (and our
    // writes may alias its inputs), so wait for it -- then let the NEXT kernel's CTAs start their own setup.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) CG_STAMP(1);

    const int cblocks = (P.Cin + CG_BK - 1) / CG_BK;
    const int kblocks = P.kh * P.kw * cblocks;
    const int tiles_per_img = P.tiles_y * P.tiles_x * P.tiles_n;
    const int total = P.B * tiles_per_img;
    const int pad_y = P.kh / 2, pad_x = P.kw / 2;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int b = tile / tiles_per_img;
                int r = tile - b * tiles_per_img;
                const int nt = r % P.tiles_n; r /= P.tiles_n;
                const int x0 = (r % P.tiles_x) * P.tw, y0 = (r / P.tiles_x) * P.th;
                int tap = 0, cb = 0, dy = -pad_y, dx = -pad_x;           // k-block = (tap, channel block), no divisions
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage;
                    const uint32_t dst = tiles_base + stage * Cfg::kStageBytes;
                    mbar_expect_tx(full, CG_A_BYTES + (P.passes == 1 ? 1 : 2) * W_BYTES);
                    if (tile == blockIdx.x && kb == 0) CG_STAMP(14);
                    tma_load_4d(dst, &map_x, full, cb * CG_BK, x0 * P.stride + dx * P.dil, y0 * P.stride + dy * P.dil, b);
                    if (tile == blockIdx.x && kb == 0) CG_STAMP(15);
                    tma_load_2d(dst + CG_A_BYTES * 2, &map_whi, full, tap * P.Cin + cb * CG_BK, nt * BN);
                    if (P.passes != 1) tma_load_2d(dst + CG_A_BYTES * 2 + W_BYTES, &map_wlo, full, tap * P.Cin + cb * CG_BK, nt * BN);
                    if (tile == blockIdx.x && kb < 16) CG_STAMP(16 + kb);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++cb == cblocks) {
                        cb = 0; ++tap;
                        if (++dx > pad_x) { dx = -pad_x; ++dy; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            // Two instructions per k-step: [main | corr] (+)= A_hi x [W_hi ; W_lo]^T  (N = 2 BN: W_hi and W_lo are
            // adjacent in the stage, one descriptor spans both), then corr += A_lo x W_hi^T (N = BN).
            constexpr uint32_t IDESC_2N = tf32_idesc(CG_BM, 2 * BN), IDESC_N = tf32_idesc(CG_BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int ch = 0;                                                  // chunks handled by this CTA so far
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                for (int kb0 = 0; kb0 < kblocks; kb0 += CG_CHUNK, ++ch) {
                    const int acc = ch & 1;
                    mbar_wait(bar_aempty + 8 * acc, ((ch >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_main = tmem_base + acc * 2 * BN, tmem_corr = tmem_main + BN;
                    const int kb1 = min(kb0 + CG_CHUNK, kblocks);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        // conv[s] completes only after the converter warps observed full[s] (acquire) and arrived
                        // (release), so it also orders the TMA-written weight tiles before this thread: one wait, not two
                        mbar_wait(bar_conv + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t src = tiles_base + stage * Cfg::kStageBytes;
                        const uint64_t a_hi = make_kmajor_sw128_desc(src), a_lo = make_kmajor_sw128_desc(src + CG_A_BYTES);
                        const uint64_t b_hl = make_kmajor_sw128_desc(src + 2 * CG_A_BYTES);     // W_hi rows, then W_lo rows
                        if (P.passes == 1) {                             // single product: A (truncated to tf32 by the tensor core) x W_hi
#pragma unroll
                            for (int k = 0; k < CG_BK / 8; ++k)
                                mma_tf32(tmem_main, a_hi + (uint64_t)(k * 2), b_hl + (uint64_t)(k * 2), IDESC_N, ((kb - kb0) | k) ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int k = 0; k < CG_BK / 8; ++k) {
                                const uint64_t adv = (uint64_t)(k * 2);
                                mma_tf32(tmem_main, a_hi + adv, b_hl + adv, IDESC_2N, ((kb - kb0) | k) ? 1u : 0u);
                                mma_tf32(tmem_corr, a_lo + adv, b_hl + adv, IDESC_N, 1u);
                            }
                        }
                        mma_commit(bar_empty + 8 * stage);
                        if (tile == blockIdx.x && kb < 16) CG_STAMP(48 + kb);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    mma_commit(bar_afull + 8 * acc);                     // this chunk (main + corrections) is complete
                }
                if (tile == blockIdx.x) CG_STAMP(8);
            }
        }
    } else if (warp < 6) {
        // ===================== converters (warps 2..5): lo part of the landed fp32 tile =====================
        const int t = threadIdx.x - 64;                                 // 0..127
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                if (t == 0 && tile == blockIdx.x && kb < 16) CG_STAMP(32 + kb);
                if (P.passes != 1) {                                   // (single-pass mode: nothing to split, just hand the stage on)
                    const uint32_t a_hi = tiles_base + stage * Cfg::kStageBytes + t * 16, a_lo = a_hi + CG_A_BYTES;
                    float4 v[CG_A_BYTES / 16 / 128];                        // 8 chunks of 16 bytes per thread
#pragma unroll
                    for (int i = 0; i < CG_A_BYTES / 16 / 128; ++i) v[i] = lds_v4(a_hi + i * 2048);
#pragma unroll
                    for (int i = 0; i < CG_A_BYTES / 16 / 128; ++i)
                        sts_v4(a_lo + i * 2048, make_float4(cg_lo_part(v[i].x), cg_lo_part(v[i].y), cg_lo_part(v[i].z), cg_lo_part(v[i].w)));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to UMMA
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_conv + 8 * stage);              // one arrival per converter warp
                if (t == 0 && tile == blockIdx.x && kb < 16) CG_STAMP(64 + kb);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 6..13) =====================
        const int q = warp & 3;                                         // TMEM lane quarter of this warp
        const int half = (warp - 6) >> 2;                               // which half of the columns
        if (half * CW >= BN) goto done;                                 // (second group idle for 32-column tiles)
        const int ncols = min(CW, BN - half * CW);                      // (32 for the second group of a 96-column tile)
        const int row = q * 32 + lane;                                  // accumulator row = pixel of the tile
        const int ty = row / P.tw, tx = row - ty * P.tw;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + half * CW;
        const uint32_t pad = tiles_base + STAGES * Cfg::kStageBytes + 256 + (warp - 6) * Cfg::kPadFloats * 4;
        int ch = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int b = tile / tiles_per_img;
            int r = tile - b * tiles_per_img;
            const int nt = r % P.tiles_n; r /= P.tiles_n;
            const int x = (r % P.tiles_x) * P.tw + tx, y = (r / P.tiles_x) * P.th + ty;
            const int n0 = nt * BN + half * CW;
            // ---- sum of the chunks (main + corrections), round-to-nearest, in registers
            float sum[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) sum[j] = 0.f;
            for (int kb0 = 0; kb0 < kblocks; kb0 += CG_CHUNK, ++ch) {
                const int acc = ch & 1;
                mbar_wait(bar_afull + 8 * acc, (ch >> 1) & 1);
                tc_fence_after();
                if (warp == 6 && lane == 0 && tile == blockIdx.x && kb0 < 32) CG_STAMP(80 + kb0 / CG_CHUNK);
#pragma unroll
                for (int c = 0; c < CW / 32; ++c) {
                    if (c * 32 >= ncols) break;                          // (warp-uniform)
                    float v[32];
                    if (P.passes != 1) {
                        tmem_ld32(lane_base + acc * 2 * BN + BN + c * 32, v);       // corrections first (small)
#pragma unroll
                        for (int j = 0; j < 32; ++j) sum[c * 32 + j] += v[j];
                    }
                    tmem_ld32(lane_base + acc * 2 * BN + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c * 32 + j] += v[j];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_aempty + 8 * acc);
            }
            if (warp == 6 && lane == 0 && tile == blockIdx.x) CG_STAMP(9);
            // ---- transpose through the pad (16 rows at a time), then bias / residual / activation / store with 8 lanes
            // ---- on each row: 128-bit coalesced side loads and stores.  One switch per tile: only the taken activation's
            // ---- (inlined, unrolled) code is ever fetched.
            const int mypix = (x < P.W && y < P.H) ? (b * P.H + y) * P.W + x : -1;       // pixel of this lane's row
            switch (P.act) {
                case CAMLI_ACT_RELU | CAMLI_ACT_FIX_NONFINITE:
                    cg_store_tile<CAMLI_ACT_RELU | CAMLI_ACT_FIX_NONFINITE, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                case CAMLI_ACT_NONE | CAMLI_ACT_FIX_NONFINITE:
                    cg_store_tile<CAMLI_ACT_NONE | CAMLI_ACT_FIX_NONFINITE, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                case CAMLI_ACT_RELU: cg_store_tile<CAMLI_ACT_RELU, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                case CAMLI_ACT_LEAKY: cg_store_tile<CAMLI_ACT_LEAKY, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                case CAMLI_ACT_TANH: cg_store_tile<CAMLI_ACT_TANH, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                case CAMLI_ACT_SIGMOID: cg_store_tile<CAMLI_ACT_SIGMOID, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                case CAMLI_ACT_GRU_GATE: cg_store_tile<CAMLI_ACT_GRU_GATE, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                case CAMLI_ACT_GRU_UPDATE: cg_store_tile<CAMLI_ACT_GRU_UPDATE, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                case CAMLI_ACT_GRU_UPDATE_FIX: cg_store_tile<CAMLI_ACT_GRU_UPDATE_FIX, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
                default: cg_store_tile<CAMLI_ACT_NONE, CW>(sum, P, pad, mypix, n0, lane, ncols); break;
            }
            if (warp == 6 && lane == 0 && tile == blockIdx.x) CG_STAMP(10);
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) CG_STAMP(11);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    }
}

__global__ void __launch_bounds__(256)
split_tf32_pair_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float h, l;
    split_tf32(__ldg(x + i), h, l);
    hi[i] = h;
    lo[i] = l;
}

int encode_map(CUtensorMap* map, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, int pixel_stride = 1) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    // traversal stride of the two pixel dimensions (a strided convolution reads every `pixel_stride`-th input pixel:
    // the box then spans pixel_stride * tile pixels and delivers tile of them)
    const cuuint32_t estr[4] = {1, (cuuint32_t)pixel_stride, (cuuint32_t)pixel_stride, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides_bytes,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int BN>
int launch_conv_gemm(const CUtensorMap& mx, const CUtensorMap& mwh, const CUtensorMap& mwl, const CgParams& P, int total,
                     cudaStream_t st) {
    using Cfg = CgCfg<BN>;
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) return (int)e;
    const int n_sms = sm_count();
    const int grid = total < n_sms ? total : n_sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(CG_THREADS);
    cfg.dynamicSmemBytes = Cfg::kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = camli_cg_pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, conv_gemm_tf32x3_kernel<BN>, mx, mwh, mwl, P);
    if (e != cudaSuccess) return (int)e;
    CAMLI_RETURN_LAUNCH_STATUS();
}

}  // namespace

// Diagnostics: attach a device buffer of >= 128 int64 that CTA 0 of every following camli_conv_gemm launch
// stamps with SM-clock values of its pipeline events (scripts/conv_gemm_timeline.py); NULL detaches.
static long long* camli_cg_timeline = nullptr;
extern "C" int camli_conv_gemm_set_timeline(long long* device_buffer) {
    camli_cg_timeline = device_buffer;
    return CAMLI_OK;
}

extern "C" int camli_conv_gemm_set_pdl(int enabled) {
    const int old = camli_cg_pdl;
    camli_cg_pdl = enabled ? 1 : 0;
    return old;
}

extern "C" int camli_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
    if (n < 0) return CAMLI_EINVAL;
    if (n == 0) return CAMLI_OK;
    if (!x || !hi || !lo) return CAMLI_EINVAL;
    split_tf32_pair_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, hi, lo, (size_t)n);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_conv_gemm(const float* x, int B, int H, int W, int Cin, int64_t ldx,
                               const float* w_hi, const float* w_lo, int Cout, int kh, int kw,
                               const float* bias, const float* residual, int64_t ldr,
                               int act, float slope, float* out, int64_t ldo, int tile_n, void* stream) {
    if (act < CAMLI_ACT_NONE || (act & ~CAMLI_ACT_FIX_NONFINITE) > CAMLI_ACT_SIGMOID) return CAMLI_EINVAL;
    return camli_conv_gemm_fused(x, B, H, W, Cin, ldx, w_hi, w_lo, Cout, kh, kw, bias, residual, ldr, act, slope, out, ldo,
                                 nullptr, 0, nullptr, 0, 0, nullptr, 0, tile_n, stream);
}

extern "C" int camli_conv_gemm_fused(const float* x, int B, int H, int W, int Cin, int64_t ldx,
                                     const float* w_hi, const float* w_lo, int Cout, int kh, int kw,
                                     const float* bias, const float* residual, int64_t ldr,
                                     int act, float slope, float* out, int64_t ldo,
                                     const float* aux1, int64_t ld1, const float* aux2, int64_t ld2, int split,
                                     float* out2, int64_t ldo2, int tile_n, void* stream) {
    return camli_conv_gemm_strided(x, B, H, W, Cin, ldx, w_hi, w_lo, Cout, kh, kw, 1, 1, bias, residual, ldr, act, slope, out, ldo,
                                   aux1, ld1, aux2, ld2, split, out2, ldo2, tile_n, stream);
}

extern "C" int camli_conv_gemm_strided(const float* x, int B, int Hin, int Win, int Cin, int64_t ldx,
                                       const float* w_hi, const float* w_lo, int Cout, int kh, int kw, int stride, int dilation,
                                       const float* bias, const float* residual, int64_t ldr,
                                       int act, float slope, float* out, int64_t ldo,
                                       const float* aux1, int64_t ld1, const float* aux2, int64_t ld2, int split,
                                       float* out2, int64_t ldo2, int tile_n, void* stream) {
    if ((stride != 1 && stride != 2) || dilation < 1 || dilation > 64) return CAMLI_EUNSUPPORTED;
    if (B < 0 || Hin < 1 || Win < 1 || Cin < 1 || Cout < 1 || kh < 1 || kw < 1 || ldx < Cin) return CAMLI_EINVAL;
    const int H = (Hin - 1) / stride + 1, W = (Win - 1) / stride + 1;       // output grid (padding k/2, odd k)
    if (residual && ldr < Cout) return CAMLI_EINVAL;
    if (act & CAMLI_ACT_FIX_NONFINITE) {            // only NONE / RELU carry the nan_to_num flag on this kernel
        const int base = act & ~CAMLI_ACT_FIX_NONFINITE;
        if (base != CAMLI_ACT_NONE && base != CAMLI_ACT_RELU) return CAMLI_EUNSUPPORTED;
    } else if (act < CAMLI_ACT_NONE || act > CAMLI_ACT_GRU_UPDATE_FIX) return CAMLI_EINVAL;
    const bool gru_update = !(act & CAMLI_ACT_FIX_NONFINITE) && act > CAMLI_ACT_GRU_GATE;
    if (act == CAMLI_ACT_GRU_GATE && (!aux1 || split < 0 || split > Cout || (split & 31) || ld1 < Cout - split)) return CAMLI_EINVAL;
    if (gru_update && (!aux1 || !aux2 || ld1 < Cout || ld2 < Cout)) return CAMLI_EINVAL;
    if (act != CAMLI_ACT_GRU_GATE) { split = Cout; out2 = nullptr; }
    if (out2 ? (ldo < split || ldo2 < Cout - split) : ldo < Cout) return CAMLI_EINVAL;
    // TMA: 16-byte aligned bases and strides; odd windows only ("same" padding)
    if ((kh & 1) == 0 || (kw & 1) == 0 || kh > 15 || kw > 15 || (Cin & 3) || (ldx & 3)) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!x || !w_hi || !w_lo || !out) return CAMLI_EINVAL;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_hi) | reinterpret_cast<uintptr_t>(w_lo)) & 15)
        return CAMLI_EINVAL;
    if ((long long)B * H * W > 2147483647LL) return CAMLI_EUNSUPPORTED;

    CgParams P;
    P.B = B; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout; P.kh = kh; P.kw = kw; P.stride = stride; P.dil = dilation;
    // pixel tile th x tw = 128 with the least padding waste (a linear layer, H = 1, gets 1 x 128)
    int best_th = 1;
    long long best = -1;
    for (int th = 1; th <= 16; th *= 2) {
        const int tw = CG_BM / th;
        const long long tiles = (long long)camli_div_up(H, th) * camli_div_up(W, tw);
        if (best < 0 || tiles < best) { best = tiles; best_th = th; }
    }
    P.th = best_th; P.tw = CG_BM / best_th;
    P.tiles_y = camli_div_up(H, P.th); P.tiles_x = camli_div_up(W, P.tw);
    // N tile: the widest tile that Cout fills.  Inside this kernel a k-step costs ~280 cycles at BN = 128 and hardly less
    // at BN = 64 / 32 (the issue micro-benchmark, scripts/probes/mma_probe.cu, gives 120 / 61 / 47 / 45 cycles per
    // instruction for N = 256 / 128 / 64 / 32: the floor of a narrow instruction plus the per-k-block costs of the pipeline
    // dominate), so a narrow tile does not shorten a CTA's k-loop much -- it multiplies the CTAs (and the activation splits).  Few, wide CTAs also leave SMs free for the kernels of the other branch's stream: the
    // C_out <= 128 convolutions of the update block occupy 68 SMs instead of 136, a point-branch linear 16 instead of 64.
    P.passes = (tile_n & CAMLI_CONV_SINGLE_PASS) ? 1 : 3;
    int bn = tile_n & 0xff;
    // (96-column tiles for the 96 / 192 / 288-channel layers -- the motion encoder's 3x3 256 -> 192, PWC's pyramid and
    // estimator levels: with 128-wide tiles a quarter of their MMA columns would be padding)
    if (bn == 0) bn = Cout > 64 ? ((Cout % 96 == 0 && Cout % 128 != 0) ? 96 : 128) : (Cout > 32 ? 64 : 32);
    if (bn != 32 && bn != 64 && bn != 96 && bn != 128) return CAMLI_EINVAL;
    P.tiles_n = camli_div_up(Cout, bn);
    P.bias = bias; P.residual = residual; P.ldr = ldr; P.out = out; P.ldo = ldo; P.act = act; P.slope = slope; P.timeline = camli_cg_timeline;
    P.aux1 = aux1; P.ld1 = ld1; P.aux2 = aux2; P.ld2 = ld2; P.split = split; P.out2 = out2; P.ldo2 = ldo2;
    const long long total = (long long)B * P.tiles_y * P.tiles_x * P.tiles_n;
    if (total > 2147483647LL) return CAMLI_EUNSUPPORTED;

    CUtensorMap mx, mwh, mwl;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)ldx * Win * 4, (cuuint64_t)ldx * Win * Hin * 4};
        const cuuint32_t box[4] = {CG_BK, (cuuint32_t)(P.tw * stride), (cuuint32_t)(P.th * stride), 1};
        int rc = encode_map(&mx, x, 4, dims, strides, box, stride);
        if (rc) return rc;
    }
    {
        const cuuint64_t ktot = (cuuint64_t)kh * kw * Cin;
        const cuuint64_t dims[2] = {ktot, (cuuint64_t)Cout};
        const cuuint64_t strides[1] = {ktot * 4};
        const cuuint32_t box[2] = {CG_BK, (cuuint32_t)bn};
        int rc = encode_map(&mwh, w_hi, 2, dims, strides, box);
        if (rc) return rc;
        rc = encode_map(&mwl, w_lo, 2, dims, strides, box);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 32: return launch_conv_gemm<32>(mx, mwh, mwl, P, (int)total, st);
        case 64: return launch_conv_gemm<64>(mx, mwh, mwl, P, (int)total, st);
        case 96: return launch_conv_gemm<96>(mx, mwh, mwl, P, (int)total, st);
        default: return launch_conv_gemm<128>(mx, mwh, mwl, P, (int)total, st);
    }
}
