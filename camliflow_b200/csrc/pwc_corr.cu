// PWC local cost volume (forward + backward) for sm_100a.
//
// Replaces correlation_forward_kernel / correlation_backward_input{1,2}_kernel
// (reference models/csrc/correlation/correlation_forward_kernel.cu:11-55,
//  correlation_backward_kernel.cu:4-89).
//
// The op is HBM-bound (2*C*4 bytes in + 81*4 bytes out per pixel against
// 2*81*C flops).  The reference launches one 32-thread CTA per pixel and does 81
// serial warp reductions, re-reading every input2 pixel 81 times from L1/L2.
// Here a CTA owns one row tile of 32 output pixels: the 9-row x 40-pixel halo of
// input2 is staged ONCE per 32-channel chunk in shared memory (transposed to
// x-fastest so a thread's 12-pixel sliding window is three LDS.128), and each
// thread keeps a 4-pixel x 9-displacement register block, so every staged value
// feeds 9 FMAs from registers.  Out-of-range displacements fall out as exact
// zeros because the halo is zero-filled (the reference relies on a pre-zeroed
// output tensor instead; ours writes every output element).
//
// Backward: grad1[p,c] = (1/C) sum_disp gO[disp][p] * in2[p+disp][c]; grad2 has the
// same form after the substitution disp' = -disp with the "flipped, shifted"
// view G'[disp'][p] = gO[-disp'][p+disp'] and in1 in place of in2, so ONE kernel
// (template FLIP) serves both gradients.
#include "common.cuh"

namespace {

constexpr int CORR_D = 4;                       // max_displacement of the tiled kernels
constexpr int CORR_DS = 2 * CORR_D + 1;         // 9
constexpr int CORR_NDISP = CORR_DS * CORR_DS;   // 81
constexpr int CORR_TX = 32;                     // output pixels per CTA (one row)
constexpr int CORR_HALO = CORR_TX + 2 * CORR_D; // 40
constexpr int CORR_CC = 32;                     // channels per staged chunk

// ---------------------------------------------------------------- forward ----
constexpr int FWD_THREADS = CORR_DS * 32;       // warp w <-> displacement row dy = w-4
constexpr int FWD_S2 = 44;                      // x-stride of one (row, channel) plane of in2 (16B aligned)
constexpr int FWD_S1 = 36;                      // x-stride of one channel plane of in1
constexpr int FWD_SMEM_FLOATS = CORR_DS * CORR_CC * FWD_S2 + CORR_CC * FWD_S1;

__global__ void __launch_bounds__(FWD_THREADS)
corr_fwd_tiled_kernel(float* __restrict__ out, const float* __restrict__ in1,
                      const float* __restrict__ in2, int C, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    float* s_in2 = smem;                                   // [9][CC][S2]
    float* s_in1 = smem + CORR_DS * CORR_CC * FWD_S2;      // [CC][S1]

    const int tid = threadIdx.x, lane = tid & 31, dyi = tid >> 5;
    const int xg = lane & 7, cs = lane >> 3;
    const int x0 = blockIdx.x * CORR_TX, y = blockIdx.y, n = blockIdx.z;

    float acc[4][CORR_DS];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int d = 0; d < CORR_DS; ++d) acc[p][d] = 0.f;

    for (int c0 = 0; c0 < C; c0 += CORR_CC) {
        __syncthreads();
        // stage in2 halo: 9 rows x 40 px x 8 float4 channel groups
        for (int it = tid; it < CORR_DS * CORR_HALO * (CORR_CC / 4); it += FWD_THREADS) {
            const int cg = it & 7, pxr = it >> 3;
            const int r = pxr / CORR_HALO, xl = pxr - r * CORR_HALO;
            const int yy = y + r - CORR_D, xx = x0 + xl - CORR_D;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                v = __ldg(reinterpret_cast<const float4*>(in2 + (((size_t)n * H + yy) * W + xx) * C + c0 + 4 * cg));
            float* dst = s_in2 + (r * CORR_CC + 4 * cg) * FWD_S2 + xl;
            dst[0] = v.x; dst[FWD_S2] = v.y; dst[2 * FWD_S2] = v.z; dst[3 * FWD_S2] = v.w;
        }
        for (int it = tid; it < CORR_TX * (CORR_CC / 4); it += FWD_THREADS) {
            const int cg = it & 7, xl = it >> 3;
            const int xx = x0 + xl;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (xx < W)
                v = __ldg(reinterpret_cast<const float4*>(in1 + (((size_t)n * H + y) * W + xx) * C + c0 + 4 * cg));
            float* dst = s_in1 + (4 * cg) * FWD_S1 + xl;
            dst[0] = v.x; dst[FWD_S1] = v.y; dst[2 * FWD_S1] = v.z; dst[3 * FWD_S1] = v.w;
        }
        __syncthreads();
#pragma unroll 2
        for (int ci = 0; ci < CORR_CC / 4; ++ci) {
            const int c = cs + 4 * ci;
            const float4 a4 = *reinterpret_cast<const float4*>(s_in1 + c * FWD_S1 + 4 * xg);
            const float* brow = s_in2 + (dyi * CORR_CC + c) * FWD_S2 + 4 * xg;
            const float4 b0 = *reinterpret_cast<const float4*>(brow);
            const float4 b1 = *reinterpret_cast<const float4*>(brow + 4);
            const float4 b2 = *reinterpret_cast<const float4*>(brow + 8);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int d = 0; d < CORR_DS; ++d) acc[p][d] = fmaf(a[p], b[p + d], acc[p][d]);
        }
    }
    // sum the 4 channel sub-slices (lanes differing in bits 3,4), then lane (xg, cs)
    // writes pixel 4*xg + cs so a warp stores 32 consecutive floats per displacement.
    const int x = x0 + 4 * xg + cs;
    const float inv_den = (float)C;
#pragma unroll
    for (int d = 0; d < CORR_DS; ++d) {
        float v[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float s = acc[p][d];
            s += __shfl_xor_sync(CAMLI_FULL_MASK, s, 8);
            s += __shfl_xor_sync(CAMLI_FULL_MASK, s, 16);
            v[p] = s;
        }
        const float mine = cs == 0 ? v[0] : cs == 1 ? v[1] : cs == 2 ? v[2] : v[3];
        if (x < W)
            out[(((size_t)n * CORR_NDISP + dyi * CORR_DS + d) * H + y) * W + x] = mine / inv_den;
    }
}

// Any C / any displacement: one thread per output element (slow path).
__global__ void corr_fwd_generic_kernel(float* __restrict__ out, const float* __restrict__ in1,
                                        const float* __restrict__ in2, int B, int C, int H, int W, int md) {
    const int ds = 2 * md + 1;
    const long long total = (long long)B * ds * ds * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int tc = (int)((i / ((long long)W * H)) % (ds * ds));
        const int n = (int)(i / ((long long)W * H * ds * ds));
        const int y2 = y + tc / ds - md, x2 = x + tc % ds - md;
        float s = 0.f;
        if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
            const float* a = in1 + (((size_t)n * H + y) * W + x) * C;
            const float* b = in2 + (((size_t)n * H + y2) * W + x2) * C;
            for (int c = 0; c < C; ++c) s = fmaf(__ldg(a + c), __ldg(b + c), s);
        }
        out[i] = s / (float)C;
    }
}

// --------------------------------------------------------------- backward ----
constexpr int BWD_THREADS = 256;               // lane <-> channel, warp <-> 4-pixel group
constexpr int BWD_SG = 36;                     // x-stride of one displacement plane of the gO tile
constexpr int BWD_SO = 33;                     // transpose buffer stride
constexpr int BWD_SMEM_FLOATS = CORR_NDISP * BWD_SG + CORR_DS * CORR_HALO * CORR_CC + CORR_CC * BWD_SO;

template <bool FLIP>
__global__ void __launch_bounds__(BWD_THREADS)
corr_bwd_tiled_kernel(float* __restrict__ grad_in, const float* __restrict__ grad_out,
                      const float* __restrict__ other, int C, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    float* s_g = smem;                                             // [81][SG]
    float* s_in = s_g + CORR_NDISP * BWD_SG;                       // [9][40][CC]
    float* s_out = s_in + CORR_DS * CORR_HALO * CORR_CC;           // [CC][SO]

    const int tid = threadIdx.x, lane = tid & 31, pg = tid >> 5;
    const int x0 = blockIdx.x * CORR_TX, y = blockIdx.y, n = blockIdx.z;

    // gO tile (FLIP: G'[tc'][x] = gO[80-tc'][y+dy'][x+dx'])
    for (int it = tid; it < CORR_NDISP * CORR_TX; it += BWD_THREADS) {
        const int xl = it & 31, tc = it >> 5;
        int src_tc = tc, yy = y, xx = x0 + xl;
        if (FLIP) {
            src_tc = CORR_NDISP - 1 - tc;
            yy = y + tc / CORR_DS - CORR_D;
            xx = x0 + xl + tc % CORR_DS - CORR_D;
        }
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W && x0 + xl < W)
            v = __ldg(grad_out + (((size_t)n * CORR_NDISP + src_tc) * H + yy) * W + xx);
        s_g[tc * BWD_SG + xl] = v;
    }

    for (int c0 = 0; c0 < C; c0 += CORR_CC) {
        __syncthreads();   // previous chunk's s_in / s_out readers are done; s_g visible
        for (int it = tid; it < CORR_DS * CORR_HALO * (CORR_CC / 4); it += BWD_THREADS) {
            const int cg = it & 7, pxr = it >> 3;
            const int r = pxr / CORR_HALO, xl = pxr - r * CORR_HALO;
            const int yy = y + r - CORR_D, xx = x0 + xl - CORR_D;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                v = __ldg(reinterpret_cast<const float4*>(other + (((size_t)n * H + yy) * W + xx) * C + c0 + 4 * cg));
            *reinterpret_cast<float4*>(s_in + (size_t)pxr * CORR_CC + 4 * cg) = v;
        }
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int dyi = 0; dyi < CORR_DS; ++dyi) {
            float b[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) b[j] = s_in[(dyi * CORR_HALO + 4 * pg + j) * CORR_CC + lane];
#pragma unroll
            for (int d = 0; d < CORR_DS; ++d) {
                const float4 g = *reinterpret_cast<const float4*>(s_g + (dyi * CORR_DS + d) * BWD_SG + 4 * pg);
                acc[0] = fmaf(g.x, b[0 + d], acc[0]);
                acc[1] = fmaf(g.y, b[1 + d], acc[1]);
                acc[2] = fmaf(g.z, b[2 + d], acc[2]);
                acc[3] = fmaf(g.w, b[3 + d], acc[3]);
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) s_out[lane * BWD_SO + 4 * pg + p] = acc[p];
        __syncthreads();
        for (int it = tid; it < CORR_CC * CORR_TX; it += BWD_THREADS) {
            const int xl = it & 31, c = it >> 5;
            if (x0 + xl < W)
                grad_in[(((size_t)n * C + c0 + c) * H + y) * W + x0 + xl] = s_out[c * BWD_SO + xl] / (float)C;
        }
    }
}

template <bool FLIP>
__global__ void corr_bwd_generic_kernel(float* __restrict__ grad_in, const float* __restrict__ grad_out,
                                        const float* __restrict__ other, int B, int C, int H, int W, int md) {
    const int ds = 2 * md + 1;
    const long long total = (long long)B * C * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int c = (int)((i / ((long long)W * H)) % C);
        const int n = (int)(i / ((long long)W * H * C));
        float s = 0.f;
        for (int ty = -md; ty <= md; ++ty)
            for (int tx = -md; tx <= md; ++tx) {
                // FLIP=false: p2 = p1 + disp, gO taken at p1 ; FLIP=true: p1 = p2 - disp, gO taken at p1
                const int yo = FLIP ? y - ty : y + ty, xo = FLIP ? x - tx : x + tx;
                if (yo < 0 || yo >= H || xo < 0 || xo >= W) continue;
                const int tc = (ty + md) * ds + (tx + md);
                const int yg = FLIP ? yo : y, xg = FLIP ? xo : x;
                s = fmaf(__ldg(grad_out + (((size_t)n * ds * ds + tc) * H + yg) * W + xg),
                         __ldg(other + (((size_t)n * H + yo) * W + xo) * C + c), s);
            }
        grad_in[i] = s / (float)C;
    }
}

bool corr_tiled_ok(int C, int md, const void* a, const void* b) {
    return md == CORR_D && C % CORR_CC == 0 &&
           (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0;
}

}  // namespace

extern "C" int camli_correlation_forward(float* out, const float* in1, const float* in2, int B, int C,
                                         int H, int W, int md, void* stream) {
    if (B < 0 || C < 1 || H < 0 || W < 0 || md < 0) return CAMLI_EINVAL;
    if (B == 0 || H == 0 || W == 0) return CAMLI_OK;
    if (!out || !in1 || !in2) return CAMLI_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (corr_tiled_ok(C, md, in1, in2) && H <= 65535 && B <= 65535) {
        const size_t smem = FWD_SMEM_FLOATS * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(corr_fwd_tiled_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid(camli_div_up(W, CORR_TX), H, B);
        corr_fwd_tiled_kernel<<<grid, FWD_THREADS, smem, st>>>(out, in1, in2, C, H, W);
    } else {
        const long long total = (long long)B * (2 * md + 1) * (2 * md + 1) * H * W;
        const int blocks = (int)(camli_div_up_ll(total, 256) < 148 * 16 ? camli_div_up_ll(total, 256) : 148 * 16);
        corr_fwd_generic_kernel<<<blocks, 256, 0, st>>>(out, in1, in2, B, C, H, W, md);
    }
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_correlation_backward(const float* grad_out, float* grad_in1, float* grad_in2,
                                          const float* in1, const float* in2, int B, int C, int H, int W,
                                          int md, void* stream) {
    if (B < 0 || C < 1 || H < 0 || W < 0 || md < 0) return CAMLI_EINVAL;
    if (B == 0 || H == 0 || W == 0) return CAMLI_OK;
    if (!grad_out || !grad_in1 || !grad_in2 || !in1 || !in2) return CAMLI_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (corr_tiled_ok(C, md, in1, in2) && H <= 65535 && B <= 65535) {
        const size_t smem = BWD_SMEM_FLOATS * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(corr_bwd_tiled_kernel<false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(corr_bwd_tiled_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid(camli_div_up(W, CORR_TX), H, B);
        corr_bwd_tiled_kernel<false><<<grid, BWD_THREADS, smem, st>>>(grad_in1, grad_out, in2, C, H, W);
        corr_bwd_tiled_kernel<true><<<grid, BWD_THREADS, smem, st>>>(grad_in2, grad_out, in1, C, H, W);
    } else {
        const long long total = (long long)B * C * H * W;
        const int blocks = (int)(camli_div_up_ll(total, 256) < 148 * 16 ? camli_div_up_ll(total, 256) : 148 * 16);
        corr_bwd_generic_kernel<false><<<blocks, 256, 0, st>>>(grad_in1, grad_out, in2, B, C, H, W, md);
        corr_bwd_generic_kernel<true><<<blocks, 256, 0, st>>>(grad_in2, grad_out, in1, B, C, H, W, md);
    }
    CAMLI_RETURN_LAUNCH_STATUS();
}
