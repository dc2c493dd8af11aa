// Backward kernels of the point-branch / fusion operators (training path, BASELINE config 5) for sm_100a.
//
// The reference differentiates these stages through torch autograd over materialised [B,C,n,k] intermediates
// (models/utils.py:130-146, models/camliraft_l_core.py:56-98, models/clfm.py:57-75).  Each backward here is ONE
// launch that recomputes the forward's small per-neighbour / per-pixel quantities in registers (the neighbour
// search included: the tables are never stored) and scatters the gradient:
//   * three-NN interpolation: d feat[idx_j] += w_j * g                       (atomic adds; weights are constants of xyz)
//   * point-correlation lookup: d vol[q, idx_j] = dcost_j                    (plain stores: (q, idx_j) is unique)
//                               + the 4 -> 32 -> 32 cost-MLP parameter gradients (per-CTA partial sums, then atomics)
//   * point-correlation pooling: d vol_in[p, idx[q, j]] += g[p, q] / k       (atomic adds)
//   * image-correlation pooling: d V0 += up(g1)/4 + up(g2)/16 + ...          (one pass over level 0)
//   * CLFM interpolation: ScoreNet (3 -> 16 -> C) parameter gradients         (thread = channel, per-CTA partial sums)
// Coordinates carry no gradient in these kernels (CamLiRAFT warps by a detached flow, models/camliraft_core.py:105;
// CLFM detaches both cross-modal inputs, models/clfm.py:34-38); the host side keeps the recompute path for callers
// that do need them (CamLiPWC's back-warp).
#include "knn_search.cuh"

namespace {

// ---------------------------------------------------------------------------------------------- three-NN
__device__ __forceinline__ float bp_tnn_weight(int lane, int k, float px, float py, float pz, float ux, float uy, float uz) {
    float w = 0.f;
    if (lane < k) {
        const float dx = px - ux, dy = py - uy, dz = pz - uz;
        w = 1.0f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-8f);
    }
    return w / camli_warp_sum(w);
}

__global__ void __launch_bounds__(KNN_WARPS * 32)
three_nn_interp_backward_kernel(int n, int m, int k, int F,
                                const float* __restrict__ query, KnnView qv, const float* __restrict__ input, KnnView iv,
                                const float* __restrict__ g, long long g_sb, long long g_sc, long long g_sp,
                                float* __restrict__ gfeat, long long f_sb, long long f_sc, long long f_sp) {
    __shared__ KnnTile tile;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    const bool active = q < n;
    const int b = blockIdx.y;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float* qp = query + b * qv.sb + q * qv.sp;
        ux = __ldg(qp); uy = __ldg(qp + qv.sd); uz = __ldg(qp + 2 * qv.sd);
    }
    const float* in = input + b * iv.sb;
    const KnnPlainPoints<3> pts{in, iv.sp, iv.sd};
    KnnList r;
    knn_cta_search<3, 1>(r, tile, pts, m, k, active, ux, uy, uz);
    if (!active) return;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (lane < k) {
        const float* p = in + r.i0 * iv.sp;
        px = __ldg(p); py = __ldg(p + iv.sd); pz = __ldg(p + 2 * iv.sd);
    }
    const float w = bp_tnn_weight(lane, k, px, py, pz, ux, uy, uz);
    for (int f0 = 0; f0 < F; f0 += 32) {
        const int f = f0 + lane;
        const float gq = f < F ? __ldg(g + b * g_sb + f * g_sc + q * g_sp) : 0.f;
        for (int j = 0; j < k; ++j) {
            const int ij = __shfl_sync(CAMLI_FULL_MASK, r.i0, j);
            const float wj = __shfl_sync(CAMLI_FULL_MASK, w, j);
            if (f < F) atomicAdd(gfeat + b * f_sb + f * f_sc + ij * f_sp, wj * gq);
        }
    }
}

// ---------------------------------------------------------------------------------------------- point-correlation lookup
constexpr int CB_K = 16, CB_H = 32, CB_MAX_LEVELS = 8, CB_QPW = 4;   // queries per warp (amortises the parameter atomics)

struct Corr3dGradLevels {
    const float* xyz2[CB_MAX_LEVELS];
    long long sb[CB_MAX_LEVELS], sp[CB_MAX_LEVELS], sd[CB_MAX_LEVELS];
    const float* vol[CB_MAX_LEVELS];      // [B,n1,n2]
    float* gvol[CB_MAX_LEVELS];           // [B,n1,n2], zero-initialised by the caller
    int n2[CB_MAX_LEVELS];
};

__global__ void __launch_bounds__(KNN_WARPS * 32)
corr3d_lookup_backward_kernel(const __grid_constant__ Corr3dGradLevels lv, int n1, const float* __restrict__ xyz1,
                              const float* __restrict__ W1, const float* __restrict__ b1,     // [32,4],[32]
                              const float* __restrict__ W2, const float* __restrict__ b2,     // [32,32],[32]
                              const float* __restrict__ g, int ld_g,                           // rows [B,n1,ld_g]
                              float* __restrict__ gW1, float* __restrict__ gb1, float* __restrict__ gW2, float* __restrict__ gb2) {
    __shared__ KnnTile tile;
    __shared__ __align__(16) float s_row[KNN_WARPS][CB_H];
    __shared__ float s_red[KNN_WARPS][CB_H + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int level = blockIdx.y, b = blockIdx.z;
    const int n2 = lv.n2[level];
    const float* x2 = lv.xyz2[level] + b * lv.sb[level];
    const long long sp = lv.sp[level], sd = lv.sd[level];
    const KnnPlainPoints<3> pts{x2, sp, sd};

    // parameters: lane = unit.  w2r = row `lane` of W2 (layer-2 unit lane), w2c = column `lane` (layer-1 unit lane)
    const float w10 = __ldg(W1 + lane * 4), w11 = __ldg(W1 + lane * 4 + 1), w12 = __ldg(W1 + lane * 4 + 2),
                w13 = __ldg(W1 + lane * 4 + 3), bb1 = __ldg(b1 + lane), bb2 = __ldg(b2 + lane);
    float w2r[CB_H], w2c[CB_H], aW2[CB_H];
#pragma unroll
    for (int a = 0; a < CB_H; ++a) {
        w2r[a] = __ldg(W2 + lane * CB_H + a);
        w2c[a] = __ldg(W2 + a * CB_H + lane);
        aW2[a] = 0.f;
    }
    float ab2 = 0.f, ab1 = 0.f, aW1[4] = {0.f, 0.f, 0.f, 0.f};

    for (int it = 0; it < CB_QPW; ++it) {
        const int q = (blockIdx.x * CB_QPW + it) * KNN_WARPS + warp;
        const bool active = q < n1;
        float ux = 0.f, uy = 0.f, uz = 0.f;
        if (active) {
            const float* qp = xyz1 + (size_t)b * 3 * n1 + q;
            ux = __ldg(qp); uy = __ldg(qp + n1); uz = __ldg(qp + 2 * n1);
        }
        KnnList r;
        knn_cta_search<3, 1>(r, tile, pts, n2, CB_K, active, ux, uy, uz);      // (all warps of the CTA take part)
        if (!active) continue;
        float in0 = 0.f, in1 = 0.f, in2 = 0.f, in3 = 0.f;
        if (lane < CB_K) {
            const float* p = x2 + r.i0 * sp;
            in0 = __ldg(p) - ux; in1 = __ldg(p + sd) - uy; in2 = __ldg(p + 2 * sd) - uz;
            in3 = __ldg(lv.vol[level] + ((size_t)b * n1 + q) * n2 + r.i0);
        }
        const float go = __ldg(g + ((size_t)b * n1 + q) * ld_g + level * CB_H + lane);    // d out[q, level*32 + lane]
        float dcost = 0.f;                                                                // lane j: d in3 of neighbour j
        for (int j = 0; j < CB_K; ++j) {
            const float a0 = __shfl_sync(CAMLI_FULL_MASK, in0, j), a1 = __shfl_sync(CAMLI_FULL_MASK, in1, j);
            const float a2 = __shfl_sync(CAMLI_FULL_MASK, in2, j), a3 = __shfl_sync(CAMLI_FULL_MASK, in3, j);
            // forward of this neighbour: h1[lane], then pre2[lane]
            const float h1 = fmaxf(fmaf(w13, a3, fmaf(w12, a2, fmaf(w11, a1, fmaf(w10, a0, bb1)))), 0.f);
            __syncwarp();
            s_row[warp][lane] = h1;
            __syncwarp();
            float pre2 = bb2;
#pragma unroll
            for (int a = 0; a < CB_H; a += 4) {
                const float4 h = *reinterpret_cast<const float4*>(&s_row[warp][a]);
                pre2 = fmaf(w2r[a], h.x, pre2); pre2 = fmaf(w2r[a + 1], h.y, pre2);
                pre2 = fmaf(w2r[a + 2], h.z, pre2); pre2 = fmaf(w2r[a + 3], h.w, pre2);
            }
            // backward: layer 2 (lane = its unit)
            const float d2 = pre2 > 0.f ? go : 0.f;
            ab2 += d2;
#pragma unroll
            for (int a = 0; a < CB_H; a += 4) {
                const float4 h = *reinterpret_cast<const float4*>(&s_row[warp][a]);
                aW2[a] = fmaf(d2, h.x, aW2[a]); aW2[a + 1] = fmaf(d2, h.y, aW2[a + 1]);
                aW2[a + 2] = fmaf(d2, h.z, aW2[a + 2]); aW2[a + 3] = fmaf(d2, h.w, aW2[a + 3]);
            }
            __syncwarp();
            s_row[warp][lane] = d2;
            __syncwarp();
            // layer 1 (lane = its unit): d h1 = relu'(h1) * sum_c W2[c][lane] * d2[c]
            float d1 = 0.f;
#pragma unroll
            for (int c = 0; c < CB_H; c += 4) {
                const float4 d = *reinterpret_cast<const float4*>(&s_row[warp][c]);
                d1 = fmaf(w2c[c], d.x, d1); d1 = fmaf(w2c[c + 1], d.y, d1);
                d1 = fmaf(w2c[c + 2], d.z, d1); d1 = fmaf(w2c[c + 3], d.w, d1);
            }
            d1 = h1 > 0.f ? d1 : 0.f;
            ab1 += d1;
            aW1[0] = fmaf(d1, a0, aW1[0]); aW1[1] = fmaf(d1, a1, aW1[1]);
            aW1[2] = fmaf(d1, a2, aW1[2]); aW1[3] = fmaf(d1, a3, aW1[3]);
            const float dc = camli_warp_sum(w13 * d1);                                   // d cost entry of neighbour j
            if (lane == j) dcost = dc;
        }
        if (lane < CB_K) lv.gvol[level][((size_t)b * n1 + q) * n2 + r.i0] = dcost;
    }

    // parameter gradients: sum the 8 warps of the CTA through shared memory, then one atomic per element and CTA
    __syncthreads();
    auto cta_sum_add = [&](float v, float* dst) {      // v: one value per (warp, lane); dst[lane] += sum over warps
        s_red[warp][lane] = v;
        __syncthreads();
        if (warp == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < KNN_WARPS; ++w) t += s_red[w][lane];
            atomicAdd(dst + lane, t);
        }
        __syncthreads();
    };
#pragma unroll
    for (int a = 0; a < CB_H; ++a) {
        // element (row = lane, col = a) of W2: dst index lane * 32 + a -> strided; use a per-a base with stride 32
        s_red[warp][lane] = aW2[a];
        __syncthreads();
        if (warp == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < KNN_WARPS; ++w) t += s_red[w][lane];
            atomicAdd(gW2 + lane * CB_H + a, t);
        }
        __syncthreads();
    }
    cta_sum_add(ab2, gb2);
    cta_sum_add(ab1, gb1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s_red[warp][lane] = aW1[i];
        __syncthreads();
        if (warp == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < KNN_WARPS; ++w) t += s_red[w][lane];
            atomicAdd(gW1 + lane * 4 + i, t);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- point-correlation pooling
constexpr int PB_ROWS = 16, PB_MAX_K = 8;

__global__ void __launch_bounds__(256)
corr3d_pool_backward_kernel(int n1, int n_in, int n_out, int k, const float* __restrict__ g,     // [B,n1,n_out]
                            const int64_t* __restrict__ idx, float* __restrict__ g_in) {          // [B,n1,n_in], zeroed
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = blockIdx.y * PB_ROWS, b = blockIdx.z;
    if (q >= n_out) return;
    int col[PB_MAX_K];
    const int64_t* ip = idx + ((size_t)b * n_out + q) * k;
#pragma unroll
    for (int j = 0; j < PB_MAX_K; ++j) col[j] = j < k ? (int)__ldg(ip + j) : 0;
    const float kf = (float)k;
    const int p1 = min(p0 + PB_ROWS, n1);
    for (int p = p0; p < p1; ++p) {
        const float v = __ldg(g + ((size_t)b * n1 + p) * n_out + q) / kf;
        float* row = g_in + ((size_t)b * n1 + p) * n_in;
#pragma unroll
        for (int j = 0; j < PB_MAX_K; ++j)
            if (j < k) atomicAdd(row + col[j], v);
    }
}

// ---------------------------------------------------------------------------------------------- image-correlation pooling
// g0 [rows, h0, w0] += sum_l up_l(g_l) / 4^l  (avg_pool2d(2, 2) backward chained over the levels; cells a floor-divided
// level does not cover receive nothing).
constexpr int QB_MAX_LEVELS = 8;
struct PoolGradLevels {
    const float* g[QB_MAX_LEVELS];
    int h[QB_MAX_LEVELS], w[QB_MAX_LEVELS];
};

__global__ void __launch_bounds__(256)
corr2d_pool_backward_kernel(float* __restrict__ g0, const __grid_constant__ PoolGradLevels lv, int n_levels, long long rows) {
    const int h0 = lv.h[0], w0 = lv.w[0];
    const long long total = rows * h0 * w0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(e % w0), y = (int)((e / w0) % h0);
        const long long r = e / ((long long)w0 * h0);
        float acc = g0[e];
        float scale = 1.f;
        int yy = y, xx = x;
        for (int l = 1; l < n_levels; ++l) {
            // cell (yy, xx) of level l-1 is pooled into (yy/2, xx/2) of level l only if that cell exists AND the 2x2
            // block is complete at every level on the way (floor division drops a trailing odd row / column)
            if ((yy >> 1) >= lv.h[l] || (xx >> 1) >= lv.w[l]) break;
            yy >>= 1; xx >>= 1;
            scale *= 0.25f;
            acc = fmaf(__ldg(lv.g[l] + (r * lv.h[l] + yy) * lv.w[l] + xx), scale, acc);
        }
        g0[e] = acc;
    }
}

// ---------------------------------------------------------------------------------------------- CLFM interpolation
// Gradients of the ScoreNet parameters (models/clfm.py:65-71): per pixel p with nearest projected point nn,
//   in = [uv[nn] - p, |uv[nn] - p|], h = leaky(W1 in + b1, 0.1) (16), s = sigmoid(W2 h + b2) (C), out = s * feat[nn].
// thread = channel c: keeps W2[c][:], accumulates dW2[c][:], db2[c] and its share of dW1 / db1 over the CTA's pixels.
constexpr int CI_HID = 16, CI_THREADS = 256, CI_PIX = 64;

__global__ void __launch_bounds__(CI_THREADS)
clfm_interp_backward_kernel(int H, int W, int N, int C, const float* __restrict__ uv,            // [B,2,N]
                            const int64_t* __restrict__ nn_idx,                                 // [B,HW]
                            const float* __restrict__ feat, long long ld_feat,                   // rows [B,N,ld]
                            const float* __restrict__ W1, const float* __restrict__ b1,        // [16,3],[16]
                            const float* __restrict__ W2, const float* __restrict__ b2,        // [C,16],[C]
                            const float* __restrict__ g,                                        // rows [B,HW,C]
                            float* __restrict__ gW1, float* __restrict__ gb1, float* __restrict__ gW2, float* __restrict__ gb2) {
    __shared__ float s_w1[CI_HID * 3], s_b1[CI_HID];
    __shared__ float s_in[CI_PIX][3];
    __shared__ int s_nn[CI_PIX];
    __shared__ float s_red[CI_THREADS / 32][CI_HID * 4];
    const int t = threadIdx.x, HW = H * W;
    const int c = blockIdx.y * CI_THREADS + t, b = blockIdx.z;
    const int p0 = blockIdx.x * CI_PIX;
    if (t < CI_HID * 3) s_w1[t] = __ldg(W1 + t);
    if (t < CI_HID) s_b1[t] = __ldg(b1 + t);
    if (t < CI_PIX) {
        const int p = p0 + t;
        int nn = 0;
        float dx = 0.f, dy = 0.f;
        if (p < HW) {
            nn = (int)__ldg(nn_idx + (size_t)b * HW + p);
            dx = __ldg(uv + ((size_t)b * 2 + 0) * N + nn) - (float)(p % W);
            dy = __ldg(uv + ((size_t)b * 2 + 1) * N + nn) - (float)(p / W);
        }
        s_nn[t] = nn;
        s_in[t][0] = dx; s_in[t][1] = dy; s_in[t][2] = sqrtf(dx * dx + dy * dy);       // linalg.norm(offset)
    }
    __syncthreads();
    const bool live = c < C;
    float w2[CI_HID], aW2[CI_HID], aW1[CI_HID * 3], ab1[CI_HID];
    float ab2 = 0.f;
    const float bb2 = live ? __ldg(b2 + c) : 0.f;
#pragma unroll
    for (int a = 0; a < CI_HID; ++a) {
        w2[a] = live ? __ldg(W2 + (size_t)c * CI_HID + a) : 0.f;
        aW2[a] = 0.f; ab1[a] = 0.f;
        aW1[a * 3] = aW1[a * 3 + 1] = aW1[a * 3 + 2] = 0.f;
    }
    const int n_pix = min(CI_PIX, HW - p0);
    for (int i = 0; i < n_pix; ++i) {
        const float x0 = s_in[i][0], x1 = s_in[i][1], x2 = s_in[i][2];
        float h[CI_HID], lk[CI_HID];
        float pre2 = bb2;
#pragma unroll
        for (int a = 0; a < CI_HID; ++a) {
            const float pre1 = fmaf(s_w1[a * 3 + 2], x2, fmaf(s_w1[a * 3 + 1], x1, fmaf(s_w1[a * 3], x0, s_b1[a])));
            lk[a] = pre1 > 0.f ? 1.f : 0.1f;
            h[a] = pre1 * lk[a];
            pre2 = fmaf(w2[a], h[a], pre2);
        }
        float d2 = 0.f;
        if (live) {
            const float s = 1.f / (1.f + expf(-pre2));
            const size_t pix = (size_t)b * HW + p0 + i;
            d2 = __ldg(g + pix * C + c) * __ldg(feat + ((size_t)b * N + s_nn[i]) * ld_feat + c) * s * (1.f - s);
        }
        ab2 += d2;
#pragma unroll
        for (int a = 0; a < CI_HID; ++a) {
            aW2[a] = fmaf(d2, h[a], aW2[a]);
            const float d1 = w2[a] * d2 * lk[a];                   // this channel's share of d pre1[a]
            ab1[a] += d1;
            aW1[a * 3] = fmaf(d1, x0, aW1[a * 3]);
            aW1[a * 3 + 1] = fmaf(d1, x1, aW1[a * 3 + 1]);
            aW1[a * 3 + 2] = fmaf(d1, x2, aW1[a * 3 + 2]);
        }
    }
    if (live) {
#pragma unroll
        for (int a = 0; a < CI_HID; ++a) atomicAdd(gW2 + (size_t)c * CI_HID + a, aW2[a]);
        atomicAdd(gb2 + c, ab2);
    }
    // layer-1 gradients: sum over the channels (threads) of the CTA, then one atomic per element
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int a = 0; a < CI_HID; ++a) {
        const float v0 = camli_warp_sum(aW1[a * 3]), v1 = camli_warp_sum(aW1[a * 3 + 1]);
        const float v2 = camli_warp_sum(aW1[a * 3 + 2]), v3 = camli_warp_sum(ab1[a]);
        if (lane == 0) { s_red[warp][a * 4] = v0; s_red[warp][a * 4 + 1] = v1; s_red[warp][a * 4 + 2] = v2; s_red[warp][a * 4 + 3] = v3; }
    }
    __syncthreads();
    if (t < CI_HID * 4) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < CI_THREADS / 32; ++w) v += s_red[w][t];
        const int a = t >> 2, i = t & 3;
        if (i < 3) atomicAdd(gW1 + a * 3 + i, v);
        else atomicAdd(gb1 + a, v);
    }
}

}  // namespace

extern "C" int camli_three_nn_interpolate_backward(int B, int n, int m, int k, int F,
                                                   const float* query_xyz, int64_t q_sb, int64_t q_sp, int64_t q_sd,
                                                   const float* input_xyz, int64_t i_sb, int64_t i_sp, int64_t i_sd,
                                                   const float* grad_out, int64_t g_sb, int64_t g_sc, int64_t g_sp,
                                                   float* grad_feat, int64_t f_sb, int64_t f_sc, int64_t f_sp, void* stream) {
    if (B < 0 || n < 0 || m < 1 || k < 1 || F < 0) return CAMLI_EINVAL;
    if (k > 32 || k > m || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || n == 0 || F == 0) return CAMLI_OK;
    if (!query_xyz || !input_xyz || !grad_out || !grad_feat) return CAMLI_EINVAL;
    dim3 grid(camli_div_up(n, KNN_WARPS), B);
    three_nn_interp_backward_kernel<<<grid, KNN_WARPS * 32, 0, (cudaStream_t)stream>>>(
        n, m, k, F, query_xyz, KnnView{q_sb, q_sp, q_sd}, input_xyz, KnnView{i_sb, i_sp, i_sd},
        grad_out, g_sb, g_sc, g_sp, grad_feat, f_sb, f_sc, f_sp);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_corr3d_lookup_backward(int B, int n1, int n_levels, const float* xyz1,
                                            const float* const* xyz2_levels_host, const int64_t* xyz2_strides_host,
                                            const int* n2_host, const float* const* volumes_host,
                                            float* const* grad_volumes_host,
                                            const float* W1, const float* b1, const float* W2, const float* b2,
                                            const float* grad_out_rows, int ld_grad,
                                            float* grad_W1, float* grad_b1, float* grad_W2, float* grad_b2, void* stream) {
    if (B < 0 || n1 < 0 || n_levels < 1 || ld_grad < n_levels * CB_H) return CAMLI_EINVAL;
    if (n_levels > CB_MAX_LEVELS || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || n1 == 0) return CAMLI_OK;
    if (!xyz1 || !xyz2_levels_host || !xyz2_strides_host || !n2_host || !volumes_host || !grad_volumes_host || !W1 || !b1 ||
        !W2 || !b2 || !grad_out_rows || !grad_W1 || !grad_b1 || !grad_W2 || !grad_b2) return CAMLI_EINVAL;
    Corr3dGradLevels lv;
    for (int l = 0; l < n_levels; ++l) {
        if (!xyz2_levels_host[l] || !volumes_host[l] || !grad_volumes_host[l]) return CAMLI_EINVAL;
        if (n2_host[l] < CB_K) return CAMLI_EUNSUPPORTED;
        lv.xyz2[l] = xyz2_levels_host[l];
        lv.sb[l] = xyz2_strides_host[3 * l]; lv.sp[l] = xyz2_strides_host[3 * l + 1]; lv.sd[l] = xyz2_strides_host[3 * l + 2];
        lv.vol[l] = volumes_host[l]; lv.gvol[l] = grad_volumes_host[l]; lv.n2[l] = n2_host[l];
    }
    dim3 grid(camli_div_up(n1, KNN_WARPS * CB_QPW), n_levels, B);
    corr3d_lookup_backward_kernel<<<grid, KNN_WARPS * 32, 0, (cudaStream_t)stream>>>(
        lv, n1, xyz1, W1, b1, W2, b2, grad_out_rows, ld_grad, grad_W1, grad_b1, grad_W2, grad_b2);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_corr3d_pool_backward(int B, int n1, int n_in, int n_out, int k, const float* grad_out,
                                          const int64_t* knn_idx, float* grad_in, void* stream) {
    if (B < 0 || n1 < 0 || n_in < 1 || n_out < 0 || k < 1) return CAMLI_EINVAL;
    if (B > 65535 || camli_div_up(n1, PB_ROWS) > 65535 || k > PB_MAX_K) return CAMLI_EUNSUPPORTED;
    if (B == 0 || n1 == 0 || n_out == 0) return CAMLI_OK;
    if (!grad_out || !knn_idx || !grad_in) return CAMLI_EINVAL;
    dim3 grid(camli_div_up(n_out, 256), camli_div_up(n1, PB_ROWS), B);
    corr3d_pool_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n1, n_in, n_out, k, grad_out, knn_idx, grad_in);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_corr2d_pool_backward(float* grad_vol0, const float* const* grad_coarser_host, int n_levels, int64_t rows,
                                          int h0, int w0, void* stream) {
    if (rows < 0 || h0 < 1 || w0 < 1 || n_levels < 1) return CAMLI_EINVAL;
    if (n_levels > QB_MAX_LEVELS) return CAMLI_EUNSUPPORTED;
    if (rows == 0 || n_levels == 1) return CAMLI_OK;
    if (!grad_vol0 || !grad_coarser_host) return CAMLI_EINVAL;
    PoolGradLevels lv;
    lv.h[0] = h0; lv.w[0] = w0; lv.g[0] = nullptr;
    for (int l = 1; l < n_levels; ++l) {
        lv.h[l] = lv.h[l - 1] / 2; lv.w[l] = lv.w[l - 1] / 2;
        if (lv.h[l] < 1 || lv.w[l] < 1) return CAMLI_EUNSUPPORTED;
        if (!grad_coarser_host[l - 1]) return CAMLI_EINVAL;
        lv.g[l] = grad_coarser_host[l - 1];
    }
    const long long total = rows * h0 * w0;
    const long long blocks = (total + 255) / 256;
    const unsigned grid = (unsigned)(blocks < 148LL * 32 ? blocks : 148LL * 32);
    corr2d_pool_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_vol0, lv, n_levels, rows);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_clfm_interp_backward(int B, int H, int W, int N, int C, const float* uv, const int64_t* nn_idx,
                                          const float* feat3d_rows, int64_t ld_feat,
                                          const float* W1, const float* b1, const float* W2, const float* b2,
                                          const float* grad_out_rows,
                                          float* grad_W1, float* grad_b1, float* grad_W2, float* grad_b2, void* stream) {
    if (B < 0 || H < 1 || W < 1 || N < 1 || C < 1 || ld_feat < C) return CAMLI_EINVAL;
    if (B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!uv || !nn_idx || !feat3d_rows || !W1 || !b1 || !W2 || !b2 || !grad_out_rows || !grad_W1 || !grad_b1 || !grad_W2 ||
        !grad_b2) return CAMLI_EINVAL;
    dim3 grid(camli_div_up(H * W, CI_PIX), camli_div_up(C, CI_THREADS), B);
    clfm_interp_backward_kernel<<<grid, CI_THREADS, 0, (cudaStream_t)stream>>>(
        H, W, N, C, uv, nn_idx, feat3d_rows, ld_feat, W1, b1, W2, b2, grad_out_rows, grad_W1, grad_b1, grad_W2, grad_b2);
    CAMLI_RETURN_LAUNCH_STATUS();
}

// =============================================================================================
// PointConv grouping stage, backward (forward: pointconv.cu; reference models/point_conv.py:56-66 under autograd).
//   out[s, o*C + c] = sum_j w_j[o] * rows[idx_j, c],   w_j = WeightNet(rows[idx_j, 0:3] - centre[s])   (3 -> 8 -> 16, leaky)
// Same mapping as the forward: one warp owns a centroid, lane j re-evaluates the WeightNet of neighbour j, the
// lanes sweep the channels of the neighbours' rows.  With G[o][c] = grad_out[s, o*C + c] in registers:
//   d rows[idx_j, c] += sum_o w_j[o] G[o][c]                       (atomics: a point is a neighbour of many centroids)
//   d w_j[o]          = sum_c G[o][c] rows[idx_j, c]               (warp sums; lane j keeps its 16)
// then lane j walks back through its WeightNet: the offset gradient goes to the neighbour's xyz columns (+) and the
// centroid (-), and the outer products that make up the WeightNet's parameter gradients are staged in shared memory,
// summed over the CTA's 8 centroids x k neighbours by one thread per parameter, kept in a register across the CTA's
// grid-stride loop and added to global memory once per CTA.
namespace {

constexpr int PCB_WARPS = 8;
constexpr int PCB_NW = 16, PCB_H = 8;
constexpr int PCB_REC = PCB_NW + PCB_H + PCB_H + 3;                       // d pre2 | h | d pre1 | offset  = 35 floats
constexpr int PCB_PARAMS = PCB_H * 3 + PCB_H + PCB_NW * PCB_H + PCB_NW;   // W1 | b1 | W2 | b2 = 176

template <int CHUNKS>
__global__ void __launch_bounds__(PCB_WARPS * 32)
pointconv_group_backward_kernel(int N, int S, int K, int k, int C,
                                const float* __restrict__ rows, long long ldr,
                                const float* __restrict__ centre, long long c_sb, long long c_sp, long long c_sd,
                                const int64_t* __restrict__ idx,
                                const float* __restrict__ W1, const float* __restrict__ b1,
                                const float* __restrict__ W2, const float* __restrict__ b2, float slope,
                                const float* __restrict__ g_out,                                   // [B,S,16*C]
                                float* __restrict__ g_rows,                                        // [B,N,C] zero-initialised
                                float* __restrict__ g_centre,                                      // [B,3,S] contiguous, every element written
                                float* __restrict__ g_params) {                                    // [176] zero-initialised
    __shared__ float s_rec[PCB_WARPS][32][PCB_REC + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const float* rb = rows + (size_t)b * N * ldr;
    float* grb = g_rows + (size_t)b * N * C;
    float param_acc = 0.f;                                                  // thread t < 176 owns parameter t
    const int groups = (S + PCB_WARPS - 1) / PCB_WARPS;
    for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const int s = grp * PCB_WARPS + warp;
        const bool live = s < S;
        float rec[PCB_REC];
#pragma unroll
        for (int i = 0; i < PCB_REC; ++i) rec[i] = 0.f;
        if (live) {
            const float* cp = centre + b * c_sb + s * c_sp;
            const float cx = __ldg(cp), cy = __ldg(cp + c_sd), cz = __ldg(cp + 2 * c_sd);
            int my = 0;
            float h[PCB_H], w[PCB_NW], off0 = 0.f, off1 = 0.f, off2 = 0.f;
#pragma unroll
            for (int a = 0; a < PCB_H; ++a) h[a] = 0.f;
#pragma unroll
            for (int o = 0; o < PCB_NW; ++o) w[o] = 0.f;
            if (lane < k) {
                my = (int)__ldg(idx + ((size_t)b * S + s) * K + lane);
                const float* p = rb + (size_t)my * ldr;
                off0 = __ldg(p) - cx; off1 = __ldg(p + 1) - cy; off2 = __ldg(p + 2) - cz;
#pragma unroll
                for (int a = 0; a < PCB_H; ++a)
                    h[a] = camli_leaky(fmaf(__ldg(W1 + a * 3 + 2), off2, fmaf(__ldg(W1 + a * 3 + 1), off1,
                                       fmaf(__ldg(W1 + a * 3), off0, __ldg(b1 + a)))), slope);
#pragma unroll
                for (int o = 0; o < PCB_NW; ++o) {
                    float acc = __ldg(b2 + o);
#pragma unroll
                    for (int a = 0; a < PCB_H; ++a) acc = fmaf(__ldg(W2 + o * PCB_H + a), h[a], acc);
                    w[o] = camli_leaky(acc, slope);
                }
            }
            // this centroid's output gradient: G[t][o] = g_out[s, o*C + t*32 + lane]
            float G[CHUNKS][PCB_NW];
            const float* gs = g_out + ((size_t)b * S + s) * PCB_NW * C;
#pragma unroll
            for (int o = 0; o < PCB_NW; ++o)
#pragma unroll
                for (int t = 0; t < CHUNKS; ++t) {
                    const int c = t * 32 + lane;
                    G[t][o] = (c < C) ? __ldg(gs + (size_t)o * C + c) : 0.f;
                }
            float dw[PCB_NW];
#pragma unroll
            for (int o = 0; o < PCB_NW; ++o) dw[o] = 0.f;
            for (int j = 0; j < k; ++j) {
                const int ij = __shfl_sync(CAMLI_FULL_MASK, my, j);
                const float* g = rb + (size_t)ij * ldr;
                float gv[CHUNKS], dr[CHUNKS];
#pragma unroll
                for (int t = 0; t < CHUNKS; ++t) {
                    const int c = t * 32 + lane;
                    gv[t] = (c < C) ? __ldg(g + c) : 0.f;
                    dr[t] = 0.f;
                }
#pragma unroll
                for (int o = 0; o < PCB_NW; ++o) {
                    const float wj = __shfl_sync(CAMLI_FULL_MASK, w[o], j);
                    float part = 0.f;
#pragma unroll
                    for (int t = 0; t < CHUNKS; ++t) {
                        dr[t] = fmaf(wj, G[t][o], dr[t]);
                        part = fmaf(G[t][o], gv[t], part);
                    }
                    part = camli_warp_sum(part);
                    if (lane == j) dw[o] = part;
                }
#pragma unroll
                for (int t = 0; t < CHUNKS; ++t) {
                    const int c = t * 32 + lane;
                    if (c < C) atomicAdd(grb + (size_t)ij * C + c, dr[t]);
                }
            }
            // lane j: back through its WeightNet (leaky': the sign of the activation is the sign of the pre-activation)
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
            if (lane < k) {
                float dh[PCB_H];
#pragma unroll
                for (int a = 0; a < PCB_H; ++a) dh[a] = 0.f;
#pragma unroll
                for (int o = 0; o < PCB_NW; ++o) {
                    const float dp2 = dw[o] * (w[o] > 0.f ? 1.f : slope);
                    rec[o] = dp2;
#pragma unroll
                    for (int a = 0; a < PCB_H; ++a) dh[a] = fmaf(__ldg(W2 + o * PCB_H + a), dp2, dh[a]);
                }
#pragma unroll
                for (int a = 0; a < PCB_H; ++a) {
                    const float dp1 = dh[a] * (h[a] > 0.f ? 1.f : slope);
                    rec[PCB_NW + a] = h[a];
                    rec[PCB_NW + PCB_H + a] = dp1;
                    d0 = fmaf(__ldg(W1 + a * 3), dp1, d0);
                    d1 = fmaf(__ldg(W1 + a * 3 + 1), dp1, d1);
                    d2 = fmaf(__ldg(W1 + a * 3 + 2), dp1, d2);
                }
                rec[PCB_NW + 2 * PCB_H] = off0; rec[PCB_NW + 2 * PCB_H + 1] = off1; rec[PCB_NW + 2 * PCB_H + 2] = off2;
                float* gp = grb + (size_t)my * C;
                atomicAdd(gp, d0); atomicAdd(gp + 1, d1); atomicAdd(gp + 2, d2);
            }
            const float c0 = camli_warp_sum(d0), c1 = camli_warp_sum(d1), c2 = camli_warp_sum(d2);
            if (lane == 0) {
                float* gc = g_centre + (size_t)b * 3 * S + s;
                gc[0] = -c0; gc[S] = -c1; gc[2 * (size_t)S] = -c2;
            }
        }
        // ---- parameter gradients of this group of centroids: one thread per parameter sums the staged records
#pragma unroll
        for (int i = 0; i < PCB_REC; ++i) s_rec[warp][lane][i] = rec[i];     // (lanes >= k and dead warps stage zeros)
        __syncthreads();
        const int t = threadIdx.x;
        if (t < PCB_PARAMS) {
            int ia, ib;                                                         // product of record fields ia * ib (ib < 0: ia alone)
            if (t < PCB_H * 3) { ia = PCB_NW + PCB_H + t / 3; ib = PCB_NW + 2 * PCB_H + t % 3; }                 // dW1[a][d] = d pre1[a] * off[d]
            else if (t < PCB_H * 4) { ia = PCB_NW + PCB_H + (t - PCB_H * 3); ib = -1; }                         // db1[a]
            else if (t < PCB_H * 4 + PCB_NW * PCB_H) { const int u = t - PCB_H * 4; ia = u / PCB_H; ib = PCB_NW + u % PCB_H; }   // dW2[o][a] = d pre2[o] * h[a]
            else { ia = t - (PCB_H * 4 + PCB_NW * PCB_H); ib = -1; }                                             // db2[o]
            float acc = 0.f;
            for (int wq = 0; wq < PCB_WARPS; ++wq)
                for (int l = 0; l < k; ++l)
                    acc += ib >= 0 ? s_rec[wq][l][ia] * s_rec[wq][l][ib] : s_rec[wq][l][ia];
            param_acc += acc;
        }
        __syncthreads();
    }
    if (threadIdx.x < PCB_PARAMS) atomicAdd(g_params + threadIdx.x, param_acc);
}

}  // namespace

extern "C" int camli_pointconv_group_backward(int B, int N, int S, int K, int k, int C,
                                              const float* rows, int64_t ld_rows,
                                              const float* centre_xyz, int64_t c_sb, int64_t c_sp, int64_t c_sd,
                                              const int64_t* knn_idx, const float* W1, const float* b1, const float* W2,
                                              const float* b2, float negative_slope, const float* grad_out,
                                              float* grad_rows, float* grad_centre, float* grad_params, void* stream) {
    if (B < 0 || N < 1 || S < 0 || k < 1 || K < k || C < 3 || ld_rows < C) return CAMLI_EINVAL;
    if (k > 32 || B > 65535 || C > 32 * 8) return CAMLI_EUNSUPPORTED;
    if (B == 0 || S == 0) return CAMLI_OK;
    if (!rows || !centre_xyz || !knn_idx || !W1 || !b1 || !W2 || !b2 || !grad_out || !grad_rows || !grad_centre || !grad_params)
        return CAMLI_EINVAL;
    const int groups = camli_div_up(S, PCB_WARPS);
    dim3 grid(groups < 148 * 2 ? groups : 148 * 2, B), block(PCB_WARPS * 32);
    cudaStream_t st = (cudaStream_t)stream;
#define CAMLI_PCB_LAUNCH(CH)                                                                                                 \
    pointconv_group_backward_kernel<CH><<<grid, block, 0, st>>>(N, S, K, k, C, rows, ld_rows, centre_xyz, c_sb, c_sp, c_sd, \
                                                                knn_idx, W1, b1, W2, b2, negative_slope, grad_out, grad_rows, \
                                                                grad_centre, grad_params)
    const int chunks = camli_div_up(C, 32);
    if (chunks <= 2) CAMLI_PCB_LAUNCH(2);
    else if (chunks <= 4) CAMLI_PCB_LAUNCH(4);
    else if (chunks <= 6) CAMLI_PCB_LAUNCH(6);
    else CAMLI_PCB_LAUNCH(8);
#undef CAMLI_PCB_LAUNCH
    CAMLI_RETURN_LAUNCH_STATUS();
}
