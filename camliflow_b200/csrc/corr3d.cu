// Point all-pairs correlation for sm_100a: pyramid pooling and the fused per-iteration lookup.
//
// Replaces Correlation3D.build_cost_volume_pyramid's pooling loop and Correlation3D.forward /
// calc_matching_cost (reference models/camliraft_l_core.py:51-101).  Per GRU iteration the
// reference runs, for each of the 4 levels: a k-NN launch (k=16), a channel-first gather of the
// neighbour coordinates, an advanced-indexing gather of the volume entries, a cat, two 1x1
// Conv2d + ReLU over [B,4->32->32,n1,k] and a sum over k -- ~40 launches and several
// [B,32,n1,16] temporaries.  Here ONE launch covers all levels: a warp searches its query's 16
// neighbours in the (warped) level cloud (bit-exact order, knn_search.cuh), gathers the 16
// offsets and volume entries, runs the 4->32->32 MLP with the hidden layer in shared memory and
// writes the k-summed 32 channels of that level straight into the [B,n1,32*L] row.
#include "knn_search.cuh"

namespace {

constexpr int C3_WARPS = KNN_WARPS;
constexpr int C3_K = 16;
constexpr int C3_H = 32;
constexpr int C3_MAX_LEVELS = 8;

struct Corr3dLevels {
    const float* xyz2[C3_MAX_LEVELS];     // [B,3,n2] views
    long long sb[C3_MAX_LEVELS], sp[C3_MAX_LEVELS], sd[C3_MAX_LEVELS];
    const float* vol[C3_MAX_LEVELS];      // [B,n1,n2] contiguous
    int n2[C3_MAX_LEVELS];
};

__global__ void __launch_bounds__(C3_WARPS * 32)
corr3d_lookup_kernel(const __grid_constant__ Corr3dLevels lv, int n1, const float* __restrict__ xyz1,   // [B,3,n1]
                     const float* __restrict__ W1, const float* __restrict__ b1,                        // [32,4],[32]
                     const float* __restrict__ W2, const float* __restrict__ b2,                        // [32,32],[32]
                     float* __restrict__ out, int ld_out) {                                             // rows [B,n1,ld]
    __shared__ KnnTile tile;
    __shared__ __align__(16) float s_h1[C3_WARPS][C3_K][C3_H];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * C3_WARPS + warp;
    const bool active = q < n1;
    const int level = blockIdx.y, b = blockIdx.z;
    const int n2 = lv.n2[level];
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float* qp = xyz1 + (size_t)b * 3 * n1 + q;
        ux = __ldg(qp); uy = __ldg(qp + n1); uz = __ldg(qp + 2 * n1);
    }
    const float* x2 = lv.xyz2[level] + b * lv.sb[level];
    const long long sp = lv.sp[level], sd = lv.sd[level];
    const KnnPlainPoints<3> pts{x2, sp, sd};
    KnnList r;
    knn_cta_search<3, 1>(r, tile, pts, n2, C3_K, active, ux, uy, uz);
    if (!active) return;

    // lanes 0..15: neighbour offset + matching cost
    float in0 = 0.f, in1 = 0.f, in2 = 0.f, in3 = 0.f;
    if (lane < C3_K) {
        const float* p = x2 + r.i0 * sp;
        in0 = __ldg(p) - ux; in1 = __ldg(p + sd) - uy; in2 = __ldg(p + 2 * sd) - uz;
        in3 = __ldg(lv.vol[level] + ((size_t)b * n1 + q) * n2 + r.i0);
    }
    // layer 1 (4 -> 32, ReLU): lane = hidden unit
    const float w10 = __ldg(W1 + lane * 4), w11 = __ldg(W1 + lane * 4 + 1), w12 = __ldg(W1 + lane * 4 + 2),
                w13 = __ldg(W1 + lane * 4 + 3), bb1 = __ldg(b1 + lane);
#pragma unroll
    for (int j = 0; j < C3_K; ++j) {
        const float a0 = __shfl_sync(CAMLI_FULL_MASK, in0, j), a1 = __shfl_sync(CAMLI_FULL_MASK, in1, j);
        const float a2 = __shfl_sync(CAMLI_FULL_MASK, in2, j), a3 = __shfl_sync(CAMLI_FULL_MASK, in3, j);
        s_h1[warp][j][lane] = fmaxf(fmaf(w13, a3, fmaf(w12, a2, fmaf(w11, a1, fmaf(w10, a0, bb1)))), 0.f);
    }
    __syncwarp();
    // layer 2 (32 -> 32, ReLU) and the sum over the k neighbours: lane = output unit
    float w2[C3_H];
#pragma unroll
    for (int a = 0; a < C3_H; a += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(W2 + lane * C3_H + a));
        w2[a] = v.x; w2[a + 1] = v.y; w2[a + 2] = v.z; w2[a + 3] = v.w;
    }
    const float bb2 = __ldg(b2 + lane);
    float total = 0.f;
#pragma unroll 4
    for (int j = 0; j < C3_K; ++j) {
        float acc = bb2;
#pragma unroll
        for (int a = 0; a < C3_H; a += 4) {
            const float4 h = *reinterpret_cast<const float4*>(&s_h1[warp][j][a]);   // broadcast
            acc = fmaf(w2[a], h.x, acc); acc = fmaf(w2[a + 1], h.y, acc);
            acc = fmaf(w2[a + 2], h.z, acc); acc = fmaf(w2[a + 3], h.w, acc);
        }
        total += fmaxf(acc, 0.f);
    }
    out[((size_t)b * n1 + q) * ld_out + level * C3_H + lane] = total;
}

// vol_out[b,p,q] = mean_j vol_in[b,p,idx[b,q,j]]   (camliraft_l_core.py:56-60).  A thread owns one output column q:
// its k neighbour columns are read once (the int64 table is the largest stream of the naive form, re-read for
// every row) and kept in registers while the CTA walks C3P_ROWS rows of the volume.
constexpr int C3P_ROWS = 16;
constexpr int C3P_MAX_K = 8;

__global__ void __launch_bounds__(256)
corr3d_pool_kernel(int n1, int n_in, int n_out, int k, const float* __restrict__ vol_in,
                   const int64_t* __restrict__ idx, float* __restrict__ vol_out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = blockIdx.y * C3P_ROWS, b = blockIdx.z;
    if (q >= n_out) return;
    int col[C3P_MAX_K];
    const int64_t* ip = idx + ((size_t)b * n_out + q) * k;
#pragma unroll
    for (int j = 0; j < C3P_MAX_K; ++j) col[j] = j < k ? (int)__ldg(ip + j) : 0;
    const float kf = (float)k;                          // torch.mean divides: keep the division
    const int p1 = min(p0 + C3P_ROWS, n1);
    for (int p = p0; p < p1; ++p) {
        const float* row = vol_in + ((size_t)b * n1 + p) * n_in;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < C3P_MAX_K; ++j)
            if (j < k) acc += __ldg(row + col[j]);
        vol_out[((size_t)b * n1 + p) * n_out + q] = acc / kf;
    }
}

}  // namespace

extern "C" int camli_corr3d_lookup(int B, int n1, int n_levels, const float* xyz1,
                                   const float* const* xyz2_levels_host, const int64_t* xyz2_strides_host,
                                   const int* n2_host, const float* const* volumes_host,
                                   const float* W1, const float* b1, const float* W2, const float* b2,
                                   float* out_rows, int ld_out, void* stream) {
    if (B < 0 || n1 < 0 || n_levels < 1 || ld_out < n_levels * C3_H) return CAMLI_EINVAL;
    if (n_levels > C3_MAX_LEVELS || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || n1 == 0) return CAMLI_OK;
    if (!xyz1 || !xyz2_levels_host || !xyz2_strides_host || !n2_host || !volumes_host || !W1 || !b1 || !W2 || !b2 ||
        !out_rows) return CAMLI_EINVAL;
    Corr3dLevels lv;
    for (int l = 0; l < n_levels; ++l) {
        if (!xyz2_levels_host[l] || !volumes_host[l]) return CAMLI_EINVAL;
        if (n2_host[l] < C3_K) return CAMLI_EUNSUPPORTED;
        lv.xyz2[l] = xyz2_levels_host[l];
        lv.sb[l] = xyz2_strides_host[3 * l]; lv.sp[l] = xyz2_strides_host[3 * l + 1]; lv.sd[l] = xyz2_strides_host[3 * l + 2];
        lv.vol[l] = volumes_host[l]; lv.n2[l] = n2_host[l];
    }
    dim3 grid(camli_div_up(n1, C3_WARPS), n_levels, B);
    corr3d_lookup_kernel<<<grid, C3_WARPS * 32, 0, (cudaStream_t)stream>>>(lv, n1, xyz1, W1, b1, W2, b2, out_rows, ld_out);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_corr3d_pool(int B, int n1, int n_in, int n_out, int k, const float* vol_in,
                                 const int64_t* knn_idx, float* vol_out, void* stream) {
    if (B < 0 || n1 < 0 || n_in < 1 || n_out < 0 || k < 1) return CAMLI_EINVAL;
    if (B > 65535 || camli_div_up(n1, C3P_ROWS) > 65535 || k > C3P_MAX_K) return CAMLI_EUNSUPPORTED;
    if (B == 0 || n1 == 0 || n_out == 0) return CAMLI_OK;
    if (!vol_in || !knn_idx || !vol_out) return CAMLI_EINVAL;
    dim3 grid(camli_div_up(n_out, 256), camli_div_up(n1, C3P_ROWS), B);
    corr3d_pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n1, n_in, n_out, k, vol_in, knn_idx, vol_out);
    CAMLI_RETURN_LAUNCH_STATUS();
}
