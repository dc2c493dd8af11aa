// Three-NN inverse-distance interpolation and 3-D back-warping for sm_100a, fused with the
// neighbour search.
//
// Replaces knn_interpolation / backwarp_3d (reference models/utils.py:130-159): 1 k-NN launch,
// 2 channel-first gathers (4-byte strided reads), norm, clamp, reciprocal, normalise, multiply,
// sum -- ~10 launches and three [B,F,n,k] temporaries per call, 5 calls per GRU iteration.
// Here one warp searches its query's k (<= 32) neighbours (bit-exact order, knn_search.cuh) and
// immediately blends the neighbours' features: w_j = (1/max(|p_j - q|, 1e-8)) / sum_j(...).
// Algorithmic bytes: B*(n+m)*12 + B*n*k*F*4 gathered + B*n*F*4 written.
#include "knn_search.cuh"

namespace {

constexpr int TNN_WARPS = KNN_WARPS;

// Weights of the k neighbours held one per lane (lanes >= k get 0).
__device__ __forceinline__ float tnn_weight(int lane, int k, float px, float py, float pz,
                                            float ux, float uy, float uz) {
    float w = 0.f;
    if (lane < k) {
        const float dx = px - ux, dy = py - uy, dz = pz - uz;
        const float d = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-8f);   // linalg.norm(...).clamp(1e-8)
        w = 1.0f / d;
    }
    const float sum = camli_warp_sum(w);
    return w / sum;
}

// out[b,f,q] = sum_j w_j * feat[b,f,idx_j]   (all tensors channel-first with explicit strides)
__global__ void __launch_bounds__(TNN_WARPS * 32)
three_nn_interp_kernel(int n, int m, int k, int F,
                       const float* __restrict__ query, KnnView qv, const float* __restrict__ input, KnnView iv,
                       const float* __restrict__ feat, long long f_sb, long long f_sc, long long f_sp,
                       float* __restrict__ out, long long o_sb, long long o_sc, long long o_sp) {
    __shared__ KnnTile tile;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * TNN_WARPS + (threadIdx.x >> 5);
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
    const float w = tnn_weight(lane, k, px, py, pz, ux, uy, uz);
    const float* fb = feat + b * f_sb;
    float* ob = out + b * o_sb + q * o_sp;
    for (int f0 = 0; f0 < F; f0 += 32) {
        const int f = f0 + lane;
        float acc = 0.f;
        for (int j = 0; j < k; ++j) {
            const int ij = __shfl_sync(CAMLI_FULL_MASK, r.i0, j);
            const float wj = __shfl_sync(CAMLI_FULL_MASK, w, j);
            if (f < F) acc += __ldg(fb + f * f_sc + ij * f_sp) * wj;   // same order as torch.sum over k
        }
        if (f < F) ob[f * o_sc] = acc;
    }
}

// xyz2_warp[b,:,q] = xyz2[b,:,q] + sum_j w_j * (-flow[b,:,idx_j]),  neighbours searched in xyz1 + flow
// (backwarp_3d, models/utils.py:149-159).  All tensors [B,3,N] channel-first, contiguous.
__global__ void __launch_bounds__(TNN_WARPS * 32)
backwarp3d_kernel(int n, int m, int k, const float* __restrict__ xyz1, const float* __restrict__ flow,
                  const float* __restrict__ xyz2, float* __restrict__ out) {
    __shared__ KnnTile tile;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * TNN_WARPS + (threadIdx.x >> 5);
    const bool active = q < n;
    const int b = blockIdx.y;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float* qp = xyz2 + (size_t)b * 3 * n + q;
        ux = __ldg(qp); uy = __ldg(qp + n); uz = __ldg(qp + 2 * n);
    }
    const float* p1 = xyz1 + (size_t)b * 3 * m;
    const float* fl = flow + (size_t)b * 3 * m;
    const KnnDisplacedPoints pts{p1, fl, m};
    KnnList r;
    knn_cta_search<3, 1>(r, tile, pts, m, k, active, ux, uy, uz);
    if (!active) return;

    float px = 0.f, py = 0.f, pz = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
    if (lane < k) {
        fx = __ldg(fl + r.i0); fy = __ldg(fl + m + r.i0); fz = __ldg(fl + 2 * m + r.i0);
        px = __fadd_rn(__ldg(p1 + r.i0), fx);
        py = __fadd_rn(__ldg(p1 + m + r.i0), fy);
        pz = __fadd_rn(__ldg(p1 + 2 * m + r.i0), fz);
    }
    const float w = tnn_weight(lane, k, px, py, pz, ux, uy, uz);
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int j = 0; j < k; ++j) {
        const float wj = __shfl_sync(CAMLI_FULL_MASK, w, j);
        ax += -__shfl_sync(CAMLI_FULL_MASK, fx, j) * wj;
        ay += -__shfl_sync(CAMLI_FULL_MASK, fy, j) * wj;
        az += -__shfl_sync(CAMLI_FULL_MASK, fz, j) * wj;
    }
    if (lane == 0) {
        float* o = out + (size_t)b * 3 * n + q;
        o[0] = ux + ax; o[n] = uy + ay; o[2 * n] = uz + az;
    }
}

}  // namespace

extern "C" int camli_three_nn_interpolate(int B, int n, int m, int k, int F,
                                          const float* query_xyz, int64_t q_sb, int64_t q_sp, int64_t q_sd,
                                          const float* input_xyz, int64_t i_sb, int64_t i_sp, int64_t i_sd,
                                          const float* input_feat, int64_t f_sb, int64_t f_sc, int64_t f_sp,
                                          float* out, int64_t o_sb, int64_t o_sc, int64_t o_sp, void* stream) {
    if (B < 0 || n < 0 || m < 1 || k < 1 || F < 0) return CAMLI_EINVAL;
    if (k > 32 || k > m || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || n == 0 || F == 0) return CAMLI_OK;
    if (!query_xyz || !input_xyz || !input_feat || !out) return CAMLI_EINVAL;
    dim3 grid(camli_div_up(n, TNN_WARPS), B);
    three_nn_interp_kernel<<<grid, TNN_WARPS * 32, 0, (cudaStream_t)stream>>>(
        n, m, k, F, query_xyz, KnnView{q_sb, q_sp, q_sd}, input_xyz, KnnView{i_sb, i_sp, i_sd},
        input_feat, f_sb, f_sc, f_sp, out, o_sb, o_sc, o_sp);
    CAMLI_RETURN_LAUNCH_STATUS();
}

extern "C" int camli_backwarp_3d(int B, int n, int m, int k, const float* xyz1, const float* flow12,
                                 const float* xyz2, float* xyz2_warp, void* stream) {
    if (B < 0 || n < 0 || m < 1 || k < 1) return CAMLI_EINVAL;
    if (k > 32 || k > m || B > 65535) return CAMLI_EUNSUPPORTED;
    if (B == 0 || n == 0) return CAMLI_OK;
    if (!xyz1 || !flow12 || !xyz2 || !xyz2_warp) return CAMLI_EINVAL;
    dim3 grid(camli_div_up(n, TNN_WARPS), B);
    backwarp3d_kernel<<<grid, TNN_WARPS * 32, 0, (cudaStream_t)stream>>>(n, m, k, xyz1, flow12, xyz2, xyz2_warp);
    CAMLI_RETURN_LAUNCH_STATUS();
}
