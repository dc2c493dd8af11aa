// Exact brute-force k-nearest-neighbour search for sm_100a.
//
// Replaces k_nearest_neighbor_{2d,3d}_kernel (reference
// models/csrc/k_nearest_neighbor/k_nearest_neighbor_kernel.cu:9-95).
//
// The reference runs ONE THREAD per query with the k-entry result list in
// dynamically indexed local memory; at the working sizes of CamLiRAFT (n = 2048
// queries) that is 8 CTAs on a 148-SM part.  Here ONE WARP owns a query: the 32
// lanes evaluate 32 candidates per step (coalesced loads), and the sorted result
// list is distributed over the lanes' registers (entry e lives in lane e%32,
// slot e/32), so an insertion is one ballot + one shuffle-shift.  Candidates are
// still consumed in ascending input order and every insertion applies the
// reference's rule literally (skip if d > worst; place after all entries <= d,
// scanning down from j = min(idx, k-1); the last entry is dropped), so results
// are bit-identical, including on exact distance ties.
#include "knn_search.cuh"

namespace {

template <int D, int SLOTS>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_warp_kernel(int n, int m, int k, const float* __restrict__ query_all, KnnView qv,
                const float* __restrict__ input_all, KnnView iv, int64_t* __restrict__ idx_all) {
    __shared__ KnnTile tile;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    const bool active = q < n;
    const int b = blockIdx.y;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float* __restrict__ qp = query_all + b * qv.sb + q * qv.sp;
        ux = __ldg(qp); uy = __ldg(qp + qv.sd); uz = (D == 3) ? __ldg(qp + 2 * qv.sd) : 0.f;
    }
    const KnnPlainPoints<D> pts{input_all + b * iv.sb, iv.sp, iv.sd};
    KnnList r;
    knn_cta_search<D, SLOTS>(r, tile, pts, m, k, active, ux, uy, uz);
    if (!active) return;
    int64_t* __restrict__ out = idx_all + ((size_t)b * n + q) * k;
    if (lane < k) out[lane] = (int64_t)r.i0;
    if (SLOTS == 2 && lane + 32 < k) out[lane + 32] = (int64_t)r.i1;
}

template <int D>
int knn_launch(int B, int n, int m, int k, const float* query, KnnView qv, const float* input, KnnView iv,
               int64_t* idx, cudaStream_t st) {
    dim3 grid(camli_div_up(n, KNN_WARPS), B);
    dim3 block(KNN_WARPS * 32);
    if (k <= 32) knn_warp_kernel<D, 1><<<grid, block, 0, st>>>(n, m, k, query, qv, input, iv, idx);
    else         knn_warp_kernel<D, 2><<<grid, block, 0, st>>>(n, m, k, query, qv, input, iv, idx);
    CAMLI_RETURN_LAUNCH_STATUS();
}

}  // namespace

static int knn_dispatch(int B, int n, int m, int k, int D, const float* query, KnnView qv, const float* input,
                        KnnView iv, int64_t* idx, void* stream) {
    if (B < 0 || n < 0 || m < 0 || k < 1) return CAMLI_EINVAL;
    if (k > 64) return CAMLI_EUNSUPPORTED;      // reference MAX_K (k_nearest_neighbor_kernel.cu:5)
    if (D != 2 && D != 3) return CAMLI_EUNSUPPORTED;
    if (B == 0 || n == 0) return CAMLI_OK;
    if (B > 65535) return CAMLI_EUNSUPPORTED;
    if (!query || !idx || (m > 0 && !input)) return CAMLI_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    return D == 3 ? knn_launch<3>(B, n, m, k, query, qv, input, iv, idx, st)
                  : knn_launch<2>(B, n, m, k, query, qv, input, iv, idx, st);
}

extern "C" int camli_k_nearest_neighbor(int B, int n, int m, int k, int D, const float* query,
                                        const float* input, int64_t* idx, void* stream) {
    const KnnView qv{(long long)n * D, D, 1}, iv{(long long)m * D, D, 1};
    return knn_dispatch(B, n, m, k, D, query, qv, input, iv, idx, stream);
}

extern "C" int camli_k_nearest_neighbor_strided(int B, int n, int m, int k, int D, const float* query,
                                                int64_t q_stride_b, int64_t q_stride_pt, int64_t q_stride_dim,
                                                const float* input, int64_t i_stride_b, int64_t i_stride_pt,
                                                int64_t i_stride_dim, int64_t* idx, void* stream) {
    const KnnView qv{q_stride_b, q_stride_pt, q_stride_dim}, iv{i_stride_b, i_stride_pt, i_stride_dim};
    return knn_dispatch(B, n, m, k, D, query, qv, input, iv, idx, stream);
}
