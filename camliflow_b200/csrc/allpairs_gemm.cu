// All-pairs feature inner product on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the fp32 `torch.matmul(fmap1^T, fmap2) / sqrt(C)` of Correlation2D.build_cost_volume_pyramid
// (reference models/raft_core.py:56-63; 34.1 GFLOP and a 266 MB result per 960x540 pair) and the
// `torch.bmm(feat1^T, feat2) / C` of Correlation3D (models/camliraft_l_core.py:52-53).
//
//   out[b, m, n] = scale * sum_k A[b, m, k] * Bm[b, n, k]        A: [B*M, K], Bm: [B*N, K] row-major fp32
//
// Precision.  The reference computes this product in fp32 (torch keeps TF32 off for matmul), and
// the volume feeds 12-32 recurrent refinements, so a plain TF32 product (10-bit mantissa) is not
// accurate enough for the 1e-3 px parity bar.  Every operand is therefore split once into
// x = hi + lo with hi = tf32(x) (round-to-nearest) and lo = tf32(x - hi), and the kernel
// accumulates hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator ("3xTF32"): the dropped lo*lo term
// is ~2^-22 relative, i.e. fp32-level accuracy at tensor-core speed.
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0    : TMA producer -- per k-block (32 fp32 = one 128-byte swizzle atom) four 128x32 tiles
//               (A_hi, A_lo, B_hi, B_lo) land in a 3-stage shared-memory ring (64 KB per stage),
//               completion on an mbarrier (expect_tx);
//   warp 1    : MMA issuer -- one elected lane issues tcgen05.mma kind::tf32, M=128 N=128 K=8,
//               12 per k-block (4 k-steps x {hi*hi, hi*lo, lo*hi}) into one of two 128-column TMEM
//               accumulators; tcgen05.commit releases the smem stage / publishes the accumulator;
//   warps 2-5 : epilogue -- tcgen05.ld (32 lanes x 32 columns per instruction), scale, 128-bit
//               stores of each thread's row segment straight to the [B,M,N] volume, overlapping
//               the next tile's MMAs through the second accumulator.
// Rows/columns past M/N are zero-filled by TMA on load and masked on store.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 32;          // BK fp32 = 128 bytes = one SWIZZLE_128B row
constexpr int GM_STAGES = 3;
constexpr int GM_TILE_BYTES = GM_BM * GM_BK * 4;             // 16 KB
constexpr int GM_STAGE_BYTES = 4 * GM_TILE_BYTES;            // A_hi, A_lo, B_hi, B_lo
constexpr int GM_THREADS = 192;
constexpr int GM_TMEM_COLS = 2 * GM_BN;                      // two fp32 accumulators
constexpr int GM_SMEM_BYTES = GM_STAGES * GM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

using namespace camli_tc;

constexpr uint32_t GM_IDESC = tf32_idesc(GM_BM, GM_BN);

__global__ void __launch_bounds__(GM_THREADS, 1)
allpairs_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                       float* __restrict__ out, int B, int M, int N, int K, float scale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GM_STAGES * GM_STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GM_STAGES + 4);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + GM_STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * GM_STAGES), bar_tempty = smem_u32(bars + 2 * GM_STAGES + 2);
    const uint32_t tiles_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GM_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {      // whole warp: TMEM allocation (and, at the end, release)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)GM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_m = (M + GM_BM - 1) / GM_BM, tiles_n = (N + GM_BN - 1) / GM_BN;
    const int total = B * tiles_m * tiles_n;
    const int kblocks = K / GM_BK;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int b = tile / (tiles_m * tiles_n), r = tile - b * tiles_m * tiles_n;
                const int row_a = b * M + (r / tiles_n) * GM_BM, row_b = b * N + (r % tiles_n) * GM_BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage;
                    const uint32_t dst = tiles_base + stage * GM_STAGE_BYTES;
                    mbar_expect_tx(full, GM_STAGE_BYTES);
                    tma_load_2d(dst + 0 * GM_TILE_BYTES, &map_a_hi, full, kb * GM_BK, row_a);
                    tma_load_2d(dst + 1 * GM_TILE_BYTES, &map_a_lo, full, kb * GM_BK, row_a);
                    tma_load_2d(dst + 2 * GM_TILE_BYTES, &map_b_hi, full, kb * GM_BK, row_b);
                    tma_load_2d(dst + 3 * GM_TILE_BYTES, &map_b_lo, full, kb * GM_BK, row_b);
                    if (++stage == GM_STAGES) { stage = 0; phase ^= 1; }
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
                const int acc = it & 1;
                mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);     // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * GM_BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t src = tiles_base + stage * GM_STAGE_BYTES;
                    const uint64_t a_hi = make_kmajor_sw128_desc(src), a_lo = make_kmajor_sw128_desc(src + GM_TILE_BYTES);
                    const uint64_t b_hi = make_kmajor_sw128_desc(src + 2 * GM_TILE_BYTES),
                                   b_lo = make_kmajor_sw128_desc(src + 3 * GM_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < GM_BK / 8; ++k) {              // UMMA_K = 8 tf32 = 32 bytes = +2 in the address field
                        const uint64_t adv = (uint64_t)(k * 2);
                        mma_tf32(tmem_d, a_hi + adv, b_hi + adv, GM_IDESC, (kb | k) ? 1u : 0u);
                        mma_tf32(tmem_d, a_hi + adv, b_lo + adv, GM_IDESC, 1u);
                        mma_tf32(tmem_d, a_lo + adv, b_hi + adv, GM_IDESC, 1u);
                    }
                    mma_commit(bar_empty + 8 * stage);                 // smem stage reusable once these MMAs retire
                    if (++stage == GM_STAGES) { stage = 0; phase ^= 1; }
                }
                mma_commit(bar_tfull + 8 * acc);                       // accumulator complete
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int b = tile / (tiles_m * tiles_n), r = tile - b * tiles_m * tiles_n;
            const int m0 = (r / tiles_n) * GM_BM, n0 = (r % tiles_n) * GM_BN;
            const int acc = it & 1;
            mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            float* __restrict__ orow = out + ((size_t)b * M + row) * N + n0;
            const bool vec_ok = ((((size_t)b * M + row) * N + n0) & 3) == 0;
#pragma unroll 1
            for (int c = 0; c < GM_BN / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + acc * GM_BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
                if (row < M) {
                    const int col0 = n0 + c * 32;
                    if (vec_ok && col0 + 32 <= N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            __stcs(reinterpret_cast<float4*>(orow + c * 32 + j),
                                   make_float4(v[j] * scale, v[j + 1] * scale, v[j + 2] * scale, v[j + 3] * scale));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < N) orow[c * 32 + j] = v[j] * scale;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)GM_TMEM_COLS) : "memory");
    }
}

// x -> (hi, lo): hi = tf32(x) round-to-nearest, lo = tf32(x - hi)
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float in[4] = {v.x, v.y, v.z, v.w}, h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[j]));
        h[j] = __uint_as_float(hb);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(in[j] - h[j]));
        l[j] = __uint_as_float(lb);
    }
    reinterpret_cast<float4*>(hi)[i] = make_float4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<float4*>(lo)[i] = make_float4(l[0], l[1], l[2], l[3]);
}

// [rows, K] row-major fp32, box = 128 rows x 32 columns (128 bytes), 128-byte swizzle, zero fill outside.
int make_operand_map(CUtensorMap* map, const float* base, long long rows, int K) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    const cuuint32_t box[2] = {GM_BK, GM_BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace

extern "C" int64_t camli_allpairs_workspace_floats(int B, int M, int N, int K) {
    return 2 * ((int64_t)B * M * K + (int64_t)B * N * K);
}

extern "C" int camli_allpairs_correlation(const float* a_rows, const float* b_rows, float* workspace, float* out,
                                          int B, int M, int N, int K, float scale, void* stream) {
    if (B < 0 || M < 1 || N < 1 || K < 1) return CAMLI_EINVAL;
    if (K % GM_BK != 0 || (long long)B * M > 2147483647LL - GM_BM || (long long)B * N > 2147483647LL - GM_BN)
        return CAMLI_EUNSUPPORTED;
    if (B == 0) return CAMLI_OK;
    if (!a_rows || !b_rows || !workspace || !out) return CAMLI_EINVAL;
    if ((reinterpret_cast<uintptr_t>(a_rows) | reinterpret_cast<uintptr_t>(b_rows) | reinterpret_cast<uintptr_t>(workspace)) & 15)
        return CAMLI_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t na = (size_t)B * M * K, nb = (size_t)B * N * K;
    float *a_hi = workspace, *a_lo = a_hi + na, *b_hi = a_lo + na, *b_lo = b_hi + nb;
    split_tf32_kernel<<<(unsigned)((na / 4 + 255) / 256), 256, 0, st>>>(a_rows, a_hi, a_lo, na / 4);
    split_tf32_kernel<<<(unsigned)((nb / 4 + 255) / 256), 256, 0, st>>>(b_rows, b_hi, b_lo, nb / 4);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;

    CUtensorMap m_ahi, m_alo, m_bhi, m_blo;
    int rc;
    if ((rc = make_operand_map(&m_ahi, a_hi, (long long)B * M, K))) return rc;
    if ((rc = make_operand_map(&m_alo, a_lo, (long long)B * M, K))) return rc;
    if ((rc = make_operand_map(&m_bhi, b_hi, (long long)B * N, K))) return rc;
    if ((rc = make_operand_map(&m_blo, b_lo, (long long)B * N, K))) return rc;

    const int n_sms = sm_count();
    e = cudaFuncSetAttribute(allpairs_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const long long total = (long long)B * ((M + GM_BM - 1) / GM_BM) * ((N + GM_BN - 1) / GM_BN);
    const int grid = (int)(total < n_sms ? total : n_sms);
    allpairs_tf32x3_kernel<<<grid, GM_THREADS, GM_SMEM_BYTES, st>>>(m_ahi, m_alo, m_bhi, m_blo, out, B, M, N, K, scale);
    CAMLI_RETURN_LAUNCH_STATUS();
}
