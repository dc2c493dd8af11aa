// Micro-benchmark of tcgen05.mma kind::tf32 issue patterns (diagnosis tool, not part of the library):
// cycles per instruction for M = 128, K = 8 as a function of N, of whether consecutive instructions
// accumulate into the SAME TMEM columns (dependent chain) or rotate over independent column ranges,
// and of the A operand source (shared memory descriptor vs TMEM).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o camliflow_b200/_build/mma_probe scripts/probes/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../camliflow_b200/csrc/tcgen05.cuh"

using namespace camli_tc;

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// N of every instruction; ROT independent accumulator ranges rotated over (1 = dependent chain); TS: A from TMEM;
// PAIR: the conv_gemm pattern -- a wide instruction (N) into [D, D+N) followed by a half-width one (N/2) into
// [D+N/2, D+N) (PAIR = 1, dependent) or into a separate range (PAIR = 2, independent)
template <int N, int ROT, int TS, int PAIR>
__global__ void __launch_bounds__(128, 1) probe_kernel(long long* out) {
    constexpr int COUNT = 64;
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 65536) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = tf32_idesc(128, N), idesc_h = tf32_idesc(128, N / 2);
        const uint64_t a = make_kmajor_sw128_desc(smem_u32(smem)), b = make_kmajor_sw128_desc(smem_u32(smem) + 16384);
        const uint32_t a_tmem = tmem + 480;
        uint32_t phase = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < COUNT; ++i) {
                const uint64_t adv = (uint64_t)((i & 3) * 2);
                if (PAIR == 0) {
                    const uint32_t d = tmem + (uint32_t)((i % ROT) * N);
                    if (TS) mma_tf32_ts(d, a_tmem, b + adv, idesc, i >= ROT ? 1u : 0u);
                    else mma_tf32(d, a + adv, b + adv, idesc, i >= ROT ? 1u : 0u);
                } else {
                    if (i & 1) mma_tf32(tmem + (PAIR == 1 ? N / 2 : N), a + adv, b + adv, idesc_h, 1u);
                    else mma_tf32(tmem, a + adv, b + adv, idesc, i >= 2 ? 1u : 0u);
                }
            }
            const long long t1 = clock64();
            mma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), phase);
            phase ^= 1;
            const long long t2 = clock64();
            out[rep * 2] = t1 - t0;
            out[rep * 2 + 1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N, int ROT, int TS, int PAIR>
void run(long long* d_out) {
    const int smem = 16384 + 65536 + 1024;
    cudaFuncSetAttribute(probe_kernel<N, ROT, TS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<N, ROT, TS, PAIR><<<1, 128, smem>>>(d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d rot=%d ts=%d pair=%d: %s\n", N, ROT, TS, PAIR, cudaGetErrorString(e)); exit(1); }
    long long h[6];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%5d %4d %3d %4d | %10.1f %10.1f\n", N, ROT, TS, PAIR, (double)h[4] / 64, (double)h[5] / 64);
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    printf("%5s %4s %3s %4s | %10s %10s | cycles per instruction (issue / complete), last of 3 repetitions\n", "N", "rot", "ts", "pair", "issue", "complete");
    run<32, 1, 0, 0>(d_out); run<32, 2, 0, 0>(d_out); run<32, 4, 0, 0>(d_out);
    run<64, 1, 0, 0>(d_out); run<64, 2, 0, 0>(d_out); run<64, 4, 0, 0>(d_out);
    run<128, 1, 0, 0>(d_out); run<128, 2, 0, 0>(d_out); run<128, 3, 0, 0>(d_out);
    run<256, 1, 0, 0>(d_out);
    run<64, 1, 1, 0>(d_out); run<64, 2, 1, 0>(d_out); run<128, 1, 1, 0>(d_out); run<128, 2, 1, 0>(d_out); run<256, 1, 1, 0>(d_out);
    run<256, 1, 0, 1>(d_out); run<256, 1, 0, 2>(d_out); run<128, 1, 0, 1>(d_out); run<128, 1, 0, 2>(d_out);
    return 0;
}
