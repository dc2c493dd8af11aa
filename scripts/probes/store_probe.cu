// Micro-benchmark (diagnosis tool): how fast can ONE CTA move a 64 KB tile from shared memory to global memory --
// (a) LSU: 256 threads, coalesced STG.128; (b) TMA: cp.async.bulk shared -> global in 1 / 16 / 128 pieces.
// Cycles from first store to completion (membar / bulk wait_group), per CTA, with 1, 16 and 148 CTAs running.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o camliflow_b200/_build/store_probe scripts/probes/store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int TILE = 65536;

__global__ void __launch_bounds__(256, 1) probe(float* out, long long* cycles, int mode, int pieces) {
    extern __shared__ __align__(128) uint8_t smem[];
    float4* s4 = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < TILE / 16; i += 256) s4[i] = make_float4(1.f, 2.f, 3.f, (float)i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    float* dst = out + (size_t)blockIdx.x * (TILE / 4);
    long long t0 = clock64();
    if (mode == 0) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < TILE / 16; i += 256) d4[i] = s4[i];
        __threadfence();
    } else {
        if (threadIdx.x == 0) {
            const int bytes = TILE / pieces;
            for (int p = 0; p < pieces; ++p) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem + p * bytes);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint8_t*>(dst) + p * bytes), "r"(src), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, (size_t)148 * TILE);
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE);
    for (int ctas : {1, 16, 68, 148})
        for (int mode = 0; mode < 2; ++mode)
            for (int pieces : {1, 16, 128}) {
                if (mode == 0 && pieces != 1) continue;
                long long h[148];
                for (int rep = 0; rep < 3; ++rep) {
                    probe<<<ctas, 256, TILE>>>(out, cyc, mode, pieces);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
                long long mx = 0, sum = 0;
                for (int i = 0; i < ctas; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
                printf("ctas %3d  %-4s pieces %3d : avg %6lld max %6lld cycles per 64 KB  (%.1f B/clk/SM)\n", ctas, mode ? "TMA" : "LSU",
                       pieces, sum / ctas, mx, 65536.0 / (sum / ctas));
            }
    return 0;
}
