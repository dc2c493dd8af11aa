// Shared device/host helpers for libcamli_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/camli_b200.h"

#define CAMLI_FULL_MASK 0xffffffffu

// Launch epilogue: report a failed launch through the C-ABI return code.
#define CAMLI_RETURN_LAUNCH_STATUS()                      \
    do {                                                  \
        cudaError_t e__ = cudaGetLastError();             \
        return e__ == cudaSuccess ? CAMLI_OK : (int)e__;  \
    } while (0)

static inline int camli_div_up(int a, int b) { return (a + b - 1) / b; }
static inline long long camli_div_up_ll(long long a, long long b) { return (a + b - 1) / b; }

// Squared distance with the exact operation order nvcc emits (sm_100 SASS of the
// reference build, checked with cuobjdump) for the reference's
// `dx*dx + dy*dy + dz*dz` (k_nearest_neighbor_kernel.cu:79,
// furthest_point_sampling_kernel.cu:62): the compiler rounds dy*dy (FMUL), fuses
// dx*dx into it and then dz*dz:  fma(dz,dz, fma(dx,dx, dy*dy)); in 2-D
// (k_nearest_neighbor_kernel.cu:35) fma(dx,dx, dy*dy).  Explicit intrinsics so no
// compiler flag can change the rounding.  The order matters: on the reference's
// own FPS self-test recipe one pair of candidates is an exact tie under one order
// and one ulp apart under the other (tests/test_oracle_kernels.py).
__device__ __forceinline__ float camli_sqdist3(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
__device__ __forceinline__ float camli_sqdist2(float dx, float dy) {
    return __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
}

__device__ __forceinline__ float camli_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CAMLI_FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ float camli_leaky(float v, float slope) { return v > 0.f ? v : v * slope; }
