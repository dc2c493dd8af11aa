/*
 * camli_b200.h -- C ABI of libcamli_b200.so (hand-written sm_100a kernels for the
 * CamLiFlow / CamLiRAFT fused 2D-3D hot path).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - every entry point takes the CUDA stream to launch on as `void* stream`
 *     (a cudaStream_t; NULL = legacy default stream) and is graph-capturable:
 *     no allocation, no host synchronisation, no host<->device copies;
 *   - return value: 0 on success, a positive cudaError_t value if a launch
 *     failed, a negative CAMLI_E* value for an argument error (nothing launched);
 *   - tensors are dense row-major in the layout given in brackets; indices are
 *     int64 where the reference returns int64 ("i64"), int32 otherwise;
 *   - "reference" citations are relative to MCG-NJU/CamLiFlow @3bf1974.
 */
#ifndef CAMLI_B200_H
#define CAMLI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAMLI_OK            0
#define CAMLI_EINVAL       -1   /* bad size / null pointer                        */
#define CAMLI_EUNSUPPORTED -2   /* valid in the reference but outside kernel limits */

/* ABI version, bumped whenever a signature changes. */
int camli_abi_version(void);

/* Human-readable text for a return code of this library (static storage). */
const char* camli_strerror(int code);

/* ------------------------------------------------------------------------- *
 * L0: the three native extensions of models/csrc
 * ------------------------------------------------------------------------- */

/*
 * Furthest point sampling.
 * Replaces furthest_point_sampling_kernel_wrapper(float* xyz, float* dists_tmp,
 *   int B, int N, int S, int64_t* out)      models/csrc/furthest_point_sampling/
 *   furthest_point_sampling.cpp:3, kernel furthest_point_sampling_kernel.cu:34-84.
 * xyz [B,N,3] f32, out [B,S] i64.  First sample is index 0; ties are resolved
 * exactly as the reference's 1024-thread shared-memory tree does (largest
 * bit-reversed (i mod 1024), then smallest i).
 * dists_tmp: [B,N] f32 scratch, only touched when N > 8192 (may be NULL
 * otherwise); contents on entry are ignored (the library initialises it).
 */
int camli_furthest_point_sampling(const float* xyz, float* dists_tmp,
                                  int B, int N, int S, int64_t* out, void* stream);

/*
 * Brute-force exact k nearest neighbours, ascending distance.
 * Replaces k_nearest_neighbor_{2d,3d}_kernel_wrapper(int b,int n,int m,int k,
 *   const float* query,const float* input,int64_t* idx)
 *   models/csrc/k_nearest_neighbor/k_nearest_neighbor.cpp:3-4, kernels
 *   k_nearest_neighbor_kernel.cu:9-95.
 * query [B,n,D], input [B,m,D] f32 with D in {2,3}; idx [B,n,k] i64; 1<=k<=64.
 * Bit-exact with the reference's sequential insertion (including its
 * equal-to-worst-replaces rule and index 0 for unfilled slots when m<k).
 */
int camli_k_nearest_neighbor(int B, int n, int m, int k, int D,
                             const float* query, const float* input,
                             int64_t* idx, void* stream);

/*
 * Same search on strided views: element strides (batch, point, dim) of query and
 * input, so the models' channel-first [B,D,N] tensors are searched in place instead
 * of through the transpose+contiguous copy of models/csrc/wrapper.py:119-122.
 */
int camli_k_nearest_neighbor_strided(int B, int n, int m, int k, int D,
                                     const float* query, int64_t q_stride_b, int64_t q_stride_pt, int64_t q_stride_dim,
                                     const float* input, int64_t i_stride_b, int64_t i_stride_pt, int64_t i_stride_dim,
                                     int64_t* idx, void* stream);

/*
 * PWC local cost volume, forward.
 * Replaces correlation_forward_kernel_wrapper(float* out,const float* in1,
 *   const float* in2,int B,int C,int H,int W,int d)
 *   models/csrc/correlation/correlation.cpp:3, kernel correlation_forward_kernel.cu:11-55.
 * in1,in2 [B,H,W,C] f32 (NHWC); out [B,(2d+1)^2,H,W] f32, fully written
 * (out-of-range displacements are written as 0; no pre-zeroing needed).
 * channel tc = (dy+d)*(2d+1)+(dx+d), value = (1/C) sum_c in1[y,x,c]*in2[y+dy,x+dx,c].
 */
int camli_correlation_forward(float* out, const float* in1, const float* in2,
                              int B, int C, int H, int W, int max_displacement,
                              void* stream);

/*
 * PWC local cost volume, backward.
 * Replaces correlation_backward_kernel_wrapper(const float* gO,float* g1,float* g2,
 *   const float* in1,const float* in2,int B,int C,int H,int W,int d)
 *   models/csrc/correlation/correlation.cpp:6-9, kernels correlation_backward_kernel.cu:4-89.
 * grad_out [B,(2d+1)^2,H,W]; in1,in2 [B,H,W,C]; grad_in1, grad_in2 [B,C,H,W] (NCHW,
 * like the reference).
 */
int camli_correlation_backward(const float* grad_out, float* grad_in1, float* grad_in2,
                               const float* in1, const float* in2,
                               int B, int C, int H, int W, int max_displacement,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAMLI_B200_H */
