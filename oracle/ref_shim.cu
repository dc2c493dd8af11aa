// ref_shim.cu -- TEST INFRASTRUCTURE ONLY.
// extern "C" doorway onto the reference's own, unmodified CUDA kernels.  The kernel
// sources are compiled where they lie under /root/reference/models/csrc (never copied
// into this repo) by oracle/build.py into oracle/_ref/libref_kernels.so; this shim
// only declares the C++ launch wrappers those files define and re-exports them with
// C linkage so ctypes can call them.  They launch on the legacy default stream.
#include <stdint.h>
#include <cuda_runtime.h>

// furthest_point_sampling/furthest_point_sampling_kernel.cu:81
void furthest_point_sampling_kernel_wrapper(float* batched_points_xyz, float* batched_dists_temp, int n_batch,
                                            int n_points, int n_samples, int64_t* batched_furthest_indices);
// k_nearest_neighbor/k_nearest_neighbor_kernel.cu:97,106
void k_nearest_neighbor_2d_kernel_wrapper(int b, int n, int m, int k, const float* query_xyz,
                                          const float* input_xyz, int64_t* indices);
void k_nearest_neighbor_3d_kernel_wrapper(int b, int n, int m, int k, const float* query_xyz,
                                          const float* input_xyz, int64_t* indices);
// correlation/correlation_forward_kernel.cu:51, correlation_backward_kernel.cu:76
void correlation_forward_kernel_wrapper(float* output, const float* input1, const float* input2, int n_batches,
                                        int in_channels, int height, int width, int max_displacement);
void correlation_backward_kernel_wrapper(const float* grad_output, float* grad_input1, float* grad_input2,
                                         const float* input1, const float* input2, int n_batches,
                                         int in_channels, int height, int width, int max_displacement);

extern "C" {
int ref_fps(float* xyz, float* dists_tmp, int B, int N, int S, int64_t* out) {
    furthest_point_sampling_kernel_wrapper(xyz, dists_tmp, B, N, S, out);
    return (int)cudaGetLastError();
}
int ref_knn(int B, int n, int m, int k, int D, const float* query, const float* input, int64_t* idx) {
    if (D == 2) k_nearest_neighbor_2d_kernel_wrapper(B, n, m, k, query, input, idx);
    else        k_nearest_neighbor_3d_kernel_wrapper(B, n, m, k, query, input, idx);
    return (int)cudaGetLastError();
}
int ref_corr_fwd(float* out, const float* in1, const float* in2, int B, int C, int H, int W, int md) {
    correlation_forward_kernel_wrapper(out, in1, in2, B, C, H, W, md);
    return (int)cudaGetLastError();
}
int ref_corr_bwd(const float* gout, float* g1, float* g2, const float* in1, const float* in2, int B, int C, int H,
                 int W, int md) {
    correlation_backward_kernel_wrapper(gout, g1, g2, in1, in2, B, C, H, W, md);
    return (int)cudaGetLastError();
}
}
