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
 * dists_tmp: [B,N] f32 scratch, only touched when N > 16384 (or N > 8192 with the cluster
 * path disabled); may be NULL otherwise; contents on entry are ignored.
 */
int camli_furthest_point_sampling(const float* xyz, float* dists_tmp,
                                  int B, int N, int S, int64_t* out, void* stream);

/* Selects the FPS kernel for 2048 < N <= 16384: 3 = one CTA per cloud over Morton-ordered buckets with an exact
 * bounding-box pruning test per warp and round (default; N <= 8192, larger clouds take 2), 2 = 8-CTA thread-block
 * cluster exchanging the per-round records with st.async + mbarrier, 1 = 8-CTA cluster with a cluster barrier per
 * round, 0 = single-CTA register kernel.  Returns the previous setting.  Results are identical. */
int camli_fps_set_cluster_path(int mode);

/* Diagnostics: a device buffer of >= 20 int64 that thread 0 of CTA 0 of the async-cluster FPS kernel stamps with
 * SM-clock values of rounds 100..103 (5 events per round); NULL detaches (default). */
int camli_fps_set_timeline(long long* device_buffer);

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


/* ------------------------------------------------------------------------- *
 * L1: fused hot-path operators of the CamLiRAFT cores
 *   "rows" = channel-last storage: [B, points, ld] f32 with ld >= channels.
 *   Arguments named *_host are HOST arrays (of device pointers / sizes).
 * ------------------------------------------------------------------------- */

/*
 * RAFT correlation lookup, all pyramid levels in one launch.
 * Replaces Correlation2D.forward (models/raft_core.py:71-107).
 * volumes_host[l]: device pointer of level l, [B, H*W, level_h[l], level_w[l]] f32;
 * coords [B,2,H,W] (x then y, in level-0 pixels); radius must be 4.
 * out: [B, n_levels*81, H, W] f32 if out_nhwc == 0, else NHWC [B, H, W, n_levels*81].
 * channel = l*81 + i*9 + j, i offsets x and j offsets y (the reference's order);
 * bilinear, align_corners, zero padding.
 */
int camli_corr2d_lookup(const float* const* volumes_host, const int* level_h_host, const int* level_w_host,
                        int n_levels, const float* coords, float* out, int B, int H, int W,
                        int radius, int out_nhwc, void* stream);

/*
 * Every coarser level of the correlation-volume pyramid from one read of level 0:
 * level l = 2x2 average pooling (stride 2, floor) of level l-1 over the trailing (h, w) axes,
 * the avg_pool2d loop of models/raft_core.py:65-68.  vol0 [rows, h0, w0]; coarser_host[l-1]:
 * device pointer of level l, [rows, h0>>l, w0>>l], l = 1..n_levels-1.
 */
int camli_corr2d_pool_pyramid(const float* vol0, float* const* coarser_host, int n_levels, int64_t rows,
                              int h0, int w0, void* stream);

/*
 * Three-NN (k <= 32) inverse-distance interpolation fused with the neighbour search.
 * Replaces knn_interpolation (models/utils.py:130-146).  All tensors are addressed through
 * element strides: query_xyz/input_xyz as [B, points, 3] views (batch, point, dim),
 * input_feat/out as [B, F, points] views (batch, channel, point).
 */
int camli_three_nn_interpolate(int B, int n, int m, int k, int F,
                               const float* query_xyz, int64_t q_sb, int64_t q_sp, int64_t q_sd,
                               const float* input_xyz, int64_t i_sb, int64_t i_sp, int64_t i_sd,
                               const float* input_feat, int64_t f_sb, int64_t f_sc, int64_t f_sp,
                               float* out, int64_t o_sb, int64_t o_sc, int64_t o_sp, void* stream);

/*
 * backwarp_3d (models/utils.py:149-159): xyz2_warp = xyz2 + interp_k(xyz1 + flow12, -flow12)(xyz2).
 * xyz1, flow12 [B,3,m]; xyz2, xyz2_warp [B,3,n]; contiguous channel-first.
 */
int camli_backwarp_3d(int B, int n, int m, int k, const float* xyz1, const float* flow12,
                      const float* xyz2, float* xyz2_warp, void* stream);

/*
 * PointConvDW WeightNet (models/point_conv.py:122-127): weights_out[b,s,j,:] =
 * relu(W3 relu(W2 relu(W1 (xyz[idx[b,s,j]] - centre[b,s]) + b1) + b2) + b3), rows [B,S,k,O].
 * xyz / centre_xyz: [B, points, 3] views (element strides batch, point, dim);
 * knn_idx [B,S,K] i64 of which the first k columns are used; W1 [8,3], W2 [32,8], W3 [O,32].
 */
int camli_pointconv_dw_weights(int B, int N, int S, int K, int k, int O,
                               const float* xyz, int64_t x_sb, int64_t x_sp, int64_t x_sd,
                               const float* centre_xyz, int64_t c_sb, int64_t c_sp, int64_t c_sd,
                               const int64_t* knn_idx,
                               const float* W1, const float* b1, const float* W2, const float* b2,
                               const float* W3, const float* b3, float* weights_out, void* stream);

/*
 * PointConvDW aggregation (models/point_conv.py:126-128):
 * out_rows[b,s,o] = max_{j<k} feat_rows[b, idx[b,s,j], o] * weights[b,s,j,o].  k <= 32.
 */
/* Layers 1-2 only of the same WeightNet (3 -> 8 -> 32, ReLU) as rows hidden_out [B,S,k,32]; the 32 -> O output
 * layer is then one camli_conv_gemm (K = 32, ReLU epilogue) over the B*S*k rows, which yields the [B,S,k,O]
 * weights of camli_pointconv_dw_weights on the tensor cores. */
int camli_pointconv_dw_hidden(int B, int N, int S, int K, int k,
                              const float* xyz, int64_t x_sb, int64_t x_sp, int64_t x_sd,
                              const float* centre_xyz, int64_t c_sb, int64_t c_sp, int64_t c_sd,
                              const int64_t* knn_idx, const float* W1, const float* b1, const float* W2,
                              const float* b2, float* hidden_out, void* stream);

int camli_pointconv_dw_gather_max(int B, int N, int S, int K, int k, int O,
                                  const float* feat_rows, int64_t ld_feat, const float* weights,
                                  const int64_t* knn_idx, float* out_rows, int64_t ld_out, void* stream);

/*
 * Point-correlation lookup, all levels in one launch (Correlation3D.forward up to the `merge`
 * convolution, models/camliraft_l_core.py:62-98): per level l, the 16 nearest points of
 * xyz2_levels[l] around every xyz1 point, MLP 4->32->32 (ReLU) of (offset, volume entry),
 * summed over the neighbours, written to out_rows[b, q, 32*l : 32*l+32].
 * xyz1 [B,3,n1] contiguous; xyz2_levels_host[l]: device pointer of a [B, n2, 3] view with
 * element strides xyz2_strides_host[3l..3l+2] = (batch, point, dim); volumes_host[l] [B,n1,n2_host[l]].
 */
int camli_corr3d_lookup(int B, int n1, int n_levels, const float* xyz1,
                        const float* const* xyz2_levels_host, const int64_t* xyz2_strides_host,
                        const int* n2_host, const float* const* volumes_host,
                        const float* W1, const float* b1, const float* W2, const float* b2,
                        float* out_rows, int ld_out, void* stream);

/*
 * One pooling step of the point-correlation pyramid (models/camliraft_l_core.py:56-60):
 * vol_out[b,p,q] = mean_{j<k} vol_in[b,p,knn_idx[b,q,j]];  vol_in [B,n1,n_in], vol_out [B,n1,n_out].
 */
int camli_corr3d_pool(int B, int n1, int n_in, int n_out, int k, const float* vol_in,
                      const int64_t* knn_idx, float* vol_out, void* stream);

/*
 * CLFM 3D->2D gather + ScoreNet (FusionAwareInterp.forward before out_conv, models/clfm.py:57-75):
 * out_rows[b,p,:] = sigmoid(W2 leaky(W1 [du,dv,|d|] + b1) + b2) * feat3d_rows[b, nn_idx[b,p], :]
 * with (du,dv) = uv[b,:,nn] - pixel(p).  uv [B,2,N]; nn_idx [B,H*W] i64; W1 [16,3]; W2 [C,16];
 * out_rows NHWC [B,H*W,C].
 */
int camli_clfm_interp(int B, int H, int W, int N, int C, const float* uv, const int64_t* nn_idx,
                      const float* feat3d_rows, int64_t ld_feat,
                      const float* W1, const float* b1, const float* W2, const float* b2,
                      float* out_rows, void* stream);

/*
 * Bilinear sample of an NHWC map at pixel coordinates uv [B,2,N] (align_corners, zeros
 * outside): grid_sample_wrapper, models/utils.py:262-269.  out_rows [B,N,ld_out].
 */
int camli_bilinear_sample_rows(int B, int H, int W, int N, int C, const float* feat_nhwc, const float* uv,
                               float* out_rows, int64_t ld_out, void* stream);

/*
 * Backward of camli_pointconv_group (the reference differentiates models/point_conv.py:56-66 through autograd).
 * grad_out [B,S,16*C] -> grad_rows [B,N,C] (ZERO-INITIALISED by the caller: scatter-add over the neighbour tables),
 * grad_centre [B,3,S] contiguous (every element written) and grad_params[176] = dW1[8,3] | db1[8] | dW2[16,8] |
 * db2[16] of the WeightNet (ZERO-INITIALISED by the caller).  Other arguments as in the forward.
 */
int camli_pointconv_group_backward(int B, int N, int S, int K, int k, int C,
                                   const float* rows, int64_t ld_rows,
                                   const float* centre_xyz, int64_t c_sb, int64_t c_sp, int64_t c_sd,
                                   const int64_t* knn_idx, const float* W1, const float* b1, const float* W2,
                                   const float* b2, float negative_slope, const float* grad_out,
                                   float* grad_rows, float* grad_centre, float* grad_params, void* stream);

/*
 * Convex up-sampling of a coarse flow field (convex_upsample, models/utils.py:191-204; callers
 * models/raft_core.py:184-197 with factor 8 and `0.25 * mask`, models/pwc_core.py:218-224 with factor 4):
 *   up[b,c,S*h+i,S*w+j] = sum_k softmax_k(scale * mask[b,h,w, k*S*S + i*S + j]) * S * flow[b,c,h+dy_k,w+dx_k]
 * k = 3*(dy+1)+(dx+1), zeros outside the map.  flow [B,2,H,W] contiguous, mask_rows NHWC [B,H,W,9*S*S],
 * up [B,2,S*H,S*W] contiguous; factor S = 4 or 8.
 */
int camli_convex_upsample(int B, int H, int W, int factor, const float* flow, const float* mask_rows, float scale,
                          float* up, void* stream);

/*
 * Its backward: grad_up [B,2,S*H,S*W] -> grad_mask_rows [B,H,W,9*S*S] (every element written) and grad_flow
 * [B,2,H,W] (every element written; gathered from tap_scratch, B*H*W*18 floats of workspace).
 */
int camli_convex_upsample_backward(int B, int H, int W, int factor, const float* flow, const float* mask_rows, float scale,
                                   const float* grad_up, float* grad_mask_rows, float* tap_scratch, float* grad_flow,
                                   void* stream);

/*
 * PointConv grouping stage (models/point_conv.py:56-66): rows [B,N,ld_rows] hold [xyz | features]
 * channel-last (xyz in columns 0..2, C columns used); out[b,s,w*C + c] =
 * sum_{j<k} WeightNet(xyz[idx[b,s,j]] - centre[b,s])[w] * rows[b, idx[b,s,j], c], w < 16, with
 * WeightNet = act(W2 act(W1 d + b1) + b2), act = leaky(negative_slope) (0 => ReLU, 1 => identity).
 * centre_xyz: [B,S,3] view with element strides (batch, point, dim).  k <= 32, C <= 256.
 */
int camli_pointconv_group(int B, int N, int S, int K, int k, int C,
                          const float* rows, int64_t ld_rows,
                          const float* centre_xyz, int64_t c_sb, int64_t c_sp, int64_t c_sd,
                          const int64_t* knn_idx, const float* W1, const float* b1, const float* W2,
                          const float* b2, float negative_slope, float* out, void* stream);

/*
 * All-pairs feature inner product on the tcgen05 tensor cores (TMA-fed, TMEM accumulators),
 * fp32-accurate through a 3xTF32 operand split:
 *     out[b,m,n] = scale * sum_k a_rows[b,m,k] * b_rows[b,n,k]
 * Replaces torch.matmul(fmap1^T, fmap2) / sqrt(C) (models/raft_core.py:56-63) and
 * torch.bmm(feat1^T, feat2) / C (models/camliraft_l_core.py:52-53).
 * a_rows [B*M, K], b_rows [B*N, K] row-major f32 (16-byte aligned); K % 32 == 0;
 * workspace: camli_allpairs_workspace_floats(B,M,N,K) floats of scratch (the hi/lo operand
 * copies); out [B,M,N] f32.
 */
int64_t camli_allpairs_workspace_floats(int B, int M, int N, int K);
int camli_allpairs_correlation(const float* a_rows, const float* b_rows, float* workspace, float* out,
                               int B, int M, int N, int K, float scale, void* stream);

/*
 * Training side of camli_conv_gemm (in the reference: the cuDNN wgrad / cuBLAS GEMMs autograd picks for every dense layer,
 * models/raft_core.py:110-197, models/mlp.py:41-128, train.py:143-171).
 *
 * camli_transpose_split: rows [P, C] (pitch ld) -> hi_t, lo_t [C, P], the tf32 hi / lo parts of the TRANSPOSED tensor
 * (pixel-contiguous = K-major for the weight-gradient GEMM).  With y_rows != NULL the rows are dL/d(layer output): they are
 * multiplied by the derivative of activation `act` (CAMLI_ACT_NONE / RELU / LEAKY / TANH / SIGMOID) evaluated at the layer
 * OUTPUT y_rows [P, >= C] (pitch ldy) first; g_rows [P, C] (optional) receives that product row-major, colsum [C] (optional,
 * zeroed by the caller) its column sums = the bias gradient.  n_shift (odd) > 1: hi_t / lo_t are [n_shift, C, P], copy j
 * holding the rows shifted by (j - n_shift/2) * shift_step pixels inside their image row of W pixels (zero outside) -- the
 * x operand of a kw-wide window, because a TMA box cannot start at a 4-byte pixel offset of the contiguous dimension.
 * xstride = 2 (stride-2 convolution): the copies keep every second pixel of a row, [n_shift, C, P / W * ceil(W/2)].
 *
 * camli_conv_wgrad: dw[n, tap*Cin + c] = sum over pixels of g[p, n] * x[p (+) tap, c] for the stride-1 "same" convolution /
 * linear layer of camli_conv_gemm (OHWI layout, [Cout, kh*kw*Cin]); operands are the [C, B, H, W] hi / lo tensors of
 * camli_transpose_split, x with n_shift = kw, shift_step = dilation.  H, W are the OUTPUT grid; a stride-2 layer (stride = 2,
 * Hin = input rows, x prepared with xstride = 2) reads input row y * 2 + dy for output row y.  3xTF32 on tcgen05, K split over the SMs, partial tiles added with 128-bit atomics (dw is zeroed
 * inside).  W % 4 == 0, Cin % 4 == 0, odd windows.  passes = 3: 3xTF32; passes = 1: the hi parts only (one tf32 product, the
 * reduced-precision mode of the bf16-autocast training step; the lo tensors may then be NULL, also in camli_transpose_split).  The data gradient needs no kernel of its own: it is camli_conv_gemm of
 * g_rows with the spatially flipped, in/out-transposed weights.
 */
int camli_transpose_split(const float* rows, int64_t ld, int64_t P, int C, const float* y_rows, int64_t ldy,
                          int act, float slope, int W, int n_shift, int shift_step, int xstride,
                          float* hi_t, float* lo_t, float* g_rows, float* colsum, void* stream);
int camli_conv_wgrad(const float* g_hi_t, const float* g_lo_t, const float* x_hi_t, const float* x_lo_t,
                     int B, int H, int W, int Cout, int Cin, int kh, int kw, int dilation, int stride, int Hin, int passes,
                     float* dw, void* stream);

/*
 * SK fusion tail (SKFusion.forward after the align layers, models/clfm.py:199-214):
 *   a = leaky(a_rows), b = leaky(b_rows)   (negative_slope; 1 => inputs already activated)
 *   w = softmax_pair(sigmoid(W_out relu(W_mid mean_p(a + b))))      [B,C,2]
 *   out_rows = a * w[..,0] + b * w[..,1]
 * a_rows, b_rows, out_rows [B,P,C] rows; w_mid [C_mid,C]; w_out [2C,C_mid];
 * partial_scratch [B,32,C] and weights_scratch [B, 2*C + C_mid] f32 scratch (the [B,C,2] blend weights, then the
 * hidden layer).  Deterministic (no atomics).
 */
int camli_sk_fusion_tail(int B, int P, int C, int C_mid, const float* a_rows, const float* b_rows,
                         float negative_slope, const float* w_mid, const float* w_out,
                         float* partial_scratch, float* weights_scratch, float* out_rows, void* stream);

/* The same with output rows of pitch ld_out (>= C) floats: the blend lands in a channel slice of a wider buffer
 * (the ConvGRU's [h | x | r*h] state), batch b's rows starting at out_rows + b * P * ld_out. */
int camli_sk_fusion_tail_strided(int B, int P, int C, int C_mid, const float* a_rows, const float* b_rows,
                                 float negative_slope, const float* w_mid, const float* w_out,
                                 float* partial_scratch, float* weights_scratch, float* out_rows, int64_t ld_out,
                                 void* stream);

/*
 * ConvGRU gate stage (models/raft_core.py:125-128 / :132-135): zr [rows,2H] = pre-activation output
 * of the merged z|r convolution, h [rows,H], x [rows,X] (all channel-last rows):
 *   z = sigmoid(zr[:, :H]);  rhx = [ sigmoid(zr[:, H:]) * h | x ]   ([rows,H], [rows,H+X]).
 */
int camli_gru_gate(int64_t rows, int H, int X, const float* zr, const float* h, const float* x, float* z,
                   float* rhx, void* stream);

/*
 * ConvGRU state update (models/raft_core.py:129,136-138): h_out = (1-z)*h + z*tanh(q) over n
 * elements; fix_nonfinite != 0 additionally applies torch.nan_to_num.
 */
int camli_gru_update(int64_t n, const float* z, const float* h, const float* q, int fix_nonfinite,
                     float* h_out, void* stream);

/* Epilogue activations of camli_conv_gemm. */
#define CAMLI_ACT_NONE    0
#define CAMLI_ACT_RELU    1
#define CAMLI_ACT_LEAKY   2   /* v > 0 ? v : v * slope */
#define CAMLI_ACT_TANH    3
#define CAMLI_ACT_SIGMOID 4
/* fused ConvGRU epilogues of camli_conv_gemm_fused (models/raft_core.py:125-138) */
#define CAMLI_ACT_GRU_GATE       5   /* sigmoid(v); columns >= split are multiplied by aux1[p, n - split] (r * h)      */
#define CAMLI_ACT_GRU_UPDATE     6   /* (1 - z) * h + z * tanh(v) with z = aux1[p, n], h = aux2[p, n]                */
#define CAMLI_ACT_GRU_UPDATE_FIX 7   /* ... followed by torch.nan_to_num                                              */
/* OR-ed onto NONE / RELU / LEAKY / TANH / SIGMOID: torch.nan_to_num of the activated value (nan -> 0, +-inf -> +-FLT_MAX),
 * the guard the reference puts behind its motion encoder and flow heads (models/raft_core.py:164,180) */
#define CAMLI_ACT_FIX_NONFINITE  16
#define CAMLI_CONV_SINGLE_PASS   0x100   /* flag in the tile_n argument of camli_conv_gemm* (see there) */
#define CAMLI_WGRAD_BF16          16      /* `passes` value of camli_conv_wgrad: bf16 operands (written by camli_transpose_split with  */
#define CAMLI_TRANSPOSE_BF16      0x100   /* xstride | CAMLI_TRANSPOSE_BF16), one kind::f16 product, fp32 accumulation; W % 8 == 0      */
#define CAMLI_WGRAD_ACCUMULATE    0x100   /* flag in `passes` of camli_conv_wgrad: dw is NOT zeroed, the result is added to it (gradient   */
                                          /* accumulation into the parameter's own buffer: no zero-fill, no separate add kernel)            */

/* x -> (hi, lo) with hi = tf32(x) (round to nearest), lo = tf32(x - hi): the operand split of the 3xTF32
 * tensor-core kernels; used once per weight tensor. */
int camli_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);

/*
 * Linear layer / stride-1 "same" convolution on channel-last activations as one TMA + tcgen05 implicit
 * GEMM with fp32-level accuracy (3xTF32) and a fused epilogue.  Replaces the cuBLAS/cuDNN call + bias +
 * norm(folded) + activation kernels behind nn.Linear / MLP1d / Conv1d(1) / Conv2d of the hot path
 * (models/mlp.py:41-128, models/point_conv.py:29,62,106, models/raft_core.py:110-197):
 *   out[p, n] = act( sum_{ky,kx,c} x[b, y+ky-kh/2, x+kx-kw/2, c] * w[n, ky, kx, c] + bias[n] + residual[p, n] )
 * x        [B,H,W,ldx] f32 (first Cin channels of every pixel are read; zero padding outside the map);
 *          a linear layer over R rows is B=1, H=1, W=R.  Cin % 4 == 0, ldx % 4 == 0, 16-byte aligned.
 * w_hi/lo  [Cout, kh*kw*Cin] f32 = camli_split_tf32 of the OHWI (channels_last) weight; kh, kw odd.
 * bias     [Cout] or NULL; residual [B*H*W, ldr] or NULL; act = CAMLI_ACT_*.
 * out      [B*H*W, ldo] f32: ldo >= Cout lets the layer write a channel slice of a wider tensor.
 * tile_n   0 = automatic; else 32 / 64 / 96 / 128 (accumulator tile width; automatic = the widest that Cout fills, 96 for Cout % 96 == 0).  | CAMLI_CONV_SINGLE_PASS: ONE tf32 product per
 *          element instead of three (operands cut to 10-bit mantissas, fp32 accumulation) -- the reduced-precision mode the
 *          training step uses under bf16 autocast (a tf32 operand is more accurate than a bf16 one); parity paths never set it.
 */
int camli_conv_gemm(const float* x, int B, int H, int W, int Cin, int64_t ldx,
                    const float* w_hi, const float* w_lo, int Cout, int kh, int kw,
                    const float* bias, const float* residual, int64_t ldr,
                    int act, float slope, float* out, int64_t ldo, int tile_n, void* stream);

/*
 * camli_conv_gemm with the ConvGRU gate / update arithmetic in its epilogue, so a ConvGRU half is two launches
 * (z|r convolution -> z and r*h; q convolution -> h') with no gate / update / concatenation kernels:
 * aux1 [B*H*W, ld1], aux2 [B*H*W, ld2] per-pixel side inputs (see CAMLI_ACT_GRU_*); with act = GRU_GATE and out2 != NULL
 * the columns n >= split are written to out2[p, n - split] (row stride ldo2) instead of out.  The bias and the residual
 * are added before the gate arithmetic.  `out` may alias aux2 (h' over h): each element is read and written by the
 * same thread.  Other arguments as camli_conv_gemm.
 */
int camli_conv_gemm_fused(const float* x, int B, int H, int W, int Cin, int64_t ldx,
                          const float* w_hi, const float* w_lo, int Cout, int kh, int kw,
                          const float* bias, const float* residual, int64_t ldr,
                          int act, float slope, float* out, int64_t ldo,
                          const float* aux1, int64_t ld1, const float* aux2, int64_t ld2, int split,
                          float* out2, int64_t ldo2, int tile_n, void* stream);

/* The same kernel for a STRIDED and / or DILATED convolution (stride 1 or 2, padding dilation * (k/2)): H_in x W_in is
 * the input grid, the output grid is ((H_in - 1) / stride + 1) x ((W_in - 1) / stride + 1); residual / out / aux are
 * indexed by output pixel.  The activation operand is fetched through a TMA tensor map whose pixel dimensions have
 * traversal stride `stride`; a dilated tap is just a larger coordinate offset of the box (reference: the stride-2
 * convolutions of the ResNet stage `layer2.0`, models/raft_core.py:10-22, and of PWC's feature pyramid,
 * models/pwc_core.py:9-29; the dilated context network, models/pwc_core.py:128-141). */
int camli_conv_gemm_strided(const float* x, int B, int H_in, int W_in, int Cin, int64_t ldx,
                            const float* w_hi, const float* w_lo, int Cout, int kh, int kw, int stride, int dilation,
                            const float* bias, const float* residual, int64_t ldr,
                            int act, float slope, float* out, int64_t ldo,
                            const float* aux1, int64_t ld1, const float* aux2, int64_t ld2, int split,
                            float* out2, int64_t ldo2, int tile_n, void* stream);

/*
 * The same operation for a handful of output channels (Cout <= 4: the last layer of the flow heads,
 * models/raft_core.py:176, models/camliraft_l_core.py:110) on the CUDA cores, plain fp32 FMA: one warp per
 * pixel.  w [Cout, kh*kw*Cin] f32 in OHWI order (not split); other arguments as camli_conv_gemm.
 */
int camli_conv_small_n(const float* x, int B, int H, int W, int Cin, int64_t ldx, const float* w, int Cout,
                       int kh, int kw, const float* bias, int act, float slope, float* out, int64_t ldo,
                       void* stream);

/*
 * Backward of camli_corr2d_lookup with respect to the volume pyramid (the coordinates carry no gradient: the
 * models look up at a detached flow, models/camliraft_core.py:105).  grad_volumes[l]: [B,H*W,h_l,w_l] f32,
 * ZERO-INITIALISED by the caller; grad_out_rows: channel-last [B,H*W,n_levels*(2r+1)^2].  Each (pixel, level)
 * footprint is written by one CTA: no atomics, deterministic.
 */
int camli_corr2d_lookup_backward(float* const* grad_volumes, const int* level_h, const int* level_w, int n_levels,
                                 const float* coords, const float* grad_out_rows, int B, int H, int W,
                                 int radius, void* stream);

/*
 * Backward of camli_pointconv_dw_gather_max: the arg-max neighbour j* of every (centroid s, channel o) is found
 * again from the saved inputs (first maximum, as torch.max); grad_feat_rows[b, idx[b,s,j*], o] += g * w (atomic),
 * grad_weights[b,s,j*,o] = g * feat.  Both gradients ZERO-INITIALISED by the caller; grad_weights may be NULL.
 */
int camli_pointconv_dw_gather_max_backward(int B, int N, int S, int K, int k, int O, const float* feat_rows,
                                           const float* weights, const int64_t* knn_idx, const float* grad_out_rows,
                                           float* grad_feat_rows, float* grad_weights, void* stream);

/*
 * The same operation for a handful of INPUT channels (Cin <= 4, windows up to 7x7: the first layer of the motion
 * encoder's flow path, models/raft_core.py:154) on the CUDA cores, plain fp32 FMA.  x may have any pixel stride
 * ldx >= Cin (no alignment requirement); w [Cout, kh*kw*Cin] f32 in OHWI order.
 */
int camli_conv_small_cin(const float* x, int B, int H, int W, int Cin, int64_t ldx, const float* w, int Cout,
                         int kh, int kw, const float* bias, int act, float slope, float* out, int64_t ldo,
                         void* stream);

/*
 * ResNet stem in one kernel: 7x7 stride-2 padding-3 convolution 3 -> 64 (eval-mode BatchNorm folded into w / bias) + ReLU
 * + 3x3 stride-2 padding-1 max-pool (reference: conv1 / bn1 / relu / maxpool of the mmdet ResNet-50 behind
 * models/raft_core.py:10-22,35-38).  x [B,H,W,ldx>=3] channel-last, w_ohwi [64,7,7,3], bias [64],
 * out [B,Hp,Wp,ldo>=64] with Hp = ((H-1)/2)/2 + 1 rounded like torch (H = 544 -> 136).  fp32 FMA on the CUDA cores.
 */
int camli_stem_conv_pool(const float* x, int B, int H, int W, int64_t ldx, const float* w_ohwi, const float* bias,
                         float* out, int64_t ldo, void* stream);

/*
 * Backward kernels of the point-branch / fusion operators (camliflow_b200/csrc/backward_point.cu).  The reference
 * differentiates these stages through torch autograd (models/utils.py:130-146, models/camliraft_l_core.py:56-98,
 * models/clfm.py:57-75, models/raft_core.py:65-68); here each gradient is one launch.  Every grad_* output must be
 * zero-initialised by the caller unless noted; coordinates receive no gradient.
 *
 * camli_three_nn_interpolate_backward : grad_feat[b,f,idx_j(q)] += w_j(q) * grad_out[b,f,q]   (same search + weights as
 *                                       camli_three_nn_interpolate; strides as there)
 * camli_corr3d_lookup_backward        : grad_volumes[l][b,q,idx_j] = d cost entry (plain stores), and the gradients of the
 *                                       4 -> 32 -> 32 cost MLP (grad_W1 [32,4], grad_b1 [32], grad_W2 [32,32], grad_b2 [32])
 * camli_corr3d_pool_backward          : grad_in[b,p,idx[b,q,j]] += grad_out[b,p,q] / k
 * camli_corr2d_pool_backward          : grad_vol0 (IN/OUT, holds level 0's own gradient) += sum_l up(grad_coarser[l-1]) / 4^l
 * camli_clfm_interp_backward          : gradients of the ScoreNet parameters (grad_W1 [16,3], grad_b1 [16], grad_W2 [C,16],
 *                                       grad_b2 [C]); grad_out_rows is [B,H*W,C] channel-last
 */
int camli_three_nn_interpolate_backward(int B, int n, int m, int k, int F,
                                        const float* query_xyz, int64_t q_sb, int64_t q_sp, int64_t q_sd,
                                        const float* input_xyz, int64_t i_sb, int64_t i_sp, int64_t i_sd,
                                        const float* grad_out, int64_t g_sb, int64_t g_sc, int64_t g_sp,
                                        float* grad_feat, int64_t f_sb, int64_t f_sc, int64_t f_sp, void* stream);
int camli_corr3d_lookup_backward(int B, int n1, int n_levels, const float* xyz1,
                                 const float* const* xyz2_levels_host, const int64_t* xyz2_strides_host,
                                 const int* n2_host, const float* const* volumes_host, float* const* grad_volumes_host,
                                 const float* W1, const float* b1, const float* W2, const float* b2,
                                 const float* grad_out_rows, int ld_grad,
                                 float* grad_W1, float* grad_b1, float* grad_W2, float* grad_b2, void* stream);
int camli_corr3d_pool_backward(int B, int n1, int n_in, int n_out, int k, const float* grad_out,
                               const int64_t* knn_idx, float* grad_in, void* stream);
int camli_corr2d_pool_backward(float* grad_vol0, const float* const* grad_coarser_host, int n_levels, int64_t rows,
                               int h0, int w0, void* stream);
int camli_clfm_interp_backward(int B, int H, int W, int N, int C, const float* uv, const int64_t* nn_idx,
                               const float* feat3d_rows, int64_t ld_feat,
                               const float* W1, const float* b1, const float* W2, const float* b2,
                               const float* grad_out_rows,
                               float* grad_W1, float* grad_b1, float* grad_W2, float* grad_b2, void* stream);

/* Tuning switch (no reference counterpart): 1 tags the footprint loads of camli_corr2d_lookup L2 evict-last (meant to keep the
 * ~10 % of the volume pyramid the refinement revisits in the L2 across iterations; measured: no gain, default 0).  Returns the
 * previous setting. */
int camli_corr2d_lookup_set_l2_keep(int enabled);

/* Tuning switch (no reference counterpart): launch the implicit-GEMM kernel with programmatic dependent launch
 * (default 0; 1: its barrier / TMEM / descriptor setup overlaps the tail of the previous kernel of the stream; it
 * still waits for that kernel's completion before touching global memory).  Returns the previous setting. */
int camli_conv_gemm_set_pdl(int enabled);

/* Diagnostics for camli_conv_gemm: a device buffer of >= 128 int64 that CTA 0 of every following launch stamps
 * with SM-clock values of its pipeline events (scripts/conv_gemm_timeline.py); NULL detaches (default). */
int camli_conv_gemm_set_timeline(long long* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* CAMLI_B200_H */
