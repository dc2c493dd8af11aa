"""Fused hot-path operators (host side).  Every function here is the doorway to one kernel
(family) of libcamli_b200.so (include/camli_b200.h).  Two layouts are used:

  * "cf"   -- the reference's channel-first logical layout ([B,C,N] point features, [B,3,N]
              coordinates, [B,C,H,W] maps) at the module boundaries;
  * "rows" -- channel-last storage ([B,N,C]; NHWC for maps) inside the fused paths, so that a
              neighbour's feature vector is one contiguous, coalesced read.

There is no CPU path and no eager-PyTorch fallback: CUDA tensors and a built library are
required.  Every operator is differentiable (camliflow_b200/grad.py): the forward is always the fused
kernel, the backward a hand-written kernel or a recompute of the operator's formula.
"""
import ctypes
import os

import torch
import torch.nn.functional as F

from . import grad, native
from .csrc import k_nearest_neighbor
from .native import i32, i64, ptr, stream


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("camliflow_b200.ops: CUDA tensors required (there is no CPU fallback)")
        if t.is_floating_point() and t.dtype != torch.float32:
            raise RuntimeError("camliflow_b200.ops: float32 tensors required, got %s" % t.dtype)


def _no_grad(name, *tensors):
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        raise NotImplementedError("camliflow_b200.ops.%s has no backward yet: run under torch.no_grad()" % name)


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _int_array(values, ctype=ctypes.c_int):
    return (ctype * len(values))(*values)


def rows_of(x_cf):
    """[B,C,N] (any strides) -> contiguous [B,N,C]; free when x_cf is already a transposed view of rows."""
    t = x_cf.transpose(1, 2)
    return t if t.is_contiguous() else t.contiguous()


def cf_of(rows):
    """[B,N,C] rows -> logical [B,C,N] view."""
    return rows.transpose(1, 2)


def nhwc_rows(x):
    """[B,C,H,W] (any strides) -> contiguous NHWC storage viewed as [B,H,W,C]."""
    return x.permute(0, 2, 3, 1).contiguous()


def nchw_view(rows_bhwc):
    """NHWC storage [B,H,W,C] -> logical [B,C,H,W] (channels_last strides)."""
    return rows_bhwc.permute(0, 3, 1, 2)


# ---------------------------------------------------------------- grouping / interpolation
def gather_points(data, idx):
    """data [B,C,N], idx [B,...] (i64) -> [B,C,...]: models/utils.py:62-80 (one-off uses only)."""
    _need_cuda(data, idx)
    B, C = data.shape[:2]
    flat = idx.reshape(B, 1, -1).expand(B, C, -1)
    return torch.gather(data, 2, flat).view([B, C] + list(idx.shape[1:]))


def knn_interpolate(input_xyz, input_feat, query_xyz, k=3):
    """Three-NN inverse-distance interpolation fused with the search (models/utils.py:130-146).
    input_xyz [B,3,m], input_feat [B,F,m], query_xyz [B,3,n] -> [B,F,n]."""
    input_xyz, input_feat, query_xyz = grad.f32(input_xyz, input_feat, query_xyz)
    _need_cuda(input_xyz, input_feat, query_xyz)
    if grad.needs_grad(input_feat) and not grad.needs_grad(input_xyz, query_xyz):
        with torch.autocast("cuda", enabled=False):          # weights are constants of the geometry: scatter kernel
            return _ThreeNN.apply(input_xyz, input_feat, query_xyz, k)
    return grad.recompute(_knn_interpolate, grad.f_knn_interpolate, input_xyz, input_feat, query_xyz, k)


class _ThreeNN(torch.autograd.Function):
    """knn_interpolate with a gradient for the features only: d feat[idx_j] += w_j * g (camli_three_nn_interpolate_backward
    repeats the search and the inverse-distance weights)."""

    @staticmethod
    def forward(ctx, input_xyz, input_feat, query_xyz, k):
        ctx.save_for_backward(input_xyz, query_xyz)
        ctx.k, ctx.feat_shape = k, input_feat.shape
        return _knn_interpolate(input_xyz, input_feat, query_xyz, k)

    @staticmethod
    def backward(ctx, g):
        input_xyz, query_xyz = ctx.saved_tensors
        B, Fc, m = ctx.feat_shape
        n = query_xyz.shape[-1]
        g = g.contiguous()
        gfeat = torch.zeros(ctx.feat_shape, dtype=torch.float32, device=g.device)
        qs, xs, gs, fs = query_xyz.stride(), input_xyz.stride(), g.stride(), gfeat.stride()
        with torch.cuda.device(g.device):
            native.call("camli_three_nn_interpolate_backward", i32(B), i32(n), i32(m), i32(ctx.k), i32(Fc),
                        ptr(query_xyz), i64(qs[0]), i64(qs[2]), i64(qs[1]),
                        ptr(input_xyz), i64(xs[0]), i64(xs[2]), i64(xs[1]),
                        ptr(g), i64(gs[0]), i64(gs[1]), i64(gs[2]),
                        ptr(gfeat), i64(fs[0]), i64(fs[1]), i64(fs[2]), stream(),
                        algo_bytes=B * ((n + m) * 12 + n * ctx.k * (Fc * 8 + 12) + n * Fc * 4), flops=B * n * m * 8)
        return None, gfeat, None, None


def _knn_interpolate(input_xyz, input_feat, query_xyz, k):
    B, Fc, m = input_feat.shape
    n = query_xyz.shape[-1]
    out = torch.empty((B, Fc, n), dtype=torch.float32, device=input_feat.device)
    qs, xs, fs, os_ = query_xyz.stride(), input_xyz.stride(), input_feat.stride(), out.stride()
    with torch.cuda.device(out.device):
        native.call("camli_three_nn_interpolate", i32(B), i32(n), i32(m), i32(k), i32(Fc),
                    ptr(query_xyz), i64(qs[0]), i64(qs[2]), i64(qs[1]),
                    ptr(input_xyz), i64(xs[0]), i64(xs[2]), i64(xs[1]),
                    ptr(input_feat), i64(fs[0]), i64(fs[1]), i64(fs[2]),
                    ptr(out), i64(os_[0]), i64(os_[1]), i64(os_[2]), stream(),
                    algo_bytes=B * ((n + m) * 12 + n * k * (Fc * 4 + 12) + n * Fc * 4), flops=B * n * m * 8)
    return out


def backwarp_3d(xyz1, xyz2, flow12, k=3):
    """xyz2 + interp(xyz1 + flow12, -flow12)(xyz2) in one kernel (models/utils.py:149-159)."""
    _need_cuda(xyz1, xyz2, flow12)
    if grad.needs_grad(xyz1, xyz2, flow12):       # (the models always warp by a detached flow; kept for completeness)
        return xyz2 + knn_interpolate(xyz1 + flow12, -flow12, xyz2, k)
    xyz1, xyz2, flow12 = xyz1.contiguous(), xyz2.contiguous(), flow12.contiguous()
    B, _, m = xyz1.shape
    n = xyz2.shape[-1]
    out = torch.empty_like(xyz2)
    with torch.cuda.device(out.device):
        native.call("camli_backwarp_3d", i32(B), i32(n), i32(m), i32(k), ptr(xyz1), ptr(flow12), ptr(xyz2), ptr(out),
                    stream(), algo_bytes=B * (m * 24 + n * 24), flops=B * n * m * 11)
    return out


# ---------------------------------------------------------------- image-side sampling
def bilinear_sample_rows(feat2d, uv):
    """feat2d [B,C,H,W] (channels_last preferred), uv [B,2,N] pixel coords -> rows [B,N,C]
    (align_corners, zero padding; models/utils.py:262-269)."""
    feat2d, uv = grad.f32(feat2d, uv)
    _need_cuda(feat2d, uv)
    return grad.recompute(_bilinear_sample_rows, grad.f_bilinear_sample_rows, feat2d, uv)


def _bilinear_sample_rows(feat2d, uv):
    B, C, H, W = feat2d.shape
    N = uv.shape[-1]
    src = nhwc_rows(feat2d)
    uv = uv.contiguous()
    out = torch.empty((B, N, C), dtype=torch.float32, device=feat2d.device)
    with torch.cuda.device(out.device):
        native.call("camli_bilinear_sample_rows", i32(B), i32(H), i32(W), i32(N), i32(C), ptr(src), ptr(uv), ptr(out),
                    i64(C), stream(), algo_bytes=B * N * (5 * C * 4 + 8))
    return out


def bilinear_sample(feat2d, uv):
    """Channel-first result [B,C,N] of bilinear_sample_rows."""
    return cf_of(bilinear_sample_rows(feat2d, uv))


class _ConvexUpsample(torch.autograd.Function):
    """camli_convex_upsample / camli_convex_upsample_backward: softmax over the taps, the 3x3 unfold, the weighted sum
    and the pixel shuffle in one kernel each way (the backward recomputes the softmax; the flow gradient is a gather
    over per-pixel tap sums: deterministic)."""

    @staticmethod
    def forward(ctx, flow, mask, s, scale):
        B, _, H, W = flow.shape
        flow = flow.contiguous()
        mask_rows = nhwc_rows(mask)
        ctx.save_for_backward(flow, mask_rows)
        ctx.s, ctx.scale = s, scale
        up = torch.empty((B, 2, H * s, W * s), dtype=torch.float32, device=flow.device)
        with torch.cuda.device(flow.device):
            native.call("camli_convex_upsample", i32(B), i32(H), i32(W), i32(s), ptr(flow), ptr(mask_rows), ctypes.c_float(scale),
                        ptr(up), stream(), algo_bytes=B * H * W * 11 * s * s * 4)
        return up

    @staticmethod
    def backward(ctx, g):
        flow, mask_rows = ctx.saved_tensors
        B, _, H, W = flow.shape
        g = g.float().contiguous()
        g_mask_rows = torch.empty_like(mask_rows)
        taps = torch.empty((B, H, W, 18), dtype=torch.float32, device=flow.device)
        g_flow = torch.empty_like(flow)
        with torch.cuda.device(flow.device):
            native.call("camli_convex_upsample_backward", i32(B), i32(H), i32(W), i32(ctx.s), ptr(flow), ptr(mask_rows),
                        ctypes.c_float(ctx.scale), ptr(g), ptr(g_mask_rows), ptr(taps), ptr(g_flow), stream(),
                        algo_bytes=B * H * W * (20 * ctx.s * ctx.s + 18) * 4)
        return g_flow, nchw_view(g_mask_rows), None, None


def convex_upsample(flow, mask, s=8, scale=1.0):
    """models/utils.py:191-204: flow [B,2,H,W], mask [B,9*s*s,H,W] -> [B,2,s*H,s*W]; the softmax runs on scale * mask
    (RAFT's up-sampler passes 0.25, models/raft_core.py:197).  fp32 like the reference's explicit casts."""
    flow, mask = grad.f32(flow, mask)
    _need_cuda(flow, mask)
    B, two, H, W = flow.shape
    if two != 2 or tuple(mask.shape) != (B, 9 * s * s, H, W) or s not in (4, 8):
        raise RuntimeError("convex_upsample: flow [B,2,H,W] and mask [B,9*s*s,H,W] with s = 4 or 8 expected, got %s / %s / s = %d"
                           % (tuple(flow.shape), tuple(mask.shape), s))
    return _ConvexUpsample.apply(flow, mask, s, float(scale))


# ---------------------------------------------------------------- all-pairs inner products
# "tcgen05": the hand-written TMA + tcgen05 3xTF32 kernel (default); "cublas": torch.bmm (library fp32
# SGEMM), kept only as the A/B comparison point for benchmarks.
ALLPAIRS_IMPL = os.environ.get("CAMLI_ALLPAIRS", "tcgen05")


def allpairs(a_rows, b_rows, scale):
    """out[b,m,n] = scale * <a_rows[b,m,:], b_rows[b,n,:]>: a_rows [B,M,K], b_rows [B,N,K] -> [B,M,N]."""
    a_rows, b_rows = grad.f32(a_rows, b_rows)
    _need_cuda(a_rows, b_rows)
    if grad.needs_grad(a_rows, b_rows):
        with torch.autocast("cuda", enabled=False):
            return _AllPairs.apply(a_rows, b_rows, scale)
    return _allpairs(a_rows, b_rows, scale)


class _AllPairs(torch.autograd.Function):
    """d/da = scale * gO @ b, d/db = scale * gO^T @ a (the transposed products of the forward GEMM)."""

    @staticmethod
    def forward(ctx, a_rows, b_rows, scale):
        ctx.save_for_backward(a_rows, b_rows)
        ctx.scale = scale
        return _allpairs(a_rows, b_rows, scale)

    @staticmethod
    def backward(ctx, g):
        a_rows, b_rows = ctx.saved_tensors
        ga = torch.bmm(g, b_rows).mul_(ctx.scale) if ctx.needs_input_grad[0] else None
        gb = torch.bmm(g.transpose(1, 2), a_rows).mul_(ctx.scale) if ctx.needs_input_grad[1] else None
        return ga, gb, None


def _allpairs(a_rows, b_rows, scale):
    B, M, K = a_rows.shape
    N = b_rows.shape[1]
    if ALLPAIRS_IMPL == "cublas" or K % 32 != 0:
        return torch.bmm(a_rows * scale, b_rows.transpose(1, 2))
    a_rows, b_rows = a_rows.contiguous(), b_rows.contiguous()
    out = torch.empty((B, M, N), dtype=torch.float32, device=a_rows.device)
    lib = native.lib()
    lib.camli_allpairs_workspace_floats.restype = ctypes.c_int64
    ws = torch.empty((lib.camli_allpairs_workspace_floats(B, M, N, K),), dtype=torch.float32, device=a_rows.device)
    with torch.cuda.device(out.device):
        native.call("camli_allpairs_correlation", ptr(a_rows), ptr(b_rows), ptr(ws), ptr(out), i32(B), i32(M), i32(N),
                    i32(K), ctypes.c_float(scale), stream(),
                    algo_bytes=B * ((M + N) * K * 4 + M * N * 4), flops=2 * B * M * N * K)
    return out


# ---------------------------------------------------------------- tensor-core linear / convolution
ACT_CODES = {None: 0, "none": 0, "relu": 1, "leaky_relu": 2, "tanh": 3, "sigmoid": 4,
             "gru_gate": 5, "gru_update": 6, "gru_update_fix": 7,
             # + torch.nan_to_num of the activated value (CAMLI_ACT_FIX_NONFINITE)
             "none_fix": 16, "relu_fix": 17}
_TC_WEIGHTS = {}


def _publish_cached_weight():
    """A cached weight is prepared by kernels on whichever stream first needs the layer, and later launches on OTHER streams
    (the forked fusion sites share one CLFM module) find it in the cache without any ordering against those kernels: finish
    them before the entry becomes visible.  Once per weight; never inside a stream capture (engines warm up eagerly first)."""
    if torch.cuda.is_available() and not torch.cuda.is_current_stream_capturing():
        torch.cuda.current_stream().synchronize()


def tc_weight(key_params, builder):
    """(w_hi, w_lo, bias) of a layer for conv_gemm: `builder()` returns the effective (weight [N, taps*Cin] in
    OHWI order, bias [N] or None) -- e.g. with an eval BatchNorm folded in -- which is split into its tf32
    hi / lo parts ONCE and cached until one of `key_params` changes (data_ptr, version)."""
    import weakref
    live = [p for p in key_params if p is not None]
    key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in live)
    slot = _TC_WEIGHTS.get(key[0][0])
    if slot is None or slot[0] != key or slot[1]() is not live[0]:      # (a freed tensor's address may be reused)
        with torch.no_grad():
            w2d, bias = builder()
            w2d = w2d.float().contiguous()
            hi, lo = torch.empty_like(w2d), torch.empty_like(w2d)
            with torch.cuda.device(w2d.device):
                native.call("camli_split_tf32", ptr(w2d), ptr(hi), ptr(lo), i64(w2d.numel()), stream())
            slot = (key, weakref.ref(live[0]), hi, lo, None if bias is None else bias.float().contiguous())
            _publish_cached_weight()
        _TC_WEIGHTS[key[0][0]] = slot
    return slot[2], slot[3], slot[4]


def _pixel_layout(t):
    """(ld, ok) of a channel-last [B,H,W,C] view whose pixels are `ld` floats apart (channel slices of a wider
    NHWC buffer qualify); ok = the TMA / vector-store alignment rules hold."""
    B, H, W, C = t.shape
    sb, sh, sw, sc = t.stride()
    ld = sw if W > 1 else (sh if H > 1 else (sb if B > 1 else C))
    ok = (sc == 1 or C == 1) and (W == 1 or sw == ld) and (H == 1 or sh == W * ld) and (B == 1 or sb == H * W * ld)
    return ld, ok and ld % 4 == 0 and t.data_ptr() % 16 == 0


# "throughput" (default): the widest N tile that C_out fills -- fewest CTAs, most SMs left to the other stream / the other
# graphs in flight.  "latency": layers of <= 74 pixel tiles with 64 < C_out <= 128 and a long k-loop use 64-wide tiles (twice
# the CTAs, ~15-30 % shorter).  A property of the graph being captured: FlowEngine(tile_policy=...) sets it around its capture.
TILE_POLICY = os.environ.get("CAMLI_TILE_POLICY", "throughput")
# Inference-side precision switch of the dense kernel (NOT a parity mode): True = one tf32 product per element instead of
# three, what torch's default `cudnn.allow_tf32 = True` gives the reference's own convolutions on an Ampere+ GPU.  The parity
# tests, smoke() and the default bench line run with False.
SINGLE_PASS_INFERENCE = os.environ.get("CAMLI_CONV_PRECISION", "fp32") == "tf32"


# CTAs a layer may spread over under the "latency" policy: one wave of the 148 SMs
LATENCY_WAVE = int(os.environ.get("CAMLI_LATENCY_WAVE", "148"))


def _latency_tile(B, Ho, Wo, Cout):
    """Tile width of a layer when ONE graph runs at a time (TILE_POLICY "latency"): the narrowest of 32 / 64 / 128
    output columns that still keeps the layer within one wave of CTAs.  The library's default (the widest tile Cout
    fills) minimises SM-time, which is what counts with several graphs in flight; alone on the GPU a 2048-point linear
    layer would sit on 16 SMs and an 8160-pixel C_out = 128 convolution on 68, with a 64 KB tile store at the end of
    each CTA (scripts/cg_time.py, CG_TIME_SMALL: 2048 x 128->128: 9.1 / 7.0 / 6.6 us at 128 / 64 / 32 columns; 8160
    pixels: 9.2 / 7.3 / 9.3 -- the 32-wide tiling would need two waves there)."""
    m_tiles = None
    th = 1
    while th <= 16:                                     # the library's pixel tile: th x (128 / th) with the least padding
        tiles = B * (-(-Ho // th)) * (-(-Wo // (128 // th)))
        m_tiles = tiles if m_tiles is None else min(m_tiles, tiles)
        th *= 2
    widest = 128 if Cout > 64 else (64 if Cout > 32 else 32)
    for bn in (32, 64, 128):
        if bn <= widest and m_tiles * (-(-Cout // bn)) <= LATENCY_WAVE:
            return bn if bn != widest else 0
    return 0


def conv_gemm_ok(x_bhwc, kh=1, kw=1):
    """Whether conv_gemm can take this input directly (else callers keep the cuDNN / cuBLAS route)."""
    if not (x_bhwc.is_cuda and x_bhwc.dtype == torch.float32 and x_bhwc.dim() == 4):
        return False
    ld, ok = _pixel_layout(x_bhwc)
    return ok and x_bhwc.shape[-1] % 4 == 0 and kh % 2 == 1 and kw % 2 == 1


def conv_gemm(x_bhwc, w_hi, w_lo, kh, kw, bias=None, act=None, slope=0.1, residual=None, out=None, tile_n=0,
              aux1=None, aux2=None, split=0, out2=None, stride=1, dilation=1, single_pass=False):
    """Linear layer / stride-1 "same" convolution + bias + residual + activation in one tcgen05 kernel
    (include/camli_b200.h: camli_conv_gemm).  x_bhwc [B,H,W,Cin] channel-last view (a linear layer over rows
    is [1,1,R,K]); w_hi/w_lo [Cout, kh*kw*Cin] from tc_weight(); residual / out [B,H,W,Cout] channel-last views
    (out may be a channel slice of a wider buffer).  act "gru_gate" / "gru_update[_fix]" fuse the ConvGRU
    arithmetic (camli_conv_gemm_fused): aux1 / aux2 [B,H,W,*] channel-last side inputs, columns >= split of a gate
    convolution go to out2.  stride 2 (padding k/2): the output grid is ceil(H/2) x ceil(W/2) (camli_conv_gemm_strided).
    single_pass: ONE tf32 product instead of three (CAMLI_CONV_SINGLE_PASS; training under bf16 autocast only).
    Returns out."""
    _need_cuda(x_bhwc, w_hi, w_lo)
    _no_grad("conv_gemm", x_bhwc, w_hi)
    B, H, W, Cin = x_bhwc.shape
    Cout = w_hi.shape[0]
    assert w_hi.shape[1] == kh * kw * Cin, (tuple(w_hi.shape), kh, kw, Cin)
    ldx, ok = _pixel_layout(x_bhwc)
    if not ok or Cin % 4:
        raise RuntimeError("conv_gemm: input must be a 16-byte aligned channel-last view with Cin % 4 == 0")
    n_out = split if out2 is not None else Cout
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if tile_n == 0 and TILE_POLICY == "latency":
        tile_n = _latency_tile(B, Ho, Wo, Cout)
    if out is None:
        out = torch.empty((B, Ho, Wo, n_out), dtype=torch.float32, device=x_bhwc.device)
    assert tuple(out.shape) == (B, Ho, Wo, n_out)
    ldo = _pixel_layout(out)[0]
    ldr = 0
    if residual is not None:
        assert tuple(residual.shape) == (B, Ho, Wo, Cout)
        ldr = _pixel_layout(residual)[0]
    with torch.cuda.device(x_bhwc.device):
        native.call("camli_conv_gemm_strided", ptr(x_bhwc), i32(B), i32(H), i32(W), i32(Cin), i64(ldx), ptr(w_hi), ptr(w_lo),
                    i32(Cout), i32(kh), i32(kw), i32(stride), i32(dilation), ptr(bias), ptr(residual), i64(ldr), i32(ACT_CODES[act]),
                    ctypes.c_float(slope), ptr(out), i64(ldo),
                    ptr(aux1), i64(_pixel_layout(aux1)[0] if aux1 is not None else 0),
                    ptr(aux2), i64(_pixel_layout(aux2)[0] if aux2 is not None else 0), i32(split),
                    ptr(out2), i64(_pixel_layout(out2)[0] if out2 is not None else 0),
                    i32(tile_n | (0x100 if (single_pass or (SINGLE_PASS_INFERENCE and not torch.is_grad_enabled())) else 0)), stream(),
                    algo_bytes=B * (H * W * Cin + Ho * Wo * Cout) * 4 + Cout * kh * kw * Cin * 4,
                    flops=2 * B * Ho * Wo * Cout * kh * kw * Cin, shape=(B * Ho * Wo, Cout, kh * kw * Cin))   # GEMM M, N, K
    return out


def conv_small_n(x_bhwc, w2d, kh, kw, bias=None, act=None, slope=0.1, out=None):
    """Stride-1 "same" convolution / linear layer with Cout <= 4 on the CUDA cores (camli_conv_small_n):
    x_bhwc [B,H,W,Cin] channel-last view, w2d [Cout, kh*kw*Cin] (OHWI, fp32)."""
    _need_cuda(x_bhwc, w2d)
    _no_grad("conv_small_n", x_bhwc, w2d)
    B, H, W, Cin = x_bhwc.shape
    Cout = w2d.shape[0]
    assert w2d.shape[1] == kh * kw * Cin and w2d.is_contiguous()
    ldx, ok = _pixel_layout(x_bhwc)
    if not ok or Cin % 4:
        raise RuntimeError("conv_small_n: input must be a 16-byte aligned channel-last view with Cin % 4 == 0")
    if out is None:
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x_bhwc.device)
    ldo = _pixel_layout(out)[0]
    with torch.cuda.device(x_bhwc.device):
        native.call("camli_conv_small_n", ptr(x_bhwc), i32(B), i32(H), i32(W), i32(Cin), i64(ldx), ptr(w2d), i32(Cout),
                    i32(kh), i32(kw), ptr(bias), i32(ACT_CODES[act]), ctypes.c_float(slope), ptr(out), i64(ldo), stream(),
                    algo_bytes=B * H * W * (Cin + Cout) * 4, flops=2 * B * H * W * Cout * kh * kw * Cin)
    return out


def conv_small_cin(x_bhwc, w2d, kh, kw, bias=None, act=None, slope=0.1, out=None):
    """Stride-1 "same" convolution with Cin <= 4 (windows up to 7x7) on the CUDA cores (camli_conv_small_cin):
    x_bhwc [B,H,W,Cin] channel-last view (any pixel stride), w2d [Cout, kh*kw*Cin] (OHWI, fp32)."""
    _need_cuda(x_bhwc, w2d)
    _no_grad("conv_small_cin", x_bhwc, w2d)
    B, H, W, Cin = x_bhwc.shape
    Cout = w2d.shape[0]
    assert w2d.shape[1] == kh * kw * Cin and w2d.is_contiguous()
    ldx = _pixel_layout(x_bhwc)[0]
    if out is None:
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x_bhwc.device)
    ldo = _pixel_layout(out)[0]
    with torch.cuda.device(x_bhwc.device):
        native.call("camli_conv_small_cin", ptr(x_bhwc), i32(B), i32(H), i32(W), i32(Cin), i64(ldx), ptr(w2d), i32(Cout),
                    i32(kh), i32(kw), ptr(bias), i32(ACT_CODES[act]), ctypes.c_float(slope), ptr(out), i64(ldo), stream(),
                    algo_bytes=B * H * W * (Cin + Cout) * 4, flops=2 * B * H * W * Cout * kh * kw * Cin)
    return out


def stem_conv_pool(x_bhwc, w_ohwi, bias):
    """ResNet stem (7x7 stride-2 convolution 3 -> 64 with folded BatchNorm, ReLU, 3x3 stride-2 max-pool) in one kernel
    (camli_stem_conv_pool): x_bhwc [B,H,W,3] channel-last view -> [B,Hp,Wp,64] channel-last."""
    _need_cuda(x_bhwc, w_ohwi, bias)
    _no_grad("stem_conv_pool", x_bhwc, w_ohwi)
    B, H, W, C = x_bhwc.shape
    assert C == 3 and tuple(w_ohwi.shape) == (64, 7, 7, 3) and w_ohwi.is_contiguous() and bias.is_contiguous()
    ldx = _pixel_layout(x_bhwc)[0]
    sb, sh, sw, sc = x_bhwc.stride()
    if not (sc == 1 and sw == ldx and sh == W * ldx and (B == 1 or sb == H * W * ldx)):
        x_bhwc = x_bhwc.contiguous()
        ldx = 3
    Hc, Wc = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    Hp, Wp = (Hc - 1) // 2 + 1, (Wc - 1) // 2 + 1
    out = torch.empty((B, Hp, Wp, 64), dtype=torch.float32, device=x_bhwc.device)
    with torch.cuda.device(out.device):
        native.call("camli_stem_conv_pool", ptr(x_bhwc), i32(B), i32(H), i32(W), i64(ldx), ptr(w_ohwi), ptr(bias), ptr(out),
                    i64(64), stream(), algo_bytes=B * (H * W * 3 + Hp * Wp * 64) * 4, flops=2 * B * Hc * Wc * 64 * 147)
    return out


def linear_rows(x, w_hi, w_lo, bias=None, act=None, slope=0.1, residual=None):
    """x [..., K] contiguous rows -> [..., N] through conv_gemm (1x1)."""
    K = x.shape[-1]
    x2 = x.reshape(1, 1, -1, K)
    r2 = None if residual is None else residual.reshape(1, 1, -1, w_hi.shape[0])
    out = conv_gemm(x2, w_hi, w_lo, 1, 1, bias, act, slope, r2)
    return out.view(*x.shape[:-1], w_hi.shape[0])


# ---------------------------------------------------------------- training side of conv_gemm (dense-layer backward)
def split_tf32(w2d):
    """(hi, lo) tf32 parts of a contiguous fp32 tensor (camli_split_tf32), uncached."""
    hi, lo = torch.empty_like(w2d), torch.empty_like(w2d)
    with torch.cuda.device(w2d.device):
        native.call("camli_split_tf32", ptr(w2d), ptr(hi), ptr(lo), i64(w2d.numel()), stream())
    return hi, lo


def transpose_split(rows, y=None, act=None, slope=0.1, want_rows=False, want_colsum=False, n_shift=1, shift_step=1, want_lo=True,
                    xstride=1, bf16=False):
    """rows [B,H,W,C] channel-last view -> (hi_t, lo_t [C, B*H*W], g_rows, colsum): the transposed tf32 parts the
    weight-gradient GEMM reads (camli_transpose_split).  With `y` (the layer output, same shape) rows is dL/dy: it is
    multiplied by act'(y) first; g_rows [B,H,W,C] = that product (operand of the data-gradient convolution), colsum [C] =
    its sum over the pixels (the bias gradient).  n_shift = kw > 1: [kw, C, B*H*W] horizontally pre-shifted copies (the x
    operand of a kw-wide window, shift_step = dilation); xstride = 2: the copies keep every second pixel of a row (the x
    operand of a stride-2 layer), [kw, C, B*H*ceil(W/2)]."""
    B, H, W, C = rows.shape
    ld, ok = _pixel_layout(rows)
    if not ok and rows.stride(-1) != 1:
        raise RuntimeError("transpose_split: channel-last rows expected")
    P = B * H * W
    P_out = B * H * ((W - 1) // xstride + 1)
    hi_t = torch.empty((n_shift, C, P_out) if (n_shift > 1 or xstride > 1) else (C, P_out),
                       dtype=torch.bfloat16 if bf16 else torch.float32, device=rows.device)      # bf16: operands of the kind::f16 product
    lo_t = torch.empty_like(hi_t) if (want_lo and not bf16) else None
    g_rows = torch.empty((B, H, W, C), dtype=torch.float32, device=rows.device) if (want_rows and y is not None) else None
    colsum = torch.zeros((C,), dtype=torch.float32, device=rows.device) if want_colsum else None
    ldy = _pixel_layout(y)[0] if y is not None else 0
    with torch.cuda.device(rows.device):
        native.call("camli_transpose_split", ptr(rows), i64(ld), i64(P), i32(C), ptr(y), i64(ldy), i32(ACT_CODES[act]),
                    ctypes.c_float(slope), i32(W), i32(n_shift), i32(shift_step), i32(xstride | (0x100 if bf16 else 0)), ptr(hi_t), ptr(lo_t),
                    ptr(g_rows), ptr(colsum),
                    stream(), algo_bytes=P * C * 4 * (1 + 2 * n_shift + (2 if y is not None else 0)))
    return hi_t, lo_t, g_rows, colsum


def conv_wgrad(g_t, x_t, B, H, W, Cout, Cin, kh, kw, dilation=1, passes=3, stride=1, Hin=None, accumulate_into=None):
    """dW [Cout, kh*kw*Cin] (OHWI) of the stride-1 "same" convolution from the transposed hi / lo operand pairs of
    transpose_split (camli_conv_wgrad: 3xTF32 on tcgen05, K split over the SMs; passes = 1: one tf32 product, lo parts unused).
    H, W: the OUTPUT grid (= the grid of g); stride 2: Hin input rows, x_t prepared with xstride = 2.
    accumulate_into: a contiguous fp32 [Cout, kh*kw*Cin] buffer the result is ADDED to (CAMLI_WGRAD_ACCUMULATE) and returned."""
    if accumulate_into is None:
        dw = torch.empty((Cout, kh * kw * Cin), dtype=torch.float32, device=g_t[0].device)
    else:
        dw = accumulate_into
        if tuple(dw.shape) != (Cout, kh * kw * Cin) or not dw.is_contiguous() or dw.dtype != torch.float32 or dw.data_ptr() % 16:
            raise RuntimeError("conv_wgrad: accumulate_into must be a 16-byte aligned contiguous fp32 [Cout, kh*kw*Cin] buffer")
        passes |= 0x100
    with torch.cuda.device(dw.device):
        native.call("camli_conv_wgrad", ptr(g_t[0]), ptr(g_t[1]), ptr(x_t[0]), ptr(x_t[1]), i32(B), i32(H), i32(W), i32(Cout),
                    i32(Cin), i32(kh), i32(kw), i32(dilation), i32(stride), i32(H if Hin is None else Hin), i32(passes), ptr(dw), stream(),
                    algo_bytes=2 * B * H * W * (Cout + Cin) * 4 + Cout * kh * kw * Cin * 4,
                    flops=2 * B * H * W * Cout * kh * kw * Cin, shape=(Cout, kh * kw * Cin, B * H * W))
    return dw


def conv_wgrad_ok(B, H, W, Cin):
    return W % 4 == 0 and Cin % 4 == 0


# ---------------------------------------------------------------- RAFT all-pairs correlation
def corr2d_build(fmap1, fmap2, num_levels):
    """All-pairs volume of two [B,C,H,W] maps scaled by 1/sqrt(C) and its 2x2 average-pooled pyramid
    (models/raft_core.py:56-68).  Returns [B,H*W,h_l,w_l] per level; the coarser levels come from one
    fused pass over level 0."""
    fmap1, fmap2 = grad.f32(fmap1, fmap2)
    _need_cuda(fmap1, fmap2)
    B, C, H, W = fmap1.shape
    a = nhwc_rows(fmap1).view(B, H * W, C)
    b = nhwc_rows(fmap2).view(B, H * W, C)
    vol = allpairs(a, b, 1.0 / C ** 0.5).view(B, H * W, H, W)
    if grad.needs_grad(vol):
        with torch.autocast("cuda", enabled=False):
            return [vol] + list(_Corr2dPool.apply(vol, num_levels))
    return _corr2d_pool(vol, num_levels)


class _Corr2dPool(torch.autograd.Function):
    """The coarser pyramid levels of `vol`; backward = the chained avg_pool2d backward as one pass
    (camli_corr2d_pool_backward)."""

    @staticmethod
    def forward(ctx, vol, num_levels):
        ctx.shape, ctx.num_levels = vol.shape, num_levels
        return tuple(_corr2d_pool(vol, num_levels)[1:])

    @staticmethod
    def backward(ctx, *gs):
        B, P, H, W = ctx.shape
        dev = next(g for g in gs if g is not None).device
        h, w, coarse = H, W, []
        for g in gs:
            h, w = h // 2, w // 2
            coarse.append(torch.zeros((B, P, h, w), dtype=torch.float32, device=dev) if g is None else g.contiguous())
        g0 = torch.zeros(ctx.shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            native.call("camli_corr2d_pool_backward", ptr(g0), _ptr_array(coarse), i32(ctx.num_levels), i64(B * P), i32(H), i32(W),
                        stream(), algo_bytes=(2 * g0.numel() + sum(c.numel() for c in coarse)) * 4)
        return g0, None


def _corr2d_pool(vol, num_levels):
    B, _, H, W = vol.shape
    pyr = [vol]
    h, w = H, W
    for _ in range(num_levels - 1):
        h, w = h // 2, w // 2
        pyr.append(torch.empty((B, H * W, h, w), dtype=torch.float32, device=vol.device))
    if num_levels > 1:
        with torch.cuda.device(vol.device):
            native.call("camli_corr2d_pool_pyramid", ptr(vol), _ptr_array(pyr[1:]), i32(num_levels), i64(B * H * W),
                        i32(H), i32(W), stream(),
                        algo_bytes=sum(v.numel() for v in pyr) * 4)
    return pyr


def corr2d_lookup(pyramid, coords, radius, channels_last=True):
    """models/raft_core.py:71-107: coords [B,2,H,W] -> logical [B, L*(2r+1)^2, H, W] (NHWC storage when
    channels_last); window index i moves x, j moves y (the reference's meshgrid quirk)."""
    coords, pyramid = grad.f32(coords), [grad.f32(v) for v in pyramid]
    _need_cuda(coords, *pyramid)
    if grad.needs_grad(coords, *pyramid):
        with torch.autocast("cuda", enabled=False):
            return _Corr2dLookup.apply(coords, radius, channels_last, *pyramid)
    return _corr2d_lookup(pyramid, coords, radius, channels_last)


class _Corr2dLookup(torch.autograd.Function):
    """Backward by the scatter kernel camli_corr2d_lookup_backward: the gradient of volume slice V_l[p] only
    comes from pixel p's own window, so every (pixel, level) footprint is written by exactly one CTA -- no
    atomics, deterministic.  The coordinates get no gradient (the models look up at a detached flow,
    models/camliraft_core.py:105)."""

    @staticmethod
    def forward(ctx, coords, radius, channels_last, *pyramid):
        ctx.save_for_backward(coords)
        ctx.radius, ctx.shapes = radius, [tuple(v.shape) for v in pyramid]
        return _corr2d_lookup(list(pyramid), coords, radius, channels_last)

    @staticmethod
    def backward(ctx, g):
        (coords,) = ctx.saved_tensors
        grads = _corr2d_lookup_backward(g, coords, ctx.radius, ctx.shapes)
        return (None, None, None) + tuple(gv if need else None for gv, need in zip(grads, ctx.needs_input_grad[3:]))


def _corr2d_lookup_backward(g, coords, radius, shapes):
    """g: logical [B, L*81, H, W] -> zero-filled gradient volumes [B,HW,h_l,w_l] with the window footprints."""
    coords = coords.float().contiguous()
    B, _, H, W = coords.shape
    g_rows = nhwc_rows(g.float())
    grads = [torch.zeros(sh, dtype=torch.float32, device=g.device) for sh in shapes]
    with torch.cuda.device(g.device):
        native.call("camli_corr2d_lookup_backward", _ptr_array(grads), _int_array([sh[-2] for sh in shapes]),
                    _int_array([sh[-1] for sh in shapes]), i32(len(shapes)), ptr(coords), ptr(g_rows), i32(B), i32(H), i32(W),
                    i32(radius), stream(),
                    algo_bytes=B * H * W * (len(shapes) * ((2 * radius + 2) ** 2 + (2 * radius + 1) ** 2) * 4 + 8))
    return grads


def _corr2d_lookup(pyramid, coords, radius, channels_last):
    coords = coords.float().contiguous()
    B, _, H, W = coords.shape
    L, n_ch = len(pyramid), len(pyramid) * (2 * radius + 1) ** 2
    if channels_last:
        store = torch.empty((B, H, W, n_ch), dtype=torch.float32, device=coords.device)
        out = nchw_view(store)
    else:
        store = out = torch.empty((B, n_ch, H, W), dtype=torch.float32, device=coords.device)
    with torch.cuda.device(coords.device):
        native.call("camli_corr2d_lookup", _ptr_array(pyramid), _int_array([v.shape[-2] for v in pyramid]),
                    _int_array([v.shape[-1] for v in pyramid]), i32(L), ptr(coords), ptr(store), i32(B), i32(H), i32(W),
                    i32(radius), i32(1 if channels_last else 0), stream(),
                    algo_bytes=B * H * W * (L * ((2 * radius + 2) ** 2 + (2 * radius + 1) ** 2) * 4 + 8))
    return out


# ---------------------------------------------------------------- point all-pairs correlation
def corr3d_build(feat1, feat2, xyzs2, k=3):
    """models/camliraft_l_core.py:51-60: feat [B,C,n] -> volumes [B,n1,n2_l]."""
    _need_cuda(feat1, feat2)
    B, C, n1 = feat1.shape
    vol = allpairs(rows_of(feat1.float()), rows_of(feat2.float()), 1.0 / C)
    pyr = [vol]
    for i in range(1, len(xyzs2)):
        idx = k_nearest_neighbor(xyzs2[i - 1].detach(), xyzs2[i].detach(), k)
        if grad.needs_grad(pyr[-1]):
            with torch.autocast("cuda", enabled=False):
                pyr.append(_Corr3dPool.apply(pyr[-1], idx))
        else:
            pyr.append(_corr3d_pool(pyr[-1], idx))
    return pyr


class _Corr3dPool(torch.autograd.Function):
    """k-NN mean pooling of the point volume; backward: d vol_in[p, idx[q,j]] += g[p,q] / k (camli_corr3d_pool_backward)."""

    @staticmethod
    def forward(ctx, vol, idx):
        ctx.save_for_backward(idx)
        ctx.shape = vol.shape
        return _corr3d_pool(vol, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, n1, n_in = ctx.shape
        n_out, k = idx.shape[1], idx.shape[2]
        g = g.contiguous()
        g_in = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            native.call("camli_corr3d_pool_backward", i32(B), i32(n1), i32(n_in), i32(n_out), i32(k), ptr(g), ptr(idx.contiguous()),
                        ptr(g_in), stream(), algo_bytes=B * n1 * (n_in + n_out * (1 + 2 * k)) * 4)
        return g_in, None


def _corr3d_pool(vol, idx):
    B, n1, n_in = vol.shape
    n_out, k = idx.shape[1], idx.shape[2]
    vol = vol.contiguous()
    nxt = torch.empty((B, n1, n_out), dtype=torch.float32, device=vol.device)
    with torch.cuda.device(vol.device):
        native.call("camli_corr3d_pool", i32(B), i32(n1), i32(n_in), i32(n_out), i32(k), ptr(vol), ptr(idx),
                    ptr(nxt), stream(), algo_bytes=B * n1 * (n_in + n_out) * 4)
    return nxt


def corr3d_lookup_rows(xyz1, xyzs2, pyramid, W1, b1, W2, b2):
    """Correlation3D.forward before `merge` (models/camliraft_l_core.py:62-98), every level in one
    launch: rows [B,n1,32*L]."""
    xyz1, xyzs2, pyramid = grad.f32(xyz1), [grad.f32(x) for x in xyzs2], [grad.f32(v) for v in pyramid]
    _need_cuda(xyz1, *xyzs2, *pyramid)
    L = len(pyramid)
    if grad.needs_grad(W1, b1, W2, b2, *pyramid) and not grad.needs_grad(xyz1, *xyzs2):
        with torch.autocast("cuda", enabled=False):          # (CamLiRAFT: the clouds are warped by a detached flow)
            return _Corr3dLookup.apply(xyz1, W1, b1, W2, b2, L, *xyzs2, *pyramid)
    return grad.recompute(_corr3d_lookup_rows, grad.f_corr3d_lookup_rows, xyz1, W1, b1, W2, b2, 16, L, *xyzs2, *pyramid)


def _xyz_strides(xyzs2):
    strides = []
    for x in xyzs2:
        sb, sd, sp = x.stride()
        strides += [sb, sp, sd]
    return strides


class _Corr3dLookup(torch.autograd.Function):
    """Point-correlation lookup with gradients for the volumes and the cost MLP (camli_corr3d_lookup_backward: one
    launch for all levels, the neighbour search repeated instead of stored)."""

    @staticmethod
    def forward(ctx, xyz1, W1, b1, W2, b2, L, *rest):
        ctx.L = L
        ctx.save_for_backward(xyz1, W1, b1, W2, b2, *rest)
        return _corr3d_lookup_rows(xyz1, W1, b1, W2, b2, 16, L, *rest)

    @staticmethod
    def backward(ctx, g):
        xyz1, W1, b1, W2, b2, *rest = ctx.saved_tensors
        L = ctx.L
        xyzs2, pyramid = list(rest[:L]), [v.contiguous() for v in rest[L:]]
        B, _, n1 = xyz1.shape
        g = g.contiguous()
        gvol = [torch.zeros_like(v) for v in pyramid]
        gW1, gb1, gW2, gb2 = (torch.zeros_like(t, memory_format=torch.contiguous_format) for t in (W1, b1, W2, b2))
        with torch.cuda.device(g.device):
            native.call("camli_corr3d_lookup_backward", i32(B), i32(n1), i32(L), ptr(xyz1.contiguous()), _ptr_array(xyzs2),
                        _int_array(_xyz_strides(xyzs2), ctypes.c_int64), _int_array([x.shape[-1] for x in xyzs2]),
                        _ptr_array(pyramid), _ptr_array(gvol),
                        ptr(W1.contiguous()), ptr(b1.contiguous()), ptr(W2.contiguous()), ptr(b2.contiguous()),
                        ptr(g), i32(g.shape[-1]), ptr(gW1), ptr(gb1), ptr(gW2), ptr(gb2), stream(),
                        algo_bytes=B * sum(n1 * 12 + x.shape[-1] * 12 + n1 * 16 * 24 + n1 * 128 for x in xyzs2),
                        flops=B * sum(n1 * x.shape[-1] * 8 + n1 * 16 * 6 * (4 * 32 + 32 * 32) for x in xyzs2))
        return (None, gW1, gb1, gW2, gb2, None) + (None,) * L + tuple(gvol)


def _corr3d_lookup_rows(xyz1, W1, b1, W2, b2, k, L, *rest):
    xyzs2, pyramid = list(rest[:L]), [v.contiguous() for v in rest[L:]]
    W1, b1, W2, b2 = W1.contiguous(), b1.contiguous(), W2.contiguous(), b2.contiguous()
    xyz1 = xyz1.contiguous()
    B, _, n1 = xyz1.shape
    out = torch.empty((B, n1, 32 * L), dtype=torch.float32, device=xyz1.device)
    strides = []
    for x in xyzs2:
        s = x.stride()
        strides += [s[0], s[2], s[1]]
    with torch.cuda.device(xyz1.device):
        native.call("camli_corr3d_lookup", i32(B), i32(n1), i32(L), ptr(xyz1), _ptr_array(xyzs2),
                    _int_array(strides, ctypes.c_int64), _int_array([x.shape[-1] for x in xyzs2]), _ptr_array(pyramid),
                    ptr(W1), ptr(b1), ptr(W2), ptr(b2), ptr(out), i32(32 * L), stream(),
                    algo_bytes=B * sum(n1 * 12 + x.shape[-1] * 12 + n1 * k * 16 + n1 * 128 for x in xyzs2),
                    flops=B * sum(n1 * x.shape[-1] * 8 + n1 * k * 2 * (4 * 32 + 32 * 32) for x in xyzs2))
    return out


# ---------------------------------------------------------------- point convolutions
def pointconv_dw_weights(xyz, sampled_xyz, knn_idx, k, weight_net):
    """WeightNet(3->8->32->O, ReLU) of every neighbour offset as rows [B,S,k,O]
    (models/point_conv.py:122-127).  Depends only on geometry + layer parameters."""
    _need_cuda(xyz, sampled_xyz, knn_idx)
    convs = weight_net.convs
    params = []
    for c in convs:
        w, b = c.folded()
        params += [w, b]
    if not grad.needs_grad(xyz, sampled_xyz, *params) and params[4].shape[0] >= 32:
        return _pointconv_dw_weights_tc(xyz, sampled_xyz, knn_idx, k, params, convs[2].conv_fn)
    return grad.recompute(_pointconv_dw_weights, grad.f_pointconv_dw_weights, xyz, sampled_xyz, knn_idx, k, *params)


def _pointconv_dw_weights_tc(xyz, sampled_xyz, knn_idx, k, params, out_layer):
    """Inference route: hidden layer [B,S,k,32] by a small kernel, the 32 -> O output layer (where the flops are)
    as one tensor-core GEMM with its bias + ReLU epilogue over all B*S*k neighbours."""
    params = [p.contiguous() for p in params]
    B, _, N = xyz.shape
    S, K = knn_idx.shape[1], knn_idx.shape[2]
    knn_idx = knn_idx.contiguous()
    hidden = torch.empty((B, S, k, 32), dtype=torch.float32, device=xyz.device)
    xs, cs = xyz.stride(), sampled_xyz.stride()
    with torch.cuda.device(xyz.device):
        native.call("camli_pointconv_dw_hidden", i32(B), i32(N), i32(S), i32(K), i32(k),
                    ptr(xyz), i64(xs[0]), i64(xs[2]), i64(xs[1]), ptr(sampled_xyz), i64(cs[0]), i64(cs[2]), i64(cs[1]),
                    ptr(knn_idx), *[ptr(p) for p in params[:4]], ptr(hidden), stream(),
                    algo_bytes=B * S * k * (32 * 4 + 8 + 12), flops=2 * B * S * k * (24 + 256))
    w_hi, w_lo, bias = tc_weight([out_layer.weight, out_layer.bias], lambda: (params[4], params[5]))
    O = params[4].shape[0]
    return linear_rows(hidden.view(B, S * k, 32), w_hi, w_lo, bias, "relu").view(B, S, k, O)


def _pointconv_dw_weights(xyz, sampled_xyz, knn_idx, k, *params):
    params = [p.contiguous() for p in params]
    B, _, N = xyz.shape
    S, K = knn_idx.shape[1], knn_idx.shape[2]
    O = params[4].shape[0]
    assert params[0].shape == (8, 3) and params[2].shape == (32, 8) and params[4].shape[1] == 32
    knn_idx = knn_idx.contiguous()
    out = torch.empty((B, S, k, O), dtype=torch.float32, device=xyz.device)
    xs, cs = xyz.stride(), sampled_xyz.stride()
    with torch.cuda.device(xyz.device):
        native.call("camli_pointconv_dw_weights", i32(B), i32(N), i32(S), i32(K), i32(k), i32(O),
                    ptr(xyz), i64(xs[0]), i64(xs[2]), i64(xs[1]), ptr(sampled_xyz), i64(cs[0]), i64(cs[2]), i64(cs[1]),
                    ptr(knn_idx), *[ptr(p) for p in params], ptr(out), stream(),
                    algo_bytes=B * S * k * (O * 4 + 8 + 12), flops=2 * B * S * k * (24 + 256 + 32 * O))
    return out


def pointconv_dw_gather_max(feat_rows, weights, knn_idx, k):
    """out[b,s,o] = max_j feat_rows[b, idx[b,s,j], o] * weights[b,s,j,o] (models/point_conv.py:126-128):
    feat_rows [B,N,O], weights [B,S,k,O], knn_idx [B,S,K>=k] -> rows [B,S,O]."""
    feat_rows, weights = grad.f32(feat_rows, weights)
    _need_cuda(feat_rows, weights, knn_idx)
    if grad.needs_grad(feat_rows, weights):
        with torch.autocast("cuda", enabled=False):
            return _DwGatherMax.apply(feat_rows, weights, knn_idx, k)
    return _pointconv_dw_gather_max(feat_rows, weights, knn_idx, k)


class _DwGatherMax(torch.autograd.Function):
    """Backward by camli_pointconv_dw_gather_max_backward: the arg-max neighbour j* of every (s, o) is found
    again from the saved inputs; d feat[idx[s,j*], o] += g * w[s,j*,o] (atomic scatter), d w[s,j*,o] = g * feat[...]."""

    @staticmethod
    def forward(ctx, feat_rows, weights, knn_idx, k):
        feat_rows, weights = feat_rows.contiguous(), weights.contiguous()
        ctx.save_for_backward(feat_rows, weights, knn_idx)
        ctx.k = k
        return _pointconv_dw_gather_max(feat_rows, weights, knn_idx, k)

    @staticmethod
    def backward(ctx, g):
        feat_rows, weights, knn_idx = ctx.saved_tensors
        B, N, O = feat_rows.shape
        S, K = knn_idx.shape[1], knn_idx.shape[2]
        g = g.contiguous()
        g_feat = torch.zeros_like(feat_rows)
        g_w = torch.zeros_like(weights) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(g.device):
            native.call("camli_pointconv_dw_gather_max_backward", i32(B), i32(N), i32(S), i32(K), i32(ctx.k), i32(O),
                        ptr(feat_rows), ptr(weights), ptr(knn_idx.contiguous()), ptr(g), ptr(g_feat), ptr(g_w), stream(),
                        algo_bytes=B * S * (ctx.k * (2 * O * 4 + 8) + 3 * O * 4))
        return g_feat if ctx.needs_input_grad[0] else None, g_w, None, None


def _pointconv_dw_gather_max(feat_rows, weights, knn_idx, k):
    feat_rows, weights = feat_rows.contiguous(), weights.contiguous()
    knn_idx = knn_idx.contiguous()
    B, N, O = feat_rows.shape
    S, K = knn_idx.shape[1], knn_idx.shape[2]
    out = torch.empty((B, S, O), dtype=torch.float32, device=feat_rows.device)
    with torch.cuda.device(out.device):
        native.call("camli_pointconv_dw_gather_max", i32(B), i32(N), i32(S), i32(K), i32(k), i32(O), ptr(feat_rows), i64(O),
                    ptr(weights), ptr(knn_idx), ptr(out), i64(O), stream(),
                    # gathered neighbour features + indices read, reduced rows written.  (The kernel also streams
                    # the cached WeightNet rows, another B*S*k*O*4 bytes the reference computes in a separate op.)
                    algo_bytes=B * S * (k * (O * 4 + 8) + O * 4), flops=B * S * k * O * 2)
    return out


def pointconv_group(rows, sampled_xyz, knn_idx, k, weight_net, negative_slope):
    """PointConv grouping (models/point_conv.py:56-66): rows [B,N,3+C] = [xyz | features] channel-last,
    sampled_xyz [B,3,S], knn_idx [B,S,K>=k]; WeightNet(3->8->16) evaluated in-kernel.
    Returns [B,S,16*(3+C)] in the order the reference's Linear expects (weight-major)."""
    rows, sampled_xyz = grad.f32(rows, sampled_xyz)
    _need_cuda(rows, sampled_xyz, knn_idx)
    (w1, b1), (w2, b2) = weight_net.convs[0].folded(), weight_net.convs[1].folded()
    if not grad.needs_grad(rows, sampled_xyz, w1, b1, w2, b2):
        return _pointconv_group(rows, sampled_xyz, knn_idx, k, w1, b1, w2, b2, negative_slope)
    return _PointConvGroup.apply(rows, sampled_xyz, knn_idx, k, w1, b1, w2, b2, negative_slope)


class _PointConvGroup(torch.autograd.Function):
    """Backward by camli_pointconv_group_backward: the WeightNet is re-evaluated per neighbour, the row gradient is a
    scatter-add over the neighbour tables, the WeightNet's parameter gradients are reduced per CTA (176 atomics each)."""

    @staticmethod
    def forward(ctx, rows, sampled_xyz, knn_idx, k, w1, b1, w2, b2, slope):
        rows, knn_idx = rows.contiguous(), knn_idx.contiguous()
        ctx.save_for_backward(rows, sampled_xyz, knn_idx, w1, b1, w2, b2)
        ctx.k, ctx.slope = k, slope
        return _pointconv_group(rows, sampled_xyz, knn_idx, k, w1, b1, w2, b2, slope)

    @staticmethod
    def backward(ctx, g):
        rows, sampled_xyz, knn_idx, w1, b1, w2, b2 = ctx.saved_tensors
        B, N, C = rows.shape
        S, K = knn_idx.shape[1], knn_idx.shape[2]
        g = g.float().contiguous()
        g_rows = torch.zeros_like(rows)
        g_centre = torch.empty((B, 3, S), dtype=torch.float32, device=rows.device)
        g_par = torch.zeros(176, dtype=torch.float32, device=rows.device)
        cs = sampled_xyz.stride()
        with torch.cuda.device(rows.device):
            native.call("camli_pointconv_group_backward", i32(B), i32(N), i32(S), i32(K), i32(ctx.k), i32(C), ptr(rows), i64(C),
                        ptr(sampled_xyz), i64(cs[0]), i64(cs[2]), i64(cs[1]), ptr(knn_idx),
                        ptr(w1.contiguous()), ptr(b1.contiguous()), ptr(w2.contiguous()), ptr(b2.contiguous()),
                        ctypes.c_float(ctx.slope), ptr(g), ptr(g_rows), ptr(g_centre), ptr(g_par), stream(),
                        algo_bytes=B * S * (ctx.k * (2 * C * 4 + 8) + 16 * C * 4), flops=4 * B * S * 16 * ctx.k * C)
        return (g_rows, g_centre, None, None, g_par[:24].view(8, 3), g_par[24:32], g_par[32:160].view(16, 8), g_par[160:],
                None)


def _pointconv_group(rows, sampled_xyz, knn_idx, k, w1, b1, w2, b2, negative_slope):
    rows = rows.contiguous()
    assert w1.shape == (8, 3) and w2.shape == (16, 8)
    knn_idx = knn_idx.contiguous()
    B, N, C = rows.shape
    S, K = knn_idx.shape[1], knn_idx.shape[2]
    out = torch.empty((B, S, 16 * C), dtype=torch.float32, device=rows.device)
    cs = sampled_xyz.stride()
    with torch.cuda.device(rows.device):
        native.call("camli_pointconv_group", i32(B), i32(N), i32(S), i32(K), i32(k), i32(C), ptr(rows), i64(C),
                    ptr(sampled_xyz), i64(cs[0]), i64(cs[2]), i64(cs[1]), ptr(knn_idx),
                    ptr(w1.contiguous()), ptr(b1.contiguous()), ptr(w2.contiguous()), ptr(b2.contiguous()),
                    ctypes.c_float(negative_slope), ptr(out), stream(),
                    algo_bytes=B * S * (k * (C * 4 + 8) + 16 * C * 4), flops=2 * B * S * 16 * k * C)
    return out


# ---------------------------------------------------------------- CLFM
def nearest_point_2d(uv, H, W):
    """Index [B,H*W] of the projected point nearest to every pixel centre (2-D k-NN, k=1;
    models/clfm.py:57-60)."""
    from .utils import mesh_grid
    B = uv.shape[0]
    grid = mesh_grid(B, H, W, uv.device).reshape(B, 2, -1)
    return k_nearest_neighbor(uv, grid, 1)[..., 0]


def clfm_interp(uv, nn_idx, feat3d_rows, score_net, H, W):
    """FusionAwareInterp before out_conv (models/clfm.py:57-75): logical [B,C,H,W], NHWC storage."""
    uv, feat3d_rows = grad.f32(uv, feat3d_rows)
    _need_cuda(uv, nn_idx, feat3d_rows)
    (w1, b1), (w2, b2) = score_net[0].folded(), score_net[1].folded()
    if grad.needs_grad(w1, b1, w2, b2) and not grad.needs_grad(uv, feat3d_rows):
        with torch.autocast("cuda", enabled=False):          # (CLFM detaches both cross-modal inputs, models/clfm.py:34-38)
            return _ClfmInterp.apply(uv, nn_idx, feat3d_rows, w1, b1, w2, b2, H, W)
    return grad.recompute(_clfm_interp, grad.f_clfm_interp, uv, nn_idx, feat3d_rows, w1, b1, w2, b2, H, W)


class _ClfmInterp(torch.autograd.Function):
    """FusionAwareInterp before out_conv with gradients for the ScoreNet parameters (camli_clfm_interp_backward)."""

    @staticmethod
    def forward(ctx, uv, nn_idx, feat3d_rows, w1, b1, w2, b2, H, W):
        ctx.save_for_backward(uv, nn_idx, feat3d_rows, w1, b1, w2, b2)
        ctx.hw = (H, W)
        return _clfm_interp(uv, nn_idx, feat3d_rows, w1, b1, w2, b2, H, W)

    @staticmethod
    def backward(ctx, g):
        uv, nn_idx, feat3d_rows, w1, b1, w2, b2 = ctx.saved_tensors
        H, W = ctx.hw
        feat3d_rows = feat3d_rows.contiguous()
        B, N, C = feat3d_rows.shape
        g_rows = nhwc_rows(g)                                  # [B,H,W,C] channel-last storage
        gw1, gb1, gw2, gb2 = (torch.zeros_like(t, memory_format=torch.contiguous_format) for t in (w1, b1, w2, b2))
        with torch.cuda.device(g.device):
            native.call("camli_clfm_interp_backward", i32(B), i32(H), i32(W), i32(N), i32(C), ptr(uv.contiguous()),
                        ptr(nn_idx.contiguous()), ptr(feat3d_rows), i64(C), ptr(w1.contiguous()), ptr(b1.contiguous()),
                        ptr(w2.contiguous()), ptr(b2.contiguous()), ptr(g_rows), ptr(gw1), ptr(gb1), ptr(gw2), ptr(gb2), stream(),
                        algo_bytes=B * H * W * (16 + 2 * C * 4), flops=2 * B * H * W * C * 16 * 6)
        return None, None, None, gw1, gb1, gw2, gb2, None, None


def _clfm_interp(uv, nn_idx, feat3d_rows, w1, b1, w2, b2, H, W):
    feat3d_rows = feat3d_rows.contiguous()
    B, N, C = feat3d_rows.shape
    uv, nn_idx = uv.contiguous(), nn_idx.contiguous()
    store = torch.empty((B, H, W, C), dtype=torch.float32, device=uv.device)
    with torch.cuda.device(uv.device):
        native.call("camli_clfm_interp", i32(B), i32(H), i32(W), i32(N), i32(C), ptr(uv), ptr(nn_idx), ptr(feat3d_rows),
                    i64(C), ptr(w1.contiguous()), ptr(b1.contiguous()), ptr(w2.contiguous()), ptr(b2.contiguous()),
                    ptr(store), stream(), algo_bytes=B * H * W * (8 + 8 + 2 * C * 4), flops=2 * B * H * W * (48 + 16 * C))
    return nchw_view(store)


# ---------------------------------------------------------------- fused pointwise stages
def sk_fusion_tail(a_rows, b_rows, negative_slope, w_mid, w_out, out=None):
    """SKFusion after the align layers (models/clfm.py:199-214) on rows [B,P,C]: activation of the two
    inputs (leaky, slope 1 = already activated), pooled mean, two bias-free FCs, pair softmax, blend.
    out: optional [B,P,C] destination view with unit channel stride and row pitch ld (batch pitch P * ld): a channel
    slice of a wider channel-last buffer."""
    _need_cuda(a_rows, b_rows, w_mid, w_out)
    _no_grad("sk_fusion_tail", a_rows, b_rows, w_mid, w_out)
    assert a_rows.is_contiguous() and b_rows.is_contiguous() and a_rows.shape == b_rows.shape
    B, P, C = a_rows.shape
    Cm = w_mid.shape[0]
    assert w_mid.shape == (Cm, C) and w_out.shape == (2 * C, Cm)
    if out is None:
        out = torch.empty_like(a_rows)
    ld = out.stride(1)
    if tuple(out.shape) != (B, P, C) or out.stride(2) != 1 or ld < C or (B > 1 and out.stride(0) != P * ld) or out.dtype != torch.float32:
        raise RuntimeError("sk_fusion_tail: out must be a [B,P,C] fp32 view with unit channel stride and batch pitch P * row pitch")
    partial = torch.empty((B, 32, C), dtype=torch.float32, device=out.device)
    weights = torch.empty((B * (2 * C + Cm),), dtype=torch.float32, device=out.device)
    with torch.cuda.device(out.device):
        native.call("camli_sk_fusion_tail_strided", i32(B), i32(P), i32(C), i32(Cm), ptr(a_rows), ptr(b_rows),
                    ctypes.c_float(negative_slope), ptr(w_mid.contiguous()), ptr(w_out.contiguous()), ptr(partial),
                    ptr(weights), ptr(out), i64(ld), stream(), algo_bytes=B * P * C * 4 * 5)
    return out


def gru_gate(zr, h, x):
    """ConvGRU gates (models/raft_core.py:125-128): zr [B,2H,..] pre-activation z|r, h [B,H,..], x [B,X,..]
    (logical channel-first maps) -> z [B,H,..], rhx = [r*h | x] [B,H+X,..], NHWC storage."""
    _need_cuda(zr, h, x)
    _no_grad("gru_gate", zr, h, x)
    B, Hc, Hh, Ww = h.shape
    X = x.shape[1]
    zr_r, h_r, x_r = nhwc_rows(zr), nhwc_rows(h), nhwc_rows(x)
    z = torch.empty((B, Hh, Ww, Hc), dtype=torch.float32, device=h.device)
    rhx = torch.empty((B, Hh, Ww, Hc + X), dtype=torch.float32, device=h.device)
    rows = B * Hh * Ww
    with torch.cuda.device(h.device):
        native.call("camli_gru_gate", i64(rows), i32(Hc), i32(X), ptr(zr_r), ptr(h_r), ptr(x_r), ptr(z), ptr(rhx), stream(),
                    algo_bytes=rows * (2 * Hc + Hc + X + Hc + Hc + X) * 4)
    return nchw_view(z), nchw_view(rhx)


def gru_update(z, h, q, fix_nonfinite=False):
    """h' = (1-z)*h + z*tanh(q) (models/raft_core.py:129,136), optionally followed by nan_to_num (:138)."""
    _need_cuda(z, h, q)
    _no_grad("gru_update", z, h, q)
    z_r, h_r, q_r = nhwc_rows(z), nhwc_rows(h), nhwc_rows(q)
    out = torch.empty_like(h_r)
    with torch.cuda.device(h.device):
        native.call("camli_gru_update", i64(out.numel()), ptr(z_r), ptr(h_r), ptr(q_r), i32(1 if fix_nonfinite else 0),
                    ptr(out), stream(), algo_bytes=out.numel() * 16)
    return nchw_view(out)


def gru_gate_rows(zr, h, x):
    """Point-branch form of gru_gate on channel-last rows: zr [..., 2H] pre-activation z|r, h [..., H],
    x [..., X] -> z [..., H], rhx = [sigmoid(r)*h | x] [..., H+X] (models/camliraft_l_core.py:129-132)."""
    _need_cuda(zr, h, x)
    _no_grad("gru_gate", zr, h, x)
    zr, h, x = zr.contiguous(), h.contiguous(), x.contiguous()
    Hc, X = h.shape[-1], x.shape[-1]
    rows = h.numel() // Hc
    z = torch.empty_like(h)
    rhx = torch.empty(h.shape[:-1] + (Hc + X,), dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        native.call("camli_gru_gate", i64(rows), i32(Hc), i32(X), ptr(zr), ptr(h), ptr(x), ptr(z), ptr(rhx), stream(),
                    algo_bytes=rows * (2 * Hc + Hc + X + Hc + Hc + X) * 4)
    return z, rhx


def gru_update_rows(z, h, q):
    """h' = (1-z)*h + z*tanh(q) on rows (models/camliraft_l_core.py:133-134)."""
    _need_cuda(z, h, q)
    _no_grad("gru_update", z, h, q)
    z, h, q = z.contiguous(), h.contiguous(), q.contiguous()
    out = torch.empty_like(h)
    with torch.cuda.device(h.device):
        native.call("camli_gru_update", i64(out.numel()), ptr(z), ptr(h), ptr(q), i32(0), ptr(out), stream(),
                    algo_bytes=out.numel() * 16)
    return out


# ---------------------------------------------------------------- per-launch profiling (bench.py)
def profile_begin():
    """Start bracketing every C-ABI launch with CUDA events on its stream (native.call)."""
    native.profile_begin()


def profile_end():
    """Stop; {entry point: {launches, avg_us, total_us, bytes, flops}} -- live measurements of this process only
    (per-launch averages of the algorithmic bytes / flops each doorway above declares)."""
    return native.profile_end()
