"""Autograd for the fused operators of camliflow_b200.ops (training path, BASELINE config 5).

Forward always runs the fused CUDA kernel.  Backward is one of
  * a hand-written backward kernel (camli_*_backward in include/camli_b200.h) where the gradient is a
    scatter / gather the library owns: PointConvDW gather-max, the RAFT correlation lookup, three-NN
    interpolation, the all-pairs product;
  * otherwise a *recompute* backward: the operator's formula (the reference's own lines, cited) is
    re-evaluated with stock differentiable torch ops on the saved inputs and differentiated by autograd.
    Nothing of the forward's intermediate tensors is kept alive between forward and backward, which is
    what lets a 10-iteration training graph of two 960x540 pairs fit comfortably.

The neighbour tables (k-NN, nearest projected point) are integer-valued and piecewise constant: they are
recomputed by the search kernels inside the backward and carry no gradient, exactly like the reference,
whose `k_nearest_neighbor` is not differentiable either (models/csrc/wrapper.py:106-127)."""
import torch
import torch.nn.functional as F

from .csrc import k_nearest_neighbor


# --------------------------------------------------------------------------------------- formulas
def gather_cf(data, idx):
    """data [B,C,N], idx [B,...] -> [B,C,...] (models/utils.py:62-80)."""
    B, C = data.shape[:2]
    flat = idx.reshape(B, 1, -1).expand(B, C, -1)
    return torch.gather(data, 2, flat).view([B, C] + list(idx.shape[1:]))


def gather_rows(rows, idx):
    """rows [B,N,C], idx [B,S,k] -> [B,S,k,C]."""
    B, N, C = rows.shape
    S, k = idx.shape[1:]
    flat = idx.reshape(B, S * k, 1).expand(B, S * k, C)
    return torch.gather(rows, 1, flat).view(B, S, k, C)


def f_knn_interpolate(input_xyz, input_feat, query_xyz, k):
    """models/utils.py:130-146."""
    idx = k_nearest_neighbor(input_xyz.detach(), query_xyz.detach(), k)
    d = torch.linalg.norm(gather_cf(input_xyz, idx) - query_xyz[..., None], dim=1).clamp(1e-8)
    w = 1.0 / d
    w = w / torch.sum(w, -1, keepdim=True)
    return torch.sum(gather_cf(input_feat, idx) * w[:, None], -1)


def f_bilinear_sample_rows(feat2d, uv):
    """models/utils.py:262-269 -> rows [B,N,C]."""
    H, W = feat2d.shape[2:]
    gx = 2.0 * uv[:, 0] / (W - 1) - 1.0
    gy = 2.0 * uv[:, 1] / (H - 1) - 1.0
    g = torch.stack([gx, gy], -1)[:, :, None, :]
    return F.grid_sample(feat2d, g, "bilinear", align_corners=True)[..., 0].transpose(1, 2)


def f_corr2d_pool(vol, num_levels):
    """models/raft_core.py:65-68: vol [B,HW,H,W] -> the list of pooled levels."""
    B, P, H, W = vol.shape
    v = vol.reshape(B * P, 1, H, W)
    pyr = [vol]
    for _ in range(num_levels - 1):
        v = F.avg_pool2d(v, 2, stride=2)
        pyr.append(v.view(B, P, v.shape[-2], v.shape[-1]))
    return pyr


def f_corr2d_lookup(coords, radius, *pyramid):
    """models/raft_core.py:71-107."""
    r = radius
    coords = coords.permute(0, 2, 3, 1).float()
    B, H, W, _ = coords.shape
    d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), -1).view(1, 2 * r + 1, 2 * r + 1, 2)
    out = []
    for i, vol in enumerate(pyramid):
        h, w = vol.shape[-2:]
        c = coords.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta
        g = torch.cat([2 * c[..., 0:1] / (w - 1) - 1, 2 * c[..., 1:2] / (h - 1) - 1], -1)
        s = F.grid_sample(vol.reshape(B * H * W, 1, h, w), g, align_corners=True)
        out.append(s.view(B, H, W, -1))
    return torch.cat(out, -1).permute(0, 3, 1, 2)


def f_corr3d_pool(vol, idx):
    """models/camliraft_l_core.py:56-60: vol [B,n1,n_in], idx [B,n_out,k] -> [B,n1,n_out]."""
    return torch.mean(gather_cf(vol, idx), -1)


def f_corr3d_lookup_rows(xyz1, W1, b1, W2, b2, k, n_levels, *rest):
    """models/camliraft_l_core.py:62-98 -> rows [B,n1,32*L]; rest = xyzs2 levels, then pyramid levels."""
    xyzs2, pyramid = rest[:n_levels], rest[n_levels:]
    costs = []
    for xyz2, vol in zip(xyzs2, pyramid):
        B, n1, n2 = vol.shape
        idx = k_nearest_neighbor(xyz2.detach(), xyz1.detach(), k)
        off = gather_cf(xyz2, idx) - xyz1[:, :, :, None]
        c = torch.gather(vol, 2, idx).view(B, 1, n1, -1)
        x = torch.cat([off, c], 1)
        x = F.relu(F.conv2d(x, W1[:, :, None, None], b1))
        x = F.relu(F.conv2d(x, W2[:, :, None, None], b2))
        costs.append(x.sum(-1))
    return torch.cat(costs, 1).transpose(1, 2)


def _pointwise_relu(x, w, b):
    """relu(conv1x1(x)) on [B,C,S,k]; on the GPU, where the layer qualifies (C % 4 == 0), forward and backward through the
    hand-written dense-layer kernels (tc.conv_train -> DenseFn) instead of cuDNN."""
    from . import tc
    y = tc.conv_train(x, w[:, :, None, None], b, act="relu") if x.is_cuda else None
    return F.relu(F.conv2d(x, w[:, :, None, None], b)) if y is None else y


def f_pointconv_dw_weights(xyz, sampled_xyz, knn_idx, k, *params):
    """models/point_conv.py:122-127 -> [B,S,k,O]."""
    x = gather_cf(xyz, knn_idx[:, :, :k]) - sampled_xyz[:, :, :, None]
    for w, b in zip(params[0::2], params[1::2]):
        x = _pointwise_relu(x, w, b)
    return x.permute(0, 2, 3, 1)


def f_pointconv_group(rows, sampled_xyz, knn_idx, k, w1, b1, w2, b2, slope):
    """models/point_conv.py:56-66 -> [B,S,16*(3+C)]."""
    idx = knn_idx[:, :, :k]
    B, S = idx.shape[:2]
    g = gather_rows(rows, idx)                                                        # [B,S,k,3+C]
    off = (g[..., :3] - sampled_xyz.transpose(1, 2)[:, :, None, :]).permute(0, 3, 1, 2)   # [B,3,S,k]
    w = F.leaky_relu(F.conv2d(off, w1[:, :, None, None], b1), slope)
    w = F.leaky_relu(F.conv2d(w, w2[:, :, None, None], b2), slope).transpose(1, 2)      # [B,S,16,k]
    return torch.matmul(w, g).reshape(B, S, -1)


def f_clfm_interp(uv, nn_idx, feat3d_rows, w1, b1, w2, b2, H, W):
    """models/clfm.py:57-75 before out_conv -> [B,C,H,W]."""
    B = uv.shape[0]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=uv.device),
                            torch.arange(W, dtype=torch.float32, device=uv.device), indexing="ij")
    grid = torch.stack([xs, ys], 0).reshape(1, 2, -1).expand(B, 2, -1)
    off = gather_cf(uv, nn_idx) - grid
    si = torch.cat([off, torch.linalg.norm(off, dim=1, keepdim=True)], 1)[..., None]
    s = F.leaky_relu(F.conv2d(si, w1[:, :, None, None], b1), 0.1)
    s = torch.sigmoid(F.conv2d(s, w2[:, :, None, None], b2))
    feat = gather_cf(feat3d_rows.transpose(1, 2), nn_idx)
    return (s[..., 0] * feat).view(B, -1, H, W)


# --------------------------------------------------------------------------------------- machinery
def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def f32(*tensors):
    """The fused operators are fp32 islands (the reference marks the same sites with explicit `.float()`
    casts): under bf16 / fp16 autocast their floating-point inputs are brought back to fp32."""
    out = tuple(t.float() if isinstance(t, torch.Tensor) and t.is_floating_point() and t.dtype != torch.float32 else t
                for t in tensors)
    return out[0] if len(out) == 1 else out


class _Recompute(torch.autograd.Function):
    """forward: `kernel(*args)` under no_grad; backward: autograd through `formula(*args)` recomputed on
    detached copies.  `args` may mix tensors and plain Python values; outputs: a tensor or a list of tensors.
    Tensor arguments are kept through save_for_backward (so autograd's version check catches an in-place write
    between forward and backward); everything else on the context."""

    @staticmethod
    def forward(ctx, kernel, formula, *args):
        ctx.formula = formula
        ctx.is_tensor = [isinstance(a, torch.Tensor) for a in args]
        ctx.plain = [None if t else a for a, t in zip(args, ctx.is_tensor)]
        ctx.save_for_backward(*[a for a, t in zip(args, ctx.is_tensor) if t])
        with torch.no_grad():
            out = kernel(*args)
        ctx.multi = isinstance(out, (list, tuple))
        return tuple(out) if ctx.multi else out

    @staticmethod
    def backward(ctx, *gouts):
        needs = ctx.needs_input_grad[2:]
        saved = iter(ctx.saved_tensors)
        args = []
        for plain, is_t, n in zip(ctx.plain, ctx.is_tensor, needs):
            a = plain
            if is_t:
                a = next(saved).detach()
                if n:
                    a.requires_grad_(True)
            args.append(a)
        with torch.enable_grad(), torch.autocast("cuda", enabled=False):
            out = ctx.formula(*args)
        outs = list(out) if isinstance(out, (list, tuple)) else [out]
        pairs = [(o, g) for o, g in zip(outs, gouts) if g is not None and o.requires_grad]
        wrt = [a for a, n in zip(args, needs) if n]
        grads = iter(torch.autograd.grad([o for o, _ in pairs], wrt, [g for _, g in pairs], allow_unused=True)
                     if pairs and wrt else [None] * len(wrt))
        return (None, None) + tuple(next(grads) if n else None for n in needs)


def recompute(kernel, formula, *args):
    """Run `kernel(*args)`; when a tensor argument requires grad, make the call differentiable through
    `formula` (see the module docstring).  fp32 island: autocast is off inside, low-precision inputs are cast up."""
    args = tuple(f32(a) if isinstance(a, torch.Tensor) else a for a in args)
    if not needs_grad(*args):
        return kernel(*args)
    with torch.autocast("cuda", enabled=False):
        out = _Recompute.apply(kernel, formula, *args)
    return list(out) if isinstance(out, tuple) else out


# --------------------------------------------------------------------------------------- dense layers
BF16_WGRAD = True  # under autocast (passes = 1): bf16 operands in the weight-gradient GEMM (A/B switch)
# Weight gradients of 1x1 layers added straight into the parameter's own .grad (the trainer's flat buffer) by the
# weight-gradient kernel -- its epilogue is an atomic add anyway -- instead of a zero-fill, a fresh tensor and autograd's
# AccumulateGrad add per use (a recurrent update block uses every weight once per refinement iteration).  Only inside
# fused_wgrad_accumulation(): torch.autograd.grad(...) and plain .backward() callers keep the functional behaviour.
# Measured on the C5 step (bench.py --workload c5, CAMLI_FUSE_WGRAD=1 / 0): 37 of 177 weight-gradient launches qualify
# (leaf 1x1 weights), the gradients agree to 3e-8, and the step gets SLOWER (97.5 -> 101.3 ms): the trainer leaves it off.
_FUSE_WGRAD = False


class fused_wgrad_accumulation:
    """with grad.fused_wgrad_accumulation(): loss.backward() -- parameters must have contiguous fp32 .grad buffers already."""

    def __init__(self, enabled=True):
        self.enabled = enabled

    def __enter__(self):
        global _FUSE_WGRAD
        self.prev, _FUSE_WGRAD = _FUSE_WGRAD, self.enabled
        return self

    def __exit__(self, *exc):
        global _FUSE_WGRAD
        _FUSE_WGRAD = self.prev
        return False


def _grad_slot(weight, O, K):
    """The [O, K] view of the leaf parameter's gradient buffer this 4-D weight (or 4-D view of a Linear / Conv1d weight)
    accumulates into, or None when the fused accumulation does not apply."""
    if not _FUSE_WGRAD:
        return None
    leaf = weight._base if weight._base is not None else weight
    g = leaf.grad
    if (not leaf.is_leaf or g is None or g.dtype != torch.float32 or leaf.dtype != torch.float32 or not g.is_contiguous()
            or g.numel() != O * K or g.data_ptr() % 16):
        return None
    return g.view(O, K)
_DENSE_W = {}      # (data_ptr, version, shape, flipped) -> (weak ref, hi, lo): lives for ONE training step (clear_dense_cache)


def clear_dense_cache():
    """Drop the split weights of the dense-layer autograd node.  trainer.train_step calls it at the start of every step (and
    CapturedTrainStep after its capture), so a cached split can never outlive the parameter values it was made from."""
    _DENSE_W.clear()


def _dense_weight(weight, flipped):
    """tf32 hi / lo parts of a conv / linear weight [O,I,kh,kw] as the OHWI matrix conv_gemm reads: [O, kh*kw*I], or --
    `flipped` -- the matrix of the DATA-gradient convolution, [I, kh*kw*O] with the window mirrored."""
    import weakref
    from . import ops
    owner = weight._base if weight._base is not None else weight       # (a Linear / Conv1d weight arrives as a 4-D view)
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), flipped)
    hit = _DENSE_W.get(key)
    if hit is None or hit[0]() is not owner:
        with torch.no_grad():
            w = weight.detach().float()
            if flipped:
                w2d = w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1).contiguous()
            else:
                w2d = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()
            hit = (weakref.ref(owner),) + ops.split_tf32(w2d)
        _DENSE_W[key] = hit
    return hit[1], hit[2]


class DenseFn(torch.autograd.Function):
    """act(conv(x, weight) + bias) for a stride-1 "same" convolution / linear layer on channel-last rows, forward AND backward on
    the hand-written tensor-core kernels: forward = camli_conv_gemm; backward = camli_transpose_split (activation derivative,
    bias gradient, K-major operands), camli_conv_gemm on the mirrored weights (data gradient), camli_conv_wgrad (weight
    gradient).  Replaces, for training, what autograd does with cuDNN / cuBLAS in the reference (train.py:143-171).
    x_rows [B,H,W,Cin] contiguous fp32, weight [O,I,kh,kw], bias [O] or None -> [B,H,W,O].
    passes = 3: every product is 3xTF32 (fp32-accurate; the parity mode).  passes = 1: one tf32 product per element in all
    three GEMMs -- the mode tc picks under bf16 / fp16 autocast, where the caller asked for reduced-precision dense layers
    (a tf32 operand keeps 10 mantissa bits, a bf16 one 7); accumulation, bias, activations and gradients stay fp32."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x_rows, weight, bias, act, slope, dilation, passes=3, stride=1):
        from . import ops
        kh, kw = weight.shape[2:]
        w_hi, w_lo = _dense_weight(weight, False)
        x_rows = x_rows.contiguous()
        y = ops.conv_gemm(x_rows, w_hi, w_lo, kh, kw, None if bias is None else bias.detach().float().contiguous(), act, slope,
                          dilation=dilation, single_pass=passes == 1, stride=stride)
        ctx.act, ctx.slope, ctx.dilation, ctx.passes, ctx.stride = act, slope, dilation, passes, stride
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x_rows, weight, y if act is not None else None)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        from . import ops
        x_rows, weight, y = ctx.saved_tensors
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        B, Hin, Win, Cin = x_rows.shape
        O, _, kh, kw = weight.shape
        gy = gy.float().contiguous()
        H, W = gy.shape[1:3]                                     # output grid (= input grid for stride 1)
        fast, st = ctx.passes == 1, ctx.stride
        # reduced-precision mode: the weight gradient multiplies bf16 operands (kind::f16, fp32 accumulation) -- half the bytes
        # through the transposed copies and L2, twice the MMA rate; what the reference's autocast does in this GEMM anyway
        wg_bf16 = fast and need_w and W % 8 == 0 and BF16_WGRAD
        g_hi, g_lo, g_rows, db = ops.transpose_split(gy, y, ctx.act, ctx.slope, want_rows=need_x, want_colsum=need_b,
                                                     want_lo=not fast, bf16=wg_bf16)
        if y is None:
            g_rows = gy
        dx = dw = None
        if need_x:
            wt_hi, wt_lo = _dense_weight(weight, True)          # (callers pad C_out to a multiple of 4: tc._pad_dense)
            if st == 2:
                # data gradient of a stride-2 layer = the stride-1 convolution of the zero-upsampled gradient with the mirrored
                # weights (3/4 of the products multiply zeros; two such layers per encoder pass)
                up = torch.zeros((B, Hin, Win, O), dtype=torch.float32, device=gy.device)
                up[:, ::2, ::2] = g_rows
                g_rows = up
            dx = ops.conv_gemm(g_rows, wt_hi, wt_lo, kh, kw, dilation=ctx.dilation, single_pass=fast)
        if need_w:
            x_hi, x_lo, _, _ = ops.transpose_split(x_rows, n_shift=kw, shift_step=ctx.dilation, want_lo=not fast, xstride=st,
                                                   bf16=wg_bf16)
            slot = _grad_slot(weight, O, Cin) if kh == 1 and kw == 1 else None      # (OHWI = OIHW only for a 1x1 window)
            dw2d = ops.conv_wgrad((g_hi, g_lo), (x_hi, x_lo), B, H, W, O, Cin, kh, kw, ctx.dilation, 16 if wg_bf16 else ctx.passes,
                                  st, Hin, accumulate_into=slot)
            dw = None if slot is not None else dw2d.view(O, kh, kw, Cin).permute(0, 3, 1, 2).to(weight.dtype)
        return dx, dw, (db.to(weight.dtype) if need_b else None), None, None, None, None, None
