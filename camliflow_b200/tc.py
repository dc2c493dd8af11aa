"""Layer-level doorway to the tensor-core implicit-GEMM kernel (ops.conv_gemm / camli_conv_gemm).

`conv2d` / `linear` evaluate an nn.Conv2d / nn.Conv1d(1) / nn.Linear -- optionally followed by an
eval-mode BatchNorm (folded), a residual add and an activation -- as ONE kernel on channel-last
data.  Without autograd that is the inference path.  Under autograd (training) a stride-1 layer runs
through grad.DenseFn: the same forward kernel, and a backward made of the hand-written transpose /
split pre-pass, the forward kernel on mirrored weights (data gradient) and the split-K weight-gradient
kernel (camli_conv_wgrad).  Only a layer the kernels do not cover (stride 2 under autograd, C_in % 4 != 0,
training-mode norm statistics) runs through torch (cuDNN / cuBLAS) -- explicitly, so callers never branch."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

ENABLED = True          # tests / A-B benchmarks flip this to route everything through torch
# dense layers under autograd: "tcgen05" = grad.DenseFn (hand-written forward + backward kernels), "library" = cuDNN / cuBLAS
TRAIN_DENSE = os.environ.get("CAMLI_TRAIN_DENSE", "tcgen05")
_DENSE_ACTS = (None, "relu", "leaky_relu", "tanh", "sigmoid")


def _train_passes():
    """3xTF32 (fp32-accurate) unless the caller runs the step under reduced-precision autocast: then one tf32 product."""
    return 1 if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") in (torch.bfloat16, torch.float16) else 3


def _train_route(x):
    return (ENABLED and TRAIN_DENSE == "tcgen05" and x.is_cuda and torch.is_grad_enabled()
            and x.dtype in (torch.float32, torch.bfloat16, torch.float16))


def _pad_dense(x_rows, weight4d, bias):
    """Zero-pad the channel dimensions the kernels need 16-byte granular (TMA rows): input channels of x and weight, output
    channels of weight and bias.  Plain differentiable pads: autograd slices the gradients back."""
    O, I = weight4d.shape[:2]
    pi, po = (-I) % 4, (-O) % 4
    if pi:
        x_rows = F.pad(x_rows, (0, pi))
        weight4d = F.pad(weight4d, (0, 0, 0, 0, 0, pi))
    if po:
        weight4d = F.pad(weight4d, (0, 0, 0, 0, 0, 0, 0, po))
        bias = None if bias is None else F.pad(bias, (0, po))
    return x_rows, weight4d, bias, O


def conv_train(x, weight, bias, stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1, act=None, slope=0.1):
    """Training route of a convolution on a logical [B,C,H,W] tensor: act(conv(x) + bias) through grad.DenseFn, or None when
    the layer is outside what the kernels cover (the caller then runs torch)."""
    from . import grad
    if not _train_route(x) or weight.dim() != 4 or act not in _DENSE_ACTS:
        return None
    O, I, kh, kw = weight.shape
    d = dilation[0]
    st = stride[0]
    if (tuple(stride) not in ((1, 1), (2, 2)) or groups != 1 or tuple(dilation) != (d, d) or kh % 2 == 0 or kw % 2 == 0
            or tuple(padding) != (d * (kh // 2), d * (kw // 2)) or x.dim() != 4 or x.shape[1] != I or I <= 4 or (st == 2 and d != 1)):
        return None
    B, _, H, W = x.shape
    if ((W - 1) // st + 1) % 4 or W % 4 or B * H * W < 128:
        return None
    rows, w4, b, O = _pad_dense(x.permute(0, 2, 3, 1), weight, bias)
    y = grad.DenseFn.apply(rows, w4, b, act, slope, d, _train_passes(), st)
    return (y if y.shape[-1] == O else y[..., :O]).permute(0, 3, 1, 2)


def linear_train(x, weight2d, bias, act=None, slope=0.1):
    """The same for a linear layer on rows x [..., K]: [..., N], or None."""
    from . import grad
    if not _train_route(x) or act not in _DENSE_ACTS:
        return None
    N, K = weight2d.shape
    if x.shape[-1] != K or K <= 4:
        return None
    rows = x.reshape(1, 1, -1, K)
    R = rows.shape[2]
    if R % 4 or R < 128:
        return None
    rows, w4, b, N = _pad_dense(rows, weight2d.reshape(N, K, 1, 1), bias)
    y = grad.DenseFn.apply(rows, w4, b, act, slope, 1, _train_passes(), 1)
    y = y if y.shape[-1] == N else y[..., :N]
    return y.reshape(*x.shape[:-1], N)


def module_train(module, x, act=None, slope=0.1):
    """Training route of an nn.Conv2d / nn.Conv1d(kernel 1) / nn.Linear applied to its usual input layout, or None."""
    if isinstance(module, nn.Conv2d):
        return conv_train(x, module.weight, module.bias, module.stride, module.padding, module.dilation, module.groups, act, slope)
    if isinstance(module, nn.Conv1d):
        if module.kernel_size != (1,) or module.stride != (1,) or module.groups != 1 or x.dim() != 3:
            return None
        y = linear_train(x.transpose(1, 2), module.weight[:, :, 0], module.bias, act, slope)
        return None if y is None else y.transpose(1, 2)
    if isinstance(module, nn.Linear):
        return linear_train(x, module.weight, module.bias, act, slope)
    return None

_TORCH_ACTS = {
    None: lambda v, s: v,
    "relu": lambda v, s: torch.relu(v),
    "leaky_relu": lambda v, s: F.leaky_relu(v, s),
    "tanh": lambda v, s: torch.tanh(v),
    "sigmoid": lambda v, s: torch.sigmoid(v),
    "relu_fix": lambda v, s: torch.nan_to_num(torch.relu(v)),
    "none_fix": lambda v, s: torch.nan_to_num(v),
}


def _bn_foldable(bn):
    return bn is None or isinstance(bn, nn.Identity) or (isinstance(bn, nn.modules.batchnorm._BatchNorm) and not bn.training)


def _fold(weight, bias, bn):
    """(weight, bias) with an eval-mode BatchNorm folded in (weight: [O, ...])."""
    if bn is None or isinstance(bn, nn.Identity):
        return weight, bias
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = weight * scale.view(-1, *([1] * (weight.dim() - 1)))
    b = bn.bias - bn.running_mean * scale
    if bias is not None:
        b = b + bias * scale
    return w, b


def _key_params(weight, bias, bn):
    if bn is None or isinstance(bn, nn.Identity):
        return [weight, bias]
    return [weight, bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]


_PLAIN = {}


def _plain_weight(key_params, builder):
    """Cached effective (weight [N, taps*Cin], bias) of a layer for the CUDA-core small-N kernel."""
    import weakref
    live = [p for p in key_params if p is not None]
    key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in live)
    slot = _PLAIN.get(key[0][0])
    if slot is None or slot[0] != key or slot[1]() is not live[0]:
        with torch.no_grad():
            w2d, bias = builder()
            slot = (key, weakref.ref(live[0]), w2d.float().contiguous(), None if bias is None else bias.float().contiguous())
            ops._publish_cached_weight()
        _PLAIN[key[0][0]] = slot
    return slot[2], slot[3]


def fused(x):
    return ENABLED and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled()


def conv2d(x, conv, act=None, slope=0.1, bn=None, residual=None, out=None):
    """act(bn(conv(x)) + residual) for an nn.Conv2d on a logical [B,C,H,W] tensor.  Fused route: channel-last
    storage in, channel-last storage out (`out`: optional [B,H,W,Cout] channel-last destination view, e.g. a
    channel slice of a concatenation buffer).  Returns the logical [B,Cout,H,W] result."""
    kh, kw = conv.kernel_size
    if (fused(x) and _bn_foldable(bn) and residual is None and conv.in_channels <= 4 and max(kh, kw) <= 7
            and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
            and conv.padding == (kh // 2, kw // 2) and kh % 2 == 1 and kw % 2 == 1):
        rows = x.permute(0, 2, 3, 1)                   # few input channels, many outputs: CUDA-core kernel
        if not ops._pixel_layout(rows)[1] and rows.stride(-1) != 1:
            rows = rows.contiguous()
        O = conv.out_channels

        def build_small():
            w, b = _fold(conv.weight, conv.bias, bn)
            return w.permute(0, 2, 3, 1).reshape(O, -1), b

        w2d, bias = _plain_weight(_key_params(conv.weight, conv.bias, bn), build_small)
        return ops.conv_small_cin(rows, w2d, kh, kw, bias, act, slope, out).permute(0, 3, 1, 2)
    dil = conv.dilation[0]
    if fused(x) and conv.in_channels > 4 and x.shape[1] == conv.in_channels and conv.in_channels % 4:
        x = padded_rows(x)                  # TMA rows are 16-byte granular: zero-pad the channels (one copy), the weight
                                            # gets matching zero columns below
    if (fused(x) and _bn_foldable(bn) and conv.stride in ((1, 1), (2, 2)) and conv.dilation == (dil, dil) and conv.groups == 1
            and conv.padding == (dil * (kh // 2), dil * (kw // 2)) and kh % 2 == 1 and kw % 2 == 1
            and x.shape[1] % 4 == 0 and 0 <= x.shape[1] - conv.in_channels < 8):
        stride = conv.stride[0]
        # `x` may carry the input channels padded with zeros up to a multiple of 4 (TMA rows are 16-byte granular):
        # the weight then gets matching zero columns (see padded_rows)
        cpad = x.shape[1] - conv.in_channels
        rows = x.permute(0, 2, 3, 1)
        if not ops.conv_gemm_ok(rows, kh, kw):
            rows = rows.contiguous()
        if ops.conv_gemm_ok(rows, kh, kw):
            O = conv.out_channels

            def build():
                w, b = _fold(conv.weight, conv.bias, bn)
                w = w.permute(0, 2, 3, 1)
                if cpad:
                    w = F.pad(w, (0, cpad))
                return w.reshape(O, -1), b

            if stride == 1 and dil == 1 and O <= 4 and residual is None and O * kh * kw * x.shape[1] * 4 <= 160 * 1024:
                w2d, bias = _plain_weight(_key_params(conv.weight, conv.bias, bn), build)
                return ops.conv_small_n(rows, w2d, kh, kw, bias, act, slope, out).permute(0, 3, 1, 2)

            w_hi, w_lo, bias = ops.tc_weight(_key_params(conv.weight, conv.bias, bn), build)
            res = None
            if residual is not None:
                res = residual.permute(0, 2, 3, 1)
                if not ops._pixel_layout(res)[1]:
                    res = res.contiguous()
            y = ops.conv_gemm(rows, w_hi, w_lo, kh, kw, bias, act, slope, res, out, stride=stride, dilation=dil)
            return y.permute(0, 3, 1, 2)
    # ---- under autograd (or outside the kernels' coverage): the training route, else torch
    fuse_act = residual is None and act in _DENSE_ACTS
    y = None
    if _train_route(x) and _bn_foldable(bn):
        w, b = _fold(conv.weight, conv.bias, bn)         # (eval-mode BatchNorm: differentiable w.r.t. its affine parameters)
        y = conv_train(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups, act if fuse_act else None, slope)
        if y is not None:
            bn = None
    elif _train_route(x):
        y = conv_train(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups, None, slope)
        fuse_act = False
    if y is None:
        y = conv(x)
        fuse_act = False
    if bn is not None:
        y = bn(y)
    if residual is not None:
        y = y + residual
    if not fuse_act:
        y = _TORCH_ACTS[act](y, slope)
    if out is not None:
        out.copy_(y.permute(0, 2, 3, 1))
        return out.permute(0, 3, 1, 2)
    return y


def padded_rows(x, multiple=4):
    """Logical [B,C,H,W] tensor -> channel-last storage with C padded by zeros to a multiple of `multiple`, returned as the
    logical [B,C+pad,H,W] view conv2d accepts for a layer whose in_channels is not 16-byte granular."""
    B, C, H, W = x.shape
    pad = (-C) % multiple
    buf = torch.empty((B, H, W, C + pad), dtype=torch.float32, device=x.device)
    buf[..., :C].copy_(x.permute(0, 2, 3, 1))
    if pad:
        buf[..., C:].zero_()
    return buf.permute(0, 3, 1, 2)


def conv2d_weights(x, weight, bias, padding, act=None, slope=0.1, out=None, residual=None):
    """Same for an explicit stride-1 (weight [O,I,kh,kw], bias) pair -- e.g. two convolutions merged into one;
    `weight` / `bias` must be long-lived tensors (they key the split cache)."""
    O, I, kh, kw = weight.shape
    if fused(x) and tuple(padding) == (kh // 2, kw // 2) and kh % 2 == 1 and kw % 2 == 1 and I % 4 == 0:
        rows = x.permute(0, 2, 3, 1)
        if not ops.conv_gemm_ok(rows, kh, kw):
            rows = rows.contiguous()
        if ops.conv_gemm_ok(rows, kh, kw):
            w_hi, w_lo, b = ops.tc_weight([weight, bias], lambda: (weight.permute(0, 2, 3, 1).reshape(O, -1), bias))
            res = None
            if residual is not None:
                res = residual.permute(0, 2, 3, 1)
                if not ops._pixel_layout(res)[1]:
                    res = res.contiguous()
            return ops.conv_gemm(rows, w_hi, w_lo, kh, kw, b, act, slope, res, out).permute(0, 3, 1, 2)
    fuse_act = residual is None and act in _DENSE_ACTS
    y = conv_train(x, weight, bias, (1, 1), tuple(padding), (1, 1), 1, act if fuse_act else None, slope)
    if y is None:
        y = F.conv2d(x, weight, bias, padding=padding)
        fuse_act = False
    if residual is not None:
        y = y + residual
    if not fuse_act:
        y = _TORCH_ACTS[act](y, slope)
    if out is not None:
        out.copy_(y.permute(0, 2, 3, 1))
        return out.permute(0, 3, 1, 2)
    return y


def linear(x, weight, bias=None, act=None, slope=0.1, bn=None):
    """act(bn(x @ weight.T + bias)) on channel-last rows x [..., K]; weight [N, K] or a 1x1 conv weight
    [N, K, 1(, 1)]."""
    w2 = weight.flatten(1)
    K = w2.shape[1]
    if fused(x) and _bn_foldable(bn) and x.shape[-1] == K:
        pad = (-K) % 4                          # TMA rows are 16-byte granular: zero-pad ragged inputs (K = 3: flow)
        rows = F.pad(x, (0, pad)) if pad else (x if x.is_contiguous() else x.contiguous())
        if rows.data_ptr() % 16 == 0:
            def build():
                w, b = _fold(weight.flatten(1), bias, bn)
                return (F.pad(w, (0, pad)) if pad else w), b

            keys = _key_params(weight, bias, bn)
            if w2.shape[0] <= 4:
                w2d, b = _plain_weight(keys, build)
                out = ops.conv_small_n(rows.reshape(1, 1, -1, K + pad), w2d, 1, 1, b, act, slope)
                return out.view(*x.shape[:-1], w2.shape[0])
            w_hi, w_lo, b = ops.tc_weight(keys, build)
            return ops.linear_rows(rows, w_hi, w_lo, b, act, slope)
    if bn is not None and not isinstance(bn, nn.Identity):
        if not _bn_foldable(bn):
            y = linear_train(x, w2, bias)
            if y is None:
                y = F.linear(x, w2, bias)
            y = bn(y.movedim(-1, 1)).movedim(1, -1)
            return _TORCH_ACTS[act](y, slope)
        w2, bias = _fold(w2, bias, bn)
    y = linear_train(x, w2, bias, act if act in _DENSE_ACTS else None, slope)
    if y is not None:
        return y if act in _DENSE_ACTS else _TORCH_ACTS[act](y, slope)
    return _TORCH_ACTS[act](F.linear(x, w2, bias), slope)
