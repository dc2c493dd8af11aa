"""1x1-convolution building blocks with the parameter names of the reference's models/mlp.py
(`conv_fn`, `norm_fn`, `convs.N`), so released checkpoints load unchanged.

One generic class covers the 1-D (point) and 2-D (image) variants; the reference spells them
out separately (mlp.py:41-128)."""
import torch
import torch.nn as nn

from . import tc


class LayerNormCF(nn.Module):
    """LayerNorm over the channel axis of a channel-first tensor (mlp.py:5-38)."""

    def __init__(self, n_channels, n_spatial, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n_channels))
        self.bias = nn.Parameter(torch.zeros(n_channels))
        self.eps = eps
        self._bshape = (1, n_channels) + (1,) * n_spatial

    def forward(self, x):
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        return self.weight.view(self._bshape) * ((x - mu) / torch.sqrt(var + self.eps)) + self.bias.view(self._bshape)


_ACTS = {
    "relu": lambda: nn.ReLU(inplace=True),
    "leaky_relu": lambda: nn.LeakyReLU(negative_slope=0.1, inplace=True),
    "sigmoid": nn.Sigmoid,
    None: nn.Identity,
}


def _make_norm(norm, n_channels, nd):
    bn, inorm = (nn.BatchNorm1d, nn.InstanceNorm1d) if nd == 1 else (nn.BatchNorm2d, nn.InstanceNorm2d)
    if norm is None:
        return nn.Identity()
    if norm == "batch_norm":
        return bn(n_channels)
    if norm == "instance_norm":
        return inorm(n_channels)
    if norm == "instance_norm_affine":
        return inorm(n_channels, affine=True)
    if norm == "layer_norm":
        return LayerNormCF(n_channels, nd)
    raise NotImplementedError("Unknown normalization function: %s" % norm)


class _ConvNormAct(nn.Module):
    ND = 1

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, groups=1,
                 norm=None, act="leaky_relu"):
        super().__init__()
        conv = nn.Conv1d if self.ND == 1 else nn.Conv2d
        self.conv_fn = conv(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                            bias=norm is None)
        self.norm_fn = _make_norm(norm, out_channels, self.ND)
        if act not in _ACTS:
            raise NotImplementedError("Unknown activation function: %s" % act)
        self.act_fn = _ACTS[act]()
        self.act = act

    def _foldable(self):
        n = self.norm_fn
        return isinstance(n, nn.Identity) or (isinstance(n, nn.modules.batchnorm._BatchNorm) and not n.training)

    def forward(self, x):
        if self.ND == 2 and tc.fused(x) and self._foldable():          # conv + norm + activation: one tcgen05 kernel
            return tc.conv2d(x, self.conv_fn, self.act, 0.1, bn=self.norm_fn)
        if isinstance(self.norm_fn, nn.Identity):
            y = tc.module_train(self.conv_fn, x, self.act)          # training: forward + backward on the tensor-core kernels
            if y is not None:
                return y
        else:
            y = tc.module_train(self.conv_fn, x)
            if y is not None:
                return self.act_fn(self.norm_fn(y))
        return self.act_fn(self.norm_fn(self.conv_fn(x)))

    def forward_rows(self, x):
        """Channel-last evaluation of a 1x1 layer: x [..., C_in] -> [..., C_out] (one GEMM + epilogue)."""
        if not self._foldable() or self.conv_fn.weight[0, 0].numel() != 1:     # training-mode norm or a real kernel window
            return self.forward(x.movedim(-1, 1)).movedim(1, -1)
        return tc.linear(x, self.conv_fn.weight, self.conv_fn.bias, self.act, 0.1, bn=self.norm_fn)

    def folded(self):
        """(weight [O,I], bias [O]) of the 1x1 convolution with an eval-mode BatchNorm folded in."""
        w = self.conv_fn.weight.flatten(1)
        b = self.conv_fn.bias if self.conv_fn.bias is not None else torch.zeros_like(w[:, 0])
        n = self.norm_fn
        if isinstance(n, nn.modules.batchnorm._BatchNorm):
            s = n.weight / torch.sqrt(n.running_var + n.eps)
            w, b = w * s[:, None], (b - n.running_mean) * s + n.bias
        elif not isinstance(n, nn.Identity):
            raise NotImplementedError("only BatchNorm (eval) folds into a convolution")
        return w, b


class Conv1dNormRelu(_ConvNormAct):
    ND = 1


class Conv2dNormRelu(_ConvNormAct):
    ND = 2


class _MLP(nn.Module):
    LAYER = Conv1dNormRelu

    def __init__(self, in_channels, mlp_channels, norm=None, act="leaky_relu"):
        super().__init__()
        assert isinstance(in_channels, int) and isinstance(mlp_channels, list)
        dims = [in_channels] + mlp_channels
        self.convs = nn.ModuleList(self.LAYER(i, o, norm=norm, act=act) for i, o in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for layer in self.convs:
            x = layer(x)
        return x

    def forward_rows(self, x):
        for layer in self.convs:
            x = layer.forward_rows(x)
        return x


class MLP1d(_MLP):
    LAYER = Conv1dNormRelu


class MLP2d(_MLP):
    LAYER = Conv2dNormRelu
