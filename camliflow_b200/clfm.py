"""CLFM: bidirectional camera<->LiDAR fusion (reference models/clfm.py), SK fusion only --
the variant every reference config uses.  3D->2D: every pixel takes the features of its
nearest projected point, gated by a ScoreNet of the pixel->point offset; 2D->3D: bilinear
sample of the image features at the projected points."""
import torch
import torch.nn as nn

from . import ops
from .mlp import Conv1dNormRelu, Conv2dNormRelu


class FusionAwareInterp(nn.Module):
    """clfm.py:43-79 (k = 1)."""

    def __init__(self, n_channels_3d, k=1, norm=None):
        super().__init__()
        if k != 1:
            raise NotImplementedError("FusionAwareInterp: the reference configs only use k=1")
        self.k = k
        self.out_conv = Conv2dNormRelu(n_channels_3d, n_channels_3d, norm=norm)
        self.score_net = nn.Sequential(Conv2dNormRelu(3, 16), Conv2dNormRelu(16, n_channels_3d, act="sigmoid"))

    def forward(self, uv, feat_2d, feat_3d, nn_idx=None):
        return self.forward_rows(uv, feat_2d.shape[-2:], ops.rows_of(feat_3d), nn_idx)

    def forward_rows(self, uv, hw, feat3d_rows, nn_idx=None):
        """feat3d_rows [B,N,C] -> [B,C,H,W] (NHWC storage).  `nn_idx` [B,HW]: the pixel -> nearest
        projected point table; it only depends on (uv, H, W), so callers looping over fusion
        sites / GRU iterations compute it once."""
        H, W = hw
        if nn_idx is None:
            nn_idx = ops.nearest_point_2d(uv, H, W)
        return self.out_conv(ops.clfm_interp(uv, nn_idx, feat3d_rows, self.score_net, H, W))


class SKFusion(nn.Module):
    """Selective-kernel blend of two aligned feature sets (clfm.py:171-214)."""

    def __init__(self, in_channels_2d, in_channels_3d, out_channels, feat_format, norm=None, reduction=1):
        super().__init__()
        if feat_format == "nchw":
            layer = Conv2dNormRelu
        elif feat_format == "ncm":
            layer = Conv1dNormRelu
        else:
            raise ValueError(feat_format)
        self.align1 = layer(in_channels_2d, out_channels, norm=norm)
        self.align2 = layer(in_channels_3d, out_channels, norm=norm)
        self.fc_mid = nn.Sequential(nn.Linear(out_channels, out_channels // reduction, bias=False),
                                    nn.ReLU(inplace=True))
        self.fc_out = nn.Sequential(nn.Linear(out_channels // reduction, out_channels * 2, bias=False), nn.Sigmoid())

    def _blend_weights(self, pooled):
        B, C = pooled.shape
        return torch.softmax(self.fc_out(self.fc_mid(pooled)).view(B, C, 2), dim=-1)

    def _fused_tail_ok(self, a):
        return a.is_cuda and not (torch.is_grad_enabled() and (a.requires_grad or self.fc_mid[0].weight.requires_grad))

    def forward(self, feat_2d, feat_3d, aligned_2d=None, out=None):
        """aligned_2d: align1(feat_2d) if the caller already started it (a handle with .join()); out: a logical [B,C,H,W]
        channel-last destination (a channel slice of a wider buffer) for the fused tail."""
        b = self.align2(feat_3d)
        a = self.align1(feat_2d) if aligned_2d is None else aligned_2d.join()
        B, C = a.shape[:2]
        if a.dim() == 4 and self._fused_tail_ok(a):        # pool + FCs + softmax + blend in 3 launches
            H, W = a.shape[-2:]
            dst = None
            if out is not None:
                dst = out.permute(0, 2, 3, 1).view(B, H * W, C)           # (raises unless the slice is viewable as rows)
            res = ops.sk_fusion_tail(ops.nhwc_rows(a).view(B, H * W, C), ops.nhwc_rows(b).view(B, H * W, C), 1.0,
                                     self.fc_mid[0].weight, self.fc_out[0].weight, out=dst)
            return out if out is not None else ops.nchw_view(res.view(B, H, W, C))
        w = self._blend_weights((a + b).flatten(2).mean(-1))
        shape = (B, C) + (1,) * (a.dim() - 2)
        return a * w[..., 0].reshape(shape) + b * w[..., 1].reshape(shape)

    def forward_rows(self, rows_2d, rows_3d):
        """'ncm' variant on channel-last point features [B,N,C]."""
        a, b = self.align1.forward_rows(rows_2d), self.align2.forward_rows(rows_3d)
        if self._fused_tail_ok(a):
            return ops.sk_fusion_tail(a.contiguous(), b.contiguous(), 1.0, self.fc_mid[0].weight, self.fc_out[0].weight)
        w = self._blend_weights((a + b).mean(1))
        return a * w[:, None, :, 0] + b * w[:, None, :, 1]


class CLFM(nn.Module):
    """clfm.py:7-40."""

    def __init__(self, in_channels_2d, in_channels_3d, fusion_fn="sk", norm=None):
        super().__init__()
        if fusion_fn != "sk":
            raise NotImplementedError("CLFM: only the 'sk' fusion of the reference configs is built")
        self.interp = FusionAwareInterp(in_channels_3d, k=1, norm=norm)
        self.mlps3d = Conv1dNormRelu(in_channels_2d, in_channels_2d, norm=norm)
        self.fuse2d = SKFusion(in_channels_2d, in_channels_3d, in_channels_2d, "nchw", norm, reduction=2)
        self.fuse3d = SKFusion(in_channels_2d, in_channels_3d, in_channels_3d, "ncm", norm, reduction=2)

    def forward(self, uv, feat_2d, feat_3d, nn_idx=None):
        """uv [B,2,N], feat_2d [B,C2,H,W], feat_3d [B,C3,N] -> (out2d [B,C2,H,W], out3d [B,C3,N])."""
        out2d, out3d_rows = self.forward_rows(uv, feat_2d, ops.rows_of(feat_3d.float()), nn_idx)
        return out2d, ops.cf_of(out3d_rows)

    def forward_rows(self, uv, feat_2d, feat3d_rows, nn_idx=None, par=None, out_2d=None):
        """Same with channel-last point features in and out.  The two directions only read the
        inputs, so `par` (a fork/join helper with .run(fn_a, fn_b)) may execute them concurrently.
        out_2d: where the fused image features go (SKFusion.forward's `out`), inference only."""
        feat_2d = feat_2d.float()

        def to_2d():
            # align1 of the selective-kernel fusion does not depend on the interpolation: beside it, not behind it
            # (forked behind the interpolation kernel: a 324-channel alignment layer is more than one wave of CTAs)
            H, W = feat_2d.shape[-2:]
            nn = ops.nearest_point_2d(uv, H, W) if nn_idx is None else nn_idx
            raw = ops.clfm_interp(uv, nn, feat3d_rows.detach(), self.interp.score_net, H, W)
            early = par.fork(lambda: self.fuse2d.align1(feat_2d), "clfm") if par is not None and hasattr(par, "fork") else None
            return self.fuse2d(feat_2d, self.interp.out_conv(raw), early, out_2d)

        def to_3d():
            sampled = ops.bilinear_sample_rows(feat_2d.detach(), uv)
            return self.fuse3d.forward_rows(self.mlps3d.forward_rows(sampled), feat3d_rows)

        if par is None:
            return to_2d(), to_3d()
        return par.run(to_2d, to_3d)
