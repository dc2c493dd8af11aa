"""CLFM: bidirectional camera<->LiDAR fusion (reference models/clfm.py), SK fusion only --
the variant every reference config uses.  3D->2D: every pixel takes the features of its
nearest projected point, gated by a ScoreNet of the pixel->point offset; 2D->3D: bilinear
sample of the image features at the projected points."""
import torch
import torch.nn as nn

from . import ops
from .mlp import Conv1dNormRelu, Conv2dNormRelu
from .utils import grid_sample_wrapper, mesh_grid


class FusionAwareInterp(nn.Module):
    """clfm.py:43-79 (k = 1)."""

    def __init__(self, n_channels_3d, k=1, norm=None):
        super().__init__()
        if k != 1:
            raise NotImplementedError("FusionAwareInterp: the reference configs only use k=1")
        self.k = k
        self.out_conv = Conv2dNormRelu(n_channels_3d, n_channels_3d, norm=norm)
        self.score_net = nn.Sequential(Conv2dNormRelu(3, 16), Conv2dNormRelu(16, n_channels_3d, act="sigmoid"))

    def forward(self, uv, feat_2d, feat_3d):
        B, _, H, W = feat_2d.shape
        nn_idx = ops.nearest_point_2d(uv, H, W)                                   # [B,HW]
        grid = mesh_grid(B, H, W, uv.device).reshape(B, 2, -1)
        off = ops.gather_points(uv, nn_idx) - grid                                # [B,2,HW]
        score_in = torch.cat([off, torch.linalg.norm(off, dim=1, keepdim=True)], dim=1)
        score = self.score_net(score_in.view(B, 3, H, W))
        final = score * ops.gather_points(feat_3d, nn_idx).view(B, -1, H, W)
        return self.out_conv(final)


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

    def forward(self, feat_2d, feat_3d):
        a, b = self.align1(feat_2d), self.align2(feat_3d)
        B, C = a.shape[:2]
        pooled = (a + b).flatten(2).mean(-1)
        w = torch.softmax(self.fc_out(self.fc_mid(pooled)).view(B, C, 2), dim=-1)
        shape = (B, C) + (1,) * (a.dim() - 2)
        return a * w[..., 0].reshape(shape) + b * w[..., 1].reshape(shape)


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

    def forward(self, uv, feat_2d, feat_3d):
        feat_2d, feat_3d = feat_2d.float(), feat_3d.float()
        interp = self.interp(uv, feat_2d.detach(), feat_3d.detach())
        out2d = self.fuse2d(feat_2d, interp)
        sampled = grid_sample_wrapper(feat_2d.detach(), uv)
        out3d = self.fuse3d(self.mlps3d(sampled.detach()), feat_3d)
        return out2d, out3d
