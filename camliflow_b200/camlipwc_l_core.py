"""3-D PWC branch (reference models/camlipwc_l_core.py): PointConv feature pyramid, PointPWC-style
learnable cost volume, PointConv flow estimator."""
import torch
import torch.nn as nn

from . import ops, tc
from .mlp import Conv1dNormRelu, MLP1d, MLP2d
from .point_conv import PointConv
from .utils import backwarp_3d, k_nearest_neighbor, knn_interpolation

PYRAMID_CHANNELS_3D = [16, 32, 64, 96, 128, 192]


class FeaturePyramid3D(nn.Module):
    """camlipwc_l_core.py:8-36."""

    def __init__(self, n_channels, norm=None, k=16):
        super().__init__()
        self.level0_mlp = MLP1d(3, [n_channels[0], n_channels[0]])
        self.pyramid_mlps = nn.ModuleList()
        self.pyramid_convs = nn.ModuleList()
        for c_in, c_out in zip(n_channels[:-1], n_channels[1:]):
            self.pyramid_mlps.append(MLP1d(c_in, [c_in, c_out]))
            self.pyramid_convs.append(PointConv(c_out, c_out, norm=norm, k=k))

    def forward(self, xyzs):
        assert len(xyzs) == len(self.pyramid_mlps) + 1
        if tc.fused(xyzs[0]):          # channel-last throughout: every 1x1 layer is one tensor-core GEMM (see Encoder3D)
            rows = self.level0_mlp.forward_rows(xyzs[0].transpose(1, 2))
            feats = [ops.cf_of(rows)]
            for i, (mlp, conv) in enumerate(zip(self.pyramid_mlps, self.pyramid_convs)):
                rows = conv.forward_rows(xyzs[i], mlp.forward_rows(rows), xyzs[i + 1])
                feats.append(ops.cf_of(rows))
            return feats
        feats = [self.level0_mlp(xyzs[0])]
        for i, (mlp, conv) in enumerate(zip(self.pyramid_mlps, self.pyramid_convs)):
            feats.append(conv(xyzs[i], mlp(feats[-1]), xyzs[i + 1]))
        return feats


class Correlation3D(nn.Module):
    """Learnable point cost volume (camlipwc_l_core.py:39-106): point-to-point costs of every xyz1
    point against its k nearest (warped) xyz2 points, aggregated patch-to-point with WeightNet2 and
    then point-to-patch over the xyz1 neighbourhood with WeightNet1."""

    def __init__(self, in_channels, out_channels, align_channels=None, k=16):
        super().__init__()
        self.k = k
        self.cost_mlp = MLP2d(3 + 2 * in_channels, [out_channels, out_channels], act="leaky_relu")
        self.weight_net1 = MLP2d(3, [8, 8, out_channels], act="relu")
        self.weight_net2 = MLP2d(3, [8, 8, out_channels], act="relu")
        self.feat_aligner = Conv1dNormRelu(out_channels, align_channels) if align_channels is not None else nn.Identity()

    def forward(self, xyz1, feat1, xyz2, feat2, knn_indices_1in1=None):
        B, C, n = feat1.shape
        idx12 = k_nearest_neighbor(input_xyz=xyz2, query_xyz=xyz1, k=self.k)
        off2 = ops.gather_points(xyz2, idx12) - xyz1[:, :, :, None]
        pair = torch.cat([feat1[:, :, :, None].expand(B, C, n, self.k), ops.gather_points(feat2, idx12), off2], dim=1)
        p2n = torch.sum(self.weight_net2(off2) * self.cost_mlp(pair), dim=3)
        if knn_indices_1in1 is None:
            idx11 = k_nearest_neighbor(input_xyz=xyz1, query_xyz=xyz1, k=self.k)
        else:
            assert knn_indices_1in1.shape[:2] == torch.Size([B, n]) and knn_indices_1in1.shape[2] >= self.k
            idx11 = knn_indices_1in1[:, :, :self.k]
        off1 = ops.gather_points(xyz1, idx11) - xyz1[:, :, :, None]
        n2n = torch.sum(self.weight_net1(off1) * ops.gather_points(p2n, idx11), dim=3)
        return self.feat_aligner(n2n)


class FlowEstimator3D(nn.Module):
    """camlipwc_l_core.py:109-139."""

    def __init__(self, n_channels, norm=None, conv_last=True, k=16):
        super().__init__()
        self.point_conv1 = PointConv(in_channels=n_channels[0], out_channels=n_channels[1], norm=norm, k=k)
        self.point_conv2 = PointConv(in_channels=n_channels[1], out_channels=n_channels[2], norm=norm, k=k)
        self.mlp = MLP1d(n_channels[2], [n_channels[2], n_channels[3]])
        self.flow_feat_dim = n_channels[3]
        self.conv_last = nn.Conv1d(n_channels[3], 3, kernel_size=1) if conv_last else None

    def forward(self, xyz, feat, knn_indices):
        feat = self.point_conv1(xyz, feat, knn_indices=knn_indices)
        feat = self.mlp(self.point_conv2(xyz, feat, knn_indices=knn_indices))
        return feat if self.conv_last is None else (feat, self.conv_last(feat))


class CamLiPWC_L_Core(nn.Module):
    """LiDAR-only CamLiPWC-L (camlipwc_l_core.py:142-210)."""

    def __init__(self, cfgs):
        super().__init__()
        self.cfgs = cfgs
        self.feature_pyramid = FeaturePyramid3D(n_channels=PYRAMID_CHANNELS_3D, norm=cfgs.norm.feature_pyramid)
        self.correlations = nn.ModuleList([nn.Identity()] + [Correlation3D(c, c, 64) for c in PYRAMID_CHANNELS_3D[1:]])
        self.pyramid_feat_aligners = nn.ModuleList(
            [nn.Identity()] + [Conv1dNormRelu(c, 64) for c in PYRAMID_CHANNELS_3D[1:]])
        self.flow_estimator = FlowEstimator3D(n_channels=[64 + 64 + 3, 128, 128, 64], norm=cfgs.norm.flow_estimator)

    def encode(self, xyzs):
        return self.feature_pyramid(xyzs)

    def decode(self, xyzs1, xyzs2, feats1_3d, feats2_3d):
        flows = []
        top = len(xyzs1) - 1
        for level in range(top, 0, -1):
            xyz1, xyz2 = xyzs1[level], xyzs2[level]
            knn1 = k_nearest_neighbor(xyz1, xyz1, k=16)
            if level == top:
                last = torch.zeros_like(xyz1)
                xyz2_warp = xyz2
            else:
                last = knn_interpolation(xyzs1[level + 1], flows[-1], xyz1)
                xyz2_warp = backwarp_3d(xyz1, xyz2, last)
            x = torch.cat([self.pyramid_feat_aligners[level](feats1_3d[level]),
                           self.correlations[level](xyz1, feats1_3d[level], xyz2_warp, feats2_3d[level], knn1),
                           last], dim=1)
            flows.append(last + self.flow_estimator(xyz1, x, knn1)[1])
        flows = [f.float() for f in flows][::-1]
        return [knn_interpolation(xyzs1[i + 1], f, xyzs1[i]) for i, f in enumerate(flows)]
