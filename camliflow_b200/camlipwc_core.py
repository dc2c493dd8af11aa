"""Fused coarse-to-fine decoder of CamLiPWC (reference models/camlipwc_core.py:124-237): a 2-D PWC
branch and a 3-D PointPWC branch refined level by level (5 levels), fused by CLFM at the pyramid,
the correlation and the estimator."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import tc
from .camlipwc_l_core import PYRAMID_CHANNELS_3D, Correlation3D, FeaturePyramid3D, FlowEstimator3D
from .clfm import CLFM
from .csrc import correlation2d
from .mlp import Conv1dNormRelu, Conv2dNormRelu
from .pwc_core import (PYRAMID_CHANNELS_2D, ContextNetwork2D, FeaturePyramid2D, FlowEstimatorDense2D,
                       FlowEstimatorLite2D, finish_flows_2d, up_mask_head, upsample2x)
from .utils import backwarp_2d, backwarp_3d, k_nearest_neighbor, knn_interpolation, project_pc2image


def _per_level(make, channels):
    return nn.ModuleList([nn.Identity()] + [make(c) for c in channels])


class CamLiPWC_Core(nn.Module):
    def __init__(self, cfgs2d, cfgs3d, cfgs):
        super().__init__()
        self.cfgs, self.cfgs2d, self.cfgs3d = cfgs, cfgs2d, cfgs3d
        corr2d = (2 * cfgs2d.max_displacement + 1) ** 2
        c2, c3 = PYRAMID_CHANNELS_2D[2:], PYRAMID_CHANNELS_3D[1:]          # per decoded level: 32,64,96,128,192

        self.branch_2d_fnet = FeaturePyramid2D(PYRAMID_CHANNELS_2D, norm=cfgs2d.norm.feature_pyramid)
        self.branch_2d_fnet_aligners = _per_level(lambda c: Conv2dNormRelu(c, 64), c2)
        est2d = FlowEstimatorLite2D if cfgs2d.lite_estimator else FlowEstimatorDense2D
        self.branch_2d_flow_estimator = est2d([64 + corr2d + 2 + 32, 128, 128, 96, 64, 32], norm=cfgs2d.norm.flow_estimator,
                                              conv_last=not cfgs.fuse_estimator)
        self.branch_2d_context_network = ContextNetwork2D(
            [self.branch_2d_flow_estimator.flow_feat_dim + 2, 128, 128, 128, 96, 64, 32], dilations=[1, 2, 4, 8, 16, 1],
            norm=cfgs2d.norm.context_network)
        self.branch_2d_up_mask_head = up_mask_head()

        self.branch_3d_fnet = FeaturePyramid3D(n_channels=PYRAMID_CHANNELS_3D, norm=cfgs3d.norm.feature_pyramid, k=cfgs3d.k)
        self.branch_3d_fnet_aligners = _per_level(lambda c: Conv1dNormRelu(c, 64), c3)
        self.branch_3d_correlations = _per_level(lambda c: Correlation3D(c, c, k=cfgs3d.k), c3)
        self.branch_3d_correlation_aligners = _per_level(lambda c: Conv1dNormRelu(c, 64), c3)
        self.branch_3d_flow_estimator = FlowEstimator3D([64 + 64 + 3 + 64, 128, 128, 64], cfgs3d.norm.flow_estimator,
                                                        conv_last=not cfgs.fuse_estimator, k=cfgs3d.k)
        if cfgs.fuse_pyramid:
            self.pyramid_clfms = _per_level(lambda c: CLFM(c, c, norm=cfgs2d.norm.feature_pyramid), c2)
        if cfgs.fuse_correlation:
            self.corr_clfms = _per_level(lambda c: CLFM(corr2d, c), c3)
        if cfgs.fuse_estimator:
            d2, d3 = self.branch_2d_flow_estimator.flow_feat_dim, self.branch_3d_flow_estimator.flow_feat_dim
            self.estimator_clfm = CLFM(d2, d3)
            self.branch_2d_conv_last = nn.Conv2d(d2, 2, kernel_size=3, stride=1, padding=1)
            self.branch_3d_conv_last = nn.Conv1d(d3, 3, kernel_size=1)

    def encode(self, image, xyzs):
        return self.branch_2d_fnet(image), self.branch_3d_fnet(xyzs)

    def decode(self, xyzs1, xyzs2, feats1_2d, feats2_2d, feats1_3d, feats2_3d, camera_info):
        assert len(xyzs1) == len(xyzs2) == len(feats1_2d) == len(feats2_2d) == len(feats1_3d) == len(feats2_3d)
        cfgs, top = self.cfgs, len(xyzs1) - 1
        flows_2d, flows_3d, feats_2d, feats_3d = [], [], [], []
        for level in range(top, 0, -1):
            xyz1, xyz2 = xyzs1[level], xyzs2[level]
            f1_2d, f2_2d, f1_3d, f2_3d = feats1_2d[level], feats2_2d[level], feats1_3d[level], feats2_3d[level]
            B, _, H, W = f1_2d.shape
            sx = (W - 1) / (camera_info["sensor_w"] - 1)
            sy = (H - 1) / (camera_info["sensor_h"] - 1)
            uv1, uv2 = project_pc2image(xyz1, camera_info), project_pc2image(xyz2, camera_info)
            uv1 = torch.stack([uv1[:, 0] * sx, uv1[:, 1] * sy], dim=1)
            uv2 = torch.stack([uv2[:, 0] * sx, uv2[:, 1] * sy], dim=1)
            knn1 = k_nearest_neighbor(xyz1, xyz1, k=self.cfgs3d.k)

            if cfgs.fuse_pyramid:
                f1_2d, f1_3d = self.pyramid_clfms[level](uv1, f1_2d, f1_3d)
                f2_2d, f2_3d = self.pyramid_clfms[level](uv2, f2_2d, f2_3d)

            if level == top:
                last_flow_2d = torch.zeros((B, 2, H, W), dtype=uv1.dtype, device=uv1.device)
                last_feat_2d = torch.zeros((B, 32, H, W), dtype=uv1.dtype, device=uv1.device)
                last_flow_3d = torch.zeros_like(xyz1)
                last_feat_3d = torch.zeros((B, 64, xyz1.shape[-1]), dtype=uv1.dtype, device=uv1.device)
                xyz2_warp, f2_2d_warp = xyz2, f2_2d
            else:
                last_flow_2d = upsample2x(flows_2d[-1], 2.0)
                last_feat_2d = upsample2x(feats_2d[-1])
                up = knn_interpolation(xyzs1[level + 1], torch.cat([flows_3d[-1], feats_3d[-1]], dim=1), xyz1)
                last_flow_3d, last_feat_3d = up[:, :3].contiguous(), up[:, 3:]
                f2_2d_warp = backwarp_2d(f2_2d, last_flow_2d, padding_mode="border")
                xyz2_warp = backwarp_3d(xyz1, xyz2, last_flow_3d)

            corr_3d = self.branch_3d_correlations[level](xyz1, f1_3d, xyz2_warp, f2_3d, knn1)
            corr_2d = F.leaky_relu(correlation2d(f1_2d, f2_2d_warp, self.cfgs2d.max_displacement), 0.1)
            if cfgs.fuse_correlation:
                corr_2d, corr_3d = self.corr_clfms[level](uv1, corr_2d, corr_3d)

            x_2d = torch.cat([corr_2d, self.branch_2d_fnet_aligners[level](f1_2d), last_flow_2d, last_feat_2d], dim=1)
            x_3d = torch.cat([self.branch_3d_correlation_aligners[level](corr_3d),
                              self.branch_3d_fnet_aligners[level](f1_3d), last_flow_3d, last_feat_3d], dim=1)
            if cfgs.fuse_estimator:
                feat_2d = self.branch_2d_flow_estimator(x_2d)
                feat_3d = self.branch_3d_flow_estimator(xyz1, x_3d, knn1)
                feat_2d, feat_3d = self.estimator_clfm(uv1, feat_2d, feat_3d)
                delta_2d, delta_3d = tc.conv2d(feat_2d, self.branch_2d_conv_last), self.branch_3d_conv_last(feat_3d)
            else:
                feat_2d, delta_2d = self.branch_2d_flow_estimator(x_2d)
                feat_3d, delta_3d = self.branch_3d_flow_estimator(xyz1, x_3d, knn1)

            flow_2d = last_flow_2d + delta_2d
            flow_3d = last_flow_3d + delta_3d
            feat_2d, delta_2d = self.branch_2d_context_network(torch.cat([feat_2d, flow_2d], dim=1))
            flow_2d = torch.clip(delta_2d + flow_2d, min=-1000, max=1000)
            flow_3d = torch.clip(flow_3d, min=-100, max=100)
            flows_2d.append(flow_2d)
            flows_3d.append(flow_3d)
            feats_2d.append(feat_2d)
            feats_3d.append(feat_3d)

        flows_2d = finish_flows_2d(flows_2d, feat_2d, self.branch_2d_up_mask_head)
        flows_3d = [f.float() for f in flows_3d][::-1]
        flows_3d = [knn_interpolation(xyzs1[i + 1], f, xyzs1[i]) for i, f in enumerate(flows_3d)]
        return flows_2d, flows_3d
