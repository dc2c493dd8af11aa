"""2-D PWC branch (reference models/pwc_core.py): strided residual feature pyramid, local cost
volume (the sm_100a kernel behind camliflow_b200.csrc.correlation2d), dense / lite flow estimators
and the dilated context network.  Parameter names follow the reference."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import tc
from .csrc import correlation2d
from .mlp import Conv2dNormRelu
from .utils import backwarp_2d, convex_upsample

PYRAMID_CHANNELS_2D = [3, 16, 32, 64, 96, 128, 192]


class ResidualBlock(nn.Module):
    """pwc_core.py:9-29."""

    def __init__(self, in_channels, out_channels, down_sample=True, norm=None):
        super().__init__()
        s = 2 if down_sample else 1
        self.down0 = Conv2dNormRelu(in_channels, out_channels, stride=2, norm=norm, act=None) if down_sample \
            else nn.Identity()
        self.conv0 = Conv2dNormRelu(in_channels, out_channels, kernel_size=3, stride=s, padding=1, norm=norm)
        self.conv1 = Conv2dNormRelu(out_channels, out_channels, kernel_size=3, stride=1, padding=1, norm=norm, act=None)

    def forward(self, x):
        if tc.fused(x) and not isinstance(self.down0, nn.Identity) and self.conv0._foldable():
            # three kernels: the strided 1x1 shortcut, the strided 3x3, and the second 3x3 with the shortcut and the
            # leaky ReLU in its epilogue (BatchNorms folded)
            r = tc.conv2d(x, self.down0.conv_fn, None, bn=self.down0.norm_fn)
            y = tc.conv2d(x, self.conv0.conv_fn, "leaky_relu", 0.1, bn=self.conv0.norm_fn)
            return tc.conv2d(y, self.conv1.conv_fn, "leaky_relu", 0.1, bn=self.conv1.norm_fn, residual=r)
        return F.leaky_relu(self.conv1(self.conv0(x)) + self.down0(x), 0.1)


class FeaturePyramid2D(nn.Module):
    """pwc_core.py:31-44: one residual block per level, each halving the resolution."""

    def __init__(self, n_channels, norm=None):
        super().__init__()
        self.pyramid_convs = nn.ModuleList(ResidualBlock(i, o, norm=norm) for i, o in zip(n_channels[:-1], n_channels[1:]))

    def forward(self, x):
        outputs = []
        for block in self.pyramid_convs:
            x = block(x)
            outputs.append(x)
        return outputs


class _FlowEstimator2D(nn.Module):
    DENSE = True

    def __init__(self, n_channels, norm=None, conv_last=True):
        super().__init__()
        c = n_channels
        if self.DENSE:       # every layer sees all earlier outputs and the input (pwc_core.py:73-125)
            ins = [sum(c[:i + 1]) for i in range(5)]
            self.flow_feat_dim = sum(c)
        else:                # lite: each layer sees the two previous outputs (pwc_core.py:47-71)
            ins = [c[0], c[1], c[1] + c[2], c[2] + c[3], c[3] + c[4]]
            self.flow_feat_dim = c[4] + c[5]
        for i in range(5):
            setattr(self, "conv%d" % (i + 1), Conv2dNormRelu(ins[i], c[i + 1], kernel_size=3, padding=1, norm=norm))
        self.conv_last = nn.Conv2d(self.flow_feat_dim, 2, kernel_size=3, stride=1, padding=1) if conv_last else None

    def forward(self, x):
        if self.DENSE:
            for i in range(5):
                x = torch.cat([getattr(self, "conv%d" % (i + 1))(x), x], dim=1)
            feat = x
        else:
            x1 = self.conv1(x)
            x2 = self.conv2(x1)
            x3 = self.conv3(torch.cat([x1, x2], dim=1))
            x4 = self.conv4(torch.cat([x2, x3], dim=1))
            x5 = self.conv5(torch.cat([x3, x4], dim=1))
            feat = torch.cat([x4, x5], dim=1)
        return feat if self.conv_last is None else (feat, tc.conv2d(feat, self.conv_last))


class FlowEstimatorDense2D(_FlowEstimator2D):
    DENSE = True


class FlowEstimatorLite2D(_FlowEstimator2D):
    DENSE = False


class ContextNetwork2D(nn.Module):
    """pwc_core.py:128-141."""

    def __init__(self, n_channels, dilations, norm=None):
        super().__init__()
        self.convs = nn.ModuleList(
            Conv2dNormRelu(i, o, kernel_size=3, padding=d, dilation=d, norm=norm)
            for i, o, d in zip(n_channels[:-1], n_channels[1:], dilations))
        self.conv_last = nn.Conv2d(n_channels[-1], 2, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        for conv in self.convs:
            x = conv(x)
        return x, tc.conv2d(x, self.conv_last)


def up_mask_head():
    return nn.Sequential(nn.Conv2d(32, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
                         nn.Conv2d(64, 4 * 4 * 9, kernel_size=1, stride=1, padding=0))


def upsample2x(t, scale=1.0):
    return F.interpolate(t * scale if scale != 1.0 else t, scale_factor=2, mode="bilinear", align_corners=True)


def finish_flows_2d(flows_2d, flow_feat, mask_head):
    """Finest level by convex up-sampling (x4), the others bilinearly (pwc_core.py:218-224)."""
    flows = [f.float() for f in flows_2d][::-1]
    mask = tc.conv2d(tc.conv2d(flow_feat, mask_head[0], "relu"), mask_head[2]) if tc.fused(flow_feat) else mask_head(flow_feat)
    flows[0] = convex_upsample(flows[0], mask, scale_factor=4)
    for i in range(1, len(flows)):
        flows[i] = F.interpolate(flows[i] * 4, scale_factor=4, mode="bilinear", align_corners=True)
    return flows


class PWCCore(nn.Module):
    """Image-only PWC-Net (pwc_core.py:144-225)."""

    def __init__(self, cfgs):
        super().__init__()
        self.cfgs = cfgs
        corr_channels = (cfgs.max_displacement * 2 + 1) ** 2
        self.feature_pyramid = FeaturePyramid2D(PYRAMID_CHANNELS_2D, norm=cfgs.norm.feature_pyramid)
        self.pyramid_feature_aligners = nn.ModuleList(
            [nn.Identity()] + [Conv2dNormRelu(c, 64) for c in PYRAMID_CHANNELS_2D[2:]])
        est = FlowEstimatorLite2D if cfgs.lite_estimator else FlowEstimatorDense2D
        self.flow_estimator = est([64 + corr_channels + 2, 128, 128, 96, 64, 32], norm=cfgs.norm.flow_estimator)
        self.context_network = ContextNetwork2D([self.flow_estimator.flow_feat_dim + 2, 128, 128, 128, 96, 64, 32],
                                                [1, 2, 4, 8, 16, 1], norm=cfgs.norm.context_network)
        self.up_mask_head = up_mask_head()

    def encode(self, image):
        return self.feature_pyramid(image)

    def decode(self, feats1_2d, feats2_2d):
        assert len(feats1_2d) == len(feats2_2d)
        flows = []
        for level in range(len(feats1_2d) - 1, 0, -1):
            f1, f2 = feats1_2d[level], feats2_2d[level]
            if not flows:
                last = torch.zeros((f1.shape[0], 2) + f1.shape[2:], dtype=f1.dtype, device=f1.device)
            else:
                last = upsample2x(flows[-1], 2.0)
                f2 = backwarp_2d(f2, last, padding_mode="border")
            corr = F.leaky_relu(correlation2d(f1, f2, self.cfgs.max_displacement), 0.1)
            feat, delta = self.flow_estimator(torch.cat([corr, self.pyramid_feature_aligners[level](f1), last], dim=1))
            flow = delta + last
            feat, delta = self.context_network(torch.cat([feat, flow], dim=1))
            flows.append(delta + flow)
        return finish_flows_2d(flows, feat, self.up_mask_head)
