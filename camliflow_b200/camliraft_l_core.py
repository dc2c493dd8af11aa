"""3-D (LiDAR) branch of CamLiRAFT (reference models/camliraft_l_core.py): PointConv encoder,
point all-pairs correlation pyramid with k-NN lookup, PointConvDW motion encoder / GRU / flow head."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, tc
from .mlp import Conv1dNormRelu, MLP1d, MLP2d
from .point_conv import PointConv, PointConvDW
from .utils import backwarp_3d, build_pc_pyramid, k_nearest_neighbor, knn_interpolation


class Encoder3D(nn.Module):
    """camliraft_l_core.py:8-37."""

    def __init__(self, n_channels, norm=None, k=16):
        super().__init__()
        self.level0_mlp = MLP1d(3, [n_channels[0], n_channels[0]])
        self.mlps = nn.ModuleList()
        self.convs = nn.ModuleList()
        for c_in, c_out in zip(n_channels[:-1], n_channels[1:]):
            self.mlps.append(MLP1d(c_in, [c_in, c_out]))
            self.convs.append(PointConv(c_out, c_out, norm=norm, k=k))

    def forward(self, xyzs, knn_tables=None):
        """`knn_tables[i]` (optional): the level i -> i+1 neighbour table [B,S,>=k]; it depends on the geometry
        only, so encoders that see the same cloud (fnet / cnet of frame 1) share one search."""
        assert len(xyzs) == len(self.mlps) + 1
        if tc.fused(xyzs[0]):
            # channel-last throughout: every 1x1 layer is one tensor-core GEMM with its activation in the epilogue (the
            # channel-first form runs them as cuDNN Conv1d + separate activation kernels)
            rows = self.level0_mlp.forward_rows(xyzs[0].transpose(1, 2))
            feats = [ops.cf_of(rows)]
            for i, (mlp, conv) in enumerate(zip(self.mlps, self.convs)):
                rows = conv.forward_rows(xyzs[i], mlp.forward_rows(rows), xyzs[i + 1], None if knn_tables is None else knn_tables[i])
                feats.append(ops.cf_of(rows))
            return feats
        feats = [self.level0_mlp(xyzs[0])]
        for i, (mlp, conv) in enumerate(zip(self.mlps, self.convs)):
            feats.append(conv(xyzs[i], mlp(feats[-1]), xyzs[i + 1], None if knn_tables is None else knn_tables[i]))
        return feats

    def neighbor_tables(self, xyzs):
        """The k-NN tables forward() would search for (one launch per level for the whole batch)."""
        return [k_nearest_neighbor(xyzs[i], xyzs[i + 1], conv.k) for i, conv in enumerate(self.convs)]


class Correlation3D(nn.Module):
    """camliraft_l_core.py:40-101; stateful between build and lookup like the reference."""

    def __init__(self, out_channels, k=16):
        super().__init__()
        self.k = k
        self.cost_mlp = MLP2d(4, [out_channels // 4, out_channels // 4], act="relu")
        self.merge = Conv1dNormRelu(out_channels, out_channels)
        self.cost_volume_pyramid = None

    def build_cost_volume_pyramid(self, feat1, feat2, xyzs2, k=3):
        self.cost_volume_pyramid = ops.corr3d_build(feat1, feat2, xyzs2, k)

    def forward(self, xyz1, xyzs2):
        """xyz1 [B,3,n1], xyzs2: the (warped) pyramid of the second cloud -> [B,128,n1]."""
        return ops.cf_of(self.forward_rows(xyz1, xyzs2))

    def forward_rows(self, xyz1, xyzs2):
        """One launch for the 4 x (k-NN + gather + cost MLP + sum over k), then the merge GEMM."""
        if self.k != 16:
            raise NotImplementedError("Correlation3D: the fused lookup is built for k=16 (all reference configs)")
        (w1, b1), (w2, b2) = self.cost_mlp.convs[0].folded(), self.cost_mlp.convs[1].folded()
        costs = ops.corr3d_lookup_rows(xyz1, xyzs2, self.cost_volume_pyramid, w1.contiguous(), b1.contiguous(),
                                       w2.contiguous(), b2.contiguous())
        return self.merge.forward_rows(costs)


class FlowHead3D(nn.Module):
    """camliraft_l_core.py:104-116."""

    def __init__(self, input_dim=128):
        super().__init__()
        self.conv1 = PointConvDW(input_dim, 128, k=32)
        self.conv2 = PointConvDW(128, 64, k=32)
        self.fc = nn.Conv1d(64, 3, kernel_size=1)

    def forward(self, xyz, features, knn_indices=None, cache=None):
        return ops.cf_of(self.forward_rows(xyz, ops.rows_of(features.float()), knn_indices, cache))

    def forward_rows(self, xyz, feat_rows, knn_indices=None, cache=None):
        f = self.conv1.forward_rows(xyz, feat_rows, knn_indices=knn_indices, cache=cache)
        f = self.conv2.forward_rows(xyz, f, knn_indices=knn_indices, cache=cache)
        return tc.linear(f, self.fc.weight, self.fc.bias)


class GRU3D(nn.Module):
    """camliraft_l_core.py:119-134."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.conv_z = PointConvDW(hidden_dim + input_dim, hidden_dim, act=None, k=4)
        self.conv_r = PointConvDW(hidden_dim + input_dim, hidden_dim, act=None, k=4)
        self.conv_q = PointConvDW(hidden_dim + input_dim, hidden_dim, act=None, k=4)

    def forward(self, xyz, h, x, knn_indices=None, cache=None):
        return ops.cf_of(self.forward_rows(xyz, ops.rows_of(h.float()), ops.rows_of(x.float()), knn_indices, cache))

    def forward_rows(self, xyz, h, x, knn_indices=None, cache=None):
        kw = dict(knn_indices=knn_indices, cache=cache)
        hx = torch.cat([h, x], dim=-1)
        if tc.fused(h) and knn_indices is not None and cache is not None and self._mergeable():
            return self._forward_fused(xyz, h, x, hx, knn_indices, cache)
        z = torch.sigmoid(self.conv_z.forward_rows(xyz, hx, **kw))
        r = torch.sigmoid(self.conv_r.forward_rows(xyz, hx, **kw))
        q = torch.tanh(self.conv_q.forward_rows(xyz, torch.cat([r * h, x], dim=-1), **kw))
        return (1 - z) * h + z * q

    def _mergeable(self):
        layers = [m.mlp.convs[0] for m in (self.conv_z, self.conv_r)]
        return all(len(m.mlp.convs) == 1 for m in (self.conv_z, self.conv_r)) and \
            all(isinstance(c.norm_fn, nn.Identity) and c.act is None for c in layers) and self.conv_z.k == self.conv_r.k

    def _forward_fused(self, xyz, h, x, hx, table, cache):
        """The z and r point convolutions read the same input over the same neighbours: ONE 256-wide GEMM, ONE
        gather-max over the concatenated WeightNet caches, then the gate / update kernels of the ConvGRU
        (6 launches instead of 19)."""
        cz, cr = self.conv_z.mlp.convs[0].conv_fn, self.conv_r.mlp.convs[0].conv_fn
        key = tuple((p.data_ptr(), p._version) for p in (cz.weight, cz.bias, cr.weight, cr.bias))
        hit = self.__dict__.get("_zr_linear")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, torch.cat([cz.weight, cr.weight], 0).contiguous(), torch.cat([cz.bias, cr.bias], 0).contiguous())
            self.__dict__["_zr_linear"] = hit
        slot = ("zr", id(self))
        w_zr = cache.get(slot)
        if w_zr is None:                       # WeightNet outputs of both layers side by side: [B,S,k,2H]
            k = self.conv_z.k
            w_zr = torch.cat([ops.pointconv_dw_weights(xyz, xyz, table, k, m.weight_net) for m in (self.conv_z, self.conv_r)], -1)
            cache[slot] = w_zr
        zr = ops.pointconv_dw_gather_max(tc.linear(hx, hit[1], hit[2]), w_zr, table, self.conv_z.k)
        z, rhx = ops.gru_gate_rows(zr, h, x)
        q = self.conv_q.forward_rows(xyz, rhx, knn_indices=table, cache=cache)
        return ops.gru_update_rows(z, h, q)


class MotionEncoder3D(nn.Module):
    """camliraft_l_core.py:137-155."""

    def __init__(self, corr_dim=128):
        super().__init__()
        self.conv_c1 = PointConvDW(corr_dim, corr_dim)
        self.conv_f1 = PointConvDW(3, 32, k=32)
        self.conv_f2 = PointConvDW(32, 16, k=16)
        self.conv = PointConvDW(corr_dim + 16, 128 - 3, k=16)

    def forward(self, xyz, flow, corr, knn_indices, cache=None):
        return ops.cf_of(self.forward_rows(xyz, ops.rows_of(flow.float()), ops.rows_of(corr.float()), knn_indices, cache))

    def forward_rows(self, xyz, flow, corr, knn_indices, cache=None):
        kw = dict(knn_indices=knn_indices, cache=cache)
        c = self.conv_c1.forward_rows(xyz, corr, **kw)
        f = self.conv_f2.forward_rows(xyz, self.conv_f1.forward_rows(xyz, flow, **kw), **kw)
        out = self.conv.forward_rows(xyz, torch.cat([c, f], dim=-1), **kw)
        return torch.cat([out, flow], dim=-1)


class CamLiRAFT_L_Core(nn.Module):
    """LiDAR-only CamLiRAFT-L (camliraft_l_core.py:158-225); also the `branch_3d` of CamLiRAFT."""

    def __init__(self, cfgs):
        super().__init__()
        self.cfgs = cfgs
        self.fnet = Encoder3D(n_channels=[64, 96, 128], norm="batch_norm", k=16)
        self.cnet = Encoder3D(n_channels=[64, 96, 128], norm="batch_norm", k=16)
        self.cnet_aligner = nn.Conv1d(128, 256, kernel_size=1)
        self.correlation = Correlation3D(out_channels=128, k=16)
        self.motion_encoder = MotionEncoder3D(corr_dim=128)
        self.gru = GRU3D(input_dim=128 + 128, hidden_dim=128)
        self.flow_head = FlowHead3D(input_dim=128)

    def forward(self, pc1, pc2):
        xyzs1, xyzs2, _, _ = build_pc_pyramid(pc1, pc2, [4096, 2048, 1024, 512, 256])
        feat1 = self.fnet(xyzs1[:3])[2]
        feat2 = self.fnet(xyzs2[:3])[2]
        featc = self.cnet_aligner(self.cnet(xyzs1[:3])[2])
        xyzs1, xyzs2 = xyzs1[2:], xyzs2[2:]
        xyz1 = xyzs1[0]
        self.correlation.build_cost_volume_pyramid(feat1, feat2, xyzs2)
        h, x = torch.split(featc, [128, 128], dim=1)
        h, x = torch.tanh(h), torch.relu(x)
        nbr = k_nearest_neighbor(xyz1, xyz1, k=32)
        n_iters = self.cfgs.n_iters_train if self.training else self.cfgs.n_iters_eval
        flow = torch.zeros_like(xyz1)
        xyzs2_warp = xyzs2
        h, x = ops.rows_of(h), ops.rows_of(x)
        cache = {}                     # iteration-invariant WeightNet outputs of the PointConvDW layers
        preds = []
        for it in range(n_iters):
            if it > 0:
                flow = flow.detach()
                xyzs2_warp = warp_pyramid(xyz1, xyzs2, flow)
            corr = self.correlation.forward_rows(xyz1, xyzs2_warp)
            motion = self.motion_encoder.forward_rows(xyz1, ops.rows_of(flow), corr, nbr, cache)
            h = self.gru.forward_rows(xyz1, h, torch.cat([x, motion], dim=-1), nbr, cache)
            flow = flow + ops.cf_of(self.flow_head.forward_rows(xyz1, h, nbr, cache))
            preds.append(flow)
        return [knn_interpolation(xyz1, p, pc1, k=3) for p in preds]


def warp_pyramid(xyz1, xyzs2, flow):
    """[backwarp_3d(xyz1, lvl, flow) for lvl in xyzs2] (camliraft_l_core.py:196).  The levels of a
    build_pc_pyramid pyramid are prefixes of one FPS order and every query is warped
    independently, so ONE search over the finest level yields all of them."""
    for coarse in xyzs2[1:]:
        assert coarse.shape[-1] <= xyzs2[0].shape[-1]
    fine = backwarp_3d(xyz1, xyzs2[0], flow)
    return [fine[:, :, :lvl.shape[-1]] for lvl in xyzs2]
