"""PointConv / PointConvDW with the nn.Module surface and parameter names of the reference's
models/point_conv.py (`weight_net.convs.N.conv_fn`, `linear`, `norm_fn`, `mlp.convs.0.conv_fn`),
computed by the fused grouping kernels of camliflow_b200.ops."""
import torch
import torch.nn as nn

from . import ops, tc
from .csrc import k_nearest_neighbor
from .mlp import MLP1d, MLP2d, LayerNormCF, _ACTS


def _neighbor_table(xyz, sampled_xyz, knn_indices, k):
    """A precomputed (possibly wider) neighbour table, or a fresh search (point_conv.py:49-55)."""
    if knn_indices is None:
        return k_nearest_neighbor(xyz, sampled_xyz, k)
    assert knn_indices.shape[:2] == torch.Size([sampled_xyz.shape[0], sampled_xyz.shape[-1]])
    assert knn_indices.shape[2] >= k
    return knn_indices




class PointConv(nn.Module):
    """Set abstraction: k-NN grouping, WeightNet(3->8->16) on the offsets, per-centroid
    [16 x k]·[k x (C+3)] contraction, Linear, norm, activation (point_conv.py:7-70)."""

    def __init__(self, in_channels, out_channels, norm=None, act="leaky_relu", k=16):
        super().__init__()
        self.k = k
        self.weight_net = MLP2d(3, [8, 16], act=act)
        self.linear = nn.Linear(16 * (in_channels + 3), out_channels)
        if norm == "batch_norm":
            self.norm_fn = nn.BatchNorm1d(out_channels, affine=True)
        elif norm == "instance_norm":
            self.norm_fn = nn.InstanceNorm1d(out_channels, affine=True)
        elif norm == "layer_norm":
            self.norm_fn = LayerNormCF(out_channels, 1)
        elif norm is None:
            self.norm_fn = nn.Identity()
        else:
            raise NotImplementedError("Unknown normalization function: %s" % norm)
        if act not in ("relu", "leaky_relu", None):
            raise NotImplementedError("Unknown activation function: %s" % act)
        self.act_fn = _ACTS[act]()
        self.act = act
        self._slope = {"relu": 0.0, "leaky_relu": 0.1, None: 1.0}[act]

    def forward_rows(self, xyz, feat_rows, sampled_xyz=None, knn_indices=None):
        """Channel-last variant: feat_rows [B,N,C] -> rows [B,S,O] (inference fast path of the encoders)."""
        return self.forward(xyz, None, sampled_xyz, knn_indices, feat_rows=feat_rows).transpose(1, 2)

    def forward(self, xyz, features, sampled_xyz=None, knn_indices=None, feat_rows=None):
        """xyz [B,3,N], features [B,C,N], sampled_xyz [B,3,S] -> [B,O,S]."""
        if sampled_xyz is None:
            sampled_xyz = xyz
        table = _neighbor_table(xyz, sampled_xyz, knn_indices, self.k)
        if feat_rows is not None:
            rows = torch.cat([xyz.transpose(1, 2), feat_rows], dim=-1)                 # [B,N,3+C]
        else:
            rows = ops.rows_of(torch.cat([xyz, features], dim=1))                      # [B,N,3+C]
        grouped = ops.pointconv_group(rows, sampled_xyz, table, self.k, self.weight_net, self._slope)
        n = self.norm_fn
        if tc.fused(grouped) and (isinstance(n, nn.Identity) or (isinstance(n, nn.BatchNorm1d) and not n.training)):
            # Linear + BatchNorm (running statistics, folded) + activation: one tensor-core kernel
            return tc.linear(grouped, self.linear.weight, self.linear.bias, self.act, 0.1, bn=n).transpose(1, 2)
        out = tc.module_train(self.linear, grouped)                                # training: grad.DenseFn
        out = (self.linear(grouped) if out is None else out).transpose(1, 2)       # [B,O,S]
        return self.act_fn(self.norm_fn(out))


class PointConvDW(nn.Module):
    """Depth-wise point convolution: MLP1d(C->O), grouping, WeightNet(3->8->32->O, relu) on the
    offsets, product, max over the k neighbours (point_conv.py:102-130)."""

    def __init__(self, in_channels, out_channels, norm=None, act="leaky_relu", k=16):
        super().__init__()
        self.k = k
        self.mlp = MLP1d(in_channels, [out_channels], norm, act)
        self.weight_net = MLP2d(3, [8, 32, out_channels], act="relu")

    def forward(self, xyz, features, sampled_xyz=None, knn_indices=None, cache=None):
        """xyz [B,3,N], features [B,C,N] -> [B,O,S] (a transposed view of channel-last storage)."""
        return ops.cf_of(self.forward_rows(xyz, ops.rows_of(features), sampled_xyz, knn_indices, cache))

    def forward_rows(self, xyz, feat_rows, sampled_xyz=None, knn_indices=None, cache=None):
        """Channel-last fast path: feat_rows [B,N,C] -> [B,S,O].  `cache` (a dict owned by the caller)
        keeps the WeightNet output, which depends only on (xyz, neighbour table, layer parameters),
        across the iterations of a recurrent loop."""
        if sampled_xyz is None:
            sampled_xyz = xyz
        table = _neighbor_table(xyz, sampled_xyz, knn_indices, self.k)
        weights = cache.get(id(self)) if cache is not None else None
        if weights is None:
            weights = ops.pointconv_dw_weights(xyz, sampled_xyz, table, self.k, self.weight_net)
            if cache is not None:
                cache[id(self)] = weights
        return ops.pointconv_dw_gather_max(self.mlp.forward_rows(feat_rows), weights, table, self.k)
