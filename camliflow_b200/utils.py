"""Host-side helpers with the names and argument meaning of the reference's models/utils.py
(grouping, point pyramid, three-NN interpolation, projection, sampling, padding); the heavy
lifting goes to the fused kernels in camliflow_b200.ops."""
import torch
import torch.nn.functional as F

from . import ops
from .csrc import furthest_point_sampling, k_nearest_neighbor  # noqa: F401  (re-exported like models/utils.py:4)


class InputPadder:
    """Replicate-pads H (bottom) and W (both sides) up to a multiple of `x` (utils.py:7-20)."""

    def __init__(self, dims, x=8):
        h, w = dims[-2:]
        ph, pw = (-h) % x, (-w) % x
        self._pad = [pw // 2, pw - pw // 2, 0, ph]

    def pad(self, *inputs):
        return [F.pad(t, self._pad, mode="replicate").contiguous() for t in inputs]

    def unpad(self, t):
        h, w = t.shape[-2:]
        return t[..., self._pad[2]:h - self._pad[3], self._pad[0]:w - self._pad[1]]


def batch_indexing(batched_data, batched_indices, layout="channel_first"):
    """Grouping (utils.py:61-104).  channel_first: data [B,C,N], idx [B,...] -> [B,C,...];
    channel_last: data [B,N,C] (or [B,N]) -> [B,...,C]."""
    if layout == "channel_first":
        return ops.gather_points(batched_data, batched_indices)
    if layout == "channel_last":
        if batched_data.dim() == 2:
            return torch.gather(batched_data, 1, batched_indices.reshape(batched_data.shape[0], -1)) \
                .view(batched_indices.shape)
        return ops.gather_points(batched_data.transpose(1, 2), batched_indices).movedim(1, -1)
    raise ValueError(layout)


def build_pc_pyramid(pc1, pc2, n_samples_list):
    """One FPS launch over both clouds, levels are prefixes of the sample order (utils.py:107-127)."""
    batch_size, _, n_points = pc1.shape
    both = torch.cat([pc1, pc2], dim=0).transpose(1, 2).contiguous()
    sel = furthest_point_sampling(both, max(n_samples_list))
    sel1, sel2 = sel[:batch_size], sel[batch_size:]
    lv0 = torch.arange(n_points, device=pc1.device)[None].expand(batch_size, n_points)
    xyzs1, xyzs2, idx1, idx2 = [pc1], [pc2], [lv0], [lv0]
    full1 = ops.gather_points(pc1, sel1)
    full2 = ops.gather_points(pc2, sel2)
    for n in n_samples_list:
        idx1.append(sel1[:, :n])
        idx2.append(sel2[:, :n])
        xyzs1.append(full1[:, :, :n].contiguous())
        xyzs2.append(full2[:, :, :n].contiguous())
    return xyzs1, xyzs2, idx1, idx2


def knn_interpolation(input_xyz, input_features, query_xyz, k=3):
    """Inverse-distance three-NN interpolation (utils.py:130-146): [B,F,m] -> [B,F,n]."""
    return ops.knn_interpolate(input_xyz, input_features, query_xyz, k)


def backwarp_3d(xyz1, xyz2, flow12, k=3):
    """utils.py:149-159."""
    return xyz2 + ops.knn_interpolate(xyz1 + flow12, -flow12, xyz2, k)


_mesh_cache = {}


def mesh_grid(n, h, w, device, channel_first=True):
    """Pixel-coordinate grid [n,2,h,w] (x then y), cached (utils.py:163-173)."""
    key = (n, h, w, str(device), channel_first)
    if key not in _mesh_cache:
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=device),
                                torch.arange(w, dtype=torch.float32, device=device), indexing="ij")
        g = torch.stack([xs, ys], 0)[None].expand(n, 2, h, w).contiguous()
        _mesh_cache[key] = g if channel_first else g.permute(0, 2, 3, 1).contiguous()
    return _mesh_cache[key]


def backwarp_2d(x, flow12, padding_mode):
    """Bilinear back-warp of an image by a flow field (utils.py:176-188)."""
    n, _, h, w = x.shape
    g = mesh_grid(n, h, w, x.device) + flow12
    gx = 2.0 * g[:, 0] / (w - 1) - 1.0
    gy = 2.0 * g[:, 1] / (h - 1) - 1.0
    return F.grid_sample(x, torch.stack([gx, gy], -1), padding_mode=padding_mode, align_corners=True)


def convex_upsample(flow, mask, scale_factor=8, mask_scale=1.0):
    """Convex-combination upsampling of a 1/s flow field (utils.py:191-204); `mask_scale` multiplies the mask logits
    first (the 0.25 of raft_core.py:197, folded into the kernel)."""
    return ops.convex_upsample(flow, mask, scale_factor, mask_scale)


def project_pc2image(pc, camera_info):
    """Pixel coordinates [B,2,N] of a channel-first cloud (utils.py:234-259)."""
    assert pc.shape[1] == 3
    cx, cy = camera_info["cx"], camera_info["cy"]
    if isinstance(cx, torch.Tensor):
        cx, cy = cx[:, None], cy[:, None]
    mode = camera_info["projection_mode"]
    if mode == "perspective":
        f = camera_info["f"][:, None]
        return torch.stack([cx + (f / pc[:, 2]) * pc[:, 0], cy + (f / pc[:, 2]) * pc[:, 1]], dim=1)
    if mode == "parallel":
        return torch.stack([pc[:, 0] + cx, pc[:, 1] + cy], dim=1)
    raise NotImplementedError(mode)


def grid_sample_wrapper(feat_2d, uv):
    """Bilinear sample of feat_2d [B,C,H,W] at pixel coordinates uv [B,2,N] -> [B,C,N]
    (align_corners, zeros outside; utils.py:262-269); always fp32."""
    return ops.bilinear_sample(feat_2d.float(), uv.float())


def resize_flow2d(flow, target_h, target_w):
    """utils.py:207-214."""
    h, w = flow.shape[2:]
    if (h, w) == (target_h, target_w):
        return flow
    flow = F.interpolate(flow, size=(target_h, target_w), mode="bilinear", align_corners=True)
    return torch.stack([flow[:, 0] * (target_w / w), flow[:, 1] * (target_h / h)], dim=1)      # (no host tensor: capturable)


def resize_to_64x(inputs, target, x=64):
    """utils.py:217-231."""
    n, c, h, w = inputs.shape
    if h % x == 0 and w % x == 0:
        return inputs, target
    rh, rw = -(-h // x) * x, -(-w // x) * x
    inputs = F.interpolate(inputs, size=(rh, rw), mode="bilinear", align_corners=True)
    if target is not None:
        target = F.interpolate(target, size=(rh, rw), mode="bilinear", align_corners=True)
        target = torch.cat([torch.stack([target[:, 0] * (rw / w), target[:, 1] * (rh / h)], dim=1), target[:, 2:]], dim=1)
    return inputs, target


def disp2pc(disp, baseline, f, cx, cy, flow=None):
    """Dense point cloud [H,W,3] of a disparity map [H,W] (reference utils.py:319-339, there in numpy on the host):
    depth = baseline * f / (disp + 1e-5), back-projected through the pinhole model, optionally at flow-displaced
    pixel positions.  Runs on whatever device `disp` lives on (the first stage of the on-GPU input pipeline)."""
    h, w = disp.shape
    depth = baseline * f / (disp + 1e-5)
    xx = torch.arange(w, dtype=torch.float32, device=disp.device)[None, :].expand(h, w)
    yy = torch.arange(h, dtype=torch.float32, device=disp.device)[:, None].expand(h, w)
    if flow is not None:
        xx, yy = xx + flow[..., 0], yy + flow[..., 1]
    return torch.stack([(xx - cx) * depth / f, (yy - cy) * depth / f, depth], dim=-1)


def densify_flow_3d(pc1, flow_3d, disp1, baseline, f, cx, cy):
    """Scene flow of every pixel of a disparity map from the sparse prediction (kitti_submission.py:89-93): the
    dense cloud of `disp1` takes the three-NN inverse-distance interpolation of `flow_3d` [3,N] given at `pc1`
    [3,N] (~466 k queries against 8192 points on KITTI: one fused search + blend launch).  Returns the dense
    cloud [3,H*W] and its flow [3,H*W]."""
    dense = disp2pc(disp1, baseline, f, cx, cy).reshape(-1, 3).t().contiguous()
    flow = knn_interpolation(pc1[None], flow_3d[None], dense[None])[0]
    return dense, flow
