"""Inverse-depth scaling: perspective <-> parallel projection of a point cloud
(reference models/ids.py).  Pointwise; stays in PyTorch."""
import torch


def _ratios(persp, paral):
    sw = (paral["sensor_w"] - 1) / (persp["sensor_w"] - 1)
    sh = (paral["sensor_h"] - 1) / (persp["sensor_h"] - 1)
    return sw, sh


def persp2paral(xyz, perspect_camera_info, parallel_camera_info):
    """ids.py:4-33: (x,y,z) -> (u*sw - (W'-1)/2, v*sh - (H'-1)/2, (f log z + 1) * min(sw,sh))."""
    f = perspect_camera_info["f"][:, None]
    cx, cy = perspect_camera_info["cx"][:, None], perspect_camera_info["cy"][:, None]
    sw, sh = _ratios(perspect_camera_info, parallel_camera_info)
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    return torch.stack([
        (cx + (f / z) * x) * sw - (parallel_camera_info["sensor_w"] - 1) / 2,
        (cy + (f / z) * y) * sh - (parallel_camera_info["sensor_h"] - 1) / 2,
        (f * torch.log(z) + 1) * min(sw, sh),
    ], dim=1)


def paral2persp(xyz, perspect_camera_info, parallel_camera_info):
    """ids.py:36-67: the inverse map."""
    f = perspect_camera_info["f"][:, None]
    cx, cy = perspect_camera_info["cx"][:, None], perspect_camera_info["cy"][:, None]
    sw, sh = _ratios(perspect_camera_info, parallel_camera_info)
    u = (xyz[:, 0] + (parallel_camera_info["sensor_w"] - 1) / 2) / sw
    v = (xyz[:, 1] + (parallel_camera_info["sensor_h"] - 1) / 2) / sh
    z = torch.exp((xyz[:, 2] / min(sw, sh) - 1) / f)
    return torch.stack([(u - cx) * z / f, (v - cy) * z / f, z], dim=1)
