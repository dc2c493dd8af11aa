"""Fused hot-path operators (host side).  Every function here is the doorway to one kernel
(family) of libcamli_b200.so; tensors keep the reference's logical layouts at this boundary
([B,C,N] point features, [B,C,H,W] maps) so the callers read like the reference's cores.

There is no CPU path: all functions require CUDA tensors and a built library.
"""
import torch
import torch.nn.functional as F

from .csrc import k_nearest_neighbor


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("camliflow_b200.ops: CUDA tensors required (there is no CPU fallback)")


# ---------------------------------------------------------------- grouping / interpolation
def gather_points(data, idx):
    """data [B,C,N], idx [B,...] (i64) -> [B,C,...]: models/utils.py:62-80."""
    _need_cuda(data, idx)
    B, C = data.shape[:2]
    flat = idx.reshape(B, 1, -1).expand(B, C, -1)
    return torch.gather(data, 2, flat).view([B, C] + list(idx.shape[1:]))


def knn_interpolate(input_xyz, input_feat, query_xyz, k=3):
    """models/utils.py:130-146."""
    _need_cuda(input_xyz, input_feat, query_xyz)
    idx = k_nearest_neighbor(input_xyz, query_xyz, k)
    d = torch.linalg.norm(gather_points(input_xyz, idx) - query_xyz[..., None], dim=1).clamp(1e-8)
    w = 1.0 / d
    w = w / torch.sum(w, -1, keepdim=True)
    return torch.sum(gather_points(input_feat, idx) * w[:, None], -1)


# ---------------------------------------------------------------- image-side sampling
def bilinear_sample(feat2d, uv):
    """feat2d [B,C,H,W], uv [B,2,N] pixel coords -> [B,C,N] (align_corners, zero padding)."""
    _need_cuda(feat2d, uv)
    H, W = feat2d.shape[2:]
    gx = 2.0 * uv[:, 0] / (W - 1) - 1.0
    gy = 2.0 * uv[:, 1] / (H - 1) - 1.0
    g = torch.stack([gx, gy], -1)[:, :, None, :]
    return F.grid_sample(feat2d, g, "bilinear", align_corners=True)[..., 0]


def convex_upsample(flow, mask, s=8):
    """models/utils.py:191-204."""
    _need_cuda(flow, mask)
    B, _, H, W = flow.shape
    mask = torch.softmax(mask.float().view(B, 1, 9, s, s, H, W), 2)
    up = F.unfold(flow.float() * s, [3, 3], padding=1).view(B, 2, 9, 1, 1, H, W)
    up = torch.sum(mask * up, 2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(B, 2, H * s, W * s)


# ---------------------------------------------------------------- RAFT all-pairs correlation
def corr2d_build(fmap1, fmap2, num_levels):
    """All-pairs volume of two [B,C,H,W] maps scaled by 1/sqrt(C) plus its 2x2 average-pooled
    pyramid over the (h2,w2) axes: models/raft_core.py:56-68.  Returns a list of [B,H*W,H_l,W_l]."""
    _need_cuda(fmap1, fmap2)
    B, C, H, W = fmap1.shape
    vol = torch.matmul(fmap1.view(B, C, H * W).transpose(1, 2), fmap2.view(B, C, H * W))
    vol = (vol / torch.sqrt(torch.tensor(float(C)))).reshape(B * H * W, 1, H, W)
    pyr = [vol]
    for _ in range(num_levels - 1):
        vol = F.avg_pool2d(vol, 2, stride=2)
        pyr.append(vol)
    return [v.view(B, H * W, v.shape[-2], v.shape[-1]) for v in pyr]


def corr2d_lookup(pyramid, coords, radius):
    """models/raft_core.py:71-107: coords [B,2,H,W] -> [B, L*(2r+1)^2, H, W]; window index i moves x,
    j moves y (the reference's meshgrid quirk)."""
    _need_cuda(coords)
    r = radius
    coords = coords.permute(0, 2, 3, 1).float()
    B, H, W, _ = coords.shape
    d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), -1).view(1, 2 * r + 1, 2 * r + 1, 2)
    out = []
    for i, vol in enumerate(pyramid):
        h, w = vol.shape[-2:]
        c = coords.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta
        g = torch.cat([2 * c[..., 0:1] / (w - 1) - 1, 2 * c[..., 1:2] / (h - 1) - 1], -1)
        s = F.grid_sample(vol.reshape(B * H * W, 1, h, w), g, align_corners=True)
        out.append(s.view(B, H, W, -1))
    return torch.cat(out, -1).permute(0, 3, 1, 2).contiguous()


# ---------------------------------------------------------------- point all-pairs correlation
def corr3d_build(feat1, feat2, xyzs2, k=3):
    """models/camliraft_l_core.py:51-60."""
    _need_cuda(feat1, feat2)
    vol = torch.bmm(feat1.float().transpose(1, 2), feat2.float()) / feat1.shape[1]
    pyr = [vol]
    for i in range(1, len(xyzs2)):
        idx = k_nearest_neighbor(xyzs2[i - 1], xyzs2[i], k)
        pyr.append(torch.mean(gather_points(pyr[i - 1], idx), -1))
    return pyr


def corr3d_gather(xyz1, xyz2, volume, k):
    """For every point of xyz1 its k nearest points of xyz2: relative offsets and the matching
    volume entries, stacked as [B,4,n1,k] (models/camliraft_l_core.py:62-76)."""
    _need_cuda(xyz1, xyz2, volume)
    B, n1, n2 = volume.shape
    idx = k_nearest_neighbor(xyz2, xyz1, k)
    off = gather_points(xyz2, idx) - xyz1[:, :, :, None]
    c = torch.gather(volume, 2, idx).view(B, 1, n1, k)
    return torch.cat([off, c], 1)


# ---------------------------------------------------------------- point convolutions
def neighbor_offsets(xyz, sampled_xyz, idx):
    """[B,3,S,k] offsets of the grouped neighbours from their centroid."""
    return gather_points(xyz, idx) - sampled_xyz[:, :, :, None]


def pointconv_dw_aggregate(feat, weights, idx):
    """max_k( feat[:, :, idx] * weights ): feat [B,O,N], weights [B,O,S,k], idx [B,S,k] -> [B,O,S]
    (models/point_conv.py:126-128)."""
    return torch.max(gather_points(feat, idx) * weights, -1)[0]


def pointconv_aggregate(feat, weights, idx):
    """Per centroid [16 x k] @ [k x C]: feat [B,C,N], weights [B,16,S,k], idx [B,S,k]
    -> [B,S,16*C] (models/point_conv.py:62-66)."""
    B, S = idx.shape[:2]
    g = gather_points(feat, idx).permute(0, 2, 3, 1)
    return torch.matmul(weights.transpose(1, 2), g).reshape(B, S, -1)


# ---------------------------------------------------------------- CLFM
def nearest_point_2d(uv, H, W):
    """Index [B,H*W] of the projected point nearest to every pixel centre (2-D k-NN, k=1;
    models/clfm.py:57-60)."""
    from .utils import mesh_grid
    B = uv.shape[0]
    grid = mesh_grid(B, H, W, uv.device).reshape(B, 2, -1)
    return k_nearest_neighbor(uv, grid, 1)[..., 0]


# ---------------------------------------------------------------- per-launch profiling (bench.py)
def profile_begin():
    from . import native
    native.profile_begin()


def profile_end(peaks_path=None):
    """Roofline record of the hand-written kernel with the largest share of the profiled
    region: achieved = algorithmic bytes per launch / average launch duration."""
    import json
    import os
    from . import native
    prof = native.profile_end()
    if not prof:
        return None
    peak, src = 6650.0, "fallback"
    if peaks_path and os.path.exists(peaks_path):
        peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    top = max(prof, key=lambda n: prof[n]["total_us"])
    rec = prof[top]
    achieved = rec["bytes"] / (rec["avg_us"] * 1e-6) / 1e9
    return {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": src, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None, "avg_us": rec["avg_us"], "launches": rec["launches"],
            "algorithmic_bytes_per_launch": rec["bytes"],
            "all": {n: {"avg_us": r["avg_us"], "launches": r["launches"], "total_us": r["total_us"],
                        "GBps": r["bytes"] / (r["avg_us"] * 1e-6) / 1e9, "GFLOPs": r["flops"] / (r["avg_us"] * 1e-6) / 1e9}
                    for n, r in prof.items()}}
