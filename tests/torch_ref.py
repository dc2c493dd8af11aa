"""Plain-PyTorch fp32 references of the fused operators in camliflow_b200.ops -- the reference's
own formulas (cited), written with stock torch ops.  TEST INFRASTRUCTURE: used by the GPU kernel
tests (kernel vs. formula on the same tensors) and by the CPU host-logic tests (patched in
place of the kernels)."""
import torch
import torch.nn.functional as F


def gather_cf(data, idx):
    B, C = data.shape[:2]
    flat = idx.reshape(B, 1, -1).expand(B, C, -1)
    return torch.gather(data, 2, flat).view([B, C] + list(idx.shape[1:]))


def knn_interpolate(input_xyz, input_feat, query_xyz, idx):
    """models/utils.py:130-146 with the neighbour indices given."""
    d = torch.linalg.norm(gather_cf(input_xyz, idx) - query_xyz[..., None], dim=1).clamp(1e-8)
    w = 1.0 / d
    w = w / torch.sum(w, -1, keepdim=True)
    return torch.sum(gather_cf(input_feat, idx) * w[:, None], -1)


def bilinear_sample(feat2d, uv):
    """models/utils.py:262-269."""
    H, W = feat2d.shape[2:]
    gx = 2.0 * uv[:, 0] / (W - 1) - 1.0
    gy = 2.0 * uv[:, 1] / (H - 1) - 1.0
    g = torch.stack([gx, gy], -1)[:, :, None, :]
    return F.grid_sample(feat2d, g, "bilinear", align_corners=True)[..., 0]


def corr2d_build(fmap1, fmap2, num_levels):
    """models/raft_core.py:56-68."""
    B, C, H, W = fmap1.shape
    vol = torch.matmul(fmap1.view(B, C, H * W).transpose(1, 2), fmap2.view(B, C, H * W))
    vol = (vol / torch.sqrt(torch.tensor(float(C)))).reshape(B * H * W, 1, H, W)
    pyr = [vol]
    for _ in range(num_levels - 1):
        vol = F.avg_pool2d(vol, 2, stride=2)
        pyr.append(vol)
    return [v.view(B, H * W, v.shape[-2], v.shape[-1]) for v in pyr]


def corr2d_lookup(pyramid, coords, r):
    """models/raft_core.py:71-107."""
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


def corr3d_pool(vol, idx):
    """models/camliraft_l_core.py:56-60."""
    return torch.mean(gather_cf(vol, idx), -1)


def corr3d_lookup(xyz1, xyzs2, pyramid, idxs, W1, b1, W2, b2):
    """models/camliraft_l_core.py:62-98 with the neighbour indices given: [B,32L,n1]."""
    costs = []
    for xyz2, vol, idx in zip(xyzs2, pyramid, idxs):
        B, n1, n2 = vol.shape
        off = gather_cf(xyz2, idx) - xyz1[:, :, :, None]
        c = torch.gather(vol, 2, idx).view(B, 1, n1, -1)
        x = torch.cat([off, c], 1)
        x = F.relu(F.conv2d(x, W1[:, :, None, None], b1))
        x = F.relu(F.conv2d(x, W2[:, :, None, None], b2))
        costs.append(x.sum(-1))
    return torch.cat(costs, 1)


def pointconv_dw_weights(xyz, sampled_xyz, idx, params):
    """models/point_conv.py:122-127: [B,S,k,O]."""
    x = gather_cf(xyz, idx) - sampled_xyz[:, :, :, None]
    for w, b in zip(params[0::2], params[1::2]):
        x = F.relu(F.conv2d(x, w[:, :, None, None], b))
    return x.permute(0, 2, 3, 1).contiguous()


def pointconv_dw_gather_max(feat_cf, weights_bsko, idx):
    """models/point_conv.py:126-128: [B,O,S]."""
    return torch.max(gather_cf(feat_cf, idx) * weights_bsko.permute(0, 3, 1, 2), -1)[0]


def clfm_interp(uv, nn_idx, feat3d, W1, b1, W2, b2, H, W):
    """models/clfm.py:57-75 before out_conv: [B,C,H,W]."""
    B = uv.shape[0]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=uv.device),
                            torch.arange(W, dtype=torch.float32, device=uv.device), indexing="ij")
    grid = torch.stack([xs, ys], 0).reshape(1, 2, -1).expand(B, 2, -1)
    off = gather_cf(uv, nn_idx) - grid
    si = torch.cat([off, torch.linalg.norm(off, dim=1, keepdim=True)], 1)[..., None]
    s = F.leaky_relu(F.conv2d(si, W1[:, :, None, None], b1), 0.1)
    s = torch.sigmoid(F.conv2d(s, W2[:, :, None, None], b2))
    return (s[..., 0] * gather_cf(feat3d, nn_idx)).view(B, -1, H, W)


def pointconv_group(xyz, feat, sampled_xyz, idx, W1, b1, W2, b2, slope):
    """models/point_conv.py:56-66: [B,S,16*(3+C)]."""
    B, S = idx.shape[:2]
    off = gather_cf(xyz, idx) - sampled_xyz[:, :, :, None]
    w = F.leaky_relu(F.conv2d(off, W1[:, :, None, None], b1), slope)
    w = F.leaky_relu(F.conv2d(w, W2[:, :, None, None], b2), slope).transpose(1, 2)      # [B,S,16,k]
    g = gather_cf(torch.cat([xyz, feat], 1), idx).permute(0, 2, 3, 1)                   # [B,S,k,3+C]
    return torch.matmul(w, g).reshape(B, S, -1)


def sk_fusion_tail(a, b, slope, w_mid, w_out):
    """models/clfm.py:199-214 on rows [B,P,C]."""
    a, b = F.leaky_relu(a, slope), F.leaky_relu(b, slope)
    B, P, C = a.shape
    w = (a + b).mean(1)
    w = torch.sigmoid(F.linear(F.relu(F.linear(w, w_mid)), w_out)).view(B, C, 2)
    w = torch.softmax(w, -1)
    return a * w[:, None, :, 0] + b * w[:, None, :, 1]


def correlation2d(input1, input2, max_displacement):
    """models/csrc/wrapper.py:41-50 (the reference's own fallback): NCHW -> [B,(2d+1)^2,H,W]."""
    H, W = input1.shape[2:]
    d = max_displacement
    padded = F.pad(input2, [d] * 4)
    vols = []
    for i in range(2 * d + 1):
        for j in range(2 * d + 1):
            vols.append(torch.mean(input1 * padded[:, :, i:i + H, j:j + W], 1, keepdim=True))
    return torch.cat(vols, 1)
