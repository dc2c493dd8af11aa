"""camliraft_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Plain-PyTorch, CPU, fp32 restatement of the reference's CamLiRAFT forward
(MCG-NJU/CamLiFlow @3bf1974): wrapper models/camliraft.py:32-73, core
models/camliraft_core.py:33-145, 2-D branch models/raft_core.py, 3-D branch
models/camliraft_l_core.py, fusion models/clfm.py, point ops models/point_conv.py,
models/utils.py, IDS models/ids.py.  It is written functionally over a flat
{name: tensor} dictionary that uses the reference's own state_dict names
(oracle/param_spec_camliraft.json), so the same seeded weights drive the
reference, this oracle and the product.

Pinning: tests/golden/make_golden_model.py runs the reference model itself (imported
from /root/reference in the build container, with the mmdet ResNet stand-in of
tests/golden/ref_harness.py) on seeded inputs and commits its outputs under
tests/golden/; tests/test_oracle_model.py checks this file against them.

Two index semantics are provided because the reference itself has two:
  * index_impl="fallback": the pure-torch FPS / k-NN of models/csrc/wrapper.py:83-96,
    115-117 -- what the reference executes on CPU tensors (and what bench.py's
    `--impl reference` / cpu_baseline leg times);
  * index_impl="kernel": the semantics of the reference's CUDA kernels (what it executes
    on a GPU), through oracle/kernels_oracle.c, itself pinned bit-exact against the
    reference kernels run on a B200 (tests/golden/l0_reference_cuda.npz).  This is the
    oracle the product's flows are compared with.
"""
import json
import os
import re
import zlib

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ parameters
def param_spec(model="camliraft"):
    with open(os.path.join(HERE, "param_spec_%s.json" % model)) as f:
        return {k: tuple(v) for k, v in json.load(f).items()}


DAMPING = {
    # CamLiRAFT: last layer of each flow head
    r"flow_head\.conv2\.weight$": 0.05, r"flow_head\.fc\.weight$": 0.05,
    # CamLiPWC: un-normalised point-geometry features grow ~5x per PointConv level and ~100x through the
    # learnable cost volume; keep activations O(1-10) and the coarse-to-fine flow updates small
    r"conv_last\.weight$": 0.01,
    r"branch_3d_fnet\.level0_mlp\.convs\.0\.conv_fn\.weight$": 0.1,
    r"pyramid_convs\.\d\.linear\.weight$": 0.2, r"point_conv[12]\.linear\.weight$": 0.2,
    r"weight_net[12]\.convs\.2\.conv_fn\.weight$": 0.1,
}


def make_params(spec, seed=0):
    """Deterministic weights keyed on the parameter NAME (independent of construction order):
    conv / linear weights ~ U(+-sqrt(3/fan_in)), biases ~ N(0, 0.05), norm scales ~ U(0.8, 1.2),
    running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5).  The last layer of every flow head is
    damped (x0.05 in CamLiRAFT, x0.002 for the `conv_last` layers of CamLiPWC, whose un-normalised
    point-geometry features are O(100)) so that the recurrent refinement of a random-weight network stays in the
    contractive, small-flow regime a trained network works in (undamped, 12 iterations amplify
    a 1-ulp perturbation to whole pixels, which would make any parity tolerance meaningless)."""
    out = {}
    for name in sorted(spec):
        shape = spec[name]
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif len(shape) == 1:   # norm scale
            t = torch.rand(shape, generator=g) * 0.4 + 0.8
        else:
            fan_in = int(np.prod(shape[1:]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * (3.0 / fan_in) ** 0.5
            for pattern, gain in DAMPING.items():
                if re.search(pattern, name):
                    t = t * gain
        out[name] = t
    return out


# ------------------------------------------------------------------ index ops
def _kernel_lib():
    import ctypes
    from oracle import build as oracle_build
    return ctypes.CDLL(oracle_build.build_kernels_oracle())


_KLIB = None


def _np_ptr(a):
    import ctypes
    return a.ctypes.data_as(ctypes.c_void_p)


def fps(xyz, n_samples, index_impl):
    """xyz [B,N,3] -> [B,S] i64 (wrapper.py:75-103)."""
    global _KLIB
    assert xyz.shape[2] == 3 and xyz.shape[1] > n_samples
    if index_impl == "kernel":
        _KLIB = _KLIB or _kernel_lib()
        a = np.ascontiguousarray(xyz.numpy(), dtype=np.float32)
        out = np.empty((a.shape[0], n_samples), dtype=np.int64)
        _KLIB.oracle_furthest_point_sampling(_np_ptr(a), a.shape[0], a.shape[1], n_samples, _np_ptr(out))
        return torch.from_numpy(out)
    B, N, _ = xyz.shape    # wrapper.py:83-96
    sel = torch.zeros(B, n_samples, dtype=torch.int64)
    dist = torch.ones(B, N) * 1e10
    cur = torch.zeros(B, dtype=torch.int64)
    rows = torch.arange(B)
    for i in range(n_samples):
        sel[:, i] = cur
        c = xyz[rows, cur, :].view(B, 1, 3)
        nd = torch.sum((xyz - c) ** 2, -1)
        m = nd < dist
        dist[m] = nd[m]
        cur = torch.max(dist, -1)[1]
    return sel


def knn(input_xyz, query_xyz, k, index_impl):
    """input [B,D,m], query [B,D,n] channel-first -> [B,n,k] i64 (wrapper.py:106-127)."""
    global _KLIB
    a = input_xyz.transpose(1, 2).contiguous()
    q = query_xyz.transpose(1, 2).contiguous()
    if index_impl == "kernel":
        _KLIB = _KLIB or _kernel_lib()
        an, qn = a.numpy(), q.numpy()
        B, n, D = qn.shape
        out = np.empty((B, n, k), dtype=np.int64)
        _KLIB.oracle_k_nearest_neighbor(B, n, an.shape[1], k, D, _np_ptr(qn), _np_ptr(an), _np_ptr(out))
        return torch.from_numpy(out)
    d = -2 * torch.matmul(q, a.permute(0, 2, 1))     # wrapper.py:60-72
    d += torch.sum(q ** 2, -1).view(q.shape[0], q.shape[1], 1)
    d += torch.sum(a ** 2, -1).view(a.shape[0], 1, a.shape[1])
    return d.topk(k, dim=2, largest=False).indices


def gather_cf(data, idx):
    """data [B,C,N], idx [B,...] -> [B,C,...] (utils.py:62-80)."""
    B, C = data.shape[:2]
    flat = idx.reshape(B, 1, -1).expand(B, C, -1)
    return torch.gather(data, 2, flat).view([B, C] + list(idx.shape[1:]))


# ------------------------------------------------------------------ small layers
def _act(x, act):
    if act == "leaky_relu":
        return F.leaky_relu(x, 0.1)
    if act == "relu":
        return F.relu(x)
    if act == "sigmoid":
        return torch.sigmoid(x)
    assert act is None
    return x


def conv_norm_act(P, pre, x, act="leaky_relu"):
    """Conv{1,2}dNormRelu, 1x1, eval mode (mlp.py:41-128): bias iff no norm; BN uses running stats."""
    w = P[pre + ".conv_fn.weight"]
    b = P.get(pre + ".conv_fn.bias")
    x = F.conv1d(x, w, b) if w.dim() == 3 else F.conv2d(x, w, b)
    if pre + ".norm_fn.running_mean" in P:
        x = F.batch_norm(x, P[pre + ".norm_fn.running_mean"], P[pre + ".norm_fn.running_var"],
                         P[pre + ".norm_fn.weight"], P[pre + ".norm_fn.bias"], False, 0.0, 1e-5)
    return _act(x, act)


def mlp(P, pre, x, n, act="leaky_relu"):
    for i in range(n):
        x = conv_norm_act(P, "%s.convs.%d" % (pre, i), x, act)
    return x


def _bn2d(P, pre, x):
    return F.batch_norm(x, P[pre + ".running_mean"], P[pre + ".running_var"], P[pre + ".weight"], P[pre + ".bias"],
                        False, 0.0, 1e-5)


def encoder2d(P, pre, x):
    """Encoder2D (raft_core.py:10-38): ResNet-50 stem + stages 1-2 (stand-in with torchvision
    layout / mmdet 'pytorch' style: stride on the 3x3 conv), eval-mode BN, then 1x1 align."""
    x = F.relu(_bn2d(P, pre + ".bn1", F.conv2d(x, P[pre + ".conv1.weight"], None, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    for layer, blocks, stride in (("layer1", 3, 1), ("layer2", 4, 2)):
        for b in range(blocks):
            p = "%s.%s.%d" % (pre, layer, b)
            s = stride if b == 0 else 1
            y = F.relu(_bn2d(P, p + ".bn1", F.conv2d(x, P[p + ".conv1.weight"])))
            y = F.relu(_bn2d(P, p + ".bn2", F.conv2d(y, P[p + ".conv2.weight"], None, s, 1)))
            y = _bn2d(P, p + ".bn3", F.conv2d(y, P[p + ".conv3.weight"]))
            if p + ".downsample.0.weight" in P:
                x = _bn2d(P, p + ".downsample.1", F.conv2d(x, P[p + ".downsample.0.weight"], None, s))
            x = F.relu(y + x)
    return conv_norm_act(P, pre + ".align", x)


# ------------------------------------------------------------------ point ops
def point_conv(P, pre, xyz, feat, sampled_xyz, k, index_impl, has_norm):
    """PointConv (point_conv.py:35-70)."""
    B, S = sampled_xyz.shape[0], sampled_xyz.shape[-1]
    feat = torch.cat([xyz, feat], 1)
    idx = knn(xyz, sampled_xyz, k, index_impl)
    off = gather_cf(xyz, idx) - sampled_xyz[:, :, :, None]
    w = mlp(P, pre + ".weight_net", off, 2).transpose(1, 2)                 # [B,S,16,k]
    g = gather_cf(feat, idx).permute(0, 2, 3, 1)                            # [B,S,k,C+3]
    out = torch.matmul(w, g).reshape(B, S, -1)
    out = F.linear(out, P[pre + ".linear.weight"], P[pre + ".linear.bias"]).transpose(1, 2)
    if has_norm:
        out = F.batch_norm(out, P[pre + ".norm_fn.running_mean"], P[pre + ".norm_fn.running_var"],
                           P[pre + ".norm_fn.weight"], P[pre + ".norm_fn.bias"], False, 0.0, 1e-5)
    return F.leaky_relu(out, 0.1)


def point_conv_dw(P, pre, xyz, feat, knn_idx, k, act="leaky_relu"):
    """PointConvDW (point_conv.py:109-130) with a precomputed (wider) neighbour table."""
    idx = knn_idx[:, :, :k]
    off = gather_cf(xyz, idx) - xyz[:, :, :, None]
    f = mlp(P, pre + ".mlp", feat, 1, act)
    f = gather_cf(f, idx) * mlp(P, pre + ".weight_net", off, 3, "relu")
    return torch.max(f, -1)[0]


def knn_interpolation(input_xyz, input_feat, query_xyz, index_impl, k=3):
    """utils.py:130-146."""
    idx = knn(input_xyz, query_xyz, k, index_impl)
    d = torch.linalg.norm(gather_cf(input_xyz, idx) - query_xyz[..., None], dim=1).clamp(1e-8)
    w = 1.0 / d
    w = w / torch.sum(w, -1, keepdim=True)
    return torch.sum(gather_cf(input_feat, idx) * w[:, None], -1)


def backwarp_3d(xyz1, xyz2, flow12, index_impl):
    """utils.py:149-159."""
    return xyz2 + knn_interpolation(xyz1 + flow12, -flow12, xyz2, index_impl)


def encoder3d(P, pre, xyzs, index_impl):
    """Encoder3D (camliraft_l_core.py:22-37), n_channels [64,96,128], batch_norm, k=16."""
    feat = mlp(P, pre + ".level0_mlp", xyzs[0], 2)
    for i in range(2):
        feat = mlp(P, "%s.mlps.%d" % (pre, i), feat, 2)
        feat = point_conv(P, "%s.convs.%d" % (pre, i), xyzs[i], feat, xyzs[i + 1], 16, index_impl, True)
    return feat


# ------------------------------------------------------------------ 2-D branch
def mesh_grid(B, H, W):
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    return torch.stack([xs, ys], 0)[None].expand(B, 2, H, W)


def corr2d_build(P, pre, f1, f2, levels=4):
    """Correlation2D.build_cost_volume_pyramid (raft_core.py:52-68)."""
    w, b = P[pre + ".fnet_aligner.weight"], P[pre + ".fnet_aligner.bias"]
    f1, f2 = F.conv2d(f1, w, b), F.conv2d(f2, w, b)
    B, C, H, W = f1.shape
    vol = torch.matmul(f1.view(B, C, H * W).transpose(1, 2), f2.view(B, C, H * W))
    vol = (vol / torch.sqrt(torch.tensor(C))).reshape(B * H * W, 1, H, W)
    pyr = [vol]
    for _ in range(levels - 1):
        vol = F.avg_pool2d(vol, 2, stride=2)
        pyr.append(vol)
    return pyr


def corr2d_lookup(pyr, coords, r=4):
    """Correlation2D.forward (raft_core.py:71-107), including its x/y-swapped window."""
    coords = coords.permute(0, 2, 3, 1)
    B, H, W, _ = coords.shape
    d = torch.linspace(-r, r, 2 * r + 1)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), -1).view(1, 2 * r + 1, 2 * r + 1, 2)
    out = []
    for i, vol in enumerate(pyr):
        c = coords.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta
        h, w = vol.shape[-2:]
        gx = 2 * c[..., 0:1] / (w - 1) - 1
        gy = 2 * c[..., 1:2] / (h - 1) - 1
        s = F.grid_sample(vol, torch.cat([gx, gy], -1), align_corners=True)
        out.append(s.view(B, H, W, -1))
    return torch.cat(out, -1).permute(0, 3, 1, 2).contiguous()


def _conv(P, name, x, pad):
    return F.conv2d(x, P[name + ".weight"], P[name + ".bias"], padding=pad)


def motion_encoder2d(P, pre, flow, corr):
    """MotionEncoder2D (raft_core.py:155-166)."""
    c = F.relu(_conv(P, pre + ".conv_c1", corr, 0))
    c = F.relu(_conv(P, pre + ".conv_c2", c, 1))
    f = F.relu(_conv(P, pre + ".conv_f1", flow, 3))
    f = F.relu(_conv(P, pre + ".conv_f2", f, 1))
    out = torch.nan_to_num(F.relu(_conv(P, pre + ".conv", torch.cat([c, f], 1), 1)))
    return torch.cat([out, flow], 1)


def gru2d(P, pre, h, x):
    """GRU2D (raft_core.py:123-139)."""
    for sfx, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(_conv(P, pre + ".convz" + sfx, hx, pad))
        r = torch.sigmoid(_conv(P, pre + ".convr" + sfx, hx, pad))
        q = torch.tanh(_conv(P, pre + ".convq" + sfx, torch.cat([r * h, x], 1), pad))
        h = (1 - z) * h + z * q
    return torch.nan_to_num(h)


def convex_upsample(flow, mask, s=8):
    """utils.py:191-204."""
    B, _, H, W = flow.shape
    mask = torch.softmax(mask.view(B, 1, 9, s, s, H, W), 2)
    up = F.unfold(flow * s, [3, 3], padding=1).view(B, 2, 9, 1, 1, H, W)
    up = torch.sum(mask * up, 2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(B, 2, H * s, W * s)


# ------------------------------------------------------------------ 3-D branch
def corr3d_build(feat1, feat2, xyzs2, index_impl, k=3):
    """Correlation3D.build_cost_volume_pyramid (camliraft_l_core.py:51-60)."""
    vol = torch.bmm(feat1.transpose(1, 2), feat2) / feat1.shape[1]
    pyr = [vol]
    for i in range(1, len(xyzs2)):
        idx = knn(xyzs2[i - 1], xyzs2[i], k, index_impl)
        pyr.append(torch.mean(gather_cf(pyr[i - 1], idx), -1))
    return pyr


def corr3d_lookup(P, pre, xyz1, xyzs2, pyr, index_impl, k=16):
    """Correlation3D.forward / calc_matching_cost (camliraft_l_core.py:62-101)."""
    costs = []
    for xyz2, vol in zip(xyzs2, pyr):
        B, n1, n2 = vol.shape
        idx = knn(xyz2, xyz1, k, index_impl)
        off = gather_cf(xyz2, idx) - xyz1.view(B, 3, n1, 1)
        c = torch.gather(vol, 2, idx).reshape(B, 1, n1, k)
        costs.append(torch.sum(mlp(P, pre + ".cost_mlp", torch.cat([off, c], 1), 2, "relu"), -1))
    return conv_norm_act(P, pre + ".merge", torch.cat(costs, 1))


def motion_encoder3d(P, pre, xyz, flow, corr, nbr):
    """MotionEncoder3D (camliraft_l_core.py:146-155)."""
    c = point_conv_dw(P, pre + ".conv_c1", xyz, corr, nbr, 16)
    f = point_conv_dw(P, pre + ".conv_f1", xyz, flow, nbr, 32)
    f = point_conv_dw(P, pre + ".conv_f2", xyz, f, nbr, 16)
    out = point_conv_dw(P, pre + ".conv", xyz, torch.cat([c, f], 1), nbr, 16)
    return torch.cat([out, flow], 1)


def gru3d(P, pre, xyz, h, x, nbr):
    """GRU3D (camliraft_l_core.py:127-134)."""
    hx = torch.cat([h, x], 1)
    z = torch.sigmoid(point_conv_dw(P, pre + ".conv_z", xyz, hx, nbr, 4, None))
    r = torch.sigmoid(point_conv_dw(P, pre + ".conv_r", xyz, hx, nbr, 4, None))
    q = torch.tanh(point_conv_dw(P, pre + ".conv_q", xyz, torch.cat([r * h, x], 1), nbr, 4, None))
    return (1 - z) * h + z * q


def flow_head3d(P, pre, xyz, h, nbr):
    """FlowHead3D (camliraft_l_core.py:111-116)."""
    f = point_conv_dw(P, pre + ".conv1", xyz, h, nbr, 32)
    f = point_conv_dw(P, pre + ".conv2", xyz, f, nbr, 32)
    return F.conv1d(f, P[pre + ".fc.weight"], P[pre + ".fc.bias"])


# ------------------------------------------------------------------ CLFM
def grid_sample_uv(feat2d, uv):
    """grid_sample_wrapper (utils.py:262-269)."""
    H, W = feat2d.shape[2:]
    gx = 2.0 * uv[:, 0] / (W - 1) - 1.0
    gy = 2.0 * uv[:, 1] / (H - 1) - 1.0
    g = torch.cat([gx[:, :, None, None], gy[:, :, None, None]], -1)
    return F.grid_sample(feat2d, g, "bilinear", align_corners=True)[..., 0]


def sk_fusion(P, pre, a, b):
    """SKFusion (clfm.py:195-214)."""
    B = a.shape[0]
    a, b = conv_norm_act(P, pre + ".align1", a), conv_norm_act(P, pre + ".align2", b)
    w = (a + b).mean(dim=tuple(range(2, a.dim())))
    w = F.relu(F.linear(w, P[pre + ".fc_mid.0.weight"]))
    w = torch.sigmoid(F.linear(w, P[pre + ".fc_out.0.weight"])).reshape(B, -1, 2)
    w = torch.softmax(w, -1)
    shape = [B, -1] + [1] * (a.dim() - 2)
    return a * w[..., 0].reshape(shape) + b * w[..., 1].reshape(shape)


def clfm(P, pre, uv, feat2d, feat3d, index_impl):
    """CLFM.forward with FusionAwareInterp k=1 and SK fusion (clfm.py:30-79)."""
    B, _, H, W = feat2d.shape
    C3 = feat3d.shape[1]
    grid = mesh_grid(B, H, W).reshape(B, 2, -1)
    idx = knn(uv, grid, 1, index_impl)
    g = gather_cf(torch.cat([uv, feat3d], 1), idx)
    off = g[:, :2] - grid[..., None]
    si = torch.cat([off, torch.linalg.norm(off, dim=1, keepdim=True)], 1)
    score = conv_norm_act(P, pre + ".interp.score_net.1", conv_norm_act(P, pre + ".interp.score_net.0", si), "sigmoid")
    interp = (score * g[:, 2:]).sum(-1).reshape(B, C3, H, W)
    interp = conv_norm_act(P, pre + ".interp.out_conv", interp)
    out2d = sk_fusion(P, pre + ".fuse2d", feat2d, interp)
    sampled = conv_norm_act(P, pre + ".mlps3d", grid_sample_uv(feat2d, uv))
    out3d = sk_fusion(P, pre + ".fuse3d", sampled, feat3d)
    return out2d, out3d


# ------------------------------------------------------------------ IDS (ids.py)
def persp2paral(xyz, f, cx, cy, ph, pw, qh, qw):
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    f, cx, cy = f[:, None], cx[:, None], cy[:, None]
    sw, sh = (qw - 1) / (pw - 1), (qh - 1) / (ph - 1)
    return torch.stack([(cx + (f / z) * x) * sw - (qw - 1) / 2,
                        (cy + (f / z) * y) * sh - (qh - 1) / 2,
                        (f * torch.log(z) + 1) * min(sw, sh)], 1)


def paral2persp(xyz, f, cx, cy, ph, pw, qh, qw):
    f, cx, cy = f[:, None], cx[:, None], cy[:, None]
    sw, sh = (qw - 1) / (pw - 1), (qh - 1) / (ph - 1)
    x = (xyz[:, 0] + (qw - 1) / 2) / sw
    y = (xyz[:, 1] + (qh - 1) / 2) / sh
    z = torch.exp((xyz[:, 2] / min(sw, sh) - 1) / f)
    return torch.stack([(x - cx) * z / f, (y - cy) * z / f, z], 1)


# ------------------------------------------------------------------ the model
def camliraft_forward(P, images, pcs, intrinsics, n_iters=12, index_impl="kernel", all_iters=False):
    """CamLiRAFT.forward (camliraft.py:32-73) in eval mode with every fusion site but `hidden` on
    (conf/model/camliraft.yaml).  images [B,6,H,W] (0..255), pcs [B,6,N], intrinsics [B,3] (f,cx,cy).
    Returns {'flow_2d': [B,2,H,W], 'flow_3d': [B,3,N]} (plus the per-iteration lists when all_iters)."""
    with torch.no_grad():
        images = images.float()
        pc1, pc2 = pcs[:, :3].float(), pcs[:, 3:].float()
        H0, W0 = images.shape[-2:]
        pad_h, pad_w = (-H0) % 8, (-W0) % 8                                   # InputPadder(x=8), utils.py:7-20
        pad = [pad_w // 2, pad_w - pad_w // 2, 0, pad_h]
        images = F.pad(images, pad, mode="replicate")
        mean = torch.tensor([123.675, 116.280, 103.530]).reshape(1, 3, 1, 1)
        std = torch.tensor([58.395, 57.120, 57.375]).reshape(1, 3, 1, 1)
        image1 = (images[:, :3] - mean) / std
        image2 = (images[:, 3:] - mean) / std
        Hp, Wp = image1.shape[-2:]
        qh, qw = round(Hp / 32), round(Wp / 32)
        cam = (intrinsics[:, 0], intrinsics[:, 1], intrinsics[:, 2], Hp, Wp, qh, qw)
        pc1 = persp2paral(pc1, *cam)
        pc2 = persp2paral(pc2, *cam)

        # ---- core (camliraft_core.py:33-145)
        B = pc1.shape[0]
        sel = fps(torch.cat([pc1, pc2], 0).transpose(1, 2), 4096, index_impl)     # utils.py:107-127
        s1, s2 = sel[:B], sel[B:]
        xyzs1 = [pc1] + [gather_cf(pc1, s1[:, :n]) for n in (4096, 2048, 1024, 512, 256)]
        xyzs2 = [pc2] + [gather_cf(pc2, s2[:, :n]) for n in (4096, 2048, 1024, 512, 256)]

        b2, b3 = "core.branch_2d", "core.branch_3d"
        f1_2d = encoder2d(P, b2 + ".fnet", image1)
        f2_2d = encoder2d(P, b2 + ".fnet", image2)
        fc_2d = encoder2d(P, b2 + ".cnet", image1)
        f1_3d = encoder3d(P, b3 + ".fnet", xyzs1[:3], index_impl)
        f2_3d = encoder3d(P, b3 + ".fnet", xyzs2[:3], index_impl)
        fc_3d = encoder3d(P, b3 + ".cnet", xyzs1[:3], index_impl)
        xyzs1, xyzs2 = xyzs1[2:], xyzs2[2:]
        xyz1, xyz2 = xyzs1[0], xyzs2[0]

        h8, w8 = f1_2d.shape[-2:]
        cxp, cyp = (qw - 1) / 2, (qh - 1) / 2                                  # parallel camera, utils.py:251-253
        scale = torch.tensor([(w8 - 1) / (qw - 1), (h8 - 1) / (qh - 1)]).reshape(1, 2, 1)
        uv1 = torch.stack([xyz1[:, 0] + cxp, xyz1[:, 1] + cyp], 1) * scale
        uv2 = torch.stack([xyz2[:, 0] + cxp, xyz2[:, 1] + cyp], 1) * scale

        f1_2d, f1_3d = clfm(P, "core.clfm_fnet", uv1, f1_2d, f1_3d, index_impl)
        f2_2d, f2_3d = clfm(P, "core.clfm_fnet", uv2, f2_2d, f2_3d, index_impl)
        fc_2d, fc_3d = clfm(P, "core.clfm_cnet", uv1, fc_2d, fc_3d, index_impl)

        fc_2d = F.conv2d(fc_2d, P[b2 + ".cnet_aligner.weight"], P[b2 + ".cnet_aligner.bias"])
        h_2d, x_2d = torch.tanh(fc_2d[:, :128]), torch.relu(fc_2d[:, 128:])
        fc_3d = F.conv1d(fc_3d, P[b3 + ".cnet_aligner.weight"], P[b3 + ".cnet_aligner.bias"])
        h_3d, x_3d = torch.tanh(fc_3d[:, :128]), torch.relu(fc_3d[:, 128:])

        pyr2d = corr2d_build(P, b2 + ".correlation", f1_2d, f2_2d)
        pyr3d = corr3d_build(f1_3d, f2_3d, xyzs2, index_impl)
        nbr = knn(xyz1, xyz1, 32, index_impl)

        grid = mesh_grid(B, h8, w8)
        flow2d = torch.zeros_like(grid)
        flow3d = torch.zeros_like(xyz1)
        xyzs2_warp = xyzs2
        preds2d, preds3d = [], []
        for it in range(n_iters):
            if it > 0:
                xyzs2_warp = [backwarp_3d(xyz1, x2, flow3d, index_impl) for x2 in xyzs2]
            c2d = corr2d_lookup(pyr2d, grid + flow2d)
            c3d = corr3d_lookup(P, b3 + ".correlation", xyz1, xyzs2_warp, pyr3d, index_impl)
            c2d, c3d = clfm(P, "core.clfm_corr", uv1, c2d, c3d, index_impl)
            m2d = motion_encoder2d(P, b2 + ".motion_encoder", flow2d, c2d)
            m3d = motion_encoder3d(P, b3 + ".motion_encoder", xyz1, flow3d, c3d, nbr)
            m2d, m3d = clfm(P, "core.clfm_motion", uv1, m2d, m3d, index_impl)
            h_2d = gru2d(P, b2 + ".gru", h_2d, torch.cat([x_2d, m2d], 1))
            h_3d = gru3d(P, b3 + ".gru", xyz1, h_3d, torch.cat([x_3d, m3d], 1), nbr)

            d2 = F.relu(_conv(P, b2 + ".flow_head.conv1", h_2d, 1))
            flow2d = flow2d + torch.nan_to_num(_conv(P, b2 + ".flow_head.conv2", d2, 1))
            last = it == n_iters - 1
            if last or all_iters:
                m = F.relu(_conv(P, b2 + ".convex_upsampler.mask.0", h_2d, 1))
                m = 0.25 * _conv(P, b2 + ".convex_upsampler.mask.2", m, 0)
                preds2d.append(convex_upsample(flow2d, m))
            flow3d = flow3d + flow_head3d(P, b3 + ".flow_head", xyz1, h_3d, nbr)
            if last or all_iters:
                preds3d.append(knn_interpolation(xyz1, flow3d, pc1, index_impl))

        # ---- wrapper tail (camliraft.py:68-73)
        preds2d = [p[..., pad[2]:p.shape[-2] - pad[3], pad[0]:p.shape[-1] - pad[1]] for p in preds2d]
        base = paral2persp(pc1, *cam)
        preds3d = [paral2persp(pc1 + p, *cam) - base for p in preds3d]
        out = {"flow_2d": preds2d[-1], "flow_3d": preds3d[-1]}
        if all_iters:
            out["flow_2d_preds"], out["flow_3d_preds"] = preds2d, preds3d
        return out


# ------------------------------------------------------------------ synthetic inputs
def synthetic_inputs(B=1, H=540, W=960, N=8192, seed=0):
    """The measurement generator of SURVEY 8(d) / BASELINE.md: random RGB, points from
    u~U[0,W-1], v~U[0,H-1], z~U[5,35] back-projected with f=1050; pc2 = pc1 + N(0,0.05^2), permuted."""
    g = torch.Generator().manual_seed(seed)
    f, cx, cy = 1050.0, (W - 1) / 2.0, (H - 1) / 2.0
    images = torch.randint(0, 256, (B, 6, H, W), generator=g).float()
    u = torch.rand((B, N), generator=g) * (W - 1)
    v = torch.rand((B, N), generator=g) * (H - 1)
    z = torch.rand((B, N), generator=g) * 30.0 + 5.0
    pc1 = torch.stack([(u - cx) * z / f, (v - cy) * z / f, z], 1)
    pc2 = pc1 + torch.randn(pc1.shape, generator=g) * 0.05
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    pc2 = torch.gather(pc2, 2, perm[:, None, :].expand(B, 3, N))
    intr = torch.tensor([[f, cx, cy]]).repeat(B, 1)
    return {"images": images, "pcs": torch.cat([pc1, pc2], 1), "intrinsics": intr}
