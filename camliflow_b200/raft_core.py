"""2-D branch of CamLiRAFT (reference models/raft_core.py): ResNet-50 stage-1/2 encoder,
all-pairs 4-D correlation (build + 4-level 9x9 lookup), separable ConvGRU, motion encoder,
flow head and convex upsampler.  Parameter names follow the reference so checkpoints load.

The reference inherits its encoder from mmdet's ResNet (absent here); the stem and
the two bottleneck stages are written out below with the same state_dict keys
(`conv1`, `bn1`, `layerN.M.{conv,bn}{1,2,3}`, `layerN.0.downsample.{0,1}`, `align.conv_fn`)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, tc
from .mlp import Conv2dNormRelu
from .utils import convex_upsample, mesh_grid


def _folded(owner, name, conv, bn):
    """(weight, bias) of `conv` followed by an eval-mode BatchNorm, folded once and cached on `owner`
    (rebuilt when a parameter changes).  Inference only."""
    key = tuple((p.data_ptr(), p._version) for p in (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var))
    cache = owner.__dict__.setdefault("_fold_cache", {})
    hit = cache.get(name)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = conv.weight * scale[:, None, None, None]
            if conv.weight.is_contiguous(memory_format=torch.channels_last):
                w = w.contiguous(memory_format=torch.channels_last)
            hit = (key, w, (bn.bias - bn.running_mean * scale).contiguous())
        cache[name] = hit
    return hit[1], hit[2]


def _fused_inference(x):
    """The encoder's BatchNorms always run on their running statistics (norm_eval), so without autograd
    every conv+BN(+residual)+ReLU group is ONE cuDNN call with a bias / add / ReLU epilogue instead of
    conv, batch-norm, add and ReLU kernels (the BN pass alone re-read and re-wrote every activation)."""
    return x.is_cuda and not torch.is_grad_enabled()


class _Bottleneck(nn.Module):
    def __init__(self, c_in, c_mid, stride):
        super().__init__()
        c_out = 4 * c_mid
        self.conv1 = nn.Conv2d(c_in, c_mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(c_mid)
        self.conv2 = nn.Conv2d(c_mid, c_mid, 3, stride, 1, bias=False)     # stride on the 3x3 ('pytorch' style)
        self.bn2 = nn.BatchNorm2d(c_mid)
        self.conv3 = nn.Conv2d(c_mid, c_out, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(c_out)
        self.downsample = None
        if stride != 1 or c_in != c_out:
            self.downsample = nn.Sequential(nn.Conv2d(c_in, c_out, 1, stride, bias=False), nn.BatchNorm2d(c_out))

    def _forward_fused(self, x):
        """conv+BN(+residual)+ReLU groups as single kernels on the tcgen05 implicit-GEMM kernel (fp32-accurate 3xTF32)."""
        # (the two stride-2 convolutions of layer2.0 included: the kernel's activation tensor map then traverses the
        # input with a pixel stride of 2)
        y = tc.conv2d(x, self.conv1, "relu", bn=self.bn1)
        y = tc.conv2d(y, self.conv2, "relu", bn=self.bn2)
        if self.downsample is not None:
            x = tc.conv2d(x, self.downsample[0], None, bn=self.downsample[1])
        return tc.conv2d(y, self.conv3, "relu", bn=self.bn3, residual=x)

    def forward(self, x):
        if _fused_inference(x):
            return self._forward_fused(x)
        if tc._train_route(x):
            # training: the BatchNorms still run on their running statistics (norm_eval), so they fold into the convolutions
            # differentiably; every stride-1 layer then runs forward AND backward on the tensor-core kernels (grad.DenseFn)
            return self._forward_fused(x)
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = F.relu(self.bn2(self.conv2(y)), inplace=True)
        y = self.bn3(self.conv3(y))
        if self.downsample is not None:
            x = self.downsample(x)
        return F.relu(y + x, inplace=True)


class Encoder2D(nn.Module):
    """1/8-resolution, 128-channel image encoder (raft_core.py:10-38); BatchNorm always runs on
    its running statistics (`norm_eval=True`, raft_core.py:18)."""

    def __init__(self, depth=50, pretrained=None):
        super().__init__()
        if depth != 50:
            raise NotImplementedError("Encoder2D: only the ResNet-50 backbone of the reference configs")
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = nn.Sequential(_Bottleneck(64, 64, 1), _Bottleneck(256, 64, 1), _Bottleneck(256, 64, 1))
        self.layer2 = nn.Sequential(_Bottleneck(256, 128, 2), *[_Bottleneck(512, 128, 1) for _ in range(3)])
        self.feat_dim = 512
        self.align = Conv2dNormRelu(self.feat_dim, 128)

    def train(self, mode=True):
        super().train(mode)
        for m in self.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.eval()
        return self

    def forward(self, x):
        if _fused_inference(x) and tc.fused(x):
            # stem convolution + folded BatchNorm + ReLU + max-pool: one CUDA-core kernel (3 input channels)
            w, b = _folded(self, "stem", self.conv1, self.bn1)
            cache = self.__dict__.setdefault("_stem_cache", {})
            if cache.get("key") is not w:
                cache["key"], cache["w"] = w, w.permute(0, 2, 3, 1).contiguous()
            x = ops.stem_conv_pool(x.permute(0, 2, 3, 1), cache["w"], b).permute(0, 3, 1, 2)
        else:
            x = F.max_pool2d(F.relu(self.bn1(self.conv1(x)), inplace=True), 3, 2, 1)
        return self.align(self.layer2(self.layer1(x)))


class Correlation2D(nn.Module):
    """All-pairs correlation pyramid + windowed bilinear lookup (raft_core.py:41-107).  Stateful
    like the reference: `build_cost_volume_pyramid` caches the volume for later `forward` calls."""

    def __init__(self, num_levels=4, radius=4):
        super().__init__()
        self.num_levels = num_levels
        self.radius = radius
        self.fnet_aligner = nn.Conv2d(128, 256, kernel_size=1)
        self.cost_volume_pyramid = None

    def build_cost_volume_pyramid(self, fmap1, fmap2):
        fmap1 = tc.conv2d(fmap1.float(), self.fnet_aligner)
        fmap2 = tc.conv2d(fmap2.float(), self.fnet_aligner)
        self.cost_volume_pyramid = ops.corr2d_build(fmap1, fmap2, self.num_levels)

    def forward(self, coords):
        return ops.corr2d_lookup(self.cost_volume_pyramid, coords, self.radius)


class GRU2D(nn.Module):
    """Separable ConvGRU: a 1x5 pass then a 5x1 pass (raft_core.py:110-139)."""

    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        c = hidden_dim + input_dim
        self.convz1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convr1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convq1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convz2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))
        self.convr2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))
        self.convq2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))

    def _half(self, h, x, convz, convr, convq):
        hx = torch.cat([h, x], dim=1)
        if tc._train_route(hx):
            # training on the GPU: z and r read the same input -- ONE 256-output convolution (the concatenation of the two
            # weights is differentiable), forward and backward on the tensor-core kernels, sigmoid in the epilogue
            zr = tc.conv2d_weights(hx, torch.cat([convz.weight, convr.weight], 0), torch.cat([convz.bias, convr.bias], 0),
                                   convz.padding, "sigmoid")
            z, r = torch.split(zr, [h.shape[1], h.shape[1]], dim=1)
        else:
            z = tc.conv2d(hx, convz, "sigmoid")
            r = tc.conv2d(hx, convr, "sigmoid")
        q = tc.conv2d(torch.cat([r * h, x], dim=1), convq, "tanh")
        return (1 - z) * h + z * q

    def _merged_zr(self, convz, convr):
        """Weights of the z and r convolutions stacked into one 256-output convolution (they read the
        same input); rebuilt only when a parameter changes."""
        key = tuple((p.data_ptr(), p._version) for p in (convz.weight, convz.bias, convr.weight, convr.bias))
        cache = self.__dict__.setdefault("_zr_cache", {})
        hit = cache.get(id(convz))
        if hit is None or hit[0] != key:
            w = torch.cat([convz.weight, convr.weight], 0)
            if convz.weight.is_contiguous(memory_format=torch.channels_last):
                w = w.contiguous(memory_format=torch.channels_last)
            hit = (key, w, torch.cat([convz.bias, convr.bias], 0))
            cache[id(convz)] = hit
        return hit[1], hit[2]

    def _half_fused(self, h, x, convz, convr, convq, last):
        """One merged z|r convolution, one gate kernel (sigmoids, r*h, [r*h | x] assembly), the q
        convolution and one update kernel -- 5 launches instead of ~12."""
        w, b = self._merged_zr(convz, convr)
        zr = tc.conv2d_weights(torch.cat([h, x], dim=1), w, b, convz.padding)
        z, rhx = ops.gru_gate(zr, h, x)
        return ops.gru_update(z, h, tc.conv2d(rhx, convq), fix_nonfinite=last)

    def _split_weights(self, convz, convr, convq, n_h, n_static):
        """The z|r and q convolutions split by input-channel group [h | x_static | x_dynamic], as tf32 hi/lo
        OHWI matrices: the x_static part (the context features, identical in every refinement iteration) is
        convolved once per forward; the per-iteration parts read the recurrent buffer [h | x_dynamic | r*h]
        (z|r from its first two groups, q from its last two, hence the [x_dynamic | r*h] column order of w_q)."""
        key = tuple((p.data_ptr(), p._version) for p in (convz.weight, convz.bias, convr.weight, convr.bias,
                                                        convq.weight, convq.bias)) + (n_h, n_static)
        cache = self.__dict__.setdefault("_split_cache", {})
        hit = cache.get(id(convz))
        if hit is None or hit[0] != key:
            with torch.no_grad():
                wzr, bzr = self._merged_zr(convz, convr)
                lo, hi = n_h, n_h + n_static
                ohwi = lambda w: w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()     # noqa: E731
                keep = [torch.cat([wzr[:, :lo], wzr[:, hi:]], 1), wzr[:, lo:hi],
                        torch.cat([convq.weight[:, hi:], convq.weight[:, :lo]], 1), convq.weight[:, lo:hi]]
                mats = [ops.tc_weight([t], lambda t=t: (ohwi(t), None))[:2] for t in keep]
                hit = (key, keep, mats, bzr.contiguous(), convq.bias.detach().contiguous())
            cache[id(convz)] = hit
        return hit[2], hit[3], hit[4]

    @staticmethod
    def _state_buffers(cache, h, X):
        """[h | x_dynamic | r*h] and z, channel-last, one pair per forward pass (kept in the caller's cache)."""
        bufs = cache.get("gru2d_buffers")
        if bufs is None:
            B, C, H, W = h.shape
            bufs = (torch.empty((B, H, W, C + X + C), dtype=torch.float32, device=h.device),
                    torch.empty((B, H, W, C), dtype=torch.float32, device=h.device))
            cache["gru2d_buffers"] = bufs
        return bufs

    def dynamic_slot(self, cache, h, X):
        """The x_dynamic group of the state buffer as a logical [B,X,H,W] map: the layer that produces the motion features
        writes them here and forward_split() finds them in place (no copy kernel on the image chain)."""
        C = h.shape[1]
        return self._state_buffers(cache, h, X)[0][..., C:C + X].permute(0, 3, 1, 2)

    def forward_split(self, h, x_static, x_dynamic, cache):
        """forward(h, cat([x_static, x_dynamic])) as four tensor-core launches per call and nothing else:
        * the x_static contributions to the z, r and q pre-activations come from `cache` (a dict owned by the
          caller for one forward pass) and are added in the convolution epilogues (384 -> 256 input channels);
        * the gate and update arithmetic runs in the epilogues too (camli_conv_gemm_fused): the z|r convolution
          writes z and r*h, the q convolution writes h' over h -- no gate, update or concatenation kernels;
        * h, x_dynamic and r*h live in one channel-last buffer [h | x_dynamic | r*h]; the returned state is a
          view of its first group and is recognised (no copy) when it comes back in the next iteration."""
        B, C, H, W = h.shape
        S, X = x_static.shape[1], x_dynamic.shape[1]
        Q, Z = self._state_buffers(cache, h, X)
        h_rows = Q[..., :C]
        if h.data_ptr() != h_rows.data_ptr():
            h_rows.copy_(h.permute(0, 2, 3, 1))
        if x_dynamic.data_ptr() != Q[..., C:C + X].data_ptr():      # (a producer that was handed dynamic_slot() wrote it in place)
            Q[..., C:C + X].copy_(x_dynamic.permute(0, 2, 3, 1))
        xs_rows = x_static.permute(0, 2, 3, 1)
        if not ops.conv_gemm_ok(xs_rows):
            xs_rows = xs_rows.contiguous()
        for half, (cz, cr, cq) in enumerate(((self.convz1, self.convr1, self.convq1), (self.convz2, self.convr2, self.convq2))):
            (w_zr, w_zr_s, w_q, w_q_s), b_zr, b_q = self._split_weights(cz, cr, cq, C, S)
            kh, kw = cz.kernel_size
            ctx = cache.get(("gru2d", half))
            if ctx is None:
                ctx = (ops.conv_gemm(xs_rows, *w_zr_s, kh, kw), ops.conv_gemm(xs_rows, *w_q_s, kh, kw))
                cache[("gru2d", half)] = ctx
            ops.conv_gemm(Q[..., :C + X], *w_zr, kh, kw, b_zr, "gru_gate", residual=ctx[0], out=Z, out2=Q[..., C + X:],
                          aux1=h_rows, split=C)
            ops.conv_gemm(Q[..., C:], *w_q, kh, kw, b_q, "gru_update_fix" if half == 1 else "gru_update", residual=ctx[1],
                          out=h_rows, aux1=Z, aux2=h_rows)
        return h_rows.permute(0, 3, 1, 2)

    def forward(self, h, x):
        if h.is_cuda and not (torch.is_grad_enabled() and (h.requires_grad or x.requires_grad or self.convz1.weight.requires_grad)):
            h = self._half_fused(h, x, self.convz1, self.convr1, self.convq1, False)
            return self._half_fused(h, x, self.convz2, self.convr2, self.convq2, True)
        h = self._half(h, x, self.convz1, self.convr1, self.convq1)
        h = self._half(h, x, self.convz2, self.convr2, self.convq2)
        return torch.nan_to_num(h)


class MotionEncoder2D(nn.Module):
    """raft_core.py:142-166."""

    def __init__(self, corr_levels, corr_radius):
        super().__init__()
        corr_planes = corr_levels * (2 * corr_radius + 1) ** 2
        self.conv_c1 = nn.Conv2d(corr_planes, 256, kernel_size=1, padding=0)
        self.conv_c2 = nn.Conv2d(256, 192, kernel_size=3, padding=1)
        self.conv_f1 = nn.Conv2d(2, 128, kernel_size=7, padding=3)
        self.conv_f2 = nn.Conv2d(128, 64, kernel_size=3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, kernel_size=3, padding=1)

    def flow_features(self, flow):
        """The flow half of the encoder (conv_f1, conv_f2), written into its slice of the [corr | flow] feature
        buffer.  It depends on the flow alone, so the fused core issues it while the point branch is still busy
        with its correlation lookup; pass the result to forward(..., cf=)."""
        B, _, H, W = flow.shape
        cf = torch.empty((B, H, W, 192 + 64), dtype=torch.float32, device=flow.device)
        tc.conv2d(tc.conv2d(flow, self.conv_f1, "relu"), self.conv_f2, "relu", out=cf[..., 192:])
        # the encoder's output buffer [out | flow] with its flow columns filled: also off the image chain
        n_out = self.conv.out_channels
        mf = torch.empty((B, H, W, n_out + 2), dtype=torch.float32, device=flow.device)
        mf[..., n_out:].copy_(flow.permute(0, 2, 3, 1))
        return cf, mf

    def forward(self, flow, corr, cf=None):
        if tc.fused(corr):
            # every convolution with its ReLU in one kernel; the two branches write straight into the halves
            # of one channel-last buffer (no torch.cat)
            cf, mf = self.flow_features(flow) if cf is None else cf
            tc.conv2d(tc.conv2d(corr, self.conv_c1, "relu"), self.conv_c2, "relu", out=cf[..., :192])
            # the last convolution (ReLU + nan_to_num in its epilogue) and the flow land in one channel-last buffer
            # [out | flow]: no nan_to_num, cat or layout-conversion kernels
            tc.conv2d(cf.permute(0, 3, 1, 2), self.conv, "relu_fix", out=mf[..., :self.conv.out_channels])
            return mf.permute(0, 3, 1, 2)
        c = tc.conv2d(tc.conv2d(corr, self.conv_c1, "relu"), self.conv_c2, "relu")
        f = tc.conv2d(tc.conv2d(flow, self.conv_f1, "relu"), self.conv_f2, "relu")
        out = torch.nan_to_num(tc.conv2d(torch.cat([c, f], dim=1), self.conv, "relu"))
        return torch.cat([out, flow], dim=1)


class FlowHead2D(nn.Module):
    """raft_core.py:169-181."""

    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, kernel_size=3, padding=1)

    def forward(self, x):
        if tc.fused(x):
            return tc.conv2d(tc.conv2d(x, self.conv1, "relu"), self.conv2, "none_fix")      # nan_to_num in the epilogue
        return torch.nan_to_num(tc.conv2d(tc.conv2d(x, self.conv1, "relu"), self.conv2).float())


class ConvexUpsampler2D(nn.Module):
    """raft_core.py:184-197."""

    def __init__(self, input_dim):
        super().__init__()
        self.mask = nn.Sequential(nn.Conv2d(input_dim, 256, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(256, 64 * 9, 1, padding=0))

    def forward(self, h, flow):
        mask = tc.conv2d(tc.conv2d(h.float(), self.mask[0], "relu"), self.mask[2])
        return convex_upsample(flow, mask, mask_scale=0.25)


class RAFTCore(nn.Module):
    """The image-only RAFT (raft_core.py:200-270); also the `branch_2d` of CamLiRAFT."""

    def __init__(self, cfgs):
        super().__init__()
        self.cfgs = cfgs
        self.hidden_dim = self.context_dim = 128
        self.corr_levels = self.corr_radius = 4
        self.fnet = Encoder2D(cfgs.backbone.depth, cfgs.backbone.pretrained)
        self.cnet = Encoder2D(cfgs.backbone.depth, cfgs.backbone.pretrained)
        self.cnet_aligner = nn.Conv2d(128, 256, kernel_size=1)
        self.correlation = Correlation2D(self.corr_levels, self.corr_radius)
        self.motion_encoder = MotionEncoder2D(self.corr_levels, self.corr_radius)
        self.gru = GRU2D(hidden_dim=self.hidden_dim, input_dim=self.hidden_dim + 128)
        self.flow_head = FlowHead2D(self.hidden_dim)
        self.convex_upsampler = ConvexUpsampler2D(self.hidden_dim)

    def forward(self, image1, image2):
        self.correlation.build_cost_volume_pyramid(self.fnet(image1), self.fnet(image2))
        h, x = torch.split(tc.conv2d(self.cnet(image1), self.cnet_aligner), [self.hidden_dim, self.context_dim], dim=1)
        h, x = torch.tanh(h), torch.relu(x)
        B, _, H, W = image1.shape
        grid = mesh_grid(B, H // 8, W // 8, device=image1.device)
        flow = torch.zeros_like(grid)
        preds = []
        n_iters = self.cfgs.n_iters_train if self.training else self.cfgs.n_iters_eval
        for _ in range(n_iters):
            flow = flow.detach()
            corr = self.correlation(grid + flow)
            h = self.gru(h, torch.cat([x, self.motion_encoder(flow, corr)], dim=1))
            flow = flow + self.flow_head(h)
            preds.append(self.convex_upsampler(h, flow))
        return preds
