"""The fused 2-D/3-D recurrent core of CamLiRAFT (reference models/camliraft_core.py:33-145):
schedules the two branches and the CLFM fusion sites around the hot loop."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, tc

from .camliraft_l_core import CamLiRAFT_L_Core, warp_pyramid
from .clfm import CLFM
from .raft_core import RAFTCore
from .utils import build_pc_pyramid, k_nearest_neighbor, knn_interpolation, mesh_grid, project_pc2image


class CamLiRAFT_Core(nn.Module):
    def __init__(self, cfgs):
        super().__init__()
        self.cfgs = cfgs
        self.corr_levels = self.corr_radius = 4
        self.branch_2d = RAFTCore(cfgs)            # attribute names are checkpoint keys
        self.branch_3d = CamLiRAFT_L_Core(cfgs)
        if cfgs.fuse_fnet:
            self.clfm_fnet = CLFM(128, 128, norm="batch_norm")
        if cfgs.fuse_cnet:
            self.clfm_cnet = CLFM(128, 128, norm="batch_norm")
        if cfgs.fuse_corr:
            self.clfm_corr = CLFM(81 * 4, 128)
        if cfgs.fuse_motion:
            self.clfm_motion = CLFM(128, 128)
        if cfgs.fuse_hidden:
            self.clfm_hidden = CLFM(128, 128)
        # Inference only needs the last refinement; the reference also materialises the
        # up-sampled prediction of every earlier iteration (needed by the training loss).
        self.all_predictions = None   # None: follow self.training
        self.two_streams = True       # overlap the image and the point branch on two CUDA streams (inference)

    def forward(self, image1, image2, pc1, pc2, camera_info):
        cfgs, b2, b3 = self.cfgs, self.branch_2d, self.branch_3d
        par = _TwoStreams(self.two_streams and image1.is_cuda and not torch.is_grad_enabled())

        # ---- encoders: the image branch and the point branch (FPS pyramid + PointConv encoders) are
        # independent until the first fusion site
        def encode_2d():
            # both frames through the feature encoder as one batch (its BatchNorms use running statistics,
            # so samples stay independent)
            f12 = b2.fnet(torch.cat([image1, image2], dim=0))
            return f12[:image1.shape[0]], f12[image1.shape[0]:], b2.cnet(image1)

        fh, fw = image1.shape[-2] // 8, image1.shape[-1] // 8        # the 1/8 feature grid of the image encoders

        def geometry(xyz1, xyz2):
            """Everything that depends on the point COORDINATES only: projected positions on the feature grid, the pixel ->
            nearest projected point tables of CLFM (computed once per cloud; the reference repeats the search at every
            fusion site and iteration) and the k = 32 self-neighbour table of the update block."""
            sh, sw = camera_info["sensor_h"], camera_info["sensor_w"]
            sx, sy = (fw - 1) / (sw - 1), (fh - 1) / (sh - 1)
            uv1, uv2 = project_pc2image(xyz1, camera_info), project_pc2image(xyz2, camera_info)
            uv1 = torch.stack([uv1[:, 0] * sx, uv1[:, 1] * sy], dim=1)
            uv2 = torch.stack([uv2[:, 0] * sx, uv2[:, 1] * sy], dim=1)
            nn1 = ops.nearest_point_2d(uv1, fh, fw)
            nn2 = ops.nearest_point_2d(uv2, fh, fw) if cfgs.fuse_fnet else None
            return uv1, uv2, nn1, nn2, k_nearest_neighbor(xyz1, xyz1, k=32)

        def encode_3d():
            xyzs1, xyzs2, _, _ = build_pc_pyramid(pc1, pc2, [4096, 2048, 1024, 512, 256])
            # (forked as soon as the pyramid exists: ~110 us of searches and glue beside the encoders instead of behind them)
            geo = par.fork(lambda: geometry(xyzs1[2], xyzs2[2]), "geo")
            B = pc1.shape[0]
            # neighbour tables depend on the geometry only: ONE search per level for both clouds, shared by the feature
            # and the context encoder of frame 1 (the reference searches again in each of the 3 encoder passes)
            both = [torch.cat([a, b], dim=0) for a, b in zip(xyzs1[:3], xyzs2[:3])]
            tables = b3.fnet.neighbor_tables(both)
            t1 = [t[:B] for t in tables]
            if not (self.training and torch.is_grad_enabled()):
                # both frames through the feature encoder as one batch (eval-mode BatchNorm: samples stay independent)
                f12 = b3.fnet(both, tables)[2]
                feats = f12[:B], f12[B:], b3.cnet(xyzs1[:3], t1)[2]
            else:
                feats = b3.fnet(xyzs1[:3], t1)[2], b3.fnet(xyzs2[:3], [t[B:] for t in tables])[2], b3.cnet(xyzs1[:3], t1)[2]
            return xyzs1, xyzs2, [ops.rows_of(f) for f in feats], geo      # channel-last point features from here on

        (feat1_2d, feat2_2d, featc_2d), (xyzs1, xyzs2, (feat1_3d, feat2_3d, featc_3d), geo) = par.run(encode_2d, encode_3d)

        xyzs1, xyzs2 = xyzs1[2:], xyzs2[2:]        # working pyramid 2048 / 1024 / 512 / 256
        xyz1, xyz2 = xyzs1[0], xyzs2[0]
        assert tuple(feat1_2d.shape[-2:]) == (fh, fw), (tuple(feat1_2d.shape), fh, fw)
        uv1, uv2, nn1, nn2, nbr = geo.join()

        # The fusion sites in front of the loop (frame 1, frame 2, context) are independent of each other: the second and
        # third run on streams of their own beside the first (each a chain of ~10 short kernels)
        site_2 = site_c = None
        if cfgs.fuse_fnet:
            site_2 = par.fork(lambda: self.clfm_fnet.forward_rows(uv2, feat2_2d, feat2_3d, nn2), "site2")
        if cfgs.fuse_cnet:
            site_c = par.fork(lambda: self.clfm_cnet.forward_rows(uv1, featc_2d, featc_3d, nn1), "sitec")
        if cfgs.fuse_fnet:
            feat1_2d, feat1_3d = self.clfm_fnet.forward_rows(uv1, feat1_2d, feat1_3d, nn1, par)
            feat2_2d, feat2_3d = site_2.join()
        if cfgs.fuse_cnet:
            featc_2d, featc_3d = site_c.join()

        def init_2d():
            conv = b2.cnet_aligner
            if tc.fused(featc_2d):
                # tanh / ReLU of the two halves in the epilogues of two 128-column launches (a strided elementwise kernel on a
                # channel slice of the 256-channel output costs more than the convolution itself)
                w_hi, w_lo, bias = ops.tc_weight([conv.weight, conv.bias], lambda: (conv.weight.flatten(1), conv.bias))
                rows = ops.nhwc_rows(featc_2d)
                h = ops.conv_gemm(rows, w_hi[:128], w_lo[:128], 1, 1, bias[:128], "tanh")
                x = ops.conv_gemm(rows, w_hi[128:], w_lo[128:], 1, 1, bias[128:], "relu")
                # (the volume build last: its persistent all-pairs GEMM holds every SM for ~190 us)
                b2.correlation.build_cost_volume_pyramid(feat1_2d, feat2_2d)
                return ops.nchw_view(h), ops.nchw_view(x)
            h, x = torch.split(tc.conv2d(featc_2d, conv), [128, 128], dim=1)
            b2.correlation.build_cost_volume_pyramid(feat1_2d, feat2_2d)
            return torch.tanh(h), torch.relu(x)

        def init_3d():
            hx = tc.linear(featc_3d, b3.cnet_aligner.weight, b3.cnet_aligner.bias)
            b3.correlation.build_cost_volume_pyramid(ops.cf_of(feat1_3d), ops.cf_of(feat2_3d), xyzs2)
            return torch.tanh(hx[..., :128]), torch.relu(hx[..., 128:])

        (h_2d, x_2d), (h_3d, x_3d) = par.run(init_2d, init_3d)

        n_iters = cfgs.n_iters_train if self.training else cfgs.n_iters_eval
        every = self.training if self.all_predictions is None else self.all_predictions
        B, _, H, W = image1.shape
        grid = mesh_grid(B, H // 8, W // 8, device=image1.device)
        flow_2d = torch.zeros_like(grid)
        flow_3d = torch.zeros_like(xyz1)
        dw_cache = {}                  # iteration-invariant WeightNet outputs of the PointConvDW layers
        gru_cache = {}                 # iteration-invariant context contributions to the ConvGRU pre-activations
        preds_2d, preds_3d = [], []
        # Software pipelining of the point branch (two-stream inference only): its back-warp + correlation lookup of
        # iteration i+1 depend on flow_3d alone, and the point branch finishes an iteration well before the image branch
        # (ConvGRU + flow head); issued right behind the point update they are done when the image chain comes round,
        # instead of standing in front of the correlation fusion (~60 us of every iteration).
        pipe = os.environ.get("CAMLI_PIPELINE_3D", "1")       # "0": program order; "force": also without streams (host tests)
        ahead = (par.enabled or pipe == "force") and not cfgs.fuse_hidden and pipe != "0"
        corr_3d_next = None
        for it in range(n_iters):
            if it > 0:
                flow_2d, flow_3d = flow_2d.detach(), flow_3d.detach()

            def corr_3d_fn(flow=None):
                flow = flow_3d if flow is None else flow
                warped = warp_pyramid(xyz1, xyzs2, flow) if (it > 0 or flow is not flow_3d) else xyzs2
                return b3.correlation.forward_rows(xyz1, warped)

            # the flow half of the motion encoder only needs flow_2d: on a stream of its own it runs beside the
            # lookups and the correlation fusion instead of in front of them (2 of the ~25 launches of the image chain)
            if corr_3d_next is None:
                corr_2d, corr_3d = par.run(lambda: b2.correlation(grid + flow_2d), corr_3d_fn)
            else:
                corr_2d, corr_3d = b2.correlation(grid + flow_2d), corr_3d_next
            # (forked behind the lookup: its 7x7 convolution fills every SM and must not stand in front of the chain)
            cf_fork = par.fork(lambda: b2.motion_encoder.flow_features(flow_2d)) if tc.fused(flow_2d) else None
            if cfgs.fuse_corr:
                corr_2d, corr_3d = self.clfm_corr.forward_rows(uv1, corr_2d, corr_3d, nn1, par)
            cf_2d = cf_fork.join() if cf_fork is not None else None

            motion_2d, motion_3d = par.run(
                lambda: b2.motion_encoder(flow_2d, corr_2d, cf_2d),
                lambda: b3.motion_encoder.forward_rows(xyz1, ops.rows_of(flow_3d), corr_3d, nbr, dw_cache))
            if cfgs.fuse_motion:
                # the fused image features are the ConvGRU's x_dynamic: written straight into its state buffer
                slot = b2.gru.dynamic_slot(gru_cache, h_2d, motion_2d.shape[1]) if tc.fused(h_2d) and not cfgs.fuse_hidden else None
                motion_2d, motion_3d = self.clfm_motion.forward_rows(uv1, motion_2d, motion_3d, nn1, par, out_2d=slot)

            last = every or it == n_iters - 1

            def update_2d():
                if tc.fused(h_2d):
                    h = b2.gru.forward_split(h_2d, x_2d, motion_2d, gru_cache)
                else:
                    h = b2.gru(h=h_2d, x=torch.cat([x_2d, motion_2d], dim=1))
                flow = flow_2d + b2.flow_head(h)
                return h, flow, (b2.convex_upsampler(h, flow) if last and not cfgs.fuse_hidden else None)

            def update_3d():
                h = b3.gru.forward_rows(xyz1, h_3d, torch.cat([x_3d, motion_3d], dim=-1), nbr, dw_cache)
                if cfgs.fuse_hidden:
                    return h, None, None, None
                flow = flow_3d + ops.cf_of(b3.flow_head.forward_rows(xyz1, h, nbr, dw_cache))
                nxt = corr_3d_fn(flow.detach()) if ahead and it < n_iters - 1 else None
                return h, flow, (knn_interpolation(xyz1, flow, pc1, k=3) if last else None), nxt

            if cfgs.fuse_hidden:       # hidden-state fusion sits between the GRUs and the flow heads
                h_2d, h_3d = b2.gru(h=h_2d, x=torch.cat([x_2d, motion_2d], dim=1)), update_3d()[0]
                h_2d, h_3d = self.clfm_hidden.forward_rows(uv1, h_2d, h_3d, nn1, par)
                flow_2d = flow_2d + b2.flow_head(h_2d)
                flow_3d = flow_3d + ops.cf_of(b3.flow_head.forward_rows(xyz1, h_3d, nbr, dw_cache))
                up_2d = b2.convex_upsampler(h_2d, flow_2d) if last else None
                up_3d = knn_interpolation(xyz1, flow_3d, pc1, k=3) if last else None
            else:
                (h_2d, flow_2d, up_2d), (h_3d, flow_3d, up_3d, corr_3d_next) = par.run(update_2d, update_3d)
            if last:
                preds_2d.append(up_2d)
                preds_3d.append(up_3d)
        return preds_2d, preds_3d


class _TwoStreams:
    """Runs two independent pieces of the forward concurrently: the first on the current stream,
    the second on a side stream, joined before returning (fork/join is graph-capturable).  At batch 1
    every kernel of the update block is far too small to fill 148 SMs, so overlapping the image and
    the point branch is worth more than any single-kernel optimisation.  Tensors that cross streams
    are kept referenced by the caller until after the join, which is what the caching allocator needs."""
    _side = {}
    _aux = {}

    def __init__(self, enabled):
        self.enabled = enabled
        if enabled:
            dev = torch.cuda.current_device()
            if dev not in _TwoStreams._side:
                # the point branch is a chain of short kernels: give its CTAs precedence whenever SMs free up
                _TwoStreams._side[dev] = torch.cuda.Stream(dev, priority=int(os.environ.get("CAMLI_SIDE_PRIORITY", "-1")))
            self.side = _TwoStreams._side[dev]

    def fork(self, fn, key="aux"):
        """Starts `fn` on a further side stream (one per key) and returns a handle whose .join() makes the current
        stream wait for it and hands back fn's result.  For work off the critical chain of a branch: the flow half
        of the motion encoder, the 2-D alignment layer of a fusion site."""
        if not self.enabled or os.environ.get("CAMLI_AUX_STREAMS", "1") == "0":     # (A/B switch: inline, in program order)
            return _Joined(fn())
        dev = torch.cuda.current_device()
        if (dev, key) not in _TwoStreams._aux:
            _TwoStreams._aux[(dev, key)] = torch.cuda.Stream(dev, priority=0)      # below the two branch streams
        s = _TwoStreams._aux[(dev, key)]
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            out = fn()
        return _Joined(out, s)

    def run(self, fn_main, fn_side):
        if not self.enabled:
            return fn_main(), fn_side()
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            b = fn_side()
        a = fn_main()
        main.wait_stream(self.side)
        return a, b


class _Joined:
    """Result of _TwoStreams.fork: .join() orders the current stream after the forked work (fork / join on streams is
    graph-capturable; the result stays referenced by the caller past the join, as the caching allocator needs)."""

    def __init__(self, out, stream=None):
        self.out, self.stream = out, stream

    def join(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.stream = None
        return self.out
