"""The fused 2-D/3-D recurrent core of CamLiRAFT (reference models/camliraft_core.py:33-145):
schedules the two branches and the CLFM fusion sites around the hot loop."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

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

    def forward(self, image1, image2, pc1, pc2, camera_info):
        cfgs, b2, b3 = self.cfgs, self.branch_2d, self.branch_3d
        xyzs1, xyzs2, _, _ = build_pc_pyramid(pc1, pc2, [4096, 2048, 1024, 512, 256])

        feat1_2d, feat2_2d, featc_2d = b2.fnet(image1), b2.fnet(image2), b2.cnet(image1)
        feat1_3d = b3.fnet(xyzs1[:3])[2]
        feat2_3d = b3.fnet(xyzs2[:3])[2]
        featc_3d = b3.cnet(xyzs1[:3])[2]

        xyzs1, xyzs2 = xyzs1[2:], xyzs2[2:]        # working pyramid 2048 / 1024 / 512 / 256
        xyz1, xyz2 = xyzs1[0], xyzs2[0]

        # projected point positions on the 1/8 feature grid
        sh, sw = camera_info["sensor_h"], camera_info["sensor_w"]
        fh, fw = feat1_2d.shape[-2:]
        sx, sy = (fw - 1) / (sw - 1), (fh - 1) / (sh - 1)
        uv1, uv2 = project_pc2image(xyz1, camera_info), project_pc2image(xyz2, camera_info)
        uv1 = torch.stack([uv1[:, 0] * sx, uv1[:, 1] * sy], dim=1)
        uv2 = torch.stack([uv2[:, 0] * sx, uv2[:, 1] * sy], dim=1)

        # channel-last point features from here on; pixel -> nearest projected point tables are
        # computed once per cloud (the reference repeats the search at every fusion site and iteration)
        feat1_3d, feat2_3d, featc_3d = ops.rows_of(feat1_3d), ops.rows_of(feat2_3d), ops.rows_of(featc_3d)
        nn1 = ops.nearest_point_2d(uv1, fh, fw)
        if cfgs.fuse_fnet:
            nn2 = ops.nearest_point_2d(uv2, fh, fw)
            feat1_2d, feat1_3d = self.clfm_fnet.forward_rows(uv1, feat1_2d, feat1_3d, nn1)
            feat2_2d, feat2_3d = self.clfm_fnet.forward_rows(uv2, feat2_2d, feat2_3d, nn2)
        if cfgs.fuse_cnet:
            featc_2d, featc_3d = self.clfm_cnet.forward_rows(uv1, featc_2d, featc_3d, nn1)

        h_2d, x_2d = torch.split(b2.cnet_aligner(featc_2d), [128, 128], dim=1)
        h_2d, x_2d = torch.tanh(h_2d), torch.relu(x_2d)
        hx_3d = F.linear(featc_3d, b3.cnet_aligner.weight.flatten(1), b3.cnet_aligner.bias)
        h_3d, x_3d = torch.tanh(hx_3d[..., :128]), torch.relu(hx_3d[..., 128:])

        b2.correlation.build_cost_volume_pyramid(feat1_2d, feat2_2d)
        b3.correlation.build_cost_volume_pyramid(ops.cf_of(feat1_3d), ops.cf_of(feat2_3d), xyzs2)
        nbr = k_nearest_neighbor(xyz1, xyz1, k=32)

        n_iters = cfgs.n_iters_train if self.training else cfgs.n_iters_eval
        every = self.training if self.all_predictions is None else self.all_predictions
        B, _, H, W = image1.shape
        grid = mesh_grid(B, H // 8, W // 8, device=image1.device)
        flow_2d = torch.zeros_like(grid)
        flow_3d = torch.zeros_like(xyz1)
        xyzs2_warp = xyzs2
        dw_cache = {}                  # iteration-invariant WeightNet outputs of the PointConvDW layers
        preds_2d, preds_3d = [], []
        for it in range(n_iters):
            if it > 0:
                flow_2d, flow_3d = flow_2d.detach(), flow_3d.detach()
                xyzs2_warp = warp_pyramid(xyz1, xyzs2, flow_3d)

            corr_2d = b2.correlation(grid + flow_2d)
            corr_3d = b3.correlation.forward_rows(xyz1, xyzs2_warp)
            if cfgs.fuse_corr:
                corr_2d, corr_3d = self.clfm_corr.forward_rows(uv1, corr_2d, corr_3d, nn1)

            motion_2d = b2.motion_encoder(flow_2d, corr_2d)
            motion_3d = b3.motion_encoder.forward_rows(xyz1, ops.rows_of(flow_3d), corr_3d, nbr, dw_cache)
            if cfgs.fuse_motion:
                motion_2d, motion_3d = self.clfm_motion.forward_rows(uv1, motion_2d, motion_3d, nn1)

            h_2d = b2.gru(h=h_2d, x=torch.cat([x_2d, motion_2d], dim=1))
            h_3d = b3.gru.forward_rows(xyz1, h_3d, torch.cat([x_3d, motion_3d], dim=-1), nbr, dw_cache)
            if cfgs.fuse_hidden:
                h_2d, h_3d = self.clfm_hidden.forward_rows(uv1, h_2d, h_3d, nn1)

            flow_2d = flow_2d + b2.flow_head(h_2d)
            flow_3d = flow_3d + ops.cf_of(b3.flow_head.forward_rows(xyz1, h_3d, nbr, dw_cache))
            if every or it == n_iters - 1:
                preds_2d.append(b2.convex_upsampler(h_2d, flow_2d))
                preds_3d.append(knn_interpolation(xyz1, flow_3d, pc1, k=3))
        return preds_2d, preds_3d
