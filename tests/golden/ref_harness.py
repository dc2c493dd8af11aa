"""Imports the REFERENCE model code from /root/reference (only available in the build
container) with stand-ins for the two third-party packages that are not installed:
mmdet's ResNet (rebuilt from torchvision layers with identical parameter names) and
mmcv's logger.  Used only to generate golden fixtures / to cross-check the oracle.
"""
import logging
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


class _AttrDict(dict):
    """Stand-in for the Hydra DictConfig the reference models read attributes from."""
    __getattr__ = dict.__getitem__

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = _AttrDict(v) if isinstance(v, dict) else v


def _install_stubs():
    if "mmdet.models.backbones" in sys.modules:
        return
    import torchvision

    class ResNet(nn.Module):   # signature of mmdet 2.14 ResNet as used by models/raft_core.py:10-22
        def __init__(self, depth=50, num_stages=2, strides=(1, 2), dilations=(1, 1), out_indices=(1,),
                     norm_eval=True, with_cp=False, init_cfg=None):
            super().__init__()
            assert depth == 50 and num_stages == 2
            net = torchvision.models.resnet50(weights=None)
            self.conv1, self.bn1, self.relu, self.maxpool = net.conv1, net.bn1, net.relu, net.maxpool
            self.layer1, self.layer2 = net.layer1, net.layer2
            self.feat_dim = 512
            self.norm_eval = norm_eval

        def init_weights(self):
            pass

        def forward(self, x):
            x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
            return (self.layer2(self.layer1(x)),)

        def train(self, mode=True):
            super().train(mode)
            if mode and self.norm_eval:
                for m in self.modules():
                    if isinstance(m, nn.modules.batchnorm._BatchNorm):
                        m.eval()
            return self

    for name in ("mmdet", "mmdet.models", "mmdet.models.backbones", "mmcv", "mmcv.utils", "mmcv.utils.logging"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["mmdet.models.backbones"].ResNet = ResNet
    sys.modules["mmcv.utils.logging"].get_logger = lambda name, *a, **k: logging.getLogger(name)


def load_reference():
    """Returns the reference's `models` package."""
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models   # noqa: E402  (the reference package)
    return models


def camliraft_cfg(n_iters=12):
    return _AttrDict(dict(name="camliraft", batch_size=8, freeze_bn=False,
                          backbone=dict(depth=50, pretrained=None),
                          n_iters_train=10, n_iters_eval=n_iters,
                          fuse_fnet=True, fuse_cnet=True, fuse_corr=True, fuse_motion=True, fuse_hidden=False,
                          loss2d=dict(gamma=0.8, order="l2-norm"), loss3d=dict(gamma=0.8, order="l2-norm")))


def camliraft_l_cfg(n_iters=4):
    return _AttrDict(dict(name="camliraft_l", batch_size=8, n_iters_train=10, n_iters_eval=n_iters,
                          ids=dict(enabled=True), loss=dict(gamma=0.8, order="l2-norm")))


def camlipwc_cfg():
    return _AttrDict(dict(name="camlipwc", batch_size=32, freeze_bn=False,
                          pwc2d=dict(norm=dict(feature_pyramid="batch_norm", flow_estimator=None, context_network=None),
                                     max_displacement=4, lite_estimator=False, fixed=False),
                          pwc3d=dict(norm=dict(feature_pyramid="batch_norm", correlation=None, flow_estimator=None),
                                     fixed=False, k=16),
                          fusion=dict(fuse_pyramid=True, fuse_correlation=True, fuse_estimator=True),
                          loss2d=dict(level_weights=[8, 4, 2, 1, 0.5], order="l2-norm"),
                          loss3d=dict(level_weights=[8, 4, 2, 1, 0.5], order="l2-norm")))
