"""Default model configurations (the values of the reference's conf/model/*.yaml), as plain
attribute dictionaries -- the models only read attributes, so a Hydra DictConfig works too."""


class AttrDict(dict):
    __getattr__ = dict.__getitem__

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = AttrDict(v) if isinstance(v, dict) else v


def camliraft_config(n_iters_eval=20, n_iters_train=10, **overrides):
    """conf/model/camliraft.yaml."""
    cfg = dict(name="camliraft", batch_size=8, freeze_bn=False, backbone=dict(depth=50, pretrained=None),
               n_iters_train=n_iters_train, n_iters_eval=n_iters_eval,
               fuse_fnet=True, fuse_cnet=True, fuse_corr=True, fuse_motion=True, fuse_hidden=False,
               loss2d=dict(gamma=0.8, order="l2-norm"), loss3d=dict(gamma=0.8, order="l2-norm"))
    cfg.update(overrides)
    return AttrDict(cfg)


def camliraft_l_config(n_iters_eval=20, n_iters_train=10, **overrides):
    """conf/model/camliraft_l.yaml."""
    cfg = dict(name="camliraft_l", batch_size=8, n_iters_train=n_iters_train, n_iters_eval=n_iters_eval,
               ids=dict(enabled=True), loss=dict(gamma=0.8, order="l2-norm"))
    cfg.update(overrides)
    return AttrDict(cfg)


def camlipwc_config(**overrides):
    """conf/model/camlipwc.yaml."""
    cfg = dict(name="camlipwc", batch_size=32, freeze_bn=False,
               pwc2d=dict(norm=dict(feature_pyramid="batch_norm", flow_estimator=None, context_network=None),
                          max_displacement=4, lite_estimator=False, fixed=False),
               pwc3d=dict(norm=dict(feature_pyramid="batch_norm", correlation=None, flow_estimator=None),
                          fixed=False, k=16),
               fusion=dict(fuse_pyramid=True, fuse_correlation=True, fuse_estimator=True),
               loss2d=dict(level_weights=[8, 4, 2, 1, 0.5], order="l2-norm"),
               loss3d=dict(level_weights=[8, 4, 2, 1, 0.5], order="l2-norm"))
    cfg.update(overrides)
    return AttrDict(cfg)
