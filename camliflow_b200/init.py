"""Name-seeded weights: every parameter / buffer is drawn from a generator keyed on its
state_dict NAME, so two differently-constructed implementations of the same architecture
(this package, the reference, the test oracle) get identical random weights for a seed.
Used for benchmarks and parity tests -- there are no checkpoints on the boxes."""
import re
import zlib

import numpy as np
import torch

# per-name gains that keep a random-weight network in a numerically tame, small-flow regime
_DAMPED = {
    # CamLiRAFT: last layer of each flow head
    r"flow_head\.conv2\.weight$": 0.05, r"flow_head\.fc\.weight$": 0.05,
    # CamLiPWC: un-normalised point-geometry features grow ~5x per PointConv level and ~100x through the
    # learnable cost volume; keep activations O(1-10) and the coarse-to-fine flow updates small
    r"conv_last\.weight$": 0.01,
    r"branch_3d_fnet\.level0_mlp\.convs\.0\.conv_fn\.weight$": 0.1,
    r"pyramid_convs\.\d\.linear\.weight$": 0.2, r"point_conv[12]\.linear\.weight$": 0.2,
    r"weight_net[12]\.convs\.2\.conv_fn\.weight$": 0.1,
}


def seeded_tensor(name, shape, seed=0):
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
    leaf = name.rsplit(".", 1)[-1]
    shape = tuple(shape)
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_var":
        return torch.rand(shape, generator=g) + 0.5
    if leaf == "running_mean":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "bias":
        return torch.randn(shape, generator=g) * 0.05
    if len(shape) == 1:
        return torch.rand(shape, generator=g) * 0.4 + 0.8
    fan_in = int(np.prod(shape[1:]))
    t = (torch.rand(shape, generator=g) * 2 - 1) * (3.0 / fan_in) ** 0.5
    for pattern, gain in _DAMPED.items():
        if re.search(pattern, name):
            t = t * gain
    return t


def seed_module_(module, seed=0):
    """In-place: fill every entry of module.state_dict() from its name."""
    sd = module.state_dict()
    module.load_state_dict({k: seeded_tensor(k, v.shape, seed) for k, v in sd.items()}, strict=True)
    return module
