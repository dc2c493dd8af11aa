"""Name-seeded weights: every parameter / buffer is drawn from a generator keyed on its
state_dict NAME, so two differently-constructed implementations of the same architecture
(this package, the reference, the test oracle) get identical random weights for a seed.
Used for benchmarks and parity tests -- there are no checkpoints on the boxes."""
import zlib

import numpy as np
import torch

_DAMPED = ("flow_head.conv2.weight", "flow_head.fc.weight")   # keep the recurrence in the small-flow regime


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
    return t * 0.05 if name.endswith(_DAMPED) else t


def seed_module_(module, seed=0):
    """In-place: fill every entry of module.state_dict() from its name."""
    sd = module.state_dict()
    module.load_state_dict({k: seeded_tensor(k, v.shape, seed) for k, v in sd.items()}, strict=True)
    return module
