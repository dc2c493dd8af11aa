"""camliflow_b200 -- Blackwell-native (sm_100a) kernels for the CamLiFlow / CamLiRAFT
fused 2D-3D hot path, behind the reference's own operator surface.

`camliflow_b200.csrc` mirrors `models/csrc` of the reference (same four functions,
same pybind-level callables); `camliflow_b200.point_conv` mirrors
`models/point_conv.py`.  Everything runs through libcamli_b200.so (see
include/camli_b200.h); there is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
