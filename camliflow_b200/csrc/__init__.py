"""Drop-in for the reference's `models/csrc` package (models/csrc/__init__.py:1):
the same four public functions, backed by libcamli_b200.so."""
from .wrapper import correlation2d, furthest_point_sampling, squared_distance, k_nearest_neighbor  # noqa: F401
