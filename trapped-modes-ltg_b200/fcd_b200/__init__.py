"""B200-native fast-checkerboard-demodulation path (hand-written CUDA behind a C ABI).

The reference's call surface lives in the sibling package ``pyfcd`` (drop-in for
``from pyfcd.fcd import fcd``); this package is the engine underneath it."""
from .engine import (HeightMapPlan, compute_height_maps, gather_height_maps, get_plan,
                     height_from_layers, resolve_height, shard_range)

__all__ = ["HeightMapPlan", "compute_height_maps", "gather_height_maps", "get_plan",
           "height_from_layers", "resolve_height", "shard_range"]
