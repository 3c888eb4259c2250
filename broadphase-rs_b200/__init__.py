"""broadphase-rs_b200 -- the B200-native (sm_100a CUDA) implementation of broadphase-rs's hot path
(Layer::extend / sort / merge / scan) behind the C ABI of include/bp.h.

Importing this package loads libbroadphase_b200.so; it raises ImportError when the library has not
been built.  There is no CPU fallback and no alternative backend.
"""
from . import _lib
from ._lib import BpError, lib

lib()  # fail loudly, at import time, if the CUDA library is missing

from .layer import (FILTER_CATEGORY, FILTER_SPHERES, FILTER_ID_PARITY, FILTER_NONE, FILTER_XOR_MASK, PICK_AABB, PICK_SPHERE, Index32_2D,  # noqa: E402
                    Index64_2D, Index64_3D, Layer, LayerBuilder, ScanFilter, device_count, plan_dist_shard_bits,
                    plan_dist_splitters, plan_radix_passes, plan_sort_finish)
from . import scenes  # noqa: E402

__all__ = ["Layer", "LayerBuilder", "ScanFilter", "Index32_2D", "Index64_2D", "Index64_3D", "BpError",
           "FILTER_NONE", "FILTER_ID_PARITY", "FILTER_XOR_MASK", "FILTER_CATEGORY", "FILTER_SPHERES", "PICK_SPHERE", "PICK_AABB", "device_count",
           "plan_radix_passes", "plan_sort_finish", "plan_dist_splitters", "plan_dist_shard_bits", "scenes", "lib"]
