"""ergo_uvo_b200 -- B200-native (sm_100a) implementation of the UVO per-frame hot path.

The product is the C-ABI shared library `libuvo_b200.so` (include/uvo_c.h, sources in ergo_uvo_b200/csrc/).
This package is the thin Python binding used by tests/ and bench.py; `vo_utility` mirrors the reference's
`uvo_libraries` function API (uvo_libraries/include/uvo_libraries/VO_utility.h:96-117).
"""
from ._lib import Camera, MonoResult, Params, StereoResult, UvoError, load, SO_PATH  # noqa: F401
from .vo_utility import *  # noqa: F401,F403
