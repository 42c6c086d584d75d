"""kaolin_b200 — B200-native (sm_100a) drop-in for Kaolin's DIB-R hot path.

Provided: ``kaolin.render.mesh.{rasterize, dibr_soft_mask, dibr_rasterization}``
(forward + backward), the four ``kaolin._C.render.mesh.*`` operators, and the steps
either side of the rasterizer in the DIB-R loop (``render.mesh.prepare_vertices``,
``render.mesh.texture_mapping``, ``metrics.render.mask_iou``); see DESIGN.md for scope.  There is no CPU fallback: every entry point
raises if the CUDA library (kaolin_b200/csrc, built in-tree) is missing or the
tensors are not on a CUDA device.
"""
__version__ = "0.1.0"

from . import render   # noqa: E402,F401
from . import metrics  # noqa: E402,F401
from . import _C       # noqa: E402,F401
