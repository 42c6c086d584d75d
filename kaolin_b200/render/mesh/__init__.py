"""Drop-in for ``kaolin.render.mesh``'s DIB-R path (kaolin/render/mesh/__init__.py:1-5)."""
from .rasterization import rasterize
from .dibr import dibr_soft_mask, dibr_rasterization

__all__ = ["rasterize", "dibr_soft_mask", "dibr_rasterization"]
