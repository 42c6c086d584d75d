"""Drop-in for ``kaolin.render.mesh``'s DIB-R path (kaolin/render/mesh/__init__.py:1-5) and the
steps either side of it (``prepare_vertices``, ``texture_mapping``: kaolin/render/mesh/utils.py)."""
from .rasterization import rasterize
from .dibr import dibr_soft_mask, dibr_rasterization
from .utils import prepare_vertices, texture_mapping
from .deftet import deftet_sparse_render
from .graphed import make_graphed_dibr_rasterization

__all__ = ["rasterize", "dibr_soft_mask", "dibr_rasterization", "prepare_vertices", "texture_mapping",
           "deftet_sparse_render", "make_graphed_dibr_rasterization"]
