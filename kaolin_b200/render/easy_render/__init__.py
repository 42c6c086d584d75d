from .mesh import mesh_rasterize_interpolate_cuda  # noqa: F401

__all__ = ["mesh_rasterize_interpolate_cuda"]
