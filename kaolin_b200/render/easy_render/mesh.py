"""``mesh_rasterize_interpolate_cuda`` — the CUDA backend of Kaolin's one-call renderer
(kaolin/render/easy_render/mesh.py:141-209) on the B200 rasterizer.

Only the rasterize + interpolate stage is provided: it is the part of ``render_mesh`` that runs
on the hot path (wide feature interpolation: normals 3 + uvs 2 + tangents 3 + features k);
material sampling and spherical-Gaussian shading (easy_render/mesh.py:102-139) consume its
outputs with ordinary PyTorch ops and are out of scope.

The function is duck-typed exactly on what the reference function touches, so a
``kaolin.rep.SurfaceMesh`` / ``kaolin.render.camera.Camera`` pair works unchanged, and so does
any object exposing the same attributes:

  mesh:    vertices (V,3), faces (F,3) long, face_normals (1,F,3,3) or (F,3,3),
           optional face_uvs (..,F,3,2), face_tangents (..,F,3,3), face_features (..,F,3,k),
           has_attribute(name), has_or_can_compute_attribute(name), as_transformed()
  camera:  extrinsics.transform(points), intrinsics.transform(points), height, width, dtype, device
"""
import warnings

import torch

from ..mesh import rasterize

__all__ = ["mesh_rasterize_interpolate_cuda"]


def _index_vertices_by_faces(vertices_features, faces):
    """ops/mesh/mesh.py:54-76 (a gather): (B,V,k), (F,3) -> (B,F,3,k)."""
    return vertices_features[:, faces]


def _face_attr(x):
    return x if x.dim() == 4 else x.unsqueeze(0)


def mesh_rasterize_interpolate_cuda(mesh, camera, normals_required=True, uvs_required=True,
                                    tangents_required=True, features_required=True):
    """Rasterization and interpolation of an unbatched mesh seen by a single camera
    (kaolin/render/easy_render/mesh.py:141-209).  Returns
    ``(face_idx, im_normals, im_tangents, im_uvs, im_features)`` — image-space values at the camera
    resolution for the attributes that are required and available, ``None`` for the others."""
    if mesh.has_attribute('transform'):
        mesh = mesh.as_transformed()
        warnings.warn("Mesh has a transform attribute, transforming to world space. "
                      "If you are rasterizing the mesh multiple times, "
                      "consider transforming it once before rasterizing it.", stacklevel=2)

    vertices_camera = camera.extrinsics.transform(mesh.vertices)
    vertices_image = camera.intrinsics.transform(vertices_camera)
    if vertices_camera.dim() == 2:
        vertices_camera, vertices_image = vertices_camera.unsqueeze(0), vertices_image.unsqueeze(0)

    face_vertices_camera = _index_vertices_by_faces(vertices_camera, mesh.faces)
    face_vertices_image = _index_vertices_by_faces(vertices_image, mesh.faces)[..., :2]

    in_face_features = []
    idx_normals = idx_uvs = idx_tangents = idx_features = -1
    current_idx = 0
    if normals_required:
        in_face_features.append(_face_attr(mesh.face_normals))
        idx_normals = current_idx
        current_idx += in_face_features[-1].shape[-1]
    if uvs_required and mesh.has_or_can_compute_attribute('face_uvs'):
        in_face_features.append(_face_attr(mesh.face_uvs))
        idx_uvs = current_idx
        current_idx += in_face_features[-1].shape[-1]
    if tangents_required and mesh.has_or_can_compute_attribute('face_tangents'):
        in_face_features.append(_face_attr(mesh.face_tangents))
        idx_tangents = current_idx
        current_idx += in_face_features[-1].shape[-1]
    if features_required and mesh.has_or_can_compute_attribute('face_features'):
        in_face_features.append(_face_attr(mesh.face_features))
        idx_features = current_idx
        current_idx += in_face_features[-1].shape[-1]

    if len(in_face_features) == 0:
        in_face_features = [torch.zeros((1,) + tuple(mesh.faces.shape) + (1,), dtype=camera.dtype,
                                        device=camera.device)]

    in_face_features = torch.cat(in_face_features, dim=-1).float()
    face_features, face_idx = rasterize(
        camera.height, camera.width,
        face_features=in_face_features,
        face_vertices_z=face_vertices_camera[..., -1].float().contiguous(),
        face_vertices_image=face_vertices_image.float().contiguous())

    im_normals = im_uvs = im_tangents = im_features = None
    if idx_normals >= 0:
        im_normals = face_features[..., idx_normals:idx_normals + 3]
    if idx_uvs >= 0:
        im_uvs = face_features[..., idx_uvs:idx_uvs + 2] % 1.
    if idx_tangents >= 0:
        im_tangents = face_features[..., idx_tangents:idx_tangents + 3]
    if idx_features >= 0:
        im_features = face_features[..., idx_features:]

    return face_idx, im_normals, im_tangents, im_uvs, im_features
