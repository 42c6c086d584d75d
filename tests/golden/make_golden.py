"""Regenerates tests/golden/*.npz from the reference (run in the build container only).

    python tests/golden/make_golden.py

Sources (nothing here is reference *source code*; these are the reference's test
vectors and outputs of its own Python code, which cannot travel to the GPU box):

* tests/samples/dibr/simple/*.pt, tests/samples/dibr/sphere/*.pt — the golden
  tensors "From Kaolin V0.10.0" that tests/python/kaolin/render/mesh/test_dibr.py
  :83-107,281-307 compare the CUDA kernels against (stored here with compact
  dtypes; face ids converted to 0-based, -1 = empty, as the tests do `.long()-1`).
* the input scenes of those tests: the literal tensors of test_dibr.py:43-62 and
  tests/samples/model.obj seen through the three fixture cameras
  (test_dibr.py:198-261 == test_rasterization.py:36-104), computed with the
  reference's own camera / indexing functions imported in place.
* `_naive_deftet_sparse_render` (kaolin/render/mesh/deftet.py:101-267), the pure
  PyTorch oracle the reference's rasterize tests use
  (test_rasterization.py:137-233), run here on the same fixtures, forward and
  autograd backward.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_import  # noqa: E402

REF = ref_import.REF
SAMPLES = os.path.join(REF, "tests", "samples")


def load_pt(*p):
    return torch.load(os.path.join(SAMPLES, *p), map_location="cpu")


def read_obj(path):
    """v / vt / f lines in file order, 0-based (as kaolin/io/obj.py:228-291)."""
    v, vt, f, ft = [], [], [], []
    for line in open(path):
        d = line.split()
        if not d:
            continue
        if d[0] == "v":
            v.append([float(x) for x in d[1:4]])
        elif d[0] == "vt":
            vt.append([float(x) for x in d[1:3]])
        elif d[0] == "f":
            parts = [p.split("/") for p in d[1:]]
            f.append([int(p[0]) - 1 for p in parts])
            ft.append([int(p[1]) - 1 for p in parts])
    uvs = torch.tensor(vt, dtype=torch.float32)
    uvs[..., 1] = 1 - uvs[..., 1]          # obj.py:293
    return (torch.tensor(v, dtype=torch.float32), torch.tensor(f, dtype=torch.long),
            uvs, torch.tensor(ft, dtype=torch.long))


def simple_scene():
    fvi = torch.tensor(
        [[[[-0.7, 0.], [0., -0.7], [0., 0.7]],
          [[-0.7, 0.], [0., 0.7], [0., -0.7]],
          [[0., -0.7], [0., 0.7], [0.7, 0.]]],
         [[[-0.7, -0.7], [0.7, -0.7], [-0.7, 0.7]],
          [[-0.7, -0.7], [0.7, -0.7], [-0.7, 0.7]],
          [[-0.7, -0.7], [0.7, -0.7], [-0.7, 0.7]]]], dtype=torch.float32)
    fvz = torch.tensor(
        [[[-2., -1., -1.], [-2.5, -3., -3.], [-2., -2., -2.]],
         [[-2., -1., -3.], [-2., -2., -2.], [-2., -3., -1.]]], dtype=torch.float32)
    return fvi, fvz


def make_simple():
    fvi, fvz = simple_scene()
    out = {"fvi": fvi.numpy(), "fvz": fvz.numpy(),
           "face_idx": load_pt("dibr", "simple", "new_face_idx_35_31.pt").numpy().astype(np.int8)}
    for sig in (7000, 70):
        for box in (0.02, 0.2):
            tag = f"35_31_{sig}_{box}"
            key = f"s{sig}_b{box}"
            out[key + "_soft_mask"] = load_pt("dibr", "simple", f"soft_mask_{tag}.pt").numpy().astype(np.float32)
            out[key + "_close_face_idx"] = (load_pt("dibr", "simple", f"close_face_idx_{tag}.pt").long() - 1).numpy().astype(np.int8)
            out[key + "_close_face_prob"] = load_pt("dibr", "simple", f"close_face_dist_{tag}.pt").numpy().astype(np.float32)
            out[key + "_close_face_dist_type"] = load_pt("dibr", "simple", f"close_face_dist_type_{tag}.pt").numpy().astype(np.uint8)
            out[key + "_grad_fvi"] = load_pt("dibr", "simple", f"grad_face_vertices_image_{tag}.pt").numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "dibr_simple.npz"), **out)
    print("dibr_simple.npz", {k: v.shape for k, v in out.items() if "7000_b0.02" in k or "_" not in k})


def model_scene(dtype=torch.float32):
    """Fixtures of test_dibr.py:198-261 / test_rasterization.py:36-104 (batch 3, no flip)."""
    cam = ref_import.module("kaolin.render.camera.legacy")
    meshops = ref_import.module("kaolin.ops.mesh.mesh")
    vertices, faces, uvs, face_uvs_idx = read_obj(os.path.join(SAMPLES, "model.obj"))
    camera_pos = torch.tensor([[0.5, 0.5, 3.], [2., 2., -2.], [3., 0.5, 0.5]], dtype=dtype)
    look_at = torch.full((3, 3), 0.5, dtype=dtype)
    camera_up = torch.tensor([[0., 1., 0.]], dtype=dtype).repeat(3, 1)
    camera_proj = cam.generate_perspective_projection(fovyangle=math.pi / 4., dtype=dtype)
    v = vertices.to(dtype).unsqueeze(0)
    vmin = v.min(dim=1, keepdims=True)[0]
    vmax = v.max(dim=1, keepdims=True)[0]
    v = (v - vmin) / (vmax - vmin)
    rot, trans = cam.generate_rotate_translate_matrices(camera_pos, look_at, camera_up)
    v_cam = cam.rotate_translate_points(v, rot, trans)
    v_img = cam.perspective_camera(v_cam, camera_proj)
    fvz = meshops.index_vertices_by_faces(v_cam[:, :, -1:], faces).squeeze(-1)
    fvi = meshops.index_vertices_by_faces(v_img, faces)
    face_uvs = meshops.index_vertices_by_faces(uvs.unsqueeze(0).to(dtype), face_uvs_idx).repeat(3, 1, 1, 1)
    # test_rasterization.py:96-101
    min_z = fvz.reshape(3, -1).min(dim=1, keepdims=True)[0]
    max_z = fvz.reshape(3, -1).max(dim=1, keepdims=True)[0]
    middle_z = (min_z + max_z) / 2.
    valid_faces = torch.all(fvz < middle_z.unsqueeze(-1), dim=-1)
    return fvz, fvi, face_uvs, valid_faces, v_cam


def make_sphere():
    fvz, fvi, face_uvs, valid_faces, _ = model_scene()
    out = {"fvi": fvi.numpy(), "fvz": fvz.numpy()}
    for sig in (7000, 70):
        for box in (0.02, 0.01):
            tag = f"35_31_{sig}_{box}"
            key = f"s{sig}_b{box}"
            out[key + "_soft_mask"] = load_pt("dibr", "sphere", f"soft_mask_{tag}.pt").numpy().astype(np.float32)
            out[key + "_close_face_idx"] = (load_pt("dibr", "sphere", f"close_face_idx_{tag}.pt").long() - 1).numpy().astype(np.int16)
            out[key + "_close_face_prob"] = load_pt("dibr", "sphere", f"close_face_dist_{tag}.pt").numpy().astype(np.float32)
            out[key + "_close_face_dist_type"] = load_pt("dibr", "sphere", f"close_face_dist_type_{tag}.pt").numpy().astype(np.uint8)
            out[key + "_grad_fvi"] = load_pt("dibr", "sphere", f"grad_face_vertices_image_{tag}.pt").numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "dibr_sphere.npz"), **out)
    print("dibr_sphere.npz written")


def make_rasterize():
    """Outputs of the reference's naive oracle on the rasterize test fixtures (32x32)."""
    deftet = ref_import.module("kaolin.render.mesh.deftet")
    H = W = 32
    fvz, fvi, face_uvs, valid_faces, v_cam = model_scene()
    B = 3
    x = (2 * torch.arange(W, dtype=torch.float32) + 1 - W) / W
    y = (H - 2 * torch.arange(H, dtype=torch.float32) - 1.) / H
    pixel_coords = torch.stack([x.reshape(1, 1, -1).repeat(B, H, 1),
                                y.reshape(1, -1, 1).repeat(B, 1, W)], dim=-1).reshape(B, -1, 2)
    min_z = v_cam[:, :, -1].min(dim=1)[0]
    max_z = v_cam[:, :, -1].max(dim=1)[0]
    render_ranges = torch.stack([min_z - 1e-2, max_z + 1e-2], dim=-1).unsqueeze(1).repeat(1, H * W, 1)
    torch.manual_seed(20260922)
    grad_out = torch.rand(B, H, W, 3)
    out = {"fvi": fvi.numpy(), "fvz": fvz.numpy(), "face_uvs": face_uvs.numpy(),
           "valid_faces": valid_faces.numpy(), "grad_out": grad_out.numpy()}
    for tag, kwargs in (("all", {}), ("valid", {"valid_faces": valid_faces})):
        a_fvi = fvi.clone().requires_grad_(True)
        a_uv = face_uvs.clone().requires_grad_(True)
        ones = torch.ones_like(face_uvs[..., :1]).requires_grad_(True)
        (g_uv, g_mask), g_idx = deftet._naive_deftet_sparse_render(
            pixel_coords, render_ranges, fvz, a_fvi, [a_uv, ones], 1, **kwargs)
        feats = torch.cat([g_uv.reshape(B, H, W, 2), g_mask.reshape(B, H, W, 1)], -1)
        feats.backward(grad_out)
        out[tag + "_face_idx"] = g_idx.reshape(B, H, W).numpy().astype(np.int16)
        out[tag + "_features"] = feats.detach().numpy()
        out[tag + "_grad_fvi"] = a_fvi.grad.numpy()
        out[tag + "_grad_uvs"] = a_uv.grad.numpy()
        out[tag + "_grad_ones"] = ones.grad.numpy()
        print(tag, "covered", float((g_idx >= 0).float().mean()))
    np.savez_compressed(os.path.join(HERE, "rasterize_model.npz"), **out)
    print("rasterize_model.npz written")


if __name__ == "__main__":
    make_simple()
    make_sphere()
    make_rasterize()
