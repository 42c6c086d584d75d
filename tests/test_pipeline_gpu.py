"""sm_100a kernels of the steps either side of the rasterizer against the
reference's golden vectors (tests/golden/pipeline.npz) and, on larger seeded inputs, against the
oracle (oracle/pipeline.py; gradients through its torch restatement).  Tolerance 1e-5 relative to
the tensor's scale (fp32; the reference runs these steps as chains of library kernels whose
summation order is not specified)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import pipeline as P
from kaolin_b200 import synthetic
from kaolin_b200.metrics.render import mask_iou
from kaolin_b200.render.mesh import dibr_rasterization, prepare_vertices, texture_mapping

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pipeline.npz"))
T = lambda k, grad=False: torch.from_numpy(np.ascontiguousarray(G[k])).to(DEV).requires_grad_(grad)


def close(a, ref, tol=1e-5):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    ref = ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else ref
    scale = max(float(np.abs(ref).max()), 1e-30)
    err = float(np.abs(a - ref).max()) / scale
    assert err <= tol, err
    return err


@pytest.mark.parametrize("tag", ["T", "Rt"])
def test_prepare_vertices_vs_reference_golden(tag):
    v = T("pv_vertices", True)
    if tag == "T":
        cam = dict(camera_transform=T("pv_transform", True))
    else:
        cam = dict(camera_rot=T("pv_rot", True), camera_trans=T("pv_trans", True))
    fvc, fvi, fn = prepare_vertices(v, T("pv_faces"), T("pv_proj"), **cam)
    close(fvc, G[f"pv_{tag}_fvc"]); close(fvi, G[f"pv_{tag}_fvi"]); close(fn, G[f"pv_{tag}_fn"], 2e-5)
    ((fvc * T("pv_w1")).sum() + (fvi * T("pv_w2")).sum() + (fn * T("pv_w3")).sum()).backward()
    close(v.grad, G[f"pv_{tag}_g_vertices"], 2e-5)
    if tag == "T":
        close(cam["camera_transform"].grad, G["pv_T_g_transform"], 2e-5)
    else:
        close(cam["camera_rot"].grad, G["pv_Rt_g_rot"], 2e-5)
        close(cam["camera_trans"].grad, G["pv_Rt_g_trans"], 2e-5)


def test_prepare_vertices_icosphere_vs_oracle_and_into_the_rasterizer():
    """20 480-face icosphere, 4 views: kernel == numpy oracle; gradients == autograd through the torch
    restatement; and the outputs feed dibr_rasterization exactly like the reference's tensors."""
    verts, faces = synthetic.icosphere(5)
    B = 4
    rng = np.random.default_rng(3)
    v = np.repeat(verts[None].astype(np.float32), B, 0) * (1 + 0.05 * rng.uniform(-1, 1, (B, verts.shape[0], 1)).astype(np.float32))
    ang = rng.uniform(0, 2 * math.pi, B)
    pos = np.stack([3 * np.cos(ang), rng.uniform(-1, 1, B), 3 * np.sin(ang)], 1).astype(np.float32)
    z = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    x = np.cross(np.array([[0., 1., 0.]], np.float32), z); x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    rot = np.stack([x, y, z], 2).astype(np.float32)
    Tm = np.concatenate([rot, -(pos[:, None] @ rot)], 1).astype(np.float32)
    proj = np.array([[1 / math.tan(math.pi / 8)], [1 / math.tan(math.pi / 8)], [-1.]], np.float32)
    tv = torch.from_numpy(v).to(DEV).requires_grad_(True)
    tT = torch.from_numpy(Tm).to(DEV).requires_grad_(True)
    tf = torch.from_numpy(faces.astype(np.int64)).to(DEV)
    tp = torch.from_numpy(proj).to(DEV)
    fvc, fvi, fn = prepare_vertices(tv, tf, tp, camera_transform=tT)
    o_fvc, o_fvi, o_fn = P.prepare_vertices(v, faces, proj, camera_transform=Tm)
    # unit normals of 0.06-long edges at distance 3: the cross product cancels ~5 digits in fp32 on BOTH
    # sides (the reference's torch kernels included), so 1e-4 is the resolution of the comparison
    close(fvc, o_fvc); close(fvi, o_fvi); close(fn, o_fn, 1e-4)
    gen = torch.Generator(device=DEV); gen.manual_seed(1)
    w1, w2, w3 = (torch.rand(t.shape, device=DEV, generator=gen) for t in (fvc, fvi, fn))
    ((fvc * w1).sum() + (fvi * w2).sum() + (fn * w3).sum()).backward()
    rv, rT = tv.detach().clone().requires_grad_(True), tT.detach().clone().requires_grad_(True)
    a, b, c = P.prepare_vertices_torch(rv, tf, tp, camera_transform=rT)
    ((a * w1).sum() + (b * w2).sum() + (c * w3).sum()).backward()
    close(tv.grad, rv.grad, 2e-4); close(tT.grad, rT.grad, 2e-4)      # through 1/|cross| of tiny faces
    # straight into the rasterizer (the DIB-R loop): z, image coordinates, normal z
    H = W = 256
    ff = torch.rand((B, faces.shape[0], 3, 3), device=DEV, generator=gen)
    feat, soft, idx = dibr_rasterization(H, W, fvc[..., 2].detach(), fvi.detach(), ff, fn[..., 2].detach())
    assert 0.2 < (idx >= 0).float().mean().item() < 0.9 and soft.shape == (B, H, W)


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_texture_mapping_vs_reference_golden(mode):
    tex, uv = T("tm_tex", True), T("tm_uv", mode == "bilinear")
    o = texture_mapping(uv, tex, mode=mode)
    assert tuple(o.shape) == G[f"tm_{mode}_out"].shape
    close(o, G[f"tm_{mode}_out"])
    (o * T("tm_gw")).sum().backward()
    close(tex.grad, G[f"tm_{mode}_g_tex"])
    if mode == "bilinear":
        close(uv.grad, G["tm_bilinear_g_uv"], 2e-5)
    sp = texture_mapping(T("tm_sparse_uv"), T("tm_tex"), mode="bilinear")
    assert tuple(sp.shape) == G["tm_sparse_out"].shape
    close(sp, G["tm_sparse_out"])


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_texture_mapping_image_vs_grid_sample(mode):
    """1024^2-image sized coordinates against the library path the reference takes
    (torch.nn.functional.grid_sample through the oracle's torch restatement)."""
    gen = torch.Generator(device=DEV); gen.manual_seed(7)
    B, C, Ht, Wt, H, W = 2, 3, 128, 96, 512, 384
    tex = torch.rand((B, C, Ht, Wt), device=DEV, generator=gen).requires_grad_(True)
    uv = (torch.rand((B, H, W, 2), device=DEV, generator=gen) * 1.2 - 0.1).requires_grad_(mode == "bilinear")
    gw = torch.rand((B, H, W, C), device=DEV, generator=gen)
    o = texture_mapping(uv, tex, mode=mode)
    (o * gw).sum().backward()
    rt, ru = tex.detach().clone().requires_grad_(True), uv.detach().clone().requires_grad_(mode == "bilinear")
    ro = P.texture_mapping_torch(ru, rt, mode=mode)
    (ro * gw).sum().backward()
    close(o, ro); close(tex.grad, rt.grad, 2e-5)
    if mode == "bilinear":
        close(uv.grad, ru.grad, 2e-5)


def test_mask_iou_vs_reference_golden_and_oracle():
    l, r = T("mi_lhs", True), T("mi_rhs", True)
    loss = mask_iou(l, r)
    assert abs(loss.item() - float(G["mi_loss"])) <= 1e-6
    (loss * float(G["mi_gscale"])).backward()
    close(l.grad, G["mi_g_lhs"]); close(r.grad, G["mi_g_rhs"])
    gen = torch.Generator(device=DEV); gen.manual_seed(2)
    a = torch.rand((5, 300, 421), device=DEV, generator=gen).requires_grad_(True)
    b = (torch.rand((5, 300, 421), device=DEV, generator=gen) > 0.6).float()
    loss = mask_iou(a, b)
    assert abs(loss.item() - float(P.mask_iou(a.detach().cpu().numpy(), b.cpu().numpy()))) <= 2e-6
    loss.backward()
    ra = a.detach().clone().requires_grad_(True)
    P.mask_iou_torch(ra, b).backward()
    close(a.grad, ra.grad, 2e-5)


def test_easy_render_cuda_backend_wide_features():
    """mesh_rasterize_interpolate_cuda (easy_render/mesh.py:141-209) on duck-typed mesh / camera objects:
    D = 3 + 2 + 3 + 5 = 13 interpolated channels through the rasterizer, forward against the oracle
    and backward (wide-D path) against autograd-free finite structure: gradients of a linear loss
    equal the oracle's analytic backward."""
    import types
    import oracle
    from kaolin_b200.render.easy_render import mesh_rasterize_interpolate_cuda
    verts, faces = synthetic.icosphere(3)
    V, F = verts.shape[0], faces.shape[0]
    gen = torch.Generator(device=DEV); gen.manual_seed(11)
    tv = torch.from_numpy(verts.astype(np.float32)).to(DEV)
    tf = torch.from_numpy(faces.astype(np.int64)).to(DEV)
    attrs = {"face_normals": torch.rand((1, F, 3, 3), device=DEV, generator=gen),
             "face_uvs": torch.rand((1, F, 3, 2), device=DEV, generator=gen) * 3 - 1,
             "face_tangents": torch.rand((1, F, 3, 3), device=DEV, generator=gen),
             "face_features": torch.rand((1, F, 3, 5), device=DEV, generator=gen).requires_grad_(True)}
    mesh = types.SimpleNamespace(vertices=tv, faces=tf, has_attribute=lambda n: False,
                                 has_or_can_compute_attribute=lambda n: n in attrs, **attrs)
    f = 1 / math.tan(math.pi / 8)
    cam = types.SimpleNamespace(
        height=120, width=96, dtype=torch.float32, device=torch.device(DEV),
        extrinsics=types.SimpleNamespace(transform=lambda p: (p + torch.tensor([0., 0., -3.], device=DEV)).unsqueeze(0)),
        intrinsics=types.SimpleNamespace(transform=lambda p: torch.stack(
            [p[..., 0] * f / -p[..., 2], p[..., 1] * f / -p[..., 2], p[..., 2]], -1)))
    face_idx, im_n, im_t, im_uv, im_f = mesh_rasterize_interpolate_cuda(mesh, cam)
    assert face_idx.shape == (1, 120, 96) and im_n.shape[-1] == 3 and im_t.shape[-1] == 3
    assert im_uv.shape[-1] == 2 and im_f.shape[-1] == 5 and 0 <= im_uv.min().item() and im_uv.max().item() < 1
    vc = (tv + torch.tensor([0., 0., -3.], device=DEV)).unsqueeze(0)
    fvc = vc[:, tf]
    fvi = torch.stack([fvc[..., 0] * f / -fvc[..., 2], fvc[..., 1] * f / -fvc[..., 2]], -1)
    ff = torch.cat([attrs["face_normals"], attrs["face_uvs"], attrs["face_tangents"], attrs["face_features"].detach()], -1)
    o_feat, o_idx, o_w = oracle.rasterize(120, 96, fvc[..., 2].cpu().numpy(), fvi.cpu().numpy(), ff.cpu().numpy(),
                                          return_weights=True)
    assert np.array_equal(face_idx.cpu().numpy(), o_idx)
    close(im_n, o_feat[..., 0:3]); close(im_t, o_feat[..., 5:8]); close(im_f, o_feat[..., 8:])
    g = torch.rand(im_f.shape, device=DEV, generator=gen)
    (im_f * g).sum().backward()
    g_full = np.zeros(o_feat.shape, np.float32); g_full[..., 8:] = g.cpu().numpy()
    _, o_gff = oracle.rasterize_backward(g_full, o_idx, o_w, fvi.cpu().numpy(), ff.cpu().numpy())
    close(attrs["face_features"].grad, o_gff[..., 8:], 3e-5)
