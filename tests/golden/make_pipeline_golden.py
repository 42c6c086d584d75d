"""Regenerates tests/golden/pipeline.npz from the reference (run in the build container only).

    python tests/golden/make_pipeline_golden.py

Outputs (and autograd gradients) of the reference's OWN functions, imported in place from
/root/reference, on small seeded inputs:
  kaolin.render.mesh.utils.prepare_vertices  (both camera conventions)      utils.py:129-175
  kaolin.render.mesh.utils.texture_mapping   ('nearest' and 'bilinear')     utils.py:22-79
  kaolin.metrics.render.mask_iou                                            render.py:18-41
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_import  # noqa: E402

utils = ref_import.module("kaolin.render.mesh.utils")
camera = ref_import.module("kaolin.render.camera.legacy")
metrics = ref_import.module("kaolin.metrics.render")

out = {}
g = torch.Generator().manual_seed(20260923)
B, V, F = 3, 40, 70
vertices = (torch.rand((B, V, 3), generator=g) - 0.5).requires_grad_(True)
faces = torch.stack([torch.randperm(V, generator=g)[:3] for _ in range(F)]).long()
pos = torch.tensor([[0., 0., 3.], [2., 1., 2.], [-2.5, 0.5, -1.]])
look = torch.zeros((B, 3)); up = torch.tensor([[0., 1., 0.]]).repeat(B, 1)
proj = camera.generate_perspective_projection(math.pi / 4, 1.0)
T = camera.generate_transformation_matrix(pos, look, up).requires_grad_(True)
R, t = camera.generate_rotate_translate_matrices(pos, look, up)
R = R.clone().requires_grad_(True); t = t.clone().requires_grad_(True)
w1 = torch.rand((B, F, 3, 3), generator=g); w2 = torch.rand((B, F, 3, 2), generator=g); w3 = torch.rand((B, F, 3), generator=g)
for tag, kw in (("T", dict(camera_transform=T)), ("Rt", dict(camera_rot=R, camera_trans=t))):
    for p in (vertices, T, R, t):
        p.grad = None
    fvc, fvi, fn = utils.prepare_vertices(vertices, faces, proj, **kw)
    ((fvc * w1).sum() + (fvi * w2).sum() + (fn * w3).sum()).backward()
    out[f"pv_{tag}_fvc"], out[f"pv_{tag}_fvi"], out[f"pv_{tag}_fn"] = fvc.detach().numpy(), fvi.detach().numpy(), fn.detach().numpy()
    out[f"pv_{tag}_g_vertices"] = vertices.grad.numpy().copy()
    if tag == "T":
        out["pv_T_g_transform"] = T.grad.numpy().copy()
    else:
        out["pv_Rt_g_rot"], out["pv_Rt_g_trans"] = R.grad.numpy().copy(), t.grad.numpy().copy()
out.update(pv_vertices=vertices.detach().numpy(), pv_faces=faces.numpy(), pv_proj=proj.numpy(),
           pv_transform=T.detach().numpy(), pv_rot=R.detach().numpy(), pv_trans=t.detach().numpy(),
           pv_w1=w1.numpy(), pv_w2=w2.numpy(), pv_w3=w3.numpy())

B, C, Ht, Wt, H, W = 2, 3, 9, 7, 11, 13
tex = torch.rand((B, C, Ht, Wt), generator=g).requires_grad_(True)
uv = (torch.rand((B, H, W, 2), generator=g) * 1.3 - 0.15)      # some outside [0,1]
uv[0, 0, 0] = torch.tensor([0., 1.]); uv[0, 0, 1] = torch.tensor([1., 0.]); uv[0, 0, 2] = torch.tensor([0.5, 0.5])
uv.requires_grad_(True)
gw = torch.rand((B, H, W, C), generator=g)
for mode in ("nearest", "bilinear"):
    tex.grad = None; uv.grad = None
    o = utils.texture_mapping(uv, tex, mode=mode)
    (o * gw).sum().backward()
    out[f"tm_{mode}_out"] = o.detach().numpy()
    out[f"tm_{mode}_g_tex"] = tex.grad.numpy().copy()
    out[f"tm_{mode}_g_uv"] = (uv.grad if uv.grad is not None else torch.zeros_like(uv)).numpy().copy()
sp = torch.rand((B, 17, 2), generator=g)
out["tm_sparse_uv"] = sp.numpy(); out["tm_sparse_out"] = utils.texture_mapping(sp, tex.detach(), mode="bilinear").numpy()
out.update(tm_tex=tex.detach().numpy(), tm_uv=uv.detach().numpy(), tm_gw=gw.numpy())

l = torch.rand((3, 10, 12), generator=g).requires_grad_(True)
r = (torch.rand((3, 10, 12), generator=g) > 0.5).float().requires_grad_(True)
loss = metrics.mask_iou(l, r)
(loss * 1.7).backward()
out.update(mi_lhs=l.detach().numpy(), mi_rhs=r.detach().numpy(), mi_loss=np.float32(loss.item()),
           mi_g_lhs=l.grad.numpy(), mi_g_rhs=r.grad.numpy(), mi_gscale=np.float32(1.7))

np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **out)
print("wrote", os.path.join(HERE, "pipeline.npz"), {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
