"""The pipeline oracle (oracle/pipeline.py: prepare_vertices, texture_mapping, mask_iou) against the
golden vectors produced by the reference's own functions (tests/golden/make_pipeline_golden.py)."""
import os

import numpy as np
import torch

from oracle import pipeline as P

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pipeline.npz"))


def test_prepare_vertices_oracle_vs_reference_golden():
    for tag, kw in (("T", dict(camera_transform=G["pv_transform"])),
                    ("Rt", dict(camera_rot=G["pv_rot"], camera_trans=G["pv_trans"]))):
        fvc, fvi, fn = P.prepare_vertices(G["pv_vertices"], G["pv_faces"], G["pv_proj"], **kw)
        np.testing.assert_allclose(fvc, G[f"pv_{tag}_fvc"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(fvi, G[f"pv_{tag}_fvi"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(fn, G[f"pv_{tag}_fn"], rtol=1e-4, atol=1e-5)


def test_prepare_vertices_torch_restatement_gradients_vs_reference_golden():
    t = lambda k, grad=False: torch.from_numpy(G[k]).requires_grad_(grad)
    v, T = t("pv_vertices", True), t("pv_transform", True)
    fvc, fvi, fn = P.prepare_vertices_torch(v, t("pv_faces"), t("pv_proj"), camera_transform=T)
    ((fvc * t("pv_w1")).sum() + (fvi * t("pv_w2")).sum() + (fn * t("pv_w3")).sum()).backward()
    np.testing.assert_allclose(v.grad.numpy(), G["pv_T_g_vertices"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(T.grad.numpy(), G["pv_T_g_transform"], rtol=1e-4, atol=1e-4)


def test_texture_mapping_oracle_vs_reference_golden():
    for mode in ("nearest", "bilinear"):
        o = P.texture_mapping(G["tm_uv"], G["tm_tex"], mode)
        np.testing.assert_allclose(o, G[f"tm_{mode}_out"], rtol=1e-5, atol=1e-6)
    o = P.texture_mapping(G["tm_sparse_uv"], G["tm_tex"], "bilinear")
    assert o.shape == G["tm_sparse_out"].shape
    np.testing.assert_allclose(o, G["tm_sparse_out"], rtol=1e-5, atol=1e-6)


def test_mask_iou_oracle_vs_reference_golden():
    assert abs(float(P.mask_iou(G["mi_lhs"], G["mi_rhs"])) - float(G["mi_loss"])) <= 1e-6
    l = torch.from_numpy(G["mi_lhs"]).requires_grad_(True)
    r = torch.from_numpy(G["mi_rhs"]).requires_grad_(True)
    (P.mask_iou_torch(l, r) * float(G["mi_gscale"])).backward()
    np.testing.assert_allclose(l.grad.numpy(), G["mi_g_lhs"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r.grad.numpy(), G["mi_g_rhs"], rtol=1e-5, atol=1e-7)


def test_deftet_oracle_vs_reference_naive_golden():
    from oracle import deftet as DT
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "deftet.npz"))
    for tag in ("soup", "layers"):
        knum = int(g[f"{tag}_knum"])
        feat, idx = DT.sparse_render(g[f"{tag}_pix"], g[f"{tag}_rr"], g[f"{tag}_fvz"], g[f"{tag}_fvi"], g[f"{tag}_ff"], knum)
        assert np.array_equal(idx, g[f"{tag}_idx"]), tag
        np.testing.assert_allclose(feat, g[f"{tag}_feat"], rtol=1e-4, atol=1e-5)
        t = lambda k, grad=False: torch.from_numpy(g[f"{tag}_{k}"]).requires_grad_(grad)
        fvi, ff = t("fvi", True), t("ff", True)
        f2, i2 = DT.sparse_render_torch(t("pix"), t("rr"), t("fvz"), fvi, ff, knum)
        assert np.array_equal(i2.numpy(), g[f"{tag}_idx"]), tag
        (f2 * t("gw")).sum().backward()
        np.testing.assert_allclose(ff.grad.numpy(), g[f"{tag}_g_ff"], rtol=1e-4, atol=1e-5)
        scale = np.abs(g[f"{tag}_g_fvi"]).max()
        assert np.abs(fvi.grad.numpy() - g[f"{tag}_g_fvi"]).max() <= 1e-4 * scale
