"""Pins the CPU oracle (oracle/dibr_oracle.c) to the reference's golden vectors.

Mirrors the assertions (and tolerances) of the reference's own tests:
  tests/python/kaolin/render/mesh/test_dibr.py:109-191   (simple scene)
  tests/python/kaolin/render/mesh/test_dibr.py:309-394   (sphere scene)
  tests/python/kaolin/render/mesh/test_rasterization.py:137-289 (naive oracle)
against tests/golden/*.npz (made by tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest

import oracle

H, W = 35, 31


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _mask_iou_grad(soft, face_idx):
    """d mask_iou(soft, shifted)/d soft — kaolin/metrics/render.py:18-41 with the
    shifted target of test_dibr.py:182-186, differentiated by hand (float64)."""
    B = soft.shape[0]
    mask = face_idx != -1
    shifted = np.zeros_like(mask)
    shifted[..., :-5] = mask[..., 5:]
    s = soft.astype(np.float64)
    r = shifted.astype(np.float64)
    up = (s * r).reshape(B, -1).sum(1)
    down = (s + r - s * r).reshape(B, -1).sum(1) + 1e-10
    d_up = r
    d_down = 1.0 - r
    g = -(d_up / down[:, None, None] - up[:, None, None] * d_down / down[:, None, None] ** 2) / B
    return g.astype(np.float32)


@pytest.mark.parametrize("sigmainv", [7000, 70])
@pytest.mark.parametrize("boxlen", [0.02, 0.2])
@pytest.mark.parametrize("multiplier", [1000, 100, 1])
@pytest.mark.parametrize("knum", [30, 20])
def test_simple_scene(golden_dir, sigmainv, boxlen, multiplier, knum):
    g = _load(golden_dir, "dibr_simple.npz")
    key = f"s{sigmainv}_b{boxlen}_"
    fvi, fvz = g["fvi"], g["fvz"]
    ff = np.zeros(fvz.shape + (1,), np.float32)
    _, face_idx = oracle.rasterize(H, W, fvz, fvi, ff)
    # rasterize KAT: simple/new_face_idx_35_31.pt
    assert np.array_equal(face_idx, g["face_idx"].astype(np.int64))
    soft, prob, cidx, ctype = oracle.dibr_soft_mask(
        fvi, face_idx, sigmainv, boxlen, knum, multiplier, return_lists=True)
    np.testing.assert_allclose(soft, g[key + "soft_mask"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(cidx, g[key + "close_face_idx"][..., :knum].astype(np.int64))
    np.testing.assert_allclose(prob, g[key + "close_face_prob"][..., :knum], rtol=1e-5, atol=1e-5)
    assert np.array_equal(ctype, g[key + "close_face_dist_type"][..., :knum])
    # backward: test_dibr.py:167-191
    gsoft = _mask_iou_grad(soft, face_idx)
    gxy = oracle.dibr_soft_mask_backward(gsoft, fvi, face_idx, sigmainv, boxlen, knum, multiplier)
    np.testing.assert_allclose(gxy, g[key + "grad_fvi"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("sigmainv", [7000, 70])
@pytest.mark.parametrize("boxlen", [0.02, 0.01])
@pytest.mark.parametrize("multiplier", [1000, 100])
@pytest.mark.parametrize("knum", [30, 40])
@pytest.mark.parametrize("batch_size", [1, 3])
@pytest.mark.parametrize("flip", [False, True])
def test_sphere_scene(golden_dir, sigmainv, boxlen, multiplier, knum, batch_size, flip):
    g = _load(golden_dir, "dibr_sphere.npz")
    key = f"s{sigmainv}_b{boxlen}_"
    fvi, fvz = g["fvi"][:batch_size], g["fvz"][:batch_size]
    if flip:  # test_dibr.py:204-209 flips the vertex order of every face
        fvi = np.ascontiguousarray(fvi[:, :, ::-1])
        fvz = np.ascontiguousarray(fvz[:, :, ::-1])
    ff = np.zeros(fvz.shape + (1,), np.float32)
    _, face_idx = oracle.rasterize(H, W, fvz, fvi, ff)
    soft, prob, cidx, ctype = oracle.dibr_soft_mask(
        fvi, face_idx, sigmainv, boxlen, knum, multiplier, return_lists=True)
    np.testing.assert_allclose(soft, g[key + "soft_mask"][:batch_size], rtol=1e-5, atol=1e-5)
    assert np.array_equal(cidx, g[key + "close_face_idx"][:batch_size, ..., :knum].astype(np.int64))
    np.testing.assert_allclose(prob, g[key + "close_face_prob"][:batch_size, ..., :knum],
                               rtol=1e-5, atol=1e-5)
    if not flip:  # (with flipped vertex order the edge/vertex ids are permuted)
        mism = ctype != g[key + "close_face_dist_type"][:batch_size, ..., :knum]
        assert mism.sum() / mism.size <= 0.01          # test_dibr.py:340-341
    gsoft = _mask_iou_grad(soft, face_idx)
    gxy = oracle.dibr_soft_mask_backward(gsoft, fvi, face_idx, sigmainv, boxlen, knum, multiplier)
    ref = g[key + "grad_fvi"][:batch_size]
    if flip:
        ref = ref[:, :, ::-1]
    if batch_size == 3:   # the stored gradient is of the batch-3 mean loss
        np.testing.assert_allclose(gxy, ref, rtol=1e-1, atol=1e-1)  # test_dibr.py:392-394
        # much tighter in aggregate than the reference's own tolerance:
        assert np.abs(gxy - ref).max() <= 2e-3 * max(1e-6, np.abs(ref).max()) + 1e-6


@pytest.mark.parametrize("tag", ["all", "valid"])
@pytest.mark.parametrize("batch_size", [1, 3])
def test_rasterize_vs_naive(golden_dir, tag, batch_size):
    """test_rasterization.py:137-289 with the naive-oracle outputs stored as fixtures."""
    g = _load(golden_dir, "rasterize_model.npz")
    b = batch_size
    fvi, fvz, uvs = g["fvi"][:b], g["fvz"][:b], g["face_uvs"][:b]
    valid = g["valid_faces"][:b] if tag == "valid" else None
    ones = np.ones_like(uvs[..., :1])
    (uv_map, mask), face_idx, w = oracle.rasterize(32, 32, fvz, fvi, [uvs, ones], valid,
                                                   return_weights=True)
    assert np.array_equal(face_idx, g[tag + "_face_idx"][:b].astype(np.int64))
    feats = np.concatenate([uv_map, mask], -1)
    np.testing.assert_allclose(feats, g[tag + "_features"][:b], rtol=1e-5, atol=1e-5)
    gxy, gff = oracle.rasterize_backward(g["grad_out"][:b], face_idx, w, fvi,
                                         np.concatenate([uvs, ones], -1))
    np.testing.assert_allclose(gxy, g[tag + "_grad_fvi"][:b], rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(gff[..., :2], g[tag + "_grad_uvs"][:b], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(gff[..., 2:], g[tag + "_grad_ones"][:b], rtol=1e-3, atol=1e-3)
    # aggregate error is far below the reference's own tolerance
    assert np.abs(gxy - g[tag + "_grad_fvi"][:b]).max() <= 1e-4 * np.abs(g[tag + "_grad_fvi"][:b]).max()
