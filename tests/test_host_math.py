"""CPU check of the exact arithmetic shared with the CUDA kernels.

kaolin_b200/csrc/dibr_math.cuh is compiled as plain C++ (tests/host_math) and
driven by brute-force loops; results must equal the oracle's *bit for bit* for
everything discrete (and for the forward floats, which use identical operation
trees).  This is what lets a CPU-only box vouch for the kernels' arithmetic: the
integer pixel rectangles, the sign-based early-out of the barycentric test, the
order-independent depth tie-break and the soft-mask distances.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from kaolin_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_math", "host_math.cpp")
LIB = os.path.join(HERE, "host_math", "_build", "libhost_math.so")
_f = ctypes.POINTER(ctypes.c_float)
_l = ctypes.POINTER(ctypes.c_int64)
_b = ctypes.POINTER(ctypes.c_uint8)
_i = ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def hm():
    hdr = os.path.join(HERE, "..", "kaolin_b200", "csrc", "dibr_math.cuh")
    if (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC),
                                                               os.path.getmtime(hdr))):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC",
                               "-shared", "-o", LIB, SRC])
    return ctypes.CDLL(LIB)


def p(a, t):
    return a.ctypes.data_as(t)


def hm_rasterize(hm, H, W, fvz, fvi, ff, valid, multiplier=1000, eps=1e-8):
    """Same packing as oracle.rasterize, but the kernel-side arithmetic."""
    B, F = fvz.shape[:2]
    D = ff.shape[-1]
    if valid is None:
        valid = np.ones((B, F), bool)
    b_idx, f_idx = np.nonzero(valid)
    first = np.zeros(B + 1, np.int64)
    np.cumsum(valid.sum(1), out=first[1:])
    xy = np.ascontiguousarray(fvi[b_idx, f_idx] * np.float32(multiplier))
    z = np.ascontiguousarray(fvz[b_idx, f_idx])
    feat = np.ascontiguousarray(ff[b_idx, f_idx])
    bbox = np.ascontiguousarray(np.concatenate([xy.min(1), xy.max(1)], 1))
    sel = np.empty((B, H, W), np.int64)
    w = np.empty((B, H, W, 3), np.float32)
    out = np.empty((B, H, W, D), np.float32)
    hm.hm_rasterize_forward(B, H, W, D, p(z, _f), p(xy, _f), p(bbox, _f), p(feat, _f),
                            p(first, _l), ctypes.c_float(multiplier), ctypes.c_float(eps),
                            p(sel, _l), p(w, _f), p(out, _f))
    face_idx = np.full(sel.shape, -1, np.int64)
    cov = sel >= 0
    face_idx[cov] = f_idx[(sel + first[:-1].reshape(-1, 1, 1))[cov]]
    return out, face_idx, w


def hm_soft(hm, fvi, face_idx, sigmainv, boxlen, knum, multiplier):
    fvi_m = np.ascontiguousarray(fvi * np.float32(multiplier))
    bb = np.ascontiguousarray(oracle._large_bboxes(fvi_m, boxlen, multiplier))
    B, F = fvi.shape[:2]
    _, H, W = face_idx.shape
    soft = np.empty((B, H, W), np.float32)
    prob = np.empty((B, H, W, knum), np.float32)
    cidx = np.empty((B, H, W, knum), np.int64)
    ctype = np.empty((B, H, W, knum), np.uint8)
    idx = np.ascontiguousarray(face_idx)
    hm.hm_soft_mask_forward(B, H, W, F, knum, p(fvi_m, _f), p(bb, _f), p(idx, _l),
                            ctypes.c_float(sigmainv), ctypes.c_float(multiplier),
                            p(soft, _f), p(prob, _f), p(cidx, _l), p(ctype, _b))
    return soft, prob, cidx, ctype, fvi_m


SCENES = [
    ("ico1_64", lambda: synthetic.icosphere_views(2, 1, seed=1), 64, 64),
    ("ico3_48x40", lambda: synthetic.icosphere_views(2, 3, seed=2), 48, 40),
    ("soup300_37x53", lambda: synthetic.triangle_soup(2, 300, seed=3, coverage=3.0), 37, 53),
    ("soup2000_96", lambda: synthetic.triangle_soup(1, 2000, seed=4), 96, 96),
]


@pytest.mark.parametrize("name,gen,H,W", SCENES, ids=[s[0] for s in SCENES])
def test_forward_bit_exact(hm, name, gen, H, W):
    fvz, fvi, fnz = gen()
    B, F = fvz.shape[:2]
    ff = synthetic.random_features(B, F, 3, seed=5)
    valid = fnz >= 0
    o_feat, o_idx, o_w = oracle.rasterize(H, W, fvz, fvi, ff, valid, return_weights=True)
    h_feat, h_idx, h_w = hm_rasterize(hm, H, W, fvz, fvi, ff, valid)
    assert (o_idx >= 0).mean() > 0.02
    assert np.array_equal(o_idx, h_idx)
    assert np.array_equal(o_w.view(np.uint32), h_w.view(np.uint32))
    assert np.array_equal(o_feat.view(np.uint32), h_feat.view(np.uint32))
    for sigmainv, boxlen, knum in ((7000, 0.02, 30), (70, 0.2, 5)):
        o = oracle.dibr_soft_mask(fvi, o_idx, sigmainv, boxlen, knum, 1000., return_lists=True)
        h = hm_soft(hm, fvi, o_idx, sigmainv, boxlen, knum, 1000.)
        assert np.array_equal(o[2], h[2])                                # close_face_idx
        assert np.array_equal(o[3], h[3])                                # dist_type
        assert np.array_equal(o[1].view(np.uint32), h[1].view(np.uint32))  # prob
        assert np.array_equal(o[0].view(np.uint32), h[0].view(np.uint32))  # soft_mask


def test_degenerate_and_ties(hm):
    """Duplicate faces (exact depth ties), zero-area faces, faces outside the image."""
    fvz, fvi, fnz = synthetic.triangle_soup(1, 40, seed=9, coverage=6.0)
    fvi = np.concatenate([fvi, fvi[:, :10], fvi[:, :5] * 0 + 0.3, fvi[:, :5] + 5.0], 1)
    fvz = np.concatenate([fvz, fvz[:, :10], fvz[:, :5], fvz[:, :5]], 1)
    B, F = fvz.shape[:2]
    ff = synthetic.random_features(B, F, 2, seed=1)
    for H, W in ((33, 29), (16, 16), (1, 1), (5, 70)):
        o_feat, o_idx, o_w = oracle.rasterize(H, W, fvz, fvi, ff, None, return_weights=True)
        h_feat, h_idx, h_w = hm_rasterize(hm, H, W, fvz, fvi, ff, None)
        assert np.array_equal(o_idx, h_idx)
        assert np.array_equal(o_w.view(np.uint32), h_w.view(np.uint32))


@pytest.mark.parametrize("multiplier", [1000.0, 100.0, 1.0, 3.7])
@pytest.mark.parametrize("shape", [(35, 31), (64, 64), (7, 129)])
def test_rect_is_exact_image_of_bbox_test(hm, multiplier, shape):
    """bbox_to_rect == the set of pixels passing the reference's float bbox test."""
    H, W = shape
    rng = np.random.default_rng(11)
    n = 400
    lo = rng.uniform(-1.3, 1.3, size=(n, 2)).astype(np.float32)
    hi = lo + rng.uniform(0, 0.5, size=(n, 2)).astype(np.float32) ** 2
    bbox = np.concatenate([lo, hi], 1) * np.float32(multiplier)
    # pixel-centre-aligned and special values
    xs = (np.float32(multiplier) / np.float32(W)) * (2 * np.arange(W) + 1 - W).astype(np.float32)
    bbox[:W, 0] = xs
    bbox[W:2 * W, 2] = xs
    bbox[-1] = [np.nan, -np.inf, np.inf, np.nan]
    bbox[-2] = [5 * multiplier, 0, 6 * multiplier, 1]
    bbox = np.ascontiguousarray(bbox.astype(np.float32))
    rects = np.empty((n, 4), np.int32)
    hm.hm_bbox_to_rect(ctypes.c_float(multiplier), W, H, p(bbox, _f), n, p(rects, _i))
    ys = (np.float32(multiplier) / np.float32(H)) * (H - 2 * np.arange(H) - 1).astype(np.float32)
    with np.errstate(invalid="ignore"):
        for i in range(n):
            inx = ~((xs < bbox[i, 0]) | (xs >= bbox[i, 2]))
            iny = ~((ys < bbox[i, 1]) | (ys >= bbox[i, 3]))
            ex = np.zeros(W, bool); ex[max(rects[i, 0], 0):max(rects[i, 1], 0)] = True
            ey = np.zeros(H, bool); ey[max(rects[i, 2], 0):max(rects[i, 3], 0)] = True
            assert np.array_equal(inx, ex), (i, bbox[i], rects[i])
            assert np.array_equal(iny, ey), (i, bbox[i], rects[i])


def test_backward_terms(hm):
    fvz, fvi, fnz = synthetic.icosphere_views(2, 2, seed=21)
    B, F = fvz.shape[:2]
    H, W = 40, 44
    ff = synthetic.random_features(B, F, 3, seed=2)
    feat, idx, w = oracle.rasterize(H, W, fvz, fvi, ff, fnz >= 0, return_weights=True)
    rng = np.random.default_rng(3)
    g = rng.uniform(size=feat.shape).astype(np.float32)
    o_gxy, o_gff = oracle.rasterize_backward(g, idx, w, fvi, ff)
    h_gxy = np.empty_like(fvi); h_gff = np.empty_like(ff)
    hm.hm_rasterize_backward(B, H, W, F, 3, p(g, _f), p(idx, _l), p(w, _f), p(fvi, _f), p(ff, _f),
                             ctypes.c_float(1e-8), p(h_gxy, _f), p(h_gff, _f))
    assert np.abs(o_gxy - h_gxy).max() <= 1e-5 * np.abs(o_gxy).max()
    assert np.abs(o_gff - h_gff).max() <= 1e-6 * max(1.0, np.abs(o_gff).max())
    soft, prob, cidx, ctype, fvi_m = hm_soft(hm, fvi, idx, 7000, 0.02, 30, 1000.)
    gs = rng.uniform(size=soft.shape).astype(np.float32)
    o_g = oracle.soft_mask_backward_op(gs, soft, idx, prob, cidx, ctype, fvi_m, 7000, 1000.)
    h_g = np.empty_like(fvi)
    hm.hm_soft_mask_backward(B, H, W, F, 30, p(gs, _f), p(soft, _f), p(idx, _l), p(prob, _f),
                             p(cidx, _l), p(ctype, _b), p(fvi_m, _f), ctypes.c_float(7000),
                             ctypes.c_float(1000.), p(h_g, _f))
    assert np.abs(o_g).max() > 0
    assert np.abs(o_g - h_g).max() <= 1e-5 * np.abs(o_g).max()
