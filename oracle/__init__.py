"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

numpy/ctypes front-end of the CPU oracle (oracle/dibr_oracle.c): a restatement
of Kaolin's DIB-R hot path.  The host-side preparation the reference does in
Python is restated here with numpy float32 arithmetic:

* ``rasterize``           <- kaolin/render/mesh/rasterization.py:273-352 (RasterizeCuda.forward)
* ``rasterize_backward``  <- kaolin/render/mesh/rasterization.py:355-371
* ``dibr_soft_mask``      <- kaolin/render/mesh/dibr.py:29-55
* ``dibr_soft_mask_backward`` <- kaolin/render/mesh/dibr.py:58-73
* ``dibr_rasterization``  <- kaolin/render/mesh/dibr.py:190-209

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  ``kaolin_b200`` never does.

Parity status: PINNED — see oracle/README.md (golden fixtures of the reference's
own tests, tests/golden/*.npz, are reproduced by tests/test_oracle_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "dibr_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libdibr_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off (no compiler-added FMA fusion)."""
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC)):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-fopenmp",
           "-ffp-contract=off", "-fno-fast-math", "-o", _LIB, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def set_threads(n):
    lib().oracle_set_threads(ctypes.c_int(int(n)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


# --------------------------------------------------------------------------
# raw operator level (same argument meaning as kaolin._C.render.mesh.*)
# --------------------------------------------------------------------------
def packed_rasterize_forward(height, width, face_vertices_z, face_vertices_image,
                             face_bboxes, face_features, first_idx_face_per_mesh,
                             multiplier, eps):
    """kaolin/csrc/render/mesh/rasterization.cpp:49-104."""
    z = _f32(face_vertices_z)
    xy = _f32(face_vertices_image)
    bb = _f32(face_bboxes)
    ff = _f32(face_features)
    first = np.ascontiguousarray(first_idx_face_per_mesh, dtype=np.int64)
    B = first.shape[0] - 1
    D = ff.shape[-1]
    idx = np.empty((B, height, width), np.int64)
    w = np.empty((B, height, width, 3), np.float32)
    out = np.empty((B, height, width, D), np.float32)
    lib().oracle_rasterize_forward(
        ctypes.c_int(B), ctypes.c_int(height), ctypes.c_int(width), ctypes.c_int(D),
        _p(z, _f32p), _p(xy, _f32p), _p(bb, _f32p), _p(ff, _f32p), _p(first, _i64p),
        ctypes.c_float(multiplier), ctypes.c_float(eps),
        _p(idx, _i64p), _p(w, _f32p), _p(out, _f32p))
    return out, idx, w


def rasterize_backward_op(grad_interpolated_features, selected_face_idx, output_weights,
                          face_vertices_image, face_features, eps):
    """kaolin/csrc/render/mesh/rasterization.cpp:106-168."""
    g = _f32(grad_interpolated_features)
    idx = np.ascontiguousarray(selected_face_idx, dtype=np.int64)
    w = _f32(output_weights)
    xy = _f32(face_vertices_image)
    ff = _f32(face_features)
    B, H, W, D = g.shape
    F = xy.shape[1]
    gxy = np.empty_like(xy)
    gff = np.empty_like(ff)
    lib().oracle_rasterize_backward(
        ctypes.c_int(B), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(F), ctypes.c_int(D),
        _p(g, _f32p), _p(idx, _i64p), _p(w, _f32p), _p(xy, _f32p), _p(ff, _f32p),
        ctypes.c_float(eps), _p(gxy, _f32p), _p(gff, _f32p))
    return gxy, gff


def soft_mask_forward_op(face_vertices_image_m, face_large_bboxes, selected_face_idx,
                         sigmainv, knum, multiplier, with_lists=True):
    """kaolin/csrc/render/mesh/dibr_soft_mask.cpp:48-108."""
    xy = _f32(face_vertices_image_m)
    bb = _f32(face_large_bboxes)
    idx = np.ascontiguousarray(selected_face_idx, dtype=np.int64)
    B, F = xy.shape[:2]
    _, H, W = idx.shape
    soft = np.empty((B, H, W), np.float32)
    prob = cidx = ctype = None
    if with_lists:
        prob = np.empty((B, H, W, knum), np.float32)
        cidx = np.empty((B, H, W, knum), np.int64)
        ctype = np.empty((B, H, W, knum), np.uint8)
    lib().oracle_soft_mask_forward(
        ctypes.c_int(B), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(F), ctypes.c_int(knum),
        _p(xy, _f32p), _p(bb, _f32p), _p(idx, _i64p),
        ctypes.c_float(sigmainv), ctypes.c_float(multiplier),
        _p(soft, _f32p), _p(prob, _f32p), _p(cidx, _i64p), _p(ctype, _u8p))
    return soft, prob, cidx, ctype


def soft_mask_backward_op(grad_soft_mask, soft_mask, selected_face_idx, close_face_prob,
                          close_face_idx, close_face_dist_type, face_vertices_image_m,
                          sigmainv, multiplier):
    """kaolin/csrc/render/mesh/dibr_soft_mask.cpp:110-183."""
    g = _f32(grad_soft_mask)
    soft = _f32(soft_mask)
    idx = np.ascontiguousarray(selected_face_idx, dtype=np.int64)
    prob = _f32(close_face_prob)
    cidx = np.ascontiguousarray(close_face_idx, dtype=np.int64)
    ctype = np.ascontiguousarray(close_face_dist_type, dtype=np.uint8)
    xy = _f32(face_vertices_image_m)
    B, F = xy.shape[:2]
    _, H, W = idx.shape
    K = cidx.shape[-1]
    gxy = np.empty_like(xy)
    lib().oracle_soft_mask_backward(
        ctypes.c_int(B), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(F), ctypes.c_int(K),
        _p(g, _f32p), _p(soft, _f32p), _p(idx, _i64p), _p(prob, _f32p), _p(cidx, _i64p),
        _p(ctype, _u8p), _p(xy, _f32p), ctypes.c_float(sigmainv), ctypes.c_float(multiplier),
        _p(gxy, _f32p))
    return gxy


# --------------------------------------------------------------------------
# public-API level (numpy restatement of the reference's Python wrappers)
# --------------------------------------------------------------------------
def rasterize(height, width, face_vertices_z, face_vertices_image, face_features,
              valid_faces=None, multiplier=None, eps=None, return_weights=False):
    """rasterization.py:273-352 + :455-467 (defaults, list concat).

    Returns (interpolated_features, face_idx[, output_weights]).
    """
    if multiplier is None:
        multiplier = 1000
    if eps is None:
        eps = 1e-8
    is_list = isinstance(face_features, (list, tuple))
    ff = np.concatenate([_f32(f) for f in face_features], -1) if is_list else _f32(face_features)
    fvz = _f32(face_vertices_z)
    fvi = _f32(face_vertices_image)
    B, F = fvz.shape[:2]
    D = ff.shape[-1]
    if valid_faces is None:
        b_idx = np.repeat(np.arange(B), F)
        f_idx = np.tile(np.arange(F), B)
        nfpm = np.full((B,), F, np.int64)
    else:
        vf = np.asarray(valid_faces, dtype=bool)
        b_idx, f_idx = np.nonzero(vf)           # row-major == torch.where order
        nfpm = vf.reshape(B, -1).sum(1).astype(np.int64)
    v_xy = fvi[b_idx, f_idx].reshape(-1, 3, 2)
    v_z = fvz[b_idx, f_idx].reshape(-1, 3)
    v_ff = ff[b_idx, f_idx].reshape(-1, 3, D)
    first = np.zeros(B + 1, np.int64)
    np.cumsum(nfpm, out=first[1:])
    v_xy = v_xy * np.float32(multiplier)                       # :320
    pmin = v_xy.min(axis=1) if v_xy.shape[0] else np.zeros((0, 2), np.float32)
    pmax = v_xy.max(axis=1) if v_xy.shape[0] else np.zeros((0, 2), np.float32)
    bboxes = np.concatenate([pmin, pmax], axis=1)              # :325-327
    out, sel, w = packed_rasterize_forward(height, width, v_z, v_xy, bboxes, v_ff,
                                           first, multiplier, eps)
    # :340-346 remap packed -> original face ids
    face_idx = np.full(sel.shape, -1, np.int64)
    covered = sel >= 0
    packed = sel + first[:-1].reshape(-1, 1, 1)
    face_idx[covered] = f_idx[packed[covered]]
    if is_list:
        outs, cur = [], 0
        for f in face_features:
            d = np.asarray(f).shape[-1]
            outs.append(out[..., cur:cur + d])
            cur += d
        out = tuple(outs)
    if return_weights:
        return out, face_idx, w
    return out, face_idx


def rasterize_backward(grad_interpolated_features, face_idx, output_weights,
                       face_vertices_image, face_features, eps=None):
    """rasterization.py:355-371 (face_vertices_image UNSCALED)."""
    if eps is None:
        eps = 1e-8
    return rasterize_backward_op(grad_interpolated_features, face_idx, output_weights,
                                 face_vertices_image, face_features, eps)


def _large_bboxes(fvi_m, boxlen, multiplier):
    """dibr.py:33-39."""
    pmin = fvi_m.min(axis=-2)
    pmax = fvi_m.max(axis=-2)
    margin = np.float32(boxlen * multiplier)
    return np.concatenate([pmin - margin, pmax + margin], axis=-1)


def dibr_soft_mask(face_vertices_image, selected_face_idx, sigmainv=7000, boxlen=0.02,
                   knum=30, multiplier=1000., return_lists=False):
    """dibr.py:29-55."""
    fvi_m = _f32(face_vertices_image) * np.float32(multiplier)
    bb = _large_bboxes(fvi_m, boxlen, multiplier)
    soft, prob, cidx, ctype = soft_mask_forward_op(
        fvi_m, bb, selected_face_idx, sigmainv, knum, multiplier, with_lists=True)
    if return_lists:
        return soft, prob, cidx, ctype
    return soft


def dibr_soft_mask_backward(grad_soft_mask, face_vertices_image, selected_face_idx,
                            sigmainv=7000, boxlen=0.02, knum=30, multiplier=1000.):
    """dibr.py:58-73 (recomputes the forward to obtain the saved K-lists)."""
    fvi_m = _f32(face_vertices_image) * np.float32(multiplier)
    bb = _large_bboxes(fvi_m, boxlen, multiplier)
    soft, prob, cidx, ctype = soft_mask_forward_op(
        fvi_m, bb, selected_face_idx, sigmainv, knum, multiplier, with_lists=True)
    return soft_mask_backward_op(grad_soft_mask, soft, selected_face_idx, prob, cidx, ctype,
                                 fvi_m, sigmainv, multiplier)


def dibr_rasterization(height, width, face_vertices_z, face_vertices_image, face_features,
                       face_normals_z, sigmainv=7000, boxlen=0.02, knum=30,
                       multiplier=None, eps=None, return_weights=False):
    """dibr.py:190-209."""
    res = rasterize(height, width, face_vertices_z, face_vertices_image, face_features,
                    np.asarray(face_normals_z) >= 0., multiplier, eps,
                    return_weights=return_weights)
    feats, face_idx = res[0], res[1]
    _m = 1000. if multiplier is None else multiplier
    soft = dibr_soft_mask(face_vertices_image, face_idx, sigmainv, boxlen, knum, _m)
    if return_weights:
        return feats, soft, face_idx, res[2]
    return feats, soft, face_idx


def dibr_rasterization_backward(grad_features, grad_soft_mask, face_idx, output_weights,
                                face_vertices_image, face_features, sigmainv=7000,
                                boxlen=0.02, knum=30, multiplier=None, eps=None):
    """Sum of both backward branches, as autograd does."""
    is_list = isinstance(face_features, (list, tuple))
    ff = np.concatenate([_f32(f) for f in face_features], -1) if is_list else _f32(face_features)
    g = np.concatenate([_f32(x) for x in grad_features], -1) if isinstance(
        grad_features, (list, tuple)) else _f32(grad_features)
    gxy_r, gff = rasterize_backward(g, face_idx, output_weights, face_vertices_image, ff, eps)
    _m = 1000. if multiplier is None else multiplier
    gxy_s = dibr_soft_mask_backward(grad_soft_mask, face_vertices_image, face_idx,
                                    sigmainv, boxlen, knum, _m)
    return gxy_r.astype(np.float64) + gxy_s.astype(np.float64), gff, gxy_r, gxy_s


# --------------------------------------------------------------------------
# bounded-sample driver for bench.py's cpu_baseline / --impl reference legs
# --------------------------------------------------------------------------
class RowSample:
    """DIB-R forward+backward of image rows [row0,row1) of every view on the CPU.

    The host-side preparation (packing, bboxes) is done once in __init__ (the
    reference does it with a dozen torch kernels per call; it is negligible next
    to the brute-force pixel loops), ``run()`` executes the four restated
    kernels on the strip with the OpenMP threads set by ``set_threads``.
    """

    def __init__(self, height, width, fvz, fvi, ff, fnz, g_feat, g_soft, row0, row1,
                 sigmainv=7000, boxlen=0.02, knum=30, multiplier=1000., eps=1e-8, strips=None):
        self.H, self.W, self.row0, self.row1 = height, width, int(row0), int(row1)
        # strips: list of (row0,row1) blocks; default one block [row0,row1)
        self.strips = [(int(a), int(b)) for a, b in strips] if strips else [(int(row0), int(row1))]
        self.m, self.eps, self.sigmainv, self.K = float(multiplier), float(eps), float(sigmainv), int(knum)
        fvz, fvi, ff = _f32(fvz), _f32(fvi), _f32(ff)
        B, F = fvz.shape[:2]
        D = ff.shape[-1]
        self.B, self.F, self.D = B, F, D
        vf = np.asarray(fnz) >= 0.
        b_idx, f_idx = np.nonzero(vf)
        self.f_idx = f_idx
        self.first = np.zeros(B + 1, np.int64)
        np.cumsum(vf.sum(1), out=self.first[1:])
        self.p_xy = np.ascontiguousarray(fvi[b_idx, f_idx] * np.float32(multiplier))
        self.p_z = np.ascontiguousarray(fvz[b_idx, f_idx])
        self.p_ff = np.ascontiguousarray(ff[b_idx, f_idx])
        self.p_bb = np.ascontiguousarray(np.concatenate([self.p_xy.min(1), self.p_xy.max(1)], 1))
        self.fvi, self.ff = fvi, ff
        self.fvi_m = np.ascontiguousarray(fvi * np.float32(multiplier))
        self.bb_large = np.ascontiguousarray(_large_bboxes(self.fvi_m, boxlen, multiplier))
        self.g_feat, self.g_soft = _f32(g_feat), _f32(g_soft)
        P = (B, height, width)
        self.sel = np.full(P, -1, np.int64)
        self.idx = np.full(P, -1, np.int64)
        self.w = np.zeros(P + (3,), np.float32)
        self.out = np.zeros(P + (D,), np.float32)
        self.soft = np.zeros(P, np.float32)
        self.prob = np.zeros(P + (knum,), np.float32)
        self.cidx = np.full(P + (knum,), -1, np.int64)
        self.ctype = np.zeros(P + (knum,), np.uint8)
        self.gxy = np.zeros_like(fvi)
        self.gxy2 = np.zeros_like(fvi)
        self.gff = np.zeros_like(ff)

    @property
    def pixels(self):
        return self.B * sum(b - a for a, b in self.strips) * self.W

    def run(self):
        for a, b in self.strips:
            self.row0, self.row1 = a, b
            out = self._run_strip()
        return out

    def _run_strip(self):
        L = lib()
        c_i, c_f = ctypes.c_int, ctypes.c_float
        B, H, W, F, D, K = self.B, self.H, self.W, self.F, self.D, self.K
        r0, r1 = c_i(self.row0), c_i(self.row1)
        L.oracle_rasterize_forward_rows(
            c_i(B), c_i(H), c_i(W), c_i(D), _p(self.p_z, _f32p), _p(self.p_xy, _f32p),
            _p(self.p_bb, _f32p), _p(self.p_ff, _f32p), _p(self.first, _i64p), c_f(self.m),
            c_f(self.eps), _p(self.sel, _i64p), _p(self.w, _f32p), _p(self.out, _f32p), r0, r1)
        rows = slice(self.row0, self.row1)
        sel = self.sel[:, rows]
        cov = sel >= 0
        packed = sel + self.first[:-1].reshape(-1, 1, 1)
        idx = np.full(sel.shape, -1, np.int64)
        idx[cov] = self.f_idx[packed[cov]]
        self.idx[:, rows] = idx
        L.oracle_soft_mask_forward_rows(
            c_i(B), c_i(H), c_i(W), c_i(F), c_i(K), _p(self.fvi_m, _f32p), _p(self.bb_large, _f32p),
            _p(self.idx, _i64p), c_f(self.sigmainv), c_f(self.m), _p(self.soft, _f32p),
            _p(self.prob, _f32p), _p(self.cidx, _i64p), _p(self.ctype, _u8p), r0, r1)
        L.oracle_rasterize_backward_rows(
            c_i(B), c_i(H), c_i(W), c_i(F), c_i(D), _p(self.g_feat, _f32p), _p(self.idx, _i64p),
            _p(self.w, _f32p), _p(self.fvi, _f32p), _p(self.ff, _f32p), c_f(self.eps),
            _p(self.gxy, _f32p), _p(self.gff, _f32p), r0, r1)
        L.oracle_soft_mask_backward_rows(
            c_i(B), c_i(H), c_i(W), c_i(F), c_i(K), _p(self.g_soft, _f32p), _p(self.soft, _f32p),
            _p(self.idx, _i64p), _p(self.prob, _f32p), _p(self.cidx, _i64p), _p(self.ctype, _u8p),
            _p(self.fvi_m, _f32p), c_f(self.sigmainv), c_f(self.m), _p(self.gxy2, _f32p), r0, r1)
        return self.gxy, self.gxy2, self.gff
