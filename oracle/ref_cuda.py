"""TEST INFRASTRUCTURE — drives the reference's own CUDA operators (oracle/_ref,
built by oracle/build_ref.py from the unmodified reference .cu/.cpp files) the way
the reference's Python wrappers do, restated in torch so that it runs on the GPU
box where /root/reference does not exist.

  rasterize            <- kaolin/render/mesh/rasterization.py:273-352 (+ :355-371 backward)
  dibr_soft_mask       <- kaolin/render/mesh/dibr.py:29-73
  dibr_rasterization   <- kaolin/render/mesh/dibr.py:190-209

Used as (i) the primary parity oracle of the `-m gpu` tests ("outputs matching
the reference CUDA path") and (ii) the "reference CUDA on the same B200" timing
row of bench.py.  Never imported by kaolin_b200.
"""
import torch

from . import build_ref

_mod = None


def available():
    return module() is not None


def module():
    global _mod
    if _mod is None:
        try:
            _mod = build_ref.load()
        except Exception:  # pragma: no cover - missing/incompatible prebuilt .so
            _mod = None
    return _mod


def rasterize_forward(height, width, fvz, fvi, ff, valid_faces=None, multiplier=1000, eps=1e-8, C=None):
    """RasterizeCuda.forward -> (interp, face_idx, weights).  ``C``: the object playing the role of
    ``kaolin._C.render.mesh`` (default: the reference's own operators, oracle/_ref)."""
    C = C or module()
    B, F = fvz.shape[:2]
    D = ff.shape[-1]
    dev = fvz.device
    ff = ff.contiguous()
    fvi = fvi.contiguous()
    if valid_faces is None:
        vidx = (torch.arange(B, device=dev).reshape(-1, 1).repeat(1, F).reshape(-1),
                torch.arange(F, device=dev).reshape(1, -1).repeat(B, 1).reshape(-1))
        v_xy = fvi.reshape(B * F, 3, 2)
        v_z = fvz.reshape(B * F, 3)
        v_ff = ff.reshape(B * F, 3, D)
        nfpm = torch.full((B,), F, dtype=torch.long, device=dev)
    else:
        vidx = torch.where(valid_faces)
        v_xy = fvi[vidx[0], vidx[1]]
        v_z = fvz[vidx[0], vidx[1]]
        v_ff = ff[vidx[0], vidx[1]]
        nfpm = torch.sum(valid_faces.reshape(B, -1), dim=1)
    first = torch.zeros(B + 1, dtype=torch.long, device=dev)
    torch.cumsum(nfpm, dim=0, out=first[1:])
    v_xy = v_xy * multiplier
    pmin = torch.min(v_xy, dim=1)[0]
    pmax = torch.max(v_xy, dim=1)[0]
    bboxes = torch.cat((pmin, pmax), dim=1)
    interp, sel, w = C.packed_rasterize_forward_cuda(
        height, width, v_z.contiguous(), v_xy.contiguous(), bboxes.contiguous(),
        v_ff.contiguous(), first.contiguous(), multiplier, eps)
    face_idx = vidx[1][(sel + first[:-1].reshape(-1, 1, 1)).reshape(-1)]
    face_idx = face_idx.reshape(sel.shape).contiguous()
    face_idx[sel == -1] = -1
    return interp, face_idx, w


def rasterize_backward(grad, interp, face_idx, w, fvi, ff, eps=1e-8, C=None):
    C = C or module()
    return C.rasterize_backward_cuda(grad.contiguous(), interp, face_idx, w,
                                     fvi.contiguous(), ff.contiguous(), eps)


def soft_mask_forward(fvi, face_idx, sigmainv=7000, boxlen=0.02, knum=30, multiplier=1000., C=None):
    """DibrSoftMaskCuda.forward -> (soft, fvi_m, prob, cidx, ctype)."""
    C = C or module()
    fvi_m = fvi.contiguous() * multiplier
    pmin = torch.min(fvi_m, dim=-2)[0]
    pmax = torch.max(fvi_m, dim=-2)[0]
    bb = torch.cat([pmin - boxlen * multiplier, pmax + boxlen * multiplier], dim=-1)
    soft, prob, cidx, ctype = C.dibr_soft_mask_forward_cuda(
        fvi_m, bb.contiguous(), face_idx.contiguous(), sigmainv, knum, multiplier)
    return soft, fvi_m, prob, cidx, ctype


def soft_mask_backward(grad_soft, soft, face_idx, prob, cidx, ctype, fvi_m, sigmainv=7000,
                       multiplier=1000., C=None):
    C = C or module()
    return C.dibr_soft_mask_backward_cuda(grad_soft.contiguous(), soft, face_idx, prob, cidx,
                                          ctype, fvi_m, sigmainv, multiplier)


def dibr_forward_backward(height, width, fvz, fvi, ff, fnz, g_feat, g_soft, sigmainv=7000,
                          boxlen=0.02, knum=30, multiplier=None, eps=None, C=None):
    """dibr_rasterization forward + both backward branches (summed as autograd does)."""
    m = 1000 if multiplier is None else multiplier
    e = 1e-8 if eps is None else eps
    interp, face_idx, w = rasterize_forward(height, width, fvz, fvi, ff, fnz >= 0., m, e, C=C)
    _m = 1000. if multiplier is None else multiplier
    soft, fvi_m, prob, cidx, ctype = soft_mask_forward(fvi, face_idx, sigmainv, boxlen, knum, _m, C=C)
    out = {"features": interp, "face_idx": face_idx, "weights": w, "soft_mask": soft}
    if g_feat is not None:
        gxy_r, gff = rasterize_backward(g_feat, interp, face_idx, w, fvi, ff, e, C=C)
        gxy_s = soft_mask_backward(g_soft, soft, face_idx, prob, cidx, ctype, fvi_m, sigmainv, _m, C=C)
        out.update(grad_fvi=gxy_r + gxy_s, grad_ff=gff, grad_fvi_raster=gxy_r, grad_fvi_soft=gxy_s)
    return out
