"""TEST INFRASTRUCTURE — restatement of Kaolin's DefTet volumetric renderer.
Never imported by kaolin_b200.

  forward_op     kaolin/csrc/render/mesh/deftet_cuda.cu:31-194 (the first knum faces in INDEX order
                 that contain the point, with depth in [min, max)) — numpy
  sparse_render  kaolin/render/mesh/deftet.py:269-334 (depth sort, w2 = 1 - w0 - w1 on hits,
                 feature interpolation) — numpy on top of forward_op
  sparse_render_torch  the same lines in torch, vectorised over (point, face); gradients through
                 autograd (the role _naive_deftet_sparse_render, deftet.py:101-267, plays for the
                 reference's own tests)

Parity pinned: tests/test_pipeline_oracle_cpu.py::test_deftet_* check these functions against
tests/golden/deftet.npz = outputs and autograd gradients of the reference's
_naive_deftet_sparse_render imported in place (tests/golden/make_deftet_golden.py).
"""
import numpy as np


def forward_op(face_vertices_z, face_vertices_image, face_bboxes, pixel_coords, render_ranges, knum, eps):
    fvz = np.asarray(face_vertices_z, np.float32); fvi = np.asarray(face_vertices_image, np.float32)
    bb = np.asarray(face_bboxes, np.float32); pix = np.asarray(pixel_coords, np.float32)
    rr = np.asarray(render_ranges, np.float32)
    B, F = fvz.shape[:2]
    P = pix.shape[1]
    idx = np.full((B, P, knum), -1, np.int64)
    depth = np.full((B, P, knum), -np.inf, np.float32)            # deftet.cpp:90-93
    w0o = np.zeros((B, P, knum), np.float32); w1o = np.zeros((B, P, knum), np.float32)
    eps = np.float32(eps)
    for b in range(B):
        ax, ay = fvi[b, :, 0, 0], fvi[b, :, 0, 1]
        bx, by = fvi[b, :, 1, 0], fvi[b, :, 1, 1]
        cx, cy = fvi[b, :, 2, 0], fvi[b, :, 2, 1]
        for p in range(P):
            x0, y0 = pix[b, p]
            inb = (x0 >= bb[b, :, 0]) & (x0 < bb[b, :, 2]) & (y0 >= bb[b, :, 1]) & (y0 < bb[b, :, 3])   # :118
            f = np.nonzero(inb)[0]
            if f.size == 0:
                continue
            aex, aey = ax[f] - x0, ay[f] - y0
            bex, bey = bx[f] - x0, by[f] - y0
            cex, cey = cx[f] - x0, cy[f] - y0
            u0 = bex * cey - bey * cex                                                                   # :133-135
            u1 = cex * aey - cey * aex
            u2 = aex * bey - aey * bex
            norm = u0 + u1 + u2
            den = norm + np.copysign(eps, norm).astype(np.float32)                                        # :137-141
            with np.errstate(divide="ignore", invalid="ignore"):
                w0, w1, w2 = u0 / den, u1 / den, u2 / den
            d = w0 * fvz[b, f, 0] + w1 * fvz[b, f, 1] + w2 * fvz[b, f, 2]                                 # :153
            ok = (w0 >= 0) & (w1 >= 0) & (w2 >= 0) & (d < rr[b, p, 1]) & (d >= rr[b, p, 0])               # :144,155
            sel = np.nonzero(ok)[0][:knum]                                                                # index order, first knum
            n = sel.size
            idx[b, p, :n] = f[sel]; depth[b, p, :n] = d[sel]; w0o[b, p, :n] = w0[sel]; w1o[b, p, :n] = w1[sel]
    return idx, depth, w0o, w1o


def sparse_render(pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features, knum=300,
                  eps=1e-8):
    fvi = np.asarray(face_vertices_image, np.float32); ff = np.asarray(face_features, np.float32)
    bb = np.concatenate([fvi.min(2), fvi.max(2)], -1)                                                     # deftet.py:288-290
    idx, depth, w0, w1 = forward_op(face_vertices_z, fvi, bb, pixel_coords, render_ranges, knum, eps)
    order = np.argsort(-depth, axis=-1, kind="stable")                                                    # :301
    idx = np.take_along_axis(idx, order, -1)
    w0 = np.take_along_axis(w0, order, -1); w1 = np.take_along_axis(w1, order, -1)
    w2 = (idx != -1).astype(np.float32) - (w0 + w1)                                                       # :305
    B, P, K = idx.shape
    pad = np.concatenate([np.zeros_like(ff[:, :1]), ff], 1)                                               # :309-312
    sel = pad[np.arange(B)[:, None, None], idx + 1]                                                       # (B,P,K,3,D)
    w = np.stack([w0, w1, w2], -1)
    return (w[..., None] * sel).sum(-2).astype(np.float32), idx                                           # :315-316


def sparse_render_torch(pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features,
                        knum=300, eps=1e-8):
    import torch
    x0 = pixel_coords[:, :, None, 0]; y0 = pixel_coords[:, :, None, 1]                    # (B,P,1)
    fvi = face_vertices_image
    ax, ay = fvi[:, None, :, 0, 0], fvi[:, None, :, 0, 1]
    bx, by = fvi[:, None, :, 1, 0], fvi[:, None, :, 1, 1]
    cx, cy = fvi[:, None, :, 2, 0], fvi[:, None, :, 2, 1]
    mn, mx = fvi.min(2)[0], fvi.max(2)[0]
    inb = (x0 >= mn[:, None, :, 0]) & (x0 < mx[:, None, :, 0]) & (y0 >= mn[:, None, :, 1]) & (y0 < mx[:, None, :, 1])
    aex, aey, bex, bey, cex, cey = ax - x0, ay - y0, bx - x0, by - y0, cx - x0, cy - y0
    u0 = bex * cey - bey * cex; u1 = cex * aey - cey * aex; u2 = aex * bey - aey * bex
    norm = u0 + u1 + u2
    den = norm + eps * torch.where(norm < 0, -torch.ones_like(norm), torch.ones_like(norm))
    w0, w1, w2 = u0 / den, u1 / den, u2 / den
    z = face_vertices_z[:, None]
    d = w0 * z[..., 0] + w1 * z[..., 1] + w2 * z[..., 2]
    ok = inb & (w0 >= 0) & (w1 >= 0) & (w2 >= 0) & (d < render_ranges[:, :, None, 1]) & (d >= render_ranges[:, :, None, 0])
    F = fvi.shape[1]
    rank = torch.cumsum(ok.long(), -1) - 1                                                # index order
    ok = ok & (rank < knum)
    dd = torch.where(ok, d, torch.full_like(d, -float("inf")))
    k = min(knum, F)
    top_d, top_i = torch.topk(dd.detach(), k, dim=-1)                                     # depth descending
    hit = torch.gather(ok, -1, top_i)
    idx = torch.where(hit, top_i, torch.full_like(top_i, -1))
    g = lambda t: torch.gather(t, -1, top_i)
    W0, W1 = g(w0) * hit, g(w1) * hit
    W2 = hit.to(W0.dtype) - (W0 + W1)                                                     # deftet.py:305 (as the CUDA wrapper defines it)
    B, P = idx.shape[:2]
    D = face_features.shape[-1]
    sel = torch.gather(face_features[:, None].expand(B, P, F, 3, D), 2,
                       top_i.clamp(min=0)[..., None, None].expand(B, P, k, 3, D))
    out = (torch.stack([W0, W1, W2], -1)[..., None] * sel).sum(-2)
    if k < knum:
        out = torch.nn.functional.pad(out, (0, 0, 0, knum - k))
        idx = torch.nn.functional.pad(idx, (0, knum - k), value=-1)
    return out, idx
