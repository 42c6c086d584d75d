"""Regenerates tests/golden/deftet.npz from the reference (run in the build container only).

    python tests/golden/make_deftet_golden.py

Outputs and autograd gradients of the reference's pure-PyTorch
``_naive_deftet_sparse_render`` (kaolin/render/mesh/deftet.py:101-267) — the oracle its own tests
compare the CUDA renderer with (tests/python/kaolin/render/mesh/test_deftet.py) — on two seeded
scenes: a random triangle soup seen at random points, and a stack of overlapping layers (several
intersections per point, some points hitting more than knum faces)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_import  # noqa: E402

deftet = ref_import.module("kaolin.render.mesh.deftet")
out = {}
g = torch.Generator().manual_seed(77)


def scene(tag, B, F, P, knum, spread, size):
    c = (torch.rand((B, F, 1, 2), generator=g) * 2 - 1) * spread
    fvi = (c + (torch.rand((B, F, 3, 2), generator=g) - 0.5) * size).requires_grad_(True)
    fvz = -(torch.rand((B, F, 3), generator=g) * 3 + 1)
    ff = torch.rand((B, F, 3, 4), generator=g).requires_grad_(True)
    pix = (torch.rand((B, P, 2), generator=g) * 2 - 1) * spread
    rr = torch.stack([torch.full((B, P), -3.5), torch.full((B, P), -1.2)], -1)
    feat, idx = deftet._naive_deftet_sparse_render(pix, rr, fvz, fvi, ff, knum, eps=1e-8)
    gw = torch.rand(feat.shape, generator=g)
    (feat * gw).sum().backward()
    out.update({f"{tag}_fvi": fvi.detach().numpy(), f"{tag}_fvz": fvz.numpy(), f"{tag}_ff": ff.detach().numpy(),
                f"{tag}_pix": pix.numpy(), f"{tag}_rr": rr.numpy(), f"{tag}_knum": np.int64(knum),
                f"{tag}_feat": feat.detach().numpy(), f"{tag}_idx": idx.numpy(), f"{tag}_gw": gw.numpy(),
                f"{tag}_g_fvi": fvi.grad.numpy(), f"{tag}_g_ff": ff.grad.numpy()})
    print(tag, "hits per point", float((idx >= 0).sum(-1).float().mean()), "max", int((idx >= 0).sum(-1).max()))


scene("soup", 2, 300, 200, 8, 0.9, 0.35)
scene("layers", 1, 120, 150, 64, 0.3, 0.9)     # many overlapping faces per point (no truncation: the naive
                                                # function keeps the knum CLOSEST, the CUDA kernel the first knum by index)
np.savez_compressed(os.path.join(HERE, "deftet.npz"), **out)
print("wrote deftet.npz")
