"""The non-default kernel variants the library keeps behind A/B switches (environment variables read
per call) must produce the same results as the default ones: forward v2 forced to one tile per CTA
(S = 1) and to 2x2 tiles on DENSE scenes (several staging rounds re-streamed per sub-tile), the
first-generation forward kernel, the warp-shuffle rasterize backward, the shuffle-reduction soft-mask
backward and the match.any-aggregated binning."""
import numpy as np
import pytest
import torch

import oracle
from kaolin_b200 import synthetic
from kaolin_b200.render.mesh import dibr_rasterization

pytestmark = pytest.mark.gpu
DEV = "cuda"

SCENES = {
    "ico4_256": (lambda: synthetic.icosphere_views(2, 4, seed=3), 256, 256),
    "soup2000_200x72": (lambda: synthetic.triangle_soup(1, 2000, seed=5), 200, 72),          # ~10^2 candidates per tile
    "soup6000_128_dense": (lambda: synthetic.triangle_soup(1, 6000, seed=6, coverage=6.0), 128, 128),  # > 256 per 32x32 tile
}
VARIANTS = [
    {"DIBR_B200_FWD": "s2"}, {"DIBR_B200_FWD": "s1"}, {"DIBR_B200_FWD": "old"},
    {"DIBR_B200_RASTER_BWD": "warp"}, {"DIBR_B200_SOFT_BWD": "dense"}, {"DIBR_B200_BIN": "warp"},
]


@pytest.mark.parametrize("scene", list(SCENES))
@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k[10:]}={v}" for k, v in e.items()))
def test_variant_equals_oracle(scene, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    gen, H, W = SCENES[scene]
    fvz, fvi, fnz = gen()
    B, F = fvz.shape[:2]
    ff = synthetic.random_features(B, F, 3, seed=11)
    rng = np.random.default_rng(12)
    g_feat = rng.uniform(size=(B, H, W, 3)).astype(np.float32)
    g_soft = rng.uniform(size=(B, H, W)).astype(np.float32)
    T = lambda a: torch.from_numpy(a).to(DEV)
    t_fvi, t_ff = T(fvi).requires_grad_(True), T(ff).requires_grad_(True)
    feat, soft, idx = dibr_rasterization(H, W, T(fvz), t_fvi, t_ff, T(fnz))
    torch.autograd.backward([feat, soft], [T(g_feat), T(g_soft)])
    o_feat, o_soft, o_idx, o_w = oracle.dibr_rasterization(H, W, fvz, fvi, ff, fnz, return_weights=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    np.testing.assert_allclose(feat.detach().cpu().numpy(), o_feat, rtol=0, atol=1e-5)
    np.testing.assert_allclose(soft.detach().cpu().numpy(), o_soft, rtol=0, atol=1e-5)
    o_gxy, o_gff, _, _ = oracle.dibr_rasterization_backward(g_feat, g_soft, o_idx, o_w, fvi, ff)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    assert rel(t_fvi.grad.cpu().numpy(), o_gxy) <= 3e-5      # vs the double-accumulating CPU oracle
    assert rel(t_ff.grad.cpu().numpy(), o_gff) <= 3e-5


def _random_scene(rng, B, F, kind):
    """Triangle soups that stress the conservative block culling of the v2 forward kernel: slivers,
    zero-area and duplicated faces, faces larger than the image, vertices far off screen, exact
    depth ties, many faces through one pixel."""
    c = rng.uniform(-1.1, 1.1, (B, F, 1, 2))
    if kind == "tiny":
        off = rng.normal(0, 0.01, (B, F, 3, 2))
    elif kind == "huge":
        off = rng.normal(0, 1.5, (B, F, 3, 2))
    elif kind == "sliver":
        d = rng.normal(0, 0.3, (B, F, 1, 2))
        off = d * rng.uniform(-1, 1, (B, F, 3, 1)) + rng.normal(0, 1e-4, (B, F, 3, 2))
    else:
        off = rng.normal(0, 0.08, (B, F, 3, 2))
    fvi = (c + off).astype(np.float32)
    fvz = rng.uniform(-3, -1, (B, F, 3)).astype(np.float32)
    if kind == "degenerate":
        fvi[:, ::7, 2] = fvi[:, ::7, 1]                 # zero area
        fvi[:, 1::7] = fvi[:, 0::7][:, :fvi[:, 1::7].shape[1]]   # duplicates -> exact depth ties
        fvz[:, 1::7] = fvz[:, 0::7][:, :fvz[:, 1::7].shape[1]]
        fvi[:, 3::11, 0] = 1e4                          # far off screen
    fnz = rng.uniform(-0.2, 1.0, (B, F)).astype(np.float32)
    return fvz, fvi, fnz


@pytest.mark.parametrize("kind", ["mixed", "tiny", "huge", "sliver", "degenerate"])
def test_forward_v2_fuzz_equals_first_generation_kernel(kind, monkeypatch):
    """face_idx, weights, features and soft mask of the v2 forward kernel (2x2 tiles per CTA, conservative
    block culling, shared-reciprocal divisions) are BIT-IDENTICAL to the first-generation tile kernel
    (per-lane walk over every rectangle hit) on adversarial scenes, image sizes that are not multiples of 32
    and other multipliers."""
    rng = np.random.default_rng({"mixed": 1, "tiny": 2, "huge": 3, "sliver": 4, "degenerate": 5}[kind])
    T = lambda a: torch.from_numpy(a).to(DEV)
    for trial, (H, W, F, mult) in enumerate([(96, 128, 700, None), (77, 203, 1500, 1000.), (160, 64, 300, 37.5),
                                             (256, 256, 4000, 1e4)]):
        fvz, fvi, fnz = _random_scene(rng, 2, F, kind)
        ff = synthetic.random_features(2, F, 3, seed=trial)
        outs = {}
        for v in ("old", "s2", "s1"):
            monkeypatch.setenv("DIBR_B200_FWD", v)
            feat, soft, idx = dibr_rasterization(H, W, T(fvz), T(fvi), T(ff), T(fnz), multiplier=mult)
            outs[v] = (feat, soft, idx)
        for v in ("s2", "s1"):
            assert torch.equal(outs[v][2], outs["old"][2]), (kind, trial, v, "face_idx")
            assert torch.equal(outs[v][0], outs["old"][0]), (kind, trial, v, "features")
            assert torch.equal(outs[v][1], outs["old"][1]), (kind, trial, v, "soft_mask")
        assert (outs["old"][2] >= 0).any()
