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
