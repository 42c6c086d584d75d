"""Parity of the sm_100a kernels (through the C ABI / public API) with the oracle.

Bars (BASELINE.json north_star): face_idx / close_face_idx / dist_type bit-exact;
forward floats within 1e-5 (they are in fact produced by the same operation
trees, so most are bit-equal); gradients within 1e-5 of the gradient's scale
(max |err| <= 1e-5 * max |ref| — the reference itself is run-to-run
non-deterministic at this level because it accumulates with float atomics).
Where oracle/_ref (the reference's own CUDA kernels, compiled in place) is
present, it is used as a second, independent oracle.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import ref_cuda
from kaolin_b200 import synthetic
from kaolin_b200 import _C as b200_C
from kaolin_b200.render.mesh import rasterize, dibr_soft_mask, dibr_rasterization

pytestmark = pytest.mark.gpu
DEV = "cuda"
# Gradients: every per-pixel term is produced by the reference's own operation tree
# (dibr_math.cuh), so the only difference left is the ORDER of the fp32 additions —
# which the reference does not fix either (float atomicAdd).  Against the oracle's
# double-precision accumulation that noise is ~1e-6 of the gradient's scale for
# meshes and reaches ~1.5e-5 on the overlapping-soup scene (thousands of pixels per
# face); the reference's own kernels show the same deviation (printed below).
GRAD_REL = 3e-5          # against the double-accumulating CPU oracle only
GRAD_REL_REF = 1e-5      # against the reference's own CUDA kernels (north_star: "within 1e-5")


def assert_grad_close(a, ref, tol=GRAD_REL_REF, what="grad"):
    """max-normalised error <= tol AND element-wise allclose(rtol=tol, atol=tol * scale)."""
    scale = max(float(np.abs(ref).max()), 1e-30)
    e = float(np.abs(a - ref).max() / scale)
    assert e <= tol, (what, e)
    assert np.allclose(a, ref, rtol=tol, atol=tol * scale), what
    return e


def T(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t.requires_grad_(True) if grad else t


def N(t):
    return t.detach().cpu().numpy()


def rel_err(a, ref):
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


SCENES = {
    "ico1_64": (lambda: synthetic.icosphere_views(1, 1, seed=1), 64, 64),          # BASELINE configs[0]
    "ico3_96x80": (lambda: synthetic.icosphere_views(2, 3, seed=2), 96, 80),
    "ico4_256": (lambda: synthetic.icosphere_views(2, 4, seed=3), 256, 256),       # configs[1] shape, B=2
    "soup300_37x53": (lambda: synthetic.triangle_soup(2, 300, seed=3, coverage=3.0), 37, 53),
    "soup_big_faces_130": (lambda: synthetic.triangle_soup(2, 60, seed=4, coverage=40.0), 130, 130),
    "soup2000_200x72": (lambda: synthetic.triangle_soup(1, 2000, seed=5), 200, 72),
}


def _run_fused(fvz, fvi, fnz, ff, H, W, g_feat, g_soft, **kw):
    t_fvi, t_ff = T(fvi, True), T(ff, True)
    feat, soft, idx = dibr_rasterization(H, W, T(fvz), t_fvi, t_ff, T(fnz), **kw)
    torch.autograd.backward([feat, soft], [T(g_feat), T(g_soft)])
    return N(feat), N(soft), N(idx), N(t_fvi.grad), N(t_ff.grad)


@pytest.mark.parametrize("name", list(SCENES))
def test_fused_dibr_vs_oracle(name):
    gen, H, W = SCENES[name]
    fvz, fvi, fnz = gen()
    B, F = fvz.shape[:2]
    D = 3
    ff = synthetic.random_features(B, F, D, seed=11)
    rng = np.random.default_rng(12)
    g_feat = rng.uniform(size=(B, H, W, D)).astype(np.float32)
    g_soft = rng.uniform(size=(B, H, W)).astype(np.float32)
    feat, soft, idx, g_fvi, g_ff = _run_fused(fvz, fvi, fnz, ff, H, W, g_feat, g_soft)

    o_feat, o_soft, o_idx, o_w = oracle.dibr_rasterization(H, W, fvz, fvi, ff, fnz, return_weights=True)
    assert (o_idx >= 0).mean() > 0.01
    assert np.array_equal(idx, o_idx)
    np.testing.assert_allclose(feat, o_feat, rtol=0, atol=1e-5)
    np.testing.assert_allclose(soft, o_soft, rtol=0, atol=1e-5)
    o_gxy, o_gff, o_gr, o_gs = oracle.dibr_rasterization_backward(g_feat, g_soft, o_idx, o_w, fvi, ff)
    assert np.abs(o_gs).max() > 0 and np.abs(o_gr).max() > 0
    assert rel_err(g_fvi, o_gxy) <= GRAD_REL, rel_err(g_fvi, o_gxy)
    assert rel_err(g_ff, o_gff) <= GRAD_REL, rel_err(g_ff, o_gff)


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref (reference CUDA build) not present")
@pytest.mark.parametrize("name", list(SCENES))
def test_fused_dibr_vs_reference_cuda(name):
    """The reference's own kernels on the same GPU: face_idx bit-exact, 1e-5 elsewhere."""
    gen, H, W = SCENES[name]
    fvz, fvi, fnz = gen()
    B, F = fvz.shape[:2]
    ff = synthetic.random_features(B, F, 3, seed=11)
    rng = np.random.default_rng(12)
    g_feat = rng.uniform(size=(B, H, W, 3)).astype(np.float32)
    g_soft = rng.uniform(size=(B, H, W)).astype(np.float32)
    feat, soft, idx, g_fvi, g_ff = _run_fused(fvz, fvi, fnz, ff, H, W, g_feat, g_soft)
    r = ref_cuda.dibr_forward_backward(H, W, T(fvz), T(fvi), T(ff), T(fnz), T(g_feat), T(g_soft))
    o_feat, o_soft, o_idx, o_w = oracle.dibr_rasterization(H, W, fvz, fvi, ff, fnz, return_weights=True)
    o_gxy, o_gff, _, _ = oracle.dibr_rasterization_backward(g_feat, g_soft, o_idx, o_w, fvi, ff)
    print(f"\n[{name}] grad_fvi rel err vs double-accumulated oracle: ours {rel_err(g_fvi, o_gxy):.2e}, "
          f"reference CUDA {rel_err(N(r['grad_fvi']), o_gxy):.2e}; ours vs reference CUDA "
          f"{rel_err(g_fvi, N(r['grad_fvi'])):.2e}; soft_mask bit-equal to reference: "
          f"{float((soft == N(r['soft_mask'])).mean()):.4f}")
    assert np.array_equal(idx, N(r["face_idx"]))
    np.testing.assert_allclose(feat, N(r["features"]), rtol=0, atol=1e-5)
    np.testing.assert_allclose(soft, N(r["soft_mask"]), rtol=0, atol=1e-5)
    assert_grad_close(g_fvi, N(r["grad_fvi"]), what="grad_fvi")     # both sides accumulate in fp32 atomics
    assert_grad_close(g_ff, N(r["grad_ff"]), what="grad_ff")


CONFIG_CASES = {
    # BASELINE.json configs at (or cut down from) their full sizes, against the reference's own
    # CUDA kernels on the same GPU.  (name: views, icosphere level, H, W, vertex jitter)
    "c2_full_8x5120f_256": (8, 4, 256, 256, 0.05),            # configs[1] exactly
    "c3_cut_4x20480f_512": (4, 5, 512, 512, 0.05),            # configs[2], 4 of 64 views
    "c4_cut_2x20480f_1024": (2, 5, 1024, 1024, 0.05),         # configs[3], 2 of 256 views (fp32 features)
    "c5_cut_1x1310720f_2048": (1, 8, 2048, 2048, 0.05 / 8),   # configs[4], 1 of 8 views, 1.3 M triangles
    # same mesh with the jitter 8x the edge length: slivers, >2000 soft-mask candidates per tile
    "c5_spiky_1x1310720f_2048": (1, 8, 2048, 2048, 0.05),
}


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref (reference CUDA build) not present")
@pytest.mark.parametrize("name", list(CONFIG_CASES))
def test_baseline_configs_vs_reference_cuda(name):
    B, level, H, W, jitter = CONFIG_CASES[name]
    fvz, fvi, fnz = synthetic.icosphere_views(B, level, seed=1234, jitter=jitter, same_mesh=(level >= 8))
    F = fvz.shape[1]
    ff = synthetic.random_features(B, F, 3, seed=99)
    gen = torch.Generator(device=DEV); gen.manual_seed(7)
    g_feat = torch.rand((B, H, W, 3), device=DEV, generator=gen)
    g_soft = torch.rand((B, H, W), device=DEV, generator=gen)
    t_fvz, t_fnz = T(fvz), T(fnz)
    t_fvi, t_ff = T(fvi, True), T(ff, True)
    feat, soft, idx = dibr_rasterization(H, W, t_fvz, t_fvi, t_ff, t_fnz)
    torch.autograd.backward([feat, soft], [g_feat, g_soft])
    r = ref_cuda.dibr_forward_backward(H, W, t_fvz, t_fvi.detach(), t_ff.detach(), t_fnz, g_feat, g_soft)
    assert torch.equal(idx, r["face_idx"])                                   # bit-exact
    assert 0.2 < (idx >= 0).float().mean().item() < 0.9
    assert (feat - r["features"]).abs().max().item() <= 1e-5
    assert (soft - r["soft_mask"]).abs().max().item() <= 1e-5
    bit_equal = (soft == r["soft_mask"]).float().mean().item()
    e_xy = rel_err(N(t_fvi.grad), N(r["grad_fvi"]))
    e_ff = rel_err(N(t_ff.grad), N(r["grad_ff"]))
    print(f"\n[{name}] face_idx exact; soft_mask bit-equal {bit_equal:.6f}; grad rel err fvi {e_xy:.2e} ff {e_ff:.2e}")
    assert bit_equal > 0.9999
    assert_grad_close(N(t_fvi.grad), N(r["grad_fvi"]), what="grad_fvi")
    assert_grad_close(N(t_ff.grad), N(r["grad_ff"]), what="grad_ff")


def test_two_call_backward_with_feature_grad_hook_equals_fused():
    """dibr_b200_backward split at the branch boundary (DIBR_B200_ACCUMULATE; the hook the
    multi-GPU overlap uses) gives the gradients of the single fused call."""
    from kaolin_b200.render.mesh import _host
    fvz, fvi, fnz = synthetic.icosphere_views(3, 4, seed=11)
    B, F = fvz.shape[:2]
    H, W = 160, 144
    ff = synthetic.random_features(B, F, 3, seed=5)
    gen = torch.Generator(device=DEV); gen.manual_seed(3)
    g_feat = torch.rand((B, H, W, 3), device=DEV, generator=gen)
    g_soft = torch.rand((B, H, W), device=DEV, generator=gen)
    grads, seen = [], []
    for hook in (None, lambda g: seen.append(g.clone())):
        t_fvi, t_ff = T(fvi, True), T(ff, True)
        feat, soft, idx = dibr_rasterization(H, W, T(fvz), t_fvi, t_ff, T(fnz))
        if hook is not None:
            soft.grad_fn.feature_grad_hook = hook       # per-node state (what OverlappedGradAllGather.attach sets)
        torch.autograd.backward([feat, soft], [g_feat, g_soft])
        grads.append((N(t_fvi.grad), N(t_ff.grad)))
    assert len(seen) == 1 and rel_err(N(seen[0]), grads[1][1]) == 0.0   # g_ff was final at the hook
    assert rel_err(grads[1][0], grads[0][0]) <= 1e-6                      # float atomics: order only
    assert rel_err(grads[1][1], grads[0][1]) <= 1e-6


def test_backward_through_one_output_only():
    """Undefined output gradients are not materialised: backward through features only or the
    soft mask only runs one branch, and the two add up to the gradient through both."""
    fvz, fvi, fnz = synthetic.icosphere_views(2, 3, seed=13)
    B, F = fvz.shape[:2]
    H, W = 96, 112
    ff = synthetic.random_features(B, F, 3, seed=6)
    gen = torch.Generator(device=DEV); gen.manual_seed(4)
    g_feat = torch.rand((B, H, W, 3), device=DEV, generator=gen)
    g_soft = torch.rand((B, H, W), device=DEV, generator=gen)
    out = {}
    for which in ("feat", "soft", "both"):
        t_fvi, t_ff = T(fvi, True), T(ff, True)
        feat, soft, idx = dibr_rasterization(H, W, T(fvz), t_fvi, t_ff, T(fnz))
        if which == "feat":
            torch.autograd.backward([feat], [g_feat])
        elif which == "soft":
            torch.autograd.backward([soft], [g_soft])
        else:
            torch.autograd.backward([feat, soft], [g_feat, g_soft])
        out[which] = (N(t_fvi.grad), None if t_ff.grad is None else N(t_ff.grad))
    assert out["soft"][1] is None or not out["soft"][1].any()
    assert rel_err(out["feat"][1], out["both"][1]) <= 1e-6
    assert rel_err(out["feat"][0] + out["soft"][0], out["both"][0]) <= 1e-5
    assert np.abs(out["soft"][0]).max() > 0 and np.abs(out["feat"][0]).max() > 0


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref (reference CUDA build) not present")
@pytest.mark.parametrize("D", [3, 5])
def test_bf16_feature_storage_vs_reference_cuda(D):
    """BASELINE configs[3]: bf16 face_features / features / grad_features, fp32 arithmetic.
    The interpolated features equal the reference's fp32 result (on the same bf16-rounded
    inputs) rounded once to bf16; everything geometric is unchanged."""
    from kaolin_b200.render.mesh import _host
    fvz, fvi, fnz = synthetic.icosphere_views(2, 4, seed=21)
    B, F = fvz.shape[:2]
    H, W = 192, 176
    ff16 = T(synthetic.random_features(B, F, D, seed=8)).to(torch.bfloat16)
    gen = torch.Generator(device=DEV); gen.manual_seed(9)
    g_feat16 = torch.rand((B, H, W, D), device=DEV, generator=gen).to(torch.bfloat16)
    g_soft = torch.rand((B, H, W), device=DEV, generator=gen)
    t_fvz, t_fnz = T(fvz), T(fnz)
    t_fvi, t_ff = T(fvi, True), ff16.clone().requires_grad_(True)
    feat, soft, idx = dibr_rasterization(H, W, t_fvz, t_fvi, t_ff, t_fnz)
    assert feat.dtype == torch.bfloat16 and soft.dtype == torch.float32
    torch.autograd.backward([feat, soft], [g_feat16, g_soft])
    assert t_ff.grad.dtype == torch.bfloat16 and t_fvi.grad.dtype == torch.float32
    r = ref_cuda.dibr_forward_backward(H, W, t_fvz, t_fvi.detach(), ff16.float(), t_fnz,
                                       g_feat16.float(), g_soft)
    assert torch.equal(idx, r["face_idx"])
    assert torch.equal(soft, r["soft_mask"])
    assert torch.equal(feat, r["features"].to(torch.bfloat16))            # one rounding, on store
    assert rel_err(N(t_fvi.grad), N(r["grad_fvi"])) <= GRAD_REL_REF
    assert rel_err(N(t_ff.grad.float()), N(r["grad_ff"])) <= 2.0 ** -8      # bf16 rounding of the result
    # the C ABI returns the fp32 accumulation itself
    feat2, idx2, wts2, soft2, ws = _host.forward(3, H, W, t_fvz, t_fvi.detach(), ff16, t_fnz, None, 1000., 1e-8,
                                                 7000., 0.02 * 1000., 30)
    g_fvi, g_ff = _host.backward(H, W, g_feat16, g_soft, idx2, wts2, soft2, t_fvi.detach(), ff16, 1000., 1e-8,
                                 7000., 0.02 * 1000., 30, ws, True)
    assert g_ff.dtype == torch.float32 and torch.equal(feat2, feat)
    assert rel_err(N(g_ff), N(r["grad_ff"])) <= GRAD_REL_REF
    assert rel_err(N(g_fvi), N(r["grad_fvi"])) <= GRAD_REL_REF
    # rasterize alone, tuple features
    (a, b), idx3 = rasterize(H, W, t_fvz, t_fvi.detach(), [ff16[..., :2], ff16[..., 2:]], t_fnz >= 0.)
    assert torch.equal(idx3, idx) and torch.equal(torch.cat([a, b], -1), feat)
    with pytest.raises(RuntimeError):   # half is not a feature storage type
        rasterize(H, W, t_fvz, t_fvi.detach(), ff16.to(torch.float16))


def test_forward_backward_capture_in_a_cuda_graph():
    """The C ABI never synchronises, allocates or touches global state, so a whole
    forward+backward step can be captured once and replayed (launch-bound sizes such as
    BASELINE configs[1] spend most of a step on launch gaps otherwise)."""
    from kaolin_b200.render.mesh import _host
    fvz, fvi, fnz = synthetic.icosphere_views(4, 4, seed=17)
    B, F = fvz.shape[:2]
    H, W = 128, 128
    ff = synthetic.random_features(B, F, 3, seed=2)
    t_fvz, t_fvi, t_ff, t_fnz = T(fvz), T(fvi), T(ff), T(fnz)
    gen = torch.Generator(device=DEV); gen.manual_seed(5)
    g_feat = torch.rand((B, H, W, 3), device=DEV, generator=gen)
    g_soft = torch.rand((B, H, W), device=DEV, generator=gen)

    def step():
        feat, idx, wts, soft, ws = _host.forward(3, H, W, t_fvz, t_fvi, t_ff, t_fnz, None, 1000., 1e-8,
                                                 7000., 20., 30)
        g_fvi, g_ff = _host.backward(H, W, g_feat, g_soft, idx, wts, soft, t_fvi, t_ff, 1000., 1e-8,
                                     7000., 20., 30, ws, True)
        return feat, idx, soft, g_fvi, g_ff

    eager = [t.clone() for t in step()]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()                                   # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = step()
    # new inputs through the same buffers: the replay must follow them
    t_fvi.mul_(0.9); t_ff.add_(0.25)
    graph.replay()
    torch.cuda.synchronize()
    replayed = [t.clone() for t in outs]
    fresh = step()
    assert torch.equal(replayed[1], fresh[1]) and torch.equal(replayed[0], fresh[0])
    assert torch.equal(replayed[2], fresh[2])
    assert rel_err(N(replayed[3]), N(fresh[3])) <= 1e-6 and rel_err(N(replayed[4]), N(fresh[4])) <= 1e-6
    assert not torch.equal(eager[0], fresh[0])   # the inputs did change


def test_composition_equals_separate_calls():
    """test_dibr.py:482-529: dibr_rasterization == rasterize + dibr_soft_mask, torch.equal."""
    fvz, fvi, fnz = synthetic.icosphere_views(3, 3, seed=7)
    B, F = fvz.shape[:2]
    uv = synthetic.random_features(B, F, 2, seed=1)
    ones = np.ones((B, F, 3, 1), np.float32)
    H, W = 70, 90
    for kw in ({}, {"sigmainv": 70, "boxlen": 0.2, "knum": 20, "multiplier": 100, "eps": 1e-7}):
        (a_uv, a_m), a_soft, a_idx = dibr_rasterization(H, W, T(fvz), T(fvi), [T(uv), T(ones)], T(fnz), **kw)
        rkw = {k: kw[k] for k in ("multiplier", "eps") if k in kw}
        (b_uv, b_m), b_idx = rasterize(H, W, T(fvz), T(fvi), [T(uv), T(ones)], T(fnz) >= 0., **rkw)
        skw = {k: kw[k] for k in ("sigmainv", "boxlen", "knum") if k in kw}
        b_soft = dibr_soft_mask(T(fvi), b_idx, multiplier=kw.get("multiplier", 1000.), **skw)
        assert torch.equal(a_idx, b_idx)
        assert torch.equal(a_uv, b_uv) and torch.equal(a_m, b_m)
        assert torch.equal(a_soft, b_soft)
        assert a_idx.dtype == torch.int64 and a_uv.shape == (B, H, W, 2) and a_m.shape == (B, H, W, 1)


@pytest.mark.parametrize("with_valid", [False, True])
def test_rasterize_api_vs_naive_golden(golden_dir, with_valid):
    """test_rasterization.py:137-289 against the stored outputs of the reference's naive oracle."""
    g = np.load(os.path.join(golden_dir, "rasterize_model.npz"))
    tag = "valid" if with_valid else "all"
    fvi, fvz, uvs = g["fvi"], g["fvz"], g["face_uvs"]
    kwargs = {"valid_faces": T(g["valid_faces"])} if with_valid else {}
    t_fvz, t_fvi, t_uv = T(fvz, True), T(fvi, True), T(uvs, True)
    t_ones = T(np.ones_like(uvs[..., :1]), True)
    (uv_map, mask), face_idx = rasterize(32, 32, t_fvz, t_fvi, [t_uv, t_ones], backend="cuda", **kwargs)
    assert torch.equal(face_idx.cpu(), torch.from_numpy(g[tag + "_face_idx"].astype(np.int64)))
    feats = torch.cat([uv_map, mask], -1)
    np.testing.assert_allclose(N(feats), g[tag + "_features"], rtol=1e-5, atol=1e-5)
    feats.backward(T(g["grad_out"]))
    assert t_fvz.grad is None or torch.all(t_fvz.grad == 0.)          # test_rasterization.py:227
    np.testing.assert_allclose(N(t_fvi.grad), g[tag + "_grad_fvi"], rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(N(t_uv.grad), g[tag + "_grad_uvs"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(N(t_ones.grad), g[tag + "_grad_ones"], rtol=1e-3, atol=1e-3)
    assert rel_err(N(t_fvi.grad), g[tag + "_grad_fvi"]) <= 1e-4


def _mask_iou(soft, face_idx):
    """kaolin/metrics/render.py:18-41 with the shifted target of test_dibr.py:182-186."""
    mask = face_idx != -1
    shifted = torch.nn.functional.pad(mask, (0, 5))[..., 5:]
    B = soft.shape[0]
    mul = soft * shifted
    add = soft + shifted
    up = torch.sum(mul.reshape(B, -1), dim=1)
    down = torch.sum((add - mul).reshape(B, -1), dim=1)
    return 1.0 - torch.mean(up / (down + 1e-10))


@pytest.mark.parametrize("sigmainv", [7000, 70])
@pytest.mark.parametrize("boxlen", [0.02, 0.2])
@pytest.mark.parametrize("multiplier", [1000, 100, 1])
@pytest.mark.parametrize("knum", [30, 20])
def test_simple_scene_golden(golden_dir, sigmainv, boxlen, multiplier, knum):
    """test_dibr.py:109-191 (TestSimpleDibrSoftMask) at the `_C` operator and API level."""
    g = np.load(os.path.join(golden_dir, "dibr_simple.npz"))
    key = f"s{sigmainv}_b{boxlen}_"
    H, W = 35, 31
    fvi, fvz = T(g["fvi"]), T(g["fvz"])
    ff = torch.zeros(fvz.shape + (1,), device=DEV)
    _, face_idx = rasterize(H, W, fvz, fvi, ff)
    assert torch.equal(face_idx.cpu(), torch.from_numpy(g["face_idx"].astype(np.int64)))
    fvi_m = fvi * multiplier
    pmin = torch.min(fvi_m, dim=-2)[0]
    pmax = torch.max(fvi_m, dim=-2)[0]
    bb = torch.cat([pmin - boxlen * multiplier, pmax + boxlen * multiplier], dim=-1)
    soft, prob, cidx, ctype = b200_C.render.mesh.dibr_soft_mask_forward_cuda(
        fvi_m, bb, face_idx, sigmainv, knum, multiplier)
    gt_soft = torch.from_numpy(g[key + "soft_mask"]).to(DEV)
    assert torch.allclose(soft, gt_soft, atol=1e-5, rtol=1e-5)
    assert torch.equal(cidx.cpu(), torch.from_numpy(g[key + "close_face_idx"][..., :knum].astype(np.int64)))
    assert torch.allclose(prob.cpu(), torch.from_numpy(g[key + "close_face_prob"][..., :knum]),
                          atol=1e-5, rtol=1e-5)
    assert torch.equal(ctype.cpu(), torch.from_numpy(g[key + "close_face_dist_type"][..., :knum]))
    # Python API forward + backward (test_dibr.py:142-191)
    t_fvi = fvi.detach().clone().requires_grad_(True)
    soft2 = dibr_soft_mask(t_fvi, face_idx, sigmainv, boxlen, knum, multiplier)
    assert torch.allclose(soft2, gt_soft, atol=1e-5, rtol=1e-5)
    _mask_iou(soft2, face_idx).backward()
    gt_grad = torch.from_numpy(g[key + "grad_fvi"]).to(DEV)
    assert torch.allclose(t_fvi.grad, gt_grad, rtol=1e-5, atol=1e-5)
    # operator-level backward from the stored K-lists
    s_req = soft.clone().requires_grad_(True)
    _mask_iou(s_req, face_idx).backward()
    g_op = b200_C.render.mesh.dibr_soft_mask_backward_cuda(
        s_req.grad.contiguous(), soft, face_idx, prob, cidx, ctype, fvi_m, sigmainv, multiplier)
    assert torch.allclose(g_op, gt_grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("sigmainv,boxlen", [(7000, 0.02), (70, 0.01)])
@pytest.mark.parametrize("knum", [30, 40])
@pytest.mark.parametrize("flip", [False, True])
def test_sphere_scene_golden(golden_dir, sigmainv, boxlen, knum, flip):
    """test_dibr.py:309-394 (TestDibrSoftMask) at the `_C` operator and API level."""
    g = np.load(os.path.join(golden_dir, "dibr_sphere.npz"))
    key = f"s{sigmainv}_b{boxlen}_"
    H, W = 35, 31
    fvi_np, fvz_np = g["fvi"], g["fvz"]
    if flip:
        fvi_np, fvz_np = fvi_np[:, :, ::-1], fvz_np[:, :, ::-1]
    fvi, fvz = T(fvi_np), T(fvz_np)
    ff = torch.zeros(fvz.shape + (1,), device=DEV)
    _, face_idx = rasterize(H, W, fvz, fvi, ff)
    for multiplier in (1000, 100):
        fvi_m = fvi * multiplier
        pmin = torch.min(fvi_m, dim=-2)[0]
        pmax = torch.max(fvi_m, dim=-2)[0]
        bb = torch.cat([pmin - boxlen * multiplier, pmax + boxlen * multiplier], dim=-1)
        soft, prob, cidx, ctype = b200_C.render.mesh.dibr_soft_mask_forward_cuda(
            fvi_m, bb, face_idx, sigmainv, knum, multiplier)
        assert torch.allclose(soft.cpu(), torch.from_numpy(g[key + "soft_mask"]), atol=1e-5, rtol=1e-5)
        assert torch.equal(cidx.cpu(), torch.from_numpy(g[key + "close_face_idx"][..., :knum].astype(np.int64)))
        assert torch.allclose(prob.cpu(), torch.from_numpy(g[key + "close_face_prob"][..., :knum]),
                              atol=1e-5, rtol=1e-5)
        if not flip:
            mism = ctype.cpu() != torch.from_numpy(g[key + "close_face_dist_type"][..., :knum])
            assert mism.sum() / mism.numel() <= 0.01
    t_fvi = fvi.detach().clone().requires_grad_(True)
    soft2 = dibr_soft_mask(t_fvi, face_idx, sigmainv, boxlen, knum, 1000)
    _mask_iou(soft2, face_idx).backward()
    ref = g[key + "grad_fvi"]
    if flip:
        ref = ref[:, :, ::-1]
    assert torch.allclose(t_fvi.grad.cpu(), torch.from_numpy(np.ascontiguousarray(ref)), rtol=1e-1, atol=1e-1)
    assert rel_err(N(t_fvi.grad), ref) <= 2e-3


def test_operator_level_vs_oracle():
    """The four `_C` operators with the reference's packed arguments (rasterization.py:308-339)."""
    fvz, fvi, fnz = synthetic.icosphere_views(3, 3, seed=31)
    B, F = fvz.shape[:2]
    D = 4
    ff = synthetic.random_features(B, F, D, seed=3)
    H, W = 50, 66
    valid = fnz >= 0
    valid[1] = False                      # a mesh with no valid face at all
    b_idx, f_idx = np.nonzero(valid)
    first = np.zeros(B + 1, np.int64)
    np.cumsum(valid.sum(1), out=first[1:])
    xy = np.ascontiguousarray(fvi[b_idx, f_idx] * np.float32(1000))
    z = np.ascontiguousarray(fvz[b_idx, f_idx])
    feat = np.ascontiguousarray(ff[b_idx, f_idx])
    bbox = np.ascontiguousarray(np.concatenate([xy.min(1), xy.max(1)], 1))
    out, sel, w = b200_C.render.mesh.packed_rasterize_forward_cuda(
        H, W, T(z), T(xy), T(bbox), T(feat), T(first), 1000, 1e-8)
    o_out, o_sel, o_w = oracle.packed_rasterize_forward(H, W, z, xy, bbox, feat, first, 1000, 1e-8)
    assert np.array_equal(N(sel), o_sel)
    assert (o_sel[1] == -1).all() and (o_sel[0] >= 0).any()
    assert np.array_equal(N(w).view(np.uint32), o_w.view(np.uint32))
    np.testing.assert_allclose(N(out), o_out, rtol=0, atol=1e-6)
    # backward operator (original ids, unscaled coordinates)
    _, face_idx, wts = oracle.rasterize(H, W, fvz, fvi, ff, valid, return_weights=True)
    rng = np.random.default_rng(5)
    g = rng.uniform(size=(B, H, W, D)).astype(np.float32)
    gxy, gff = b200_C.render.mesh.rasterize_backward_cuda(
        T(g), T(np.zeros_like(g)), T(face_idx), T(wts), T(fvi), T(ff), 1e-8)
    o_gxy, o_gff = oracle.rasterize_backward(g, face_idx, wts, fvi, ff)
    assert rel_err(N(gxy), o_gxy) <= GRAD_REL and rel_err(N(gff), o_gff) <= GRAD_REL


@pytest.mark.parametrize("D", [1, 2, 5, 9])
def test_feature_dims(D):
    fvz, fvi, fnz = synthetic.icosphere_views(1, 2, seed=8)
    B, F = fvz.shape[:2]
    ff = synthetic.random_features(B, F, D, seed=2)
    H, W = 48, 48
    rng = np.random.default_rng(1)
    g = rng.uniform(size=(B, H, W, D)).astype(np.float32)
    t_fvi, t_ff = T(fvi, True), T(ff, True)
    feat, idx = rasterize(H, W, T(fvz), t_fvi, t_ff)
    feat.backward(T(g))
    o_feat, o_idx, o_w = oracle.rasterize(H, W, fvz, fvi, ff, None, return_weights=True)
    assert np.array_equal(N(idx), o_idx)
    np.testing.assert_allclose(N(feat), o_feat, rtol=0, atol=1e-6)
    o_gxy, o_gff = oracle.rasterize_backward(g, o_idx, o_w, fvi, ff)
    assert rel_err(N(t_fvi.grad), o_gxy) <= GRAD_REL and rel_err(N(t_ff.grad), o_gff) <= GRAD_REL


def test_edge_cases():
    """Ragged / degenerate inputs: exact ties, zero-area faces, faces off screen, all faces
    culled, tiny and non-multiple-of-16 images, knum = 1."""
    fvz, fvi, fnz = synthetic.triangle_soup(1, 40, seed=9, coverage=6.0)
    fvi = np.concatenate([fvi, fvi[:, :10], fvi[:, :5] * 0 + 0.3, fvi[:, :5] + 5.0], 1)
    fvz = np.concatenate([fvz, fvz[:, :10], fvz[:, :5], fvz[:, :5]], 1)
    fnz = np.ones(fvz.shape[:2], np.float32)
    B, F = fvz.shape[:2]
    ff = synthetic.random_features(B, F, 2, seed=1)
    for H, W in ((33, 29), (16, 16), (1, 1), (5, 70), (17, 15)):
        for knum in (1, 30):
            feat, soft, idx = dibr_rasterization(H, W, T(fvz), T(fvi), T(ff), T(fnz), knum=knum)
            o_feat, o_soft, o_idx = oracle.dibr_rasterization(H, W, fvz, fvi, ff, fnz, knum=knum)
            assert np.array_equal(N(idx), o_idx), (H, W)
            np.testing.assert_allclose(N(feat), o_feat, rtol=0, atol=1e-6)
            np.testing.assert_allclose(N(soft), o_soft, rtol=0, atol=1e-6)
    # every face back-facing: nothing rasterized, soft mask still sees all faces (dibr.py:200-208)
    feat, soft, idx = dibr_rasterization(40, 40, T(fvz), T(fvi), T(ff), T(-fnz))
    o_feat, o_soft, o_idx = oracle.dibr_rasterization(40, 40, fvz, fvi, ff, -fnz)
    assert (N(idx) == -1).all() and np.array_equal(N(idx), o_idx)
    np.testing.assert_allclose(N(soft), o_soft, rtol=0, atol=1e-6)
    assert (o_soft > 0).any()


def test_soft_mask_dense_overflow_path():
    """> 1024 enlarged faces over one tile: exercises the windowed first-K walk."""
    rng = np.random.default_rng(4)
    F = 3000
    c = rng.uniform(-0.05, 0.05, size=(1, F, 1, 2))
    fvi = (c + rng.normal(scale=0.004, size=(1, F, 3, 2))).astype(np.float32)
    fvz = rng.uniform(-3, -1, size=(1, F, 3)).astype(np.float32)
    fnz = np.ones((1, F), np.float32)
    ff = synthetic.random_features(1, F, 1, seed=0)
    H = W = 64
    for knum in (30, 2000):
        t_fvi = T(fvi, True)
        feat, soft, idx = dibr_rasterization(H, W, T(fvz), t_fvi, T(ff), T(fnz), boxlen=0.1, knum=knum)
        o_feat, o_soft, o_idx, o_w = oracle.dibr_rasterization(H, W, fvz, fvi, ff, fnz, boxlen=0.1,
                                                              knum=knum, return_weights=True)
        assert np.array_equal(N(idx), o_idx)
        np.testing.assert_allclose(N(soft), o_soft, rtol=0, atol=1e-5)
        g_soft = rng.uniform(size=(1, H, W)).astype(np.float32)
        soft.backward(T(g_soft))
        o_g = oracle.dibr_soft_mask_backward(g_soft, fvi, o_idx, 7000, 0.1, knum, 1000.)
        assert rel_err(N(t_fvi.grad), o_g) <= 2e-5


def test_dense_mesh_single_branch_calls():
    """Aggregated binning (>= 32 faces per tile) with only one bin set in use: rasterize alone
    and dibr_soft_mask alone on a dense mesh equal the fused call."""
    fvz, fvi, fnz = synthetic.icosphere_views(1, 6, seed=5, jitter=0.0125)     # 81 920 faces
    B, F = fvz.shape[:2]
    H = W = 160                                                               # 100 tiles
    ff = synthetic.random_features(B, F, 2, seed=4)
    feat, soft, idx = dibr_rasterization(H, W, T(fvz), T(fvi), T(ff), T(fnz))
    feat_r, idx_r = rasterize(H, W, T(fvz), T(fvi), T(ff), T(fnz) >= 0.)
    soft_s = dibr_soft_mask(T(fvi), idx_r)
    assert torch.equal(idx_r, idx) and torch.equal(feat_r, feat) and torch.equal(soft_s, soft)
    assert 0.2 < (idx >= 0).float().mean().item() < 0.9
    o_feat, o_soft, o_idx = oracle.dibr_rasterization(H, W, fvz, fvi, ff, fnz)
    assert np.array_equal(N(idx), o_idx)
    np.testing.assert_allclose(N(soft), o_soft, rtol=0, atol=1e-5)


@pytest.mark.parametrize("cache_tiles", [0, 3, 10 ** 9])
def test_soft_backward_cache_and_recompute_paths_agree(cache_tiles, monkeypatch):
    """The soft-mask backward streams over the hit cache filled by forward; tiles that
    do not fit are recomputed.  No cache, a 3-tile cache and a full cache must all match."""
    from kaolin_b200.render.mesh import _host
    monkeypatch.setattr(_host, "CACHE_TILE_FRACTION", 1.0 if cache_tiles > 3 else 0.0)
    monkeypatch.setattr(_host, "CACHE_MIN_TILES", min(cache_tiles, 3))
    fvz, fvi, fnz = synthetic.icosphere_views(2, 3, seed=2)
    B, F = fvz.shape[:2]
    H, W = 96, 80
    ff = synthetic.random_features(B, F, 3, seed=11)
    rng = np.random.default_rng(12)
    g_feat = rng.uniform(size=(B, H, W, 3)).astype(np.float32)
    g_soft = rng.uniform(size=(B, H, W)).astype(np.float32)
    feat, soft, idx, g_fvi, g_ff = _run_fused(fvz, fvi, fnz, ff, H, W, g_feat, g_soft)
    o_feat, o_soft, o_idx, o_w = oracle.dibr_rasterization(H, W, fvz, fvi, ff, fnz, return_weights=True)
    o_gxy, o_gff, _, _ = oracle.dibr_rasterization_backward(g_feat, g_soft, o_idx, o_w, fvi, ff)
    assert np.array_equal(idx, o_idx)
    np.testing.assert_allclose(soft, o_soft, rtol=0, atol=1e-6)
    assert rel_err(g_fvi, o_gxy) <= GRAD_REL
    # standalone dibr_soft_mask goes through the same cache
    t_fvi = T(fvi, True)
    s2 = dibr_soft_mask(t_fvi, T(o_idx))
    s2.backward(T(g_soft))
    o_gs = oracle.dibr_soft_mask_backward(g_soft, fvi, o_idx)
    assert rel_err(N(t_fvi.grad), o_gs) <= GRAD_REL


def test_soft_mask_windowed_path_many_tiles():
    """Sub-pixel triangles, > 1024 enlarged faces over every silhouette tile (the regime of
    BASELINE configs[4]): the index-windowed walk runs in many CTAs at once."""
    fvz, fvi, fnz = synthetic.icosphere_views(2, 6, seed=5, same_mesh=True)      # 81 920 faces
    B, F = fvz.shape[:2]
    H = W = 256
    ff = synthetic.random_features(B, F, 2, seed=1)
    t_fvi = T(fvi, True)
    feat, soft, idx = dibr_rasterization(H, W, T(fvz), t_fvi, T(ff), T(fnz), boxlen=0.05)
    o_feat, o_soft, o_idx = oracle.dibr_rasterization(H, W, fvz, fvi, ff, fnz, boxlen=0.05)
    assert np.array_equal(N(idx), o_idx)
    np.testing.assert_allclose(N(soft), o_soft, rtol=0, atol=1e-5)
    rng = np.random.default_rng(2)
    g_soft = rng.uniform(size=(B, H, W)).astype(np.float32)
    soft.backward(T(g_soft))
    o_g = oracle.dibr_soft_mask_backward(g_soft, fvi, o_idx, 7000, 0.05, 30, 1000.)
    assert rel_err(N(t_fvi.grad), o_g) <= GRAD_REL


def test_empty_mesh():
    """No faces at all (the reference's kernels loop over zero faces: background everywhere)."""
    z = lambda *s: torch.zeros(s, device=DEV).requires_grad_(True)
    fvi, ff = z(2, 0, 3, 2), z(2, 0, 3, 3)
    feat, soft, idx = dibr_rasterization(40, 56, z(2, 0, 3), fvi, ff, z(2, 0))
    assert feat.shape == (2, 40, 56, 3) and soft.shape == (2, 40, 56) and idx.shape == (2, 40, 56)
    assert (idx == -1).all() and (feat == 0).all() and (soft == 0).all()
    (feat.sum() + soft.sum()).backward()
    assert fvi.grad.shape == fvi.shape and ff.grad.shape == ff.shape
    out, idx = rasterize(40, 56, z(2, 0, 3), z(2, 0, 3, 2), z(2, 0, 3, 1))
    assert (idx == -1).all() and (out == 0).all()
    soft = dibr_soft_mask(z(2, 0, 3, 2), idx)
    assert (soft == 0).all()


def test_empty_view_shard():
    """batch < world size leaves a rank with zero views (multi_gpu.shard_range): empty outputs and gradients."""
    fvz, fvi, fnz = synthetic.icosphere_views(1, 1, seed=1)
    ff = synthetic.random_features(1, fvz.shape[1], 2)
    for dt in (torch.float32, torch.float64):
        t_fvi, t_ff = T(fvi)[:0].to(dt).requires_grad_(True), T(ff)[:0].to(dt).requires_grad_(True)
        feat, soft, idx = dibr_rasterization(24, 32, T(fvz)[:0].to(dt), t_fvi, t_ff, T(fnz)[:0].to(dt))
        assert feat.shape == (0, 24, 32, 2) and soft.shape == (0, 24, 32) and idx.shape == (0, 24, 32)
        assert feat.dtype == dt and idx.dtype == torch.int64
        (feat.sum() + soft.sum()).backward()
        assert t_fvi.grad.shape == t_fvi.shape and t_ff.grad.shape == t_ff.shape


def test_errors_like_reference():
    fvz, fvi, fnz = synthetic.icosphere_views(1, 1, seed=1)
    ff = synthetic.random_features(1, fvz.shape[1], 2)
    with pytest.raises(ValueError):
        rasterize(8, 8, T(fvz), T(fvi), T(ff), backend="nvdiffrast")
    with pytest.raises(ValueError):
        dibr_rasterization(8, 8, T(fvz), T(fvi), T(ff), T(fnz), rast_backend="nvdiffrast_fwd")
    with pytest.raises(RuntimeError):   # CPU tensors: no CPU path (rasterization.cpp:95-102)
        rasterize(8, 8, torch.from_numpy(fvz), torch.from_numpy(fvi), torch.from_numpy(ff))
    # float64 callers are served by the <double> instantiation: tests/test_f64_gpu.py
    out64, _ = rasterize(8, 8, T(fvz).double(), T(fvi).double(), T(ff).double())
    assert out64.dtype == torch.float64
    with pytest.raises(RuntimeError):   # half precision geometry is not a reference dtype either
        rasterize(8, 8, T(fvz).half(), T(fvi).half(), T(ff).half())
    with pytest.raises(RuntimeError):   # non-contiguous operator argument (checkAllContiguous)
        b200_C.render.mesh.rasterize_backward_cuda(
            torch.zeros(1, 8, 8, 2, device=DEV).transpose(1, 2), torch.zeros(1, 8, 8, 2, device=DEV),
            torch.zeros(1, 8, 8, dtype=torch.long, device=DEV), torch.zeros(1, 8, 8, 3, device=DEV),
            T(fvi), T(ff), 1e-8)


def test_full_size_properties():
    """BASELINE configs[3] per-GPU shape cut to 4 views (1024^2, 20480 faces): properties that
    need no oracle — determinism of the forward, composition equality, background consistency,
    interpolation of constant features, translation of the image by whole tiles."""
    fvz, fvi, fnz = synthetic.icosphere_views(4, 5, seed=77)
    B, F = fvz.shape[:2]
    H = W = 1024
    ones = np.ones((B, F, 3, 1), np.float32)
    uv = synthetic.random_features(B, F, 2, seed=3)
    args = (T(fvz), T(fvi), [T(uv), T(ones)], T(fnz))
    (uv1, m1), s1, i1 = dibr_rasterization(H, W, *args)
    (uv2, m2), s2, i2 = dibr_rasterization(H, W, *args)
    assert torch.equal(i1, i2) and torch.equal(uv1, uv2) and torch.equal(s1, s2)
    cov = i1 >= 0
    assert 0.2 < cov.float().mean().item() < 0.8
    # constant feature interpolates to w0+w1+w2 = 1 (within rounding) on covered, 0 elsewhere
    assert torch.all((m1[..., 0] - 1).abs()[cov] < 1e-5) and torch.all(m1[..., 0][~cov] == 0)
    assert torch.all(s1[cov] == 1) and torch.all((s1 >= 0) & (s1 <= 1))
    # only front-facing faces are drawn
    fn = T(fnz)
    assert torch.all(torch.gather(fn, 1, i1.clamp(min=0).reshape(B, -1)).reshape(B, H, W)[cov] >= 0)
    # the soft mask decays away from the silhouette: pixels farther than boxlen from any face are 0
    (b_uv, b_m), b_idx = rasterize(H, W, args[0], args[1], args[2], fn >= 0)
    assert torch.equal(b_idx, i1) and torch.equal(b_uv, uv1)
    # one view against the CPU oracle on a 128x1024 strip would take minutes; instead compare a
    # 1-view 256x256 render of the same mesh (same faces, different sampling) exactly
    f2, s2b, i2b = dibr_rasterization(256, 256, args[0][:1], args[1][:1], T(uv[:1]), args[3][:1])
    o_f, o_s, o_i = oracle.dibr_rasterization(256, 256, fvz[:1], fvi[:1], uv[:1], fnz[:1])
    assert np.array_equal(N(i2b), o_i)
    np.testing.assert_allclose(N(s2b), o_s, rtol=0, atol=1e-5)
    np.testing.assert_allclose(N(f2), o_f, rtol=0, atol=1e-5)


def test_graphed_fast_path_equals_eager():
    """make_graphed_dibr_rasterization: forward and backward replayed from CUDA graphs give the eager
    results (bit-equal images; gradients up to the atomics order) on new input VALUES of the same shape."""
    from kaolin_b200.render.mesh import make_graphed_dibr_rasterization
    H, W = 128, 160
    fvz, fvi, fnz = synthetic.icosphere_views(2, 3, seed=51)
    ff = synthetic.random_features(2, fvz.shape[1], 3, seed=52)
    f = make_graphed_dibr_rasterization(H, W, T(fvz), T(fvi, True), T(ff, True), T(fnz))
    gen = torch.Generator(device=DEV); gen.manual_seed(53)
    g_feat = torch.rand((2, H, W, 3), device=DEV, generator=gen)
    g_soft = torch.rand((2, H, W), device=DEV, generator=gen)
    for seed in (61, 62):                              # two different scenes through the same graphs
        fvz, fvi, fnz = synthetic.icosphere_views(2, 3, seed=seed)
        a_fvi, a_ff = T(fvi, True), T(ff, True)
        feat, soft, idx = f(T(fvz), a_fvi, a_ff, T(fnz))
        torch.autograd.backward([feat, soft], [g_feat, g_soft])
        b_fvi, b_ff = T(fvi, True), T(ff, True)
        feat2, soft2, idx2 = dibr_rasterization(H, W, T(fvz), b_fvi, b_ff, T(fnz))
        torch.autograd.backward([feat2, soft2], [g_feat, g_soft])
        assert torch.equal(idx, idx2) and torch.equal(soft, soft2) and torch.equal(feat, feat2)
        assert rel_err(N(a_fvi.grad), N(b_fvi.grad)) <= 1e-6 and rel_err(N(a_ff.grad), N(b_ff.grad)) <= 1e-6


@pytest.mark.parametrize("variant", ["rows", "warp", "bf16"])
def test_view_chunked_backward_equals_full_backward(variant, monkeypatch):
    """dibr_b200_backward_views: the backward of views [0,2), [2,5) written into shared full-batch buffers
    (with the feature-gradient hook between the branches) equals one full backward."""
    from kaolin_b200.render.mesh import _host
    if variant == "warp":
        monkeypatch.setenv("DIBR_B200_RASTER_BWD", "warp")
    fvz, fvi, fnz = synthetic.icosphere_views(5, 4, seed=71)
    B, F = fvz.shape[:2]
    H, W = 144, 160
    dt = torch.bfloat16 if variant == "bf16" else torch.float32
    ff = T(synthetic.random_features(B, F, 3, seed=72)).to(dt)
    gen = torch.Generator(device=DEV); gen.manual_seed(73)
    g_feat = torch.rand((B, H, W, 3), device=DEV, generator=gen).to(dt)
    g_soft = torch.rand((B, H, W), device=DEV, generator=gen)
    t_fvz, t_fvi, t_fnz = T(fvz), T(fvi), T(fnz)
    feat, idx, wts, soft, ws = _host.forward(3, H, W, t_fvz, t_fvi, ff, t_fnz, None, 1000., 1e-8, 7000., 20., 30)
    ref_fvi, ref_ff = _host.backward(H, W, g_feat, g_soft, idx, wts, soft, t_fvi, ff, 1000., 1e-8, 7000., 20., 30, ws, True)
    g_fvi = torch.full_like(t_fvi, float("nan"))
    g_ff = torch.full(ff.shape, float("nan"), dtype=torch.float32, device=DEV)
    hooks = []
    for c0, c1 in ((0, 2), (2, 5)):
        _host.backward(H, W, g_feat, g_soft, idx, wts, soft, t_fvi, ff, 1000., 1e-8, 7000., 20., 30, ws, True,
                       feature_grad_hook=lambda g, c0=c0, c1=c1: hooks.append(g[c0:c1].clone()),
                       views=(c0, c1), out=(g_fvi, g_ff))
        assert torch.isnan(g_fvi[c1:]).all() and not torch.isnan(g_fvi[:c1]).any()   # only these views were written
    assert rel_err(N(g_fvi), N(ref_fvi)) <= 1e-6 and rel_err(N(g_ff), N(ref_ff)) <= 1e-6
    assert rel_err(N(torch.cat(hooks)), N(ref_ff)) <= 1e-6          # g_ff was final at each hook
