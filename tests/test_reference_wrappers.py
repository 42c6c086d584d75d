"""INTEGRATION.md Option B, checked: the reference's OWN Python wrappers
(kaolin/render/mesh/rasterization.py, dibr.py) run on top of ``kaolin_b200._C``.

* CPU part (this container only — needs /root/reference): the reference modules are imported in
  place with ``kaolin._C`` replaced by ``kaolin_b200._C``; calling ``rasterize`` /
  ``dibr_soft_mask`` must travel through RasterizeCuda / DibrSoftMaskCuda down to the ctypes shim
  with the right arity and argument order, i.e. fail with the shim's "no CPU path" RuntimeError —
  not with a TypeError / AttributeError.
* GPU part (`-m gpu`, no reference on that box): the torch restatement of those wrappers
  (oracle/ref_cuda.py, line-for-line rasterization.py:290-346 / dibr.py:31-72) driven with
  ``C = kaolin_b200._C.render.mesh`` must reproduce the fused public API: operator boundary and
  fused path agree bit-for-bit on images, and on gradients up to atomics order."""
import numpy as np
import pytest
import torch

from oracle import ref_import
from kaolin_b200 import _C as b200_C
from kaolin_b200 import synthetic


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_reference_wrappers_reach_the_shim_with_the_right_arity():
    import inspect
    ref_import.setup(b200_C)
    rast = ref_import.module("kaolin.render.mesh.rasterization")
    dibr = ref_import.module("kaolin.render.mesh.dibr")
    assert rast._C is b200_C and dibr._C is b200_C
    fvz, fvi, fnz = synthetic.icosphere_views(1, 1, seed=1)
    ff = synthetic.random_features(1, fvz.shape[1], 3, seed=2)
    t = torch.from_numpy
    with pytest.raises(RuntimeError, match="GPU|CUDA|no CPU path"):
        rast.rasterize(32, 32, t(fvz), t(fvi), t(ff), backend="cuda")
    with pytest.raises(RuntimeError, match="GPU|CUDA|no CPU path"):
        dibr.dibr_soft_mask(t(fvi), torch.full((1, 32, 32), -1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="GPU|CUDA|no CPU path"):
        dibr.dibr_rasterization(32, 32, t(fvz), t(fvi), t(ff), t(fnz))
    dt = ref_import.module("kaolin.render.mesh.deftet")
    assert dt._C is b200_C
    with pytest.raises(RuntimeError, match="GPU|CUDA|no CPU path"):        # deftet.py:292-299 -> the shim
        dt.deftet_sparse_render(torch.zeros(1, 5, 2), torch.zeros(1, 5, 2), t(fvz), t(fvi), t(ff), knum=4)
    assert len(inspect.signature(b200_C.render.mesh.deftet_sparse_render_forward_cuda).parameters) == 7
    assert len(inspect.signature(b200_C.render.mesh.deftet_sparse_render_backward_cuda).parameters) == 6
    # backward operators: arity of the shim == arity of the reference's call sites
    sig = lambda f: len(inspect.signature(f).parameters)
    assert sig(b200_C.render.mesh.packed_rasterize_forward_cuda) == 9      # rasterization.py:329-339
    assert sig(b200_C.render.mesh.rasterize_backward_cuda) == 7            # rasterization.py:360-368
    assert sig(b200_C.render.mesh.dibr_soft_mask_forward_cuda) == 6        # dibr.py:40-48
    assert sig(b200_C.render.mesh.dibr_soft_mask_backward_cuda) == 9       # dibr.py:63-72


@pytest.mark.gpu
def test_wrapper_logic_on_b200_operators_equals_fused_api():
    from oracle import ref_cuda
    from kaolin_b200.render.mesh import dibr_rasterization
    dev = "cuda"
    fvz, fvi, fnz = synthetic.icosphere_views(2, 4, seed=21)
    H, W = 192, 160
    ff = synthetic.random_features(2, fvz.shape[1], 3, seed=22)
    T = lambda a: torch.from_numpy(a).to(dev)
    gen = torch.Generator(device=dev); gen.manual_seed(23)
    g_feat = torch.rand((2, H, W, 3), device=dev, generator=gen)
    g_soft = torch.rand((2, H, W), device=dev, generator=gen)
    r = ref_cuda.dibr_forward_backward(H, W, T(fvz), T(fvi), T(ff), T(fnz), g_feat, g_soft,
                                       C=b200_C.render.mesh)
    t_fvi, t_ff = T(fvi).requires_grad_(True), T(ff).requires_grad_(True)
    feat, soft, idx = dibr_rasterization(H, W, T(fvz), t_fvi, t_ff, T(fnz))
    torch.autograd.backward([feat, soft], [g_feat, g_soft])
    assert torch.equal(idx, r["face_idx"]) and torch.equal(soft, r["soft_mask"]) and torch.equal(feat, r["features"])
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(t_fvi.grad, r["grad_fvi"]) <= 1e-5 and rel(t_ff.grad, r["grad_ff"]) <= 1e-5
    if ref_cuda.available():     # and both equal the reference's own operators
        rr = ref_cuda.dibr_forward_backward(H, W, T(fvz), T(fvi), T(ff), T(fnz), g_feat, g_soft)
        assert torch.equal(rr["face_idx"], r["face_idx"])
        assert (rr["soft_mask"] - r["soft_mask"]).abs().max().item() <= 1e-5


@pytest.mark.gpu
def test_float64_operator_callers_are_served_in_fp32():
    """The reference dispatches double in its operators too (AT_DISPATCH_FLOATING_TYPES).  The public
    API has a real float64 instantiation (tests/test_f64_gpu.py); the packed ``_C`` OPERATOR shims serve
    double callers by casting (fp32 arithmetic, float64 outputs) - this pins that contract."""
    from oracle import ref_cuda
    dev = "cuda"
    fvz, fvi, fnz = synthetic.icosphere_views(2, 3, seed=31)
    H, W = 96, 128
    ff = synthetic.random_features(2, fvz.shape[1], 3, seed=32)
    D = lambda a: torch.from_numpy(a).to(dev).double()
    gen = torch.Generator(device=dev); gen.manual_seed(33)
    g_feat = torch.rand((2, H, W, 3), device=dev, generator=gen, dtype=torch.float64)
    g_soft = torch.rand((2, H, W), device=dev, generator=gen, dtype=torch.float64)
    ours = ref_cuda.dibr_forward_backward(H, W, D(fvz), D(fvi), D(ff), D(fnz), g_feat, g_soft, C=b200_C.render.mesh)
    for k in ("features", "weights", "soft_mask", "grad_fvi", "grad_ff"):
        assert ours[k].dtype == torch.float64, k
    assert ours["face_idx"].dtype == torch.int64
    if ref_cuda.available():     # against the reference's <double> kernels: fp32-level agreement
        r = ref_cuda.dibr_forward_backward(H, W, D(fvz), D(fvi), D(ff), D(fnz), g_feat, g_soft)
        same = ours["face_idx"] == r["face_idx"]
        agree = same.float().mean().item()
        print(f"\nfp64 operator callers: face_idx agreement with the reference's double kernels {agree:.6f}")
        assert agree >= 0.999
        assert (ours["features"] - r["features"])[same].abs().max().item() <= 1e-4
        assert (ours["soft_mask"] - r["soft_mask"])[same].abs().max().item() <= 1e-4
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        assert rel(ours["grad_ff"], r["grad_ff"]) <= 1e-3


def _binding():
    from integration import build_binding
    return build_binding.load()


@pytest.mark.skipif(_binding() is None, reason="integration/_build/kaolin_b200_binding.so not built")
def test_option_a_binding_loads_and_checks_like_the_reference():
    """INTEGRATION.md Option A compiled for real (integration/kaolin_binding.cpp): the pybind11 module
    exports the four operator names of bindings.cpp:111-115 and rejects CPU tensors through the same
    at::checkAllSameGPU the reference wrappers use (rasterization.cpp:70-72)."""
    m = _binding()
    for name in ("packed_rasterize_forward_cuda", "rasterize_backward_cuda", "dibr_soft_mask_forward_cuda",
                 "dibr_soft_mask_backward_cuda"):
        assert hasattr(m, name)
    with pytest.raises(RuntimeError, match="expected it to be on GPU"):
        m.packed_rasterize_forward_cuda(8, 8, torch.zeros(4, 3), torch.zeros(4, 3, 2), torch.zeros(4, 4),
                                        torch.zeros(4, 3, 2), torch.tensor([0, 4]), 1000., 1e-8)
    with pytest.raises(RuntimeError, match="expected it to be on GPU"):
        m.dibr_soft_mask_forward_cuda(torch.zeros(1, 4, 3, 2), torch.zeros(1, 4, 4),
                                      torch.zeros(1, 8, 8, dtype=torch.long), 7000., 30, 1000.)


@pytest.mark.gpu
@pytest.mark.skipif(_binding() is None, reason="integration/_build/kaolin_b200_binding.so not built")
def test_option_a_binding_results():
    """The reference's wrapper logic (oracle/ref_cuda.py) on top of the compiled Option A binding ==
    the same logic on top of the ctypes shim == the fused public API."""
    from oracle import ref_cuda
    m = _binding()
    dev = "cuda"
    fvz, fvi, fnz = synthetic.icosphere_views(2, 4, seed=41)
    H, W = 160, 176
    ff = synthetic.random_features(2, fvz.shape[1], 3, seed=42)
    T = lambda a: torch.from_numpy(a).to(dev)
    gen = torch.Generator(device=dev); gen.manual_seed(43)
    g_feat = torch.rand((2, H, W, 3), device=dev, generator=gen)
    g_soft = torch.rand((2, H, W), device=dev, generator=gen)
    a = ref_cuda.dibr_forward_backward(H, W, T(fvz), T(fvi), T(ff), T(fnz), g_feat, g_soft, C=m)
    b = ref_cuda.dibr_forward_backward(H, W, T(fvz), T(fvi), T(ff), T(fnz), g_feat, g_soft, C=b200_C.render.mesh)
    for k in ("face_idx", "soft_mask", "features", "weights"):
        assert torch.equal(a[k], b[k]), k
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
    assert rel(a["grad_fvi"], b["grad_fvi"]) <= 1e-5 and rel(a["grad_ff"], b["grad_ff"]) <= 1e-5
    if ref_cuda.available():
        r = ref_cuda.dibr_forward_backward(H, W, T(fvz), T(fvi), T(ff), T(fnz), g_feat, g_soft)
        assert torch.equal(a["face_idx"], r["face_idx"])
        assert rel(a["grad_fvi"], r["grad_fvi"]) <= 1e-5 and rel(a["grad_ff"], r["grad_ff"]) <= 1e-5
