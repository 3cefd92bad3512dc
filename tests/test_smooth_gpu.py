"""the edge-aware smoothness term (losses/loss.py:61-176) and its gradient, through the C ABI
(upr_edge_smooth_loss_f32) and the autograd drop-in, against the golden vectors of the unmodified reference, the NumPy
oracle and the reference's own torch formulation run on the GPU.  Tolerances: loss 5e-6 relative, gradient 2e-6 of its
largest element (fp32 products, fp64 sums; the reference sums in fp32)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def native():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from retinex_image_enhancement_b200 import native as nv
    return nv


def smooth_cases():
    spec = importlib.util.spec_from_file_location("make_golden_smooth", os.path.join(GOLDEN_DIR, "make_golden_smooth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.smooth_cases()


def check(loss3, grad, want_loss, want_grad):
    assert abs(float(loss3[0]) - float(want_loss)) <= 5e-6 * abs(float(want_loss))
    assert abs(float(loss3[1]) + float(loss3[2]) - float(loss3[0])) <= 1e-6 * abs(float(loss3[0]))
    g = grad.cpu().numpy()
    assert np.abs(g - want_grad).max() <= 2e-6 * np.abs(want_grad).max() + 1e-12
    assert np.array_equal(g == 0, want_grad == 0)


def test_golden_reference_vectors(native):
    gold = np.load(os.path.join(GOLDEN_DIR, "smooth_loss.npz"))
    for name, illu, img, lam, alpha in smooth_cases():
        loss3, grad = native.edge_smooth_loss(torch.from_numpy(illu).cuda(), torch.from_numpy(img).cuda(), lam, alpha)
        check(loss3, grad, gold[f"{name}_loss"], gold[f"{name}_grad"])


@pytest.mark.parametrize("b,ci,cs,h,w", [(8, 1, 3, 256, 256), (2, 1, 3, 640, 640), (3, 3, 3, 97, 131), (1, 1, 1, 2, 517), (5, 2, 4, 301, 2)])
def test_against_oracle(native, b, ci, cs, h, w):
    rng = np.random.default_rng(b * 1000 + h)
    illu = rng.random((b, ci, h, w), dtype=np.float32)
    img = rng.random((b, cs, h, w), dtype=np.float32) * np.float32(0.5)
    loss, _lh, _lv, grad = O.edge_smooth_loss(illu, img, 10.0, 1.0)
    loss3, g = native.edge_smooth_loss(torch.from_numpy(illu).cuda(), torch.from_numpy(img).cuda(), 10.0, 1.0)
    check(loss3, g, loss, grad)
    # without the gradient: same value, no gradient buffer
    loss3b, none = native.edge_smooth_loss(torch.from_numpy(illu).cuda(), torch.from_numpy(img).cuda(), 10.0, 1.0, want_grad=False)
    assert none is None and torch.equal(loss3b, loss3)


def test_autograd_drop_in_matches_stock_formulation(native):
    from retinex_image_enhancement_b200.losses.loss import EdgeAwareSmoothnessLoss
    g = torch.Generator(device="cuda").manual_seed(5)
    illu = torch.rand((4, 1, 128, 160), device="cuda", generator=g)
    img = torch.rand((4, 3, 128, 160), device="cuda", generator=g) * 0.4
    mod = EdgeAwareSmoothnessLoss(lambda_val=10.0, alpha=1.0)
    a = illu.clone().requires_grad_(True)
    (mod(a, img) * 0.37).backward()                       # an upstream factor, like the dynamic weight of loss.py:724
    b = illu.clone().requires_grad_(True)
    (mod._stock(b, img) * 0.37).backward()                # the reference's own sequence of torch ops, autograd
    assert abs(float(mod(illu, img)) - float(mod._stock(illu, img))) <= 5e-6 * float(mod._stock(illu, img))
    assert (a.grad - b.grad).abs().max() <= 2e-6 * b.grad.abs().max()
    # deterministic: two calls, identical bits
    l1, g1 = native.edge_smooth_loss(illu, img)
    l2, g2 = native.edge_smooth_loss(illu, img)
    assert torch.equal(l1, l2) and torch.equal(g1, g2)
    # no gradient requested -> plain value; img_low with a gradient -> stock path reaches it
    assert not mod(illu, img).requires_grad
    s = img.clone().requires_grad_(True)
    mod(illu, s).backward()
    # (the reference's formulation itself yields NaN there wherever the Sobel response is exactly zero: d sqrt(0))
    assert s.grad is not None and torch.nan_to_num(s.grad).abs().sum() > 0


def test_argument_errors(native):
    with pytest.raises(RuntimeError):
        native.edge_smooth_loss(torch.zeros((1, 1, 8, 8)), torch.zeros((1, 3, 8, 8)))                       # host tensors
    with pytest.raises(ValueError):
        native.edge_smooth_loss(torch.zeros((1, 1, 8, 8), device="cuda"), torch.zeros((1, 3, 8, 9), device="cuda"))
    with pytest.raises(native.UprError):
        native.edge_smooth_loss(torch.zeros((1, 1, 1, 8), device="cuda"), torch.zeros((1, 3, 1, 8), device="cuda"))   # h < 2


def test_accelerate_reference_total_loss_stand_in(native):
    """accelerate_reference_total_loss on a stand-in that looks up the same names as losses/loss.py:673 and :707-717."""
    import sys
    import types
    from retinex_image_enhancement_b200.losses import loss as L

    mod = types.ModuleType("fake_reference_losses")
    mod.calculate_texture_complexity = lambda img, method="tv": (_ for _ in ()).throw(AssertionError("stock path used"))

    class Smooth(torch.nn.Module):
        lambda_val, alpha = 6.0, 0.5

        def forward(self, illu, img):
            raise AssertionError("stock smoothness used")

    def forward(self, img_low, illu_map):
        loss_smooth = self.smoothness_loss(illu_map, img_low)
        c = mod.calculate_texture_complexity(img_low, method=self.texture_method)      # module-global lookup, like :707
        w = torch.clamp(self.weight_smooth * (1.0 - torch.mean(c) * 0.8), 0.1, 5.0)
        return w * loss_smooth, w

    Total = type("TotalLoss", (torch.nn.Module,), {"forward": forward, "__module__": mod.__name__})
    sys.modules[mod.__name__] = mod
    try:
        t = Total()
        t.smoothness_loss, t.texture_method, t.weight_smooth = Smooth(), "tv", 1.0
        t = L.accelerate_reference_total_loss(t)
        rng = np.random.default_rng(8)
        img = rng.random((4, 3, 64, 96), dtype=np.float32)
        illu = rng.random((4, 1, 64, 96), dtype=np.float32)
        it = torch.from_numpy(illu).cuda().requires_grad_(True)
        total, w = t(torch.from_numpy(img).cuda(), it)
        total.backward()
        want_w = O.dynamic_smooth_weight(O.texture_tv(img), 1.0)
        want_loss, _h, _v, want_grad = O.edge_smooth_loss(illu, img, 6.0, 0.5)
        assert abs(float(w) - want_w) <= 1e-6
        assert abs(float(total) - want_w * float(want_loss)) <= 5e-6 * abs(want_w * float(want_loss))
        assert np.abs(it.grad.cpu().numpy() - want_w * want_grad).max() <= 2e-6 * np.abs(want_w * want_grad).max()
    finally:
        del sys.modules[mod.__name__]
