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


# ---- exposure / colour / spatial-consistency losses of the enhanced image (one fused evaluation) ---------------------
def _enh_cases():
    spec = importlib.util.spec_from_file_location("make_golden_smooth", os.path.join(GOLDEN_DIR, "make_golden_smooth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.enh_cases()


def _unit(k):
    u = torch.zeros(3, device="cuda")
    u[k] = 1.0
    return u


def test_enhanced_losses_golden_reference_vectors(native):
    """Loss values and per-term gradients against AdaptiveExposureLoss / ColorLoss / SpatialConsistencyLoss of the unmodified
    reference (colour: 2e-4 relative, limited by the reference's own fp32 channel means; the others 1e-5)."""
    gold = np.load(os.path.join(GOLDEN_DIR, "smooth_loss.npz"))
    for name, enh, low in _enh_cases():
        e, l = torch.from_numpy(enh).cuda(), torch.from_numpy(low).cuda()
        losses, saved = native.enhanced_image_losses(e, l)
        for k, tag in enumerate(("exp", "col", "spa")):
            rtol = 2e-4 if tag == "col" else 1e-5
            want, g = float(gold[f"{name}_{tag}_loss"]), gold[f"{name}_{tag}_grad"]
            assert abs(float(losses[k]) - want) <= rtol * abs(want), (name, tag)
            got = native.enhanced_image_losses_grad(e, l, saved, _unit(k)).cpu().numpy()
            assert np.abs(got - g).max() <= rtol * np.abs(g).max() + 1e-12, (name, tag)


@pytest.mark.parametrize("b,h,w", [(8, 256, 256), (2, 640, 640), (3, 97, 131), (1, 16, 2048)])
def test_enhanced_losses_against_oracle(native, b, h, w):
    rng = np.random.default_rng(b * 7 + h)
    enh = rng.random((b, 3, h, w), dtype=np.float32)
    low = rng.random((b, 3, h, w), dtype=np.float32) * np.float32(0.4)
    (l_exp, l_col, l_spa), grads = O.enhanced_image_losses(enh, low)
    e, l = torch.from_numpy(enh).cuda(), torch.from_numpy(low).cuda()
    losses, saved = native.enhanced_image_losses(e, l)
    for k, want in enumerate((l_exp, l_col, l_spa)):
        assert abs(float(losses[k]) - float(want)) <= 2e-6 * abs(float(want)) + 1e-12
        got = native.enhanced_image_losses_grad(e, l, saved, _unit(k)).cpu().numpy()
        assert np.abs(got - grads[k]).max() <= 2e-6 * np.abs(grads[k]).max() + 1e-12
    # a weighted combination in one pass
    up = torch.tensor([0.7, 2.5, 1.3], device="cuda")
    got = native.enhanced_image_losses_grad(e, l, saved, up).cpu().numpy()
    want = np.float32(0.7) * grads[0] + np.float32(2.5) * grads[1] + np.float32(1.3) * grads[2]
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


def test_enhanced_losses_autograd_and_total_loss_views(native):
    """EnhancedImageLosses: the three drop-in modules share one evaluation; autograd through a weighted sum equals the stock
    torch formulation of the three losses."""
    import torch.nn.functional as F
    from retinex_image_enhancement_b200.losses.loss import EnhancedImageLosses
    g = torch.Generator(device="cuda").manual_seed(9)
    enh = torch.rand((4, 3, 96, 128), device="cuda", generator=g)
    low = torch.rand((4, 3, 96, 128), device="cuda", generator=g) * 0.3

    def stock(e):
        gm = torch.mean(torch.mean(low, dim=1, keepdim=True))
        target = 0.6 + (0.8 - 0.6) * (1 - gm)
        l_exp = torch.mean(torch.abs(F.avg_pool2d(torch.mean(e, dim=1, keepdim=True), 16, 16) - target))
        mr, mg, mb = (torch.mean(e[:, c]) for c in range(3))
        l_col = (mr - mg) ** 2 + (mr - mb) ** 2 + (mg - mb) ** 2
        dh = (e[..., :-1] - e[..., 1:]) - (low[..., :-1] - low[..., 1:])
        dv = (e[..., :-1, :] - e[..., 1:, :]) - (low[..., :-1, :] - low[..., 1:, :])
        return l_exp, l_col, torch.mean(dh ** 2) + torch.mean(dv ** 2)

    fused = EnhancedImageLosses()
    exposure, color, spatial = fused.exposure(), fused.color(), fused.spatial()
    a = enh.clone().requires_grad_(True)
    le, lc = exposure(a, low), color(a)                              # the order TotalLoss.forward uses (loss.py:672-675)
    val = fused._val
    assert val is not None and le is val[0] and lc is val[1]
    ls = spatial(a, low)
    assert ls is val[2]                                               # one evaluation served all three terms ...
    assert fused._val is None and fused._enh is None                  # ... and was then released (no pinned graph / tensors)
    (10.0 * le + 5.0 * lc + 1.0 * ls).backward()
    b = enh.clone().requires_grad_(True)
    se, sc, ss = stock(b)
    (10.0 * se + 5.0 * sc + 1.0 * ss).backward()
    for got, want, rtol in ((le, se, 1e-5), (lc, sc, 2e-4), (ls, ss, 1e-5)):
        assert abs(float(got) - float(want)) <= rtol * abs(float(want))
    assert (a.grad - b.grad).abs().max() <= 2e-5 * b.grad.abs().max()
    # a new tensor or an in-place update triggers a new evaluation; ColorLoss alone after that pairs with the same input image
    a2 = enh.clone().requires_grad_(True)
    exposure(a2, low)
    second = fused._val
    assert second is not val and color(a2) is second[1]
    with torch.no_grad():
        low.mul_(0.5)
    assert spatial(a2, low) is not second[2]
    with torch.no_grad():                                             # an evaluation without autograd is never reused with it
        e_ng = exposure(a2, low)
    assert exposure(a2, low) is not e_ng and exposure(a2, low).requires_grad
    fused.clear()
    assert fused._val is None and fused._enh is None
    with pytest.raises(ValueError):
        native.enhanced_image_losses(torch.zeros((1, 1, 32, 32), device="cuda"), torch.zeros((1, 1, 32, 32), device="cuda"))
    with pytest.raises(native.UprError):
        native.enhanced_image_losses(torch.zeros((1, 3, 8, 32), device="cuda"), torch.zeros((1, 3, 8, 32), device="cuda"))


def test_loss_terms_under_autocast_with_half_inputs(native):
    """trainers/train.py:72-77 evaluates the criterion under torch.autocast: the CNN's outputs arrive as fp16.  The autograd
    functions cast to fp32 on entry (custom_fwd), the gradients come back in the inputs' dtype."""
    from retinex_image_enhancement_b200.losses.loss import EdgeAwareSmoothnessLoss, EnhancedImageLosses
    g = torch.Generator(device="cuda").manual_seed(3)
    low = torch.rand((2, 3, 64, 64), device="cuda", generator=g)
    enh16 = torch.rand((2, 3, 64, 64), device="cuda", generator=g).half().requires_grad_(True)
    illu16 = torch.rand((2, 1, 64, 64), device="cuda", generator=g).half().requires_grad_(True)
    fused = EnhancedImageLosses()
    with torch.autocast("cuda"):
        total = EdgeAwareSmoothnessLoss()(illu16, low) + fused.exposure()(enh16, low) + fused.color()(enh16) + fused.spatial()(enh16, low)
    total.backward()
    assert total.dtype == torch.float32 and enh16.grad.dtype == torch.float16 and illu16.grad.dtype == torch.float16
    want, _h, _v, gi = O.edge_smooth_loss(illu16.detach().float().cpu().numpy(), low.cpu().numpy())
    (le, lc, ls), ge = O.enhanced_image_losses(enh16.detach().float().cpu().numpy(), low.cpu().numpy())
    assert abs(float(total) - float(want + le + lc + ls)) <= 1e-5 * float(want + le + lc + ls)
    assert np.abs(illu16.grad.float().cpu().numpy() - gi).max() <= 2e-3 * np.abs(gi).max()          # fp16 rounding of the result
    assert np.abs(enh16.grad.float().cpu().numpy() - (ge[0] + ge[1] + ge[2])).max() <= 2e-3 * np.abs(ge[0] + ge[1] + ge[2]).max()
