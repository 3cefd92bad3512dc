"""Out-of-bounds and race checks that do not need compute-sanitizer (the tool is closed on this pool: scripts/dev/sanitize.sh is
what we would run; `compute-sanitizer` answers "closed on this pool" there).

  * guard bands: every output tensor and every workspace is carved out of a larger allocation whose margins carry a canary
    pattern; after the call the margins must be untouched (an out-of-bounds WRITE of any kernel shows up here);
  * inputs sit at the very end of their allocation with a poisoned (NaN) margin on both sides: a kernel that READS past either end
    of an input changes a result that is compared with the oracle;
  * determinism: results are bit-identical over repeated calls while another stream keeps the GPU busy -- the ticket hand-overs,
    byte-counter histograms, ordered-key min/max atomics and work queues must not depend on scheduling.
Shapes are the awkward ones: ragged widths, band / segment / strip edges, single rows."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
CANARY = 0xA5
MARGIN = 1 << 16


@pytest.fixture(scope="module")
def native():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from retinex_image_enhancement_b200 import native as nat
    assert nat.lib().upr_device_check() == 0
    return nat


class Guarded:
    """A tensor of `shape` / `dtype` in the middle of a canary-filled allocation."""

    def __init__(self, shape, dtype=torch.float32):
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 256
        self.raw = torch.full((MARGIN + nbytes + pad + MARGIN,), CANARY, dtype=torch.uint8, device="cuda")
        self.nbytes = nbytes
        self.t = self.raw[MARGIN:MARGIN + nbytes].view(dtype).view(shape)

    def intact(self):
        return bool((self.raw[:MARGIN] == CANARY).all()) and bool((self.raw[MARGIN + self.nbytes:] == CANARY).all())


def poisoned_input(a):
    """`a` on the device with NaN float margins directly before and after it."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    raw = torch.full((4096 + a.size + 4096,), float("nan"), dtype=torch.float32, device="cuda")
    t = raw[4096:4096 + a.size].view(a.shape)
    t.copy_(torch.from_numpy(a))
    return t, raw


SHAPES = [(2, 64, 96), (1, 67, 93), (3, 33, 481), (1, 136, 256), (2, 128, 248), (1, 17, 8), (1, 1080, 1920)]


@pytest.mark.parametrize("n,h,w", SHAPES)
def test_guard_bands_and_poisoned_margins(native, monkeypatch, n, h, w):
    x = np.concatenate([O.kat_input(900 + i, h, w, ("uniform", "dark", "ramp")[i % 3]) for i in range(n)])
    enh = np.random.default_rng(h * w).random((n, 3, h, w), dtype=np.float32) * np.float32(1.2)
    illu = (np.random.default_rng(h + w).random((n, 1, h, w), dtype=np.float32) * 0.9 + 0.05).astype(np.float32)
    xd, xraw = poisoned_input(x)
    ed, eraw = poisoned_input(enh)
    il, iraw = poisoned_input(illu)
    guards = []

    # every workspace of the call under test is a Guarded allocation
    def guarded_ws(nbytes, device):
        g = Guarded((max(nbytes, 1),), torch.uint8)
        g.t.zero_()
        guards.append(g)
        return g.t
    monkeypatch.setattr(native, "workspace", guarded_ws)
    monkeypatch.setattr(native, "zero_workspace", lambda tag, nbytes, device: guarded_ws(nbytes, device))

    def out(shape, dtype=torch.float32):
        g = Guarded(shape, dtype)
        guards.append(g)
        return g.t

    small = h * w <= 400 * 600
    got = native.clahe_lab(xd, out=out((n, 3, h, w)))
    if small:
        for i in range(n):
            assert np.array_equal(got[i:i + 1].cpu().numpy(), O.clahe_lab(x[i:i + 1]))
    native.retinex_clahe(xd, il, ed, out=out((n, 3, h, w)))
    native.retinex_clahe_u8(xd, il, ed, out=out((n, h, w, 3), torch.uint8))
    native.clahe_lab_f32_u8(xd, out=out((n, h, w, 3), torch.uint8))
    x8 = Guarded((n, h, w, 3), torch.uint8)
    x8.t.copy_((xd * 255).to(torch.uint8).permute(0, 2, 3, 1))
    guards.append(x8)
    native.clahe_lab_u8(x8.t, out=out((n, h, w, 3), torch.uint8))
    ca = native.content_aware_apply(xd, ed, out=out((n, 3, h, w)))
    chain, _g = native.content_multiscale_apply(xd, ed, out=out((n, 3, h, w)))
    sal, att = native.saliency(xd), native.attention(xd)
    if h >= 4 and w >= 4:
        m, gain = native.multiscale_stats(xd)
        mg, _ = native.multiscale_stats(xd, force_generic=True)
        native.multiscale_enhance(xd, ed, out=out((n, 3, h, w)))
    native.scale_clamp(ed, torch.ones(n, device="cuda"), out=out((n, 3, h, w)))
    native.quantize_u8(ed, out=out((n, h, w, 3), torch.uint8))
    native.quantize_u8(il, out=out((n, h, w, 1), torch.uint8))
    native.brightness_hist(xd)
    tv = native.texture_complexity(xd, "tv")
    native.texture_complexity(xd, "edge_density")
    native.retinex_recombine(xd, il, ed)
    native.edge_smooth_loss(il, xd)
    if h >= 16 and w >= 16:
        losses, saved = native.enhanced_image_losses(ed, xd)
        native.enhanced_image_losses_grad(ed, xd, saved, torch.ones(3, device="cuda"))
    torch.cuda.synchronize()
    for g in guards:
        assert g.intact(), "a kernel wrote outside its buffer"
    for raw, a in ((xraw, x), (eraw, enh), (iraw, illu)):       # nobody wrote into the inputs' margins either
        assert bool(torch.isnan(raw[:4096]).all()) and bool(torch.isnan(raw[4096 + a.size:]).all())
    # a read past an input's end would have pulled NaN into these
    assert not bool(torch.isnan(ca).any()) and not bool(torch.isnan(chain).any()) and not bool(torch.isnan(sal).any())
    assert not bool(torch.isnan(att).any()) and not bool(torch.isnan(tv).any())
    if small:
        for i in range(n):
            np.testing.assert_allclose(att[i:i + 1].cpu().numpy(), O.attention(x[i:i + 1]), rtol=0, atol=2e-6)
            np.testing.assert_allclose(tv[i:i + 1].cpu().numpy(), O.texture_tv(x[i:i + 1]), rtol=2e-6)
            if h >= 4 and w >= 4:
                np.testing.assert_allclose(m[i].cpu().numpy(), O.multiscale_means(x[i:i + 1])[0], rtol=2e-6)
                np.testing.assert_allclose(mg[i].cpu().numpy(), O.multiscale_means(x[i:i + 1])[0], rtol=2e-6)


def test_results_do_not_depend_on_scheduling(native):
    """20 repetitions of every op with a second stream hammering the GPU: bit-identical results (ticket hand-overs, byte-counter
    histograms with `red.shared`, ordered-key atomics, the persistent map kernel's work queue, the two-stream chunk schedule)."""
    n, h, w = 6, 1080, 1920
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((n, 3, h, w), device="cuda", generator=g) * 0.8
    x[1] = 0.3                                           # constant frame: every pixel of a tile in one histogram bin
    enh = torch.rand((n, 3, h, w), device="cuda", generator=g)
    illu = torch.rand((n, 1, h, w), device="cuda", generator=g) * 0.9 + 0.05
    noise = torch.rand((64, 1024, 1024), device="cuda")
    side = torch.cuda.Stream()

    def run():
        return (native.clahe_lab(x), native.retinex_clahe(x, illu, enh), native.content_aware_apply(x, enh), native.attention(x),
                native.multiscale_stats(x)[0], native.texture_complexity(x, "tv"), native.texture_complexity(x, "edge_density"),
                native.edge_smooth_loss(illu, x)[0], native.brightness_hist(x))

    ref = [t.clone() for t in run()]
    for rep in range(20):
        with torch.cuda.stream(side):
            for _ in range(1 + rep % 3):
                noise.mul_(1.0001)
        for a, b in zip(run(), ref):
            assert torch.equal(a, b), rep
    torch.cuda.synchronize()
