"""GPU checks of the reference-facing drivers: enhance_single_image / enhance_batch_images / predict_single_image and
AdaptiveParameterAdjuster.apply_adaptive_enhancement, with a stub CNN so that results are reproducible (the
reference itself runs a randomly initialised network, SURVEY finding 5)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return "cuda"


class StubModel(torch.nn.Module):
    """enhanced = sqrt(x) (a fixed brightening), illu = mean over channels."""

    def forward(self, x):
        return torch.sqrt(x), x, x.mean(dim=1, keepdim=True)


def _write_png(path, h, w, seed):
    from PIL import Image
    a = np.random.default_rng(seed).integers(0, 90, (h, w, 3), dtype=np.uint8)
    Image.fromarray(a).save(path)
    return a


def _read_png(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


def test_apply_adaptive_enhancement_matches_oracle(cuda):
    from retinex_image_enhancement_b200.enhancers.adaptive_params import AdaptiveParameterAdjuster
    x = O.kat_input(2, 400, 600, "dark")
    out, illu = AdaptiveParameterAdjuster().apply_adaptive_enhancement(StubModel(), torch.from_numpy(x), cuda)
    assert out.is_cuda and illu.shape == (1, 1, 400, 600)
    ref = O.clahe_lab(np.sqrt(x))
    assert np.array_equal(out.cpu().numpy(), ref)


def test_enhance_drivers_write_the_reference_file_set(cuda, tmp_path):
    from retinex_image_enhancement_b200.enhancers.simple_enhance import enhance_batch_images, enhance_single_image
    src, dst_b, dst_s = tmp_path / "in", tmp_path / "out_batch", tmp_path / "out_single"
    src.mkdir()
    imgs = {name: _write_png(str(src / f"{name}.png"), h, w, i) for i, (name, h, w) in
            enumerate([("a", 120, 160), ("b", 120, 160), ("c", 96, 128), ("d", 120, 160)])}
    (src / "notes.txt").write_text("ignored")
    model = StubModel()
    enhance_batch_images(str(src), str(dst_b), cuda, model=model, batch_size=2)
    for name, a in imgs.items():
        for suffix in ("enhanced", "illumination", "comparison"):
            assert os.path.exists(dst_b / f"{name}_{suffix}.png"), (name, suffix)
        x = (a.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None]
        ref = O.clahe_lab(np.sqrt(x))[0].transpose(1, 2, 0)
        assert np.array_equal(_read_png(dst_b / f"{name}_enhanced.png"), (np.clip(ref, 0, 1) * 255).astype(np.uint8))
        cmp_img = _read_png(dst_b / f"{name}_comparison.png")
        assert cmp_img.shape == (a.shape[0], 2 * a.shape[1], 3) and np.array_equal(cmp_img[:, : a.shape[1]], a)
    # single-image path gives the same bytes as the batched path
    enhance_single_image(model, str(src / "c.png"), str(dst_s), cuda)
    assert np.array_equal(_read_png(dst_s / "c_enhanced.png"), _read_png(dst_b / "c_enhanced.png"))
    # the two flags the reference drops are honoured
    enhance_single_image(model, str(src / "a.png"), str(tmp_path / "ms"), cuda, enable_multi_scale=True)
    enhance_single_image(model, str(src / "a.png"), str(tmp_path / "ca"), cuda, enable_content_aware=True)
    x = (imgs["a"].astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None]
    _, f_ref = O.multiscale_means(x)
    ref_ms = (np.clip(O.scale_clamp(np.sqrt(x), np.float32(f_ref))[0].transpose(1, 2, 0), 0, 1) * 255).astype(np.uint8)
    got_ms = _read_png(tmp_path / "ms" / "a_enhanced.png")
    assert np.abs(got_ms.astype(int) - ref_ms.astype(int)).max() <= 1          # <= 1 LSB on uint8 output
    assert os.path.exists(tmp_path / "ca" / "a_enhanced.png")


def test_predict_and_cli(cuda, tmp_path):
    from retinex_image_enhancement_b200 import cli
    from retinex_image_enhancement_b200.predictors.predict import predict_single_image
    _write_png(str(tmp_path / "p.png"), 64, 96, 5)
    predict_single_image(StubModel(), str(tmp_path / "p.png"), str(tmp_path / "pred"), cuda)
    assert _read_png(tmp_path / "pred" / "p_comparison.png").shape == (64, 3 * 96, 3)
    cli.main(["--mode", "enhance", "--input_path", str(tmp_path / "p.png"), "--output_dir", str(tmp_path / "cli"), "--device", "cuda"])
    assert os.path.exists(tmp_path / "cli" / "p_enhanced.png")           # single-file mode works (reference: TypeError)
    cli.simple_enhance_main(["--input", str(tmp_path / "p.png"), "--output", str(tmp_path / "cli2"), "--content_aware"])
    assert os.path.exists(tmp_path / "cli2" / "p_illumination.png")


def test_model_inference_uses_fused_recombine(cuda):
    from retinex_image_enhancement_b200.models.model import UP_Retinex
    torch.manual_seed(0)
    m = UP_Retinex().to(cuda).eval()
    x = torch.rand(2, 3, 64, 96, device=cuda)
    with torch.no_grad():
        enhanced, reflectance, illu = m(x)
        e = m.enhancement_map(x)
    r_ref = x / (illu + 1e-6)
    assert torch.equal(reflectance, r_ref)
    assert torch.allclose(enhanced, r_ref * e + (1 - r_ref) * e ** 2, rtol=3e-7, atol=1e-30)


def test_apply_adaptive_enhancement_fused_path_matches_unfused(cuda):
    """With a model that exposes forward_maps() the adjuster takes the fused recombination+CLAHE kernel; the result must equal
    model(x)[0] -> apply_clahe_enhancement bit for bit, and the oracle's recombination + CLAHE."""
    from retinex_image_enhancement_b200 import native
    from retinex_image_enhancement_b200.enhancers.adaptive_params import AdaptiveParameterAdjuster
    from retinex_image_enhancement_b200.models.model import UP_Retinex
    torch.manual_seed(5)
    model = UP_Retinex().to(cuda).eval()
    x = torch.from_numpy(O.kat_input(2, 400, 600, "dark"))
    adj = AdaptiveParameterAdjuster()
    out, illu = adj.apply_adaptive_enhancement(model, x, cuda)
    with torch.no_grad():
        enhanced, _r, illu2 = model(x.to(cuda))
    ref = native.clahe_lab(enhanced.contiguous())
    assert torch.equal(out, ref) and torch.equal(illu, illu2)
    with torch.no_grad():
        il, e = model.forward_maps(x.to(cuda))
    _, e_ref = O.retinex_recombine(x.numpy(), il.cpu().numpy(), e.cpu().numpy())
    assert np.array_equal(out.cpu().numpy(), O.clahe_lab(e_ref))
