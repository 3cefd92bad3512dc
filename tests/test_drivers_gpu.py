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


class MapsStub(torch.nn.Module):
    """A model with forward_maps(): pointwise, hence bit-identical whatever the batch composition (cuDNN convolutions are not)."""

    def forward_maps(self, x):
        return x.mean(dim=1, keepdim=True) * 0.5 + 0.25, torch.sqrt(x)

    def forward(self, x):
        from retinex_image_enhancement_b200.models.model import retinex_recombine
        illu, e = self.forward_maps(x)
        reflectance, enhanced = retinex_recombine(x.contiguous(), illu.contiguous(), e.contiguous())
        return enhanced, reflectance, illu


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
    # predict mode refuses to run without a checkpoint file (main.py:152-157) instead of writing the output of random weights
    rc = cli.main(["--mode", "predict", "--input_path", str(tmp_path / "p.png"), "--output_dir", str(tmp_path / "nockpt"),
                   "--checkpoint", str(tmp_path / "missing.pth")])
    assert rc == 1 and not os.path.exists(tmp_path / "nockpt" / "p_enhanced.png")
    # ... and loads the trainer's checkpoint format (trainers/train.py:134-162) into the same module tree
    from retinex_image_enhancement_b200.models.model import UP_Retinex
    torch.manual_seed(9)
    trained = UP_Retinex(use_preact=False, use_aspp=False)
    torch.save({"epoch": 7, "model_state_dict": trained.state_dict(), "optimizer_state_dict": {}}, tmp_path / "best_model.pth")
    rc = cli.main(["--mode", "predict", "--input_path", str(tmp_path / "p.png"), "--output_dir", str(tmp_path / "ckpt"),
                   "--checkpoint", str(tmp_path / "best_model.pth")])
    assert rc == 0 and os.path.exists(tmp_path / "ckpt" / "p_enhanced.png")
    x = torch.from_numpy((_read_png(tmp_path / "p.png").astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None]).cuda()
    with torch.no_grad():
        want = trained.cuda().eval()(x)[0]
    assert np.array_equal(_read_png(tmp_path / "ckpt" / "p_enhanced.png"),
                          (want.clamp(0, 1) * 255).to(torch.uint8)[0].permute(1, 2, 0).cpu().numpy())
    with pytest.raises(RuntimeError, match="use_preact"):                 # architecture flags that do not match the checkpoint
        cli.main(["--mode", "predict", "--input_path", str(tmp_path / "p.png"), "--output_dir", str(tmp_path / "ckpt2"),
                  "--checkpoint", str(tmp_path / "best_model.pth"), "--use_preact"])


def test_batch_driver_pipeline(cuda, tmp_path):
    """enhance_batch_images: decode prefetch -> pinned staging -> one device batch per shape group -> u8 results -> PNG pool.
    Inputs that share a stem are written one after the other (the last one in list order wins, like the reference's loop);
    --max_size letterboxes on the device; the fused Retinex + CLAHE + u8 kernel serves models with forward_maps()."""
    from retinex_image_enhancement_b200.enhancers.simple_enhance import (enhance_batch_images, enhance_frames_u8,
                                                                         enhance_single_image, load_image)
    from retinex_image_enhancement_b200.models.model import UP_Retinex
    from PIL import Image
    src = tmp_path / "in"
    src.mkdir()
    for i in range(7):
        _write_png(str(src / f"f{i}.png"), 128, 192, 40 + i)
    dup = np.random.default_rng(50).integers(0, 200, (128, 192, 3), dtype=np.uint8)
    Image.fromarray(dup).save(src / "f3.bmp")                 # same stem as f3.png; sorted order: f3.bmp, f3.png -> the png wins
    _write_png(str(src / "wide.png"), 160, 640, 60)
    model = MapsStub().to(cuda).eval()
    enhance_batch_images(str(src), str(tmp_path / "out"), cuda, model=model, batch_size=3)
    for name in [f"f{i}" for i in range(7)] + ["wide"]:
        low, _ = load_image(str(src / f"{name}.png"), device=cuda)
        enh8, illu8 = enhance_frames_u8(model, low)
        assert np.array_equal(_read_png(tmp_path / "out" / f"{name}_enhanced.png"), enh8[0].cpu().numpy()), name
        assert np.array_equal(_read_png(tmp_path / "out" / f"{name}_illumination.png")[:, :, :1], illu8[0].cpu().numpy()), name
        cmp_img = _read_png(tmp_path / "out" / f"{name}_comparison.png")
        assert np.array_equal(cmp_img[:, : cmp_img.shape[1] // 2], _read_png(src / f"{name}.png"))
    # the batch driver's bytes are the single-image driver's bytes
    enhance_single_image(model, str(src / "f5.png"), str(tmp_path / "single"), cuda)
    assert np.array_equal(_read_png(tmp_path / "single" / "f5_enhanced.png"), _read_png(tmp_path / "out" / "f5_enhanced.png"))
    # --max_size: letterbox (down-scale + 114 border to a multiple of 32) on the device, same bytes as the single-image path
    enhance_batch_images(str(src), str(tmp_path / "lb"), cuda, max_size=96, model=model, batch_size=4)
    enhance_single_image(model, str(src / "wide.png"), str(tmp_path / "lb1"), cuda, max_size=96)
    assert np.array_equal(_read_png(tmp_path / "lb" / "wide_enhanced.png"), _read_png(tmp_path / "lb1" / "wide_enhanced.png"))
    assert np.array_equal(_read_png(tmp_path / "lb" / "wide_comparison.png"), _read_png(tmp_path / "lb1" / "wide_comparison.png"))
    # the real CNN (reference module tree, random weights) through the same pipeline, and the other two enhancers
    torch.manual_seed(3)
    cnn = UP_Retinex(use_preact=False, use_aspp=False).to(cuda).eval()
    enhance_batch_images(str(src), str(tmp_path / "cnn"), cuda, model=cnn, batch_size=4)
    enhance_batch_images(str(src), str(tmp_path / "ca"), cuda, model=model, batch_size=4, enable_content_aware=True)
    enhance_batch_images(str(src), str(tmp_path / "ms"), cuda, model=model, batch_size=4, enable_multi_scale=True)
    for d in ("cnn", "ca", "ms"):
        assert _read_png(tmp_path / d / "wide_enhanced.png").shape == (160, 640, 3) and os.path.exists(tmp_path / d / "f6_comparison.png")
    enhance_single_image(model, str(src / "f2.png"), str(tmp_path / "ca1"), cuda, enable_content_aware=True)
    assert np.array_equal(_read_png(tmp_path / "ca1" / "f2_enhanced.png"), _read_png(tmp_path / "ca" / "f2_enhanced.png"))


def test_predict_batch_pipeline(cuda, tmp_path):
    """predict_batch runs on the same decode -> device batch -> PNG pipeline: same bytes as predict_single_image, file by file."""
    from retinex_image_enhancement_b200.predictors.predict import predict_batch, predict_single_image
    src = tmp_path / "in"
    src.mkdir()
    for i in range(5):
        _write_png(str(src / f"p{i}.png"), 96, 160, 70 + i)
    _write_png(str(src / "q.png"), 64, 64, 80)
    model = MapsStub().to(cuda).eval()
    predict_batch(model, str(src), str(tmp_path / "batch"), cuda, batch_size=2)
    for name in [f"p{i}" for i in range(5)] + ["q"]:
        predict_single_image(model, str(src / f"{name}.png"), str(tmp_path / "single"), cuda)
        for suffix in ("enhanced", "illumination", "comparison"):
            assert np.array_equal(_read_png(tmp_path / "batch" / f"{name}_{suffix}.png"), _read_png(tmp_path / "single" / f"{name}_{suffix}.png")), (name, suffix)
    predict_batch(model, str(src), str(tmp_path / "nocmp"), cuda, save_comparison=False)
    assert os.path.exists(tmp_path / "nocmp" / "q_enhanced.png") and not os.path.exists(tmp_path / "nocmp" / "q_comparison.png")


def test_adaptive_parameters_are_lazy(cuda):
    from retinex_image_enhancement_b200.enhancers.adaptive_params import AdaptiveParameterAdjuster
    adj = AdaptiveParameterAdjuster()
    assert adj.last_parameters() is None
    x = torch.from_numpy(O.kat_input(2, 400, 600, "dark"))
    adj.apply_adaptive_enhancement(StubModel(), x, cuda)
    assert adj.last_parameters() == O.adjust_parameters(x.numpy()) == adj.adjust_parameters(x)


def test_ranks_bind_to_their_own_gpu(cuda, monkeypatch):
    from retinex_image_enhancement_b200.enhancers import simple_enhance as S
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    assert S.resolve_device("cuda") == f"cuda:{torch.cuda.current_device()}" and S.resolve_device("cuda:0") == "cuda:0"
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("LOCAL_RANK", str(torch.cuda.device_count()))
    with pytest.raises(RuntimeError, match="LOCAL_RANK"):
        S.bind_rank_to_gpu()
    if torch.cuda.device_count() >= 2:
        before = torch.cuda.current_device()
        monkeypatch.setenv("LOCAL_RANK", "1")
        try:
            assert S.resolve_device("cuda") == "cuda:1" and torch.cuda.current_device() == 1
            assert torch.zeros(1, device=S.resolve_device(None)).device.index == 1
        finally:
            torch.cuda.set_device(before)


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
