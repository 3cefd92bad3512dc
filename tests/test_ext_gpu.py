"""GPU tests of the EXTENSION ops (Gaussian blur / pyramid, log-domain SSR/MSR, gamma).  The reference does not
contain these operations, so the oracle is OpenCV / NumPy directly (cv2 4.13 in this image).  fp32 filters accumulate in a
different order than OpenCV's SIMD code: tolerance 2e-6 relative to the data range; the log-domain maps 2e-5 absolute."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def ext():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from retinex_image_enhancement_b200 import extensions, native
    assert native.lib().upr_device_check() == 0
    return extensions


def _img(seed, n, c, h, w):
    return np.random.default_rng(seed).random((n, c, h, w), dtype=np.float32)


def _cv_blur(x, k, sigma):
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        for j in range(x.shape[1]):
            out[i, j] = cv2.GaussianBlur(x[i, j], (k, k), sigma, borderType=cv2.BORDER_REFLECT_101)
    return out


@pytest.mark.parametrize("shape", [(1, 3, 1080, 1920), (2, 3, 200, 320), (1, 1, 64, 128), (1, 3, 97, 131), (1, 2, 40, 36), (1, 1, 5, 9)])
@pytest.mark.parametrize("k,sigma", [(3, 0.0), (5, 0.0), (7, 0.0), (15, 0.0), (15, 2.6), (31, 5.0), (9, 1.2)])
def test_gaussian_blur_vs_cv2(ext, shape, k, sigma):
    """TMA path (rows 16-byte aligned, >= 32 px), plain path (odd widths, tiny images), every border, multi-plane batches."""
    x = _img(k * 100 + shape[2], *shape)
    got = ext.gaussian_blur(torch.from_numpy(x).cuda(), k, sigma).cpu().numpy()
    np.testing.assert_allclose(got, _cv_blur(x, k, sigma), rtol=0, atol=2e-6)


@pytest.mark.parametrize("shape", [(1, 3, 256, 384), (1, 1, 101, 67)])
def test_msr_vs_cv2_numpy(ext, shape):
    x = _img(9, *shape) * 0.9 + 0.02
    ks, sg, wt, eps = (7, 15, 31), (0.0, 3.0, 5.0), (0.5, 0.3, 0.2), 1e-3
    ref = np.zeros_like(x)
    for k, s, w_ in zip(ks, sg, wt):
        ref += np.float32(w_) * (np.log(x + np.float32(eps)) - np.log(_cv_blur(x, k, s) + np.float32(eps)))
    got = ext.multi_scale_retinex(torch.from_numpy(x).cuda(), ks, sg, wt, eps).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)
    ssr = ext.single_scale_retinex(torch.from_numpy(x).cuda(), 15, 0.0, eps).cpu().numpy()
    np.testing.assert_allclose(ssr, np.log(x + np.float32(eps)) - np.log(_cv_blur(x, 15, 0.0) + np.float32(eps)), rtol=0, atol=2e-5)


@pytest.mark.parametrize("shape", [(1, 3, 1080, 1920), (2, 2, 33, 47), (1, 1, 2, 2), (1, 1, 1, 7)])
def test_pyramid_vs_cv2(ext, shape):
    x = _img(11, *shape)
    pyr = ext.gaussian_pyramid(torch.from_numpy(x).cuda(), 3)
    ref = x
    for lvl in range(1, 3):
        ref = np.stack([np.stack([cv2.pyrDown(ref[i, j]) for j in range(ref.shape[1])]) for i in range(ref.shape[0])])
        assert tuple(pyr[lvl].shape) == ref.shape
        np.testing.assert_allclose(pyr[lvl].cpu().numpy(), ref, rtol=0, atol=2e-6)


def test_gamma_vs_numpy(ext):
    x = (_img(12, 1, 3, 120, 200) * 1.4 - 0.2).astype(np.float32)
    for g in (0.45, 1.0, 2.2):
        got = ext.gamma_correct(torch.from_numpy(x).cuda(), g).cpu().numpy()
        np.testing.assert_allclose(got, np.power(np.clip(x, 0, 1), np.float32(g)), rtol=2e-6, atol=1e-7)


def test_extension_errors(ext):
    x = torch.zeros(1, 3, 64, 64, device="cuda")
    with pytest.raises(RuntimeError):
        ext.gaussian_blur(x, 4)          # even kernel
    with pytest.raises(RuntimeError):
        ext.gaussian_blur(x, 33)         # radius > 15
    with pytest.raises(RuntimeError):
        ext.gaussian_blur(torch.zeros(1, 3, 8, 8), 3)   # CPU tensor: no CPU path


def test_tma_and_plain_tile_fill_agree(ext):
    """The TMA-staged tile (zero fill + reflect patch) and the plain reflect-indexed fill must give bit-identical results.  The
    plain fill serves inputs TMA cannot describe -- here the same frames at an address that is not 16-byte aligned."""
    x = torch.from_numpy(_img(21, 2, 3, 270, 484)).cuda()
    shifted = torch.empty(x.numel() + 1, device="cuda")[1:].view(x.shape)
    shifted.copy_(x)
    assert shifted.data_ptr() % 16 != 0 and x.data_ptr() % 16 == 0
    for k in (5, 31):
        assert torch.equal(ext.gaussian_blur(x, k, 0.0), ext.gaussian_blur(shifted, k, 0.0))
    y, ys = x * 0.9 + 0.05, shifted * 0.9 + 0.05
    ys2 = torch.empty(ys.numel() + 1, device="cuda")[1:].view(ys.shape)
    ys2.copy_(ys)
    assert torch.equal(ext.multi_scale_retinex(y, (7, 15, 31)), ext.multi_scale_retinex(ys2, (7, 15, 31)))
