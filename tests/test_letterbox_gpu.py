"""GPU parity of the letterbox kernel against the reference's own recipe evaluated by OpenCV
(oracle/cv2_chain.letterbox_ref == utils/letterbox.py, pinned in tests/test_oracle_pin.py): bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
pytest.importorskip("cv2")
from oracle import cv2_chain  # noqa: E402


@pytest.fixture(scope="module")
def native():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from retinex_image_enhancement_b200 import native as nat
    assert nat.lib().upr_device_check() == 0
    return nat


@pytest.mark.parametrize("h,w,new_shape", [(480, 640, 256), (1000, 1024, 640), (1080, 1920, 512), (123, 457, 96), (300, 517, 640),
                                            (64, 64, 64), (2160, 3840, 1280), (777, 1333, (389, 667))])
def test_letterbox_tensor_cuda_matches_reference_recipe(native, h, w, new_shape):
    from retinex_image_enhancement_b200.utils.letterbox import letterbox_tensor
    x = np.random.default_rng(h + w).random((3, h, w), dtype=np.float32)
    ref, ratio, pad = cv2_chain.letterbox_ref(x, new_shape, auto=True, scaleup=False)
    got, ratio2, pad2 = letterbox_tensor(torch.from_numpy(x).cuda(), new_shape=new_shape, auto=True, scaleup=False)
    assert got.is_cuda and tuple(got.shape) == ref.shape
    assert np.array_equal(got.cpu().numpy(), ref)
    assert ratio == ratio2 and tuple(pad) == tuple(pad2)


def test_letterbox_u8_input_and_out_of_range_floats(native):
    """The decoded-file entry (u8 HWC) equals the f32 entry on k/255 inputs; floats outside [0,1] follow numpy's cast."""
    from retinex_image_enhancement_b200.utils.letterbox import letterbox_geometry
    rng = np.random.default_rng(5)
    u8 = rng.integers(0, 256, (2, 300, 420, 3), dtype=np.uint8)
    f32 = np.ascontiguousarray(np.transpose(u8, (0, 3, 1, 2)).astype(np.float32) / 255.0)
    (rh, rw), (top, bottom, left, right), _, _ = letterbox_geometry(300, 420, 256, auto=True, scaleup=False)
    a = native.letterbox(torch.from_numpy(u8).cuda(), (rh, rw), top, left, (rh + top + bottom, rw + left + right))
    b = native.letterbox(torch.from_numpy(f32).cuda(), (rh, rw), top, left, (rh + top + bottom, rw + left + right))
    assert torch.equal(a, b)
    for i in range(2):
        assert np.array_equal(a[i].cpu().numpy(), cv2_chain.letterbox_ref(f32[i], 256, auto=True, scaleup=False)[0])
    wild = (rng.random((3, 90, 130), dtype=np.float32) * 3 - 1).astype(np.float32)
    wild[0, 3, 4] = np.nan
    with np.errstate(invalid="ignore"):
        ref = cv2_chain.letterbox_ref(wild, 64, auto=True, scaleup=False)[0]
    from retinex_image_enhancement_b200.utils.letterbox import letterbox_tensor
    got = letterbox_tensor(torch.from_numpy(wild).cuda(), new_shape=64, auto=True, scaleup=False)[0]
    assert np.array_equal(got.cpu().numpy(), ref)


def test_letterbox_upscale_refused_and_host_path(native):
    x = torch.rand(1, 3, 40, 50, device="cuda")
    with pytest.raises(RuntimeError):
        native.letterbox(x, (80, 100), 0, 0, (80, 100))
    from retinex_image_enhancement_b200.utils.letterbox import letterbox_tensor
    got, _, _ = letterbox_tensor(x[0], new_shape=128, auto=False, scaleup=True)      # up-scaling: OpenCV on the host
    ref, _, _ = cv2_chain.letterbox_ref(x[0].cpu().numpy(), 128, auto=False, scaleup=True)
    assert np.array_equal(got.cpu().numpy(), ref)


def test_load_image_on_device_equals_host_path(native, tmp_path):
    from PIL import Image
    from retinex_image_enhancement_b200.enhancers.simple_enhance import load_image
    a = np.random.default_rng(8).integers(0, 256, (333, 500, 3), dtype=np.uint8)
    path = str(tmp_path / "img.png")
    Image.fromarray(a).save(path)
    for max_size in (None, 256, 640):
        host, size_h = load_image(path, max_size)
        dev, size_d = load_image(path, max_size, device="cuda")
        assert dev.is_cuda and size_h == size_d == (500, 333)
        assert torch.equal(host, dev.cpu())
