"""Packed u8 boundary of the CLAHE-in-Lab op (upr_clahe_lab_u8, upr_clahe_lab_f32_u8) against the oracle.

The reference (enhancers/adaptive_params.py:121-169) works on f32 tensors; on x = u8 / 255 its quantisation is the identity
and save_image's (y * 255).astype(u8) of its output recovers the u8 result (tests/test_oracle_pin.py pins both casts), so the
u8 entry must equal: oracle(u8 / 255) * 255 truncated.  Bit-exact, incl. Lab intermediate, histograms and LUTs.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from retinex_image_enhancement_b200 import native as nv
    return nv


def u8_frames(seed, n, h, w, kind="uniform"):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    if kind == "dark":
        x = (x.astype(np.float32) * 0.3).astype(np.uint8)
    elif kind == "const":
        x[:] = 77
    elif kind == "ramp":
        xs = (np.arange(w) * 255 // max(w - 1, 1)).astype(np.uint8)
        ys = (np.arange(h) * 255 // max(h - 1, 1)).astype(np.uint8)
        x[..., 0] = xs[None, None, :]
        x[..., 1] = ys[None, :, None]
        x[..., 2] = ((xs[None, :].astype(np.int32) + ys[:, None]) // 2).astype(np.uint8)[None]
    return x


def oracle_u8(x_u8, clip=2.0, tiles=(8, 8)):
    """Reference semantics on a u8 frame: f32 = u8 / 255 -> oracle -> * 255 truncated; also returns the oracle's taps."""
    outs, taps = [], []
    for i in range(x_u8.shape[0]):
        xf = (x_u8[i].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None]
        ref, tp = O.clahe_lab(np.ascontiguousarray(xf), clip, tiles, taps=True)
        outs.append((ref[0] * np.float32(255)).astype(np.uint8).transpose(1, 2, 0))
        taps.append(tp)
    return np.stack(outs), taps


def check_u8(native, x_u8, clip=2.0, tiles=(8, 8)):
    ref, taps = oracle_u8(x_u8, clip, tiles)
    xd = torch.from_numpy(x_u8).cuda()
    out = native.clahe_lab_u8(xd, clip, tiles)
    n, h, w, _ = x_u8.shape
    hist, lut, lab = native.clahe_debug((n, 3, h, w), tiles)
    torch.cuda.synchronize()
    for i in range(n):
        assert np.array_equal(lab[i].cpu().numpy(), taps[i]["lab"]), f"Lab intermediate differs (frame {i})"
        assert np.array_equal(hist[i].cpu().numpy(), taps[i]["hist"]), f"histograms differ (frame {i})"
        assert np.array_equal(lut[i].cpu().numpy(), taps[i]["lut"]), f"LUTs differ (frame {i})"
    assert np.array_equal(out.cpu().numpy(), ref)
    return xd, out


@pytest.mark.parametrize("h,w,kind", [(1080, 1920, "uniform"), (1080, 1920, "dark"), (2160, 3840, "ramp"), (256, 256, "uniform"),
                                      (400, 600, "uniform"), (403, 601, "dark"), (64, 2048, "const"), (72, 96, "ramp")])
def test_u8_against_oracle(native, h, w, kind):
    check_u8(native, u8_frames(10, 1, h, w, kind))


def test_u8_batch_other_grids_and_in_place(native):
    x = np.concatenate([u8_frames(20 + i, 1, 480, 640, k) for i, k in enumerate(["uniform", "dark", "ramp", "const", "uniform"])])
    check_u8(native, x)
    check_u8(native, x[:2], 3.5, (4, 4))
    check_u8(native, x[:2], 0.0, (8, 8))          # clip 0: plain histogram equalisation
    check_u8(native, u8_frames(31, 2, 96, 2048), 2.0, (1, 3))
    xd = torch.from_numpy(x).cuda()
    ref = native.clahe_lab_u8(xd).clone()
    native.clahe_lab_u8(xd, out=xd)               # in place
    assert torch.equal(xd, ref)
    empty = torch.empty((0, 480, 640, 3), dtype=torch.uint8, device="cuda")
    assert native.clahe_lab_u8(empty).shape == (0, 480, 640, 3)


def test_u8_equals_f32_path(native):
    """u8 -> u8 == quantise(f32 op(u8 / 255)) and f32 -> u8 == quantise(f32 op(x)) on a 1080p batch (vector kernels) and a ragged shape."""
    for n, h, w in ((5, 1080, 1920), (2, 403, 601)):
        x8 = torch.from_numpy(u8_frames(40, n, h, w)).cuda()
        xf = (x8.float() / 255.0).permute(0, 3, 1, 2).contiguous()
        yf = native.clahe_lab(xf)
        want = (yf * 255.0).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
        assert torch.equal(native.clahe_lab_u8(x8), want)
        assert torch.equal(native.clahe_lab_f32_u8(xf), want)
        # arbitrary floats (not multiples of 1/255) through the f32 -> u8 entry
        g = torch.Generator(device="cuda").manual_seed(41)
        xr = torch.rand((n, 3, h, w), device="cuda", generator=g)
        want_r = (native.clahe_lab(xr) * 255.0).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
        assert torch.equal(native.clahe_lab_f32_u8(xr), want_r)


def test_u8_host_entry(native):
    """upr_clahe_lab_u8_host (pinned or pageable host frames, chunked pipeline) == the device entry."""
    x = u8_frames(50, 7, 480, 640)
    want = native.clahe_lab_u8(torch.from_numpy(x).cuda()).cpu()
    for chunk in (0, 1, 3):
        assert torch.equal(native.clahe_lab_u8_host(torch.from_numpy(x), frames_per_chunk=chunk), want)
    pinned = torch.from_numpy(x).pin_memory()
    out = torch.empty(x.shape, dtype=torch.uint8, pin_memory=True)
    assert native.clahe_lab_u8_host(pinned, out=out) is out and torch.equal(out, want)
    with pytest.raises(TypeError):
        native.clahe_lab_u8_host(torch.zeros((1, 8, 8, 3)))


def test_u8_argument_errors(native):
    with pytest.raises(TypeError):
        native.clahe_lab_u8(torch.zeros((1, 8, 8, 3), device="cuda"))                       # f32
    with pytest.raises(TypeError):
        native.clahe_lab_u8(torch.zeros((1, 3, 8, 8), dtype=torch.uint8, device="cuda"))    # NCHW
    with pytest.raises(RuntimeError):
        native.clahe_lab_u8(torch.zeros((1, 8, 8, 3), dtype=torch.uint8))                   # host tensor
    with pytest.raises(ValueError):
        native.clahe_lab_f32_u8(torch.zeros((1, 3, 8, 8), device="cuda"), out=torch.zeros((1, 3, 8, 8), dtype=torch.uint8, device="cuda"))
