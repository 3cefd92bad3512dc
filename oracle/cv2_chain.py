"""The reference's CPU call sequence for the CLAHE-in-Lab op, restated on the same third-party library.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's ``cpu_baseline`` leg and by
``bench.py --impl reference``; never imported by the product package.

The reference (/root/reference/enhancers/adaptive_params.py:121-169) is pure Python; all arithmetic is done
by NumPy and by OpenCV (``opencv-python-headless``, un-vendored, unpinned by requirements.txt; container pin
4.13.0.92).  /root/reference does not exist on the GPU box, so the CPU arm times this restatement of the same
library calls in the same order -- per frame: CHW->HWC transpose (:139), ``*255`` + uint8 cast (:142), RGB2BGR
(:142), BGR2LAB (:145), split (:146), createCLAHE(2.0,(8,8)).apply (:149-152), merge (:155), LAB2BGR (:158),
BGR2RGB (:161), float32 ``/255`` (:164), permute back to CHW (:167).  tests/test_oracle_pin.py checks that this
chain and the C oracle agree bit for bit, and the golden fixtures tie both to the unmodified reference.
"""
from __future__ import annotations

import numpy as np


def available() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def clahe_lab_frame(chw: np.ndarray, clip_limit: float = 2.0, tiles=(8, 8)) -> np.ndarray:
    """One frame [3,H,W] f32 -> [3,H,W] f32 (a permuted view of the HWC result, like the reference)."""
    import cv2
    hwc = np.transpose(chw, (1, 2, 0))
    with np.errstate(invalid="ignore"):
        bgr = cv2.cvtColor((hwc * 255).astype(np.uint8), cv2.COLOR_RGB2BGR)
    l, a, b = cv2.split(cv2.cvtColor(bgr, cv2.COLOR_BGR2LAB))
    l2 = cv2.createCLAHE(clipLimit=clip_limit, tileGridSize=tuple(tiles)).apply(l)
    rgb = cv2.cvtColor(cv2.cvtColor(cv2.merge((l2, a, b)), cv2.COLOR_LAB2BGR), cv2.COLOR_BGR2RGB)
    return np.transpose(rgb.astype(np.float32) / 255.0, (2, 0, 1))


def clahe_lab_frame_u8(hwc_rgb: np.ndarray, clip_limit: float = 2.0, tiles=(8, 8)) -> np.ndarray:
    """The OpenCV part of the chain alone (adaptive_params.py:142-161) on a packed u8 RGB frame [H,W,3] -> [H,W,3] u8: the CPU
    counterpart of upr_clahe_lab_u8 (the float casts on either side of it are exact on such frames)."""
    import cv2
    bgr = cv2.cvtColor(hwc_rgb, cv2.COLOR_RGB2BGR)
    l, a, b = cv2.split(cv2.cvtColor(bgr, cv2.COLOR_BGR2LAB))
    l2 = cv2.createCLAHE(clipLimit=clip_limit, tileGridSize=tuple(tiles)).apply(l)
    return cv2.cvtColor(cv2.cvtColor(cv2.merge((l2, a, b)), cv2.COLOR_LAB2BGR), cv2.COLOR_BGR2RGB)


def clahe_lab_batch_u8(frames: np.ndarray, workers: int = 1) -> list:
    """[N,H,W,3] u8 -> list of N u8 results (thread pool like clahe_lab_batch)."""
    if workers <= 1:
        return [clahe_lab_frame_u8(f) for f in frames]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(clahe_lab_frame_u8, frames))


def clahe_lab_batch(frames: np.ndarray, workers: int = 1) -> list:
    """[N,3,H,W] -> list of N results.  workers > 1 runs frames on a thread pool (NumPy and OpenCV release
    the GIL), which is how the CPU arm uses every host core; the reference itself loops serially
    (enhancers/simple_enhance.py:237-243) and relies on OpenCV's internal threading."""
    if workers <= 1:
        return [clahe_lab_frame(f) for f in frames]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(clahe_lab_frame, frames))


# ---------------------------------------------------------------------------------------------------
# letterbox: the reference's utils/letterbox.py:9-102 call sequence, and the pinned fixed-point recipe
# ---------------------------------------------------------------------------------------------------
def letterbox_ref(chw: np.ndarray, new_shape, color=(114, 114, 114), auto=True, scale_fill=False, scaleup=True):
    """utils/letterbox.py letterbox_tensor on a [C,H,W] f32 array, OpenCV doing the pixels: -> ([C,H',W'] f32, ratio, (dw, dh))."""
    import cv2
    img = (np.transpose(chw, (1, 2, 0)) * 255).astype(np.uint8)
    shape = img.shape[:2]
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    ratio = r, r
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, 32), np.mod(dh, 32)
    elif scale_fill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
        ratio = new_shape[1] / shape[1], new_shape[0] / shape[0]
    dw /= 2
    dh /= 2
    if shape[::-1] != new_unpad:
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    img = cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)
    if img.ndim == 2:
        img = img[:, :, None]
    return np.ascontiguousarray(np.transpose(img.astype(np.float32) / 255.0, (2, 0, 1))), ratio, (dw, dh)


def resize_linear_u8_fixed(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src u8 [H,W,C], (dw, dh), INTER_LINEAR) restated (8-bit fixed point, 11-bit coefficients); pinned against
    the binary for down-scaling (tests/test_oracle_pin.py) -- the specification the GPU kernel implements."""
    sh, sw = src.shape[:2]

    def coeffs(dn, sn, scale):
        idx = np.empty(dn, np.int64)
        a0 = np.empty(dn, np.int64)
        a1 = np.empty(dn, np.int64)
        for d in range(dn):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - np.float32(s))
            if s < 0:
                f, s = np.float32(0), 0
            if s >= sn - 1:
                f, s = np.float32(0), sn - 1
            idx[d] = s
            a0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
            a1[d] = int(np.rint(np.float32(f * np.float32(2048))))
        return idx, a0, a1

    xi, xa0, xa1 = coeffs(dw, sw, sw / dw)
    yi, ya0, ya1 = coeffs(dh, sh, sh / dh)
    s64 = src.reshape(sh, sw, -1).astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    rows = s64[:, xi, :] * xa0[None, :, None] + s64[:, x1, :] * xa1[None, :, None]
    y1 = np.minimum(yi + 1, sh - 1)
    out = (((ya0[:, None, None] * (rows[yi] >> 4)) >> 16) + ((ya1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dh, dw) + src.shape[2:])
