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


def clahe_lab_batch(frames: np.ndarray, workers: int = 1) -> list:
    """[N,3,H,W] -> list of N results.  workers > 1 runs frames on a thread pool (NumPy and OpenCV release
    the GIL), which is how the CPU arm uses every host core; the reference itself loops serially
    (enhancers/simple_enhance.py:237-243) and relies on OpenCV's internal threading."""
    if workers <= 1:
        return [clahe_lab_frame(f) for f in frames]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(clahe_lab_frame, frames))
