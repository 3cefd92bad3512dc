"""Letterbox pre-processing boundary (reference: utils/letterbox.py:9-102).

With ``new_shape`` equal to the image's own shape -- the enhance default (enhancers/simple_enhance.py:53-58) -- the
reference's letterbox is a lossless uint8 round trip of a k/255 grid, i.e. the identity (SURVEY.md section 2, row
12, verified for all 256 k), so it is skipped.  With ``--max_size`` the YOLO recipe -- ratio =
min(new/old, 1), cv2.INTER_LINEAR resize on uint8, constant 114 border, mod-32 padding -- has its geometry computed
here exactly as the reference does; the pixels are produced by ``upr_letterbox_f32`` / ``upr_letterbox_u8_f32`` when
the tensor lives on a CUDA device and the image is not up-scaled (bit-exact against cv2), and by OpenCV on the host
otherwise (host tensors: the reference's own behaviour; up-scaling: cv2's 8-bit up-scaler is not reproduced).
"""
from __future__ import annotations

import numpy as np
import torch


def letterbox_geometry(h: int, w: int, new_shape, auto=True, scale_fill=False, scaleup=True):
    """The size arithmetic of utils/letterbox.py:26-57 alone: -> ((rh, rw), (top, bottom, left, right), ratio, (dw, dh))."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    new_shape = tuple(int(v) for v in new_shape)
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    ratio = (r, r)
    unpad = (int(round(w * r)), int(round(h * r)))
    dw, dh = new_shape[1] - unpad[0], new_shape[0] - unpad[1]
    if auto:
        dw, dh = dw % 32, dh % 32
    elif scale_fill:
        dw, dh, unpad = 0.0, 0.0, (new_shape[1], new_shape[0])
        ratio = (new_shape[1] / w, new_shape[0] / h)
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return (unpad[1], unpad[0]), (top, bottom, left, right), ratio, (dw, dh)


def letterbox_tensor(img_tensor: torch.Tensor, new_shape=640, color=(114, 114, 114), auto=True, scale_fill=False,
                     scaleup=True):
    """[C,H,W] f32 in [0,1] -> (letterboxed [C,H',W'] f32, (rw, rh), (dw, dh)).  CUDA tensors stay on the device."""
    c, h, w = img_tensor.shape
    (rh, rw), (top, bottom, left, right), ratio, (dw, dh) = letterbox_geometry(h, w, new_shape, auto, scale_fill, scaleup)
    if img_tensor.is_cuda and rw <= w and rh <= h:
        from .. import native
        out = native.letterbox(img_tensor.detach().to(torch.float32).unsqueeze(0), (rh, rw), top, left,
                               (rh + top + bottom, rw + left + right), color)
        return out[0], ratio, (dw, dh)
    if (h, w) == (rh, rw) and top == bottom == left == right == 0:
        # identity: float -> uint8 -> float of a k/255 grid round-trips; for general floats apply the same quantisation
        q = (img_tensor.detach().cpu() * 255).to(torch.uint8).to(torch.float32) / 255.0
        return q, ratio, (dw, dh)
    import cv2
    hwc = (img_tensor.detach().cpu().numpy().transpose(1, 2, 0) * 255).astype(np.uint8)
    if (h, w) != (rh, rw):
        hwc = cv2.resize(hwc, (rw, rh), interpolation=cv2.INTER_LINEAR)
    hwc = cv2.copyMakeBorder(hwc, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)
    if hwc.ndim == 2:
        hwc = hwc[:, :, None]
    return torch.from_numpy(hwc.astype(np.float32).transpose(2, 0, 1) / 255.0), ratio, (dw, dh)
