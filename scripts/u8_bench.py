#!/usr/bin/env python3
"""Developer benchmark of the packed u8 boundary of the CLAHE op (upr_clahe_lab_u8 / upr_clahe_lab_f32_u8) on 64 x 1080p."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retinex_image_enhancement_b200 import native  # noqa: E402
from scripts.quick_bench import make_batch, time_op  # noqa: E402

n, h, w = 64, 1080, 1920
xf, _ = make_batch(n, h, w)
x8 = (xf * 255.0).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
o8 = torch.empty_like(x8)
of = torch.empty_like(xf)
px = n * h * w
res = {"f32_f32_ms": time_op(lambda: native.clahe_lab(xf, out=of), 20)[0],
       "u8_u8_ms": time_op(lambda: native.clahe_lab_u8(x8, out=o8), 20)[0],
       "f32_u8_ms": time_op(lambda: native.clahe_lab_f32_u8(xf, out=o8), 20)[0]}
res["u8_u8_gpix_s"] = px / res["u8_u8_ms"] / 1e6
res["f32_u8_gpix_s"] = px / res["f32_u8_ms"] / 1e6
print(json.dumps(res))
