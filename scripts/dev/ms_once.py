#!/usr/bin/env python3
"""One warm + a few profiled calls of the multi-scale statistics kernel on 16 x 4K (for `ncu -k regex:k_ms_stream`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from retinex_image_enhancement_b200 import native  # noqa: E402

x = torch.rand((16, 3, 2160, 3840), device="cuda") * 0.6
for _ in range(4):
    native.multiscale_stats(x)
torch.cuda.synchronize()
