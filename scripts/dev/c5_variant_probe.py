#!/usr/bin/env python3
"""c5 chain (content-aware + multi-scale, one call) and content-aware apply on 16 x 4K with an alternative library build
(argument = path of the .so); prints one JSON line.  Development probe for compile-time kernel variants."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from retinex_image_enhancement_b200 import native  # noqa: E402
if len(sys.argv) > 1:
    native.LIB_PATH = sys.argv[1]


def time_ms(fn, iters=15):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return round(statistics.median(ts), 4)


n, h, w = 16, 2160, 3840
x = torch.rand((n, 3, h, w), device="cuda") * 0.6
enh = torch.rand((n, 3, h, w), device="cuda")
out = torch.empty_like(enh)
res = {"lib": os.path.basename(native.LIB_PATH),
       "chain_one_call": time_ms(lambda: native.content_multiscale_apply(x, enh, out=out)),
       "content_aware_apply": time_ms(lambda: native.content_aware_apply(x, enh, out=out)),
       "saliency": time_ms(lambda: native.saliency(x)),
       "checksum": float(out.double().sum().item())}
print(json.dumps(res))
