#!/usr/bin/env python3
"""Times saliency / content-aware / multi-scale / CLAHE ops with an alternative build of the library (UPR_LIB=path), 16 x 4K and
64 x 1080p.  Development probe for compile-time variants (launch bounds, macros); the product always loads the in-tree .so."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from retinex_image_enhancement_b200 import native  # noqa: E402
if os.environ.get("UPR_LIB"):
    native.LIB_PATH = os.environ["UPR_LIB"]


def time_ms(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return round(statistics.median(ts), 4)


res = {"lib": native.LIB_PATH}
for name, (n, h, w) in {"16x4k": (16, 2160, 3840), "64x1080p": (64, 1080, 1920)}.items():
    x = torch.rand((n, 3, h, w), device="cuda") * 0.6
    enh = torch.rand((n, 3, h, w), device="cuda")
    out = torch.empty_like(enh)
    res[name] = {"saliency": time_ms(lambda: native.saliency(x)), "content_aware_apply": time_ms(lambda: native.content_aware_apply(x, enh, out=out)),
                 "multiscale_stats": time_ms(lambda: native.multiscale_stats(x)), "clahe": time_ms(lambda: native.clahe_lab(x, out=out))}
    del x, enh, out
print(json.dumps(res))
