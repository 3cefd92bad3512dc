#!/usr/bin/env python3
"""Gain/clamp pass (24 B/px): separate output vs in place, small vs large batches (does the 256-frame working set cost bandwidth?)."""
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
    return statistics.median(ts)


res = {}
h, w = 2160, 3840
for n in (3, 16, 64):
    enh = torch.rand((n, 3, h, w), device="cuda")
    out = torch.empty_like(enh)
    gain = torch.full((n,), 1.03, device="cuda")
    by = 24.0 * n * h * w
    t_sep = time_ms(lambda: native.scale_clamp(enh, gain, out=out))
    t_inp = time_ms(lambda: native.scale_clamp(enh, gain, out=enh))
    t_copy = time_ms(lambda: out.copy_(enh))
    res[n] = {"separate_ms": round(t_sep, 4), "separate_TBs": round(by / t_sep / 1e9, 3), "in_place_ms": round(t_inp, 4),
              "in_place_TBs": round(by / t_inp / 1e9, 3), "torch_copy_TBs": round(by / t_copy / 1e9, 3)}
    del enh, out
    torch.cuda.empty_cache()
print(json.dumps({'lib': os.path.basename(native.LIB_PATH), **{str(k): v for k, v in res.items()}}))
