#!/usr/bin/env python3
"""Developer micro-benchmark of upr_clahe_lab_f32 (not the contract bench; see bench.py)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retinex_image_enhancement_b200 import native  # noqa: E402


def make_batch(n, h, w, seed=1000):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand((n, 3, h, w), device="cuda", generator=g)
    kinds = []
    for i in range(n):
        k = i % 4
        if k in (0, 2):
            x[i] *= 0.3; kinds.append("dark")
        elif k == 1:
            kinds.append("uniform")
        else:
            if (i // 4) % 2 == 0:
                x[i] = 0.3; kinds.append("const")
            else:
                xs = torch.linspace(0, 1, w, device="cuda")[None, :].expand(h, w)
                ys = torch.linspace(0, 1, h, device="cuda")[:, None].expand(h, w)
                x[i, 0], x[i, 1], x[i, 2] = xs, ys, (xs + ys) / 2; kinds.append("ramp")
    return x, kinds


def time_op(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--kinds", default="mix")
    a = ap.parse_args()
    x, kinds = make_batch(a.n, a.h, a.w)
    if a.kinds != "mix":
        g = torch.Generator(device="cuda").manual_seed(7)
        x = torch.rand(x.shape, device="cuda", generator=g)
        if a.kinds == "dark":
            x *= 0.3
        elif a.kinds == "const":
            x.fill_(0.3)
        elif a.kinds == "smooth":
            # natural-image-like content: low-pass noise (16x bilinear up-sampling of coarse noise) + 2 % fine noise
            coarse = torch.rand((a.n, 3, a.h // 16 + 1, a.w // 16 + 1), device="cuda", generator=g)
            x = torch.nn.functional.interpolate(coarse, size=(a.h, a.w), mode="bicubic", align_corners=False).clamp_(0, 1)
            x = (x * 0.98 + 0.02 * torch.rand(x.shape, device="cuda", generator=g)).contiguous()
    out = torch.empty_like(x)
    med, best = time_op(lambda: native.clahe_lab(x, out=out), a.iters)
    px = a.n * a.h * a.w
    peak = 6548.8
    res = {"op": "clahe_lab_f32", "n": a.n, "h": a.h, "w": a.w, "kinds": a.kinds, "ms_median": med, "ms_best": best,
           "gpix_s": px / med / 1e6, "alg_GBs": 24 * px / med / 1e6, "frac_of_measured_peak": 24 * px / med / 1e6 / peak}
    # copy roofline on the same box, same bytes (12 B/px read + 12 B/px write)
    med_c, best_c = time_op(lambda: out.copy_(x), a.iters)
    res["copy_ms"] = med_c
    res["copy_GBs"] = 24 * px / med_c / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
