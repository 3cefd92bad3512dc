#!/usr/bin/env python3
"""How fast are the ops on NATURAL image content?  The bench inputs are noise / ramps; the table gathers of the CLAHE
kernels conflict less when neighbouring pixels are alike.  Frames = the reference's photograph crops (tests/golden/natural.npz)
mirrored and tiled up to 1080p / 4K."""
import json, os, statistics, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from retinex_image_enhancement_b200 import native  # noqa: E402

PEAK = 6548.8


def time_ms(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts)


def mirror_tile(x, h, w):
    _, _, h0, w0 = x.shape
    row = np.concatenate([x, x[..., ::-1]], axis=3)
    blk = np.concatenate([row, row[:, :, ::-1]], axis=2)
    ry, rx = -(-h // (2 * h0)), -(-w // (2 * w0))
    return np.ascontiguousarray(np.tile(blk, (1, 1, ry, rx))[:, :, :h, :w])


def main():
    arrays = np.load(os.path.join(ROOT, "tests", "golden", "natural.npz"))
    names = ["road_stripe", "asphalt_ragged", "dark_edge"]
    res = {}
    for label, (n, h, w) in {"64x1080p": (64, 1080, 1920), "16x4k": (16, 2160, 3840)}.items():
        frames = []
        for nm in names:
            u8 = arrays[f"{nm}_u8"]
            x = (u8.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None]
            frames.append(torch.from_numpy(mirror_tile(x, h, w)).cuda())
        nat = torch.cat([frames[i % 3] for i in range(n)]).contiguous()
        g = torch.Generator(device="cuda").manual_seed(3)
        noise = torch.rand((n, 3, h, w), device="cuda", generator=g) * 0.7
        enh = torch.rand((n, 3, h, w), device="cuda", generator=g)
        out = torch.empty_like(nat)
        px = n * h * w
        r = {}
        for kind, x in (("natural", nat), ("noise", noise)):
            t_c = time_ms(lambda: native.clahe_lab(x, out=out))
            t_ca = time_ms(lambda: native.content_aware_apply(x, enh, out=out))
            t_ms = time_ms(lambda: native.multiscale_stats(x))
            r[kind] = {"clahe_ms": round(t_c, 4), "clahe_frac_24Bpx": round(24.0 * px / t_c / 1e6 / PEAK, 4),
                       "content_aware_ms": round(t_ca, 4), "content_aware_frac_36Bpx": round(36.0 * px / t_ca / 1e6 / PEAK, 4),
                       "multiscale_stats_ms": round(t_ms, 4)}
        res[label] = r
        del nat, noise, enh, out, frames
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
