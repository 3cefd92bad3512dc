#!/usr/bin/env python3
"""Times the CLAHE op (64 x 1080p, 16 x 4K) with an alternative build of the library (UPR_LIB=path) and checks its output against
the in-tree library's (bit-equal).  Development probe for compile-time kernel variants."""
import ctypes, json, os, statistics, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from retinex_image_enhancement_b200 import native  # noqa: E402


def time_ms(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return round(statistics.median(ts), 4)


def main():
    if len(sys.argv) > 1:      # child: time one library, print the result and a checksum
        native.LIB_PATH = sys.argv[1]
        res = {"lib": os.path.basename(sys.argv[1])}
        for name, (n, h, w) in {"64x1080p": (64, 1080, 1920), "16x4k": (16, 2160, 3840)}.items():
            g = torch.Generator(device="cuda").manual_seed(5)
            x = torch.rand((n, 3, h, w), device="cuda", generator=g) * 0.7
            out = torch.empty_like(x)
            res[name] = {"clahe_ms": time_ms(lambda: native.clahe_lab(x, out=out)),
                         "sum": float(out.double().sum().item()), "sha": hash(out[0].cpu().numpy().tobytes()) & 0xffffffff}
            del x, out
        print(json.dumps(res))
        return
    libs = [native.LIB_PATH] + sorted(p for p in sys.argv[1:] if p) + sorted(
        os.path.join(os.path.dirname(native.LIB_PATH), f) for f in os.listdir(os.path.dirname(native.LIB_PATH))
        if f.startswith("libupretinex_b200_") and f.endswith(".so"))
    for lib in libs:
        out = subprocess.run([sys.executable, __file__, lib], capture_output=True, text=True, env=dict(os.environ, PYTHONHASHSEED="0"))
        print(out.stdout.strip() or out.stderr[-1500:])


if __name__ == "__main__":
    main()
