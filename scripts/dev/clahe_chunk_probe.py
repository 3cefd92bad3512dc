#!/usr/bin/env python3
"""Does a two-stream chunk schedule (K1 + K3 of a chunk back to back on one of two streams) pay for CLAHE / fused Retinex + CLAHE?
Whole batch in one call vs chunks of F frames on 2 streams, both as CUDA graphs.  16 x 4K and 64 x 1080p."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from retinex_image_enhancement_b200 import native  # noqa: E402


def time_graph(build, iters=15):
    build(); build()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        build()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return round(statistics.median(ts), 4)


def main():
    side = [torch.cuda.Stream(), torch.cuda.Stream()]
    res = {}
    for label, (n, h, w, fpcs) in {"16x4k": (16, 2160, 3840, (2, 3, 4, 8)), "64x1080p": (64, 1080, 1920, (8, 12, 16, 32))}.items():
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.rand((n, 3, h, w), device="cuda", generator=g) * 0.7
        illu = torch.rand((n, 1, h, w), device="cuda", generator=g) * 0.8 + 0.1
        e = torch.rand((n, 3, h, w), device="cuda", generator=g)
        out = torch.empty_like(x)

        def chunked(fn, fpc):
            def run():
                cur = torch.cuda.current_stream()
                fork = torch.cuda.Event(); fork.record(cur)
                for s in side:
                    s.wait_event(fork)
                for k, a in enumerate(range(0, n, fpc)):
                    b = min(a + fpc, n)
                    with torch.cuda.stream(side[k % 2]):
                        fn(a, b)
                for s in side:
                    ev = torch.cuda.Event(); ev.record(s); cur.wait_event(ev)
            return run

        plain = lambda a, b: native.clahe_lab(x[a:b], out=out[a:b])
        fused = lambda a, b: native.retinex_clahe(x[a:b], illu[a:b], e[a:b], out=out[a:b])
        r = {"clahe_whole": time_graph(lambda: plain(0, n)), "fused_whole": time_graph(lambda: fused(0, n))}
        for fpc in fpcs:
            r[f"clahe_chunks_of_{fpc}"] = time_graph(chunked(plain, fpc))
            r[f"fused_chunks_of_{fpc}"] = time_graph(chunked(fused, fpc))
        res[label] = r
        del x, illu, e, out
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
