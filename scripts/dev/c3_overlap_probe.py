#!/usr/bin/env python3
"""Design probe for BASELINE config 3 (multi-scale statistics + gain): does a two-stream chunk schedule (statistics of chunk i+1
under the gain pass of chunk i) beat statistics-then-gain over the whole batch?  Both captured as CUDA graphs (no launch overhead)."""
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
    n, h, w = 16, 2160, 3840
    x = torch.rand((n, 3, h, w), device="cuda") * 0.6
    enh = torch.rand((n, 3, h, w), device="cuda")
    out = torch.empty_like(enh)
    side = [torch.cuda.Stream(), torch.cuda.Stream()]
    res = {}

    def sequential():
        _m, g = native.multiscale_stats(x)
        native.scale_clamp(enh, g, out=out)

    def chunked(fpc, nstreams):
        def run():
            cur = torch.cuda.current_stream()
            fork = torch.cuda.Event(); fork.record(cur)
            for s in side[:nstreams]:
                s.wait_event(fork)
            for k, a in enumerate(range(0, n, fpc)):
                b = min(a + fpc, n)
                with torch.cuda.stream(side[k % nstreams]):
                    _m, g = native.multiscale_stats(x[a:b])
                    native.scale_clamp(enh[a:b], g, out=out[a:b])
            for s in side[:nstreams]:
                ev = torch.cuda.Event(); ev.record(s); cur.wait_event(ev)
        return run

    res["sequential_whole_batch"] = time_graph(sequential)
    for fpc in (2, 3, 4, 8):
        for ns in (1, 2):
            res[f"chunks_of_{fpc}_on_{ns}_streams"] = time_graph(chunked(fpc, ns))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
