#!/usr/bin/env python3
"""Design probe: does a frame-at-a-time schedule (the three content-aware kernels per FRAME, frames interleaved over S streams)
beat one launch per phase over the whole batch?  Per frame the intermediate plane stays L2-resident between the phases and the
issue-bound blur of one frame overlaps the memory-bound phases of another.  Everything is captured in one CUDA graph, so Python /
ctypes overhead is outside the timed region.  Also runs the same experiment for CLAHE (K1 -> K3 per frame)."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from retinex_image_enhancement_b200 import native  # noqa: E402


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


def graph_of(per_frame, n, nstreams, group):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for s in streams:                                   # warm every (stream, workspace) pair outside the capture
        with torch.cuda.stream(s):
            per_frame(0, min(group, n))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    with torch.cuda.stream(cap):
        with torch.cuda.graph(g, stream=cap):
            fork = torch.cuda.Event(); fork.record(cap)
            for s in streams:
                s.wait_event(fork)
            for k, f0 in enumerate(range(0, n, group)):
                with torch.cuda.stream(streams[k % nstreams]):
                    per_frame(f0, min(f0 + group, n))
            for s in streams:
                ev = torch.cuda.Event(); ev.record(s); cap.wait_event(ev)
    return g


GROUPS = tuple(int(v) for v in os.environ.get('FPC', '1,2,4').split(','))
STREAMS = tuple(int(v) for v in os.environ.get('STREAMS', '1,2,3,4').split(','))


def main():
    res = {}
    for name, (n, h, w) in {"4k": (16, 2160, 3840), "1080p": (64, 1080, 1920)}.items():
        x = torch.rand((n, 3, h, w), device="cuda") * 0.6
        enh = torch.rand((n, 3, h, w), device="cuda")
        out = torch.empty_like(enh)
        base_ca = time_ms(lambda: native.content_aware_apply(x, enh, out=out))
        base_cl = time_ms(lambda: native.clahe_lab(x, out=out))
        r = {"content_aware_batch_ms": base_ca, "clahe_batch_ms": base_cl, "variants": []}
        for group in GROUPS:
            for ns in STREAMS:
                if group * ns > n:
                    continue
                try:
                    g1 = graph_of(lambda a, b: native.content_aware_apply(x[a:b], enh[a:b], out=out[a:b]), n, ns, group)
                    t1 = time_ms(g1.replay)
                    g2 = graph_of(lambda a, b: native.clahe_lab(x[a:b], out=out[a:b]), n, ns, group)
                    t2 = time_ms(g2.replay)
                    r["variants"].append({"frames_per_call": group, "streams": ns, "content_aware_ms": t1, "clahe_ms": t2})
                    del g1, g2
                except Exception as e:  # pragma: no cover
                    r["variants"].append({"frames_per_call": group, "streams": ns, "error": repr(e)})
                native.release_workspaces()
        res[name] = r
        del x, enh, out
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
