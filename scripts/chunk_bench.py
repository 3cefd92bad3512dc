#!/usr/bin/env python3
"""Experiment: upr_clahe_lab_f32 over 64 x 1080p called in chunks of F frames on one stream, the chunk's workspace (hence
its u8 Lab intermediate) re-used by every chunk -- does keeping the Lab intermediate L2-resident pay for the smaller launches?"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retinex_image_enhancement_b200 import native  # noqa: E402
from scripts.quick_bench import make_batch, time_op  # noqa: E402

n, h, w = 64, 1080, 1920
x, _ = make_batch(n, h, w)
ref = torch.empty_like(x)
out = torch.empty_like(x)
native.clahe_lab(x, out=ref)
res = {"variant": os.environ.get("UPR_CLAHE_VARIANT"), "whole_ms": time_op(lambda: native.clahe_lab(x, out=ref), 20)[0]}
for F in (32, 16, 8, 4, 2):
    native.release_workspaces()

    def run():
        for i in range(0, n, F):
            native.clahe_lab(x[i:i + F], out=out[i:i + F])

    run()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    res[f"chunk{F}_ms"] = time_op(run, 20)[0]
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run()
        with torch.cuda.graph(g, stream=s):
            run()
    torch.cuda.synchronize()
    res[f"chunk{F}_graph_ms"] = time_op(g.replay, 20)[0]
print(json.dumps(res))
