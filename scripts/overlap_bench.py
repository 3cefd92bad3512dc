#!/usr/bin/env python3
"""Experiment: does running the histogram kernel (K1, L1/shared data pipe bound) of chunk i+1 concurrently with the map
kernel (K3, issue bound) of chunk i beat running the two back to back?  Two streams, the K3 stream at high priority.

    python scripts/overlap_bench.py --chunks 4

Prints one JSON line: sequential op time vs. pipelined time on the same 64 x 1080p batch.  Result (profiles/r3_clahe.md): the
pipelined schedule is slower (0.96-1.15 ms against 0.83 ms), also with the map kernel limited to one CTA per SM through a
development switch (UPR_K3_CTAS, since removed from the library; the variable is ignored now).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retinex_image_enhancement_b200 import native  # noqa: E402
from scripts.quick_bench import make_batch, time_op  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    h, w = 1080, 1920
    x, _ = make_batch(a.n, h, w)
    out = torch.empty_like(x)
    ref = torch.empty_like(x)
    saved = os.environ.pop("UPR_K3_CTAS", None)
    native.clahe_lab(x, out=ref)
    seq = time_op(lambda: native.clahe_lab(x, out=ref), a.iters)[0]
    if saved is not None:
        os.environ["UPR_K3_CTAS"] = saved
    L = native.lib()
    per = a.n // a.chunks
    nbytes = L.upr_clahe_workspace_bytes(per, h, w, 8, 8)
    wss = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(a.chunks)]
    lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
    s1 = torch.cuda.Stream(priority=0)
    s2 = torch.cuda.Stream(priority=-1)
    evs = [torch.cuda.Event() for _ in range(a.chunks)]
    fork = torch.cuda.Event()
    join1, join2 = torch.cuda.Event(), torch.cuda.Event()

    def piped():
        cur = torch.cuda.current_stream()
        fork.record(cur)
        s1.wait_event(fork)
        s2.wait_event(fork)
        for i in range(a.chunks):
            xi, oi = x[i * per:(i + 1) * per], out[i * per:(i + 1) * per]
            native.check(L.upr_clahe_lab_stages_f32(xi.data_ptr(), oi.data_ptr(), per, h, w, 2.0, 8, 8, wss[i].data_ptr(),
                                                    wss[i].numel(), 1, s1.cuda_stream), "k1")
            evs[i].record(s1)
            s2.wait_event(evs[i])
            native.check(L.upr_clahe_lab_stages_f32(xi.data_ptr(), oi.data_ptr(), per, h, w, 2.0, 8, 8, wss[i].data_ptr(),
                                                    wss[i].numel(), 2, s2.cuda_stream), "k3")
        join1.record(s1)
        join2.record(s2)
        cur.wait_event(join1)
        cur.wait_event(join2)

    piped()
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    pip = time_op(piped, a.iters)[0]
    print(json.dumps({"variant": os.environ.get("UPR_CLAHE_VARIANT"), "k3_ctas": saved, "chunks": a.chunks, "sequential_ms": seq,
                      "pipelined_ms": pip, "identical": same}))


if __name__ == "__main__":
    main()
