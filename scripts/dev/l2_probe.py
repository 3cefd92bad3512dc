"""ncu target: CLAHE over 64 x 1080p in chunks of F frames (one workspace re-used), a handful of launches only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from retinex_image_enhancement_b200 import native
from scripts.quick_bench import make_batch
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, h, w = 32, 1080, 1920
x, _ = make_batch(n, h, w)
out = torch.empty_like(x)
for rep in range(2):
    for i in range(0, n, F):
        native.clahe_lab(x[i:i + F], out=out[i:i + F])
torch.cuda.synchronize()
