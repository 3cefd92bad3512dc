"""PCIe ceiling on the box: pinned H2D alone, D2H alone, both directions at once; then the host-buffer CLAHE entry at several chunk sizes."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from retinex_image_enhancement_b200 import native
if os.environ.get('UPR_LIB'):
    native.LIB_PATH = os.environ['UPR_LIB']

n, h, w = 64, 1080, 1920
hx = torch.rand((n, 3, h, w), dtype=torch.float32).pin_memory()
hy = torch.empty_like(hx).pin_memory()
dx = torch.empty((n, 3, h, w), device="cuda"); dy = torch.rand((n, 3, h, w), device="cuda")
gb = hx.numel() * 4 / 1e9
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best

def h2d():
    with torch.cuda.stream(s1): dx.copy_(hx, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hy.copy_(dy, non_blocking=True)
def both():
    h2d(); d2h()
res = {"h2d_GBs": gb / wall(h2d), "d2h_GBs": gb / wall(d2h), "both_GBs_per_dir": gb / wall(both)}
for c in (1, 2, 3, 4, 6, 8):
    t = wall(lambda: native.clahe_lab_host(hx, out=hy, frames_per_chunk=c), 3)
    res[f"host_api_chunk{c}_ms"] = t * 1e3
print(json.dumps(res))
