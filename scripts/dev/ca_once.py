"""One content-aware call over 16 x 4K (for ncu captures of its kernels)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from retinex_image_enhancement_b200 import native
n = int(os.environ.get("N", "16"))
x = torch.rand((n, 3, 2160, 3840), device="cuda") * 0.6
enh = torch.rand((n, 3, 2160, 3840), device="cuda")
out = torch.empty_like(enh)
for _ in range(2):
    native.content_multiscale_apply(x, enh, out=out)
torch.cuda.synchronize()
