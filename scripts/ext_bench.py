"""Developer micro-benchmark of the extension ops (Gaussian blur / MSR over 16 4K frames)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retinex_image_enhancement_b200 import extensions as E
from scripts.quick_bench import time_op
x = torch.rand(16, 3, 2160, 3840, device="cuda") * 0.9 + 0.05
px = x.numel()
res = {}
for name, fn in (("blur5", lambda: E.gaussian_blur(x, 5)), ("blur15", lambda: E.gaussian_blur(x, 15)), ("blur31", lambda: E.gaussian_blur(x, 31, 5.0)),
                 ("msr_7_15_31", lambda: E.multi_scale_retinex(x, (7, 15, 31))), ("pyr_down", lambda: E.pyr_down(x)), ("gamma", lambda: E.gamma_correct(x, 0.45))):
    med, _ = time_op(fn, 10)
    res[name] = {"ms": round(med, 4), "GBs_8B_per_elem": round(8 * px / med / 1e6, 1)}
print(json.dumps(res))
