"""Developer micro-benchmark of upr_retinex_clahe_f32 (fused Retinex recombination + CLAHE)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retinex_image_enhancement_b200 import native
from scripts.quick_bench import make_batch, time_op
x, _ = make_batch(64, 1080, 1920)
e = torch.rand_like(x)
illu = (x[:, :1] * 0.5 + 0.25).contiguous()
out = torch.empty_like(x)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
med, best = time_op(lambda: native.retinex_clahe(x, illu, e, out=out), iters)
px = x.shape[0] * 1080 * 1920
print(json.dumps({"op": "retinex_clahe", "ms": med, "best": best, "alg_GBs": 40 * px / med / 1e6, "frac": 40 * px / med / 1e6 / 6548.8}))
