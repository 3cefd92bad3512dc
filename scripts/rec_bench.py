import sys, os, torch, json
sys.path.insert(0, os.getcwd())
from retinex_image_enhancement_b200 import native
from scripts.quick_bench import time_op
x=torch.rand(64,3,1080,1920,device="cuda"); e=torch.rand_like(x); illu=(x[:, :1]*0.5+0.25).contiguous()
print("recombine", time_op(lambda: native.retinex_recombine(x, illu, e, want_reflectance=False), 20))
print("recombine+refl", time_op(lambda: native.retinex_recombine(x, illu, e, want_reflectance=True), 20))
print("decompose", time_op(lambda: native.retinex_decompose(x, illu), 20))
out=torch.empty_like(x)
g=torch.ones(64,device="cuda")
print("scale_clamp", time_op(lambda: native.scale_clamp(e, g, out=out), 20))
print("copy", time_op(lambda: out.copy_(x), 20))
