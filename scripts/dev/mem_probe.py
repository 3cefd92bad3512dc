import torch, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from scripts.quick_bench import time_op
x = torch.rand(64, 3, 1080, 1920, device="cuda"); out = torch.empty_like(x)
gb = x.numel() * 4 / 1e9
res = {}
res["fill"] = gb / time_op(lambda: out.fill_(0.5), 20)[0] * 1e3
res["memset"] = gb / time_op(lambda: out.zero_(), 20)[0] * 1e3
res["copy(r+w)"] = 2 * gb / time_op(lambda: out.copy_(x), 20)[0] * 1e3
res["read(sum)"] = gb / time_op(lambda: x.sum(), 20)[0] * 1e3
res["read(max)"] = gb / time_op(lambda: x.max(), 20)[0] * 1e3
print(json.dumps({k: round(v, 1) for k, v in res.items()}), "GB/s")
