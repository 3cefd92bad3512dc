import sys, os, torch, json
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
from retinex_image_enhancement_b200 import native
from scripts.quick_bench import make_batch, time_op
x,_=make_batch(64,1080,1920)
out=torch.empty_like(x)
native.clahe_lab(x,out=out)
r={}
for m,name in ((1,"k1"),(2,"k3"),(3,"op")):
    r[name]=time_op(lambda: native.clahe_lab(x,out=out,stage_mask=m),20)[0]
print(json.dumps(r))
