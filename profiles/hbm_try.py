import os, sys, json
sys.path[:0]=['/root/repo','/root/repo/tests']
import torch
import bench
dev=torch.device("cuda",0)
for rk in ("24","1"):
    os.environ["DPK_LEAF_MMA_MIN_RK"]=rk
    r=bench.bench_hbm_bound(dev, 65536)
    print(rk, r["ms_per_step"], r["frac"])
from deeprob_kit_b200 import _lib
from deeprob_kit_b200.spn.models import GaussianRatSpn
m=GaussianRatSpn(784, rg_depth=3, rg_repetitions=1, rg_batch=8, rg_sum=8, random_state=42).eval().to(dev)
x=torch.randn(65536,784,device=dev)
with torch.no_grad():
    for _ in range(3): m(x)
    torch.cuda.synchronize(); _lib.profile_read(); _lib.profile_enable(True)
    for _ in range(10): m(x)
    torch.cuda.synchronize(); _lib.profile_enable(False)
    ms,c=_lib.profile_read()
print({k:round(v/10,4) for k,v in ms.items() if v>0})
