"""Dev tool: time bench.py's hbm_bound_check workload (narrow RAT-SPN, R=1, K=8) with per-category kernel times."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
def leaf_ms(tag):
    from deeprob_kit_b200 import _lib
    from deeprob_kit_b200.spn.models import GaussianRatSpn
    torch.manual_seed(0)
    model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=1, rg_batch=8, rg_sum=8, random_state=42).eval().to(dev)
    x = torch.randn(65536, 784, device=dev)
    with torch.no_grad():
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        _lib.profile_read()
        _lib.profile_enable(True)
        for _ in range(20):
            model(x)
        torch.cuda.synchronize()
        _lib.profile_enable(False)
        ms, launches = _lib.profile_read()
    print('RES', tag, {k: round(v / 20, 4) for k, v in ms.items() if v > 0})


for env in ({}, {"DPK_STREAM_DBG": "3"}, {"DPK_STREAM_DBG": "8"}, {"DPK_STREAM_MT": "128"}):
    os.environ.update(env)
    leaf_ms(str(env))
    for k in env:
        del os.environ[k]
print("stream on", json.dumps(bench.bench_hbm_bound(dev, 65536, 50)))
os.environ["DPK_LEAF_STREAM"] = "0"
print("stream off", json.dumps(bench.bench_hbm_bound(dev, 65536, 50)))
del os.environ["DPK_LEAF_STREAM"]
from deeprob_kit_b200 import _lib
from deeprob_kit_b200.spn.models import GaussianRatSpn
model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=1, rg_batch=8, rg_sum=8, random_state=42).eval().to(dev)
x = torch.randn(65536, 784, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    _lib.profile_read()
    _lib.profile_enable(True)
    for _ in range(20):
        model(x)
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    ms, launches = _lib.profile_read()
    print({k: round(v / 20, 4) for k, v in ms.items() if v > 0})
