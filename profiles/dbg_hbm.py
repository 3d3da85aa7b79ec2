"""Dev tool: the hbm_bound_check model at full batch -- run, synchronise, compare with the exact CUDA-core path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from deeprob_kit_b200.spn.models import GaussianRatSpn
dev = torch.device("cuda:0")
B = int(os.environ.get("DBG_B", "65536"))
model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=1, rg_batch=8, rg_sum=8, random_state=42).eval().to(dev)
x = torch.randn(B, 784, device=dev)
with torch.no_grad():
    try:
        for _ in range(int(os.environ.get('DBG_N', '5'))):
            out = model(x)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("DBGH FAIL", str(e).splitlines()[0][:150])
        sys.exit(0)
    saved = {k: os.environ.get(k) for k in ("DPK_LEAF_STREAM", "DPK_TREE_MMA", "DPK_LEAF_MMA")}
    os.environ.update({"DPK_LEAF_STREAM": "0", "DPK_TREE_MMA": "0", "DPK_LEAF_MMA": "0"})
    ref = model(x)
    print("DBGH OK maxrel", float(((out - ref).abs() / ref.abs().clamp_min(1)).max()))
