"""Dev tool (GPU box): sweep the leaf-kernel ring knobs (DPK_LEAF_STAGES / DPK_LEAF_CH) on BASELINE config 2
and print the CUDA-event time of each kernel category.   python profiles/tune_leaf.py [optimize_scale]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from deeprob_kit_b200 import _lib  # noqa: E402
from deeprob_kit_b200.spn.models import GaussianRatSpn  # noqa: E402

opt = len(sys.argv) > 1 and sys.argv[1] == "optimize_scale"
torch.manual_seed(0)
m = GaussianRatSpn(784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10, random_state=42, optimize_scale=opt).eval().cuda()
x = torch.randn(65536, 784, device="cuda")
ref = None
for stages in (2, 3, 4):
    for ch in (0, 7, 10, 14, 20, 25):
        os.environ["DPK_LEAF_STAGES"] = str(stages)
        os.environ["DPK_LEAF_CH"] = str(ch)
        with torch.no_grad():
            for _ in range(3):
                y = m(x)
            torch.cuda.synchronize()
            _lib.profile_read()
            _lib.profile_enable(True)
            for _ in range(10):
                y = m(x)
            torch.cuda.synchronize()
            _lib.profile_enable(False)
            ms, n = _lib.profile_read()
        if ref is None:
            ref = y.clone()
        print("stages %d ch %2d  leaf %.4f ms  einsum %.4f root %.4f  maxdiff %.2e" % (
            stages, ch, ms["ratspn_leaf"] / 10, ms["ratspn_einsum"] / 10, ms["ratspn_root"] / 10,
            float((y - ref).abs().max())), flush=True)
