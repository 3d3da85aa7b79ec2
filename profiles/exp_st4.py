"""Experiment: leaf kernel with 4 samples/lane (128-sample tile) vs 2 samples/lane on a D=392 model (fits both)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from deeprob_kit_b200 import _lib
from deeprob_kit_b200.spn.models import GaussianRatSpn
for opt in (False, True):
    torch.manual_seed(0)
    m = GaussianRatSpn(392, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10, random_state=42, optimize_scale=opt).eval().cuda()
    x = torch.randn(131072, 392, device="cuda")
    ref = None
    for no128 in ("1", "0"):
        os.environ["DPK_LEAF_NO128"] = no128
        with torch.no_grad():
            for _ in range(3):
                y = m(x)
            torch.cuda.synchronize(); _lib.profile_read(); _lib.profile_enable(True)
            for _ in range(10):
                y = m(x)
            torch.cuda.synchronize(); _lib.profile_enable(False)
            ms, n = _lib.profile_read()
        ref = y if ref is None else ref
        fma = 2 * 392 * 10 * 16 * 131072
        print("optimize_scale=%s tile=%s leaf %.4f ms  %.1f TFLOP/s  maxdiff %.1e" % (opt, "64" if no128 == "1" else "128", ms["ratspn_leaf"] / 10, 2 * fma / (ms["ratspn_leaf"] / 10 * 1e-3) / 1e12, float((y - ref).abs().max())), flush=True)
