"""Dev tool: end-to-end rate of streaming.log_prob_host at config 2 for several chunk sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from deeprob_kit_b200.spn.models import GaussianRatSpn
from deeprob_kit_b200.spn.streaming import log_prob_host
dev = torch.device("cuda:0")
model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10, random_state=42).eval().to(dev)
B = 65536
xh = torch.randn(B, 784).pin_memory()
oh = torch.empty(B, 1).pin_memory()
for chunk in (4096, 8192, 16384, 32768, 65536):
    for _ in range(3):
        log_prob_host(model, xh, chunk=chunk, out_host=oh)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        log_prob_host(model, xh, chunk=chunk, out_host=oh)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print("E2E chunk %6d: %.3f ms  %.3e evals/s  (%.1f GB/s of input)" % (chunk, ms, B * 784 / ms * 1e3, B * 784 * 4 / ms / 1e6))
