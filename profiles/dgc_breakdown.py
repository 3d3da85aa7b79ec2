"""Dev tool: per-layer CUDA-event times of a DGC-SPN forward (config 3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from deeprob_kit_b200.spn.models import DgcSpn
torch.manual_seed(0)
m = DgcSpn((1, 28, 28), n_batch=8, sum_channels=8, depthwise=True).cuda().eval()
x = torch.randn(32768, 1, 28, 28, device="cuda")
with torch.no_grad():
    for _ in range(2):
        m(x)
    layers = [("leaf", m.base_layer)] + [(type(l).__name__ + str(i), l) for i, l in enumerate(m.layers)] + [("root", m.root_layer)]
    h = x
    tot = 0
    for name, l in layers:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = l(h); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1); tot += t
        gb = (h.numel() + out.numel()) * 4 / 1e9
        print("%-24s in %-22s out %-22s %8.3f ms  %7.1f GB/s" % (name, tuple(h.shape[1:]), tuple(out.shape[1:]), t, gb / (t * 1e-3)))
        h = out
    print("total %.3f ms" % tot)
