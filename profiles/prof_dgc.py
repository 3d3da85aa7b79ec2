"""Launch list of one DgcSpn log-prob at batch 32768 (config 3; `--mnist` = the MNIST-example setting):
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/prof_dgc.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

from deeprob_kit_b200.spn.models import DgcSpn  # noqa: E402

torch.manual_seed(0)
if "--mnist" in sys.argv:
    m = DgcSpn((1, 28, 28), n_batch=16, sum_channels=32, depthwise=True, n_pooling=2).cuda().eval()
else:
    m = DgcSpn((1, 28, 28), n_batch=8, sum_channels=8, depthwise=True).cuda().eval()
print([(type(l).__name__, tuple(l.out_features)) for l in m.layers])
x = torch.randn(32768, 1, 28, 28, device="cuda")
with torch.no_grad():
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
