"""Launch list of one RealNVP1d(3072, 8 flows, depth 2, units 512) log-prob at batch 16384 (config 4):
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/prof_flow.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

from deeprob_kit_b200.flows.models import RealNVP1d  # noqa: E402

torch.manual_seed(0)
m = RealNVP1d(3072, n_flows=8, depth=2, units=512).cuda().eval()
x = torch.rand(16384, 3072, device="cuda")
with torch.no_grad():
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
