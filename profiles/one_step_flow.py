"""One training step (forward + backward) of BASELINE config 4, RealNVP1d(3072, n_flows=8, depth=2, units=512) at batch
16384, bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off` (launch list / traffic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import warnings

import torch

warnings.simplefilter("ignore")
from deeprob_kit_b200.flows.models import RealNVP1d  # noqa: E402

torch.manual_seed(0)
m = RealNVP1d(3072, n_flows=8, depth=2, units=512).cuda().train()
with torch.no_grad():
    for n_, p in m.named_parameters():
        if "scale_act" in n_:
            p.fill_(0.5)
x = torch.rand(16384, 3072, device="cuda")


def step():
    m.zero_grad(set_to_none=True)
    m(x).sum().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
