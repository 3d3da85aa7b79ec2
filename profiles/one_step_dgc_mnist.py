"""One forward+backward step of the MNIST-example DGC-SPN (n_batch=16, sum_channels=32, depthwise, 2 pooling levels) at
batch 8192, bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from deeprob_kit_b200.spn.models import DgcSpn  # noqa: E402

torch.manual_seed(0)
m = DgcSpn((1, 28, 28), n_batch=16, sum_channels=32, depthwise=True, n_pooling=2).cuda().train()
x = torch.randn(8192, 1, 28, 28, device="cuda")


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
