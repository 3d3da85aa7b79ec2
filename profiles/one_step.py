"""One eager step of the bench workload bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`:
  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none --csv --log-file gpurun_out/traffic.csv python profiles/one_step.py [--mode infer|em|train|hbm]
A number printed under ncu is never a bench value; this script prints none."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from deeprob_kit_b200.spn import em  # noqa: E402
from deeprob_kit_b200.spn.models import GaussianRatSpn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="infer", choices=["infer", "em", "train", "hbm"])
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if args.mode == "hbm":
    model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=1, rg_batch=8, rg_sum=8, random_state=42).eval().to(dev)
else:
    model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10, random_state=42,
                           optimize_scale=(args.mode == "train")).to(dev)
x = torch.randn(args.batch, 784, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))


def step():
    if args.mode in ("infer", "hbm"):
        with torch.no_grad():
            model.eval()(x)
    elif args.mode == "em":
        em.em_step(model, x, 0.5)
    else:
        model.zero_grad(set_to_none=True)
        model.loss(model(x)).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
