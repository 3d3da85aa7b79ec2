"""BASELINE config 5: RAT-SPN (D=784, depth 3, R=16, K=O=10) batch-EM step, batch 65536 per GPU, one NCCL all-reduce
of the flat sufficient statistics per step.  Launch like bench.py:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/bench_em.py [--steps K]
Prints one JSON line on rank 0 (time per EM step = E-step forward + backward statistics + all-reduce + M-step)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from deeprob_kit_b200 import _lib  # noqa: E402
from deeprob_kit_b200.spn import em  # noqa: E402
from deeprob_kit_b200.spn.models import GaussianRatSpn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=65536)
args = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
torch.manual_seed(0)
model = GaussianRatSpn(784, rg_depth=3, rg_repetitions=16, rg_batch=10, rg_sum=10, random_state=42, optimize_scale=True).to(dev)
x = torch.randn(args.batch, 784, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
lls = []
for _ in range(args.warmup):
    lls.append(em.em_step(model, x, 0.5))
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
_lib.profile_read()
_lib.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    lls.append(em.em_step(model, x, 0.5))
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
_lib.profile_enable(False)
ms, cnt = _lib.profile_read()
t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# every rank must hold identical parameters after identical M-steps
chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
same = True
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((hi - lo).abs() <= 1e-9 * hi.abs())
if rank == 0:
    step_ms = float(t[0]) / args.steps
    print(json.dumps({"metric": "RAT-SPN D=784 batch-EM step", "n_gpus": world, "batch_per_gpu": args.batch,
                      "ms_per_em_step": step_ms, "samples_per_s": world * args.batch / (step_ms * 1e-3),
                      "mean_ll_trajectory": [round(v, 3) for v in lls], "replicas_identical": same,
                      "kernel_ms_per_step": {k: round(v / args.steps, 4) for k, v in ms.items() if v > 0}}), flush=True)
if world > 1:
    dist.destroy_process_group()
