"""Data-parallel training glue: one rank per GPU, the batch sharded, parameters replicated.

The reference's training loop (deeprob/torch/routines.py:158-166: zero_grad -> model(inputs) -> model.loss ->
backward -> optimizer.step -> apply_constraints) knows nothing about ranks.  `distribute(model)` makes that loop
data-parallel WITHOUT modifying it: parameters and buffers are broadcast from rank 0, and every parameter gets a
post-accumulate-grad hook; when the last gradient of a backward pass has landed, all gradients are packed into ONE
flat fp32 buffer, summed with a single all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests), averaged
and unpacked -- before `optimizer.step()` reads them.  The SPN models produce all their parameter gradients from one
autograd node, so bucketing buys nothing: one latency-bound collective of the whole model (1.4 MB at BASELINE config 2)
per step, launched on the stream the backward ran on.
"""
from typing import List, Optional

import torch
import torch.distributed as dist


class GradientAllReduce:
    """Installs the hooks described above; keep the object alive as long as the model trains (`remove()` undoes it)."""

    def __init__(self, model: torch.nn.Module, group=None, average: bool = True):
        self.group, self.average = group, average
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        self._pending = 0
        self._seen = set()
        self._flat: Optional[torch.Tensor] = None
        self.calls = 0                       # number of collectives launched (one per backward pass)
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _hook(self, p: torch.nn.Parameter) -> None:
        self._seen.add(id(p))
        if len(self._seen) < len(self.params):
            return
        self._seen.clear()
        self.reduce()

    @torch.no_grad()
    def reduce(self) -> None:
        """Pack -> all_reduce(sum) -> average -> unpack, for the parameters that have a gradient."""
        world = self.world_size()
        with_grad = [p for p in self.params if p.grad is not None]
        if world == 1 or not with_grad:
            return
        n = sum(p.grad.numel() for p in with_grad)
        ref = with_grad[0].grad
        if self._flat is None or self._flat.numel() != n or self._flat.device != ref.device:
            self._flat = torch.empty(n, dtype=torch.float32, device=ref.device)
        off = 0
        for p in with_grad:
            k = p.grad.numel()
            self._flat[off:off + k].copy_(p.grad.reshape(-1))
            off += k
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        if self.average:
            self._flat.div_(world)
        off = 0
        for p in with_grad:
            k = p.grad.numel()
            p.grad.copy_(self._flat[off:off + k].view_as(p.grad))
            off += k

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def distribute(model: torch.nn.Module, group=None, src: int = 0) -> GradientAllReduce:
    """Replicate `model` from rank `src` and make its backward all-reduce the gradients (see the module docstring)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t, src=src, group=group)
    return GradientAllReduce(model, group)


def shard(batch: torch.Tensor, group=None) -> torch.Tensor:
    """This rank's contiguous slice of a global batch (the last ranks get the remainder)."""
    if not (dist.is_available() and dist.is_initialized()):
        return batch
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (batch.shape[0] + world - 1) // world
    return batch[rank * per:(rank + 1) * per]
