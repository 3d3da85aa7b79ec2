"""Batch EM for the tensorised RAT-SPN (extension: the reference has EM only for its NumPy node-graph SPNs,
deeprob/spn/learning/em.py:18-113).  One step = E-step statistics from the CUDA backward pass
(`RatSpn.em_statistics`, csrc/ratspn_bwd.cu), ONE all-reduce of the flat statistics vector when the batch is
sharded over ranks, and the node-level M-step formulas applied to the parameter tensors:

  sum / root weights  n = counts + eps32 ; w' = n / sum n ; w <- (1-eta) w + eta w'          (structure/node.py:100-111)
  Gaussian leaves     mu' = S1 / (S0 + eps32) ; sd' = max(sqrt((S2 - 2 mu' S1 + mu'^2 S0) / (S0 + eps32)), 1e-5)
                      mu <- (1-eta) mu + eta mu' ; sd likewise                               (structure/leaf.py:536-545)
  Bernoulli leaves    p' = (S1 + a) / (S0 + 2a), a = eps16 ; p <- (1-eta) p + eta p'          (structure/leaf.py:167-174)

`counts` are the posterior counts w * sum_b exp(child_ll - root_ll + log-grad) (em.py:99-102) and S0/S1/S2 the
posterior-weighted leaf moments (em.py:105-107) over the observed (non-NaN) entries.
The M-step is plain tensor arithmetic on parameter-sized tensors and runs on whatever device they live on.
"""
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

_EPS32 = float(np.finfo(np.float32).eps)
_EPS16 = float(np.finfo(np.float16).eps)


def _stat_tensors(stats: Dict) -> List[torch.Tensor]:
    out = list(stats["sum_counts"]) + [stats["root_counts"], stats["s0"], stats["s1"]]
    if stats.get("s2") is not None:
        out.append(stats["s2"])
    return out


def pack_statistics(stats: Dict, n_samples: int) -> torch.Tensor:
    """[sum LL, N, sum-level counts..., root counts, S0, S1, (S2)] as one flat fp32 vector (one all-reduce)."""
    parts = _stat_tensors(stats)
    head = torch.stack([stats["ll"].sum().float().reshape(()),
                        torch.tensor(float(n_samples), device=parts[0].device)])
    return torch.cat([head] + [p.reshape(-1).float() for p in parts])


def unpack_statistics(flat: torch.Tensor, like: Dict) -> Dict:
    parts = _stat_tensors(like)
    out, off = [], 2
    for p in parts:
        out.append(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    n_sum = len(like["sum_counts"])
    res = {"ll_sum": flat[0], "n": flat[1], "sum_counts": out[:n_sum], "root_counts": out[n_sum], "s0": out[n_sum + 1],
           "s1": out[n_sum + 2], "s2": out[n_sum + 3] if like.get("s2") is not None else None}
    return res


def all_reduce_statistics(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the flat statistics over the ranks (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def shard_bounds(n_samples: int, rank: int, world_size: int):
    """Contiguous batch shard [lo, hi) of `rank` (parameters are replicated; the forward needs no exchange)."""
    per = (n_samples + world_size - 1) // world_size
    lo = min(n_samples, rank * per)
    return lo, min(n_samples, lo + per)


@torch.no_grad()
def m_step(model, stats: Dict, step_size: float = 0.5) -> None:
    """Update the parameters of a RatSpn in place from (globally reduced) statistics."""
    eta = float(step_size)

    def update_logits(weight: torch.Tensor, counts: torch.Tensor):
        w_old = torch.softmax(weight, dim=-1)
        n = counts + _EPS32
        w_new = (1.0 - eta) * w_old + eta * n / n.sum(dim=-1, keepdim=True)
        weight.copy_(torch.log(w_new))

    sums = [layer for layer in model.layers if hasattr(layer, "weight") and isinstance(layer.weight, torch.nn.Parameter)]
    for layer, counts in zip(sums, stats["sum_counts"]):
        update_logits(layer.weight, counts)
    update_logits(model.root_layer.weight, stats["root_counts"])

    base = model.base_layer
    s0, s1 = stats["s0"], stats["s1"]
    live = torch.arange(base.dimension, device=s0.device)[None, None, :] < base._region_len.to(s0.device)[:, None, None]
    if hasattr(base, "logits"):
        p_old = torch.sigmoid(base.logits)
        p_new = (1.0 - eta) * p_old + eta * (s1 + _EPS16) / (s0 + 2.0 * _EPS16)
        p_new = p_new.clamp(1e-7, 1.0 - 1e-7)
        base.logits.copy_(torch.where(live, torch.log(p_new) - torch.log1p(-p_new), base.logits))
    else:
        s2 = stats.get("s2")
        total = s0 + _EPS32
        mean = s1 / total
        base.loc.copy_(torch.where(live, (1.0 - eta) * base.loc + eta * mean, base.loc))
        # A frozen unit scale (the default optimize_scale=False, layers/ratspn.py:198-207) stays frozen: the E-step
        # then returns no second moment (s2 is None) and only the means are re-estimated.
        if s2 is not None and base.scale.requires_grad:
            var = (s2 - 2.0 * mean * s1 + mean * mean * s0) / total
            std = torch.sqrt(var.clamp_min(0.0)).clamp_min(1e-5)
            base.scale.copy_(torch.where(live, (1.0 - eta) * base.scale + eta * std, base.scale))


def em_step(model, x: torch.Tensor, step_size: float = 0.5, group=None, timing: Optional[Dict] = None) -> float:
    """One (optionally batch-sharded) EM step on this rank's shard `x`; returns the global mean log-likelihood.
    `timing` (optional dict) collects a pair of CUDA events around the all-reduce of every step under
    "allreduce" and the size of the exchanged vector under "bytes" (bench.py reports them)."""
    stats = model.em_statistics(x)
    flat = pack_statistics(stats, x.shape[0])
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flat = all_reduce_statistics(flat, group)
        e1.record()
        timing.setdefault("allreduce", []).append((e0, e1))
        timing["bytes"] = flat.numel() * 4
    else:
        flat = all_reduce_statistics(flat, group)
    glob = unpack_statistics(flat, stats)
    m_step(model, glob, step_size)
    return float(glob["ll_sum"] / glob["n"])
