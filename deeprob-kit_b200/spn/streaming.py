"""Host-resident batches: pipelined host->device copy, log-likelihood kernels and device->host copy.

`log_prob_host` is the end-to-end entry point for data that lives in (pinned) host memory, e.g. a
DataLoader batch as in deeprob/torch/routines.py:158-166 where the reference does
`inputs.to(device)` -> `model(inputs)` -> `.cpu()`.  The batch is cut in chunks that alternate over
two CUDA streams so the PCIe copies overlap the kernels.
"""
from typing import Optional

import torch


def log_prob_host(model, x_host: torch.Tensor, chunk: int = 16384, out_host: Optional[torch.Tensor] = None,
                  device=None, wait: bool = True) -> torch.Tensor:
    """x_host (B, ...) float32 CPU tensor (pinned for async copies) -> (B, C) CPU tensor of log-likelihoods.

    With `wait=True` (default) the call returns only after the last device->host copy has landed, so the result
    can be read immediately, like the `.cpu()` of the reference loop.  `wait=False` returns as soon as the work is
    queued; the caller's current CUDA stream is ordered after the copies (synchronize it, or the device, before
    touching the returned pinned tensor)."""
    if x_host.is_cuda:
        raise ValueError("log_prob_host expects a host tensor; call model(x) for device tensors")
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("the model must live on a CUDA device (no CPU path)")
    n = x_host.shape[0]
    if n == 0:
        with torch.no_grad():
            y = model(torch.empty((0, *x_host.shape[1:]), dtype=torch.float32, device=dev))
        return out_host[:0] if out_host is not None else torch.empty((0, *y.shape[1:]), dtype=torch.float32)
    cache = model.__dict__.setdefault("_host_pipeline", {})
    key = (str(dev), chunk, tuple(x_host.shape[1:]))
    if key not in cache:
        cache[key] = {
            "streams": [torch.cuda.Stream(dev), torch.cuda.Stream(dev)],
            "bufs": [torch.empty((chunk, *x_host.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)],
            "done": [torch.cuda.Event(), torch.cuda.Event()],
        }
    streams, bufs, done = cache[key]["streams"], cache[key]["bufs"], cache[key]["done"]
    cur = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(cur)
    used = set()
    with torch.no_grad():
        for i, start in enumerate(range(0, n, chunk)):
            m = min(chunk, n - start)
            s = streams[i % 2]
            used.add(i % 2)
            with torch.cuda.stream(s):
                xb = bufs[i % 2][:m]
                xb.copy_(x_host[start:start + m], non_blocking=True)
                y = model(xb)
                if out_host is None:
                    out_host = torch.empty((n, *y.shape[1:]), dtype=torch.float32, pin_memory=True)
                out_host[start:start + m].copy_(y, non_blocking=True)
    for i in used:
        done[i].record(streams[i])
        cur.wait_event(done[i])
    if wait:
        for i in used:
            done[i].synchronize()       # host-side: the pinned result is complete when we return
    return out_host
