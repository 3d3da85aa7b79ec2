"""Host-resident batches: pipelined host->device copy, log-likelihood kernels and device->host copy.

`log_prob_host` is the end-to-end entry point for data that lives in (pinned) host memory, e.g. a
DataLoader batch as in deeprob/torch/routines.py:158-166 where the reference does
`inputs.to(device)` -> `model(inputs)` -> `.cpu()`.  The batch is cut in chunks that alternate over
two CUDA streams so the PCIe copies overlap the kernels.
"""
from typing import Optional

import torch


def log_prob_host(model, x_host: torch.Tensor, chunk: int = 8192, out_host: Optional[torch.Tensor] = None,
                  device=None, wait: bool = True) -> torch.Tensor:
    """x_host (B, ...) float32 CPU tensor (pinned for async copies) -> (B, C) CPU tensor of log-likelihoods.

    One stream issues every host->device copy back to back into a ring of three chunk buffers (the copy engine never
    waits for a kernel), a second stream runs the kernels and the device->host copy of each chunk as its copy lands.
    The path is PCIe bound (205 MB per 65 536 x 784 batch at ~55 GB/s = 3.7 ms against 0.55 ms of kernels), so what is
    left to hide is the last chunk's kernels: the default chunk of 8 192 rows keeps that tail below 0.1 ms.

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
    nbuf = 3
    cache = model.__dict__.setdefault("_host_pipeline", {})
    key = (str(dev), chunk, tuple(x_host.shape[1:]))
    if key not in cache:
        cache[key] = {
            "copy": torch.cuda.Stream(dev), "compute": torch.cuda.Stream(dev),
            "bufs": [torch.empty((chunk, *x_host.shape[1:]), dtype=torch.float32, device=dev) for _ in range(nbuf)],
            "landed": [torch.cuda.Event() for _ in range(nbuf)],     # host->device copy into buffer i complete
            "free": [torch.cuda.Event() for _ in range(nbuf)],       # kernels that read buffer i complete
            "done": torch.cuda.Event(),
        }
    c = cache[key]
    s_copy, s_comp, bufs, landed, free = c["copy"], c["compute"], c["bufs"], c["landed"], c["free"]
    cur = torch.cuda.current_stream(dev)
    s_copy.wait_stream(cur)
    s_comp.wait_stream(cur)
    with torch.no_grad():
        for i, start in enumerate(range(0, n, chunk)):
            m = min(chunk, n - start)
            b = i % nbuf
            with torch.cuda.stream(s_copy):
                if i >= nbuf:
                    s_copy.wait_event(free[b])
                xb = bufs[b][:m]
                xb.copy_(x_host[start:start + m], non_blocking=True)
                landed[b].record(s_copy)
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(landed[b])
                y = model(xb)
                free[b].record(s_comp)
                if out_host is None:
                    out_host = torch.empty((n, *y.shape[1:]), dtype=torch.float32, pin_memory=True)
                out_host[start:start + m].copy_(y, non_blocking=True)
    c["done"].record(s_comp)
    cur.wait_event(c["done"])
    # the ring buffers are reused by the next call: its copies are issued on s_copy after cur -> after `done`
    if wait:
        c["done"].synchronize()         # host-side: the pinned result is complete when we return
    return out_host
