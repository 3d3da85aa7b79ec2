"""Parameter initialisers (interface of deeprob/torch/initializers.py:7-31); cold, host-side."""
import torch
from torch import distributions


def dirichlet_(tensor: torch.Tensor, alpha: float = 1.0, log_space: bool = True, dim: int = -1):
    """Fill `tensor` with symmetric-Dirichlet(alpha) draws along `dim` (their logs if `log_space`)."""
    nd = tensor.dim()
    if nd == 0:
        raise ValueError("Singleton tensors are not valid")
    if dim not in range(-nd, nd - 1):  # same accepted range as the reference (:21-25)
        raise IndexError(
            "Dimension out of range (expected to be in range of [{}, {}], but got {})".format(-nd, nd - 1, dim)
        )
    axis = dim % nd
    with torch.no_grad():
        batch = [n for i, n in enumerate(tensor.shape) if i != axis]
        draws = distributions.Dirichlet(torch.full([tensor.shape[axis]], alpha)).sample(batch)
        if log_space:
            draws = draws.log()
        tensor.copy_(draws.transpose(axis, -1))
