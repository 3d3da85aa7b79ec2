"""CPU (gloo, world_size 2): the host side of the batch-sharded EM step -- statistics packing, the single
all-reduce and the M-step -- gives every rank the parameters a single process gets from the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deeprob_kit_b200.spn import em
from deeprob_kit_b200.spn.models import BernoulliRatSpn, GaussianRatSpn


def _model(kind):
    torch.manual_seed(3)
    if kind == "gaussian":
        return GaussianRatSpn(19, rg_depth=2, rg_repetitions=3, rg_batch=4, rg_sum=3, random_state=42, optimize_scale=True)
    return BernoulliRatSpn(19, rg_depth=2, rg_repetitions=3, rg_batch=4, rg_sum=3, random_state=42)


def _fake_stats(model, seed, n):
    """Synthetic but structurally valid E-step statistics of a shard of `n` samples."""
    rng = np.random.RandomState(seed)
    t = lambda shape: torch.from_numpy(rng.gamma(2.0, 1.0, size=tuple(shape)).astype(np.float32))  # noqa: E731
    base = model.base_layer
    shape = (base.in_regions, base.out_channels, base.dimension)
    s0 = t(shape) * n / 4
    mean = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
    stats = {"ll": torch.from_numpy(rng.standard_normal((n, 1)).astype(np.float32)) - 20.0,
             "sum_counts": [t(l.weight.shape) for l in model._sum_layers()], "root_counts": t(model.root_layer.weight.shape),
             "s0": s0, "s1": s0 * mean, "s2": (s0 * (mean * mean + 0.3)) if hasattr(base, "scale") else None}
    return stats


def _worker(rank, world, port, kind, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model(kind)
        n = 40 + 8 * rank
        stats = _fake_stats(model, 100 + rank, n)
        flat = em.all_reduce_statistics(em.pack_statistics(stats, n))
        glob = em.unpack_statistics(flat, stats)
        em.m_step(model, glob, 0.5)
        res = {k: v.clone() for k, v in model.state_dict().items()}
        res["__mean_ll"] = torch.tensor(float(glob["ll_sum"] / glob["n"]))
        out[rank] = res
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("kind", ["gaussian", "bernoulli"])
def test_sharded_em_step_matches_single_process(kind):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), kind, out), nprocs=world, join=True)
        states = [out[r] for r in range(world)]
    # single process, whole batch = the sum of both shards' statistics
    model = _model(kind)
    shards = [_fake_stats(model, 100 + r, 40 + 8 * r) for r in range(world)]
    flat = sum(em.pack_statistics(s, 40 + 8 * r) for r, s in enumerate(shards))
    em.m_step(model, em.unpack_statistics(flat, shards[0]), 0.5)
    ref = model.state_dict()
    for r in range(world):
        for k, v in ref.items():
            assert torch.allclose(states[r][k].float(), v.float(), rtol=1e-5, atol=1e-6), (r, k)
        assert abs(float(states[r]["__mean_ll"]) - float(flat[0] / flat[1])) < 1e-4
    # M-step sanity: mixtures stay normalised, scales positive
    for layer in model._sum_layers():
        assert torch.allclose(layer.weight.exp().sum(-1), torch.ones_like(layer.weight[..., 0]), atol=1e-5)
    assert torch.allclose(model.root_layer.weight.exp().sum(-1), torch.ones(1), atol=1e-5)
    if kind == "gaussian":
        assert bool((model.base_layer.scale > 0).all())


def test_shard_bounds_cover_the_batch():
    for n, w in ((10, 3), (65536, 8), (5, 8), (0, 2)):
        spans = [em.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
