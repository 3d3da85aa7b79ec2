"""Data-parallel training glue (deeprob_kit_b200.torch.parallel).  CPU, gloo, world_size 2: the gradient hooks turn
an unmodified single-rank training loop into a data-parallel one -- replicas stay identical and equal a single
process trained on the whole batch.  GPU (needs 2 devices): the same through NCCL with a RAT-SPN and the fused loss."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deeprob_kit_b200.torch import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _net():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(24, 6, generator=g), torch.randn(24, 1, generator=g)


def _train(model, x, y, steps=4):
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    with torch.enable_grad():
        for _ in range(steps):                       # the reference loop: routines.py:158-166
            opt.zero_grad()
            loss = ((model(x) - y) ** 2).mean()
            loss.backward()
            opt.step()
    return model


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _net()
        if rank == 1:                                 # replicas start different: distribute() must broadcast rank 0
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        hooks = parallel.distribute(model)
        x, y = _data()
        _train(model, parallel.shard(x), parallel.shard(y))
        out[rank] = ({k: v.clone() for k, v in model.state_dict().items()}, hooks.calls)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_matches_single_process():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = [out[r] for r in range(world)]
    x, y = _data()
    ref = _train(_net(), x, y).state_dict()          # equal shards: mean of shard means == mean over the batch
    for r in range(world):
        state, calls = res[r]
        assert calls == 4                             # exactly one collective per step
        for k, v in ref.items():
            assert torch.allclose(state[k], v, rtol=1e-5, atol=1e-6), (r, k)


def test_single_rank_is_a_no_op():
    model = _net()
    hooks = parallel.distribute(model)
    x, y = _data()
    _train(model, x, y)
    assert hooks.calls == 0
    ref = _train(_net(), x, y).state_dict()
    for k, v in model.state_dict().items():
        assert torch.equal(v, ref[k])
    hooks.remove()


def _gpu_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from deeprob_kit_b200.spn.models import GaussianRatSpn
        torch.manual_seed(3 + rank)                   # different initialisations: rank 0 wins
        model = GaussianRatSpn(16, out_classes=3, rg_depth=2, rg_repetitions=3, rg_batch=4, rg_sum=2, random_state=42,
                               optimize_scale=True).cuda().train()
        hooks = parallel.distribute(model)
        g = torch.Generator().manual_seed(0)
        x, y = torch.randn(64, 16, generator=g), torch.randint(0, 3, (64,), generator=g)
        xs, ys = parallel.shard(x).cuda(), parallel.shard(y).cuda()
        opt = torch.optim.Adam(model.parameters(), lr=1e-2)
        losses = []
        with torch.enable_grad():
            for _ in range(5):
                opt.zero_grad()
                loss = model.loss(model(xs), ys)
                loss.backward()
                opt.step()
                model.apply_constraints()
                losses.append(loss.detach())
        out[rank] = ({k: v.cpu() for k, v in model.state_dict().items()}, hooks.calls, [float(l) for l in losses])
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_ratspn_data_parallel_training_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gpu_worker, args=(2, _free_port(), out), nprocs=2, join=True)
        a, b = out[0], out[1]
    assert a[1] == 5 and b[1] == 5
    for k, v in a[0].items():
        assert torch.equal(v, b[0][k]), k             # replicas bit-identical after 5 steps


@pytest.mark.gpu
def test_fused_loss_matches_torch():
    from deeprob_kit_b200.spn._engine import nll_loss
    g = torch.Generator().manual_seed(2)
    for c in (1, 4):
        ll = (torch.randn(300, c, generator=g) * 3 - 40).cuda()
        y = torch.randint(0, c, (300,), generator=g).cuda()
        with torch.enable_grad():
            a = ll.clone().requires_grad_(True)
            la = nll_loss(a, y if c > 1 else None)
            (la * 2.5).backward()
            b = ll.clone().requires_grad_(True)
            lb = -b.mean() if c == 1 else torch.nn.functional.nll_loss(torch.log_softmax(b, dim=1), y)
            (lb * 2.5).backward()
        assert abs(float(la) - float(lb)) < 1e-4 * abs(float(lb))
        assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-7)
