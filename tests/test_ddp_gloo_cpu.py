"""world_size-2 gloo test of the data-parallel semantics the engine relies on: sharding the batch, summing
the flat gradient buffer and scaling by 1/world reproduces the single-process gradient when mean-type losses are
averaged and the SUM-type HSIC term is multiplied by world (engine.TrainEngine.loss)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 5))


def _loss(model, x, y, beta, world):
    out = model(x)
    ce = torch.nn.functional.cross_entropy(out, y)          # mean over the (local) batch
    hsic_like = (out ** 2).sum()                            # SUM over the batch, like utils.py:28-31
    return ce + beta * world * hsic_like


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(123)
    X, Y = torch.randn(8, 12), torch.randint(0, 5, (8,))
    model = _model()
    params = list(model.parameters())
    # flat buffers exactly as engine.TrainEngine builds them (16-byte aligned slices, grads are views)
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]
    flat, gflat = torch.zeros(sum(sizes)), torch.zeros(sum(sizes))
    off = 0
    for p, n in zip(params, sizes):
        sl = flat[off:off + p.numel()].view_as(p); sl.copy_(p.data); p.data = sl
        p.grad = gflat[off:off + p.numel()].view_as(p); off += n
    dist.broadcast(flat, src=0)
    shard = slice(rank * 4, (rank + 1) * 4)
    _loss(model, X[shard], Y[shard], 1e-2, world).backward()
    dist.all_reduce(gflat, op=dist.ReduceOp.SUM)
    gflat *= 1.0 / world
    if rank == 0:
        ref = _model()
        _loss(ref, X, Y, 1e-2, 1).backward()
        got = torch.cat([p.grad.reshape(-1) for p in params])
        exp = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        ret["err"] = float((got - exp).norm() / exp.norm())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradient_equals_global_batch_gradient():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["err"] < 1e-6, ret["err"]
