"""N > 1 host logic on CPU: world size 2 over gloo (127.0.0.1).  Checks that the bucketed hook-based gradient
exchange and the flat variant both produce the mean of the per-rank gradients, including parameters that got
no gradient on one step and several backward passes per step (the reference trainers run 2-3 forwards per
backward)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pwa_b200  # noqa: F401  (host library must load without a GPU)
    from pwa_b200.ddp import BucketedGradSync, allreduce_gradients_flat
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    unused = torch.nn.Parameter(torch.ones(3))
    params = list(model.parameters()) + [unused]
    sync = BucketedGradSync(params, bucket_bytes=300)          # tiny buckets -> several of them
    assert len(sync.buckets) > 2
    ok = True
    for step in range(2):
        for p in params:
            p.grad = None
        g = torch.Generator().manual_seed(100 * step + rank)
        x = torch.randn(5, 8, generator=g)
        model(x).square().sum().backward()
        local = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in params]
        sync.finish()
        # expected: mean over ranks of the local gradients
        for p, l in zip(params, local):
            t = l.clone()
            dist.all_reduce(t)
            ok = ok and torch.allclose(p.grad, t / world, atol=1e-6)
    sync.remove()
    for p in params:
        p.grad = None
    g = torch.Generator().manual_seed(7 + rank)
    model(torch.randn(5, 8, generator=g)).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    allreduce_gradients_flat(list(model.parameters()))
    for p, l in zip(model.parameters(), local):
        t = l.clone()
        dist.all_reduce(t)
        ok = ok and torch.allclose(p.grad, t / world, atol=1e-6)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_bucketed_grad_sync_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
