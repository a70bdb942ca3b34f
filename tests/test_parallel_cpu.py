"""World-size-2 gloo tests of the data-parallel host logic: frame sharding, the flat gradient all-reduce and the
bucketed reducer that overlaps the reduction with the backward (hooks, bucket readiness, unused parameters)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pillarnext_b200.parallel import BucketedGradAllReduce, FlatGradAllReduce, shard_frames
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    data = torch.randn(6, 8, generator=torch.Generator().manual_seed(1))
    mine = shard_frames(6, rank, world)
    loss = net(data[mine]).pow(2).sum() / 6 * world        # mean over the global batch once averaged over ranks
    loss.backward()
    FlatGradAllReduce(net.parameters())()
    grads = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    # the bucketed, hook-driven reducer: two buckets in backward order + one parameter that never gets a gradient
    net.zero_grad(set_to_none=True)
    unused = torch.nn.Parameter(torch.ones(3))
    red = BucketedGradAllReduce([list(net[2].parameters()), list(net[0].parameters()) + [unused]])
    for _ in range(2):                                     # twice: the bookkeeping resets between steps
        net.zero_grad(set_to_none=True)
        loss = net(data[mine]).pow(2).sum() / 6 * world
        loss.backward()
        red.finish()
    g2 = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(g2, grads, atol=1e-6) and unused.grad is not None and float(unused.grad.abs().sum()) == 0.0
    red.remove()
    q.put((rank, mine, grads.tolist()))   # plain floats: a tensor would travel as a shared fd that dies with the worker
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res = sorted((r, m, torch.tensor(g)) for r, m, g in res)
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3, 5]
    assert torch.allclose(res[0][2], res[1][2])
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    data = torch.randn(6, 8, generator=torch.Generator().manual_seed(1))
    (net(data).pow(2).sum() / 6).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(res[0][2], ref, atol=1e-6)
