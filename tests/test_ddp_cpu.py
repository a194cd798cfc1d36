"""CPU, world_size 2 over gloo: the data-parallel plumbing of the training step (flat gradient arena, per-bucket
asynchronous all-reduce fired by gradient-ready hooks, unused-parameter handling, fused n_pos all-reduce)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from embodiedscan_b200.dense_heads import FCAF3DHeadRotMat
    from embodiedscan_b200.engine import DataParallelReducer, FlatArena, broadcast_parameters
    torch.manual_seed(rank)          # different init per rank: broadcast must make them equal
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    arena = FlatArena(net, bucket_bytes=256)
    broadcast_parameters(arena)
    red = DataParallelReducer(arena)
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 8)
    out = net[2](net[1](net[0](x)))            # net[3] unused this step
    out.pow(2).sum().backward()
    red.finish()
    # fused scalar reduce of the head
    n_pos = FCAF3DHeadRotMat._reduce_mean(type('H', (), {'process_group': None})(), torch.tensor([3.0 + rank, 1.0]))
    q.put((rank, arena.flat.clone(), arena.grad.clone(), len(arena.buckets), n_pos))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(30)
    (_, f0, g0, nb, n0), (_, f1, g1, _, n1) = res
    assert nb > 1, 'the test must exercise several buckets'
    assert torch.equal(f0, f1), 'parameters equal after broadcast'
    assert torch.equal(g0, g1), 'every rank holds the same summed gradient'
    assert float(g0.abs().sum()) > 0
    assert torch.allclose(n0, torch.tensor([3.5, 1.0])) and torch.equal(n0, n1)

    # single-process reference: sum of the two per-rank gradients
    sys.path.insert(0, ROOT)
    from embodiedscan_b200.engine import FlatArena
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    arena = FlatArena(net, bucket_bytes=256)
    assert torch.equal(arena.flat, f0)
    for r in range(2):
        torch.manual_seed(100 + r)
        x = torch.randn(5, 8)
        net[2](net[1](net[0](x))).pow(2).sum().backward()
    assert torch.allclose(arena.grad, g0, atol=1e-6)
