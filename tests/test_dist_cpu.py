"""CPU, gloo, world size 2: the N>1 plumbing of bench.py (replicas: round-robin image sharding, barrier, max-over-ranks
timing).  No GPU, no kernels."""
import os
import socket
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import relnet_b200
    from relnet_b200 import replicas
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = replicas.shard_images(7, rank, world)
    ms = 10.0 + 5.0 * rank                       # rank 1 is the slow one
    dist.barrier()
    ms_max = replicas.max_over_ranks(ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, ms_max, gathered, replicas.images_per_second(7, ms_max)))
    dist.barrier()
    dist.destroy_process_group()


def test_replica_sharding_and_max_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, t0, g0, v0), (r1, m1, t1, g1, v1) = res
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5]
    assert sorted(m0 + m1) == list(range(7))                 # every image exactly once
    assert t0 == t1 == 15.0                                  # max over ranks, identical on every rank
    assert g0 == g1 == [m0, m1]
    assert abs(v0 - 7 / 0.015) < 1e-6


def test_single_process_identity():
    import relnet_b200
    from relnet_b200 import replicas
    assert replicas.max_over_ranks(3.5) == 3.5
    assert replicas.shard_images(3, 0, 1) == [0, 1, 2]


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import relnet_b200
    from relnet_b200 import replicas
    dist.init_process_group('gloo', rank=rank, world_size=world)
    shapes = {'query_1_weight': (8, 8), 'query_1_bias': (8,), 'conv1_weight': (4, 3, 7, 7), 'pair_pos_fc1_1_weight': (16, 64)}
    b = replicas.GradientBucket(shapes, device='cpu', frozen=('conv1',))
    for i, n in enumerate(b.names):                      # each rank "computes" its own gradient in place
        b.views[n].fill_(float(rank + 1) * (i + 1))
    b.allreduce()
    params = {n: torch.ones(shapes[n]) for n in b.names}
    replicas.sgd_step(params, b, lr=0.1, momentum=0.0, wd=0.0)
    q.put((rank, b.names, [float(b.views[n].flatten()[0]) for n in b.names], [float(params[n].flatten()[0]) for n in b.names],
           b.flat.numel()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_sum():
    """training exchange: one SUM allreduce over a flat bucket (reference: kvstore device, rescale_grad 1.0); frozen params
    have no window; the update sees the summed gradient on every rank"""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, names, g, pvals, numel in res:
        assert names == ['query_1_weight', 'query_1_bias', 'pair_pos_fc1_1_weight']          # conv1 is frozen
        assert g == [3.0, 6.0, 9.0]                                                          # (1 + 2) * (i + 1)
        assert all(abs(p - (1.0 - 0.1 * gg)) < 1e-6 for p, gg in zip(pvals, g))
        assert numel % 64 == 0
