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


def _train_block_worker(rank, world, port, q):
    """bench.train_block on gloo with a stand-in training step whose graph capture FAILS ON RANK 1 ONLY: every rank must issue
    the same sequence of collectives (no deadlock), take the same job-wide decision (eager), and the allreduce of the real
    GradientBucket must equal the sum of the ranks' buckets."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    import time
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import relnet_b200
    from relnet_b200 import replicas
    import bench

    class Ev(object):
        def __init__(self, enable_timing=True): self.t = None
        def record(self): self.t = time.time()
        def elapsed_time(self, o): return (o.t - self.t) * 1e3
    torch.cuda.Event = Ev
    torch.cuda.synchronize = lambda *a, **k: None
    dist.init_process_group('gloo', rank=rank, world_size=world)

    class TS(object):
        def __init__(self):
            self.bucket = replicas.GradientBucket({'w': (64, 64), 'b': (64,)}, device='cpu')
            self.graph, self.last, self.allreduces = None, {'rois': 308, 'cls': 1.0}, 0
        def forward_backward(self, im, info):
            self.bucket.flat.add_(float(rank + 1))
        def step(self, images, info):
            self.bucket.zero_()
            for im in images:
                self.forward_backward(im, info)
            a, b = Ev(), Ev()
            a.record(); self.bucket.allreduce(); self.allreduces += 1; b.record()
            return (a, b)
        def capture(self, images, info):
            if rank == 1:
                raise RuntimeError('capture failed here')
            self.graph = 'captured'
    ts = TS()
    reported = []
    r = bench.train_block(types.SimpleNamespace(steps=6), ts, [torch.zeros(1)], torch.zeros(1), 'cpu', world, rank, True, 'mode',
                          report=reported.append)
    q.put((rank, r['launch'], ts.graph, ts.allreduces, r['reduced_vs_sum_of_ranks_rel'], r['collectives_per_step'], len(reported)))
    dist.barrier()
    dist.destroy_process_group()


def test_train_block_collectives_stay_symmetric_when_capture_fails_on_one_rank():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_block_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, l0, g0, n0, c0, k0, rep0), (r1, l1, g1, n1, c1, k1, rep1) = res
    assert g0 is None and g1 is None                      # one decision for the whole job: nobody replays a graph
    assert 'another rank' in l0 and 'this rank' in l1 and l0.startswith('eager') and l1.startswith('eager')
    assert n0 == n1                                       # same number of allreduces on both ranks
    assert c0 < 1e-12 and c1 < 1e-12 and k0 == k1 == 1    # allreduced bucket == sum of the ranks' buckets
    assert rep0 == rep1 == 1                              # the eager result was reported before the capture attempt
