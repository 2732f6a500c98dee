"""One-process-per-GPU plumbing for the test-time path: images are independent, so ranks are replicas -- image i of the
global batch goes to rank i % world (the reference gives each executor its own slice: core/loader.py:561-588,
DataParallelExecutorGroup.py:336-360) and the only cross-rank traffic is the timing barrier / max-reduce of bench.py.
Training adds exactly one exchange per step: a SUM allreduce of all trainable gradients (reference: kvstore 'device',
rescale_grad = 1.0, train_end2end.py:167; frozen parameters excluded, FIXED_PARAMS in the yaml :23-29).  GradientBucket
lays every gradient out in ONE flat fp32 buffer and hands out views, so the backward kernels (rn_relation_bwd,
rn_learn_nms_bwd, ... -- they write to whatever device pointers they are given) produce the bucket in place and the step
ends with a single ncclAllReduce over NVLink, no pack / unpack copies."""
import torch
import torch.distributed as dist


def shard_images(num_images, rank, world):
    """indices of the global batch this rank processes (round-robin, like the reference's per-GPU slices)"""
    return list(range(rank, num_images, world))


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (device tensor for NCCL, CPU tensor for gloo); identity without a group"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    backend = dist.get_backend()
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if backend == 'nccl' else 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def images_per_second(total_images, ms_max):
    """whole-job throughput: all images of all ranks over the slowest rank's time"""
    return total_images / (ms_max / 1e3)


class GradientBucket(object):
    """One flat fp32 buffer holding every trainable gradient; ``views[name]`` are the per-parameter windows.

    shapes: ordered mapping name -> shape.  frozen: names (or name prefixes, like the reference's FIXED_PARAMS) to leave
    out of the bucket.  ``allreduce()`` sums the whole bucket over the ranks with one collective."""

    def __init__(self, shapes, device='cuda', frozen=()):
        self.names = [n for n in shapes if not any(n.startswith(f) for f in frozen)]
        sizes = [int(torch.Size(shapes[n]).numel()) for n in self.names]
        # 256-byte aligned windows: the kernels' vector loads/stores and cuBLAS like aligned bases
        offs, total = [], 0
        for sz in sizes:
            offs.append(total)
            total += (sz + 63) // 64 * 64
        self.flat = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        self.views = {n: self.flat[o:o + sz].view(torch.Size(shapes[n])) for n, o, sz in zip(self.names, offs, sizes)}

    def zero_(self):
        self.flat.zero_()

    def allreduce(self, async_op=False):
        """SUM over ranks (the reference's semantic); no-op without a process group.  Returns the work handle if async."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)


def sgd_step(params, bucket, lr, momentum=0.9, wd=0.0001, state=None):
    """The reference optimizer update (mx 'sgd': momentum 0.9, wd 1e-4, rescale_grad 1.0, train_end2end.py:163-170) applied
    to the parameters that own a window in the bucket, as fused torch foreach ops on the flat views."""
    state = state if state is not None else {}
    ps = [params[n] for n in bucket.names]
    gs = [bucket.views[n] for n in bucket.names]
    ms = [state.setdefault(n, torch.zeros_like(params[n])) for n in bucket.names]
    torch._foreach_mul_(ms, momentum)
    torch._foreach_add_(ms, [g + wd * p for g, p in zip(gs, ps)], alpha=-lr)
    torch._foreach_add_(ps, ms)
    return state
