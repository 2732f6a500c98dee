"""One-process-per-GPU plumbing for the test-time path: images are independent, so ranks are replicas -- image i of the
global batch goes to rank i % world (the reference gives each executor its own slice: core/loader.py:561-588,
DataParallelExecutorGroup.py:336-360) and the only cross-rank traffic is the timing barrier / max-reduce of bench.py.
(The gradient allreduce of training arrives with the backward passes)"""
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
