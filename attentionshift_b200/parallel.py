"""Multi-GPU plumbing for the hot path: one process per GPU, the image batch is the only sharded axis,
there is NO data-path collective -- every stage from Attention.forward to pseudo_gt_masks is per image
(reference: batch is an outer python loop, RH:2267 / RH:2332; launch line run_train.py:9).
torch.distributed is used for rendezvous, a barrier and the max-over-ranks reduction of the device timings only."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl' and device is not None:
            kw['device_id'] = device
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_images(n_images, rank, world):
    """Contiguous, balanced slice of the global image batch for this rank (remainder spread over the first ranks);
    the slices of all ranks partition range(n_images)."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values, device=None):
    """Element-wise maximum of a list of floats over all ranks (device timings: the slowest rank defines the step)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sum_over_ranks(values, device=None):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()
