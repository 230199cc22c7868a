"""Multi-GPU plumbing for the inference hot path: one process per GPU, the global batch sharded by rank, NO
collective on the data path (every image pair is independent end to end: eval-mode BN, per-image NMS --
reference utils/utils_bbox.py:100).  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to
collect timings and detections after the step.

The reference's own multi-GPU inference is nn.DataParallel (reference yolo_mul.py:60-62), whose gather would
concatenate the non-batched `anchors (2,A)` / `strides (1,A)` outputs along dim 0 and break decode_box; this
module replaces it.
"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world):
    """[lo, hi) of the image pairs rank `rank` processes: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(int(global_batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(tensor, rank=None, world=None):
    """This rank's slice of a batch-major tensor."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(tensor.shape[0], rank, world)
    return tensor[lo:hi]


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (multi-GPU timings are reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_detections(local_results):
    """local_results: this rank's list (one entry per image: None or ndarray (n,6)).  Returns the global list in
    batch order on every rank.  Variable-length, so it goes through all_gather_object (off the timed path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_results)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, list(local_results))
    return [r for part in parts for r in part]
