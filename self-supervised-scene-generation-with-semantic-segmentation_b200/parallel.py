"""One-process-per-GPU sharding of the raycast path.

The units of work -- (chunk, view) renderings and their backward -- are independent: a rank reads only its own chunks'
voxels and writes only its own images and voxel gradients.  There is therefore no collective on the data path; ranks
only agree on who renders what and reduce their timings / counters.  (The one real collective of a training step, the
all-reduce of the generator's gradients, belongs to DDP around the generator, not to this path.)

Works with any initialised ``torch.distributed`` backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None, device=None):
    """Initialise the default process group from torchrun's environment (RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, world).  A single process (WORLD_SIZE unset or 1) initialises nothing."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if (device is not None and torch.device(device).type == "cuda") else "gloo"
        kwargs = {}
        if backend == "nccl" and device is not None:
            kwargs["device_id"] = torch.device(device)
        dist.init_process_group(backend, **kwargs)
    return rank, world


def shard_range(num_units, rank, world):
    """Contiguous, balanced share of ``range(num_units)`` for ``rank``: sizes differ by at most one, earlier ranks
    get the longer shares.  Used for chunk batches (keeps a chunk's views on one GPU so its voxels are read once)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    per, extra = divmod(int(num_units), int(world))
    first = rank * per + min(rank, extra)
    return range(first, first + per + (1 if rank < extra else 0))


def shard_round_robin(num_units, rank, world):
    """Units rank, rank + world, ...: the dealing order of sliding-window chunks of a room
    (reference test_scene_as_chunks.py:156-157 enumerates them row by row)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return range(rank, int(num_units), int(world))


def _scalar_reduce(value, op, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value, device=None):
    """Timing rule of the bench: a multi-GPU step takes as long as its slowest rank."""
    return _scalar_reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value, device=None):
    """Whole-job totals (rays, chunks) from per-rank counts."""
    return _scalar_reduce(value, dist.ReduceOp.SUM, device)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
