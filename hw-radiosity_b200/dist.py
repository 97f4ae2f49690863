"""Multi-GPU launcher glue (one process per GPU, torch.distributed for the plumbing).

The data path needs no collective: receivers are sharded, and the only exchange - the shooter
selection - happens inside the kernels through peer-mapped mailboxes (include/radk.h, multi-GPU
section).  torch.distributed is used once, to all-gather the 64-byte CUDA IPC handles that introduce
the ranks to each other, and for barriers / result gathers in bench.py.
"""
import numpy as np

from .api import Context, default_opts


def init(backend="nccl"):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    import os

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def sharded_context(model, rank, world, device, **opts):
    """Create this rank's context and connect it to its peers (collective call)."""
    ctx = Context(model, opts=default_opts(rank=rank, world=world, device=device, **opts))
    if world > 1:
        import torch.distributed as dist
        handles = [None] * world
        dist.all_gather_object(handles, ctx.peer_handle())
        ctx.peer_connect(handles)
        dist.barrier()
    return ctx


def gather_state(ctx, rank, world):
    """Rank 0 gets the full per-slot (accB, deltaB); other ranks get None.  (Result collection, not the hot path.)"""
    accB, deltaB, parent = ctx.download()
    if world == 1:
        return accB, deltaB
    import torch.distributed as dist
    own = ctx.owned_slots()
    parts = [None] * world if rank == 0 else None
    dist.gather_object((own, accB[own], deltaB[own]), parts, dst=0)
    if rank != 0:
        return None
    for slots, a, d in parts:
        accB[slots], deltaB[slots] = a, d
    return accB, deltaB
