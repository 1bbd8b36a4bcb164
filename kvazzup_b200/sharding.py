"""Placement of independent participant streams on the GPUs of one node.

The reference keeps one encoder plus one decoder graph per peer (filtergraph.h:94-108) and the
streams share no data, so the multi-GPU plan is a static partition: stream s runs on rank
s mod world_size.  No collective is on the data path; torch.distributed is used only to agree on
totals (tests: gloo, world_size 2, on CPU).
"""
from __future__ import annotations


def streams_of_rank(n_streams: int, rank: int, world: int) -> list[int]:
    if world <= 0 or not 0 <= rank < world or n_streams < 0:
        raise ValueError("bad rank / world / stream count")
    return list(range(rank, n_streams, world))


def rank_of_stream(stream_id: int, world: int) -> int:
    return stream_id % world


def gather_totals(local_frames: int, local_seconds: float):
    """(sum of frames, max of seconds) over all ranks; identity when not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_frames, local_seconds
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    f = torch.tensor([float(local_frames)], device=dev)
    s = torch.tensor([float(local_seconds)], device=dev)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    return int(f.item()), float(s.item())
