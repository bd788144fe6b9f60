"""Multi-GPU layout of the detect path: replicas only, one video stream per GPU, no data-path collective.

The reference runs N independent single-GPU processes for multi-GPU inference (SPEED.md:61 "4 / 1*";
obj_detect_tracking.py:241-242 asserts one GPU per process); tracker state is per video and per class
(obj_detect_tracking.py:547-558), so streams never exchange data.  The only cross-rank operations are the
ones the benchmark contract needs: a barrier and a max-over-ranks of the timed region."""
from __future__ import annotations


def streams_for_rank(n_streams: int, rank: int, world: int) -> list:
    """Stream i is served by rank i % world (weak scaling: BASELINE config 4 = 8 streams on 8 GPUs)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return [s for s in range(n_streams) if s % world == rank]


def max_over_ranks(values, device=None):
    """Element-wise MAX over ranks of a list of floats (the timed region is the slowest rank's)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def aggregate_fps(frames_per_rank: int, seconds: float, world: int) -> float:
    """Whole-job throughput: all ranks' frames over the slowest rank's time."""
    return frames_per_rank * world / seconds
