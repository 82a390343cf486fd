"""Multi-GPU partitioning of the hot path: independent video streams, one per GPU.

Each rank owns one stream (its own prompt, seed, KV ring and cross cache) and runs the pipeline
unchanged; there is NO data-path collective.  torch.distributed is used only to (a) line ranks up
before / after a timed region and (b) combine per-rank device times with MAX, the way the reference
uses NCCL only for rank bookkeeping and a barrier (inference.py:43-48, 155-156).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def stream_assignment(num_streams: int, rank: int, world: int):
    """Stream ids handled by `rank` when `num_streams` videos are spread over `world` GPUs
    (round-robin, like DistributedSampler in the reference's inference.py:145-149)."""
    return list(range(rank, num_streams, world))


def stream_seeds(stream_id: int):
    """(noise seed, prompt seed) of a stream: every stream is an independent sequence."""
    return stream_id, 100 + stream_id


def barrier(device: Optional[torch.device] = None):
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value_ms: float, device: torch.device) -> float:
    t = torch.tensor([float(value_ms)], dtype=torch.float64,
                     device=device if device.type == "cuda" else torch.device("cpu"))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: float, elapsed_ms_local: float, device: torch.device) -> dict:
    """Whole-job throughput: all ranks' units divided by the slowest rank's device time."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    ms = max_over_ranks(elapsed_ms_local, device)
    return {"world": world, "ms": ms, "value": world * units_per_rank / (ms * 1e-3)}
