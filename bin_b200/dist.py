"""Multi-GPU plumbing for the BIN hot path: one process per GPU, windows are the parallel unit.

SURVEY.md 8e: every window's forward is a pure function of its 6 frames and the weights (state is
re-zeroed per call, RDN.py:423-434), so inference shards over windows with NO collective in the
loop; the only communication is one broadcast of the parameters from rank 0 at start-up (NCCL over
NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_windows(n_windows: int, rank: int, world: int) -> List[int]:
    """Window w -> rank w mod world (round-robin keeps consecutive frames spread evenly)."""
    return list(range(rank, n_windows, world))


def broadcast_weights(module: torch.nn.Module, src: int = 0) -> int:
    """One flat broadcast of the 540 unique tensors (11.44 M fp32 = 45.8 MB).  Returns bytes sent."""
    params = [p for p in module.parameters()]
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))      # in-place: bumps ._version -> packed blobs refresh
            off += n
    return flat.numel() * flat.element_size()


def max_over_ranks(value: float, device) -> float:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
