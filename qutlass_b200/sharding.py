"""Multi-GPU plumbing for the hot path: the M (row) dimension shards embarrassingly.

One process per GPU (torch.distributed).  Weights are quantised once on rank 0 and broadcast ONCE at
setup (packed e2m1 + blocked scales); every rank then quantises and multiplies only its own activation
rows.  There is no collective in the steady-state loop, so nothing to fuse a collective into
(the reference itself is single-GPU only).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_rows(m: int, world: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """Contiguous row range [start, start+rows) owned by `rank`.

    Shards are multiples of `align` rows (so every shard's blocked scale buffer is self-contained: the
    block-scaled layout works in 128-row blocks) except possibly the last non-empty one."""
    assert world >= 1 and 0 <= rank < world
    blocks = (m + align - 1) // align
    base, extra = divmod(blocks, world)
    my_blocks = base + (1 if rank < extra else 0)
    start_block = rank * base + min(rank, extra)
    start = min(start_block * align, m)
    end = min((start_block + my_blocks) * align, m)
    return start, end - start


def broadcast_weights(wq: torch.Tensor, wsf_blocked: torch.Tensor, src: int = 0, group=None) -> None:
    """ONE setup-time broadcast of the quantised weights (in place).  float8 scale tensors travel as bytes."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.broadcast(wq, src, group=group)
    dist.broadcast(wsf_blocked.view(torch.uint8), src, group=group)
