"""
Multi-GPU figure sharding. The path shards over independent figures: rank r takes
``items[r::world]`` — the reference's own scheme (examples/eval.py:80-83,125) — every rank owns a full
engine, the only collective on the model side is ONE broadcast of the weight arena at load
(``detikzify_b200.model.load(..., broadcast=True)``) and results are gathered once at the end
(reference: ``dist.all_gather_object`` + interleave, examples/eval.py:85-93,132-133).
No per-step collective exists because no tensor is shared between figures.
"""
from __future__ import annotations

from itertools import count
from typing import Any, Iterable, List, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(items: Sequence[Any], r: int | None = None, n: int | None = None) -> List[Any]:
    """Striped shard of ``items`` for rank ``r`` of ``n``."""
    r = rank() if r is None else r
    n = world() if n is None else n
    return list(items)[r::n]


def interleave(chunks: Iterable[Sequence[Any]]) -> List[Any]:
    """Inverse of ``shard`` over all ranks: [c0[0], c1[0], ..., c0[1], c1[1], ...] (ragged tails kept)."""
    chunks = [list(c) for c in chunks]
    out: List[Any] = []
    for i in count():
        row = [c[i] for c in chunks if i < len(c)]
        if not row:
            return out
        out.extend(row)


def gather_results(local: Sequence[Any]) -> List[Any]:
    """Gather every rank's results (python objects) and restore dataset order."""
    if world() == 1:
        return list(local)
    gathered: List[Any] = [None] * world()
    dist.all_gather_object(gathered, list(local))
    return interleave(gathered)


def broadcast_arena(arena: torch.Tensor | None, nbytes: int, device: torch.device, src: int = 0) -> torch.Tensor:
    """Single broadcast of the contiguous bf16 weight arena from ``src`` (NCCL over NVLink on GPUs, gloo on CPU)."""
    if rank() == src:
        assert arena is not None
        buf = arena.to(device)
    else:
        buf = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=device)
    if world() > 1:
        dist.broadcast(buf, src=src)
    return buf
