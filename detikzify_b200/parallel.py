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


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def pin_to_gpu_numa(device_index: int) -> dict:
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off (one process per GPU: the thread that polls the
    mapped token ring and feeds streamers then never crosses the socket interconnect). Best effort: returns what was done."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set(_parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return info


def host_threads() -> dict:
    """CPU threads this process may actually use: the scheduler affinity mask capped by the cgroup CPU quota (a container
    on a 64-core host may own 8 of them; sizing a thread pool by the host's core count oversubscribes 8x)."""
    import math
    import os
    aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        txt = open("/sys/fs/cgroup/cpu.max").read().split()
        if txt and txt[0] != "max":
            quota = float(txt[0]) / float(txt[1])
    except (OSError, ValueError, IndexError):
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except (OSError, ValueError):
            pass
    use = aff if quota is None else max(1, min(aff, int(math.ceil(quota))))
    return {"affinity": aff, "cgroup_quota": quota, "host_logical": os.cpu_count(), "use": use}
