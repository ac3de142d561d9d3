"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: one weight-arena broadcast from rank 0, striped
figure shards, one object gather + interleave at the end — the N>1 path of bench.py / examples."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from detikzify_b200 import _lib
        from detikzify_b200.engine import pack_arena, to_c_config
        from detikzify_b200.model.configuration import preset
        from detikzify_b200.model.weights import random_init
        from detikzify_b200.parallel import broadcast_arena, gather_results, shard
        cfg = preset("tiny")
        nbytes = _lib.load_library().dtk_arena_bytes(C.byref(to_c_config(cfg)))
        arena = pack_arena(cfg, random_init(cfg, seed=0)) if rank == 0 else None   # only rank 0 materialises weights
        got = broadcast_arena(arena, nbytes, torch.device("cpu"))
        checksum = int(got.view(torch.int16).to(torch.int64).sum())
        figures = list(range(7))
        mine = shard(figures)
        results = [(f, f * f) for f in mine]            # stand-in for per-figure generations
        merged = gather_results(results)
        q.put((rank, checksum, mine, merged))
    finally:
        dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, c0, m0, g0), (r1, c1, m1, g1) = outs
    assert c0 == c1 and c0 != 0                          # identical arena bytes on both ranks
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5]
    assert g0 == g1 == [(f, f * f) for f in range(7)]    # dataset order restored
