"""CPU tests: the C-ABI library loads and exports every symbol include/avi_b200.h declares (no compute calls), the product
path refuses to run without CUDA, and the clip-sharding helpers behave under a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avi_talking_b200 import _lib, shard


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = _lib.load(check_symbols=True)
    names = _lib.declared_symbols()
    assert len(names) >= 25 and "avi_gemm_bf16_tc" in names and "avi_flame_lbs_fwd" in names
    assert lib.avi_version() == 100
    assert lib.avi_last_error() == b""


def test_no_cpu_fallback():
    from avi_talking_b200 import ops
    x = torch.zeros(4, 8)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        ops.layernorm(x, torch.ones(8), torch.zeros(8))
    with pytest.raises(RuntimeError):
        ops.linear(x, torch.zeros(3, 8), None)


def test_clip_shard_partitions():
    for n, w in ((512, 8), (64, 1), (7, 4), (3, 8)):
        parts = [shard.clip_shard(n, r, w) for r in range(w)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.clip_shard(4, 4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, lr, w = shard.env_rank_world()
        mine = shard.clip_shard(9, r, w)                      # ragged: 5 + 4 clips
        counts = shard.gather_clip_counts(len(mine))
        ms = shard.max_over_ranks([10.0 + rank, 3.0 - rank])
        dist.barrier()
        out.put((rank, mine, counts, ms))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, c0, t0), (r1, m1, c1, t1) = res
    assert m0 == [0, 2, 4, 6, 8] and m1 == [1, 3, 5, 7]
    assert c0 == c1 == [5, 4]
    assert t0 == t1 == [11.0, 3.0]
