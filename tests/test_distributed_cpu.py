"""The N > 1 host logic (clipcap_b200/distributed.py) on CPU: world_size-2 gloo processes."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clipcap_b200.distributed import all_gather_prefix, gather_tokens, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 1024, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, K, d, EL = 3, 4, 8, 5
        g = torch.Generator().manual_seed(11)
        full = torch.randn(world * B, K, d, generator=g)          # what a single process would compute
        lo, hi = shard_range(world * B, rank, world)
        local = full[lo:hi].clone()
        out = all_gather_prefix(local)
        ok = torch.equal(out, full)                                # multi-rank result == single-rank result, exactly
        pre = torch.empty_like(full)
        ok &= all_gather_prefix(local, out=pre) is pre and torch.equal(pre, full)
        # asynchronous form (what caption_step uses to run the all-gather behind the decode): same bytes after wait()
        pre2 = torch.zeros_like(full)
        got, work = all_gather_prefix(local, out=pre2, async_op=True)
        work.wait()
        ok &= got is pre2 and torch.equal(pre2, full)
        toks_full = torch.arange(world * B * EL, dtype=torch.int32).view(world * B, EL)
        lens_full = torch.arange(world * B, dtype=torch.int32)
        toks, lens = gather_tokens(toks_full[lo:hi].clone(), lens_full[lo:hi].clone())
        ok &= torch.equal(toks, toks_full) and torch.equal(lens, lens_full)
        try:
            all_gather_prefix(local, out=torch.empty(1))
            ok = False
        except ValueError:
            pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_prefix_all_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_single_process_is_identity():
    x = torch.randn(2, 3, 4)
    assert all_gather_prefix(x) is x
    same, work = all_gather_prefix(x, async_op=True)
    assert same is x and work is None
    t, l = gather_tokens(torch.zeros(2, 5, dtype=torch.int32), torch.zeros(2, dtype=torch.int32))
    assert t.shape == (2, 5) and l.shape == (2,)
