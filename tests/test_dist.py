"""world_size-2 gloo test (CPU) of the N>1 host logic: GOP sharding and the statistics reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spatiotemporalentropymodel_b200.dist import reduce_stats, shard_units, summarize


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, T, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_units(n_units, rank, world)
        # each unit (GOP) contributes deterministic per-frame statistics
        stats = torch.zeros((3, T), dtype=torch.float64)
        for u in mine:
            g = torch.Generator().manual_seed(u)
            stats += torch.rand((3, T), generator=g, dtype=torch.float64) * 1000
        reduce_stats(stats)
        out_q.put((rank, mine, stats))
    finally:
        dist.destroy_process_group()


def test_shard_units_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(u for r in range(world) for u in shard_units(15, r, world))
        assert seen == list(range(15))
        sizes = [len(shard_units(15, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_units(4, 2, 2)


def test_two_rank_reduction_matches_single_process():
    world, n_units, T = 2, 7, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = torch.zeros((3, T), dtype=torch.float64)
    for u in range(n_units):
        g = torch.Generator().manual_seed(u)
        expect += torch.rand((3, T), generator=g, dtype=torch.float64) * 1000
    units = sorted(u for _, mine, _ in results for u in mine)
    assert units == list(range(n_units))
    for _, _, stats in results:
        assert torch.allclose(stats, expect, rtol=1e-12)
    bpp, psnr = summarize(expect, n_units * T, 1920 * 1080)
    assert bpp > 0 and psnr > 0


def test_reduce_stats_is_identity_without_process_group():
    s = torch.arange(6, dtype=torch.float64).reshape(3, 2)
    assert torch.equal(reduce_stats(s.clone()), s)
