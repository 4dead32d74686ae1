"""Multi-process host logic of the sample-sharded path, on CPU with the gloo backend (world_size 2).
The data path itself has no collective; what is tested is the partition and the throughput reduction
(frames of all ranks / slowest rank) that bench.py reports."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mm_training_b200.sharding import aggregate_throughput, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 32, 33, 64):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_aggregate_single_process():
    fps, ms, n = aggregate_throughput(32, 2.0)
    assert n == 32 and ms == 2.0 and abs(fps - 16000.0) < 1e-9


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        b, e = shard_range(9, rank, world)                      # 9 frames over 2 ranks -> 5 + 4
        # every rank reduces its own shard of a per-frame quantity; no rank needs another's frames
        frames = torch.arange(9, dtype=torch.float64)[b:e]
        local = float((frames * frames).sum())
        fps, ms, n = aggregate_throughput(e - b, 1.0 + rank)    # rank 1 is the slow one: 2 ms
        tot = torch.tensor([local], dtype=torch.float64)
        dist.all_reduce(tot)                                    # test-only check that the shards tile the batch
        out[rank] = (b, e, fps, ms, n, float(tot.item()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert (res[0][0], res[0][1]) == (0, 5) and (res[1][0], res[1][1]) == (5, 9)
    for r in range(world):
        _, _, fps, ms, n, tot = res[r]
        assert n == 9 and ms == 2.0 and abs(fps - 4500.0) < 1e-9      # all frames / slowest rank
        assert tot == float(sum(i * i for i in range(9)))
