"""Multi-process (world_size 2, gloo, CPU) coverage of the N>1 host logic: batch-axis sharding of independent
stereo sequences, per-rank recurrent state, and the max-over-ranks / whole-job throughput reduction that
bench.py reports (no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from temporalstereo_b200 import shard


def test_shard_range_partitions_exactly():
    for total in (0, 1, 2, 5, 16, 32, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
                assert b0 == a1                                   # contiguous, no overlap, no gap
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    # the BASELINE configs: 16 sequences on 4 GPUs, 32 on 8
    assert shard.shard_range(16, 4, 3) == (12, 16) and shard.shard_range(32, 8, 0) == (0, 4)
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_shard_batch_rejects_ragged():
    a, b = torch.zeros(4, 3), torch.zeros(5, 3)
    with pytest.raises(ValueError):
        shard.shard_batch([a, b], 2, 0)
    assert [t.shape[0] for t in shard.shard_batch([a, a], 2, 1)] == [2, 2]
    assert [t.shape[0] for t in shard.shard_batch([b], 2, 0)] == [3]
    assert [t.shape[0] for t in shard.shard_batch([torch.zeros(1, 2)], 2, 1)] == [0]       # empty shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert shard.env_world() == (rank, rank, world)
        # the "sequence batch": every rank builds the same global batch and keeps only its slice
        g = torch.Generator().manual_seed(0)
        frames = torch.randn(total, 3, 8, 8, generator=g)
        ids = torch.arange(total)
        mine, my_ids = shard.shard_batch([frames, ids], world, rank)
        # a stand-in for the per-frame hot path with a recurrent state: state_t = state_{t-1} + mean(frame)
        state = torch.zeros(mine.shape[0])
        for _ in range(3):
            state = state + mine.mean(dim=(1, 2, 3))
        # no data-path collective was needed; gather only to CHECK the shards against the unsharded result
        gathered = [None] * world
        dist.all_gather_object(gathered, (my_ids.tolist(), state.tolist()))
        all_ids = [i for ids_, _ in gathered for i in ids_]
        assert all_ids == list(range(total)), all_ids
        want = 3 * frames.mean(dim=(1, 2, 3))
        got = torch.tensor([v for _, st in gathered for v in st])
        assert torch.allclose(got, want, atol=1e-6)
        # timing reduction: the job is as slow as its slowest rank; throughput counts every rank's units
        shard.barrier(dist)
        fps, ms, units = shard.aggregate_throughput(int(mine.shape[0]) * 3, 10.0 * (rank + 1), dist)
        assert ms == 10.0 * world and units == 3 * total
        assert abs(fps - units / (ms * 1e-3)) < 1e-9
        assert shard.max_over_ranks(float(rank), dist) == world - 1
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 16])
def test_two_rank_gloo_sharding(total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"
