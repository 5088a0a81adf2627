"""CPU, world_size 2 over gloo: frame sharding covers every frame exactly once and the metric reduce is exact."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bnerv_b200.shard import frame_indices, norm_index, reduce_metric


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = frame_indices(n_frames, rank, world)
    # a per-frame "metric" that depends only on the frame index -> the global mean is known in closed form
    local_sum = sum(10.0 + norm_index(i, n_frames) for i in mine)
    mean, count = reduce_metric(local_sum, len(mine))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mean, count, gathered))
    dist.destroy_process_group()


def test_two_rank_shard_union_and_metric_mean():
    world, n_frames = 2, 13                      # ragged: 7 + 6 frames
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    expect = sum(10.0 + (i + 1) / n_frames for i in range(n_frames)) / n_frames
    for rank, mean, count, gathered in res:
        assert count == n_frames
        assert abs(mean - expect) < 1e-12
        assert sorted(gathered[0] + gathered[1]) == list(range(n_frames))
        assert not set(gathered[0]) & set(gathered[1])


def test_single_process_reduce_is_identity():
    mean, count = reduce_metric(30.0, 3)
    assert mean == 10.0 and count == 3
    assert frame_indices(5, 0, 1) == [0, 1, 2, 3, 4] and frame_indices(5, 3, 8) == [3] and frame_indices(2, 5, 8) == []
