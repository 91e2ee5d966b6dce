"""N>1 path on CPU: world_size-2 gloo run of the sharding helpers (block partition, gather of per-frame results in
frame order, max-over-ranks timing, whole-job throughput).  The per-shard 'detector' here is the CPU oracle, used
only as a stand-in payload for the host-side plumbing under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isaac_ros_apriltag_b200 import sharding


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 256, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [sharding.shard_bounds(n, r, world)[1] - sharding.shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from isaac_ros_apriltag_b200 import synth
    from oracle import oracle as O
    rng = np.random.default_rng(7)
    frames = np.stack([synth.make_frame(rng, 320, 240, [("tag36h11", i)], side_px=(60, 100), noise_sigma=0.5)[0] for i in range(5)])

    def detect_fn(fs):
        orc = O.Oracle(("tag36h11",))
        return [[d["id"] for d in orc.detect(f)] for f in fs]

    res = sharding.run_sharded(detect_fn, frames)
    t = sharding.max_over_ranks(1.0 + rank)          # slowest rank defines the time
    thr = sharding.whole_job_throughput(10.0, 1.0 + rank)
    if rank == 0:
        q.put((res, t, thr))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res, t, thr = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [[0], [1], [2], [3], [4]]  # frame order preserved across shards (3 + 2 split)
    assert t == 2.0
    assert thr == pytest.approx(20.0 / 2.0)
