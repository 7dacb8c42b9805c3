"""CPU tests of the N>1 host logic: object sharding and the bench's cross-rank reduction over gloo
(world_size 2)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from forge_b200.synthetic import shard_objects


def test_shard_objects_partitions_exactly():
    for n in (1, 4, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_objects(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    # each rank "measures" a different time; the job-level figure is total rays / max time
    my_ms = 10.0 * (rank + 1)
    total_ms, e2e_ms, k1 = bench.reduce_max([my_ms, 2 * my_ms, 0.5 * my_ms], world, torch.device("cpu"))
    lo, hi = shard_objects(9, rank, world)
    n = torch.tensor([hi - lo])
    dist.all_reduce(n)
    q.put((rank, total_ms, e2e_ms, k1, int(n)))
    dist.barrier()
    dist.destroy_process_group()


def test_cross_rank_reduction_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total_ms, e2e_ms, k1, n in out:
        assert total_ms == 20.0 and e2e_ms == 40.0 and k1 == 10.0      # max over ranks
        assert n == 9                                                  # shards cover every object once
