"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path (map sharding, max-over-ranks, result gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wavefunctiondiffusion_b200.utils.sharding import balance_by_cost, gather_results, max_over_ranks, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_balance_by_cost_config3():
    # BASELINE config 3: 8 tilesets x 8 maps, cost = F_step per map
    f = [737.3, 1018.2, 952.1, 983.2, 703.4, 1043.8, 1152.0, 1269.1]
    costs = [c for c in f for _ in range(8)]
    parts = balance_by_cost(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(64))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) / (sum(loads) / 8) < 1.05


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(5, world, rank)
        local = torch.arange(lo, hi, dtype=torch.float32) * 10
        counts = [shard_range(5, world, r)[1] - shard_range(5, world, r)[0] for r in range(world)]
        allv = gather_results(local, counts)
        tiles = gather_results(torch.full((hi - lo, 2, 2), rank, dtype=torch.int64), counts)
        m = max_over_ranks(1.0 + rank, torch.device("cpu"))
        q.put((rank, allv.tolist(), tiles[:, 0, 0].tolist(), m))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_max():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, allv, tiles, m in res:
        assert allv == [0.0, 10.0, 20.0, 30.0, 40.0]
        assert tiles == [0, 0, 0, 1, 1]
        assert m == 2.0
