"""Frame sharding host logic, including a world_size-2 gloo run on CPU."""
import os

import pytest
import torch

from mpv_prescalers_b200.sharding import ROW_HALO, max_over_ranks, rank_slice, row_bands, shard_bounds, sum_over_ranks


def test_shard_bounds_partition():
    for n in (0, 1, 5, 64, 65):
        for world in (1, 2, 3, 8):
            b = shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [s - a for a, s in b]
            assert max(sizes) - min(sizes) <= 1
    assert rank_slice(10, 1, 4) == slice(3, 6)
    with pytest.raises(ValueError):
        shard_bounds(4, 0)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = torch.arange(7 * 4, dtype=torch.float32).reshape(7, 4)
        mine = frames[rank_slice(7, rank, world)]
        # every rank processes only its own frames; the only communication is timing metadata
        t = max_over_ranks(1.0 + rank)
        total = sum_over_ranks(float(mine.shape[0]))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine[:, 0].tolist())
        q.put((rank, t, total, gathered))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_frame_sharding():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t, total, gathered in res:
        assert t == 2.0 and total == 7.0
        flat = [v for part in gathered for v in part]
        assert flat == [float(4 * i) for i in range(7)]  # contiguous, disjoint, complete


def test_row_bands_cover_the_frame_with_halo():
    for h in (1, 7, 64, 2160):
        for parts in (1, 2, 3, 8):
            bands = row_bands(h, parts)
            own = [(a, b) for a, b, _, _ in bands]
            assert own[0][0] == 0 and own[-1][1] == h and all(own[i][1] == own[i + 1][0] for i in range(parts - 1))
            for a, b, s0, s1 in bands:
                if b > a:
                    assert s0 == max(a - ROW_HALO, 0) and s1 == min(b + ROW_HALO, h)   # halo, clipped at true borders only
