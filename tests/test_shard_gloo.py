"""Multi-rank host logic (work partition + final gather) on CPU with the gloo backend."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_everything_once():
    from tetraear_b200 import shard
    for nch, nseg, world in ((800, 5, 8), (1, 183, 2), (3, 7, 4), (8, 1, 8), (2, 9, 8)):
        seen = []
        for r in range(world):
            seen += shard.partition(nch, nseg, world, r)
        assert sorted(seen) == sorted(shard.work_items(nch, nseg)), (nch, nseg, world)
        if nch >= world:
            for r in range(world):       # a channel never spans ranks, its segments stay ordered
                items = shard.partition(nch, nseg, world, r)
                assert all(c % world == r for c, _ in items)
                assert items == sorted(items)


def test_segments_overlap_by_one_burst():
    from tetraear_b200 import shard
    burst = 255 * 83 * 13
    seg = shard.segment_bounds(20_000_000, 4_000_000, burst)
    assert seg[0][0] == 0 and seg[-1][1] == 20_000_000
    for (a0, a1), (b0, b1) in zip(seg, seg[1:]):
        assert a1 - b0 == burst or b1 == 20_000_000
    assert shard.dedup_frames([(0, 10, "a"), (0, 10, "b"), (1, 10, "c")]) == [(0, 10, "a"), (1, 10, "c")]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tetraear_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    items = shard.partition(5, 4, world, rank)
    # fake per-item burst records; item (1, 2) is reported twice (segment overlap) by its owner
    recs = [(c, s * 1000 + 7, f"r{rank}") for c, s in items]
    if (1, 2) in items:
        recs.append((1, 2007, "dup"))
    merged = shard.gather_frames(recs)
    if rank == 0:
        np.save(os.path.join(out_dir, "n.npy"), np.array([len(merged)]))
        assert merged == sorted(merged, key=lambda r: (r[0], r[1]))
        assert len({(c, p) for c, p, _ in merged}) == len(merged) == 20
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world_size_2(tmp_path, monkeypatch):
    # spawned children unpickle `_worker` by module name: the repo root must be importable there
    monkeypatch.setenv("PYTHONPATH", ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert int(np.load(tmp_path / "n.npy")[0]) == 20
