"""world_size-2 gloo test of the ray-sharding + hit-gather host logic (bvh_b200/multi_gpu.py) on CPU.
The tracer is a stand-in (the oracle, which tests may use) because there is no GPU here; what is being
tested is the sharding arithmetic, chunking and the all-gather layout, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bvh_b200.multi_gpu import chunk_bounds, shard_range


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 64, 1000, 10_004_569):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert chunk_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert chunk_bounds(2, 4) == [(0, 1), (1, 2)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from bvh_b200 import scenes
        from bvh_b200.multi_gpu import ShardedTracer
        from oracle.pyoracle import TIE_LOWEST_ID, Oracle
        oracle = Oracle()
        tris = scenes.soup(1500)
        bb, cc = oracle.tri_bboxes_centers(tris)
        tree = oracle.build(bb, cc, quality="low")          # every rank builds the same (replicated) BVH
        oracle.set_triangles(tree, tris)
        all_rays = scenes.make_primary("soup", 40, 40)
        begin, end = rank * 800, (rank + 1) * 800          # equal contiguous shards
        rays = all_rays[begin:end]

        def trace(b, e, out):
            ids, t, u, v = oracle.trace(tree, rays[b:e], flags=TIE_LOWEST_ID)
            rec = np.stack([ids.view(np.int32), t.view(np.int32), u.view(np.int32), v.view(np.int32)], axis=1)
            out.copy_(torch.from_numpy(rec))

        tracer = ShardedTracer(800, 4, torch.int32, torch.device("cpu"), trace, chunks=3)
        tracer.step()
        g = tracer.global_hits().numpy()
        ids, t, u, v = oracle.trace(tree, all_rays, flags=TIE_LOWEST_ID)
        want = np.stack([ids.view(np.int32), t.view(np.int32), u.view(np.int32), v.view(np.int32)], axis=1)
        ok = np.array_equal(g, want)
        open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_reassembles_global_order(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"
