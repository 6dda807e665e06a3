"""Multi-rank GPU test of the fused gather (bvh_b200/multi_gpu.py FusedGatherTracer): every rank traces its shard,
the traversal kernel delivers the hit records into every rank's symmetric-memory buffer, and a NON-OWNER rank
compares a sample of the records it received with the compiled reference (oracle/_ref; the plain-C oracle when
that is absent).  Needs at least two GPUs on the box (skipped otherwise) — SURVEY.md §8(e), BASELINE configs[3]."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["BVH_TEST_ROOT"])
import bvh_b200.api as api
from bvh_b200 import scenes
from bvh_b200.multi_gpu import FusedGatherTracer
from oracle.pyoracle import TIE_LOWEST_ID, ANY_HIT, Oracle, Ref, ref_available

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); api.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
mode, kernel = os.environ["BVH_TEST_MODE"], os.environ["BVH_TEST_KERNEL"]
n_tris, img = 200_000, 1024
tris = scenes.soup(n_tris)
rows = img // world
all_rays = scenes.make_primary("soup", img, img)
if kernel == "wide":                       # incoherent order inside every shard: stragglers, evicted staging slots
    for r in range(world):
        seg = all_rays[r * rows * img:(r + 1) * rows * img]
        seg[:] = seg[np.random.default_rng(r).permutation(seg.shape[0])]
mine = np.ascontiguousarray(all_rays[rank * rows * img:(rank + 1) * rows * img])
bvh = api.Bvh.build_triangles(tris)
d_rays = torch.from_numpy(mine).to(device)
if mode in ("direct", "multicast"):        # one store / one multimem store per record; the other modes stage 32 records per bulk copy
    api.set_option("gather_staging", 0)
flags = api.DEVICE_POINTERS | (api.KERNEL_WIDE if kernel == "wide" else 0)
tracer = FusedGatherTracer(bvh, d_rays, 4, flags=flags, mode="multicast" if mode.startswith("multicast") else "peer")
for _ in range(3):                         # three steps: both halves of the double buffer are used
    tracer.step()
tracer.check()
torch.cuda.synchronize()
got = tracer.global_hits().cpu().numpy().view(api.HIT3F).reshape(-1)
# a sample of the records of the OTHER ranks, checked against the reference
rng = np.random.default_rng(100 + rank)
others = np.concatenate([np.arange(r * rows * img, (r + 1) * rows * img) for r in range(world) if r != rank])
sample = np.sort(rng.choice(others, size=50_000, replace=False))
if ref_available():
    ref = Ref()
    bb, cc = ref.tri_bboxes_centers(tris)
    tree = ref.build(bb, cc, quality="high", threads=0)
    ref.set_triangles(tree, tris)
    ids, t, u, v = ref.trace(tree, all_rays[sample], flags=TIE_LOWEST_ID, threads=0)
else:
    orc = Oracle()
    bb, cc = orc.tri_bboxes_centers(tris)
    tree = orc.build(bb, cc, quality="low")
    orc.set_triangles(tree, tris)
    ids, t, u, v = orc.trace(tree, all_rays[sample], flags=TIE_LOWEST_ID)
rec = got[sample]
ok = (np.array_equal(rec["prim_id"], ids) and np.array_equal(rec["t"].view(np.uint32), t.view(np.uint32))
      and np.array_equal(rec["u"].view(np.uint32), u.view(np.uint32)) and np.array_equal(rec["v"].view(np.uint32), v.view(np.uint32)))
bad = int((rec["prim_id"] != ids).sum())
print(f"rank {rank}: mode {tracer.mode}/{mode} kernel {kernel} sample {sample.shape[0]} mismatching ids {bad} hits {(ids != 0xFFFFFFFF).mean():.3f} -> {'ok' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.gpu
@pytest.mark.parametrize("mode,kernel", [("peer", "auto"), ("peer", "wide"), ("direct", "auto"), ("multicast", "auto"), ("multicast_staged", "auto")])
def test_gathered_records_on_a_non_owner_rank_match_the_reference(gpu_lib, tmp_path, mode, kernel):
    world = min(gpu_lib.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs on the box")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BVH_TEST_ROOT=ROOT, BVH_TEST_MODE=mode, BVH_TEST_KERNEL=kernel)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    if mode.startswith("multicast") and "no multicast address" in (res.stdout + res.stderr):
        pytest.skip("the fabric offers no multicast address")
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("-> ok") == world, res.stdout
