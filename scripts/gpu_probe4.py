"""Probe: LBVH build time (1M-triangle soup, device-resident vertices, CUDA events) for every hierarchy-kernel
variant; also checks that all variants produce the same tree."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
import bvh_b200.api as api
from bvh_b200 import scenes

api.lib().bvh_cuda_set_stream(None)
tris = scenes.soup(1_000_000)
verts = torch.from_numpy(tris).cuda()
ref = None
for variant in ("global", "thread64", "thread128", "thread256", "rounds128", "rounds256", "thread256"):
    os.environ["BVH_B200_HIERARCHY"] = variant
    def build():
        return api.Bvh.build_triangles(verts.data_ptr(), count=tris.shape[0], dtype=np.float32, flags=api.DEVICE_POINTERS)
    for _ in range(3): b = build()
    torch.cuda.synchronize()
    times = []
    for _ in range(25):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); b = build(); e.record(); e.synchronize()
        times.append(s.elapsed_time(e))
    arrays = b.arrays()
    if ref is None: ref = arrays
    same = all(np.array_equal(x, y) for x, y in zip(arrays, ref))
    print(f"{variant:10s}: median {np.median(times)*1e3:7.1f} us  min {np.min(times)*1e3:7.1f} us   same tree: {same}", flush=True)
