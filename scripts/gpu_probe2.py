"""Repeated full-size traces with one kernel variant, compared with the one-lane kernel; host-timed."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bvh_b200.api as api
from bvh_b200 import scenes
kind = sys.argv[1] if len(sys.argv) > 1 else "pair"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
flag = {"simple": api.KERNEL_SIMPLE, "notma": api.KERNEL_NO_TMA, "tma": api.KERNEL_TMA, "pair": api.KERNEL_PAIR}[kind]
import torch
tris = scenes.soup(1_000_000)
rays_np = scenes.make_primary("soup", 3163, 3163)
bvh = api.Bvh.build_triangles(tris)
rays = torch.from_numpy(rays_np).cuda()
ref = torch.empty((rays.shape[0], 4), dtype=torch.int32, device="cuda")
out = torch.empty_like(ref)
def run(dst, fl):
    if api.lib().bvh3f_intersect_rays(bvh.handle, rays.data_ptr(), rays.shape[0], dst.data_ptr(), api.DEVICE_POINTERS | fl):
        raise SystemExit("intersect: " + api.last_error())
    bvh.sync()
run(ref, api.KERNEL_NO_TMA)
print("reference trace done", flush=True)
for i in range(reps):
    out.zero_()
    t0 = time.perf_counter()
    run(out, flag)
    dt = time.perf_counter() - t0
    print(f"{kind} rep {i}: {dt*1e3:.2f} ms  {rays.shape[0]/dt/1e6:.0f} Mrays/s  equal={bool(torch.equal(out, ref))}", flush=True)
