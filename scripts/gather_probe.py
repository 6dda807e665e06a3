"""Single-GPU cost of the gather variants of the traversal kernel: the plain call, the gather entry point with ONE target
(this GPU's own array) delivered by staged bulk copies, by one store per record, and into two local arrays (what a
second rank costs in instructions, without NVLink)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bvh_b200.api as api
from bvh_b200 import scenes

api.set_device(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); api.set_stream(stream.cuda_stream)
tris = scenes.soup(1_000_000)
rays_np = scenes.make_primary("soup", 3163, 3163)
n = rays_np.shape[0]
rays = torch.from_numpy(rays_np).to(dev)
bvh = api.Bvh.build_triangles(tris)
a = torch.empty((n, 4), dtype=torch.int32, device=dev); b = torch.empty_like(a); c = torch.empty_like(a)
L = api.lib()

def timed(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    return float(np.median([s.elapsed_time(e) for s, e in ev]))

def plain():
    if L.bvh3f_intersect_rays(bvh.handle, rays.data_ptr(), n, a.data_ptr(), api.DEVICE_POINTERS): raise SystemExit(api.last_error())
t0 = timed(plain)
ref = a.clone()
print(f"plain                      {t0:.3f} ms  {n / t0 / 1e3:.1f} Mrays/s")
for staging in (1, 0):
    api.set_option("gather_staging", staging)
    for name, targets in (("1 target", [b]), ("2 targets", [b, c])):
        def g():
            bvh.intersect_rays_gather(rays.data_ptr(), n, [t.data_ptr() for t in targets], 0, flags=api.DEVICE_POINTERS)
        t = timed(g)
        ok = all(torch.equal(x, ref) for x in targets)
        print(f"gather staging={staging} {name:10s} {t:.3f} ms  {n / t / 1e3:.1f} Mrays/s  kernel {bvh.properties()['last_kernel']}  identical {ok}")
