"""Probe: effect of the order of the rays inside the batch on the
default traversal kernel (raster rows vs 8x4 / 4x8 / 16x2 pixel tiles vs Morton order)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
import bvh_b200.api as api
from bvh_b200 import scenes

W = H = 3163
tris = scenes.soup(1_000_000)
api.lib().bvh_cuda_set_stream(None)          # legacy default stream: torch events then bracket the kernels
bvh = api.Bvh.build_triangles(tris)

def timed(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    bvh.sync(); torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        times.append(a.elapsed_time(b))
    return float(np.median(times))

rays_np = scenes.make_primary("soup", W, H)
x, y = np.meshgrid(np.arange(W), np.arange(H))          # row-major pixel coordinates
x = x.ravel(); y = y.ravel()
def tile_order(tw, th):
    key = ((y // th) * ((W + tw - 1) // tw) + (x // tw)) * (tw * th) + (y % th) * tw + (x % tw)
    return np.argsort(key, kind="stable")
def morton_order():
    def spread(v):
        v = v.astype(np.uint64)
        v = (v | (v << 8)) & 0x00FF00FF; v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555
        return v
    return np.argsort(spread(x) | (spread(y) << 1), kind="stable")
orders = {"raster": None, "tile8x4": tile_order(8, 4), "tile4x8": tile_order(4, 8), "tile16x2": tile_order(16, 2),
          "tile8x8": tile_order(8, 8), "morton": morton_order()}
ref = None
for name, perm in orders.items():
    r = rays_np if perm is None else rays_np[perm]
    rays = torch.from_numpy(np.ascontiguousarray(r)).cuda()
    out = torch.empty((rays.shape[0], 4), dtype=torch.int32, device="cuda")
    def run():
        if api.lib().bvh3f_intersect_rays(bvh.handle, rays.data_ptr(), rays.shape[0], out.data_ptr(), api.DEVICE_POINTERS):
            raise SystemExit("intersect: " + api.last_error())
    ms = timed(run, reps=15)
    got = out.cpu().numpy()
    if perm is None: ref = got
    else:
        back = np.empty_like(got); back[perm] = got
        assert np.array_equal(back, ref), name
    print(f"{name:9s}: {ms:.3f} ms  {rays.shape[0]/ms/1e3:.0f} Mrays/s", flush=True)
