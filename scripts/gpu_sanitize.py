"""Small build + trace for compute-sanitizer runs (memcheck / racecheck of the build kernels)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bvh_b200.api as api
from bvh_b200 import scenes
for n, dtype in ((1, np.float32), (2, np.float32), (300, np.float32), (5000, np.float32), (3000, np.float64)):
    tris = scenes.soup(n, seed=3).astype(dtype)
    bvh = api.Bvh.build_triangles(tris)
    rays = scenes.make_primary("soup", 32, 32).astype(dtype)
    hits = bvh.intersect_rays(rays)
    b, ix, ids = bvh.arrays()
    print(n, dtype.__name__, "nodes", b.shape[0], "hit", float((hits["prim_id"] != 0xFFFFFFFF).mean()), flush=True)
