"""The re-authored C++ surface (include/bvh/v2/*.h) on the host: Bvh::intersect / refit / serialize /
extract_bvh / ThreadPool / ParallelExecutor against the golden vectors of the unmodified reference.
No GPU involved (DefaultBuilder::build, the only GPU entry of the header, is exercised by the -m gpu tests
through the reference's own example programs)."""
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe():
    out = os.path.join(ROOT, "tests", "_build", "cxx_surface_test")
    src = os.path.join(ROOT, "tests", "cxx_surface_test.cpp")
    hdr = os.path.join(ROOT, "include", "bvh", "v2", "b200_surface.h")
    lib = os.path.join(ROOT, "bvh_b200", "libbvh_c.so")
    if not os.path.exists(lib):
        import bvh_b200.build_ext as b
        b.build()
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(out):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=gnu++20", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-Wall", "-I", os.path.join(ROOT, "include"),
                               src, "-o", out, "-L", os.path.join(ROOT, "bvh_b200"), "-lbvh_c",
                               "-Wl,-rpath," + os.path.join(ROOT, "bvh_b200"), "-lpthread"])
    return out


@pytest.mark.parametrize("name", ["soup2k_f32", "grid2k_f32", "box12_f32", "soup1k_f64"])
def test_header_surface_matches_reference(exe, tmp_path, name):
    g = golden(name)
    dtype = g["tris"].dtype
    rec = np.dtype([("bounds", dtype, 6), ("index", np.uint32 if dtype == np.float32 else np.uint64)])
    nodes = np.zeros(g["ref_bounds"].shape[0], rec)
    nodes["bounds"] = g["ref_bounds"]
    nodes["index"] = g["ref_index"]
    assert rec.itemsize == (28 if dtype == np.float32 else 56)
    path = tmp_path / "case.bin"
    with open(path, "wb") as f:
        f.write(np.array([nodes.shape[0], g["ref_prim_ids"].shape[0], g["rays"].shape[0]], np.uint64).tobytes())
        f.write(nodes.tobytes())
        f.write(g["ref_prim_ids"].astype(np.uint64).tobytes())
        f.write(np.ascontiguousarray(g["tris"]).tobytes())
        f.write(np.ascontiguousarray(g["rays"]).tobytes())
        for mode in ("last", "any", "robust"):
            ids, t = g[f"{mode}_ids"], g[f"{mode}_t"]
            if mode == "any":       # reference driver semantics for any-hit ids are tie-rule independent per leaf order
                pass
            f.write(ids.astype(np.uint32).tobytes())
            f.write(t.astype(dtype).tobytes())
    out = subprocess.run([exe, str(path), "d" if dtype == np.float64 else "f"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
