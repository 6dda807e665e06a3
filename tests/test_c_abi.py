"""The C-ABI library loads and exports every symbol include/*.h declares (no compute: no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bvh_b200", "libbvh_c.so")

LEGACY_PER_SUFFIX = ["build", "destroy", "save", "load", "get_node", "get_prim_id", "get_prim_count", "get_node_count",
                     "append_node", "remove_last_node", "refit", "optimize", "intersect_ray_any",
                     "intersect_ray_any_robust", "intersect_ray", "intersect_ray_robust"]
NODE_PER_SUFFIX = ["is_leaf", "get_prim_count", "set_prim_count", "get_first_id", "set_first_id", "get_bbox", "set_bbox"]
BATCHED_3D = ["build_triangles", "set_triangles", "refit_triangles", "intersect_rays", "intersect_rays_gather", "intersect_rays_stats", "sync", "get_depth", "get_property", "get_prim_ids"]
RUNTIME = ["bvh_last_error", "bvh_cuda_device_count", "bvh_cuda_set_device", "bvh_cuda_set_stream", "bvh_cuda_reset_stream", "bvh_host_alloc",
           "bvh_host_free", "bvh_cuda_trim", "bvh_set_option", "bvh_optimize_nodes", "bvh_thread_pool_create", "bvh_thread_pool_destroy"]


def expected_symbols():
    names = list(RUNTIME)
    for s in ("2f", "3f", "2d", "3d"):
        names += [f"bvh{s}_{f}" for f in LEGACY_PER_SUFFIX]
        names += [f"bvh_node{s}_{f}" for f in NODE_PER_SUFFIX]
    for s in ("3f", "3d"):
        names += [f"bvh{s}_{f}" for f in BATCHED_3D]
    return names


@pytest.fixture(scope="module")
def library():
    if not os.path.exists(LIB):
        import bvh_b200.build_ext as b
        b.build()
    return C.CDLL(LIB)


def test_reference_abi_has_94_symbols():
    legacy = [n for n in expected_symbols() if not any(n.endswith("_" + b) for b in BATCHED_3D) and n not in RUNTIME[:10]]
    assert len(legacy) == 94          # 23 functions x 4 suffixes + 2 pool functions


def test_every_declared_symbol_is_exported(library):
    for name in expected_symbols():
        assert hasattr(library, name), f"{name} is declared in include/ but not exported"


def test_headers_declare_what_we_expect():
    """The macro-stamped headers expand (with gcc -E) to exactly the expected set of BVH_API functions."""
    src = "#include <bvh_b200.h>\n"
    out = subprocess.run(["gcc", "-E", "-P", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"], input=src,
                         capture_output=True, text=True, check=True).stdout
    declared = set(re.findall(r"\b(bvh(?:_node)?(?:[23][fd])?_[a-z_]+)\s*\(", out))
    assert declared == set(expected_symbols())


def test_header_compiles_as_c11_and_pod_sizes():
    src = r"""
    #include <bvh_b200.h>
    _Static_assert(sizeof(struct bvh_vec3f) == 12, "vec3f");
    _Static_assert(sizeof(struct bvh_bbox3f) == 24, "bbox3f");
    _Static_assert(sizeof(struct bvh_ray3f) == 32, "ray3f");
    _Static_assert(sizeof(struct bvh_ray3d) == 64, "ray3d");
    _Static_assert(sizeof(struct bvh_build_config) == 32, "config");
    _Static_assert(sizeof(struct bvh_hit3f) == 16, "hit3f");
    _Static_assert(sizeof(struct bvh_hit3d) == 32, "hit3d");
    _Static_assert(sizeof(struct bvh_ray_stats) == 12, "stats");
    int main(void) { return BVH_ROOT_INDEX; }
    """
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                   input=src, text=True, check=True)


def test_no_device_means_loud_failure(library):
    """Without a GPU the batched path must fail (there is no CPU fallback)."""
    library.bvh_cuda_device_count.restype = C.c_int
    if library.bvh_cuda_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    library.bvh3f_build_triangles.restype = C.c_void_p
    library.bvh3f_build_triangles.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint]
    verts = (C.c_float * 9)(0, 0, 0, 1, 0, 0, 0, 1, 0)
    assert library.bvh3f_build_triangles(verts, 1, None, 0) is None
    library.bvh_last_error.restype = C.c_char_p
    assert library.bvh_last_error()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under bvh_b200/ or include/ may reference it."""
    for base in ("bvh_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "pyoracle" not in text and "liboracle" not in text and "bvh_oracle" not in text and "_ref/" not in text, f


def test_set_option_knows_every_documented_switch(library):
    """bvh_set_option: every switch include/bvh_b200.h documents is accepted (needs no GPU), unknown names are an error."""
    import re
    header = open(os.path.join(ROOT, "include", "bvh_b200.h")).read()
    doc = header[header.index("Process-wide switches"):header.index("BVH_API int bvh_set_option")]
    names = re.findall(r'"([a-z0-9_]+)"', doc)
    assert {"refill_min", "inner_budget", "chunk_rays", "variant", "sah_treelets", "smem_carveout"} <= set(names)
    library.bvh_set_option.argtypes = [C.c_char_p, C.c_long]
    defaults = {"morton_bits": 0, "sah_treelets": -1, "hierarchy": 128, "e2e_chunks": 0, "variant": 0, "use_wide": 0, "inner_budget": 8,
                "refill_min": 8, "chunk_rays": 64, "wide_budget": 4, "watchdog": 1 << 26, "gather_staging": 1, "sort_onesweep": 1,
                "treelet_blocks": 3, "stack_round": 2, "smem_carveout": -1}
    for name in names:
        assert name in defaults, f"{name}: documented but its default is not listed in this test"
        assert library.bvh_set_option(name.encode(), defaults[name]) == 0, name
    assert library.bvh_set_option(b"no_such_switch", 1) != 0
    library.bvh_last_error.restype = C.c_char_p
    assert b"unknown option" in library.bvh_last_error()
