"""ctypes bindings for the two CPU checkers — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

* ``Oracle``  -> oracle/liboracle.so      (plain-C restatement, bvh_oracle.c; always available)
* ``Ref``     -> oracle/_ref/libbvh_ref.so (the unmodified reference compiled in place; present when
  the build container had /root/reference — the .so travels to the GPU box, the sources do not)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product (bvh_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

ANY_HIT, ROBUST, TIE_LOWEST_ID = 1, 2, 4
INVALID_ID = 0xFFFFFFFF

_P = C.c_void_p
_SZ = C.c_size_t


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def _sfx(dtype):
    return "3f" if np.dtype(dtype) == np.float32 else "3d"


def _ct(dtype):
    return C.c_float if np.dtype(dtype) == np.float32 else C.c_double


def build_oracle_lib(force: bool = False) -> str:
    path = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("bvh_oracle.c", "bvh_oracle.h", "bvh_oracle_impl.inc")]
    if force or not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, path])
    return path


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libbvh_ref.so"))


class _Tree:
    """A BVH living inside one of the checker libraries."""

    def __init__(self, owner, handle, dtype):
        self.owner, self.handle, self.dtype = owner, handle, np.dtype(dtype)
        self._tris = None

    def __del__(self):
        try:
            self.owner._free(self)
        except Exception:
            pass

    @property
    def node_count(self):
        return self.owner._node_count(self)

    @property
    def prim_count(self):
        return self.owner._prim_count(self)

    def arrays(self):
        """(bounds[n,6] as minx,maxx,miny,maxy,minz,maxz; index_values[n] u64; prim_ids[p] u64)"""
        return self.owner._arrays(self)


class Oracle:
    """Plain-C restatement (oracle/bvh_oracle.c)."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle_lib())
        L = self.lib
        for s, ct in (("3f", C.c_float), ("3d", C.c_double)):
            getattr(L, f"orc_build_binned{s}").restype = _P
            getattr(L, f"orc_build_binned{s}").argtypes = [_P, _P, _SZ, _SZ, _SZ]
            getattr(L, f"orc_build_sweep{s}").restype = _P
            getattr(L, f"orc_build_sweep{s}").argtypes = [_P, _P, _SZ, _SZ, _SZ]
            getattr(L, f"orc_bvh_from_arrays{s}").restype = _P
            getattr(L, f"orc_bvh_from_arrays{s}").argtypes = [_P, _P, _SZ, _P, _SZ]
            getattr(L, f"orc_bvh_get_arrays{s}").argtypes = [_P, _P, _P, _P]
            getattr(L, f"orc_bvh_node_count{s}").restype = _SZ
            getattr(L, f"orc_bvh_node_count{s}").argtypes = [_P]
            getattr(L, f"orc_bvh_prim_count{s}").restype = _SZ
            getattr(L, f"orc_bvh_prim_count{s}").argtypes = [_P]
            getattr(L, f"orc_bvh_free{s}").argtypes = [_P]
            getattr(L, f"orc_precompute_tris{s}").argtypes = [_P, _P, _P]
            getattr(L, f"orc_trace{s}").argtypes = [_P, _P, _P, _SZ, C.c_uint, _P, _P, _P, _P, _P]
            getattr(L, f"orc_brute_force{s}").argtypes = [_P, _SZ, _P, _SZ, C.c_uint, _P, _P, _P, _P]
            getattr(L, f"orc_refit{s}").argtypes = [_P]
            getattr(L, f"orc_serialize{s}").restype = _SZ
            getattr(L, f"orc_serialize{s}").argtypes = [_P, _P, _SZ]
            getattr(L, f"orc_deserialize{s}").restype = _P
            getattr(L, f"orc_deserialize{s}").argtypes = [_P, _SZ]
            getattr(L, f"orc_check_invariants{s}").restype = C.c_int
            getattr(L, f"orc_check_invariants{s}").argtypes = [_P, _SZ]
            getattr(L, f"orc_sah_cost{s}").restype = C.c_double
            getattr(L, f"orc_sah_cost{s}").argtypes = [_P]
            getattr(L, f"orc_tri_bboxes_centers{s}").argtypes = [_P, _SZ, _P, _P]
        L.orc_morton_encode32.restype = C.c_uint32
        L.orc_morton_encode32.argtypes = [C.c_uint32] * 3
        L.orc_morton_encode64.restype = C.c_uint64
        L.orc_morton_encode64.argtypes = [C.c_uint64] * 3
        L.orc_fast_mul_add_is_fma.restype = C.c_int

    # -- helpers ------------------------------------------------------------------------------
    def _f(self, name, dtype):
        return getattr(self.lib, name + _sfx(dtype))

    def _free(self, tree):
        if tree.handle:
            self._f("orc_bvh_free", tree.dtype)(tree.handle)
            tree.handle = None

    def _node_count(self, tree):
        return self._f("orc_bvh_node_count", tree.dtype)(tree.handle)

    def _prim_count(self, tree):
        return self._f("orc_bvh_prim_count", tree.dtype)(tree.handle)

    def _arrays(self, tree):
        n, p = tree.node_count, tree.prim_count
        bounds = np.empty((n, 6), tree.dtype)
        idx = np.empty(n, np.uint64)
        ids = np.empty(p, np.uint64)
        self._f("orc_bvh_get_arrays", tree.dtype)(tree.handle, _ptr(bounds), _ptr(idx), _ptr(ids))
        return bounds, idx, ids

    # -- API ----------------------------------------------------------------------------------
    def tri_bboxes_centers(self, verts):
        verts = np.ascontiguousarray(verts)
        n = verts.shape[0]
        bb = np.empty((n, 6), verts.dtype)
        cc = np.empty((n, 3), verts.dtype)
        self._f("orc_tri_bboxes_centers", verts.dtype)(_ptr(verts), n, _ptr(bb), _ptr(cc))
        return bb, cc

    def build(self, bboxes, centers, quality="low", min_leaf=0, max_leaf=0):
        """quality 'low' -> BinnedSahBuilder, 'medium' -> SweepSahBuilder (serial DefaultBuilder paths)."""
        bboxes = np.ascontiguousarray(bboxes)
        centers = np.ascontiguousarray(centers, dtype=bboxes.dtype)
        fn = self._f("orc_build_binned" if quality == "low" else "orc_build_sweep", bboxes.dtype)
        return _Tree(self, fn(_ptr(bboxes), _ptr(centers), bboxes.shape[0], min_leaf, max_leaf), bboxes.dtype)

    def from_arrays(self, bounds, index_values, prim_ids):
        bounds = np.ascontiguousarray(bounds)
        idx = np.ascontiguousarray(index_values, dtype=np.uint64)
        ids = np.ascontiguousarray(prim_ids, dtype=np.uint64)
        h = self._f("orc_bvh_from_arrays", bounds.dtype)(_ptr(bounds), _ptr(idx), idx.shape[0], _ptr(ids), ids.shape[0])
        return _Tree(self, h, bounds.dtype)

    def set_triangles(self, tree, verts):
        verts = np.ascontiguousarray(verts, dtype=tree.dtype)
        tris = np.empty((tree.prim_count, 12), tree.dtype)
        self._f("orc_precompute_tris", tree.dtype)(tree.handle, _ptr(verts), _ptr(tris))
        tree._tris = tris

    def trace(self, tree, rays, flags=TIE_LOWEST_ID, stats=False):
        rays = np.ascontiguousarray(rays, dtype=tree.dtype)
        m = rays.shape[0]
        ids = np.empty(m, np.uint32)
        t, u, v = (np.empty(m, tree.dtype) for _ in range(3))
        st = np.zeros((m, 3), np.uint32) if stats else None
        self._f("orc_trace", tree.dtype)(tree.handle, _ptr(tree._tris), _ptr(rays), m, flags,
                                         _ptr(ids), _ptr(t), _ptr(u), _ptr(v), _ptr(st))
        return (ids, t, u, v, st) if stats else (ids, t, u, v)

    def brute_force(self, verts, rays, flags=0):
        verts = np.ascontiguousarray(verts)
        rays = np.ascontiguousarray(rays, dtype=verts.dtype)
        m = rays.shape[0]
        ids = np.empty(m, np.uint32)
        t, u, v = (np.empty(m, verts.dtype) for _ in range(3))
        self._f("orc_brute_force", verts.dtype)(_ptr(verts), verts.shape[0], _ptr(rays), m, flags,
                                                _ptr(ids), _ptr(t), _ptr(u), _ptr(v))
        return ids, t, u, v

    def refit(self, tree):
        self._f("orc_refit", tree.dtype)(tree.handle)

    def serialize(self, tree) -> bytes:
        fn = self._f("orc_serialize", tree.dtype)
        size = fn(tree.handle, None, 0)
        buf = np.empty(size, np.uint8)
        fn(tree.handle, _ptr(buf), size)
        return buf.tobytes()

    def deserialize(self, data: bytes, dtype=np.float32):
        buf = np.frombuffer(data, np.uint8).copy()
        return _Tree(self, self._f("orc_deserialize", dtype)(_ptr(buf), buf.shape[0]), dtype)

    def check_invariants(self, tree, max_leaf_size=0) -> int:
        return self._f("orc_check_invariants", tree.dtype)(tree.handle, max_leaf_size)

    def sah_cost(self, tree) -> float:
        return self._f("orc_sah_cost", tree.dtype)(tree.handle)

    def morton_encode(self, x, y, z, bits=32):
        return (self.lib.orc_morton_encode32 if bits == 32 else self.lib.orc_morton_encode64)(x, y, z)

    def fast_mul_add_is_fma(self) -> bool:
        return bool(self.lib.orc_fast_mul_add_is_fma())


QUALITY = {"low": 0, "medium": 1, "high": 2}


class Ref:
    """The unmodified reference behind oracle/ref_driver.cpp (oracle/_ref/libbvh_ref.so)."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libbvh_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with `make -C oracle ref` where /root/reference exists)")
        self.lib = L = C.CDLL(path)
        for s, ct in (("3f", C.c_float), ("3d", C.c_double)):
            getattr(L, f"ref_build{s}").restype = _P
            getattr(L, f"ref_build{s}").argtypes = [_P, _P, _SZ, C.c_int, C.c_int, _SZ, _SZ]
            getattr(L, f"ref_time_build{s}").restype = C.c_double
            getattr(L, f"ref_time_build{s}").argtypes = [_P, _P, _SZ, C.c_int, C.c_int]
            getattr(L, f"ref_destroy{s}").argtypes = [_P]
            getattr(L, f"ref_node_count{s}").restype = _SZ
            getattr(L, f"ref_node_count{s}").argtypes = [_P]
            getattr(L, f"ref_prim_count{s}").restype = _SZ
            getattr(L, f"ref_prim_count{s}").argtypes = [_P]
            getattr(L, f"ref_get_nodes{s}").argtypes = [_P, _P, _P]
            getattr(L, f"ref_get_prim_ids{s}").argtypes = [_P, _P]
            getattr(L, f"ref_from_nodes{s}").restype = _P
            getattr(L, f"ref_from_nodes{s}").argtypes = [_P, _P, _SZ, _P, _SZ]
            getattr(L, f"ref_set_triangles{s}").argtypes = [_P, _P]
            getattr(L, f"ref_trace{s}").restype = C.c_double
            getattr(L, f"ref_trace{s}").argtypes = [_P, _P, _SZ, C.c_uint, C.c_int, _P, _P, _P, _P, _P]
            getattr(L, f"ref_refit{s}").argtypes = [_P]
            getattr(L, f"ref_serialize{s}").restype = _SZ
            getattr(L, f"ref_serialize{s}").argtypes = [_P, _P, _SZ]
            getattr(L, f"ref_deserialize{s}").restype = _P
            getattr(L, f"ref_deserialize{s}").argtypes = [_P, _SZ]
            getattr(L, f"ref_tri_bboxes_centers{s}").argtypes = [_P, _SZ, _P, _P]
        L.ref_thread_count.restype = C.c_int
        L.ref_thread_count.argtypes = [C.c_int]
        L.ref_fast_mul_add_is_fma.restype = C.c_int
        L.ref_morton_encode32.restype = C.c_uint32
        L.ref_morton_encode32.argtypes = [C.c_uint32] * 3
        L.ref_morton_encode64.restype = C.c_uint64
        L.ref_morton_encode64.argtypes = [C.c_uint64] * 3

    def _f(self, name, dtype):
        return getattr(self.lib, name + _sfx(dtype))

    def _free(self, tree):
        if tree.handle:
            self._f("ref_destroy", tree.dtype)(tree.handle)
            tree.handle = None

    def _node_count(self, tree):
        return self._f("ref_node_count", tree.dtype)(tree.handle)

    def _prim_count(self, tree):
        return self._f("ref_prim_count", tree.dtype)(tree.handle)

    def _arrays(self, tree):
        n, p = tree.node_count, tree.prim_count
        bounds = np.empty((n, 6), tree.dtype)
        idx = np.empty(n, np.uint64)
        ids = np.empty(p, np.uint64)
        self._f("ref_get_nodes", tree.dtype)(tree.handle, _ptr(bounds), _ptr(idx))
        self._f("ref_get_prim_ids", tree.dtype)(tree.handle, _ptr(ids))
        return bounds, idx, ids

    def tri_bboxes_centers(self, verts):
        verts = np.ascontiguousarray(verts)
        n = verts.shape[0]
        bb = np.empty((n, 6), verts.dtype)
        cc = np.empty((n, 3), verts.dtype)
        self._f("ref_tri_bboxes_centers", verts.dtype)(_ptr(verts), n, _ptr(bb), _ptr(cc))
        return bb, cc

    def build(self, bboxes, centers, quality="high", threads=-1, min_leaf=0, max_leaf=0):
        """threads < 0: serial overload (default_builder.h:49-62); 0: pool with hardware_concurrency;
        k > 0: pool with k threads (default_builder.h:33-46)."""
        bboxes = np.ascontiguousarray(bboxes)
        centers = np.ascontiguousarray(centers, dtype=bboxes.dtype)
        h = self._f("ref_build", bboxes.dtype)(_ptr(bboxes), _ptr(centers), bboxes.shape[0],
                                               QUALITY[quality], threads, min_leaf, max_leaf)
        return _Tree(self, h, bboxes.dtype)

    def time_build(self, bboxes, centers, quality="high", threads=0) -> float:
        bboxes = np.ascontiguousarray(bboxes)
        centers = np.ascontiguousarray(centers, dtype=bboxes.dtype)
        return self._f("ref_time_build", bboxes.dtype)(_ptr(bboxes), _ptr(centers), bboxes.shape[0],
                                                    QUALITY[quality], threads)

    def from_arrays(self, bounds, index_values, prim_ids):
        bounds = np.ascontiguousarray(bounds)
        idx = np.ascontiguousarray(index_values, dtype=np.uint64)
        ids = np.ascontiguousarray(prim_ids, dtype=np.uint64)
        h = self._f("ref_from_nodes", bounds.dtype)(_ptr(bounds), _ptr(idx), idx.shape[0], _ptr(ids), ids.shape[0])
        return _Tree(self, h, bounds.dtype)

    def set_triangles(self, tree, verts):
        verts = np.ascontiguousarray(verts, dtype=tree.dtype)
        self._f("ref_set_triangles", tree.dtype)(tree.handle, _ptr(verts))
        tree._tris = True

    def trace(self, tree, rays, flags=TIE_LOWEST_ID, threads=-1, stats=False, outputs=True):
        rays = np.ascontiguousarray(rays, dtype=tree.dtype)
        m = rays.shape[0]
        ids = np.empty(m, np.uint32) if outputs else None
        t, u, v = ((np.empty(m, tree.dtype) for _ in range(3)) if outputs else (None, None, None))
        st = np.zeros((m, 3), np.uint32) if stats else None
        secs = self._f("ref_trace", tree.dtype)(tree.handle, _ptr(rays), m, flags, threads,
                                               _ptr(ids), _ptr(t), _ptr(u), _ptr(v), _ptr(st))
        self.last_trace_seconds = secs
        return (ids, t, u, v, st) if stats else (ids, t, u, v)

    def refit(self, tree):
        self._f("ref_refit", tree.dtype)(tree.handle)

    def serialize(self, tree) -> bytes:
        fn = self._f("ref_serialize", tree.dtype)
        size = fn(tree.handle, None, 0)
        buf = np.empty(size, np.uint8)
        fn(tree.handle, _ptr(buf), size)
        return buf.tobytes()

    def deserialize(self, data: bytes, dtype=np.float32):
        buf = np.frombuffer(data, np.uint8).copy()
        return _Tree(self, self._f("ref_deserialize", dtype)(_ptr(buf), buf.shape[0]), dtype)

    def thread_count(self, threads=0) -> int:
        return self.lib.ref_thread_count(threads)

    def fast_mul_add_is_fma(self) -> bool:
        return bool(self.lib.ref_fast_mul_add_is_fma())

    def morton_encode(self, x, y, z, bits=32):
        return (self.lib.ref_morton_encode32 if bits == 32 else self.lib.ref_morton_encode64)(x, y, z)
