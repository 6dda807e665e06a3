"""Seeded synthetic meshes and ray batches.

Everything here is plain numpy on the host: the CUDA path, the oracle and the compiled reference
all consume the *same bytes*, so generator details (numpy's MT19937 stream rather than libstdc++'s
distributions) do not enter parity.  Triangles are ``(n, 9)`` arrays ``p0 p1 p2``; rays are
``(m, 8)`` arrays ``org3 dir3 tmin tmax`` — the memory layout of ``bvh_ray3f`` / ``bvh_ray3d``
(reference c_api/bvh.h:70-73) and of ``Ray<T,3>`` (reference ray.h:16-27).
"""
from __future__ import annotations

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)
DBL_MAX = float(np.finfo(np.float64).max)


def _tmax(dtype) -> float:
    return FLT_MAX if np.dtype(dtype) == np.float32 else DBL_MAX


def soup(n: int, seed: int = 12345, dtype=np.float32) -> np.ndarray:
    """n random small triangles in the unit cube: centre c~U[0,1)^3, vertices c+(U-0.5)*e with
    e = 1.5/cbrt(n).  Tie-free with probability 1, so hit ids are well defined for any tree."""
    rng = np.random.RandomState(seed)
    c = rng.random_sample((n, 1, 3))
    e = 1.5 / np.cbrt(float(n))
    v = c + (rng.random_sample((n, 3, 3)) - 0.5) * e
    return np.ascontiguousarray(v.reshape(n, 9).astype(dtype))


def grid(n: int, dtype=np.float32) -> np.ndarray:
    """Connected sine height-field of ~n triangles (2*k*k with k=floor(sqrt(n/2))): vertices
    (i/k, 0.05*sin(12 i/k)*cos(9 j/k), j/k), two triangles per cell.  Shared edges and vertices give
    exact-t ties, which is what the canonical lowest-id tie-break is for."""
    k = int(np.floor(np.sqrt(n / 2.0)))
    i = np.arange(k + 1, dtype=np.float64) / k
    x, z = np.meshgrid(i, i, indexing="ij")
    y = 0.05 * np.sin(12.0 * x) * np.cos(9.0 * z)
    p = np.stack([x, y, z], axis=-1).astype(dtype)          # (k+1, k+1, 3)
    p00, p10, p01, p11 = p[:-1, :-1], p[1:, :-1], p[:-1, 1:], p[1:, 1:]
    t0 = np.concatenate([p00, p10, p11], axis=-1)            # (k, k, 9)
    t1 = np.concatenate([p00, p11, p01], axis=-1)
    tris = np.stack([t0, t1], axis=2).reshape(-1, 9)
    return np.ascontiguousarray(tris)


def box12(dtype=np.float32) -> np.ndarray:
    """Unit cube [0,1]^3 as 12 triangles (BASELINE config 1)."""
    c = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], dtype=dtype)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = []
    for a, b, cc, d in quads:
        tris.append(np.concatenate([c[a], c[b], c[cc]]))
        tris.append(np.concatenate([c[a], c[cc], c[d]]))
    return np.ascontiguousarray(np.stack(tris).astype(dtype))


def _normalize(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.sqrt((v * v).sum())


def primary_rays(width: int, height: int, eye, direction, up=(0.0, 1.0, 0.0), dtype=np.float32,
                 pixel_offset: float = 0.0, y_begin: int = 0, y_end: int | None = None) -> np.ndarray:
    """Pinhole camera of the reference benchmark (test/benchmark.cpp:340-358): unnormalised
    ``dir + u*right + v*up`` with ``u = 2x/W-1``, ``v = 2y/H-1``, tmin 0, tmax = max scalar, rays in
    row-major pixel order.  ``y_begin:y_end`` selects a band of rows (ray sharding across GPUs)."""
    dt = np.dtype(dtype).type
    d = _normalize(direction).astype(dtype)
    upv = np.asarray(up, dtype=dtype)
    right = np.cross(d.astype(np.float64), upv.astype(np.float64))
    right = _normalize(right).astype(dtype)
    upv = np.cross(right.astype(np.float64), d.astype(np.float64)).astype(dtype)
    if y_end is None:
        y_end = height
    xs = (np.arange(width, dtype=dtype) + dt(pixel_offset))
    ys = (np.arange(y_begin, y_end, dtype=dtype) + dt(pixel_offset))
    u = dt(2) * xs / dt(width) - dt(1)
    v = dt(2) * ys / dt(height) - dt(1)
    rays = np.empty((y_end - y_begin, width, 8), dtype=dtype)
    rays[..., 0:3] = np.asarray(eye, dtype=dtype)
    rays[..., 3:6] = d + u[None, :, None] * right + v[:, None, None] * upv
    rays[..., 6] = 0
    rays[..., 7] = _tmax(dtype)
    return rays.reshape(-1, 8)


# Cameras used by the tests and the bench for each mesh family.
CAMERAS = {
    "soup": dict(eye=(0.5, 0.5, -0.55), direction=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0)),
    "grid": dict(eye=(0.5, 0.35, 0.2), direction=(0.0, -1.0, 0.4472136), up=(0.0, 1.0, 0.0)),
    "box12": dict(eye=(0.45, 0.55, -0.6), direction=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0)),
}


def incoherent_rays(tris: np.ndarray, m: int, seed: int = 12345, tmax: float = 0.25) -> np.ndarray:
    """AO-style incoherent rays: a random point on a random triangle, pushed 1e-3 along the unit
    normal, with a cosine-weighted direction in that normal's hemisphere and a short tmax.  The
    batch is already in random order (no two consecutive rays are spatially related)."""
    dtype = tris.dtype
    rng = np.random.RandomState(seed)
    n = tris.shape[0]
    ids = rng.randint(0, n, size=m)
    t = tris[ids].astype(np.float64).reshape(m, 3, 3)
    r1 = np.sqrt(rng.random_sample(m))
    r2 = rng.random_sample(m)
    b0, b1, b2 = 1.0 - r1, r1 * (1.0 - r2), r1 * r2
    p = b0[:, None] * t[:, 0] + b1[:, None] * t[:, 1] + b2[:, None] * t[:, 2]
    nrm = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    ln = np.sqrt((nrm * nrm).sum(axis=1, keepdims=True))
    nrm = np.where(ln > 0, nrm / np.maximum(ln, 1e-300), np.array([0.0, 1.0, 0.0]))
    flip = rng.random_sample(m) < 0.5
    nrm[flip] *= -1.0
    # orthonormal basis around nrm
    a = np.where(np.abs(nrm[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    tx = np.cross(nrm, a)
    tx /= np.sqrt((tx * tx).sum(axis=1, keepdims=True))
    ty = np.cross(nrm, tx)
    s1, s2 = rng.random_sample(m), rng.random_sample(m)
    rr, phi = np.sqrt(s1), 2.0 * np.pi * s2
    d = (rr * np.cos(phi))[:, None] * tx + (rr * np.sin(phi))[:, None] * ty + np.sqrt(1.0 - s1)[:, None] * nrm
    rays = np.empty((m, 8), dtype=dtype)
    rays[:, 0:3] = p + 1e-3 * nrm
    rays[:, 3:6] = d
    rays[:, 6] = 0
    rays[:, 7] = tmax
    return rays


def make_mesh(kind: str, n: int, seed: int = 12345, dtype=np.float32) -> np.ndarray:
    if kind == "soup":
        return soup(n, seed, dtype)
    if kind == "grid":
        return grid(n, dtype)
    if kind == "box12":
        return box12(dtype)
    raise ValueError(f"unknown mesh kind {kind!r}")


def make_primary(kind: str, width: int, height: int, dtype=np.float32, **kw) -> np.ndarray:
    return primary_rays(width, height, dtype=dtype, **CAMERAS[kind], **kw)
