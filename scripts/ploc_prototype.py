"""Quality evaluation of a PLOC-style builder (parallel locally-ordered clustering over the Morton order)
against the LBVH pass and the reference's High-quality tree, on the CPU: traversal steps per ray, measured by
the host emulation of the traversal kernel.  Prototype only (numpy); the product path is CUDA.

    python scripts/ploc_prototype.py [n_tris] [radius]
"""
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import scenes
from tests.helpers import HostEmul
from oracle.pyoracle import Ref, TIE_LOWEST_ID


def half_area(lo, hi):
    d = hi - lo
    return (d[:, 0] + d[:, 1]) * d[:, 2] + d[:, 0] * d[:, 1]


def ploc(lo, hi, radius, max_leaf=8):
    """lo/hi: leaf boxes in Morton order.  Returns (left, right, blo, bhi, count, is_leaf, root) over node ids;
    leaves are ids 0..n-1."""
    n = lo.shape[0]
    cap = 2 * n
    nlo = np.zeros((cap, 3), np.float32); nhi = np.zeros((cap, 3), np.float32)
    left = np.full(cap, -1, np.int64); right = np.full(cap, -1, np.int64)
    count = np.zeros(cap, np.int64); cost = np.zeros(cap, np.float32); collapsed = np.zeros(cap, bool)
    nlo[:n] = lo; nhi[:n] = hi; count[:n] = 1; cost[:n] = half_area(lo, hi); collapsed[:n] = True
    ids = np.arange(n)
    next_id = n
    it = 0
    while ids.size > 1:
        C = ids.size
        clo, chi = nlo[ids], nhi[ids]
        best = np.full(C, np.inf, np.float32); nn = np.full(C, -1, np.int64)
        for off in list(range(-radius, 0)) + list(range(1, radius + 1)):
            j = np.arange(C) + off
            ok = (j >= 0) & (j < C)
            jj = np.clip(j, 0, C - 1)
            a = half_area(np.minimum(clo, clo[jj]), np.maximum(chi, chi[jj]))
            a = np.where(ok, a, np.inf).astype(np.float32)
            better = a < best          # ties -> earlier offset (smaller j)
            best = np.where(better, a, best); nn = np.where(better, jj, nn)
        i = np.arange(C)
        mutual = nn[nn] == i
        lead = mutual & (i < nn)
        gone = mutual & (i > nn)
        li = i[lead]; lj = nn[lead]
        m = li.size
        a_id, b_id = ids[li], ids[lj]
        swap = half_area(nlo[a_id], nhi[a_id]) < half_area(nlo[b_id], nhi[b_id])
        l_id = np.where(swap, b_id, a_id); r_id = np.where(swap, a_id, b_id)
        new = next_id + np.arange(m)
        next_id += m
        left[new] = l_id; right[new] = r_id
        nlo[new] = np.minimum(nlo[a_id], nlo[b_id]); nhi[new] = np.maximum(nhi[a_id], nhi[b_id])
        count[new] = count[a_id] + count[b_id]
        area = half_area(nlo[new], nhi[new])
        split_cost = area + cost[a_id] + cost[b_id]
        leaf_cost = area * count[new]
        col = (count[new] <= max_leaf) & (leaf_cost <= split_cost)
        collapsed[new] = col
        cost[new] = np.where(col, leaf_cost, split_cost)
        ids = ids.copy(); ids[li] = new
        ids = ids[~gone]
        it += 1
    return left, right, nlo, nhi, count, collapsed, int(ids[0]), it


def to_reference_layout(left, right, nlo, nhi, count, collapsed, root, order):
    """DFS; emits reference-layout arrays (bounds n x 6 [minx,maxx,...], index values, prim_ids)."""
    bounds, index, prim_ids = [], [], []
    def emit(nid):
        bounds.append([nlo[nid, 0], nhi[nid, 0], nlo[nid, 1], nhi[nid, 1], nlo[nid, 2], nhi[nid, 2]])
        index.append(0)
        return len(index) - 1
    def leaves_of(nid):
        out, st = [], [nid]
        while st:
            x = st.pop()
            if left[x] < 0: out.append(x)
            else: st.append(right[x]); st.append(left[x])
        return out
    emit(root)
    stack = [(root, 0)]
    while stack:
        nid, dst = stack.pop()
        if collapsed[nid]:
            ls = leaves_of(nid)
            index[dst] = (len(prim_ids) << 4) | len(ls)
            prim_ids.extend(order[l] for l in ls)
        else:
            a = emit(left[nid]); b = emit(right[nid])
            index[dst] = a << 4
            stack.append((right[nid], b)); stack.append((left[nid], a))
    return np.array(bounds, np.float32), np.array(index, np.uint64), np.array(prim_ids, np.uint32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    radii = [int(x) for x in sys.argv[2:]] or [8, 16]
    emul, ref = HostEmul(), Ref()
    for kind in ("soup", "grid"):
        tris = scenes.soup(n, seed=12345) if kind == "soup" else scenes.grid(n)
        cam = scenes.CAMERAS[kind]
        rays = scenes.primary_rays(256, 256, **cam)
        lb = emul.build(tris=tris)
        ids, t, u, v, st = emul.trace(lb, rays, TIE_LOWEST_ID)
        print(f"{kind} n={tris.shape[0]}: LBVH       inner {st[:,0].mean():7.2f} tris {st[:,1].mean():6.2f}  hit {np.mean(ids != 0xFFFFFFFF):.3f}")
        bb, cc = ref.tri_bboxes_centers(tris)
        for q in ("high", "low"):
            rt = ref.build(bb, cc, quality=q)
            ref.set_triangles(rt, tris)
            out = ref.trace(rt, rays, TIE_LOWEST_ID, stats=True)
            rst = out[-1]
            print(f"{kind}: reference {q:5s} inner {rst[:,0].mean():7.2f} tris {rst[:,1].mean():6.2f}")
        order = lb["prim_ids"]
        v9 = tris.reshape(-1, 3, 3)[order]
        lo, hi = v9.min(axis=1), v9.max(axis=1)
        for r in radii:
            tree = ploc(lo, hi, r)
            b, ix, pid = to_reference_layout(*tree[:7], order)
            dev = emul.from_reference(b, ix, pid, tris)
            ids2, t2, u2, v2, st2 = emul.trace(dev, rays, TIE_LOWEST_ID)
            assert np.array_equal(ids, ids2) or np.mean(ids != ids2) < 1e-3, np.mean(ids != ids2)
            print(f"{kind}: PLOC r={r:2d}  inner {st2[:,0].mean():7.2f} tris {st2[:,1].mean():6.2f}  iterations {tree[7]} nodes {b.shape[0]}")


if __name__ == "__main__":
    main()
