"""Prototype (CPU, host emulation of the device build): a second SAH level over the roots of the rebuilt treelets.
Usage: python scripts/two_level_prototype.py grid|soup N.  Result at 1M triangles (300x300 primary rays): grid 25.60 -> 24.67 inner
steps per ray (-3.6 %) with super-treelets of <= 64 units, -6.9 % with one SAH level over all units; soup -0.6 % / -0.2 %.
Not adopted: the headline scene gains nothing."""
import sys, time
import numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tests.helpers import HostEmul
from bvh_b200 import scenes

def half_area(b):
    d = b[1::2] - b[0::2]
    return (d[0] + d[1]) * d[2] + d[0] * d[1]

def run(kind, n, U=64, weight="units", img=300):
    emul = HostEmul()
    tris = scenes.make_mesh(kind, n)
    t = emul.build(tris=tris, quality="high")
    bounds, index = emul.compact(t)
    prim_ids = t["prim_ids"]
    N = bounds.shape[0]
    count = (index & 15).astype(np.int64)
    first = (index >> 4).astype(np.int64)
    # prim counts and unit counts bottom-up (children have larger indices? not guaranteed) -> iterative DFS post-order
    prims = np.zeros(N, np.int64); parent = np.full(N, -1, np.int64)
    order = []
    stack = [0]
    while stack:
        i = stack.pop(); order.append(i)
        if count[i] == 0:
            c = first[i]; parent[c] = i; parent[c + 1] = i; stack += [c, c + 1]
    for i in reversed(order):
        prims[i] = count[i] if count[i] else prims[first[i]] + prims[first[i] + 1]
    is_unit = np.zeros(N, bool)
    units = np.zeros(N, np.int64)
    for i in reversed(order):
        if prims[i] <= 64:
            is_unit[i] = parent[i] < 0 or prims[parent[i]] > 64
            units[i] = 1
        else:
            units[i] = units[first[i]] + units[first[i] + 1]
    supers = [i for i in order if prims[i] > 64 and 3 <= units[i] <= U and (parent[i] < 0 or units[parent[i]] > U)]
    nb = bounds.astype(np.float64).copy(); ni = index.copy()
    def collect(root):
        us, pairs, st = [], [], [root]
        while st:
            i = st.pop()
            if is_unit[i]: us.append(i)
            else:
                pairs.append(first[i]); st += [first[i] + 1, first[i]]
        return us, pairs
    rebuilt = 0
    for root in supers:
        us, pairs = collect(root)
        ub = np.array([bounds[u] for u in us], np.float64)          # unit boxes
        urec = [(bounds[u].copy(), index[u]) for u in us]
        w = np.ones(len(us)) if weight == "units" else np.array([prims[u] for u in us], np.float64)
        cen = (ub[:, 0::2] + ub[:, 1::2]) * 0.5
        free = list(pairs)
        def box_of(ids):
            b = ub[ids]
            o = np.empty(6); o[0::2] = b[:, 0::2].min(0); o[1::2] = b[:, 1::2].max(0); return o
        def build(ids, slot):
            if len(ids) == 1:
                nb[slot] = urec[ids[0]][0]; ni[slot] = urec[ids[0]][1]; return
            best = (np.inf, None, None)
            for a in range(3):
                o = sorted(ids, key=lambda u: (cen[u, a], u))
                # prefix / suffix
                pre = []; cur = None
                for u in o:
                    bb = ub[u]
                    cur = bb.copy() if cur is None else np.concatenate([np.minimum(cur[0::2], bb[0::2]), np.maximum(cur[1::2], bb[1::2])]).reshape(2, 3).T.reshape(-1)
                    pre.append(cur.copy())
                suf = [None] * len(o); cur = None
                for k in range(len(o) - 1, -1, -1):
                    bb = ub[o[k]]
                    cur = bb.copy() if cur is None else np.concatenate([np.minimum(cur[0::2], bb[0::2]), np.maximum(cur[1::2], bb[1::2])]).reshape(2, 3).T.reshape(-1)
                    suf[k] = cur.copy()
                wl = np.cumsum([w[u] for u in o])
                for k in range(1, len(o)):
                    c = half_area(pre[k - 1]) * wl[k - 1] + half_area(suf[k]) * (wl[-1] - wl[k - 1])
                    if c < best[0]: best = (c, o, k)
            _, o, k = best
            L, R = o[:k], o[k:]
            bl, br = box_of(L), box_of(R)
            if half_area(bl) < half_area(br): L, R, bl, br = R, L, br, bl
            pair = free.pop()
            b = box_of(ids)
            nb[slot] = b; ni[slot] = np.uint64(pair) << np.uint64(4)
            build(L, pair); build(R, pair + 1)
        build(list(range(len(us))), root)
        rebuilt += 1
    rays = scenes.make_primary(kind, img, img)
    flags = 4
    base = emul.from_reference(bounds, index, prim_ids, tris)
    new = emul.from_reference(nb.astype(np.float32), ni, prim_ids, tris)
    a = emul.trace(base, rays, flags); b = emul.trace(new, rays, flags)
    same = (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all()
    print(f"{kind}-{n} U={U} weight={weight}: supers {rebuilt}, inner steps {a[4][:,0].mean():.2f} -> {b[4][:,0].mean():.2f} ({(b[4][:,0].mean()/a[4][:,0].mean()-1)*100:+.1f}%), tri tests {a[4][:,2].mean():.2f} -> {b[4][:,2].mean():.2f}, hits identical {same}")

if __name__ == "__main__":
    kind, n = sys.argv[1], int(sys.argv[2])
    for U in (64, 512, 100000000):
        for wt in ("prims",):
            t0 = time.time(); run(kind, n, U, wt); print(f"  ({time.time()-t0:.0f} s)")
