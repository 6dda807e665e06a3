"""CPU check of the device code's logic: the headers the CUDA kernels are made of are compiled by g++
(tests/host_emul.cpp) and compared with the oracle.  Covers the LBVH bottom-up pass (structure
invariants, box unions, SAH leaf collapse, depth bound) and the traversal stack machine (same-tree
bit-exactness incl. step counters; cross-tree exactness under the canonical tie-break)."""
import numpy as np
import pytest

from bvh_b200 import scenes
from oracle.pyoracle import ANY_HIT, ROBUST, TIE_LOWEST_ID
from tests.conftest import golden
from tests.helpers import INVALID, assert_hits_equal

MODES = (("lowest", TIE_LOWEST_ID), ("last", 0), ("any", ANY_HIT | TIE_LOWEST_ID), ("robust", ROBUST | TIE_LOWEST_ID))


def _tree_depth(index_values):
    depth, stack = 0, [(0, 0)]
    while stack:
        i, d = stack.pop()
        v = int(index_values[i])
        if v & 15:
            continue
        depth = max(depth, d + 1)
        stack.append((v >> 4, d + 1))
        stack.append(((v >> 4) + 1, d + 1))
    return depth


@pytest.mark.parametrize("kind,n,dtype,bits", [
    ("soup", 1, np.float32, 30), ("soup", 2, np.float32, 30), ("soup", 3, np.float32, 30), ("soup", 17, np.float32, 30),
    ("soup", 5000, np.float32, 30), ("soup", 5000, np.float32, 63), ("grid", 5000, np.float32, 30),
    ("box12", 12, np.float32, 30), ("soup", 2000, np.float64, 30), ("soup", 2000, np.float64, 63)])
def test_lbvh_structure_and_parity(emul, oracle, kind, n, dtype, bits):
    tris = scenes.make_mesh(kind, n, dtype=dtype)
    n = tris.shape[0]
    tree = emul.build(tris=tris, morton_bits=bits)
    bounds, index_values = emul.compact(tree)
    otree = oracle.from_arrays(bounds, index_values, tree["prim_ids"])
    assert oracle.check_invariants(otree, 8) == 0
    assert tree["depth"] == _tree_depth(index_values)
    assert sorted(tree["prim_ids"].tolist()) == list(range(n))
    # inner boxes are exactly what the reference's refit computes from the leaf boxes
    before = otree.arrays()[0]
    oracle.refit(otree)
    assert (otree.arrays()[0] == before).all()
    # SATO: left child has the larger (or equal) half-area
    for i in np.nonzero((index_values & 15) == 0)[0]:
        f = int(index_values[i]) >> 4
        d = bounds[[f, f + 1], 1::2] - bounds[[f, f + 1], 0::2]
        area = (d[:, 0] + d[:, 1]) * d[:, 2] + d[:, 0] * d[:, 1]
        assert area[0] >= area[1]
    # traversal
    oracle.set_triangles(otree, tris)
    rays = scenes.make_primary(kind, 64, 64, dtype=dtype)
    for _, flags in MODES:
        got = emul.trace(tree, rays, flags)
        want = oracle.trace(otree, rays, flags=flags, stats=True)
        assert_hits_equal(got[:4], want[:4], f"{kind}/{n}/{flags}")
        assert (got[4] == want[4]).all()
    if n * rays.shape[0] <= 4_000_000:
        assert_hits_equal(emul.trace(tree, rays, TIE_LOWEST_ID)[:4], oracle.brute_force(tris, rays), "brute force")


@pytest.mark.parametrize("name", ["soup2k_f32", "grid2k_f32", "box12_f32", "soup1k_f64", "soup_incoherent_f32"])
def test_against_golden(emul, name):
    """LBVH tree vs the reference's tree (golden outputs): identical under the canonical tie-break;
    reference tree run through the device stack machine: identical in every mode, counters included."""
    g = golden(name)
    tris, rays = g["tris"], g["rays"]
    lbvh = emul.build(tris=tris)
    ids, t, u, v, _ = emul.trace(lbvh, rays, TIE_LOWEST_ID)
    assert_hits_equal((ids, t, u, v), tuple(g[f"lowest_{k}"] for k in ("ids", "t", "u", "v")), name)
    any_ids = emul.trace(lbvh, rays, ANY_HIT | TIE_LOWEST_ID)[0]
    assert ((any_ids != INVALID) == (g["any_ids"] != INVALID)).all()
    same = emul.from_reference(g["ref_bounds"], g["ref_index"], g["ref_prim_ids"], tris)
    for mode, flags in MODES:
        got = emul.trace(same, rays, flags)
        assert_hits_equal(got[:4], tuple(g[f"{mode}_{k}"] for k in ("ids", "t", "u", "v")), f"{name}/{mode}")
        assert (got[4] == g[f"{mode}_stats"]).all()


def test_duplicate_and_degenerate_primitives(emul, oracle):
    """All-identical Morton keys (balanced by the index tie-break), zero-area and repeated triangles."""
    base = scenes.soup(1, seed=5)
    tris = np.repeat(base, 300, axis=0)
    tris[100:200, 3:] = tris[100:200, :3].repeat(2, axis=0).reshape(100, 6)[:, :6]     # degenerate: p1 = p2 = p0
    tree = emul.build(tris=tris)
    bounds, index_values = emul.compact(tree)
    otree = oracle.from_arrays(bounds, index_values, tree["prim_ids"])
    assert oracle.check_invariants(otree, 8) == 0
    assert tree["depth"] <= 12
    rays = scenes.make_primary("soup", 32, 32)
    got = emul.trace(tree, rays, TIE_LOWEST_ID)
    assert_hits_equal(got[:4], oracle.brute_force(tris, rays), "duplicates")
    hit = got[0] != INVALID
    assert (got[0][hit] == got[0][hit].min()).all()          # lowest id among identical triangles


def test_axis_parallel_rays(emul, oracle):
    """Zero direction components give +-FLT_MAX inverse directions and inf-inf NaNs in the slab test;
    the reference swallows them through argument order (utils.h:40-43, node.h:112-115)."""
    tris = scenes.grid(800)
    tree = emul.build(tris=tris)
    bounds, index_values = emul.compact(tree)
    otree = oracle.from_arrays(bounds, index_values, tree["prim_ids"])
    oracle.set_triangles(otree, tris)
    rng = np.random.RandomState(11)
    m = 2000
    rays = np.zeros((m, 8), np.float32)
    rays[:, 0] = rng.uniform(0, 1, m); rays[:, 1] = 0.5; rays[:, 2] = rng.uniform(0, 1, m)
    rays[:, 4] = -1.0                                         # straight down: dir = (0, -1, 0)
    rays[: m // 2, 0] = np.round(rays[: m // 2, 0] * 20) / 20  # many exactly on grid lines / vertices
    rays[:, 7] = np.finfo(np.float32).max
    for _, flags in MODES:
        got = emul.trace(tree, rays, flags)
        want = oracle.trace(otree, rays, flags=flags, stats=True)
        assert_hits_equal(got[:4], want[:4], f"axis-parallel/{flags}")
        assert (got[4] == want[4]).all()
    # Rays lying exactly in box faces are where the FAST slab test is not watertight (the result then
    # depends on the tree, in the reference too); the ROBUST test is, so it must agree with brute force.
    assert_hits_equal(emul.trace(tree, rays, ROBUST | TIE_LOWEST_ID)[:4], oracle.brute_force(tris, rays), "axis-parallel brute force")


def test_nan_and_degenerate_rays(emul, oracle):
    tris = scenes.soup(2000)
    tree = emul.build(tris=tris)
    bounds, index_values = emul.compact(tree)
    otree = oracle.from_arrays(bounds, index_values, tree["prim_ids"])
    oracle.set_triangles(otree, tris)
    rays = scenes.make_primary("soup", 40, 40).copy()
    rays[0::10, 6] = np.nan
    rays[1::10, 7] = np.nan
    rays[2::10, 3:6] = 0
    rays[3::10, 6], rays[3::10, 7] = 1.0, 0.5
    rays[4::10, 7] = np.inf
    rays[6::10, 3] = 0
    for _, flags in MODES:
        got = emul.trace(tree, rays, flags)
        want = oracle.trace(otree, rays, flags=flags)
        assert (got[0] == want[0]).all()
        ok = ~np.isnan(rays[:, 7])
        assert_hits_equal(tuple(x[ok] for x in got[:4]), tuple(x[ok] for x in want), f"nan/{flags}")


def test_boxes_and_centres_input(emul, oracle):
    tris = scenes.soup(3000)
    bb, cc = oracle.tri_bboxes_centers(tris)
    a = emul.build(tris=tris)
    b = emul.build(bboxes=bb, centers=cc)
    assert (a["nodes"] == b["nodes"]).all() and (a["prim_ids"] == b["prim_ids"]).all()


def test_leaf_size_config(emul, oracle):
    tris = scenes.soup(4000)
    for min_leaf, max_leaf in ((1, 1), (1, 4), (4, 8), (1, 15)):
        tree = emul.build(tris=tris, min_leaf=min_leaf, max_leaf=max_leaf)
        bounds, index_values = emul.compact(tree)
        otree = oracle.from_arrays(bounds, index_values, tree["prim_ids"])
        assert oracle.check_invariants(otree, max_leaf) == 0
        counts = index_values & 15
        assert counts.max() <= max_leaf
        if max_leaf == 1:
            assert index_values.shape[0] == 2 * tris.shape[0] - 1


@pytest.mark.parametrize("kind,n", [("soup", 1), ("soup", 2), ("soup", 5), ("soup", 20000), ("grid", 20000), ("box12", 12)])
def test_wide_tree_is_conservative_and_equivalent(emul, oracle, kind, n):
    """The compressed 4-wide tree derived from the binary one: every child box contains the binary box it
    was quantised from, and traversing it gives the binary traversal's canonical closest hit bit for bit
    (and the same occlusion answer), in about half as many inner steps."""
    tris = scenes.make_mesh(kind, n)
    tree = emul.build(tris=tris)
    wide, levels = emul.wide_build(tree)
    assert 1 <= levels <= tree["depth"] + 1
    # decode every child box and compare with the binary node it came from (walk both trees together)
    nodes = tree["nodes"]
    words = wide.view(np.uint32)
    def decode(w, c):
        origin = w[0:3].view(np.float32)
        exps = [(int(w[3]) >> (8 * k)) & 0xFF for k in range(3)]
        cell = np.array([np.uint32(e << 23) for e in exps], np.uint32).view(np.float32)
        lo = np.array([origin[a] + np.float32((int(w[4 + a]) >> (8 * c)) & 0xFF) * cell[a] for a in range(3)], np.float32)
        hi = np.array([origin[a] + np.float32((int(w[7 + a]) >> (8 * c)) & 0xFF) * cell[a] for a in range(3)], np.float32)
        return lo, hi
    checked = 0
    rng = np.random.RandomState(0)
    for wi in rng.choice(words.shape[0], size=min(300, words.shape[0]), replace=False):
        w = words[wi]
        count = (int(w[3]) >> 24) & 0xFF
        assert 1 <= count <= 4
        for c in range(count):
            ref = int(w[10 + c])
            lo, hi = decode(w, c)
            if ref & 15:      # leaf: find the binary leaf with the same index value and check containment
                match = np.nonzero(nodes.view(np.uint32)[:, 6] == ref)[0]
                assert match.size >= 1
                b = nodes[match[0]][:6]
                assert (lo <= b[0::2]).all() and (hi >= b[1::2]).all()
                checked += 1
    assert checked > 0 or n <= 2
    rays = scenes.make_primary(kind, 97, 95)          # odd sizes: no ray lies exactly in a vertex plane
    binary = emul.trace(tree, rays, TIE_LOWEST_ID)
    got = emul.wide_trace(tree, wide, rays, 0)
    assert_hits_equal(got[:4], binary[:4], f"wide/{kind}/{n}")
    # rays lying exactly in box faces (even image width: the centre column has dir.x == 0 on a vertex plane):
    # the wide test is conservative, so it must agree with the watertight ROBUST binary traversal
    rays_d = scenes.make_primary(kind, 96, 96)
    assert_hits_equal(emul.wide_trace(tree, wide, rays_d, 0)[:4], emul.trace(tree, rays_d, ROBUST | TIE_LOWEST_ID)[:4], "wide degenerate")
    occl = emul.wide_trace(tree, wide, rays, 1)
    assert ((occl[0] != INVALID) == (binary[0] != INVALID)).all()
    if n >= 1000:
        assert got[4].mean() < 0.75 * binary[4][:, 0].mean()      # fewer node fetches per ray
    rays2 = scenes.incoherent_rays(tris, 3000) if n >= 5 else rays
    assert_hits_equal(emul.wide_trace(tree, wide, rays2, 0)[:4], emul.trace(tree, rays2, TIE_LOWEST_ID)[:4], "wide incoherent")


@pytest.mark.parametrize("kind,n,dtype", [("soup", 5000, np.float32), ("grid", 4232, np.float32), ("soup", 1777, np.float64)])
def test_block_local_phase_builds_the_same_tree(emul, kind, n, dtype):
    """The hierarchy kernel merges inside a block of consecutive leaves through shared memory first and only
    then through global memory; every schedule must give the same nodes, bit for bit."""
    tris = (scenes.soup(n, seed=5) if kind == "soup" else scenes.grid(46)).astype(dtype)
    try:
        emul.set_block(0, 0)
        ref = emul.build(tris=tris)
        for leaves in (2, 3, 32, 256, 1 << 20):
            for order in (0, 1, 2):
                emul.set_block(leaves, order)
                got = emul.build(tris=tris)
                assert got["depth"] == ref["depth"]
                assert np.array_equal(got["nodes"].view(np.uint8), ref["nodes"].view(np.uint8)), (leaves, order)
                assert np.array_equal(got["prim_ids"], ref["prim_ids"])
    finally:
        emul.set_block(0, 0)


@pytest.mark.parametrize("kind,n,dtype", [("soup", 20000, np.float32), ("grid", 20000, np.float32), ("soup", 6000, np.float64),
                                           ("soup", 200, np.float32), ("soup", 3, np.float32), ("soup", 2, np.float32), ("soup", 1, np.float32)])
def test_sah_treelet_pass_keeps_results_and_saves_steps(emul, oracle, kind, n, dtype):
    """The experimental second build pass (treelet_warp.cuh: SAH rebuild of every maximal LBVH subtree of at most
    256 / 128 primitives) must leave a valid reference-layout BVH that answers every ray like the plain LBVH
    (canonical tie-break), with the recorded depth still bounding the traversal stack, and with fewer steps."""
    tris = (scenes.soup(n, seed=7) if kind == "soup" else scenes.make_mesh(kind, n)).astype(dtype)
    rays = scenes.make_primary(kind, 96, 96, dtype=dtype)
    plain = emul.build(tris=tris)
    try:
        emul.set_treelets(True)
        tree = emul.build(tris=tris)
    finally:
        emul.set_treelets(False)
    assert np.array_equal(np.sort(tree["prim_ids"]), np.arange(tris.shape[0]))
    bounds, index_values = emul.compact(tree)
    otree = oracle.from_arrays(bounds, index_values, tree["prim_ids"])
    assert oracle.check_invariants(otree, 8) == 0
    before = otree.arrays()[0]
    oracle.refit(otree)                                   # every inner box is exactly the union of its children
    assert (otree.arrays()[0] == before).all()
    # depth bound: walk the compact tree
    first, count = (index_values >> 4).astype(np.int64), (index_values & 15)
    depth, stack = 0, [(0, 0)]
    while stack:
        i, d = stack.pop()
        if count[i] == 0:
            depth = max(depth, d + 1)
            stack.append((first[i], d + 1)); stack.append((first[i] + 1, d + 1))
    assert depth <= tree["depth"]
    for flags in (TIE_LOWEST_ID, TIE_LOWEST_ID | ANY_HIT):
        a, b = emul.trace(plain, rays, flags), emul.trace(tree, rays, flags)
        if flags & ANY_HIT:
            assert np.array_equal(a[0] != INVALID, b[0] != INVALID)
        else:
            assert_hits_equal(b[:4], a[:4], f"treelets {kind}-{n}")
    if tris.shape[0] >= 6000:
        steps_plain, steps_tree = emul.trace(plain, rays, TIE_LOWEST_ID)[4][:, 0].mean(), emul.trace(tree, rays, TIE_LOWEST_ID)[4][:, 0].mean()
        assert steps_tree < 0.95 * steps_plain, (steps_plain, steps_tree)


def test_sah_treelet_pass_on_boxes_and_duplicates(emul, oracle):
    """Boxes + centres input (the reference's own build entry point) and heavily duplicated primitives."""
    tris = scenes.soup(3000, seed=1)
    tris = np.concatenate([tris, tris[:500], tris[:500]])                 # 1000 exact duplicates
    bb, cc = oracle.tri_bboxes_centers(tris)
    try:
        emul.set_treelets(True)
        tree = emul.build(bboxes=bb, centers=cc)
    finally:
        emul.set_treelets(False)
    assert np.array_equal(np.sort(tree["prim_ids"]), np.arange(tris.shape[0]))
    bounds, index_values = emul.compact(tree)
    otree = oracle.from_arrays(bounds, index_values, tree["prim_ids"])
    assert oracle.check_invariants(otree, 8) == 0


@pytest.mark.parametrize("kind,n,dtype", [("soup", 30000, np.float32), ("grid", 30000, np.float32), ("soup", 5000, np.float64)])
def test_sah_treelet_phases_have_no_order_dependence(emul, kind, n, dtype):
    """On the device the iterations of a phase are the threads of a block; running them in descending instead of
    ascending order on the host must give the same tree (compacted: the order in which splits draw their node
    pairs is free), or a phase reads something another iteration of the same phase writes."""
    tris = (scenes.soup(n, seed=11) if kind == "soup" else scenes.make_mesh(kind, n)).astype(dtype)
    try:
        emul.set_treelets(True)
        a = emul.build(tris=tris)
        emul.set_treelets(True, reversed_phases=True)
        b = emul.build(tris=tris)
    finally:
        emul.set_treelets(False)
    (ab, ai), (bb, bi) = emul.compact(a), emul.compact(b)
    assert ab.shape == bb.shape and (ab == bb).all() and (ai == bi).all()
    assert (a["prim_ids"] == b["prim_ids"]).all() and a["depth"] == b["depth"]


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_emulation_against_the_oracle(emul, oracle, seed):
    """Random small scenes (clustered, duplicated, flat and degenerate triangles; random, axis-parallel and
    zero-direction rays; both scalar types; both Morton widths; with and without the treelet pass): the device
    code run on the host must agree with the reference algorithm on the same tree, bit for bit, counters included."""
    rng = np.random.RandomState(1000 + seed)
    dtype = np.float64 if seed % 4 == 3 else np.float32
    n = int(rng.choice([1, 2, 3, 7, 33, 200, 900, 2500]))
    centres = rng.rand(max(1, n // 20), 3) * rng.choice([1.0, 100.0, 1e-3])
    base = centres[rng.randint(0, centres.shape[0], n)]
    tris = (base[:, None, :] + (rng.rand(n, 3, 3) - 0.5) * rng.choice([0.02, 0.5])).reshape(n, 9)
    if n >= 7:
        tris[rng.randint(0, n, n // 7)] = tris[rng.randint(0, n, n // 7)]         # exact duplicates
        flat = rng.randint(0, n, n // 7)
        tris[flat, 2::3] = tris[flat, 2:3]                                        # axis-aligned (flat in z)
        deg = rng.randint(0, n, max(1, n // 20))
        tris[deg, 3:6] = tris[deg, 0:3]                                           # zero-area
    tris = tris.astype(dtype)
    m = 300
    org = (rng.rand(m, 3) * 1.6 - 0.3) * float(np.abs(tris).max() + 1e-3)
    target = tris.reshape(n, 3, 3).mean(axis=1)[rng.randint(0, n, m)]
    d = target - org + (rng.rand(m, 3) - 0.5) * 0.05
    d[:20, 0] = 0; d[20:30, :2] = 0; d[30:33] = 0                                 # axis-parallel and zero directions
    rays = np.concatenate([org, d, np.zeros((m, 1)), np.full((m, 1), rng.choice([1.0, 2.0, np.finfo(np.float32).max]))], axis=1).astype(dtype)
    try:
        emul.set_treelets(seed % 2 == 1)
        tree = emul.build(tris=tris, morton_bits=63 if seed % 3 == 2 else 30,
                          min_leaf=int(rng.choice([1, 2])), max_leaf=int(rng.choice([1, 4, 8, 15])))
    finally:
        emul.set_treelets(False)
    bounds, index_values = emul.compact(tree)
    otree = oracle.from_arrays(bounds, index_values, tree["prim_ids"])
    assert oracle.check_invariants(otree, 15) == 0
    oracle.set_triangles(otree, tris)
    for _, flags in MODES:
        got = emul.trace(tree, rays, flags)
        want = oracle.trace(otree, rays, flags=flags, stats=True)
        assert_hits_equal(got[:4], want[:4], f"seed {seed} flags {flags}")
        assert (got[4] == want[4]).all()
    robust = [f for _, f in MODES if f & ROBUST and not f & ANY_HIT and f & TIE_LOWEST_ID]
    if robust:
        assert_hits_equal(emul.trace(tree, rays, robust[0])[:4], oracle.brute_force(tris, rays, flags=robust[0] & ~ROBUST), f"seed {seed} brute force")


def test_sah_treelet_pass_survives_nan_and_inf_vertices(emul):
    """NaN / infinite vertices (undefined in the reference) must not break the pass: the result is still a
    permutation of the primitives, and rays get the same answers as from the plain LBVH."""
    tris = scenes.soup(1500, seed=21)
    tris[7, 4] = np.nan; tris[400, 0] = np.inf; tris[401, 8] = -np.inf; tris[900, :] = np.nan
    rays = scenes.make_primary("soup", 64, 64)
    plain = emul.build(tris=tris)
    try:
        emul.set_treelets(True)
        tree = emul.build(tris=tris)
    finally:
        emul.set_treelets(False)
    assert np.array_equal(np.sort(tree["prim_ids"]), np.arange(tris.shape[0]))
    a, b = emul.trace(plain, rays, TIE_LOWEST_ID), emul.trace(tree, rays, TIE_LOWEST_ID)
    assert_hits_equal(b[:4], a[:4], "treelets with NaN / inf vertices")


def test_wide_tree_with_degenerate_rays(emul):
    """Zero direction components turn slab values of the wide step into NaNs that are ignored; an UNUSED child slot
    (inverted box) must still never be entered (it was, before the slot count was checked: a wild leaf reference).
    NaN intervals, empty and inverted intervals, infinite tmax: the wide path reports the binary path's hits."""
    tris = scenes.soup(3000)
    rays = scenes.make_primary("soup", 40, 40).copy()
    nan, inf = np.float32(np.nan), np.float32(np.inf)
    rays[0::10, 6] = nan
    rays[1::10, 7] = nan
    rays[2::10, 3:6] = 0
    rays[3::10, 6], rays[3::10, 7] = 1.0, 0.5
    rays[4::10, 7] = inf
    rays[5::10, 6] = rays[5::10, 7] = 0.75
    rays[6::10, 3] = 0
    rays[7::10, 3:5] = 0
    for quality in ("low", "high"):
        tree = emul.build(tris=tris, quality=quality)
        wide, _ = emul.wide_build(tree)
        binary = emul.trace(tree, rays, 4)
        got = emul.wide_trace(tree, wide, rays, 0)
        assert (got[0] == binary[0]).all()
        hit = binary[0] != INVALID
        assert (got[1][hit].view(np.uint32) == binary[1][hit].view(np.uint32)).all()
        any_got = emul.wide_trace(tree, wide, rays, 1)
        assert ((any_got[0] != INVALID) == (emul.trace(tree, rays, 1 | 4)[0] != INVALID)).all()
