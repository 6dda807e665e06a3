// tests/host_emul.cpp — CPU emulation of the device code, FOR TESTS ONLY.
//
// Compiles the very headers the CUDA kernels are made of (bvh_b200/csrc/core.cuh, build_core.cuh,
// traverse_core.cuh) with g++ -ffp-contract=off and runs the per-thread functions sequentially, so
// that the logic of the LBVH bottom-up pass and of the traversal stack machine can be checked against
// the oracle on a machine without a GPU.  It is not part of the product and is never loaded by it;
// the GPU tests (pytest -m gpu) exercise the real kernels through the C ABI.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <memory>
#include <vector>

#include "build_core.cuh"
#include "treelet_warp.cuh"
#include "traverse_core.cuh"
#include "wide_bvh.cuh"

using namespace bvhb200;

namespace {

template <typename U> struct HostStack {
    U data[256];
    uint32_t sp = 0;
    void push(U v) { data[sp++] = v; }
    U pop() { return data[--sp]; }
    bool empty() const { return sp == 0; }
    bool try_pop(U& out) { if (sp == 0) return false; out = data[--sp]; return true; }
};

static int g_block_leaves = 0, g_block_order = 0, g_treelets = 0, g_last_treelets = 0;
static size_t g_last_slots = 0;

template <typename T, typename K>
uint32_t emul_build(const T* verts, const T* bboxes, const T* centers, uint32_t n, uint32_t min_leaf, uint32_t max_leaf,
                    DevNode<T>* nodes /* 2n slots */, uint32_t* prim_ids, DevTri<T>* tris, uint32_t* depth_out) {
    using R = Real<T>;
    // K1: centre bounds
    std::vector<T> cs(3 * (size_t)n);
    T mn[3] = { R::max(), R::max(), R::max() }, mx[3] = { R::neg(R::max()), R::neg(R::max()), R::neg(R::max()) };
    for (uint32_t i = 0; i < n; ++i) {
        T c[3];
        if (verts) { T bmin[3], bmax[3]; tri_bounds_center(verts + 9 * (size_t)i, bmin, bmax, c); }
        else for (int k = 0; k < 3; ++k) c[k] = centers[3 * (size_t)i + k];
        for (int k = 0; k < 3; ++k) { cs[3 * (size_t)i + k] = c[k]; mn[k] = robust_min(mn[k], c[k]); mx[k] = robust_max(mx[k], c[k]); }
    }
    // K2: keys
    const GridXform<T> g = make_grid_xform(mn, mx, MortonTraits<K>::bits_per_axis);
    std::vector<K> keys(n);
    for (uint32_t i = 0; i < n; ++i) keys[i] = morton_key<T, K>(&cs[3 * (size_t)i], g);
    // K3: stable sort by key
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&] (uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<K> sorted(n);
    for (uint32_t i = 0; i < n; ++i) { sorted[i] = keys[order[i]]; prim_ids[i] = order[i]; }
    // K4: bottom-up, leaves in order (any order is valid: the second arrival continues)
    std::vector<int> flags(n, -1);
    uint32_t info[4] = { 0, 0, 0, 0 };
    std::memset(nodes, 0, 2 * (size_t)n * sizeof(DevNode<T>));
    BuildParams<T> p;
    p.nodes = nodes; p.flags = flags.data(); p.info = info; p.n = n;
    p.min_leaf = min_leaf; p.max_leaf = max_leaf;
    std::vector<uint32_t> alive(n, 1u);                      // morton_kernel initialises the liveness of every pair
    p.alive = alive.data();
    std::vector<Treelet> list((size_t)n / 3 + 1);
    uint32_t list_count = 0;
    if (g_treelets && n > 2) { p.treelets = list.data(); p.treelet_count = &list_count; p.treelet_max = (uint32_t)TreeletCfg<T>::kMaxPrims; }
    auto leaf_box = [&] (uint32_t i, T bmin[3], T bmax[3]) {
        const uint32_t id = order[i];
        if (verts) {
            T c[3];
            tri_bounds_center(verts + 9 * (size_t)id, bmin, bmax, c);
            if (tris) tris[i] = precompute_tri(verts + 9 * (size_t)id);
        } else {
            for (int k = 0; k < 3; ++k) { bmin[k] = bboxes[6 * (size_t)id + k]; bmax[k] = bboxes[6 * (size_t)id + 3 + k]; }
        }
    };
    if (g_block_leaves <= 0 || n == 1) {
        for (uint32_t i = 0; i < n; ++i) {
            T bmin[3], bmax[3];
            leaf_box(i, bmin, bmax);
            build_bottom_up<T, K, HostSync>(p, sorted.data(), i, bmin, bmax);
        }
    } else {
        // the device kernel's two phases (block-local merges through "shared" slots, then global), one block at a
        // time; within a block the leaves run in the order g_block_order selects (0 ascending, 1 descending,
        // 2 interleaved), which changes who arrives first at every boundary
        const uint32_t B = (uint32_t)g_block_leaves;
        std::vector<DevNode<T>> lnodes(2 * (size_t)B);
        std::vector<int> lflags(B), linfo(B);
        struct Pending { ClimbState<T> s; uint32_t parent, side; int outcome; };
        std::vector<Pending> pend(B);
        for (uint32_t i0 = 0; i0 < n; i0 += B) {
            const uint32_t iend = i0 + B < n ? i0 + B : n;
            std::fill(lflags.begin(), lflags.end(), -1);
            std::fill(linfo.begin(), linfo.end(), 0);
            LocalSlots<T> loc { lnodes.data(), lflags.data(), linfo.data() };
            const uint32_t m = iend - i0;
            for (uint32_t k = 0; k < m; ++k) {
                uint32_t j = k;
                if (g_block_order == 1) j = m - 1 - k;
                else if (g_block_order == 2) j = (k & 1) ? m - 1 - k / 2 : k / 2;
                T bmin[3], bmax[3];
                leaf_box(i0 + j, bmin, bmax);
                Pending& q = pend[j];
                climb_init(q.s, i0 + j, bmin, bmax);
                do q.outcome = local_step<T, K, HostSync>(p, sorted.data(), q.s, loc, i0, iend, q.parent, q.side); while (q.outcome == kStepCarry);
            }
            for (uint32_t j = 0; j < m; ++j) {                 // after the block barrier
                Pending& q = pend[j];
                if (q.outcome == kStepRetired && linfo[q.parent - i0] != 0) continue;     // matched, or the root
                climb_global<T, K, HostSync>(p, sorted.data(), q.s, q.parent, q.side, published_record(q.s));
            }
        }
    }
    if (g_treelets && n > 2) {
        // the second pass (treelet_warp.cuh): rebuild every maximal subtree of 3..kMaxPrims primitives, as listed by
        // the bottom-up pass above
        list.resize(list_count);
        auto shared = std::make_unique<TreeletShared<T>>();
        const T* leaf_src = verts ? verts : bboxes;
        for (const Treelet& t : list) {
            if (g_treelets == 2)          // lanes in descending order inside every step (shared-memory hazard check)
                treelet_rebuild<T, HostLanesReversed>(*shared, t, nodes, prim_ids, verts ? tris : nullptr, leaf_src, centers, verts ? 0 : 1,
                                                      min_leaf, max_leaf, info, info[0], alive.data());
            else
                treelet_rebuild<T, HostLanes>(*shared, t, nodes, prim_ids, verts ? tris : nullptr, leaf_src, centers, verts ? 0 : 1,
                                              min_leaf, max_leaf, info, info[0], alive.data());
        }
        g_last_treelets = (int)list.size();
    }
    // compaction (lbvh_build.cu compact_*_kernel): the live pairs, in order of split position
    {
        const uint32_t pairs = n - 1;
        std::vector<uint32_t> rank(n, 0u);
        uint32_t live = 0;
        for (uint32_t q = 0; q < pairs; ++q) { rank[q] = live; live += alive[q]; }
        std::vector<DevNode<T>> dense(2 * (size_t)live + 2);
        std::memset(dense.data(), 0, sizeof(DevNode<T>));
        dense[1] = compact_remap(nodes[1], rank.data());
        for (uint32_t q = 0; q < pairs; ++q) {
            if (!alive[q]) continue;
            dense[child_slot(rank[q], 0)] = compact_remap(nodes[child_slot(q, 0)], rank.data());
            dense[child_slot(rank[q], 1)] = compact_remap(nodes[child_slot(q, 1)], rank.data());
        }
        std::memset(nodes, 0, 2 * (size_t)n * sizeof(DevNode<T>));
        std::memcpy(nodes, dense.data(), dense.size() * sizeof(DevNode<T>));
        g_last_slots = dense.size();
    }
    *depth_out = info[0] + info[2];
    return info[1];
}

// reference layout out of the dense device array (slot = reference index + 1): what c_api.cu download_mirror copies
template <typename T>
size_t emul_compact(const DevNode<T>* dev, size_t slots, T* bounds, uint64_t* index_values, size_t cap) {
    const size_t count = slots - 1;
    if (count <= cap) for (size_t i = 0; i < count; ++i) {
        std::memcpy(bounds + 6 * i, dev[i + 1].bounds, 6 * sizeof(T));
        index_values[i] = dev[i + 1].index;
    }
    return count;
}

template <typename T>
void emul_trace(const DevNode<T>* nodes, const DevTri<T>* tris, const uint32_t* prim_ids, const T* rays, size_t m,
                unsigned flags, uint32_t* ids, T* ts, T* us, T* vs, uint32_t* stats) {
    using U = typename Real<T>::UInt;
    const bool any = flags & 1u, robust = flags & 2u, lowest = flags & 4u;
    for (size_t i = 0; i < m; ++i) {
        RayCtx<T> r;
        for (int k = 0; k < 3; ++k) { r.org[k] = rays[8 * i + k]; r.dir[k] = rays[8 * i + 3 + k]; }
        r.tmin = rays[8 * i + 6]; r.tmax = rays[8 * i + 7];
        const T tmax_in = r.tmax;
        HitState<T> hit { kInvalidId, r.tmax, (T)0, (T)0 };
        HostStack<U> stack;
        uint32_t st[3] = { 0, 0, 0 };
        const U root = nodes[1].index;
        if (robust) {
            ray_prologue<T, true>(r);
            if (any) traverse_ray<T, true, true>(nodes, tris, prim_ids, lowest, root, r, hit, stack, st);
            else     traverse_ray<T, false, true>(nodes, tris, prim_ids, lowest, root, r, hit, stack, st);
        } else {
            ray_prologue<T, false>(r);
            if (any) traverse_ray<T, true, false>(nodes, tris, prim_ids, lowest, root, r, hit, stack, st);
            else     traverse_ray<T, false, false>(nodes, tris, prim_ids, lowest, root, r, hit, stack, st);
        }
        const bool was_hit = hit.slot != kInvalidId;
        ids[i] = was_hit ? prim_ids[hit.slot] : kInvalidId;
        ts[i] = was_hit ? hit.t : tmax_in; us[i] = was_hit ? hit.u : (T)0; vs[i] = was_hit ? hit.v : (T)0;
        if (stats) { stats[3 * i] = st[0]; stats[3 * i + 1] = st[1]; stats[3 * i + 2] = st[2]; }
    }
}

// upload path: dense reference arrays -> device layout (slot = index + 1)
template <typename T>
void emul_from_reference(const T* bounds, const uint64_t* index_values, size_t node_count, DevNode<T>* dev) {
    using U = typename Real<T>::UInt;
    std::memset(dev, 0, sizeof(DevNode<T>));
    for (size_t i = 0; i < node_count; ++i) {
        std::memcpy(dev[i + 1].bounds, bounds + 6 * i, 6 * sizeof(T));
        dev[i + 1].index = (U)index_values[i];
        dev[i + 1].pad = 0;
    }
}

// wide tree: the same collapse the device runs level by level (lbvh_build.cu wide_collapse_kernel), sequentially
size_t emul_wide_build_impl(const DevNode<float>* nodes, WideNode* wide, size_t cap, uint32_t* levels_out) {
    std::vector<std::pair<uint32_t, uint32_t>> frontier { { 1u, 0u } }, next;
    uint32_t count = 1, levels = 0;
    while (!frontier.empty()) {
        ++levels;
        next.clear();
        for (auto [slot_self, wide_index] : frontier) {
            uint32_t slot[4];
            const int used = wide_gather_children(nodes, slot_self, slot);
            WideNode w;
            bool is_inner[4];
            wide_encode(nodes, slot, used, w, is_inner);
            for (int c = 0; c < used; ++c) {
                if (!is_inner[c]) continue;
                const uint32_t idx = count++;
                w.child[c] = idx << kPrimCountBits;
                next.emplace_back(slot[c], idx);
            }
            if (wide_index < cap) wide[wide_index] = w;
        }
        frontier.swap(next);
    }
    *levels_out = levels;
    return count;
}

void emul_wide_trace_impl(const WideNode* wide, const DevTri<float>* tris, const uint32_t* prim_ids, const float* rays, size_t m,
                          unsigned flags, uint32_t* ids, float* ts, float* us, float* vs, uint32_t* steps) {
    const bool any = flags & 1u;
    for (size_t i = 0; i < m; ++i) {
        RayCtx<float> r;
        for (int k = 0; k < 3; ++k) { r.org[k] = rays[8 * i + k]; r.dir[k] = rays[8 * i + 3 + k]; }
        r.tmin = rays[8 * i + 6]; r.tmax = rays[8 * i + 7];
        const float tmax_in = r.tmax;
        HitState<float> hit { kInvalidId, r.tmax, 0.f, 0.f };
        uint32_t n_steps = 0;
        if (!ray_interval_is_nan(r)) {
            wide_ray_setup(r);
            HostStack<uint32_t> stack;
            uint32_t top = 0;
            for (;;) {
                bool alive = true;
                while (index_count(top) == 0) {
                    uint32_t w[16];
                    std::memcpy(w, wide + (top >> kPrimCountBits), 64);
                    ++n_steps;
                    const bool ok = any ? wide_step<true>(w, r, top, stack) : wide_step<false>(w, r, top, stack);
                    if (!ok) { alive = false; break; }
                }
                if (!alive) break;
                leaf_step<float>(tris, prim_ids, true, top, r, hit, nullptr);
                if (any && hit.slot != kInvalidId) break;
                if (stack.empty()) break;
                top = stack.pop();
            }
        }
        const bool was_hit = hit.slot != kInvalidId;
        ids[i] = was_hit ? prim_ids[hit.slot] : kInvalidId;
        ts[i] = was_hit ? hit.t : tmax_in; us[i] = was_hit ? hit.u : 0.f; vs[i] = was_hit ? hit.v : 0.f;
        if (steps) steps[i] = n_steps;
    }
}

} // namespace

#define EMUL_API(T, S) \
    uint32_t emul_build##S(const T* verts, const T* bboxes, const T* centers, uint32_t n, uint32_t min_leaf, uint32_t max_leaf, \
                           int morton_bits, void* nodes, uint32_t* prim_ids, void* tris, uint32_t* depth) { \
        if (morton_bits <= 30) return emul_build<T, uint32_t>(verts, bboxes, centers, n, min_leaf, max_leaf, (DevNode<T>*)nodes, prim_ids, (DevTri<T>*)tris, depth); \
        return emul_build<T, uint64_t>(verts, bboxes, centers, n, min_leaf, max_leaf, (DevNode<T>*)nodes, prim_ids, (DevTri<T>*)tris, depth); } \
    size_t emul_compact##S(const void* dev, size_t slots, T* bounds, uint64_t* index_values, size_t cap) { \
        return emul_compact<T>((const DevNode<T>*)dev, slots, bounds, index_values, cap); } \
    void emul_trace##S(const void* nodes, const void* tris, const uint32_t* prim_ids, const T* rays, size_t m, unsigned flags, \
                       uint32_t* ids, T* ts, T* us, T* vs, uint32_t* stats) { \
        emul_trace<T>((const DevNode<T>*)nodes, (const DevTri<T>*)tris, prim_ids, rays, m, flags, ids, ts, us, vs, stats); } \
    void emul_from_reference##S(const T* bounds, const uint64_t* index_values, size_t node_count, void* dev) { \
        emul_from_reference<T>(bounds, index_values, node_count, (DevNode<T>*)dev); } \
    void emul_precompute##S(const T* verts, const uint32_t* prim_ids, size_t n, void* tris) { \
        for (size_t i = 0; i < n; ++i) ((DevTri<T>*)tris)[i] = precompute_tri(verts + 9 * (size_t)prim_ids[i]); }

extern "C" {
size_t emul_wide_build(const void* nodes, void* wide, size_t cap, uint32_t* levels) {
    return emul_wide_build_impl((const DevNode<float>*)nodes, (WideNode*)wide, cap, levels);
}
void emul_wide_trace(const void* wide, const void* tris, const uint32_t* prim_ids, const float* rays, size_t m, unsigned flags,
                     uint32_t* ids, float* ts, float* us, float* vs, uint32_t* steps) {
    emul_wide_trace_impl((const WideNode*)wide, (const DevTri<float>*)tris, prim_ids, rays, m, flags, ids, ts, us, vs, steps);
}
EMUL_API(float, 3f)
EMUL_API(double, 3d)
void emul_set_treelets(int on) { g_treelets = on; }
int emul_last_treelet_count() { return g_last_treelets; }
size_t emul_last_node_slots() { return g_last_slots; }
void emul_set_block(int leaves, int order) { g_block_leaves = leaves; g_block_order = order; }
uint32_t emul_morton30(uint32_t x, uint32_t y, uint32_t z) { return MortonTraits<uint32_t>::encode(x, y, z); }
uint64_t emul_morton63(uint64_t x, uint64_t y, uint64_t z) { return MortonTraits<uint64_t>::encode(x, y, z); }
size_t emul_sizeof_node(int is_double) { return is_double ? sizeof(DevNode<double>) : sizeof(DevNode<float>); }
size_t emul_sizeof_tri(int is_double) { return is_double ? sizeof(DevTri<double>) : sizeof(DevTri<float>); }
}
