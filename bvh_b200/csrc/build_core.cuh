// bvh_b200/csrc/build_core.cuh — LBVH construction logic shared by the CUDA kernels and the host
// emulation used by the CPU tests.
//
// What it replaces: DefaultBuilder<Node>::build and the builders behind it (reference
// default_builder.h:33-62, mini_tree_builder.h, binned_sah_builder.h, sweep_sah_builder.h).  The
// output keeps the reference's data structure invariants: root at index 0,
// sibling pairs adjacent with the left child at an odd index (bvh.h:34-54), packed leaf/inner index
// (index.h:51-53), prim_ids a permutation, parent boxes enclosing child boxes computed with the
// reference's own extend() order (bbox.h:23-27, bvh.h:213-217), larger-area child on the left
// (SATO, top_down_sah_builder.h:101-108) and leaves of at most max_leaf_size primitives decided by
// the reference's SAH leaf-cost model (split_heuristic.h:30-38).
//
// Pipeline (lbvh_build.cu):  centre bounds -> Morton keys (quantisation as mini_tree_builder.h:170-183,
// bit interleave as utils.h:103-120) -> radix sort -> ONE bottom-up pass (this file) that links the
// hierarchy, unions the boxes, applies the SAH leaf collapse and writes every node directly into its
// final reference-layout slot.  Node numbering: internal node p is the split between sorted
// primitives p and p+1 and its two children live at reference indices 2p+1 and 2p+2; the root's own
// record is index 0.
#pragma once

#include "core.cuh"

namespace bvhb200 {

// ---- Morton codes ---------------------------------------------------------------------------
// utils.h:103-120 (split_bits / morton_encode) specialised to 10 bits per axis in 32 bits and
// 21 bits per axis in 64 bits.
BVH_HD uint32_t spread_bits_10(uint32_t x) {
    x &= 0x000003FFu;
    x = (x | (x << 16)) & 0xFF0000FFu;
    x = (x | (x << 8))  & 0x0F00F00Fu;
    x = (x | (x << 4))  & 0xC30C30C3u;
    x = (x | (x << 2))  & 0x49249249u;
    return x;
}
BVH_HD uint64_t spread_bits_21(uint64_t x) {
    x &= 0x1FFFFFull;
    x = (x | (x << 32)) & 0x001F00000000FFFFull;
    x = (x | (x << 16)) & 0x001F0000FF0000FFull;
    x = (x | (x << 8))  & 0x100F00F00F00F00Full;
    x = (x | (x << 4))  & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2))  & 0x1249249249249249ull;
    return x;
}

template <typename K> struct MortonTraits;
template <> struct MortonTraits<uint32_t> {
    static constexpr int bits_per_axis = 10;
    static BVH_HD uint32_t encode(uint32_t x, uint32_t y, uint32_t z) {
        return spread_bits_10(x) | (spread_bits_10(y) << 1) | (spread_bits_10(z) << 2);
    }
};
template <> struct MortonTraits<uint64_t> {
    static constexpr int bits_per_axis = 21;
    static BVH_HD uint64_t encode(uint64_t x, uint64_t y, uint64_t z) {
        return spread_bits_21(x) | (spread_bits_21(y) << 1) | (spread_bits_21(z) << 2);
    }
};

// Grid placement of a centre, as mini_tree_builder.h:170-183 does it: scale = dim * safe_inverse(diag),
// offset = -min * scale, p = max(fma(c, scale, offset), 0), cell = min(dim - 1, (size_t)p).
template <typename T> struct GridXform { T scale[3], offset[3]; };

template <typename T> BVH_HD GridXform<T> make_grid_xform(const T cmin[3], const T cmax[3], int bits_per_axis,
                                                         bool cubic_cells = true) {
    using R = Real<T>;
    GridXform<T> g;
    const T dim = (T)((uint64_t)1 << bits_per_axis);
    if (cubic_cells) {
        // One cell size for all three axes (the largest extent spans the grid).  With per-axis scaling a
        // flat mesh wastes every third key bit on splits across its thin dimension; cubic cells keep the
        // Morton splits spatially isotropic, which is what the SAH builders of the reference achieve by
        // weighing areas (measured: -30 % inner steps on the height-field mesh, neutral on the soup).
        T extent = R::sub(cmax[0], cmin[0]);
        for (int a = 1; a < 3; ++a) extent = robust_max(R::sub(cmax[a], cmin[a]), extent);
        const T scale = R::mul(dim, safe_inverse(extent));
        for (int a = 0; a < 3; ++a) { g.scale[a] = scale; g.offset[a] = R::mul(R::neg(cmin[a]), scale); }
        return g;
    }
    for (int a = 0; a < 3; ++a) {
        g.scale[a] = R::mul(dim, safe_inverse(R::sub(cmax[a], cmin[a])));
        g.offset[a] = R::mul(R::neg(cmin[a]), g.scale[a]);
    }
    return g;
}

template <typename T, typename K> BVH_HD K morton_key(const T c[3], const GridXform<T>& g) {
    using R = Real<T>;
    const uint64_t last = ((uint64_t)1 << MortonTraits<K>::bits_per_axis) - 1;
    uint64_t q[3];
    for (int a = 0; a < 3; ++a) {
        T p = robust_max(R::fma(c[a], g.scale[a], g.offset[a]), (T)0);
        // saturating conversion; p is never negative or NaN here
        uint64_t cell = p >= (T)last ? last : (uint64_t)p;
        q[a] = cell;
    }
    return MortonTraits<K>::encode((K)q[0], (K)q[1], (K)q[2]);
}

// ---- Hierarchy ------------------------------------------------------------------------------
// "Distance" between sorted neighbours k and k+1: XOR of the keys, ties broken by XOR of the
// positions (a ruler sequence, so runs of identical keys become balanced subtrees).  Smaller means
// more similar; the bottom-up pass merges across the smaller boundary first, so the boundary with
// the LARGEST distance ends up at the root.
template <typename K> struct Delta { K hi; uint32_t lo; };
template <typename K> BVH_HD bool delta_less(const Delta<K>& a, const Delta<K>& b) {
    return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}
template <typename K> BVH_HD Delta<K> delta_at(const K* __restrict__ keys, uint32_t k) {
    return Delta<K>{ (K)(keys[k] ^ keys[k + 1]), k ^ (k + 1) };
}

// While the tree is being built, the spare word of a node record carries what the SECOND arrival at the
// parent needs from the first: the subtree's SAH cost and its depth.  Keeping them in the record itself
// (instead of a side array) means one 32-byte store and one 32-byte load per node on the critical path.
// float nodes have a 32-bit spare word: the cost's low 7 mantissa bits are replaced by the depth (<= 127;
// the cost only feeds the leaf-collapse comparison); double nodes have 64 bits: float cost + depth.
// The traversal never reads this word and the host mirror drops it.
template <typename T> struct AuxPack;
template <> struct AuxPack<float> {
    static BVH_HD uint32_t pack(float cost, uint32_t depth) {
        return (Real<float>::bits(cost) & ~0x7Fu) | (depth > 127u ? 127u : depth);
    }
    static BVH_HD float cost(uint32_t w) { return Real<float>::from_bits(w & ~0x7Fu); }
    static BVH_HD uint32_t depth(uint32_t w) { return w & 0x7Fu; }
};
template <> struct AuxPack<double> {
    static BVH_HD uint64_t pack(double cost, uint32_t depth) {
        return ((uint64_t)Real<float>::bits((float)cost) << 32) | depth;
    }
    static BVH_HD double cost(uint64_t w) { return (double)Real<float>::from_bits((uint32_t)(w >> 32)); }
    static BVH_HD uint32_t depth(uint64_t w) { return (uint32_t)w; }
};

// A bottom subtree handed to the SAH treelet pass (treelet_warp.cuh): device slot of its root record and the
// range [l, r] of sorted primitives it covers.
struct Treelet { uint32_t slot, l, r; };

template <typename T> struct BuildParams {
    DevNode<T>* nodes;            // 2n device slots (slot = reference index + 1)
    int* flags;                   // n-1, initialised to -1
    uint32_t* info;               // [0] tree depth (max stack entries needed), [1] root split position
    uint32_t n;
    uint32_t min_leaf, max_leaf;  // TopDownSahBuilder::Config, top_down_sah_builder.h:27-40
    // Treelet discovery (quality Medium / High): every MAXIMAL subtree of 3..treelet_max primitives is appended
    // to treelets[] by the thread that merges it into a larger parent.  treelet_max == 0 switches it off.
    Treelet* treelets = nullptr;
    uint32_t* treelet_count = nullptr;
    uint32_t treelet_max = 0;
    // Liveness of the sibling pairs (pair p = the two children of split position p, reference indices 2p+1 and
    // 2p+2): initialised to 1 for every p; a merge that collapses its subtree into a leaf clears the pairs the
    // subtree owned.  The compaction pass (lbvh_build.cu compact_tree) keeps exactly the live pairs.  nullptr: off.
    uint32_t* alive = nullptr;
};

template <typename T> BVH_HD void append_treelet(const BuildParams<T>& p, uint32_t slot, uint32_t l, uint32_t r) {
#if defined(__CUDA_ARCH__)
    const uint32_t at = atomicAdd(p.treelet_count, 1u);
#else
    const uint32_t at = (*p.treelet_count)++;
#endif
    p.treelets[at] = Treelet { slot, l, r };
}

template <typename T> BVH_HD void write_node(DevNode<T>* dst, const T bmin[3], const T bmax[3],
                                             typename Real<T>::UInt index, typename Real<T>::UInt pad = 0) {
    DevNode<T> n;
    for (int k = 0; k < 3; ++k) { n.bounds[2 * k] = bmin[k]; n.bounds[2 * k + 1] = bmax[k]; }
    n.index = index; n.pad = pad;
#if defined(__CUDA_ARCH__)
    // two (float) / four (double) 128-bit stores
    const uint4* s = reinterpret_cast<const uint4*>(&n);
    uint4* d = reinterpret_cast<uint4*>(dst);
    #pragma unroll
    for (int k = 0; k < (int)(sizeof(DevNode<T>) / 16); ++k) d[k] = s[k];
#else
    *dst = n;
#endif
}

// Compaction (lbvh_build.cu compact_* kernels; tests/host_emul.cpp): pair p moves to pair rank[p] = number of live
// pairs before it, so an inner node whose children sit at reference indices 2p+1, 2p+2 now points at 2*rank[p]+1;
// the spare word (cost + depth while building) is cleared.
template <typename T> BVH_HD DevNode<T> compact_remap(DevNode<T> n, const uint32_t* __restrict__ rank) {
    using U = typename Real<T>::UInt;
    if (index_count(n.index) == 0) {
        const uint32_t pair = (uint32_t)((index_first(n.index) - 1) >> 1);
        n.index = make_index<U>((U)(2 * (size_t)rank[pair] + 1), 0);
    }
    n.pad = 0;
    return n;
}

// Memory-ordering hooks: the device version uses __threadfence / atomicExch / L2 loads; the host
// emulation runs leaves sequentially, so plain accesses suffice.
struct HostSync {
    static inline void fence() {}
    static inline int exchange(int* p, int v) { int o = *p; *p = v; return o; }
    template <typename X> static inline X load(const X* p) { return *p; }
    static inline void store_max(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
};

// ---- the bottom-up pass, in pieces -----------------------------------------------------------------
// State of a node on its way up: the range of sorted primitives it covers, its box, its packed index
// (leaf or inner), SAH cost and depth.
template <typename T> struct ClimbState {
    uint32_t l, r;
    T bmin[3], bmax[3];
    typename Real<T>::UInt index;
    T cost;
    uint32_t depth;
};

template <typename T> BVH_HD void climb_init(ClimbState<T>& s, uint32_t i, const T bmin[3], const T bmax[3]) {
    using U = typename Real<T>::UInt;
    s.l = s.r = i;
    for (int k = 0; k < 3; ++k) { s.bmin[k] = bmin[k]; s.bmax[k] = bmax[k]; }
    s.index = make_index<U>((U)i, 1);                       // Index::make_leaf(i, 1)
    s.cost = half_area(bmin, bmax);                         // leaf cost = half_area * prim_count
    s.depth = 0;
}

// Choose the parent: merge across the more similar boundary (ties -> left boundary).  side 0: the node is
// the child covering [l, parent]; side 1: the child covering [parent+1, r].
template <typename K> BVH_HD void choose_parent(const K* __restrict__ keys, uint32_t n, uint32_t l, uint32_t r,
                                                uint32_t& parent, uint32_t& side) {
    if (l == 0 || (r != n - 1 && delta_less(delta_at(keys, r), delta_at(keys, l - 1)))) { parent = r; side = 0; }
    else { parent = l - 1; side = 1; }
}

// Device slot of child `side` of internal node `parent` (reference index 2p+1+side, +1 device shift).
BVH_HD size_t child_slot(uint32_t parent, uint32_t side) { return 2 * (size_t)parent + 1 + side + 1; }

// Publishes the node as child `side` of `parent`: its final record goes to its final slot.  Returns the
// record (the cost travels rounded to what the packed word holds, so both arrivals see the same value).
template <typename T> BVH_HD DevNode<T> publish_child(const BuildParams<T>& p, ClimbState<T>& s, uint32_t parent, uint32_t side) {
    using U = typename Real<T>::UInt;
    const U aux = AuxPack<T>::pack(s.cost, s.depth);
    s.cost = AuxPack<T>::cost(aux);
    DevNode<T> rec;
    for (int k = 0; k < 3; ++k) { rec.bounds[2 * k] = s.bmin[k]; rec.bounds[2 * k + 1] = s.bmax[k]; }
    rec.index = s.index; rec.pad = aux;
    write_node(p.nodes + child_slot(parent, side), s.bmin, s.bmax, s.index, aux);
    return rec;
}

// The second arrival at `parent` merges with its sibling's record `sn` (whose far range boundary is
// `other`): SATO swap of the two records if needed, box union in memory order, SAH leaf-collapse decision.
// Returns true when the merged node is the root (its record is then written to reference index 0).
template <typename T> BVH_HD bool merge_into_parent(const BuildParams<T>& p, ClimbState<T>& s, uint32_t parent, uint32_t side,
                                                    const DevNode<T>& own, const DevNode<T>& sn, uint32_t other) {
    using R = Real<T>;
    using U = typename R::UInt;
    const T sib_cost = AuxPack<T>::cost(sn.pad);
    const uint32_t sib_depth = AuxPack<T>::depth(sn.pad);
    T smin[3] = { sn.bounds[0], sn.bounds[2], sn.bounds[4] };
    T smax[3] = { sn.bounds[1], sn.bounds[3], sn.bounds[5] };
    const T own_area = half_area(s.bmin, s.bmax), sib_area = half_area(smin, smax);

    // SATO: the child with the larger half-area must be the LEFT one (top_down_sah_builder.h:101-108).
    const bool own_is_left = side == 0;
    const T left_area = own_is_left ? own_area : sib_area, right_area = own_is_left ? sib_area : own_area;
    const bool swap = left_area < right_area;
    if (swap) {
        write_node(p.nodes + child_slot(parent, 1 - side), s.bmin, s.bmax, own.index, own.pad);
        write_node(p.nodes + child_slot(parent, side), smin, smax, sn.index, sn.pad);
    }
    const uint32_t own_l = s.l, own_r = s.r;
    if (side == 0) s.r = other; else s.l = other;
    if (p.treelet_max != 0 && s.r - s.l + 1 > p.treelet_max) {
        // the parent is too large for a treelet: each child of 3..treelet_max primitives is a maximal one
        const uint32_t own_count = own_r - own_l + 1, sib_count = (s.r - s.l + 1) - own_count;
        const uint32_t own_side = swap ? 1 - side : side;
        if (own_count >= 3 && own_count <= p.treelet_max) append_treelet(p, (uint32_t)child_slot(parent, own_side), own_l, own_r);
        if (sib_count >= 3 && sib_count <= p.treelet_max) {
            const uint32_t sib_l = side == 0 ? parent + 1 : other, sib_r = side == 0 ? other : parent;
            append_treelet(p, (uint32_t)child_slot(parent, 1 - own_side), sib_l, sib_r);
        }
    }
    // parent box = left.get_bbox().extend(right.get_bbox()) (bvh.h:213-217), left/right as they now sit in memory
    const bool own_first = swap ? !own_is_left : own_is_left;
    for (int k = 0; k < 3; ++k) {
        const T a_min = own_first ? s.bmin[k] : smin[k], b_min = own_first ? smin[k] : s.bmin[k];
        const T a_max = own_first ? s.bmax[k] : smax[k], b_max = own_first ? smax[k] : s.bmax[k];
        s.bmin[k] = robust_min(a_min, b_min);
        s.bmax[k] = robust_max(a_max, b_max);
    }
    const uint32_t count = s.r - s.l + 1;
    const T area = half_area(s.bmin, s.bmax);
    const T split_cost = R::add(area, R::add(s.cost, sib_cost));   // node cost_ratio 1 (split_heuristic.h:18-24)
    const T leaf_cost = R::mul(area, (T)count);                   // get_leaf_cost, split_heuristic.h:30-33
    const uint32_t sub_depth = (s.depth > sib_depth ? s.depth : sib_depth) + 1;
    if (count <= p.max_leaf && (count <= p.min_leaf || leaf_cost <= split_cost)) {
        if (p.alive) {                                      // the pairs of the collapsed subtree are dead
            if (s.depth == 0 && sib_depth == 0) p.alive[parent] = 0;            // both children were leaves: only this pair
            else for (uint32_t q = s.l; q < s.r; ++q) p.alive[q] = 0;
        }
        s.index = make_index<U>((U)s.l, count);             // collapse the subtree into one leaf
        s.cost = leaf_cost; s.depth = 0;
    } else {
        s.index = make_index<U>((U)(2 * (size_t)parent + 1), 0);   // Index::make_inner(first child)
        s.cost = split_cost; s.depth = sub_depth;
    }
    if (s.l == 0 && s.r == p.n - 1) {                       // this is the root: reference index 0
        write_node(p.nodes + 1, s.bmin, s.bmax, s.index);
        p.info[0] = s.depth; p.info[1] = parent;
        if (p.treelet_max != 0 && p.n >= 3 && p.n <= p.treelet_max) append_treelet(p, 1u, 0u, p.n - 1);   // the whole tree is one treelet
        return true;
    }
    return false;
}

// Climb through GLOBAL memory from a node that has just been published as child `side` of `parent` (record
// `own`): exchange the arrival flag; the first arrival stops, the second merges and goes on.
template <typename T, typename K, typename Sync>
BVH_HD void climb_global(const BuildParams<T>& p, const K* __restrict__ keys, ClimbState<T>& s,
                         uint32_t parent, uint32_t side, DevNode<T> own) {
    for (;;) {
        Sync::fence();
        const int other = Sync::exchange(p.flags + parent, (int)(side == 0 ? s.l : s.r));
        if (other < 0) return;                              // first to arrive: the sibling will carry on
        Sync::fence();
        const DevNode<T> sn = Sync::load(p.nodes + child_slot(parent, 1 - side));
        if (merge_into_parent(p, s, parent, side, own, sn, (uint32_t)other)) return;
        choose_parent(keys, p.n, s.l, s.r, parent, side);
        own = publish_child(p, s, parent, side);
    }
}

// The whole pass for sorted leaf `i` through global memory only (host emulation; reference formulation).
template <typename T, typename K, typename Sync>
BVH_HD void build_bottom_up(const BuildParams<T>& p, const K* __restrict__ keys, uint32_t i,
                            T bmin[3], T bmax[3]) {
    ClimbState<T> s;
    climb_init(s, i, bmin, bmax);
    if (p.n == 1) {                                         // the root is a leaf (bvh.h:128 handles it)
        write_node(p.nodes + 1, bmin, bmax, s.index);
        p.info[0] = 0; p.info[1] = 0;
        return;
    }
    uint32_t parent, side;
    choose_parent(keys, p.n, s.l, s.r, parent, side);
    const DevNode<T> own = publish_child(p, s, parent, side);
    climb_global<T, K, Sync>(p, keys, s, parent, side, own);
}

// ---- block-local first phase -------------------------------------------------------------------------
// A block of consecutive sorted leaves [i0, iend) first merges, through SHARED memory, every pair of nodes
// that meets at a boundary strictly inside the block (the merges of the final tree whose both children lie
// in the block — about 95 % of all merges for 128-leaf blocks); only nodes whose parent boundary is a block
// wall, or whose sibling never showed up locally, continue through global memory (climb_global).  The tree
// is the same as with the global-only pass: the same merges happen, only where the two arrivals meet differs.
template <typename T> struct LocalSlots {
    DevNode<T>* nodes;      // [2 * block_leaves]: record of child `side` of boundary p at 2 * (p - i0) + side
    int* flags;             // [block_leaves]: -1, or the far boundary of the first arrival
    int* info;              // [block_leaves]: non-zero once a second arrival has matched the first
};

enum StepOutcome : int { kStepRetired = 0, kStepCarry = 1, kStepWall = 2 };

// One step of one node in the local phase: choose the parent, publish, meet the sibling if it is there.
//   kStepRetired  first arrival at a local boundary (its record waits in loc.nodes; after the block barrier
//                 loc.info[parent - i0] tells whether a sibling came), or the root was written
//   kStepCarry    merged with its sibling: `s` is now the parent node, to be stepped again
//   kStepWall     the chosen boundary is a block wall: (s, parent, side) continues with climb_global
// In the last two cases and for an unmatched first arrival the node's record is already published in global
// memory and equals published_record(s).
template <typename T, typename K, typename LocalSync>
BVH_HD int local_step(const BuildParams<T>& p, const K* keys, ClimbState<T>& s, const LocalSlots<T>& loc,
                      uint32_t i0, uint32_t iend, uint32_t& parent, uint32_t& side) {
    choose_parent(keys, p.n, s.l, s.r, parent, side);
    const DevNode<T> own = publish_child(p, s, parent, side);
    if (parent < i0 || parent + 1 >= iend) return kStepWall;
    const uint32_t q = parent - i0;
    loc.nodes[2 * q + side] = own;
    LocalSync::fence();                                     // record before flag (release) ...
    const int other = LocalSync::exchange(loc.flags + q, (int)(side == 0 ? s.l : s.r));
    if (other < 0) return kStepRetired;
    LocalSync::fence();                                     // ... flag before the sibling's record (acquire)
    loc.info[q] = 2;
    const DevNode<T> sn = loc.nodes[2 * q + (1 - side)];
    return merge_into_parent(p, s, parent, side, own, sn, (uint32_t)other) ? kStepRetired : kStepCarry;
}

// The record a node has published as child `side` of `parent` (what publish_child wrote).
template <typename T> BVH_HD DevNode<T> published_record(const ClimbState<T>& s) {
    DevNode<T> rec;
    for (int k = 0; k < 3; ++k) { rec.bounds[2 * k] = s.bmin[k]; rec.bounds[2 * k + 1] = s.bmax[k]; }
    rec.index = s.index;
    rec.pad = AuxPack<T>::pack(s.cost, s.depth);
    return rec;
}

} // namespace bvhb200
