// bvh_b200/csrc/traverse_core.cuh — the per-ray stack machine.
//
// A faithful restatement of Bvh::traverse_top_down / Bvh::intersect (reference bvh.h:124-182) for a
// single ray: near/far ordering by entry distance with ties going LEFT first (`>` at bvh.h:180),
// far child pushed, pop on a double miss, any-hit leaves as soon as a leaf reports a hit.  The leaf
// convention is the one every reference caller uses (benchmark.cpp:281-292): test each primitive of
// the leaf in order and shrink tmax on a hit.
//
// The same code runs in three places: the one-thread-per-ray kernel, the persistent kernel (which
// splits it into an inner phase and a leaf phase, traverse.cu) and the host emulation compiled by
// g++ for the CPU tests (tests/host_emul.cpp).
#pragma once

#include "core.cuh"

namespace bvhb200 {

// Two adjacent nodes (a sibling pair) in registers.
template <typename T> struct NodePair {
    T lb[6], rb[6];
    typename Real<T>::UInt li, ri;
};

#if defined(__CUDA_ARCH__)
// float: one 64-byte aligned pair = 4 x 128-bit loads through the read-only path
__device__ __forceinline__ void load_pair(const DevNode<float>* __restrict__ p, NodePair<float>& o) {
    const float4* q = reinterpret_cast<const float4*>(p);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    o.lb[0] = a.x; o.lb[1] = a.y; o.lb[2] = a.z; o.lb[3] = a.w; o.lb[4] = b.x; o.lb[5] = b.y;
    o.li = __float_as_uint(b.z);
    o.rb[0] = c.x; o.rb[1] = c.y; o.rb[2] = c.z; o.rb[3] = c.w; o.rb[4] = d.x; o.rb[5] = d.y;
    o.ri = __float_as_uint(d.z);
}
// double: one 128-byte aligned pair = 8 x 128-bit loads (the 4th and 8th carry index + pad)
__device__ __forceinline__ void load_pair(const DevNode<double>* __restrict__ p, NodePair<double>& o) {
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const ulonglong2 ai = __ldg(reinterpret_cast<const ulonglong2*>(q + 3));
    const double2 d = __ldg(q + 4), e = __ldg(q + 5), f = __ldg(q + 6);
    const ulonglong2 di = __ldg(reinterpret_cast<const ulonglong2*>(q + 7));
    o.lb[0] = a.x; o.lb[1] = a.y; o.lb[2] = b.x; o.lb[3] = b.y; o.lb[4] = c.x; o.lb[5] = c.y;
    o.li = ai.x;
    o.rb[0] = d.x; o.rb[1] = d.y; o.rb[2] = e.x; o.rb[3] = e.y; o.rb[4] = f.x; o.rb[5] = f.y;
    o.ri = di.x;
}
__device__ __forceinline__ void load_tri(const DevTri<float>* __restrict__ p, DevTri<float>& o) {
    const float4* q = reinterpret_cast<const float4*>(p);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    o.p0[0] = a.x; o.p0[1] = a.y; o.p0[2] = a.z; o.e1[0] = a.w;
    o.e1[1] = b.x; o.e1[2] = b.y; o.e2[0] = b.z; o.e2[1] = b.w;
    o.e2[2] = c.x; o.n[0] = c.y; o.n[1] = c.z; o.n[2] = c.w;
}
__device__ __forceinline__ void load_tri(const DevTri<double>* __restrict__ p, DevTri<double>& o) {
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4), f = __ldg(q + 5);
    o.p0[0] = a.x; o.p0[1] = a.y; o.p0[2] = b.x; o.e1[0] = b.y; o.e1[1] = c.x; o.e1[2] = c.y;
    o.e2[0] = d.x; o.e2[1] = d.y; o.e2[2] = e.x; o.n[0] = e.y; o.n[1] = f.x; o.n[2] = f.y;
}
#else
template <typename T> inline void load_pair(const DevNode<T>* p, NodePair<T>& o) {
    for (int k = 0; k < 6; ++k) { o.lb[k] = p[0].bounds[k]; o.rb[k] = p[1].bounds[k]; }
    o.li = p[0].index; o.ri = p[1].index;
}
template <typename T> inline void load_tri(const DevTri<T>* p, DevTri<T>& o) { o = *p; }
#endif

// One inner step (reference bvh.h:132-150 with the lambda of :167-181).  Returns false when both
// children were missed and the stack is empty, i.e. the traversal is over.
template <typename T, bool kAny, bool kRobust, typename Stack>
BVH_HD bool inner_step(const DevNode<T>* __restrict__ nodes, const RayCtx<T>& r,
                       typename Real<T>::UInt& top, Stack& stack) {
    using U = typename Real<T>::UInt;
    NodePair<T> pair;
    load_pair(nodes + (size_t)index_first(top) + 1, pair);      // device slot = reference index + 1
    T l0, l1, r0, r1;
    node_test<T, kRobust>(pair.lb, r, l0, l1);
    node_test<T, kRobust>(pair.rb, r, r0, r1);
    const bool hit_left = l0 <= l1, hit_right = r0 <= r1;
    if (hit_left) {
        U near_index = pair.li;
        if (hit_right) {
            U far_index = pair.ri;
            if (!kAny && l0 > r0) { U tmp = near_index; near_index = far_index; far_index = tmp; }
            stack.push(far_index);
        }
        top = near_index;
    } else if (hit_right) {
        top = pair.ri;
    } else {
        if (stack.empty()) return false;
        top = stack.pop();
    }
    return true;
}

// Leaf processing (benchmark.cpp:281-292).  stats (nullable): [1] leaves, [2] triangle tests.
template <typename T>
BVH_HD void leaf_step(const DevTri<T>* __restrict__ tris, const uint32_t* __restrict__ prim_ids, bool lowest_id,
                      typename Real<T>::UInt top, RayCtx<T>& r, HitState<T>& hit, uint32_t* stats) {
    const uint32_t first = (uint32_t)index_first(top), count = index_count(top);
    if (stats) { stats[1] += 1; stats[2] += count; }
    for (uint32_t i = first; i < first + count; ++i) {
        DevTri<T> tri;
        load_tri(tris + i, tri);
        tri_test<T>(tri, i, prim_ids, lowest_id, r, hit);
    }
}

// Whole traversal for one ray.  stats (nullable): {inner steps, leaves, triangle tests}, the same
// three counters the reference exposes through its InnerFn hook (bvh.h:168, benchmark.cpp:282-296).
template <typename T, bool kAny, bool kRobust, typename Stack>
BVH_HD void traverse_ray(const DevNode<T>* __restrict__ nodes, const DevTri<T>* __restrict__ tris,
                         const uint32_t* __restrict__ prim_ids, bool lowest_id, typename Real<T>::UInt root_index,
                         RayCtx<T>& r, HitState<T>& hit, Stack& stack, uint32_t* stats) {
    typename Real<T>::UInt top = root_index;                    // get_root().index, bvh.h:54
    for (;;) {
        bool alive = true;
        while (index_count(top) == 0) {
            if (stats) stats[0] += 1;
            if (!inner_step<T, kAny, kRobust>(nodes, r, top, stack)) { alive = false; break; }
        }
        if (!alive) break;
        leaf_step<T>(tris, prim_ids, lowest_id, top, r, hit, stats);
        if (kAny && hit.slot != kInvalidId) break;              // bvh.h:153-155
        if (stack.empty()) break;
        top = stack.pop();
    }
}

} // namespace bvhb200
