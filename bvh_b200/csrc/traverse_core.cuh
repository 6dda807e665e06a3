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

#if defined(__CUDACC__)
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&w)[8]) {
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}
// Node loads do NOT allocate their line in L1 (ld.global.nc.L1::no_allocate).  Measured on the B200 (profiles/
// r02_run10to11_*): soup-1M +6.4 %, incoherent c3 +8 %, grid-1M -0.5 %.  Deep nodes are used once per ray, and a pending miss
// that must allocate a line is limited by the L1's capacity: with 28 KB of L1 (the 228 KB shared-memory carve-out) the
// incoherent batch ran at 1.0 instead of 1.7 Grays/s — the lines in flight, not the hits, were what the larger L1 bought.
#ifndef BVH_NODE_NA
#define BVH_NODE_NA 1
#endif
#ifndef BVH_TRI_NA
#define BVH_TRI_NA 0
#endif
__device__ __forceinline__ void ldg256_na(const void* p, uint32_t (&w)[8]) {
    asm("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}
__device__ __forceinline__ void ldg256_node(const void* p, uint32_t (&w)[8]) {
#if BVH_NODE_NA
    ldg256_na(p, w);
#else
    ldg256(p, w);
#endif
}
__device__ __forceinline__ float4 ldg128_tri(const float4* p) {
#if BVH_TRI_NA
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ void ldg256(const void* p, unsigned long long (&w)[4]) {
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p));
}
__device__ __forceinline__ void ldg256_node(const void* p, unsigned long long (&w)[4]) {
#if BVH_NODE_NA
    asm("ld.global.nc.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p));
#else
    ldg256(p, w);
#endif
}
#endif

#if defined(__CUDA_ARCH__)
// sm_100 has 256-bit global loads (SASS LDG.E.256): a node is ONE load, a sibling pair two, from one
// naturally aligned 64-byte (float) / 128-byte (double) block, through the read-only path.  Halving the
// number of load instructions matters because with divergent rays every load costs one L1 tag lookup per
// distinct line touched by the warp (profiles/: the traversal is L1-tag bound, not DRAM bound).
__device__ __forceinline__ void load_pair(const DevNode<float>* __restrict__ p, NodePair<float>& o) {
    uint32_t a[8], b[8];
    ldg256_node(p, a);
    ldg256_node(p + 1, b);
    #pragma unroll
    for (int k = 0; k < 6; ++k) { o.lb[k] = __uint_as_float(a[k]); o.rb[k] = __uint_as_float(b[k]); }
    o.li = a[6]; o.ri = b[6];
}
__device__ __forceinline__ void load_pair(const DevNode<double>* __restrict__ p, NodePair<double>& o) {
    unsigned long long a[4], b[4], c[4], d[4];
    const unsigned char* q = reinterpret_cast<const unsigned char*>(p);
    ldg256_node(q, a); ldg256_node(q + 32, b); ldg256_node(q + 64, c); ldg256_node(q + 96, d);
    #pragma unroll
    for (int k = 0; k < 4; ++k) { o.lb[k] = __longlong_as_double((long long)a[k]); o.rb[k] = __longlong_as_double((long long)c[k]); }
    o.lb[4] = __longlong_as_double((long long)b[0]); o.lb[5] = __longlong_as_double((long long)b[1]); o.li = b[2];
    o.rb[4] = __longlong_as_double((long long)d[0]); o.rb[5] = __longlong_as_double((long long)d[1]); o.ri = d[2];
}
// One node (for the lane-pair kernel: each lane of a pair fetches one child of the sibling pair).
__device__ __forceinline__ void load_node(const DevNode<float>* __restrict__ p, float (&b)[6], uint32_t& index) {
    uint32_t a[8];
    ldg256(p, a);
    #pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = __uint_as_float(a[k]);
    index = a[6];
}
__device__ __forceinline__ void load_node(const DevNode<double>* __restrict__ p, double (&b)[6], uint64_t& index) {
    unsigned long long a[4], c[4];
    const unsigned char* q = reinterpret_cast<const unsigned char*>(p);
    ldg256(q, a); ldg256(q + 32, c);
    #pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = __longlong_as_double((long long)a[k]);
    b[4] = __longlong_as_double((long long)c[0]); b[5] = __longlong_as_double((long long)c[1]);
    index = c[2];
}
#if BVH_TRI_PAD
__device__ __forceinline__ void load_tri(const DevTri<float>* __restrict__ p, DevTri<float>& o) {
    uint32_t a[8], b[8];
    ldg256(p, a);
    ldg256(reinterpret_cast<const unsigned char*>(p) + 32, b);
    o.p0[0] = __uint_as_float(a[0]); o.p0[1] = __uint_as_float(a[1]); o.p0[2] = __uint_as_float(a[2]);
    o.e1[0] = __uint_as_float(a[3]); o.e1[1] = __uint_as_float(a[4]); o.e1[2] = __uint_as_float(a[5]);
    o.e2[0] = __uint_as_float(a[6]); o.e2[1] = __uint_as_float(a[7]); o.e2[2] = __uint_as_float(b[0]);
    o.n[0] = __uint_as_float(b[1]); o.n[1] = __uint_as_float(b[2]); o.n[2] = __uint_as_float(b[3]);
}
__device__ __forceinline__ void load_tri(const DevTri<double>* __restrict__ p, DevTri<double>& o) {
    unsigned long long a[4], b[4], c[4];
    const unsigned char* q = reinterpret_cast<const unsigned char*>(p);
    ldg256(q, a); ldg256(q + 32, b); ldg256(q + 64, c);
    o.p0[0] = __longlong_as_double((long long)a[0]); o.p0[1] = __longlong_as_double((long long)a[1]);
    o.p0[2] = __longlong_as_double((long long)a[2]); o.e1[0] = __longlong_as_double((long long)a[3]);
    o.e1[1] = __longlong_as_double((long long)b[0]); o.e1[2] = __longlong_as_double((long long)b[1]);
    o.e2[0] = __longlong_as_double((long long)b[2]); o.e2[1] = __longlong_as_double((long long)b[3]);
    o.e2[2] = __longlong_as_double((long long)c[0]); o.n[0] = __longlong_as_double((long long)c[1]);
    o.n[1] = __longlong_as_double((long long)c[2]); o.n[2] = __longlong_as_double((long long)c[3]);
}
#else
__device__ __forceinline__ void load_tri(const DevTri<float>* __restrict__ p, DevTri<float>& o) {
    const float4* q = reinterpret_cast<const float4*>(p);
    const float4 a = ldg128_tri(q), b = ldg128_tri(q + 1), c = ldg128_tri(q + 2);
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
#endif
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
    // Same decisions as the reference's nested branches (bvh.h:167-181), written as selects: on the device the lanes of a
    // warp take all four outcomes at once, and one predicated path costs fewer issue slots than four divergent ones.
    const bool both = hit_left && hit_right;
    const bool right_first = !kAny && both && l0 > r0;            // ties go left first (`>` at bvh.h:180)
    const U near_index = (hit_left && !right_first) ? pair.li : pair.ri;
    const U far_index = right_first ? pair.li : pair.ri;
    if (both) stack.push(far_index);
    if (hit_left || hit_right) { top = near_index; return true; }
    return stack.try_pop(top);
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
    if (ray_interval_is_nan(r)) return;                         // can never hit (node.h:110-115, tri.h:69)
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
        if (!stack.try_pop(top)) break;
    }
}

} // namespace bvhb200
