// bvh_b200/csrc/treelet_sah.cuh — SAH rebuild of the bottom of the LBVH ("treelets"): the second build pass of
// DefaultBuilder::Quality Medium and High (Low keeps the plain LBVH).
//
// Why: measured, the 11-23 % more traversal steps the LBVH needs compared with the reference's SAH
// trees come from the bottom levels: keeping the LBVH above and rebuilding every maximal subtree of at most
// kMaxPrims primitives with the reference's sweep-SAH rule recovers the reference's step counts on the soup and
// about half of the gap on the height-field, while a SAH top level over LBVH subtrees recovers nothing.
//
// What: every maximal LBVH subtree of 3..kMaxPrims primitives (listed by the bottom-up pass itself, build_core.cuh
// merge_into_parent) is rebuilt top-down by ONE WARP with the greedy rule of the reference's SweepSahBuilder /
// TopDownSahBuilder (sweep_sah_builder.h:57-139, top_down_sah_builder.h:74-131): primitives sorted once along each
// axis, per node the split minimising area(L)*|L| + area(R)*|R| over the three axes and all positions, leaf when no
// split beats area*(count - 1) and count <= max_leaf_size, median split on the largest axis otherwise, stable
// partition of the other two orders, larger-area child first (SATO).  All nodes of one LEVEL of the treelet are
// processed together: the three orders keep each node's primitives in one contiguous segment, and prefix / suffix
// boxes, split costs, minima and partitions are segmented scans over the whole array.  The subtree is written into
// the node slots the LBVH subtree owned (pairs l .. r-1 of the sorted range [l, r]), primitives are re-ordered
// inside [l, r] only, and the subtree's box is unchanged, so nothing above the treelet moves.
//
// The algorithm is written as PHASES over array positions, parametrised by an execution policy: on the device a
// phase is a lane-strided loop of the warp followed by __syncwarp() (the scratch of a treelet lives in the warp's
// slice of shared memory), in the host emulation a plain loop — the same source, the same arithmetic (Real<T> ops,
// no contraction), hence the same tree.  (Round 1 ran one 256-thread block per 256-primitive treelet: ~540
// __syncthreads-separated phases per treelet, 4.3 ms per million triangles on the B200; with 64-primitive treelets
// a warp covers the positions in two strides and a phase costs a __syncwarp.)
#pragma once

#include "build_core.cuh"

namespace bvhb200 {

// kMaxPrims: largest treelet; kChunk: positions scanned sequentially by one iteration of the blocked scans
// (kMaxPrims / 32: one chunk per lane).
template <typename T> struct TreeletCfg { static constexpr int kMaxPrims = 64; static constexpr int kChunk = kMaxPrims / 32; };

// ---- execution policies --------------------------------------------------------------------------------
struct HostExec {
    template <typename F> static void phase(uint32_t n, F f) { for (uint32_t i = 0; i < n; ++i) f(i); }
    static uint32_t atomic_add(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
    static void atomic_max(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
};
// Same phases, iterations in descending order: a phase whose iterations depended on each other (a hazard between
// the threads of the block on the device) would give a different tree (tests/test_host_emulation.py).
struct HostExecReversed {
    template <typename F> static void phase(uint32_t n, F f) { for (uint32_t i = n; i-- > 0;) f(i); }
    static uint32_t atomic_add(uint32_t* p, uint32_t v) { return HostExec::atomic_add(p, v); }
    static void atomic_max(uint32_t* p, uint32_t v) { HostExec::atomic_max(p, v); }
};
#if defined(__CUDACC__)
struct WarpExec {
    template <typename F> static __device__ __forceinline__ void phase(uint32_t n, F f) {
        for (uint32_t i = threadIdx.x & 31u; i < n; i += 32u) f(i);
        __syncwarp();
    }
    static __device__ __forceinline__ uint32_t atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
    static __device__ __forceinline__ void atomic_max(uint32_t* p, uint32_t v) { atomicMax(p, v); }
};
#endif

// ---- scratch of one treelet (shared memory on the device) ------------------------------------------
// Segmented scans are BLOCKED: a chunk of kChunk consecutive positions is scanned sequentially by one
// iteration, the chunk aggregates by a short data-parallel scan, and the carry is applied per position —
// O(n) work per scan instead of the O(n log n) of a flat data-parallel scan.
template <typename T, int S> struct TreeletScratch {
    static constexpr int kChunk = S / 32 > 0 ? S / 32 : 1;
    static constexpr int NC = S / kChunk;
    T box[6][S];                         // per PRIMITIVE (local index): minx,maxx,miny,maxy,minz,maxz
    T centre[3][S];
    uint32_t old_ids[S];                 // prim_ids[l + i] before the rebuild
    uint16_t order[3][S];                // local primitive indices sorted along each axis, segment by segment
    uint16_t tmp16[S];
    uint16_t seg_begin[S], seg_end[S];   // per POSITION: its segment; seg_end == 0: the segment became a leaf
    T pre[6][S], suf[6][S];              // boxes of [segment begin, pos] and of [pos, segment end) on the current axis
    T right_cost[S];                     // cost of [pos, segment end) as a leaf
    T cand_cost[S];                      // cost of splitting after pos, then its running first minimum
    uint16_t cand_pos[S];                // ... the position that goes with it; later the running sums of the side flags
    // chunk aggregates (ping-pong) of the blocked scans
    T cpre[2][6][NC], csuf[2][6][NC], ccost[2][NC];
    uint16_t cpos[2][NC];
    uint8_t cfl[2][NC], cbl[2][NC];      // the chunk contains a forward / backward reset
    // per SEGMENT, stored at the position of its head (= seg_begin)
    T nbox[6][S];
    T leaf_cost[S], best_cost[S], best_larea[S], best_rarea[S];
    uint16_t best_pos[S];
    uint8_t best_axis[S], split[S], seg_depth[S];
    uint32_t dst_slot[S], left_slot[S], right_slot[S];
    uint8_t side[S];                     // per PRIMITIVE: 1 = goes to the left part
    uint32_t counters[4];                // [0] unused, [1] live segments, [2] treelet depth, [3] longest live segment
};
static_assert(sizeof(TreeletScratch<float, TreeletCfg<float>::kMaxPrims>) <= 48 * 1024, "one warp's slice of shared memory");
static_assert(sizeof(TreeletScratch<double, TreeletCfg<double>::kMaxPrims>) <= 48 * 1024, "one warp's slice of shared memory");

// Monotone map of a scalar to an unsigned integer (negative values reversed below the positive ones): a total
// order even when a centre is a NaN, so the rank sort below always yields a permutation.
template <typename T> BVH_HD typename Real<T>::UInt treelet_sort_key(T x) {
    using U = typename Real<T>::UInt;
    const U bits = Real<T>::bits(x), sign = (U)1 << (8 * sizeof(U) - 1);
    return (bits & sign) ? (U)~bits : (U)(bits | sign);
}

template <typename T> BVH_HD T treelet_inf() { return Real<T>::from_bits(sizeof(T) == 4 ? (typename Real<T>::UInt)0x7F800000u : (typename Real<T>::UInt)0x7FF0000000000000ull); }

template <typename T, int S> BVH_HD T treelet_half_area(const T (&b)[6][S], uint32_t pos) {
    const T mn[3] = { b[0][pos], b[2][pos], b[4][pos] }, mx[3] = { b[1][pos], b[3][pos], b[5][pos] };
    return half_area(mn, mx);
}

// Rebuilds one treelet.  `leaf_src`: vertices (n x 9, leaf_mode 0) or boxes (n x 6, leaf_mode 1); `centre_src`:
// vertices or centres (n x 3), indexed by ORIGINAL primitive id.  `tris` (leaf_mode 0 only) receives the
// BVH-order triangle records of the range.  info[2] collects by how much a treelet got deeper than the subtree
// it replaces (the traversal stack is sized from info[0] + info[2]).
template <typename T, int S, typename Exec>
BVH_HD void treelet_rebuild(TreeletScratch<T, S>& w, const Treelet& t, DevNode<T>* __restrict__ nodes,
                            uint32_t* __restrict__ prim_ids, DevTri<T>* __restrict__ tris,
                            const T* __restrict__ leaf_src, const T* __restrict__ centre_src, int leaf_mode,
                            uint32_t min_leaf, uint32_t max_leaf, uint32_t* __restrict__ info, uint32_t lbvh_depth,
                            uint32_t* __restrict__ alive = nullptr) {
    using R = Real<T>;
    using U = typename R::UInt;
    constexpr uint32_t C = TreeletScratch<T, S>::kChunk;
    const uint32_t n = t.r - t.l + 1, l = t.l;
    const uint32_t nc = (n + C - 1) / C;                     // chunks in use
    const T inf = treelet_inf<T>();
    // depth of the subtree being replaced (its root record still carries it; the tree's root record does not)
    const uint32_t old_depth = t.slot == 1 ? lbvh_depth : AuxPack<T>::depth(nodes[t.slot].pad);

    // a scan running forward restarts at a segment head, one running backward at a segment's last position;
    // positions of finished segments always restart
    auto fwd_reset = [&] (uint32_t p) { return w.seg_end[p] == 0 || w.seg_begin[p] == p; };
    auto bwd_reset = [&] (uint32_t p) { return w.seg_end[p] == 0 || p + 1 == w.seg_end[p]; };

    // ---- load the primitives ----
    Exec::phase(n, [&] (uint32_t i) {
        const uint32_t id = prim_ids[l + i];
        w.old_ids[i] = id;
        T bmin[3], bmax[3], c[3];
        if (leaf_mode == 0) {
            T v[9];
            for (int k = 0; k < 9; ++k) v[k] = leaf_src[9 * (size_t)id + k];
            tri_bounds_center(v, bmin, bmax, c);
        } else {
            for (int k = 0; k < 3; ++k) {
                bmin[k] = leaf_src[6 * (size_t)id + k]; bmax[k] = leaf_src[6 * (size_t)id + 3 + k];
                c[k] = centre_src[3 * (size_t)id + k];
            }
        }
        for (int k = 0; k < 3; ++k) { w.box[2 * k][i] = bmin[k]; w.box[2 * k + 1][i] = bmax[k]; w.centre[k][i] = c[k]; }
        w.seg_begin[i] = 0; w.seg_end[i] = (uint16_t)n;
        if (alive && i + 1 < n) alive[l + i] = 0u;           // the subtree's pairs: the splits below mark the ones they use
        if (i == 0) {
            w.dst_slot[0] = t.slot; w.seg_depth[0] = 0;
            w.counters[0] = 0; w.counters[1] = 1; w.counters[2] = 0; w.counters[3] = n;
        }
    });
    // ---- sort once along each axis: rank of a primitive = number of primitives before it in (centre, index) order
    Exec::phase(3 * n, [&] (uint32_t x) {
        const uint32_t a = x / n, i = x - a * n;
        const U ki = treelet_sort_key(w.centre[a][i]);
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const U kj = treelet_sort_key(w.centre[a][j]);
            rank += (kj < ki || (kj == ki && j < i)) ? 1u : 0u;
        }
        w.order[a][rank] = (uint16_t)i;
    });

    // ---- level by level ----
    while (w.counters[1] != 0) {
        const uint32_t reach = (w.counters[3] + C - 1) / C;   // a segment's head chunk is at most this many chunks before its last chunk
        for (int a = 0; a < 3; ++a) {
            // ---- prefix boxes (pre) and suffix boxes (suf) of every live segment along axis a ----
            Exec::phase(n, [&] (uint32_t pos) {
                const uint32_t prim = w.order[a][pos];
                for (int c = 0; c < 6; ++c) { const T v = w.box[c][prim]; w.pre[c][pos] = v; w.suf[c][pos] = v; }
            });
            Exec::phase(2 * nc, [&] (uint32_t x) {                                       // inside each chunk, sequentially
                const uint32_t c = x < nc ? x : x - nc, first = c * C, last = (first + C < n ? first + C : n) - 1;
                bool flag = false;
                if (x < nc) {
                    for (uint32_t p = first; p <= last; ++p) {
                        const bool reset = fwd_reset(p);
                        flag = flag || reset;
                        if (p > first && !reset)
                            for (int k = 0; k < 6; k += 2) {
                                w.pre[k][p] = robust_min(w.pre[k][p], w.pre[k][p - 1]);
                                w.pre[k + 1][p] = robust_max(w.pre[k + 1][p], w.pre[k + 1][p - 1]);
                            }
                    }
                    for (int k = 0; k < 6; ++k) w.cpre[0][k][c] = w.pre[k][last];
                    w.cfl[0][c] = flag;
                } else {
                    for (uint32_t p = last + 1; p-- > first;) {
                        const bool reset = bwd_reset(p);
                        flag = flag || reset;
                        if (p < last && !reset)
                            for (int k = 0; k < 6; k += 2) {
                                w.suf[k][p] = robust_min(w.suf[k][p], w.suf[k][p + 1]);
                                w.suf[k + 1][p] = robust_max(w.suf[k + 1][p], w.suf[k + 1][p + 1]);
                            }
                    }
                    for (int k = 0; k < 6; ++k) w.csuf[0][k][c] = w.suf[k][first];
                    w.cbl[0][c] = flag;
                }
            });
            int cur = 0;
            for (uint32_t d = 1; d <= reach && d < nc; d <<= 1) {                        // across the chunks, data-parallel
                Exec::phase(2 * nc, [&] (uint32_t x) {
                    if (x < nc) {
                        const uint32_t c = x;
                        const bool take = c >= d && !w.cfl[cur][c];
                        for (int k = 0; k < 6; k += 2) {
                            const T mn = w.cpre[cur][k][c], mx = w.cpre[cur][k + 1][c];
                            w.cpre[cur ^ 1][k][c]     = take ? robust_min(mn, w.cpre[cur][k][c - d]) : mn;
                            w.cpre[cur ^ 1][k + 1][c] = take ? robust_max(mx, w.cpre[cur][k + 1][c - d]) : mx;
                        }
                        w.cfl[cur ^ 1][c] = take ? w.cfl[cur][c - d] : w.cfl[cur][c];
                    } else {
                        const uint32_t c = x - nc;
                        const bool take = c + d < nc && !w.cbl[cur][c];
                        for (int k = 0; k < 6; k += 2) {
                            const T mn = w.csuf[cur][k][c], mx = w.csuf[cur][k + 1][c];
                            w.csuf[cur ^ 1][k][c]     = take ? robust_min(mn, w.csuf[cur][k][c + d]) : mn;
                            w.csuf[cur ^ 1][k + 1][c] = take ? robust_max(mx, w.csuf[cur][k + 1][c + d]) : mx;
                        }
                        w.cbl[cur ^ 1][c] = take ? w.cbl[cur][c + d] : w.cbl[cur][c];
                    }
                });
                cur ^= 1;
            }
            Exec::phase(n, [&] (uint32_t pos) {                                          // carries in; leaf costs; the node's box at the head
                const uint32_t se = w.seg_end[pos];
                if (se == 0) return;
                const uint32_t c = pos / C, sb = w.seg_begin[pos];
                if (c > 0 && sb < c * C)
                    for (int k = 0; k < 6; k += 2) {
                        w.pre[k][pos] = robust_min(w.pre[k][pos], w.cpre[cur][k][c - 1]);
                        w.pre[k + 1][pos] = robust_max(w.pre[k + 1][pos], w.cpre[cur][k + 1][c - 1]);
                    }
                if (c + 1 < nc && se > (c + 1) * C)
                    for (int k = 0; k < 6; k += 2) {
                        w.suf[k][pos] = robust_min(w.suf[k][pos], w.csuf[cur][k][c + 1]);
                        w.suf[k + 1][pos] = robust_max(w.suf[k + 1][pos], w.csuf[cur][k + 1][c + 1]);
                    }
                const T area = treelet_half_area(w.suf, pos);
                w.right_cost[pos] = R::mul(area, (T)(se - pos));                         // get_leaf_cost, split_heuristic.h:30-33
                if (a == 0 && sb == pos) {                                               // the head: suf is the node's box
                    for (int k = 0; k < 6; ++k) w.nbox[k][pos] = w.suf[k][pos];
                    const T lc = R::mul(area, (T)(se - pos - 1));                        // get_non_split_cost, :35-38 (cost_ratio 1)
                    w.leaf_cost[pos] = lc; w.best_cost[pos] = lc;
                    w.best_pos[pos] = (uint16_t)((pos + se + 1) / 2); w.best_axis[pos] = 0;      // sweep_sah_builder.h:111
                    w.best_larea[pos] = (T)0; w.best_rarea[pos] = (T)0;
                }
            });
            // ---- cost of splitting after pos (left = [begin, pos], right = [pos + 1, end)); first minimum per segment ----
            Exec::phase(n, [&] (uint32_t pos) {
                const uint32_t se = w.seg_end[pos];
                T cost = inf;
                if (se != 0 && pos + 1 < se)
                    cost = R::add(R::mul(treelet_half_area(w.pre, pos), (T)(pos + 1 - w.seg_begin[pos])), w.right_cost[pos + 1]);
                w.cand_cost[pos] = cost; w.cand_pos[pos] = (uint16_t)(pos + 1);
            });
            Exec::phase(nc, [&] (uint32_t c) {
                const uint32_t first = c * C, last = (first + C < n ? first + C : n) - 1;
                bool flag = false;
                for (uint32_t p = first; p <= last; ++p) {
                    const bool reset = fwd_reset(p);
                    flag = flag || reset;
                    if (p > first && !reset && w.cand_cost[p - 1] <= w.cand_cost[p]) {    // ties: the earlier position (strict < in the reference's sweep)
                        w.cand_cost[p] = w.cand_cost[p - 1]; w.cand_pos[p] = w.cand_pos[p - 1];
                    }
                }
                w.ccost[0][c] = w.cand_cost[last]; w.cpos[0][c] = w.cand_pos[last]; w.cfl[0][c] = flag;
            });
            cur = 0;
            for (uint32_t d = 1; d <= reach && d < nc; d <<= 1) {
                Exec::phase(nc, [&] (uint32_t c) {
                    T cost = w.ccost[cur][c]; uint16_t k = w.cpos[cur][c]; uint8_t flag = w.cfl[cur][c];
                    if (c >= d && !flag) {
                        if (w.ccost[cur][c - d] <= cost) { cost = w.ccost[cur][c - d]; k = w.cpos[cur][c - d]; }
                        flag = w.cfl[cur][c - d];
                    }
                    w.ccost[cur ^ 1][c] = cost; w.cpos[cur ^ 1][c] = k; w.cfl[cur ^ 1][c] = flag;
                });
                cur ^= 1;
            }
            Exec::phase(n, [&] (uint32_t pos) {                                          // heads: keep the best axis (first one on ties)
                const uint32_t se = w.seg_end[pos];
                if (se == 0 || w.seg_begin[pos] != pos || se - pos < 2) return;
                const uint32_t last = se - 1, c = last / C;
                T cost = w.cand_cost[last]; uint32_t k = w.cand_pos[last];
                if (c > 0 && pos < c * C && w.ccost[cur][c - 1] <= cost) { cost = w.ccost[cur][c - 1]; k = w.cpos[cur][c - 1]; }
                if (cost < w.best_cost[pos]) {
                    w.best_cost[pos] = cost; w.best_pos[pos] = (uint16_t)k; w.best_axis[pos] = (uint8_t)a;
                    w.best_larea[pos] = treelet_half_area(w.pre, k - 1);
                    w.best_rarea[pos] = treelet_half_area(w.suf, k);
                }
            });
        }

        // ---- decide every live segment: leaf, SAH split or median fallback; write the node record ----
        Exec::phase(n, [&] (uint32_t pos) {
            const uint32_t se = w.seg_end[pos];
            if (se == 0 || w.seg_begin[pos] != pos) return;
            const uint32_t count = se - pos;
            bool do_split = false;
            if (count > min_leaf) {                                                      // top_down_sah_builder.h:89
                if (w.best_cost[pos] < w.leaf_cost[pos]) do_split = true;                // sweep_sah_builder.h:116
                else if (count > max_leaf) {                                             // :117-123: median on the largest axis
                    const T d0 = R::sub(w.nbox[1][pos], w.nbox[0][pos]), d1 = R::sub(w.nbox[3][pos], w.nbox[2][pos]),
                            d2 = R::sub(w.nbox[5][pos], w.nbox[4][pos]);
                    uint8_t axis = 0;                                                    // Vec::get_largest_axis
                    if (d0 < d1) axis = 1;
                    if ((axis == 0 ? d0 : d1) < d2) axis = 2;
                    w.best_pos[pos] = (uint16_t)((pos + se + 1) / 2); w.best_axis[pos] = axis;
                    w.best_larea[pos] = (T)0; w.best_rarea[pos] = (T)0;                  // (no SATO swap for the fallback)
                    do_split = true;
                }
            }
            const T bmin[3] = { w.nbox[0][pos], w.nbox[2][pos], w.nbox[4][pos] }, bmax[3] = { w.nbox[1][pos], w.nbox[3][pos], w.nbox[5][pos] };
            if (do_split) {
                // one of the pairs l .. r-1 the LBVH subtree owned: the one of the split boundary, as in the LBVH numbering
                // (every inner node of a tree over a contiguous range splits at a different boundary)
                const uint32_t pair = l + (uint32_t)w.best_pos[pos] - 1u;
                if (alive) alive[pair] = 1u;
                const bool swap = w.best_larea[pos] < w.best_rarea[pos];                 // SATO, top_down_sah_builder.h:101-108
                w.left_slot[pos]  = (uint32_t)child_slot(pair, swap ? 1 : 0);
                w.right_slot[pos] = (uint32_t)child_slot(pair, swap ? 0 : 1);
                write_node(nodes + w.dst_slot[pos], bmin, bmax, make_index<U>((U)(2 * (size_t)pair + 1), 0));
                w.split[pos] = 1;
            } else {
                write_node(nodes + w.dst_slot[pos], bmin, bmax, make_index<U>((U)(l + pos), count));
                Exec::atomic_max(&w.counters[2], (uint32_t)w.seg_depth[pos]);
                w.split[pos] = 0;
            }
        });

        // ---- partition: mark sides on the split axis, stable-partition the two other orders ----
        Exec::phase(n, [&] (uint32_t pos) {
            const uint32_t se = w.seg_end[pos];
            if (se == 0) return;
            const uint32_t h = w.seg_begin[pos];
            if (w.split[h]) w.side[w.order[w.best_axis[h]][pos]] = pos < w.best_pos[h] ? 1 : 0;
        });
        for (int b = 0; b < 3; ++b) {
            auto moves = [&] (uint32_t pos) {                                            // does position pos of order b get partitioned?
                if (w.seg_end[pos] == 0) return false;
                const uint32_t h = w.seg_begin[pos];
                return w.split[h] != 0 && w.best_axis[h] != b;
            };
            Exec::phase(nc, [&] (uint32_t c) {                                           // running count of left-goers inside each chunk
                const uint32_t first = c * C, last = (first + C < n ? first + C : n) - 1;
                bool flag = false;
                for (uint32_t p = first; p <= last; ++p) {
                    const bool reset = fwd_reset(p);
                    flag = flag || reset;
                    uint16_t v = moves(p) ? w.side[w.order[b][p]] : (uint16_t)0;
                    if (p > first && !reset) v = (uint16_t)(v + w.cand_pos[p - 1]);
                    w.cand_pos[p] = v;
                }
                w.cpos[0][c] = w.cand_pos[last]; w.cfl[0][c] = flag;
            });
            int cur = 0;
            for (uint32_t d = 1; d <= reach && d < nc; d <<= 1) {
                Exec::phase(nc, [&] (uint32_t c) {
                    uint16_t v = w.cpos[cur][c]; uint8_t flag = w.cfl[cur][c];
                    if (c >= d && !flag) { v = (uint16_t)(v + w.cpos[cur][c - d]); flag = w.cfl[cur][c - d]; }
                    w.cpos[cur ^ 1][c] = v; w.cfl[cur ^ 1][c] = flag;
                });
                cur ^= 1;
            }
            Exec::phase(n, [&] (uint32_t pos) {
                uint32_t dst = pos;
                if (moves(pos)) {
                    const uint32_t h = w.seg_begin[pos], c = pos / C, k = w.best_pos[h];
                    const uint32_t f = w.side[w.order[b][pos]];
                    const uint32_t incl = w.cand_pos[pos] + (c > 0 && h < c * C ? w.cpos[cur][c - 1] : 0u);
                    const uint32_t before = incl - f;
                    dst = f ? h + before : k + (pos - h - before);
                }
                w.tmp16[dst] = w.order[b][pos];
            });
            Exec::phase(n, [&] (uint32_t pos) { w.order[b][pos] = w.tmp16[pos]; });
        }
        // ---- the children become the segments of the next level ----
        Exec::phase(n, [&] (uint32_t pos) {                                              // heads hand slots and depth to both children
            const uint32_t se = w.seg_end[pos];
            if (se == 0 || w.seg_begin[pos] != pos || !w.split[pos]) return;
            const uint32_t k = w.best_pos[pos];
            const uint8_t depth = (uint8_t)(w.seg_depth[pos] + 1);
            w.dst_slot[k] = w.right_slot[pos]; w.seg_depth[k] = depth;
            w.dst_slot[pos] = w.left_slot[pos]; w.seg_depth[pos] = depth;
        });
        Exec::phase(n, [&] (uint32_t pos) {                                              // positions move to their child segment
            if (pos == 0) { w.counters[1] = 0; w.counters[3] = 0; }
            const uint32_t se = w.seg_end[pos];
            if (se == 0) return;
            const uint32_t h = w.seg_begin[pos];                                         // (reads the OLD head's decision only)
            if (!w.split[h]) { w.seg_end[pos] = 0; return; }                             // its node became a leaf: finished
            const uint32_t k = w.best_pos[h];
            if (pos < k) w.seg_end[pos] = (uint16_t)k; else w.seg_begin[pos] = (uint16_t)k;
        });
        Exec::phase(n, [&] (uint32_t pos) {
            const uint32_t se = w.seg_end[pos];
            if (se != 0 && w.seg_begin[pos] == pos) { Exec::atomic_add(&w.counters[1], 1u); Exec::atomic_max(&w.counters[3], se - pos); }
        });
    }

    // ---- final primitive order = order along axis 0 (every leaf's primitives are contiguous in all three) ----
    Exec::phase(n, [&] (uint32_t pos) {
        const uint32_t id = w.old_ids[w.order[0][pos]];
        prim_ids[l + pos] = id;
        if (leaf_mode == 0 && tris) {
            T v[9];
            for (int k = 0; k < 9; ++k) v[k] = leaf_src[9 * (size_t)id + k];
            tris[l + pos] = precompute_tri(v);
        }
        if (pos == 0 && w.counters[2] > old_depth) Exec::atomic_max(info + 2, w.counters[2] - old_depth);
    });
}

} // namespace bvhb200
