// bvh_b200/csrc/reinsertion.h — subtree reinsertion on a host tree in the reference layout.
//
// What bvhNN_optimize / ReinsertionOptimizer<Node>::optimize do in the reference
// (reinsertion_optimizer.h:27-30,88-267; c_api/bvh_impl.h:223-233): per iteration, take the batch of nodes
// with the largest half-area, find for each the position in the tree where re-attaching it shrinks the summed
// node area most (branch and bound over the tree, read-only, :107-188), then apply the non-conflicting moves
// in order of decreasing gain (:190-267).  The tree is the caller-visible host mirror (or a caller-owned
// Bvh<Node>), i.e. host data by contract, so this pass runs on the host: the search of a batch is spread over
// plain std::threads (the reference spreads it over its ThreadPool), the application is sequential as in the
// reference.  The arithmetic follows the reference's order (half-area (d0+d1)*d2+d0*d1, bbox.h:32-38; union
// via robust_min/max, bbox.h:23-27) and the selection uses the same standard-library heap / sort calls, so that
// on the same input tree the result is the reference's tree, node for node (tests/test_optimize.py).
#pragma once

#include <algorithm>
#include <array>
#include <cstddef>
#include <functional>
#include <limits>
#include <thread>
#include <utility>
#include <vector>

#include "core.cuh"

namespace bvhb200 {

// NodeT: { T bounds[2 * kDim] as [min0,max0,min1,max1,...]; UInt index; } (reference Node<T, kDim>, node.h:31-37)
template <typename T, int kDim, typename NodeT>
class Reinserter {
public:
    Reinserter(std::vector<NodeT>& nodes, size_t threads) : nodes_(nodes), threads_(threads < 1 ? 1 : threads) {}

    // Returns false (tree untouched) when the node array is not a well-formed tree.
    bool run(T batch_size_ratio, size_t max_iter_count) {
        const size_t count = nodes_.size();
        if (count < 3) return true;
        if (!link_parents()) return false;
        const size_t batch = std::max<size_t>(1, (size_t)((T)count * batch_size_ratio));
        std::vector<Move> moves;
        std::vector<bool> locked(count);
        for (size_t iter = 0; iter < max_iter_count; ++iter) {
            const std::vector<Pick> picks = largest_nodes(batch);
            std::fill(locked.begin(), locked.end(), false);
            moves.assign(picks.size(), Move {});
            parallel_for(picks.size(), [&] (size_t i) { moves[i] = best_move(picks[i].node); });
            moves.erase(std::remove_if(moves.begin(), moves.end(), [] (const Move& m) { return m.gain <= 0; }), moves.end());
            std::sort(moves.begin(), moves.end(), std::greater<> {});
            for (const Move& m : moves) {
                const std::array<size_t, 5> involved { m.to, m.from, sibling(m.from), parent_[m.to], parent_[m.from] };
                bool busy = false;
                for (size_t i : involved) busy = busy || locked[i];
                if (busy) continue;
                for (size_t i : involved) locked[i] = true;
                apply(m.from, m.to);
            }
        }
        return true;
    }

private:
    struct Pick {
        size_t node = 0;
        T area = -std::numeric_limits<T>::max();
        bool operator>(const Pick& o) const { return area > o.area; }
    };
    struct Move {
        size_t from = 0, to = 0;
        T gain = (T)0;
        bool operator>(const Move& o) const { return gain > o.gain; }
    };
    struct Box { T lo[kDim], hi[kDim]; };

    std::vector<NodeT>& nodes_;
    std::vector<size_t> parent_;
    size_t threads_;

    static bool is_leaf(const NodeT& n) { return index_count(n.index) != 0; }
    static size_t first_child(const NodeT& n) { return (size_t)index_first(n.index); }
    static size_t sibling(size_t i) { return (i & 1) ? i + 1 : i - 1; }          // bvh.h:34-41
    static size_t left_of_pair(size_t i) { return (i & 1) ? i : i - 1; }         // bvh.h:43-46

    static Box box_of(const NodeT& n) {
        Box b;
        for (int k = 0; k < kDim; ++k) { b.lo[k] = n.bounds[2 * k]; b.hi[k] = n.bounds[2 * k + 1]; }
        return b;
    }
    static void grow(Box& a, const Box& b) {
        for (int k = 0; k < kDim; ++k) { a.lo[k] = robust_min(a.lo[k], b.lo[k]); a.hi[k] = robust_max(a.hi[k], b.hi[k]); }
    }
    static T area_of(const Box& b) {
        if constexpr (kDim == 3) {
            const T d0 = b.hi[0] - b.lo[0], d1 = b.hi[1] - b.lo[1], d2 = b.hi[2] - b.lo[2];
            return (d0 + d1) * d2 + d0 * d1;
        } else {
            return (b.hi[0] - b.lo[0]) + (b.hi[1] - b.lo[1]);
        }
    }
    T area_of(size_t i) const { return area_of(box_of(nodes_[i])); }

    bool link_parents() {
        const size_t count = nodes_.size(), none = ~(size_t)0;
        parent_.assign(count, none);
        parent_[0] = 0;
        for (size_t i = 0; i < count; ++i) {
            if (is_leaf(nodes_[i])) continue;
            const size_t c = first_child(nodes_[i]);
            if (c == 0 || (c & 1) == 0 || c + 1 >= count) return false;          // left child at an odd index, both in range
            if (parent_[c] != none || parent_[c + 1] != none) return false;      // two parents: not a tree
            parent_[c] = i; parent_[c + 1] = i;
        }
        for (size_t i = 0; i < count; ++i) if (parent_[i] == none) return false;  // unreachable node
        // every node has exactly one parent; with count - 1 edges the graph is a tree iff it has no cycle,
        // i.e. iff walking up from any node reaches the root: check by depth-bounded ascent with memoisation
        std::vector<unsigned char> rooted(count, 0);
        rooted[0] = 1;
        for (size_t i = 1; i < count; ++i) {
            size_t j = i, steps = 0;
            while (!rooted[j]) { j = parent_[j]; if (++steps > count) return false; }
            for (j = i; !rooted[j]; j = parent_[j]) rooted[j] = 1;
        }
        return true;
    }

    template <typename F> void parallel_for(size_t n, F body) {
        const size_t workers = std::min(threads_, (n + 63) / 64);
        if (workers <= 1) { for (size_t i = 0; i < n; ++i) body(i); return; }
        std::vector<std::thread> pool;
        pool.reserve(workers);
        for (size_t w = 0; w < workers; ++w)
            pool.emplace_back([=, &body] { for (size_t i = n * w / workers, e = n * (w + 1) / workers; i < e; ++i) body(i); });
        for (auto& t : pool) t.join();
    }

    // The `want` nodes of largest half-area, root excluded, kept in a min-heap while scanning
    // (reinsertion_optimizer.h:88-105).
    std::vector<Pick> largest_nodes(size_t want) const {
        const size_t count = nodes_.size(), head = std::min(count, want + 1);
        std::vector<Pick> heap;
        heap.reserve(head);
        for (size_t i = 1; i < head; ++i) heap.push_back(Pick { i, area_of(i) });
        std::make_heap(heap.begin(), heap.end(), std::greater<> {});
        for (size_t i = head; i < count; ++i) {
            const T a = area_of(i);
            if (heap.front().area < a) {
                std::pop_heap(heap.begin(), heap.end(), std::greater<> {});
                heap.back() = Pick { i, a };
                std::push_heap(heap.begin(), heap.end(), std::greater<> {});
            }
        }
        return heap;
    }

    // Branch and bound for the best new position of `node` (reinsertion_optimizer.h:107-188).  Walking up from
    // the node's parent, `budget` is the area freed so far by taking the node out (its parent disappears, the
    // ancestors shrink to the union of the remaining siblings); every subtree hanging off that path is searched
    // top-down, a position costing the area of (destination U node) plus the growth of the destination's
    // ancestors inside that subtree, which is carried down as a reduced budget.
    Move best_move(size_t node) const {
        Move best;
        best.from = node;
        const Box node_box = box_of(nodes_[node]);
        const T node_area = area_of(node_box);
        const size_t first_parent = parent_[node];
        T budget = area_of(first_parent);
        size_t side = sibling(node), pivot = first_parent;
        Box remaining = box_of(nodes_[side]);                                    // what the ancestors still have to enclose
        std::vector<std::pair<T, size_t>> todo;
        do {
            todo.emplace_back(budget, side);
            while (!todo.empty()) {
                const std::pair<T, size_t> item = todo.back();
                todo.pop_back();
                if (item.first - node_area <= best.gain) continue;               // cannot beat the best even with zero growth
                const NodeT& dst = nodes_[item.second];
                Box merged = box_of(dst);
                grow(merged, node_box);
                const T gain = item.first - area_of(merged);
                if (gain > best.gain) { best.to = item.second; best.gain = gain; }
                if (!is_leaf(dst)) {
                    const T below = gain + area_of(box_of(dst));
                    todo.emplace_back(below, first_child(dst));
                    todo.emplace_back(below, first_child(dst) + 1);
                }
            }
            if (pivot != first_parent) {
                grow(remaining, box_of(nodes_[side]));
                budget += area_of(pivot) - area_of(remaining);
            }
            side = sibling(pivot);
            pivot = parent_[pivot];
        } while (pivot != 0);
        if (best.to == sibling(best.from) || best.to == parent_[best.from]) best = Move {};
        return best;
    }

    // reinsertion_optimizer.h:190-216: the sibling moves up into the parent's place, the freed pair receives
    // the destination (in the sibling's slot) next to the moved node, and the destination's slot becomes their parent.
    void apply(size_t from, size_t to) {
        const size_t sib = sibling(from), par = parent_[from];
        const NodeT sib_node = nodes_[sib], dst_node = nodes_[to];
        nodes_[to].index = make_index<decltype(nodes_[to].index)>((decltype(nodes_[to].index))left_of_pair(from), 0);
        nodes_[sib] = dst_node;
        nodes_[par] = sib_node;
        if (!is_leaf(dst_node)) { parent_[first_child(dst_node)] = sib; parent_[first_child(dst_node) + 1] = sib; }
        if (!is_leaf(sib_node)) { parent_[first_child(sib_node)] = par; parent_[first_child(sib_node) + 1] = par; }
        parent_[sib] = to;
        parent_[from] = to;
        refit_upwards(to);
        refit_upwards(par);
    }

    void refit_upwards(size_t i) {
        do {
            NodeT& n = nodes_[i];
            if (!is_leaf(n)) {
                Box b = box_of(nodes_[first_child(n)]);
                grow(b, box_of(nodes_[first_child(n) + 1]));
                for (int k = 0; k < kDim; ++k) { n.bounds[2 * k] = b.lo[k]; n.bounds[2 * k + 1] = b.hi[k]; }
            }
            i = parent_[i];
        } while (i != 0);
    }
};

} // namespace bvhb200
