// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" driver around the UNMODIFIED reference headers, included in place from
// /root/reference/src (-I/root/reference/src; see oracle/Makefile).  Nothing from the reference is
// copied into this repository: this file only *calls* the reference's public API
//   bvh::v2::DefaultBuilder<Node>::build      (default_builder.h:33-62)
//   bvh::v2::Bvh<Node>::intersect             (bvh.h:159-182)
//   bvh::v2::Bvh<Node>::refit / serialize     (bvh.h:210-242)
//   bvh::v2::PrecomputedTri<T>::intersect     (tri.h:55-74)
// the way test/benchmark.cpp:202-298 and test/simple_example.cpp:52-92 do, and flattens the
// results into plain arrays so that python (ctypes) can compare them with the CUDA path.
//
// The result is oracle/_ref/libbvh_ref.so (git-ignored; it travels to the GPU box, the reference
// sources do not).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may load it.
//
// Build flags are pinned by oracle/Makefile: -O3 -DNDEBUG -march=x86-64-v3 -ffp-contract=off.
// -ffp-contract=off makes dot/cross/Moeller-Trumbore unfused (SURVEY.md §7 "hard parts");
// -march=x86-64-v3 defines __FP_FAST_FMAF so fast_mul_add (utils.h:73-81) is a true std::fma in
// the ray/box test, which is what the device code reproduces with __fmaf_rn.

#include <bvh/v2/bvh.h>
#include <bvh/v2/vec.h>
#include <bvh/v2/ray.h>
#include <bvh/v2/node.h>
#include <bvh/v2/default_builder.h>
#include <bvh/v2/thread_pool.h>
#include <bvh/v2/executor.h>
#include <bvh/v2/stack.h>
#include <bvh/v2/tri.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <vector>

namespace {

constexpr uint32_t kInvalid = 0xFFFFFFFFu;

// flags for ref_trace*
constexpr unsigned kAnyHit      = 1u << 0;
constexpr unsigned kRobust      = 1u << 1;
constexpr unsigned kTieLowestId = 1u << 2;   // canonical, tree-independent tie-break
                                             // otherwise: reference example semantics, last visited wins
constexpr unsigned kDynamic     = 1u << 8;   // pool workers claim 4096-ray blocks from a shared counter instead
                                             // of the executor's static n/threads split (bench.py's CPU arm)

template <typename T>
struct Types {
    using Scalar = T;
    using Vec3   = bvh::v2::Vec<T, 3>;
    using BBox   = bvh::v2::BBox<T, 3>;
    using Tri    = bvh::v2::Tri<T, 3>;
    using Node   = bvh::v2::Node<T, 3>;
    using Bvh    = bvh::v2::Bvh<Node>;
    using Ray    = bvh::v2::Ray<T, 3>;
    using PTri   = bvh::v2::PrecomputedTri<T>;
    using Index  = typename Node::Index;
    using IndexT = typename Index::Type;
};

template <typename T>
struct Accel {
    typename Types<T>::Bvh bvh;
    std::vector<typename Types<T>::PTri> tris;   // permuted: tris[i] <- prims[bvh.prim_ids[i]]
};

bvh::v2::ThreadPool* get_pool(int threads) {
    // threads < 0: serial API; 0: hardware_concurrency (thread_pool.h:82-86); >0: that many.
    static bvh::v2::ThreadPool* pool = nullptr;
    static int pool_threads = -2;
    if (threads < 0) return nullptr;
    if (!pool || pool_threads != threads) {
        delete pool;
        pool = new bvh::v2::ThreadPool(static_cast<size_t>(threads));
        pool_threads = threads;
    }
    return pool;
}

template <typename T>
Accel<T>* build(const T* bboxes, const T* centers, size_t n, int quality, int threads,
                size_t min_leaf, size_t max_leaf)
{
    using Ty = Types<T>;
    std::vector<typename Ty::BBox> bb(n);
    std::vector<typename Ty::Vec3> cc(n);
    for (size_t i = 0; i < n; ++i) {
        bb[i] = typename Ty::BBox(
            typename Ty::Vec3(bboxes[6 * i + 0], bboxes[6 * i + 1], bboxes[6 * i + 2]),
            typename Ty::Vec3(bboxes[6 * i + 3], bboxes[6 * i + 4], bboxes[6 * i + 5]));
        cc[i] = typename Ty::Vec3(centers[3 * i + 0], centers[3 * i + 1], centers[3 * i + 2]);
    }
    typename bvh::v2::DefaultBuilder<typename Ty::Node>::Config config;
    config.quality = static_cast<typename bvh::v2::DefaultBuilder<typename Ty::Node>::Quality>(quality);
    if (min_leaf) config.min_leaf_size = min_leaf;
    if (max_leaf) config.max_leaf_size = max_leaf;
    auto accel = new Accel<T>();
    if (auto pool = get_pool(threads))
        accel->bvh = bvh::v2::DefaultBuilder<typename Ty::Node>::build(*pool, bb, cc, config);
    else
        accel->bvh = bvh::v2::DefaultBuilder<typename Ty::Node>::build(bb, cc, config);
    return accel;
}

template <typename T>
void set_triangles(Accel<T>* accel, const T* verts /* n x 9, original order */) {
    using Ty = Types<T>;
    size_t n = accel->bvh.prim_ids.size();
    accel->tris.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const T* v = verts + 9 * accel->bvh.prim_ids[i];
        accel->tris[i] = typename Ty::PTri(
            typename Ty::Vec3(v[0], v[1], v[2]),
            typename Ty::Vec3(v[3], v[4], v[5]),
            typename Ty::Vec3(v[6], v[7], v[8]));
    }
}

template <typename T>
struct Hit { uint32_t id; T t, u, v; };

template <typename T, bool AnyHit, bool Robust, bool TieLowestId>
inline Hit<T> trace_one(const Accel<T>& accel, typename Types<T>::Ray ray, uint32_t* stats) {
    using Ty = Types<T>;
    Hit<T> hit { kInvalid, ray.tmax, 0, 0 };
    bvh::v2::SmallStack<typename Ty::Index, 64> stack;
    accel.bvh.template intersect<AnyHit, Robust>(ray, accel.bvh.get_root().index, stack,
        [&] (size_t begin, size_t end) {
            if (stats) stats[1]++;
            for (size_t i = begin; i < end; ++i) {
                if (stats) stats[2]++;
                if (auto h = accel.tris[i].intersect(ray)) {
                    auto [t, u, v] = *h;
                    uint32_t orig = static_cast<uint32_t>(accel.bvh.prim_ids[i]);
                    if constexpr (TieLowestId) {
                        if (!(t < hit.t || orig < hit.id)) continue;
                    }
                    ray.tmax = t;
                    hit = Hit<T> { orig, t, u, v };
                }
            }
            return hit.id != kInvalid;
        },
        [&] (auto&&, auto&&) { if (stats) stats[0]++; });
    return hit;
}

template <typename T>
using TraceFn = Hit<T> (*)(const Accel<T>&, typename Types<T>::Ray, uint32_t*);

template <typename T>
TraceFn<T> select_trace(unsigned flags) {
    static const TraceFn<T> fns[8] = {
        trace_one<T, false, false, false>, trace_one<T, true, false, false>,
        trace_one<T, false, true,  false>, trace_one<T, true, true,  false>,
        trace_one<T, false, false, true >, trace_one<T, true, false, true >,
        trace_one<T, false, true,  true >, trace_one<T, true, true,  true >,
    };
    return fns[flags & 7u];
}

// rays: m x 8 (org3, dir3, tmin, tmax); outputs may be null.  stats (if non-null): m x 3 u32
// {inner steps, leaves, triangle tests} per ray (InnerFn hook, bvh.h:168; benchmark.cpp:282-296).
template <typename T>
double trace(const Accel<T>* accel, const T* rays, size_t m, unsigned flags, int threads,
             uint32_t* ids, T* ts, T* us, T* vs, uint32_t* stats)
{
    using Ty = Types<T>;
    auto fn = select_trace<T>(flags);
    auto body = [&] (size_t begin, size_t end) {
        for (size_t i = begin; i < end; ++i) {
            const T* r = rays + 8 * i;
            typename Ty::Ray ray(
                typename Ty::Vec3(r[0], r[1], r[2]),
                typename Ty::Vec3(r[3], r[4], r[5]), r[6], r[7]);
            auto hit = fn(*accel, ray, stats ? stats + 3 * i : nullptr);
            if (ids) ids[i] = hit.id;
            if (ts)  ts[i] = hit.t;
            if (us)  us[i] = hit.u;
            if (vs)  vs[i] = hit.v;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (auto pool = get_pool(threads)) {
        // The reference ships no multithreaded ray loop; this is its own executor (executor.h:51-61)
        // applied to rays, the closest faithful "own multithreaded CPU path" (SURVEY §8d).
        bvh::v2::ParallelExecutor executor(*pool);
        if (flags & kDynamic) {
            // Same pool, same executor, but every worker claims blocks of consecutive rays until the batch is
            // drained: a camera image has rows of very different cost, and a static split leaves most threads
            // idle while the slowest chunk finishes (run-to-run spread of 4x on 128-thread hosts).
            constexpr size_t kBlockRays = 4096;
            std::atomic<size_t> next { 0 };
            const size_t workers = pool->get_thread_count();
            bvh::v2::ParallelExecutor one_task_per_worker(*pool, 1);      // (the default threshold of 1024 would run it serially)
            one_task_per_worker.for_each(0, workers, [&] (size_t, size_t) {
                for (;;) {
                    const size_t b = next.fetch_add(kBlockRays, std::memory_order_relaxed);
                    if (b >= m) break;
                    body(b, b + kBlockRays < m ? b + kBlockRays : m);
                }
            });
        } else {
            executor.for_each(0, m, body);
        }
    } else {
        body(0, m);
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

template <typename T>
void get_nodes(const Accel<T>* accel, T* bounds /* count x 6: minx,maxx,miny,maxy,minz,maxz */,
               uint64_t* index_values)
{
    const auto& nodes = accel->bvh.nodes;
    for (size_t i = 0; i < nodes.size(); ++i) {
        for (int k = 0; k < 6; ++k) bounds[6 * i + k] = nodes[i].bounds[k];
        index_values[i] = nodes[i].index.value;
    }
}

template <typename T>
Accel<T>* from_nodes(const T* bounds, const uint64_t* index_values, size_t node_count,
                     const uint64_t* prim_ids, size_t prim_count)
{
    using Ty = Types<T>;
    auto accel = new Accel<T>();
    accel->bvh.nodes.resize(node_count);
    for (size_t i = 0; i < node_count; ++i) {
        for (int k = 0; k < 6; ++k) accel->bvh.nodes[i].bounds[k] = bounds[6 * i + k];
        accel->bvh.nodes[i].index = typename Ty::Index(static_cast<typename Ty::IndexT>(index_values[i]));
    }
    accel->bvh.prim_ids.assign(prim_ids, prim_ids + prim_count);
    return accel;
}

template <typename T>
size_t serialize(const Accel<T>* accel, unsigned char* out, size_t cap) {
    std::ostringstream os(std::ios::binary);
    bvh::v2::StdOutputStream stream(os);
    accel->bvh.serialize(stream);
    auto s = os.str();
    if (out && cap >= s.size()) std::memcpy(out, s.data(), s.size());
    return s.size();
}

template <typename T>
Accel<T>* deserialize(const unsigned char* data, size_t size) {
    std::istringstream is(std::string(reinterpret_cast<const char*>(data), size), std::ios::binary);
    bvh::v2::StdInputStream stream(is);
    auto accel = new Accel<T>();
    accel->bvh = Types<T>::Bvh::deserialize(stream);
    return accel;
}

// Tri::get_bbox / get_center (tri.h:24-25) as callers use them (benchmark.cpp:205-212).
template <typename T>
void tri_bboxes_centers(const T* verts, size_t n, T* bboxes, T* centers) {
    using Ty = Types<T>;
    for (size_t i = 0; i < n; ++i) {
        const T* v = verts + 9 * i;
        typename Ty::Tri tri(
            typename Ty::Vec3(v[0], v[1], v[2]),
            typename Ty::Vec3(v[3], v[4], v[5]),
            typename Ty::Vec3(v[6], v[7], v[8]));
        auto bbox = tri.get_bbox();
        auto c = tri.get_center();
        for (int k = 0; k < 3; ++k) {
            bboxes[6 * i + k] = bbox.min[k];
            bboxes[6 * i + 3 + k] = bbox.max[k];
            centers[3 * i + k] = c[k];
        }
    }
}

} // namespace

#define REF_API(T, S) \
    void* ref_build##S(const T* bboxes, const T* centers, size_t n, int quality, int threads, \
                       size_t min_leaf, size_t max_leaf) { \
        return build<T>(bboxes, centers, n, quality, threads, min_leaf, max_leaf); } \
    double ref_time_build##S(const T* bboxes, const T* centers, size_t n, int quality, int threads) { \
        auto t0 = std::chrono::steady_clock::now(); \
        auto accel = build<T>(bboxes, centers, n, quality, threads, 0, 0); \
        auto t1 = std::chrono::steady_clock::now(); \
        delete accel; \
        return std::chrono::duration<double>(t1 - t0).count(); } \
    void ref_destroy##S(void* h) { delete static_cast<Accel<T>*>(h); } \
    size_t ref_node_count##S(const void* h) { return static_cast<const Accel<T>*>(h)->bvh.nodes.size(); } \
    size_t ref_prim_count##S(const void* h) { return static_cast<const Accel<T>*>(h)->bvh.prim_ids.size(); } \
    void ref_get_nodes##S(const void* h, T* bounds, uint64_t* index_values) { \
        get_nodes<T>(static_cast<const Accel<T>*>(h), bounds, index_values); } \
    void ref_get_prim_ids##S(const void* h, uint64_t* out) { \
        auto& ids = static_cast<const Accel<T>*>(h)->bvh.prim_ids; \
        for (size_t i = 0; i < ids.size(); ++i) out[i] = ids[i]; } \
    void* ref_from_nodes##S(const T* bounds, const uint64_t* index_values, size_t node_count, \
                            const uint64_t* prim_ids, size_t prim_count) { \
        return from_nodes<T>(bounds, index_values, node_count, prim_ids, prim_count); } \
    void ref_set_triangles##S(void* h, const T* verts) { set_triangles<T>(static_cast<Accel<T>*>(h), verts); } \
    double ref_trace##S(const void* h, const T* rays, size_t m, unsigned flags, int threads, \
                        uint32_t* ids, T* ts, T* us, T* vs, uint32_t* stats) { \
        return trace<T>(static_cast<const Accel<T>*>(h), rays, m, flags, threads, ids, ts, us, vs, stats); } \
    void ref_refit##S(void* h) { static_cast<Accel<T>*>(h)->bvh.refit(); } \
    size_t ref_serialize##S(const void* h, unsigned char* out, size_t cap) { \
        return serialize<T>(static_cast<const Accel<T>*>(h), out, cap); } \
    void* ref_deserialize##S(const unsigned char* data, size_t size) { return deserialize<T>(data, size); } \
    void ref_tri_bboxes_centers##S(const T* verts, size_t n, T* bboxes, T* centers) { \
        tri_bboxes_centers<T>(verts, n, bboxes, centers); }

extern "C" {

REF_API(float, 3f)
REF_API(double, 3d)

int ref_thread_count(int threads) {
    auto pool = get_pool(threads);
    return pool ? static_cast<int>(pool->get_thread_count()) : 1;
}

// utils.h:117-120
uint32_t ref_morton_encode32(uint32_t x, uint32_t y, uint32_t z) { return bvh::v2::morton_encode<uint32_t>(x, y, z); }
uint64_t ref_morton_encode64(uint64_t x, uint64_t y, uint64_t z) { return bvh::v2::morton_encode<uint64_t>(x, y, z); }

// 1 iff the ray/box test in this build is a fused multiply-add (utils.h:75-76).
int ref_fast_mul_add_is_fma(void) {
#ifdef FP_FAST_FMAF
    return 1;
#else
    return 0;
#endif
}

} // extern "C"
