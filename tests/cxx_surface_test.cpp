// tests/cxx_surface_test.cpp — host-only check of include/bvh/v2/b200_surface.h (no GPU needed).
// Reads a tree + triangles + rays + expected hits dumped by tests/test_cxx_surface.py from the golden
// fixtures of the unmodified reference, runs Bvh::intersect exactly as the reference's callers do
// (test/benchmark.cpp:277-298), plus refit / serialize / extract_bvh, and reports mismatches.
#include <bvh/v2/bvh.h>
#include <bvh/v2/stack.h>
#include <bvh/v2/tri.h>
#include <bvh/v2/executor.h>
#include <bvh/v2/thread_pool.h>

#include <cstdio>
#include <fstream>
#include <sstream>
#include <vector>

template <typename T>
static int run(std::ifstream& in) {
    using Node = bvh::v2::Node<T, 3>;
    using Bvh = bvh::v2::Bvh<Node>;
    using Vec3 = bvh::v2::Vec<T, 3>;
    using Ray = bvh::v2::Ray<T, 3>;
    using PTri = bvh::v2::PrecomputedTri<T>;
    static_assert(sizeof(Node) == (sizeof(T) == 4 ? 28 : 56));
    uint64_t node_count, prim_count, ray_count;
    in.read((char*)&node_count, 8); in.read((char*)&prim_count, 8); in.read((char*)&ray_count, 8);
    Bvh bvh;
    bvh.nodes.resize(node_count);
    in.read((char*)bvh.nodes.data(), node_count * sizeof(Node));
    std::vector<uint64_t> ids(prim_count);
    in.read((char*)ids.data(), prim_count * 8);
    bvh.prim_ids.assign(ids.begin(), ids.end());
    std::vector<T> verts(9 * prim_count), rays(8 * ray_count);
    in.read((char*)verts.data(), verts.size() * sizeof(T));
    in.read((char*)rays.data(), rays.size() * sizeof(T));
    int failures = 0;
    std::vector<PTri> tris(prim_count);
    for (size_t i = 0; i < prim_count; ++i) {
        const T* v = &verts[9 * bvh.prim_ids[i]];
        tris[i] = PTri(Vec3(v[0], v[1], v[2]), Vec3(v[3], v[4], v[5]), Vec3(v[6], v[7], v[8]));
    }
    for (int mode = 0; mode < 3; ++mode) {          // 0: closest (last visited), 1: any, 2: robust closest
        std::vector<uint32_t> want_ids(ray_count);
        std::vector<T> want_t(ray_count);
        in.read((char*)want_ids.data(), ray_count * 4);
        in.read((char*)want_t.data(), ray_count * sizeof(T));
        bvh::v2::ThreadPool pool(2);
        bvh::v2::ParallelExecutor executor(pool);
        std::vector<int> bad(ray_count, 0);
        executor.for_each(0, ray_count, [&](size_t begin, size_t end) {
            for (size_t i = begin; i < end; ++i) {
                const T* r = &rays[8 * i];
                Ray ray(Vec3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5]), r[6], r[7]);
                size_t prim = SIZE_MAX;
                bvh::v2::SmallStack<typename Bvh::Index, 64> stack;
                auto leaf = [&](size_t b, size_t e) {
                    for (size_t k = b; k < e; ++k)
                        if (auto hit = tris[k].intersect(ray)) { ray.tmax = std::get<0>(*hit); prim = bvh.prim_ids[k]; }
                    return prim != SIZE_MAX;
                };
                if (mode == 0) bvh.template intersect<false, false>(ray, bvh.get_root().index, stack, leaf);
                if (mode == 1) bvh.template intersect<true, false>(ray, bvh.get_root().index, stack, leaf);
                if (mode == 2) bvh.template intersect<false, true>(ray, bvh.get_root().index, stack, leaf);
                const uint32_t got = prim == SIZE_MAX ? 0xFFFFFFFFu : (uint32_t)prim;
                if (got != want_ids[i] || std::memcmp(&ray.tmax, &want_t[i], sizeof(T)) != 0) bad[i] = 1;
            }
        });
        int n = 0; for (int b : bad) n += b;
        std::printf("mode %d mismatches %d\n", mode, n);
        failures += n;
    }
    // refit leaves a consistent tree unchanged; serialize/deserialize round-trips; extract_bvh keeps leaves
    Bvh copy = Bvh::deserialize(*[&] { static std::stringstream ss; bvh::v2::StdOutputStream os(ss); bvh.serialize(os);
                                      static bvh::v2::StdInputStream is(ss); return &is; }());
    if (!(copy == bvh)) { std::puts("serialize round trip differs"); ++failures; }
    copy.refit();
    if (!(copy == bvh)) { std::puts("refit changed a consistent tree"); ++failures; }
    if (node_count > 1) {
        Bvh sub = bvh.extract_bvh(1);
        size_t leaves = 0; for (auto& n : sub.nodes) if (n.is_leaf()) leaves += n.index.prim_count();
        if (leaves != sub.prim_ids.size()) { std::puts("extract_bvh lost primitives"); ++failures; }
    }
    return failures;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    if (!in) return 2;
    const int failures = argv[2][0] == 'd' ? run<double>(in) : run<float>(in);
    std::printf("failures %d\n", failures);
    return failures ? 1 : 0;
}
