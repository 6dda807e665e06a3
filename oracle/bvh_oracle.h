/* oracle/bvh_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the madmann91/bvh v2 algorithms that sit on the hot path
 * (SURVEY.md §8(a)), used only as the parity checker for the CUDA implementation:
 *   tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The shipped library (bvh_b200/csrc, libbvh_c.so) never links, loads or calls this.
 *
 * Parity status: PINNED — tests/test_oracle_pinning.py checks this restatement against
 *   (1) the reference's own known answers (simple_example, the 44-byte serialize file, the
 *       Cornell-box node/intersection counts) and
 *   (2) the unmodified reference compiled in place (oracle/_ref/libbvh_ref.so) on seeded scenes,
 *       node-array-exact for the serial Low/Medium builders and bit-exact for ids/t/u/v.
 *
 * Every function cites the reference file:line it follows.  Suffix 3f = Node<float,3>
 * (28-byte nodes, 32-bit index), 3d = Node<double,3> (56-byte nodes, 64-bit index).
 */
#ifndef BVH_ORACLE_H
#define BVH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INVALID_ID 0xFFFFFFFFu

/* trace flags */
#define ORC_ANY_HIT        1u   /* Bvh::intersect<IsAnyHit=true>  (bvh.h:153-155,180) */
#define ORC_ROBUST         2u   /* Node::intersect_robust          (node.h:68-77)      */
#define ORC_TIE_LOWEST_ID  4u   /* canonical tie-break: t < best || (t == best && id < best_id) */
                                /* default: reference example semantics, last visited wins      */

/* node.h:18-37 — bounds laid out [minx,maxx,miny,maxy,minz,maxz], then the packed index */
typedef struct { float  bounds[6]; uint32_t index; } orc_node3f;   /* 28 bytes */
typedef struct { double bounds[6]; uint64_t index; } orc_node3d;   /* 56 bytes */

/* tri.h:30-37 */
typedef struct { float  p0[3], e1[3], e2[3], n[3]; } orc_ptri3f;   /* 48 bytes */
typedef struct { double p0[3], e1[3], e2[3], n[3]; } orc_ptri3d;   /* 96 bytes */

/* bvh.h:17-23 */
typedef struct { orc_node3f* nodes; size_t node_count; size_t* prim_ids; size_t prim_count; } orc_bvh3f;
typedef struct { orc_node3d* nodes; size_t node_count; size_t* prim_ids; size_t prim_count; } orc_bvh3d;

typedef struct { uint32_t id; float  t, u, v; } orc_hit3f;
typedef struct { uint32_t id; double t, u, v; } orc_hit3d;

#define ORC_DECLARE(T, S) \
    /* Tri::get_bbox / get_center, tri.h:24-25; bboxes n x 6 (min3,max3), centers n x 3 */ \
    void orc_tri_bboxes_centers##S(const T* verts, size_t n, T* bboxes, T* centers); \
    /* BinnedSahBuilder (binned_sah_builder.h:82-155) on TopDownSahBuilder (top_down_sah_builder.h:74-131): \
       DefaultBuilder serial path for Quality::Low (default_builder.h:54-55) */ \
    orc_bvh##S* orc_build_binned##S(const T* bboxes, const T* centers, size_t n, size_t min_leaf, size_t max_leaf); \
    /* SweepSahBuilder (sweep_sah_builder.h:57-139): serial path for Quality::Medium (default_builder.h:57) */ \
    orc_bvh##S* orc_build_sweep##S(const T* bboxes, const T* centers, size_t n, size_t min_leaf, size_t max_leaf); \
    orc_bvh##S* orc_bvh_from_arrays##S(const T* bounds, const uint64_t* index_values, size_t node_count, \
                                      const uint64_t* prim_ids, size_t prim_count); \
    void orc_bvh_get_arrays##S(const orc_bvh##S*, T* bounds, uint64_t* index_values, uint64_t* prim_ids); \
    size_t orc_bvh_node_count##S(const orc_bvh##S*); \
    size_t orc_bvh_prim_count##S(const orc_bvh##S*); \
    void orc_bvh_free##S(orc_bvh##S*); \
    /* PrecomputedTri ctor, tri.h:35-37, in BVH order: out[i] <- verts[prim_ids[i]] (benchmark.cpp:221-225) */ \
    void orc_precompute_tris##S(const orc_bvh##S*, const T* verts, orc_ptri##S* out); \
    /* Bvh::intersect, bvh.h:159-182 + leaf convention of benchmark.cpp:281-292; stats (nullable): \
       {inner steps, leaves, triangle tests} */ \
    orc_hit##S orc_intersect##S(const orc_bvh##S*, const orc_ptri##S* tris, const T ray[8], unsigned flags, uint32_t stats[3]); \
    void orc_trace##S(const orc_bvh##S*, const orc_ptri##S* tris, const T* rays, size_t m, unsigned flags, \
                      uint32_t* ids, T* ts, T* us, T* vs, uint32_t* stats); \
    /* tree-free O(n*m) closest/any hit with the canonical tie-break; small cases only */ \
    void orc_brute_force##S(const T* verts, size_t n, const T* rays, size_t m, unsigned flags, \
                            uint32_t* ids, T* ts, T* us, T* vs); \
    /* Bvh::refit with no leaf function, bvh.h:184-218 (what bvhNN_refit does, c_api/bvh_impl.h:218-221) */ \
    void orc_refit##S(orc_bvh##S*); \
    /* Bvh::serialize / deserialize, bvh.h:220-242 + node.h:90-102 */ \
    size_t orc_serialize##S(const orc_bvh##S*, unsigned char* out, size_t cap); \
    orc_bvh##S* orc_deserialize##S(const unsigned char* data, size_t size); \
    /* structural invariants I1-I6 of SURVEY.md §8(a) A1; returns 0 when all hold, else a code */ \
    int orc_check_invariants##S(const orc_bvh##S*, size_t max_leaf_size); \
    /* sum over inner nodes of half_area (SAH cost proxy, bbox.h:32-38), for reporting tree quality */ \
    double orc_sah_cost##S(const orc_bvh##S*);

ORC_DECLARE(float, 3f)
ORC_DECLARE(double, 3d)

/* utils.h:103-120 */
uint32_t orc_morton_encode32(uint32_t x, uint32_t y, uint32_t z);
uint64_t orc_morton_encode64(uint64_t x, uint64_t y, uint64_t z);
/* 1 iff fast_mul_add is a fused multiply-add in this build (utils.h:75-76) */
int orc_fast_mul_add_is_fma(void);

#ifdef __cplusplus
}
#endif
#endif
