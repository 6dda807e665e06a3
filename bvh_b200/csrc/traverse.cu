// bvh_b200/csrc/traverse.cu — batched closest-hit / any-hit ray traversal kernels (sm_100a).
//
// Replaces the caller-side ray loop around Bvh::intersect (reference bvh.h:159-182 driven by
// test/benchmark.cpp:340-393 / c_api/bvh_impl.h:235-250) with two kernels over a whole ray batch:
//
//   trace_persistent_kernel  persistent warps.  Each warp owns a private chunk of consecutive rays
//                            claimed with one global atomicAdd per chunk; lanes whose ray finished are
//                            refilled from the chunk by ballot + prefix rank (active-mask compaction),
//                            so a warp keeps 32 live rays until the batch drains.  The body is a
//                            while-while loop: an inner-node phase (every live lane descends until it
//                            holds a leaf), a reconvergence point, then a leaf phase (Moeller-Trumbore
//                            in registers).  The traversal stack lives in shared memory, laid out
//                            [entry][thread] so that a warp's accesses never bank-conflict.
//   trace_simple_kernel      one thread per ray, same stack machine; also the statistics variant that
//                            counts inner steps / leaves / triangle tests per ray (the reference's
//                            InnerFn hook, bvh.h:168) which defines the algorithmic bytes of DESIGN.md.
//
// Both kernels execute the reference's per-ray algorithm exactly (traverse_core.cuh), so on the same
// tree they visit nodes in the same order as the CPU code and produce bit-identical ids, t, u, v.
#include "engine.h"
#include "traverse_core.cuh"

namespace bvhb200 {

namespace {

constexpr int kTraceBlock = 128;
constexpr int kChunkRays = 128;          // rays claimed per global atomic by one warp

// Shared-memory stack: entry k of thread t lives at base[k * stride + t] (bank = t mod 32).
template <typename U> struct SmemStack {
    U* base;
    uint32_t stride;
    uint32_t sp;
    __device__ __forceinline__ void push(U v) { base[sp * stride] = v; ++sp; }
    __device__ __forceinline__ U pop() { --sp; return base[sp * stride]; }
    __device__ __forceinline__ bool empty() const { return sp == 0; }
};

__device__ __forceinline__ void load_ray(const DevRay<float>* __restrict__ rays, size_t i, RayCtx<float>& r) {
    const float4* q = reinterpret_cast<const float4*>(rays + i);
    const float4 a = __ldcs(q), b = __ldcs(q + 1);           // streamed once: evict-first
    r.org[0] = a.x; r.org[1] = a.y; r.org[2] = a.z; r.dir[0] = a.w;
    r.dir[1] = b.x; r.dir[2] = b.y; r.tmin = b.z; r.tmax = b.w;
}
__device__ __forceinline__ void load_ray(const DevRay<double>* __restrict__ rays, size_t i, RayCtx<double>& r) {
    const double2* q = reinterpret_cast<const double2*>(rays + i);
    const double2 a = __ldcs(q), b = __ldcs(q + 1), c = __ldcs(q + 2), d = __ldcs(q + 3);
    r.org[0] = a.x; r.org[1] = a.y; r.org[2] = b.x; r.dir[0] = b.y;
    r.dir[1] = c.x; r.dir[2] = c.y; r.tmin = d.x; r.tmax = d.y;
}

// A miss reports id = all ones (BVH_INVALID_PRIM_ID), t = the ray's tmax, u = v = 0.
__device__ __forceinline__ void store_hit(DevHit<float>* __restrict__ hits, size_t i, const HitState<float>& h,
                                          float tmax, const uint32_t* __restrict__ prim_ids) {
    uint4 o;
    const bool was_hit = h.slot != kInvalidId;
    o.x = was_hit ? prim_ids[h.slot] : kInvalidId;
    o.y = __float_as_uint(was_hit ? h.t : tmax);
    o.z = __float_as_uint(was_hit ? h.u : 0.f);
    o.w = __float_as_uint(was_hit ? h.v : 0.f);
    __stcs(reinterpret_cast<uint4*>(hits + i), o);
}
__device__ __forceinline__ void store_hit(DevHit<double>* __restrict__ hits, size_t i, const HitState<double>& h,
                                          double tmax, const uint32_t* __restrict__ prim_ids) {
    const bool was_hit = h.slot != kInvalidId;
    ulonglong2 a, b;
    a.x = was_hit ? (unsigned long long)prim_ids[h.slot] : ~0ull;
    a.y = (unsigned long long)__double_as_longlong(was_hit ? h.t : tmax);
    b.x = (unsigned long long)__double_as_longlong(was_hit ? h.u : 0.0);
    b.y = (unsigned long long)__double_as_longlong(was_hit ? h.v : 0.0);
    ulonglong2* d = reinterpret_cast<ulonglong2*>(hits + i);
    __stcs(d, a);
    __stcs(d + 1, b);
}

template <typename T> struct TraceArgs {
    const DevNode<T>* nodes;
    const DevTri<T>* tris;
    const uint32_t* prim_ids;
    const DevRay<T>* rays;
    DevHit<T>* hits;
    unsigned long long n;
    unsigned long long* next_ray;     // persistent kernel: global ray cursor
    uint32_t* ray_stats;              // statistics variant: n x 3
    uint32_t stack_entries;
    int lowest_id;
};

template <typename T, bool kAny, bool kRobust, bool kStats>
__global__ void __launch_bounds__(kTraceBlock)
trace_simple_kernel(TraceArgs<T> a) {
    using U = typename Real<T>::UInt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned long long i = (unsigned long long)blockIdx.x * kTraceBlock + threadIdx.x;
    if (i >= a.n) return;
    SmemStack<U> stack { reinterpret_cast<U*>(smem_raw) + threadIdx.x, kTraceBlock, 0 };
    RayCtx<T> r;
    load_ray(a.rays, i, r);
    ray_prologue<T, kRobust>(r);
    HitState<T> hit { kInvalidId, r.tmax, (T)0, (T)0 };
    const T tmax_in = r.tmax;
    uint32_t stats[3] = { 0, 0, 0 };
    const U root_index = a.nodes[1].index;
    traverse_ray<T, kAny, kRobust>(a.nodes, a.tris, a.prim_ids, a.lowest_id != 0, root_index, r, hit, stack,
                                   kStats ? stats : nullptr);
    store_hit(a.hits, i, hit, tmax_in, a.prim_ids);
    if (kStats) {
        a.ray_stats[3 * i + 0] = stats[0];
        a.ray_stats[3 * i + 1] = stats[1];
        a.ray_stats[3 * i + 2] = stats[2];
    }
}

template <typename T, bool kAny, bool kRobust>
__global__ void __launch_bounds__(kTraceBlock)
trace_persistent_kernel(TraceArgs<T> a) {
    using U = typename Real<T>::UInt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned kFull = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    SmemStack<U> stack { reinterpret_cast<U*>(smem_raw) + threadIdx.x, kTraceBlock, 0 };
    const bool lowest_id = a.lowest_id != 0;
    const U root_index = a.nodes[1].index;

    // warp-uniform cursor over the warp's private chunk of rays
    unsigned long long chunk_pos = 0, chunk_end = 0;
    bool exhausted = false;

    bool has_ray = false;
    unsigned long long ray_index = 0;
    RayCtx<T> r;
    HitState<T> hit;
    T tmax_in = (T)0;
    U top = 0;

    for (;;) {
        // ---- refill idle lanes from the private chunk (claiming a new chunk when it runs dry) ----
        unsigned idle = __ballot_sync(kFull, !has_ray);
        while (idle != 0u && !exhausted) {
            if (chunk_pos == chunk_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.next_ray, (unsigned long long)kChunkRays);
                base = __shfl_sync(kFull, base, 0);
                if (base >= a.n) { exhausted = true; break; }
                chunk_pos = base;
                chunk_end = base + kChunkRays < a.n ? base + kChunkRays : a.n;
            }
            const unsigned avail = (unsigned)(chunk_end - chunk_pos);
            const unsigned want = __popc(idle);
            const unsigned take = want < avail ? want : avail;
            const unsigned rank = __popc(idle & lt_mask);
            if (!has_ray && rank < take) {
                ray_index = chunk_pos + rank;
                load_ray(a.rays, ray_index, r);
                ray_prologue<T, kRobust>(r);
                tmax_in = r.tmax;
                hit.slot = kInvalidId; hit.t = r.tmax; hit.u = (T)0; hit.v = (T)0;
                top = root_index;
                stack.sp = 0;
                has_ray = true;
            }
            chunk_pos += take;
            idle = __ballot_sync(kFull, !has_ray);
        }
        if (__ballot_sync(kFull, has_ray) == 0u) break;      // nothing in flight and nothing left

        // ---- inner phase: descend until this lane holds a leaf (or its ray is finished) ----------
        if (has_ray) {
            while (index_count(top) == 0) {
                if (!inner_step<T, kAny, kRobust>(a.nodes, r, top, stack)) { has_ray = false; break; }
            }
            if (!has_ray) store_hit(a.hits, ray_index, hit, tmax_in, a.prim_ids);
        }
        __syncwarp();

        // ---- leaf phase ---------------------------------------------------------------------------
        if (has_ray) {
            leaf_step<T>(a.tris, a.prim_ids, lowest_id, top, r, hit, nullptr);
            if ((kAny && hit.slot != kInvalidId) || stack.empty()) {
                store_hit(a.hits, ray_index, hit, tmax_in, a.prim_ids);
                has_ray = false;
            } else {
                top = stack.pop();
            }
        }
        __syncwarp();
    }
}

template <typename KernelT>
int configure_smem(KernelT kernel, size_t smem_bytes) {
    if (smem_bytes > 48 * 1024)
        BVH_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    return 0;
}

template <typename T, bool kAny, bool kRobust>
int launch(const TraceArgs<T>& args, bool simple, bool stats, int device, cudaStream_t stream) {
    using U = typename Real<T>::UInt;
    const size_t smem = (size_t)args.stack_entries * kTraceBlock * sizeof(U);
    if (smem > 200 * 1024) { set_error("trace: tree too deep for the shared-memory stack"); return -1; }
    if (stats || simple) {
        const unsigned long long blocks = (args.n + kTraceBlock - 1) / kTraceBlock;
        if (blocks > 0x7FFFFFFFull) { set_error("trace: batch too large for one launch"); return -1; }
        if (stats) {
            if (configure_smem(trace_simple_kernel<T, kAny, kRobust, true>, smem)) return -1;
            trace_simple_kernel<T, kAny, kRobust, true><<<(unsigned)blocks, kTraceBlock, smem, stream>>>(args);
        } else {
            if (configure_smem(trace_simple_kernel<T, kAny, kRobust, false>, smem)) return -1;
            trace_simple_kernel<T, kAny, kRobust, false><<<(unsigned)blocks, kTraceBlock, smem, stream>>>(args);
        }
    } else {
        auto kernel = trace_persistent_kernel<T, kAny, kRobust>;
        if (configure_smem(kernel, smem)) return -1;
        int sm_count = 148, per_sm = 1;
        BVH_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
        BVH_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTraceBlock, smem));
        if (per_sm < 1) per_sm = 1;
        unsigned long long grid = (unsigned long long)sm_count * per_sm;      // one resident wave
        const unsigned long long max_useful = (args.n + 31) / 32 / (kTraceBlock / 32) + 1;
        if (grid > max_useful) grid = max_useful;
        BVH_CUDA_TRY(cudaMemsetAsync(args.next_ray, 0, sizeof(unsigned long long), stream));
        kernel<<<(unsigned)grid, kTraceBlock, smem, stream>>>(args);
    }
    BVH_CUDA_TRY(cudaGetLastError());
    return 0;
}

} // namespace

template <typename T>
int trace_rays(const DeviceBvh<T>& bvh, const DevRay<T>* d_rays, DevHit<T>* d_hits, size_t n,
               unsigned flags, uint32_t* d_ray_stats, cudaStream_t stream) {
    if (n == 0) return 0;
    if (!bvh.nodes || !bvh.tris) { set_error("trace: the BVH has no triangles attached (bvhNN_set_triangles)"); return -1; }
    TraceArgs<T> args;
    args.nodes = bvh.nodes; args.tris = bvh.tris; args.prim_ids = bvh.prim_ids;
    args.rays = d_rays; args.hits = d_hits; args.n = n;
    args.ray_stats = d_ray_stats;
    args.lowest_id = (flags & kTraceLastVisited) ? 0 : 1;
    uint32_t entries = bvh.depth + 1;
    entries = (entries + 7u) & ~7u;
    if (entries < 16) entries = 16;
    args.stack_entries = entries;
    args.next_ray = nullptr;
    const bool simple = (flags & kTraceSimple) != 0, stats = d_ray_stats != nullptr;
    void* cursor = nullptr;
    if (!simple && !stats) {
        if (device_alloc(&cursor, sizeof(unsigned long long), stream)) return -1;
        args.next_ray = static_cast<unsigned long long*>(cursor);
    }
    int rc;
    const bool any = (flags & kTraceAnyHit) != 0, robust = (flags & kTraceRobust) != 0;
    if (any) rc = robust ? launch<T, true, true>(args, simple, stats, bvh.device, stream)
                         : launch<T, true, false>(args, simple, stats, bvh.device, stream);
    else     rc = robust ? launch<T, false, true>(args, simple, stats, bvh.device, stream)
                         : launch<T, false, false>(args, simple, stats, bvh.device, stream);
    if (cursor) device_free(cursor, stream);
    return rc;
}

template int trace_rays<float>(const DeviceBvh<float>&, const DevRay<float>*, DevHit<float>*, size_t, unsigned, uint32_t*, cudaStream_t);
template int trace_rays<double>(const DeviceBvh<double>&, const DevRay<double>*, DevHit<double>*, size_t, unsigned, uint32_t*, cudaStream_t);

} // namespace bvhb200
