// bvh_b200/csrc/radix_sort.cuh — hand-written LSD radix sort of (Morton key, primitive id) pairs.
//
// 8-bit digits, stable, three kernels per pass:
//   rs_tile_hist   per-tile digit histogram, stored digit-major  [digit][tile]
//   rs_scan        exclusive scan of that array: element (d, t) becomes the global output position of
//                  the first key of tile t with digit d (= keys with a smaller digit anywhere + keys
//                  with the same digit in earlier tiles)
//   rs_scatter     re-reads the tile, ranks keys of equal digit stably (warp-level multi-split with
//                  __match_any_sync, then a per-digit prefix across the tile's warps) and scatters
// A tile is 512 threads x 16 keys = 8192 keys, warp-striped so every load instruction of a warp
// reads 32 consecutive keys (128 B / 256 B, fully coalesced).
//
// Algorithmic HBM bytes per key per pass (32-bit key, 32-bit value): hist 4 + scatter read 8 +
// scatter write 8 = 20 (the first pass synthesises the identity permutation instead of reading it).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace bvhb200 {

constexpr int kRsBlock = 512;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsBlock * kRsItems;     // 8192
constexpr int kRsBins = 256;
constexpr int kRsWarps = kRsBlock / 32;          // 16
constexpr int kRsScanBlock = 1024;

template <typename K>
__global__ void __launch_bounds__(kRsBlock)
rs_tile_hist_kernel(const K* __restrict__ keys, uint32_t n, int shift,
                    uint32_t* __restrict__ tile_hist, uint32_t num_tiles) {
    __shared__ uint32_t hist[kRsBins];
    const uint32_t tile = blockIdx.x;
    if (threadIdx.x < kRsBins) hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = tile * (uint32_t)kRsTile;
    #pragma unroll 4
    for (int k = threadIdx.x; k < kRsTile; k += kRsBlock) {
        const uint32_t idx = base + k;
        if (idx < n) atomicAdd(&hist[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kRsBins) tile_hist[(size_t)threadIdx.x * num_tiles + tile] = hist[threadIdx.x];
}

// Single-block exclusive scan of `count` uint32 values, in place.
__global__ void __launch_bounds__(kRsScanBlock)
rs_scan_kernel(uint32_t* __restrict__ data, uint32_t count) {
    __shared__ uint32_t warp_sums[kRsScanBlock / 32];
    const uint32_t per = (count + kRsScanBlock - 1) / kRsScanBlock;
    const uint32_t begin = threadIdx.x * per;
    const uint32_t end = begin + per < count ? begin + per : count;
    uint32_t sum = 0;
    for (uint32_t i = begin; i < end; ++i) sum += data[i];
    // block-wide exclusive scan of `sum`
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
        uint32_t wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= (unsigned)o) wi += v;
        }
        warp_sums[lane] = wi - w;           // exclusive prefix of the warp totals
    }
    __syncthreads();
    uint32_t running = warp_sums[warp] + (incl - sum);
    for (uint32_t i = begin; i < end; ++i) {
        const uint32_t v = data[i];
        data[i] = running;
        running += v;
    }
}

template <typename K, bool kIota>
__global__ void __launch_bounds__(kRsBlock)
rs_scatter_kernel(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  K* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                  const uint32_t* __restrict__ tile_offsets, uint32_t num_tiles) {
    __shared__ uint32_t warp_hist[kRsWarps][kRsBins];     // 16 KB
    __shared__ uint32_t bin_base[kRsBins];
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t tile = blockIdx.x;

    for (int k = tid; k < kRsWarps * kRsBins; k += kRsBlock) (&warp_hist[0][0])[k] = 0;
    __syncthreads();

    const uint32_t warp_base = tile * (uint32_t)kRsTile + warp * (32u * kRsItems);
    K key[kRsItems];
    uint32_t val[kRsItems];
    uint32_t rank[kRsItems];
    #pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t idx = warp_base + j * 32 + lane;
        const bool valid = idx < n;
        key[j] = valid ? keys_in[idx] : (K)0;
        val[j] = kIota ? idx : (valid ? vals_in[idx] : 0u);
    }
    #pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t idx = warp_base + j * 32 + lane;
        const bool valid = idx < n;
        const uint32_t d = valid ? ((uint32_t)(key[j] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t r = __popc(peers & lt);
        uint32_t base = 0;
        if (valid) base = warp_hist[warp][d];
        __syncwarp();
        if (valid && r == 0) warp_hist[warp][d] = base + __popc(peers);
        __syncwarp();
        rank[j] = base + r;
    }
    __syncthreads();
    if (tid < kRsBins) {
        uint32_t running = 0;
        #pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const uint32_t c = warp_hist[w][tid];
            warp_hist[w][tid] = running;
            running += c;
        }
        bin_base[tid] = tile_offsets[(size_t)tid * num_tiles + tile];
    }
    __syncthreads();
    #pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t idx = warp_base + j * 32 + lane;
        if (idx < n) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & 255u;
            const uint32_t pos = bin_base[d] + warp_hist[warp][d] + rank[j];
            keys_out[pos] = key[j];
            vals_out[pos] = val[j];
        }
    }
}

// Sorts n (key, value) pairs by the low `key_bits` bits of the key.  keys_a/vals_a hold the input
// (vals_a is ignored: the value of element i is i) and, because the number of passes is even for
// 30- and 63-bit keys, also the output.  tile_hist needs 256 * ceil(n / 8192) uint32.
template <typename K>
inline cudaError_t radix_sort_pairs(K* keys_a, uint32_t* vals_a, K* keys_b, uint32_t* vals_b,
                                    uint32_t* tile_hist, uint32_t n, int key_bits, cudaStream_t stream) {
    const uint32_t num_tiles = (n + kRsTile - 1) / kRsTile;
    int passes = (key_bits + 7) / 8;
    if (passes & 1) ++passes;                              // keep the result in buffer A
    K* kin = keys_a; K* kout = keys_b;
    uint32_t* vin = vals_a; uint32_t* vout = vals_b;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = pass * 8;
        rs_tile_hist_kernel<K><<<num_tiles, kRsBlock, 0, stream>>>(kin, n, shift, tile_hist, num_tiles);
        rs_scan_kernel<<<1, kRsScanBlock, 0, stream>>>(tile_hist, num_tiles * (uint32_t)kRsBins);
        if (pass == 0)
            rs_scatter_kernel<K, true><<<num_tiles, kRsBlock, 0, stream>>>(kin, vin, kout, vout, n, shift, tile_hist, num_tiles);
        else
            rs_scatter_kernel<K, false><<<num_tiles, kRsBlock, 0, stream>>>(kin, vin, kout, vout, n, shift, tile_hist, num_tiles);
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    return cudaGetLastError();
}

} // namespace bvhb200
