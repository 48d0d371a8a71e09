// scan.cuh -- device-wide exclusive prefix sum (reduce / scan-tiles / apply), hand-written.
//
// Used for: QNAME run ids, kept-record ranks (stable compaction in reference
// output order), multi-mapper CSR offsets, output byte offsets, coverage
// diff-array -> depth.  HBM-bound: each pass streams the (small) SoA column once.
#pragma once
#include "common.cuh"

namespace msg {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE  = SCAN_BLOCK * SCAN_ITEMS;

template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *s_warp, T &block_total)
{
    // v: per-thread value; returns exclusive prefix within the block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = (lane < SCAN_BLOCK / 32) ? s_warp[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < SCAN_BLOCK / 32; o <<= 1) {
            T u = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += u;
        }
        if (lane < SCAN_BLOCK / 32) s_warp[lane] = winc - w;     // exclusive warp offsets
        if (lane == SCAN_BLOCK / 32 - 1) s_warp[SCAN_BLOCK / 32] = winc;
    }
    __syncthreads();
    block_total = s_warp[SCAN_BLOCK / 32];
    T r = s_warp[wid] + inc - v;
    __syncthreads();
    return r;
}

template <class T, class In>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_reduce_kernel(In in, uint64_t n, T *tile_sums)
{
    __shared__ T s_warp[SCAN_BLOCK / 32 + 1];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    T v = T(0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t i = base + k;
        if (i < n) v += in(i);
    }
    T total;
    block_exclusive_scan<T>(v, s_warp, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile_sums in place; total -> *total
template <class T>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tiles_kernel(T *tile_sums, uint32_t ntiles, T *total)
{
    __shared__ T s_warp[SCAN_BLOCK / 32 + 1];
    T carry = T(0);
    for (uint32_t b = 0; b < ntiles; b += SCAN_BLOCK) {
        uint32_t i = b + threadIdx.x;
        T v = (i < ntiles) ? tile_sums[i] : T(0);
        T bt;
        T ex = block_exclusive_scan<T>(v, s_warp, bt);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += bt;
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// out(i, exclusive_prefix, own_value)
template <class T, class In, class Out>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_apply_kernel(In in, Out out, uint64_t n, const T *tile_sums)
{
    __shared__ T s_warp[SCAN_BLOCK / 32 + 1];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    T vals[SCAN_ITEMS];
    T v = T(0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t i = base + k;
        vals[k] = (i < n) ? in(i) : T(0);
        v += vals[k];
    }
    T total;
    T ex = block_exclusive_scan<T>(v, s_warp, total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t i = base + k;
        if (i < n) out(i, ex, vals[k]);
        ex += vals[k];
    }
}

} // namespace msg
