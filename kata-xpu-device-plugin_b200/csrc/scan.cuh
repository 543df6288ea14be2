// scan.cuh -- exclusive prefix sum over uint32 (reduce / scan-of-partials / downsweep).
// Used where output sizes are data dependent (name gather, CDI fragments, busIndex).
#pragma once
#include "common.cuh"

namespace kxscan {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;                        // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 items per block

__device__ __forceinline__ uint32_t warp_incl(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, v, d);
        if (kx_lane() >= (uint32_t)d) v += y;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *tot
__device__ __forceinline__ uint32_t block_excl(uint32_t v, uint32_t *tot) {
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t total;
    uint32_t incl = warp_incl(v);
    uint32_t w = threadIdx.x >> 5;
    if (kx_lane() == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t x = kx_lane() < SCAN_THREADS / 32 ? wsum[kx_lane()] : 0;
        uint32_t xi = warp_incl(x);
        if (kx_lane() < SCAN_THREADS / 32) wsum[kx_lane()] = xi - x;
        if (kx_lane() == SCAN_THREADS / 32 - 1) total = xi;
    }
    __syncthreads();
    uint32_t r = wsum[w] + incl - v;
    *tot = total;
    __syncthreads();
    return r;
}

static __global__ void __launch_bounds__(SCAN_THREADS) reduce_kernel(const uint32_t *__restrict__ in, size_t n,
                                                              unsigned long long *__restrict__ partial) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    block_excl(s, &tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// single block: exclusive scan of the per-block partials (64-bit), writes grand total.
static __global__ void __launch_bounds__(1024) partial_scan_kernel(unsigned long long *partial, size_t nb,
                                                             unsigned long long *total_out) {
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t b0 = 0; b0 < nb; b0 += 1024) {
        size_t i = b0 + threadIdx.x;
        unsigned long long v = i < nb ? partial[i] : 0, incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
            if (kx_lane() >= (uint32_t)d) incl += y;
        }
        if (kx_lane() == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long x = wsum[threadIdx.x], xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long y = __shfl_up_sync(0xffffffffu, xi, d);
                if (kx_lane() >= (uint32_t)d) xi += y;
            }
            wsum[threadIdx.x] = xi - x;
        }
        __syncthreads();
        unsigned long long excl = carry + wsum[threadIdx.x >> 5] + incl - v;
        if (i < nb) partial[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// out[i] = exclusive prefix (64-bit capable through partial, stored as OutT)
template <typename OutT>
__global__ void __launch_bounds__(SCAN_THREADS) downsweep_kernel(const uint32_t *__restrict__ in, size_t n,
                                                                 const unsigned long long *__restrict__ partial,
                                                                 OutT *__restrict__ out) {
    // blocked arrangement so that each thread owns SCAN_ITEMS consecutive items
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + k;
        v[k] = i < n ? in[i] : 0;
        s += v[k];
    }
    uint32_t tot;
    uint32_t ex = block_excl(s, &tot);
    unsigned long long run = partial[blockIdx.x] + ex;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + k;
        if (i < n) out[i] = (OutT)run;
        run += v[k];
    }
}

static inline size_t scratch_items(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

// d_partial: scratch of scratch_items(n) u64; d_total: optional u64 (device)
template <typename OutT>
static inline void exclusive_scan(kxpu_ctx *ctx, const uint32_t *d_in, size_t n, OutT *d_out,
                                  unsigned long long *d_partial, unsigned long long *d_total) {
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 0) nb = 1;
    reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, d_partial);
    partial_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_partial, nb, d_total);
    downsweep_kernel<OutT><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, d_partial, d_out);
    ctx->launches += 3;
}

}  // namespace kxscan
