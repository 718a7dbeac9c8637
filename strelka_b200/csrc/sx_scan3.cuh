// sx_scan3.cuh -- exclusive prefix sums of three uint32 arrays at once (K7's per-read alignment / segment / key counts, K7b's per-region
// segment / insert-byte counts).  Tile = K7_SCAN_THREADS * K7_SCAN_ITEMS elements per block:
//   k7_scan_tiles : in-tile exclusive scan in place, tile totals to sums[3][n_tiles]
//   k7_scan_sums  : (one block) exclusive scan of the tile totals in place, grand totals to totals[3]
//   the caller's finish kernel adds sums[q][tile] to every element of the tile.
#pragma once

#include <stdint.h>

constexpr int K7_SCAN_THREADS = 256;
constexpr int K7_SCAN_ITEMS = 8; // elements per thread

static __global__ void __launch_bounds__(K7_SCAN_THREADS) k7_scan_tiles(const uint32_t n, uint32_t* a0, uint32_t* a1, uint32_t* a2, uint32_t* sums, const uint32_t n_tiles)
{
    __shared__ uint32_t warp_sum[3][K7_SCAN_THREADS / 32];
    uint32_t* arr[3] = {a0, a1, a2};
    const uint32_t base(blockIdx.x * K7_SCAN_THREADS * K7_SCAN_ITEMS + threadIdx.x * K7_SCAN_ITEMS);
    const uint32_t lane(threadIdx.x & 31), warp(threadIdx.x >> 5);
    for (int q = 0; q < 3; ++q)
    {
        uint32_t vals[K7_SCAN_ITEMS];
        uint32_t sum(0);
        for (int i = 0; i < K7_SCAN_ITEMS; ++i)
        {
            vals[i] = (base + i < n) ? arr[q][base + i] : 0u;
            sum += vals[i];
        }
        uint32_t incl(sum);
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t y(__shfl_up_sync(0xffffffffu, incl, d));
            if ((int)lane >= d) incl += y;
        }
        if (lane == 31) warp_sum[q][warp] = incl;
        __syncthreads();
        uint32_t warp_off(0);
        for (uint32_t w = 0; w < warp; ++w) warp_off += warp_sum[q][w];
        uint32_t run(warp_off + incl - sum);
        for (int i = 0; i < K7_SCAN_ITEMS; ++i)
        {
            if (base + i < n) arr[q][base + i] = run;
            run += vals[i];
        }
        if (threadIdx.x == K7_SCAN_THREADS - 1) sums[(size_t)q * n_tiles + blockIdx.x] = run;
    }
}

static __global__ void __launch_bounds__(K7_SCAN_THREADS) k7_scan_sums(uint32_t* sums, const uint32_t n_tiles, uint32_t* __restrict__ totals)
{
    // a single block walks the tile totals in chunks of blockDim.x, carrying the running sum
    __shared__ uint32_t warp_sum[K7_SCAN_THREADS / 32];
    __shared__ uint32_t carry;
    const uint32_t lane(threadIdx.x & 31), warp(threadIdx.x >> 5);
    for (int q = 0; q < 3; ++q)
    {
        uint32_t* s(sums + (size_t)q * n_tiles);
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (uint32_t b0 = 0; b0 < n_tiles; b0 += K7_SCAN_THREADS)
        {
            const uint32_t i(b0 + threadIdx.x);
            const uint32_t x(i < n_tiles ? s[i] : 0u);
            uint32_t incl(x);
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t y(__shfl_up_sync(0xffffffffu, incl, d));
                if ((int)lane >= d) incl += y;
            }
            if (lane == 31) warp_sum[warp] = incl;
            __syncthreads();
            uint32_t warp_off(0);
            for (uint32_t w = 0; w < warp; ++w) warp_off += warp_sum[w];
            const uint32_t c(carry);
            if (i < n_tiles) s[i] = c + warp_off + incl - x;
            __syncthreads();
            if (threadIdx.x == K7_SCAN_THREADS - 1) carry = c + warp_off + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) totals[q] = carry;
        __syncthreads();
    }
}

