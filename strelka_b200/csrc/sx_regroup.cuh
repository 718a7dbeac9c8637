// sx_regroup.cuh -- a dense work list regrouped by an 8-bit class (block-local counting sort): the two generic kernels.
//
// The thread-per-read kernels of the realignment path (K7 search, K6 score_indels) walk data-dependent loops; reads of one region at
// neighbouring positions differ in how much they have to do, so a warp that takes 32 list entries in read order executes every lane's
// path in turn.  Entries of one class side by side give warps of like work.  Which thread handles which read never matters to the
// results (every output is placed by read index).  The caller computes class[i] for list entry i and the histogram hist[256]; then
//   sx_regroup_scan_kernel     hist[c] -> first slot of class c
//   sx_regroup_scatter_kernel  out[slot] = list[i], slots handed out per block and class (order inside a class is arbitrary)
#pragma once

#include <stdint.h>

namespace
{
constexpr uint32_t SX_RG_CLASSES = 256;
constexpr int SX_RG_THREADS = 256, SX_RG_ITEMS = 4;

__global__ void __launch_bounds__(SX_RG_CLASSES) sx_regroup_scan_kernel(uint32_t* __restrict__ hist) // hist[c] -> first slot of class c (exclusive prefix sum)
{
    __shared__ uint32_t s[SX_RG_CLASSES];
    s[threadIdx.x] = hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t run(0);
        for (uint32_t c = 0; c < SX_RG_CLASSES; ++c)
        {
            const uint32_t x(s[c]);
            s[c] = run;
            run += x;
        }
    }
    __syncthreads();
    hist[threadIdx.x] = s[threadIdx.x];
}

__global__ void __launch_bounds__(SX_RG_THREADS) sx_regroup_scatter_kernel(const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list, const uint8_t* __restrict__ cls,
                                                                          uint32_t* __restrict__ cursor, uint32_t* __restrict__ out)
{
    __shared__ uint32_t s_cnt[SX_RG_CLASSES], s_base[SX_RG_CLASSES];
    const uint32_t n(*n_list);
    const uint32_t chunk(SX_RG_THREADS * SX_RG_ITEMS);
    for (uint32_t base = blockIdx.x * chunk; base < n; base += gridDim.x * chunk) // (block-uniform)
    {
        s_cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t c[SX_RG_ITEMS], at[SX_RG_ITEMS];
#pragma unroll
        for (int j = 0; j < SX_RG_ITEMS; ++j)
        {
            const uint32_t i(base + j * SX_RG_THREADS + threadIdx.x);
            c[j] = i < n ? cls[i] : 0xffffffffu;
            at[j] = c[j] != 0xffffffffu ? atomicAdd(&s_cnt[c[j]], 1u) : 0u;
        }
        __syncthreads();
        s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], s_cnt[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SX_RG_ITEMS; ++j)
        {
            const uint32_t i(base + j * SX_RG_THREADS + threadIdx.x);
            if (c[j] != 0xffffffffu) out[s_base[c[j]] + at[j]] = list[i];
        }
        __syncthreads();
    }
}

// class[i] = min(off[r + 1] - off[r], 255) for list entry i = read r (the reads' CSR offsets of candidate alignments: the loops of K6 / K9 run over
// a read's alignments), + the histogram of the classes
__global__ void __launch_bounds__(SX_RG_THREADS) sx_regroup_class_by_count_kernel(const uint32_t* __restrict__ off, const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list,
                                                                                  uint8_t* __restrict__ cls, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_cnt[SX_RG_CLASSES];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n(*n_list);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t r(list[i]);
        const uint32_t c(min(off[r + 1] - off[r], SX_RG_CLASSES - 1u));
        cls[i] = (uint8_t)c;
        atomicAdd(&s_cnt[c], 1u);
    }
    __syncthreads();
    if (s_cnt[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_cnt[threadIdx.x]);
}
} // namespace
