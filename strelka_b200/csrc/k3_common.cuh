// k3_common.cuh -- pieces shared by the two GlobalAligner kernels (k3_global_align.cu: one warp per matrix; k3_group.cu: one
// 8-lane group per matrix).  Semantics follow /root/reference/src/c++/lib/alignment/ (AlignerBase.hh:71-92, AlignerUtil.hh:47-80).
#pragma once

#include "sx_internal.h"

namespace k3
{
constexpr unsigned FULL = 0xffffffffu;
constexpr int BAD = -10000; // badVal, GlobalAlignerImpl.hh:58
enum { ST_MATCH = 0, ST_DELETE = 1, ST_INSERT = 2 };
enum { CIG_M = 0, CIG_I = 1, CIG_D = 2, CIG_S = 4, CIG_EQ = 7, CIG_X = 8 };

// first-argument-wins maximum of three with the index of the winner (AlignerBase::max3)
__device__ __forceinline__ uint32_t max3(int& mx, int v0, int v1, int v2)
{
    mx = v0;
    uint32_t p = 0;
    if (v1 > v0)
    {
        mx = v1;
        p = 1;
    }
    if (v2 > mx)
    {
        mx = v2;
        p = 2;
    }
    return p;
}

struct bt_state // BackTrace<int>
{
    int max;
    int state;
    uint32_t queryBegin, refBegin;
    bool isInit;
};
__device__ __forceinline__ void update_bt(bt_state& b, int v, uint32_t refIndex, uint32_t queryIndex, int state)
{
    if (!b.isInit || v > b.max)
    {
        b.max = v;
        b.refBegin = refIndex;
        b.queryBegin = queryIndex;
        b.isInit = true;
        b.state = state;
    }
}

// scores of DP row 0 at matrix column c >= 1 (GlobalAlignerImpl.hh:104-126) and of the initial column at DP row `row` (:69-88)
__device__ __forceinline__ void row0_scores(const sx_ga_scores& sc, int c, int& m, int& d, int& i)
{
    m = sc.isRequireEdgeDeletion ? BAD : 0;
    d = sc.isRequireEdgeDeletion ? sc.open + c * sc.extend : BAD;
    i = BAD;
}
__device__ __forceinline__ void col0_scores(const sx_ga_scores& sc, int row, int& m, int& d, int& i)
{
    m = row * sc.offEdge;
    d = BAD;
    i = sc.isAllowEdgeInsertion ? sc.open + row * sc.extend : BAD;
}
__host__ __device__ __forceinline__ uint32_t pad16u(uint32_t x) { return (x + 15u) & ~15u; }

// per-warp shared-memory slot for a (Q, R) problem
__host__ __device__ __forceinline__ uint32_t k3_slot_bytes(uint32_t Q, uint32_t R)
{
    const uint32_t T = (Q + 31) / 32;
    uint32_t o = 0;
    o += pad16u((Q + 1) * (R + 1)); // pointer matrix
    o += 3u * 32u * T * 4u;         // final-column score strips (and the working strips of the long-query fallback)
    o += pad16u(Q + 8);             // query
    o += pad16u(R + 8);             // ref
    o += pad16u(Q + R + 8);         // traceback steps
    return o;
}

} // namespace k3

// k3_group.cu: classifies the batch, runs every matrix with Q <= 128 and R <= 255 in 8-lane groups, and returns the list of the
// remaining (large) problems for the warp-per-matrix kernel.
int sx_k3_group_run(sx_ctx* ctx, const sx_ga_scores* sc, const sx_ga_batch* d, sx_ga_result* res_dev, uint32_t* cigar_dev, const uint32_t** large_order_dev,
                    uint32_t* n_large, uint32_t* large_slot_need);
