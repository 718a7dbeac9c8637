// k3_global_align.cu -- K3 global_align: affine-gap global alignment of a haplotype to a reference segment, with traceback.
//
// Replaces GlobalAligner<int>::align   (/root/reference/src/c++/lib/alignment/GlobalAlignerImpl.hh:36-228),
//          backTraceAlignment          (alignment/SingleRefAlignerSharedImpl.hh:80-170) and
//          apath_add_seqmatch          (blt_util/align_path_impl.hh:36-86)
// as called from ActiveRegionProcessor::discoverIndelsAndMismatches (starling_common/ActiveRegionProcessor.cpp:591).
//
// One warp per DP matrix.  Lane l owns the strip of T = ceil(Q/32) consecutive query rows; the warp sweeps the matrix as an
// anti-diagonal wavefront (lane l is at reference column t-l at step t).  The three state scores of the row above and of the
// diagonal cross strips through registers and a warp shuffle (3 x shfl.up per step); the lane's own previous-column scores sit
// in a bank-conflict-free shared-memory strip; the 3 x 2-bit back pointers of every cell go to a (Q+1) x (R+1) byte matrix in
// shared memory, so HBM sees only Q+R input bytes and the result.  Scores are int32 and max3 keeps the reference's
// first-argument-wins tie rule (alignment/AlignerBase.hh:71-92), so score, beginPos and CIGAR are bit-exact.
#include "sx_internal.h"

#include <algorithm>

namespace
{
constexpr unsigned FULL = 0xffffffffu;
constexpr int BAD = -10000; // badVal, GlobalAlignerImpl.hh:58
enum { ST_MATCH = 0, ST_DELETE = 1, ST_INSERT = 2 };
enum { CIG_M = 0, CIG_I = 1, CIG_D = 2, CIG_S = 4, CIG_EQ = 7, CIG_X = 8 };

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

struct bt_state // BackTrace<int>, alignment/AlignerUtil.hh:47-80
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

__host__ __device__ __forceinline__ uint32_t k3_smem_bytes(uint32_t Q, uint32_t R)
{
    const uint32_t T = (Q + 31) / 32;
    uint32_t o = 0;
    o += ((Q + 1) * (R + 1) + 15u) & ~15u; // pointer matrix
    o += 3u * 32u * T * 4u;                // score strips
    o += (Q + 15u) & ~15u;                 // query
    o += (R + 15u) & ~15u;                 // ref
    o += (Q + R + 2 + 15u) & ~15u;         // traceback steps
    return o;
}

__global__ void __launch_bounds__(32) k3_global_align_kernel(const char* __restrict__ query_pool, const char* __restrict__ ref_pool,
                                                             const uint32_t* __restrict__ query_off, const uint32_t* __restrict__ ref_off, uint32_t n,
                                                             uint32_t max_ops, sx_ga_scores sc, sx_ga_result* __restrict__ res, uint32_t* __restrict__ cigar,
                                                             uint32_t smem_bytes)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x;
    for (uint32_t prob = blockIdx.x; prob < n; prob += gridDim.x)
    {
        const uint32_t Q = query_off[prob + 1] - query_off[prob];
        const uint32_t R = ref_off[prob + 1] - ref_off[prob];
        if (Q == 0 || R == 0 || k3_smem_bytes(Q, R) > smem_bytes)
        {
            if (lane == 0)
            {
                res[prob].score = 0;
                res[prob].beginPos = 0;
                res[prob].n_ops = 0;
                res[prob].status = 2;
            }
            continue;
        }
        const uint32_t T = (Q + 31) / 32;
        const uint32_t W = R + 1; // pointer matrix row pitch
        uint8_t* ptr = smem;
        uint32_t o = ((Q + 1) * W + 15u) & ~15u;
        int* sM = reinterpret_cast<int*>(smem + o);
        int* sD = sM + 32 * T;
        int* sI = sD + 32 * T;
        o += 3u * 32u * T * 4u;
        char* qs = reinterpret_cast<char*>(smem + o);
        o += (Q + 15u) & ~15u;
        char* rs = reinterpret_cast<char*>(smem + o);
        o += (R + 15u) & ~15u;
        uint8_t* steps = smem + o;

        for (uint32_t i = lane; i < Q; i += 32) qs[i] = query_pool[query_off[prob] + i];
        for (uint32_t i = lane; i < R; i += 32) rs[i] = ref_pool[ref_off[prob] + i];
        // column 0 of the pointer matrix and its scores (GlobalAlignerImpl.hh:69-88)
        for (uint32_t qi = lane; qi <= Q; qi += 32) ptr[qi * W] = static_cast<uint8_t>(ST_MATCH | (ST_MATCH << 2) | ((sc.isAllowEdgeInsertion ? ST_INSERT : ST_MATCH) << 4));
        // row 0 of the pointer matrix, columns 1..R (:104-126)
        for (uint32_t j = lane + 1; j <= R; j += 32) ptr[j] = static_cast<uint8_t>(ST_MATCH | ((sc.isRequireEdgeDeletion ? ST_DELETE : ST_MATCH) << 2) | (ST_MATCH << 4));
        for (uint32_t r = 0; r < T; ++r)
        {
            const uint32_t qi = lane * T + r; // DP row qi+1
            sM[r * 32 + lane] = static_cast<int>(qi + 1) * sc.offEdge;
            sD[r * 32 + lane] = BAD;
            sI[r * 32 + lane] = sc.isAllowEdgeInsertion ? sc.open + static_cast<int>(qi + 1) * sc.extend : BAD;
        }
        __syncwarp();

        // wavefront
        int sendM = 0, sendD = 0, sendI = 0;          // bottom-row scores of the column this lane finished last step
        int prevRecvM = 0, prevRecvD = 0, prevRecvI = 0; // what arrived one step earlier (the diagonal for r = 0)
        const uint32_t last_lane = (Q - 1) / T;        // lane owning DP row Q
        bt_state colbt{0, ST_MATCH, 0, 0, false};      // row-Q match candidates per column (only when !isRequireEdgeDeletion)
        const uint32_t n_steps = R + last_lane;
        for (uint32_t t = 0; t < n_steps; ++t)
        {
            const int recvM = __shfl_up_sync(FULL, sendM, 1);
            const int recvD = __shfl_up_sync(FULL, sendD, 1);
            const int recvI = __shfl_up_sync(FULL, sendI, 1);
            const int j = static_cast<int>(t) - static_cast<int>(lane);
            if (j >= 0 && j < static_cast<int>(R) && lane <= last_lane)
            {
                const char rc = rs[j];
                int upM, upD, upI, dgM, dgD, dgI;
                if (lane == 0)
                {
                    // DP row 0 at matrix column j+1 (:104-126) and at matrix column j
                    upM = sc.isRequireEdgeDeletion ? BAD : 0;
                    upD = sc.isRequireEdgeDeletion ? sc.open + (j + 1) * sc.extend : BAD;
                    upI = BAD;
                    if (j == 0)
                    {
                        dgM = 0;
                        dgD = BAD;
                        dgI = sc.isAllowEdgeInsertion ? sc.open : BAD;
                    }
                    else
                    {
                        dgM = sc.isRequireEdgeDeletion ? BAD : 0;
                        dgD = sc.isRequireEdgeDeletion ? sc.open + j * sc.extend : BAD;
                        dgI = BAD;
                    }
                }
                else
                {
                    upM = recvM;
                    upD = recvD;
                    upI = recvI;
                    if (j == 0)
                    {
                        const int row = static_cast<int>(lane * T); // DP row above this strip, in the initial column
                        dgM = row * sc.offEdge;
                        dgD = BAD;
                        dgI = sc.isAllowEdgeInsertion ? sc.open + row * sc.extend : BAD;
                    }
                    else
                    {
                        dgM = prevRecvM;
                        dgD = prevRecvD;
                        dgI = prevRecvI;
                    }
                }
                int m = 0, d = 0, ins = 0;
                for (uint32_t r = 0; r < T; ++r)
                {
                    const uint32_t qi = lane * T + r;
                    if (qi >= Q) break;
                    const uint32_t idx = r * 32 + lane;
                    const int lfM = sM[idx], lfD = sD[idx], lfI = sI[idx];
                    const uint32_t pm = max3(m, dgM, dgD, dgI);
                    m += (qs[qi] == rc) ? sc.match : sc.mismatch;
                    const uint32_t pd = max3(d, lfM + sc.open, lfD, lfI + sc.insertDelete);
                    d += sc.extend;
                    if (j == 0) d = BAD;
                    const uint32_t pi = max3(ins, upM + sc.open, BAD, upI);
                    ins += sc.extend;
                    if (qi == 0) ins = BAD;
                    ptr[(qi + 1) * W + (j + 1)] = static_cast<uint8_t>(pm | (pd << 2) | (pi << 4));
                    dgM = lfM;
                    dgD = lfD;
                    dgI = lfI;
                    upM = m;
                    upD = d;
                    upI = ins;
                    sM[idx] = m;
                    sD[idx] = d;
                    sI[idx] = ins;
                }
                sendM = m;
                sendD = d;
                sendI = ins;
                if (lane == last_lane && !sc.isRequireEdgeDeletion) update_bt(colbt, m, j + 1, Q, ST_MATCH); // :170-175 (m is row Q's match here)
            }
            prevRecvM = recvM;
            prevRecvD = recvD;
            prevRecvI = recvI;
        }
        __syncwarp();

        // backtrace start selection (:178-209) + traceback + '='/'X' expansion: serial, lane 0
        colbt.max = __shfl_sync(FULL, colbt.max, last_lane);
        colbt.refBegin = __shfl_sync(FULL, colbt.refBegin, last_lane);
        colbt.isInit = __shfl_sync(FULL, colbt.isInit ? 1 : 0, last_lane) != 0;
        if (lane == 0)
        {
            bt_state bt = colbt;
            bt.queryBegin = Q;
            bt.state = ST_MATCH;
            const uint32_t lastIdx = ((Q - 1) % T) * 32 + last_lane;
            if (sc.isRequireEdgeDeletion)
            {
                update_bt(bt, sM[lastIdx], R, Q, ST_MATCH);
                update_bt(bt, sD[lastIdx], R, Q, ST_DELETE);
            }
            if (sc.isAllowEdgeInsertion) update_bt(bt, sI[lastIdx], R, Q, ST_INSERT);
            for (uint32_t queryIndex = 0; queryIndex < Q; ++queryIndex)
            {
                int mval;
                if (queryIndex == 0) mval = sc.isRequireEdgeDeletion ? BAD : 0; // DP row 0 at the last column
                else
                {
                    const uint32_t qi = queryIndex - 1;
                    mval = sM[(qi % T) * 32 + (qi / T)];
                }
                update_bt(bt, mval + static_cast<int>(Q - queryIndex) * sc.offEdge, R, queryIndex, ST_MATCH);
            }

            // traceback: one step code per move, written in reverse
            uint32_t nsteps = 0;
            uint32_t qb = bt.queryBegin, rb = bt.refBegin;
            int state = bt.state;
            const uint32_t trailing_clip = (qb < Q) ? (Q - qb) : 0;
            while (true)
            {
                const uint8_t pv = ptr[qb * W + rb];
                const int next = (pv >> (2 * state)) & 3;
                if (state == ST_MATCH)
                {
                    if (qb < 1 || rb < 1) break;
                    steps[nsteps++] = CIG_M;
                    --qb;
                    --rb;
                }
                else if (state == ST_DELETE)
                {
                    if (rb < 1) break;
                    steps[nsteps++] = CIG_D;
                    --rb;
                }
                else
                {
                    if (qb < 1) break;
                    steps[nsteps++] = CIG_I;
                    --qb;
                }
                state = next;
            }
            const uint32_t leading_clip = qb;
            const uint32_t beginPos = rb;

            // forward emission with run merging; M is split into '=' / 'X' ('N' on either side is a mismatch)
            uint32_t* cg = cigar + static_cast<size_t>(prob) * max_ops;
            uint32_t n_ops = 0;
            int cur_type = -1;
            uint32_t cur_len = 0;
            auto flush = [&]() {
                if (cur_type >= 0)
                {
                    if (n_ops < max_ops) cg[n_ops] = (cur_len << 4) | static_cast<uint32_t>(cur_type);
                    ++n_ops;
                }
            };
            auto push = [&](int type, uint32_t len, bool mergeable) {
                if (mergeable && type == cur_type)
                {
                    cur_len += len;
                    return;
                }
                flush();
                cur_type = type;
                cur_len = len;
            };
            if (leading_clip) push(CIG_S, leading_clip, false);
            uint32_t qi = leading_clip, rix = beginPos;
            // raw segments (before seqmatch) are maximal runs of one step code; the reference never merges across raw segments of
            // different type, and '='/'X' runs are merged by apath_append within and across raw M segments only if adjacent
            for (uint32_t k = nsteps; k-- > 0;)
            {
                const int s = steps[k];
                if (s == CIG_M)
                {
                    bool eq = (qs[qi] == rs[rix]);
                    if (qs[qi] == 'N' || rs[rix] == 'N') eq = false;
                    push(eq ? CIG_EQ : CIG_X, 1, true);
                    ++qi;
                    ++rix;
                }
                else if (s == CIG_D)
                {
                    push(CIG_D, 1, true);
                    ++rix;
                }
                else
                {
                    push(CIG_I, 1, true);
                    ++qi;
                }
            }
            if (trailing_clip)
            {
                // a trailing soft clip is its own segment even if the path is otherwise empty
                flush();
                cur_type = CIG_S;
                cur_len = trailing_clip;
            }
            flush();
            res[prob].score = bt.max;
            res[prob].beginPos = static_cast<int>(beginPos);
            res[prob].n_ops = n_ops;
            res[prob].status = n_ops > max_ops ? 1u : 0u;
        }
        __syncwarp();
    }
}

__global__ void k3_smem_need_kernel(const uint32_t* __restrict__ query_off, const uint32_t* __restrict__ ref_off, uint32_t n, uint32_t* __restrict__ out)
{
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        m = max(m, k3_smem_bytes(query_off[i + 1] - query_off[i], ref_off[i + 1] - ref_off[i]));
    for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(FULL, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

int k3_run(sx_ctx* ctx, const sx_ga_scores* sc, const sx_ga_batch* d, sx_ga_result* res_dev, uint32_t* cigar_dev, uint32_t need)
{
    size_t smem = std::min<size_t>(need, ctx->smem_optin); // problems that do not fit report status 2
    if (smem > 48 * 1024) SX_CUDA(ctx, cudaFuncSetAttribute(k3_global_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx->smem_optin)));
    const int grid = static_cast<int>(std::min<uint32_t>(d->n, (uint32_t)ctx->sm_count * 32));
    k3_global_align_kernel<<<grid, 32, smem, ctx->s_compute>>>(d->query, d->ref, d->query_off, d->ref_off, d->n, d->max_ops, *sc, res_dev, cigar_dev, (uint32_t)smem);
    SX_CUDA(ctx, cudaGetLastError());
    return SX_OK;
}
} // namespace

extern "C" int sx_global_align_dev(sx_ctx* ctx, const sx_ga_scores* sc, const sx_ga_batch* d, sx_ga_result* res_dev, uint32_t* cigar_dev)
{
    if (!ctx) return SX_ERR_ARG;
    ctx->timing = sx_timing{};
    if (!sc || !d || !res_dev || !cigar_dev) return sx_fail(ctx, SX_ERR_ARG, "sx_global_align_dev: NULL argument");
    if (d->n == 0) return SX_OK;
    SX_CUDA(ctx, cudaSetDevice(ctx->device));
    sx_kernel_timer t(ctx);
    uint32_t* dneed = nullptr;
    int rc = sx_ensure(ctx, 20, sizeof(uint32_t), reinterpret_cast<void**>(&dneed));
    if (rc) return rc;
    SX_CUDA(ctx, cudaMemsetAsync(dneed, 0, 4, ctx->s_compute));
    k3_smem_need_kernel<<<std::min<uint32_t>((d->n + 255) / 256, 1184), 256, 0, ctx->s_compute>>>(d->query_off, d->ref_off, d->n, dneed);
    uint32_t need = 0;
    SX_CUDA(ctx, cudaMemcpyAsync(&need, dneed, 4, cudaMemcpyDeviceToHost, ctx->s_compute));
    SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    if ((rc = k3_run(ctx, sc, d, res_dev, cigar_dev, need))) return rc;
    t.stop(2);
    return t.finish();
}

extern "C" int sx_global_align(sx_ctx* ctx, const sx_ga_scores* sc, const sx_ga_batch* b, sx_ga_result* res_host, uint32_t* cigar_host)
{
    if (!ctx) return SX_ERR_ARG;
    ctx->timing = sx_timing{};
    if (!sc || !b || !res_host || !cigar_host || !b->query_off || !b->ref_off) return sx_fail(ctx, SX_ERR_ARG, "sx_global_align: NULL argument");
    if (b->n == 0) return SX_OK;
    SX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t need = 0;
    for (uint32_t i = 0; i < b->n; ++i)
    {
        const uint32_t Q = b->query_off[i + 1] - b->query_off[i], R = b->ref_off[i + 1] - b->ref_off[i];
        if (Q == 0 || R == 0) return sx_fail(ctx, SX_ERR_ARG, "sx_global_align: empty query or reference in problem %u (asserted at GlobalAlignerImpl.hh:47-48)", i);
        need = std::max(need, k3_smem_bytes(Q, R));
    }
    SX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->s_compute));
    sx_ga_batch d = *b;
    void* p = nullptr;
    int rc;
    const size_t qbytes = b->query_off[b->n], rbytes = b->ref_off[b->n];
#define SX_UP(slot, field, type, bytes)                                                    \
    if ((rc = sx_ensure(ctx, slot, (bytes) + 16, &p))) return rc;                            \
    SX_CUDA(ctx, cudaMemcpyAsync(p, b->field, (bytes), cudaMemcpyHostToDevice, ctx->s_compute)); \
    d.field = static_cast<type>(p);
    SX_UP(9, query, const char*, qbytes)
    SX_UP(10, ref, const char*, rbytes)
    SX_UP(11, query_off, const uint32_t*, (size_t)(b->n + 1) * 4)
    SX_UP(12, ref_off, const uint32_t*, (size_t)(b->n + 1) * 4)
#undef SX_UP
    sx_ga_result* d_res = nullptr;
    uint32_t* d_cig = nullptr;
    if ((rc = sx_ensure(ctx, 13, (size_t)b->n * sizeof(sx_ga_result), reinterpret_cast<void**>(&d_res)))) return rc;
    if ((rc = sx_ensure(ctx, 14, (size_t)b->n * b->max_ops * 4 + 16, reinterpret_cast<void**>(&d_cig)))) return rc;
    SX_CUDA(ctx, cudaMemsetAsync(d_cig, 0, (size_t)b->n * b->max_ops * 4, ctx->s_compute)); // unused cigar slots read as 0
    if ((rc = k3_run(ctx, sc, &d, d_res, d_cig, need))) return rc;
    SX_CUDA(ctx, cudaMemcpyAsync(res_host, d_res, (size_t)b->n * sizeof(sx_ga_result), cudaMemcpyDeviceToHost, ctx->s_compute));
    SX_CUDA(ctx, cudaMemcpyAsync(cigar_host, d_cig, (size_t)b->n * b->max_ops * 4, cudaMemcpyDeviceToHost, ctx->s_compute));
    SX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->s_compute));
    SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
    ctx->timing.kernel_ms = ms;
    ctx->timing.launches = 1;
    ctx->total_launches += 1;
    return SX_OK;
}
