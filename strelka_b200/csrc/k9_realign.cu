// k9_realign.cu -- K9 choose_realignment: from the scored candidate alignments to rseg.realignment, one read per thread.
//
// Replaces (include/strelka_b200.h, "K9 choose_realignment") the tail of scoreCandidateAlignments
// (starling_common/starling_read_align.cpp:1573-1741) with finishRealignment :1411-1450 and starling_read_align_clipper.cpp.  Per-read
// body: k9_core.cuh.
//
// Shape of the work: per read two or three passes over its few alignments' segments (arg-max, pool preference, conflict marking) and
// a per-read-position map of at most 1024 entries kept in local memory (interleaved per lane by the hardware).  Reads in the scores
// where K1 wrote them and the alignments where K7 wrote them; writes one (pos, path) per read where K4 reads a read's best alignment.
// HBM-bound by construction (~40 B per alignment in, ~30 B per read out).  Launches: slots per read -> scan -> choose.

#include "k9_core.cuh"
#include "sx_internal.h"
#include "sx_scan3.cuh"

#include <algorithm>

namespace
{
constexpr int K9_CAP_BIT = 1 << 19;

__global__ void k9_slots_kernel(const sx_realign_batch b, uint32_t* __restrict__ cnt, uint32_t* __restrict__ z1, uint32_t* __restrict__ z2, uint32_t* __restrict__ read_region)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n_reads; r += gridDim.x * blockDim.x)
    {
        cnt[r] = k9_slots(b, r);
        z1[r] = z2[r] = 0;
    }
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < b.n_regions; g += gridDim.x * blockDim.x)
        for (uint32_t r = b.region_read_off[g]; r < b.region_read_off[g + 1]; ++r) read_region[r] = g;
}

__global__ void __launch_bounds__(K7_SCAN_THREADS) k9_finish_kernel(const uint32_t n, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ sums, const uint32_t* __restrict__ totals,
                                                                   const sx_realign_out o, int* __restrict__ status)
{
    const uint32_t tile(blockIdx.x), base(tile * K7_SCAN_THREADS * K7_SCAN_ITEMS + threadIdx.x * K7_SCAN_ITEMS);
    const uint32_t off(sums[tile]);
    for (int i = 0; i < K7_SCAN_ITEMS; ++i)
        if (base + i < n)
        {
            const uint32_t x(cnt[base + i] + off);
            cnt[base + i] = x;
            o.seg_off[base + i] = x;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        o.seg_off[n] = totals[0];
        o.totals[0] = totals[0];
        if (totals[0] > o.cap_segs) atomicOr(status, K9_CAP_BIT);
    }
}

__global__ void __launch_bounds__(128) k9_choose_kernel(const k9_view v, const uint32_t* __restrict__ read_region, const uint32_t* __restrict__ off, const uint32_t* __restrict__ totals,
                                                        const sx_realign_out o)
{
    if (totals[0] > o.cap_segs) return;
    uint8_t type[K9_MAX_READ];
    int32_t pos[K9_MAX_READ];
    k9_scratch S = {type, pos};
    const uint32_t n(v.b.n_reads);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    {
        const uint32_t s0(off[r]), s1(r + 1 < n ? off[r + 1] : totals[0]);
        int32_t p;
        uint16_t ns;
        uint32_t best;
        const uint32_t st(k9_read(v, read_region[r], r, S, o.segs + s0, s1 - s0, p, ns, best));
        if (!(st & SX_REALIGN_ST_REALIGNED)) k9_fallback(v.b, r, o.segs + s0, s1 - s0, p, ns);
        o.pos[r] = p;
        o.n_seg[r] = ns;
        o.status[r] = (uint8_t)st;
        o.best_aln[r] = best;
    }
}

int k9_run(sx_ctx* ctx, const sx_realign_batch* d, const double* lnp, const sx_realign_out* o, unsigned* launches)
{
    cudaStream_t st(ctx->s_compute);
    const uint32_t n(d->n_reads);
    int rc;
    uint32_t *cnt(nullptr), *z1(nullptr), *z2(nullptr), *read_region(nullptr), *sums(nullptr);
    if ((rc = sx_ensure(ctx, 58, (size_t)n * 4 + 16, reinterpret_cast<void**>(&read_region)))) return rc;
    if ((rc = sx_ensure(ctx, 60, (size_t)n * 4 + 16, reinterpret_cast<void**>(&cnt)))) return rc;
    if ((rc = sx_ensure(ctx, 61, (size_t)n * 4 + 16, reinterpret_cast<void**>(&z1)))) return rc;
    if ((rc = sx_ensure(ctx, 62, (size_t)n * 4 + 16, reinterpret_cast<void**>(&z2)))) return rc;
    const uint32_t tile(K7_SCAN_THREADS * K7_SCAN_ITEMS), n_tiles((n + tile - 1) / tile);
    if ((rc = sx_ensure(ctx, 63, ((size_t)3 * n_tiles + 4) * 4, reinterpret_cast<void**>(&sums)))) return rc;
    uint32_t* totals(sums + (size_t)3 * n_tiles);
    const int cap(ctx->sm_count * 16);
    const auto grid = [cap](const uint32_t m) { return (unsigned)std::max(1, std::min<int>((int)((m + 127) / 128), cap)); };
    k9_slots_kernel<<<grid(std::max(n, d->n_regions)), 128, 0, st>>>(*d, cnt, z1, z2, read_region);
    SX_CUDA(ctx, cudaGetLastError());
    k7_scan_tiles<<<n_tiles, K7_SCAN_THREADS, 0, st>>>(n, cnt, z1, z2, sums, n_tiles);
    SX_CUDA(ctx, cudaGetLastError());
    k7_scan_sums<<<1, K7_SCAN_THREADS, 0, st>>>(sums, n_tiles, totals);
    SX_CUDA(ctx, cudaGetLastError());
    k9_finish_kernel<<<n_tiles, K7_SCAN_THREADS, 0, st>>>(n, cnt, sums, totals, *o, ctx->d_status);
    SX_CUDA(ctx, cudaGetLastError());
    k9_view v;
    v.b = *d;
    v.lnp = lnp;
    k9_choose_kernel<<<grid(n), 128, 0, st>>>(v, read_region, cnt, totals, *o);
    SX_CUDA(ctx, cudaGetLastError());
    *launches = 5;
    return SX_OK;
}

int k9_finish(sx_ctx* ctx, const char* what, const uint32_t* totals_host)
{
    int st(0);
    SX_CUDA(ctx, cudaMemcpyAsync(&st, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_compute));
    SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    if (st & K9_CAP_BIT)
    {
        cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->s_compute);
        if (totals_host) return sx_fail(ctx, SX_ERR_CAPACITY, "%s: cap_segs too small: %u segment slots needed", what, totals_host[0]);
        return sx_fail(ctx, SX_ERR_CAPACITY, "%s: cap_segs too small (totals[0] holds the needed size)", what);
    }
    return sx_check_status(ctx, what);
}

int k9_check_args(sx_ctx* ctx, const sx_realign_batch* b, const double* lnp, const sx_realign_out* o, const char* what)
{
    if (!b || !o) return sx_fail(ctx, SX_ERR_ARG, "%s: NULL argument", what);
    if (!o->totals || !o->seg_off) return sx_fail(ctx, SX_ERR_ARG, "%s: NULL output array", what);
    if (b->n_reads == 0) return SX_OK;
    if (!b->region_read_off || !b->region_key_off || !b->aln_off || !b->aln_pos || !b->aln_seg_off || !b->segs || !b->aln_key_off || !b->aln_keys || !b->read_len ||
        (b->n_alns && !lnp) || !o->pos || !o->n_seg || !o->status || !o->best_aln || !o->segs)
        return sx_fail(ctx, SX_ERR_ARG, "%s: NULL array", what);
    if (b->n_regions == 0) return sx_fail(ctx, SX_ERR_ARG, "%s: reads without a region", what);
    return SX_OK;
}
} // namespace

extern "C" int sx_choose_realignment_dev(sx_ctx* ctx, const sx_realign_batch* d, const double* lnp_dev, sx_realign_out* out_dev)
{
    if (!ctx) return SX_ERR_ARG;
    ctx->timing = sx_timing{};
    int rc;
    if ((rc = k9_check_args(ctx, d, lnp_dev, out_dev, "sx_choose_realignment_dev"))) return rc;
    SX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d->n_reads == 0)
    {
        SX_CUDA(ctx, cudaMemsetAsync(out_dev->totals, 0, 4, ctx->s_compute));
        SX_CUDA(ctx, cudaMemsetAsync(out_dev->seg_off, 0, 4, ctx->s_compute));
        SX_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
        return SX_OK;
    }
    sx_kernel_timer t(ctx);
    unsigned launches(0);
    if ((rc = k9_run(ctx, d, lnp_dev, out_dev, &launches))) return rc;
    t.stop(launches);
    if ((rc = t.finish())) return rc;
    return k9_finish(ctx, "sx_choose_realignment", nullptr);
}

extern "C" int sx_choose_realignment(sx_ctx* ctx, const sx_realign_batch* b, const double* lnp_host, sx_realign_out* out_host)
{
    if (!ctx) return SX_ERR_ARG;
    ctx->timing = sx_timing{};
    int rc;
    if ((rc = k9_check_args(ctx, b, lnp_host, out_host, "sx_choose_realignment"))) return rc;
    if (b->n_reads == 0)
    {
        out_host->totals[0] = 0;
        out_host->seg_off[0] = 0;
        return SX_OK;
    }
    if (b->aln_off[b->n_reads] != b->n_alns) return sx_fail(ctx, SX_ERR_ARG, "sx_choose_realignment: aln_off does not end at n_alns");
    SX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st(ctx->s_compute);
    SX_CUDA(ctx, cudaEventRecord(ctx->ev_a, st));
    sx_realign_batch d(*b);
    void* p(nullptr);
    const size_t n_segs(b->aln_seg_off[b->n_alns]), n_keys(b->aln_key_off[b->n_alns]), n_win(b->region_key_off[b->n_regions]);
#define SX_UPX(slot, dst, src, type, bytes)                                                \
    if ((rc = sx_ensure(ctx, slot, (size_t)(bytes) + 16, &p))) return rc;                   \
    if (bytes) SX_CUDA(ctx, cudaMemcpyAsync(p, (src), (bytes), cudaMemcpyHostToDevice, st)); \
    dst = static_cast<type>(p);
    SX_UPX(0, d.region_read_off, b->region_read_off, const uint32_t*, (size_t)(b->n_regions + 1) * 4)
    SX_UPX(1, d.region_key_off, b->region_key_off, const uint32_t*, (size_t)(b->n_regions + 1) * 4)
    SX_UPX(2, d.keys, b->keys, const sx_indel_key*, n_win * sizeof(sx_indel_key))
    SX_UPX(3, d.aln_off, b->aln_off, const uint32_t*, (size_t)(b->n_reads + 1) * 4)
    SX_UPX(4, d.aln_pos, b->aln_pos, const int32_t*, (size_t)b->n_alns * 4)
    SX_UPX(5, d.aln_seg_off, b->aln_seg_off, const uint32_t*, ((size_t)b->n_alns + 1) * 4)
    SX_UPX(6, d.segs, b->segs, const sx_aln_seg*, n_segs * sizeof(sx_aln_seg))
    SX_UPX(7, d.aln_key_off, b->aln_key_off, const uint32_t*, ((size_t)b->n_alns + 1) * 4)
    SX_UPX(8, d.aln_keys, b->aln_keys, const uint16_t*, n_keys * 2)
    SX_UPX(9, d.read_len, b->read_len, const uint16_t*, (size_t)b->n_reads * 2)
    if (b->pin_flags)
    {
        SX_UPX(10, d.pin_flags, b->pin_flags, const uint8_t*, (size_t)b->n_reads)
    }
    const double* d_lnp(nullptr);
    SX_UPX(11, d_lnp, lnp_host, const double*, (size_t)b->n_alns * 8)
#undef SX_UPX
    sx_realign_out o(*out_host);
    if ((rc = sx_ensure(ctx, 12, 16, reinterpret_cast<void**>(&o.totals)))) return rc;
    if ((rc = sx_ensure(ctx, 13, (size_t)(b->n_reads + 1) * 4 + 16, reinterpret_cast<void**>(&o.seg_off)))) return rc;
    if ((rc = sx_ensure(ctx, 14, (size_t)b->n_reads * 4 + 16, reinterpret_cast<void**>(&o.pos)))) return rc;
    if ((rc = sx_ensure(ctx, 15, (size_t)b->n_reads * 2 + 16, reinterpret_cast<void**>(&o.n_seg)))) return rc;
    if ((rc = sx_ensure(ctx, 16, (size_t)b->n_reads + 16, reinterpret_cast<void**>(&o.status)))) return rc;
    if ((rc = sx_ensure(ctx, 17, (size_t)b->n_reads * 4 + 16, reinterpret_cast<void**>(&o.best_aln)))) return rc;
    if ((rc = sx_ensure(ctx, 18, (size_t)o.cap_segs * sizeof(sx_aln_seg) + 16, reinterpret_cast<void**>(&o.segs)))) return rc;
    unsigned launches(0);
    if ((rc = k9_run(ctx, &d, d_lnp, &o, &launches))) return rc;
    SX_CUDA(ctx, cudaMemcpyAsync(out_host->totals, o.totals, 4, cudaMemcpyDeviceToHost, st));
    SX_CUDA(ctx, cudaMemcpyAsync(out_host->seg_off, o.seg_off, (size_t)(b->n_reads + 1) * 4, cudaMemcpyDeviceToHost, st));
    SX_CUDA(ctx, cudaStreamSynchronize(st));
    if (out_host->totals[0] <= o.cap_segs)
    {
        SX_CUDA(ctx, cudaMemcpyAsync(out_host->pos, o.pos, (size_t)b->n_reads * 4, cudaMemcpyDeviceToHost, st));
        SX_CUDA(ctx, cudaMemcpyAsync(out_host->n_seg, o.n_seg, (size_t)b->n_reads * 2, cudaMemcpyDeviceToHost, st));
        SX_CUDA(ctx, cudaMemcpyAsync(out_host->status, o.status, (size_t)b->n_reads, cudaMemcpyDeviceToHost, st));
        SX_CUDA(ctx, cudaMemcpyAsync(out_host->best_aln, o.best_aln, (size_t)b->n_reads * 4, cudaMemcpyDeviceToHost, st));
        SX_CUDA(ctx, cudaMemcpyAsync(out_host->segs, o.segs, (size_t)out_host->totals[0] * sizeof(sx_aln_seg), cudaMemcpyDeviceToHost, st));
    }
    SX_CUDA(ctx, cudaEventRecord(ctx->ev_b, st));
    SX_CUDA(ctx, cudaStreamSynchronize(st));
    float ms(0);
    cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
    ctx->timing.kernel_ms = ms;
    ctx->timing.launches = launches;
    ctx->total_launches += launches;
    return k9_finish(ctx, "sx_choose_realignment", out_host->totals);
}

// asynchronous launcher for the device-resident pipeline (sx_pipeline.cu)
int sx_k9_run(sx_ctx* ctx, const sx_realign_batch* d, const double* lnp, const sx_realign_out* o, unsigned* launches)
{
    if (d->n_reads == 0) return SX_OK;
    unsigned l(0);
    const int rc(k9_run(ctx, d, lnp, o, &l));
    *launches += l;
    return rc;
}
