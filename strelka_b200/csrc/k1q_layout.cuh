// k1q_layout.cuh -- shared-memory layout of the K1 fast path (k1_score4.cu), shared with the sizing code in k1_score.cu.
#pragma once

#include "sx_internal.h"

namespace k1q
{
constexpr uint32_t KQ_THREADS = 128;
// records and entries carry 16-bit shared-window addresses, and the term table must sit below 0x7f00 (see the kernel): regions that
// need more than this go to the general kernel (k1_score.cu)
constexpr uint32_t KQ_MAX_SMEM = 60u * 1024u;
constexpr uint32_t KQ_LUT_BYTES = 80; // 9 x 8
constexpr uint32_t KQ_TAB_RESERVE = 2048; // 4 pages x 16 rows x 2 doubles, placed on a 1 KB boundary of the shared window inside this reserve

__host__ __device__ __forceinline__ uint32_t pad16(uint32_t x) { return (x + 15u) & ~15u; }

struct layout
{
    uint32_t lut, e8, tab, alns, segs, recs, ref, refp, ins, seq, qual, rlen, soff, ent, total;
    uint32_t n_reads, n_alns, n_segs, seg_bytes, ref_bytes, refp_bytes, ins_bytes, seq_bytes, qual_bytes;
};

// bytes of the alignment-header slice as the kernels stage it: wide headers are 16 bytes each (+ the next one, which closes the last
// segment list); sx_aln8 headers are copied from the 16-byte boundary at or below the region's first header
__host__ __device__ __forceinline__ uint32_t aln_slice_bytes(uint32_t aln_begin, uint32_t n_alns, uint32_t fmt)
{
    return (fmt & SX_FMT_ALN8) ? pad16((n_alns + (aln_begin & 1u)) * 8u) : (n_alns + 1) * 16u;
}

__host__ __device__ __forceinline__ layout make_layout(const sx_region& r0, const sx_region& r1, uint32_t fmt)
{
    layout L;
    L.n_reads = r1.read_begin - r0.read_begin;
    L.n_alns = r1.aln_begin - r0.aln_begin;
    L.n_segs = r1.seg_begin - r0.seg_begin;
    L.seg_bytes = pad16(L.n_segs * ((fmt & SX_FMT_SEG2) ? 2u : 4u));
    L.ref_bytes = pad16(r0.ref_len);
    L.ins_bytes = pad16(r1.ins_begin - r0.ins_begin);
    L.seq_bytes = pad16(static_cast<uint32_t>(r1.seq_off - r0.seq_off));
    L.qual_bytes = (fmt & SX_FMT_BASEQ) ? 0u : pad16(static_cast<uint32_t>(r1.qual_off - r0.qual_off)); // BASEQ: qualities ride in the base nibbles
    L.refp_bytes = (fmt & SX_FMT_REF4) ? pad16((r0.ref_len + 1u) / 2u) : 0u;                         // REF4: packed window, unpacked into `ref`
    uint32_t o = 16; // mbarrier
    L.lut = o;
    o += KQ_LUT_BYTES;
    L.e8 = o;
    o += 256;
    L.tab = o;
    o += KQ_TAB_RESERVE;
    L.alns = o;
    o += aln_slice_bytes(r0.aln_begin, L.n_alns, fmt);
    L.segs = o;
    o += L.seg_bytes;
    L.recs = o; // one 8-byte record per segment + one END per alignment
    o += pad16((L.n_segs + L.n_alns) * 8u);
    L.ref = o;
    o += L.ref_bytes;
    L.refp = o;
    o += L.refp_bytes;
    L.ins = o;
    o += L.ins_bytes;
    L.seq = o;
    o += L.seq_bytes;
    L.qual = o;
    o += L.qual_bytes;
    L.rlen = o;
    o += pad16(L.n_reads * 2u);
    L.soff = o;
    o += pad16((L.n_reads + 1) * 4u);
    L.ent = o;
    // one byte per base (two per packed byte) + slack: the chunk loop loads (and masks off) up to 11 bytes past a run
    o += L.seq_bytes * 2u + 32u;
    L.total = o;
    return L;
}
} // namespace k1q

int sx_k1q_launch(sx_ctx* ctx, const sx_align_batch* dev, uint32_t region_begin, uint32_t region_end, double* lnp_dev, size_t smem_bytes, cudaStream_t st);
