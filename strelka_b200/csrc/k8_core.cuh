// k8_core.cuh -- the per-alignment body of K7b link_alignments (include/strelka_b200.h): one candidate alignment as K7 wrote it
// (reference path segments, window indices of its indel keys, edge keys) -> the same alignment as K1 reads it (flattened segments,
// the bases every inserted segment is scored against, non-candidate flags).
//
// This is the segment walk of scoreCandidateAlignment (starling_common/starling_read_align_score.cpp:289-499) with the three
// container look-ups resolved: getMatchingIndelKey :177-224, getInsertSeq :229-256 with the leading-edge tail rule :334-338 /
// :394-398, IndelBuffer::isCandidateIndel :473-475.  __host__ __device__ so that tests/cpp/k8_core_host.cpp can run exactly this
// code on the CPU (a test of the device logic; the library has no host execution path).
#pragma once

#include "strelka_b200.h"

#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define K8_HD __host__ __device__ __forceinline__
#else
#define K8_HD inline
#endif

enum
{
    K8_ST_NOKEY = 1, // a path gap that matches no key of its alignment (assert(isFound), score.cpp:222) / an edge gap without an edge key
    K8_ST_KIND = 2,  // "Can't handle cigar code" (score.cpp:461-466)
};

struct k8_view
{
    sx_enum_batch b;      // window keys, region offsets
    sx_enum_out e;        // K7's output
    const uint32_t* key_ins_off;
    const char* key_ins;
};

// K7's path segment kinds (ALIGNPATH::align_t) as sx_score_indels_batch wants them
K8_HD uint8_t k8_k6_kind(const unsigned t)
{
    switch (t)
    {
    case SX_AP_MATCH:
    case SX_AP_SEQ_MATCH:
    case SX_AP_SEQ_MISMATCH: return SX_SEG_MATCH;
    case SX_AP_INSERT: return SX_SEG_INSERT;
    case SX_AP_DELETE: return SX_SEG_DELETE;
    case SX_AP_SKIP: return SX_SEG_SKIP;
    case SX_AP_SOFT_CLIP: return SX_SEG_SOFTCLIP;
    default: return SX_SEG_HARDCLIP;
    }
}

K8_HD bool k8_align_match(const unsigned t) { return t == SX_AP_MATCH || t == SX_AP_SEQ_MATCH || t == SX_AP_SEQ_MISMATCH; }

// Walks alignment a (of a read of `region`).  With segs == nullptr it only counts; otherwise it writes the K1 segments to segs[] and
// the insert bytes to ins[].  Returns a K8_ST_* status (0 = ok).
K8_HD uint32_t k8_walk(const k8_view& v, const uint32_t region, const uint32_t a, uint32_t& n_seg, uint32_t& n_ins, sx_aln_seg* segs, char* ins)
{
    const uint32_t k0(v.b.region_key_off[region]);
    const sx_indel_key* win(v.b.keys + k0);
    const uint32_t s0(v.e.aln_seg_off[a]), aps(v.e.aln_seg_off[a + 1] - s0);
    const sx_aln_seg* path(v.e.segs + s0);
    const uint16_t* keys(v.e.aln_keys + v.e.aln_key_off[a]);
    const uint32_t n_keys(v.e.aln_key_off[a + 1] - v.e.aln_key_off[a]);
    // get_match_edge_segments, blt_util/align_path.cpp:736-752
    uint32_t first(aps), last(aps);
    for (uint32_t i = 0; i < aps; ++i)
        if (k8_align_match(path[i].kind))
        {
            if (first == aps) first = i;
            last = i;
        }
    n_seg = n_ins = 0;
    int32_t ref_head(v.e.aln_pos[a]);
    uint32_t i(0);
    while (i < aps)
    {
        const unsigned t(path[i].kind);
        const uint32_t len(path[i].len);
        uint32_t step(1);
        // a run of adjacent insert / delete segments with both kinds present is a swap (is_segment_swap_start, align_path.cpp:868-895)
        uint32_t j(i), insLen(0), delLen(0);
        for (; j < aps && (path[j].kind == SX_AP_INSERT || path[j].kind == SX_AP_DELETE); ++j) (path[j].kind == SX_AP_INSERT ? insLen : delLen) += path[j].len;
        const bool swap(insLen && delLen);
        const bool gap(swap || t == SX_AP_SEQ_MISMATCH || t == SX_AP_INSERT || t == SX_AP_DELETE);
        uint32_t del(0), insl(0);
        if (swap)
        {
            del = delLen;
            insl = insLen;
            step = j - i;
        }
        else if (t == SX_AP_SEQ_MISMATCH) del = insl = len;
        else if (t == SX_AP_INSERT) insl = len;
        else if (t == SX_AP_DELETE) del = len;
        if (gap)
        {
            // getMatchingIndelKey
            uint32_t w(SX_NO_KEY);
            if (i < first) w = v.e.aln_lead_key[a];
            else if (i > last) w = v.e.aln_trail_key[a];
            else
                for (uint32_t q = 0; q < n_keys; ++q)
                {
                    const sx_indel_key& k(win[keys[q]]);
                    if (k.pos == ref_head && k.del_len == del && k.ins_len == insl)
                    {
                        w = keys[q];
                        break;
                    }
                    if (k.pos > ref_head) break;
                }
            if (w == SX_NO_KEY) return K8_ST_NOKEY;
            const sx_indel_key& k(win[w]);
            const uint8_t flag((k.flags & SX_IKF_CANDIDATE) ? 0 : SX_SEGF_NONCANDIDATE);
            if (insl) // the inserted bases: the key's sequence, for a leading-edge segment its tail (:334-338, :394-398)
            {
                if (segs)
                {
                    const char* seq(v.key_ins + v.key_ins_off[k0 + w]);
                    const int32_t seqLen((int32_t)(v.key_ins_off[k0 + w + 1] - v.key_ins_off[k0 + w]));
                    const int32_t head(i < first ? seqLen - (int32_t)len : 0);
                    for (uint32_t x = 0; x < insl; ++x)
                    {
                        const int32_t p(head + (int32_t)x);
                        ins[n_ins + x] = (p >= 0 && p < seqLen) ? seq[p] : 'N'; // string_bam_seq::get_char out of range
                    }
                    segs[n_seg] = sx_aln_seg{(uint16_t)insl, SX_SEG_INSERT, flag};
                }
                n_seg++;
                n_ins += insl;
            }
            if (swap || t == SX_AP_SEQ_MISMATCH)
            {
                if (segs) segs[n_seg] = sx_aln_seg{(uint16_t)del, SX_SEG_REFSKIP, 0};
                n_seg++;
            }
            else if (t == SX_AP_DELETE)
            {
                if (segs) segs[n_seg] = sx_aln_seg{(uint16_t)del, SX_SEG_REFSKIP, flag};
                n_seg++;
            }
            ref_head += (int32_t)del;
        }
        else
        {
            uint8_t kind;
            if (t == SX_AP_MATCH || t == SX_AP_SEQ_MATCH)
            {
                kind = SX_SEG_MATCH;
                ref_head += (int32_t)len;
            }
            else if (t == SX_AP_SKIP)
            {
                kind = SX_SEG_REFSKIP;
                ref_head += (int32_t)len;
            }
            else if (t == SX_AP_SOFT_CLIP) kind = SX_SEG_SOFTCLIP;
            else if (t == SX_AP_HARD_CLIP) kind = SX_SEG_HARDCLIP;
            else return K8_ST_KIND;
            if (segs) segs[n_seg] = sx_aln_seg{(uint16_t)len, kind, 0};
            n_seg++;
        }
        i += step;
    }
    return 0;
}
