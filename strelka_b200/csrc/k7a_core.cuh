// k7a_core.cuh -- the per-read body of K7a alignment_indels (include/strelka_b200.h): which window entries a read's input alignment
// already contains.  Restates getAlignmentIndels(cal, ref, rseg, maxIndelSize, includeMismatches = true)
// (starling_common/CandidateAlignment.cpp:58-173) and the edge keys of getCandidateAlignment (starling_read_align.cpp:1481-1522) on
// K1's read / reference pools, with IndelKeys as window indices.  __host__ __device__: tests/cpp/k7a_core_host.cpp runs exactly this
// code on the CPU against the reference's own functions.
#pragma once

#include "strelka_b200.h"

#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define K7A_HD __host__ __device__ __forceinline__
#else
#define K7A_HD inline
#endif

#define K7A_MAX_KEYS 64u // keys of one input alignment (indels + window mismatches); more: the surplus is reported as SX_NO_KEY

struct k7a_view
{
    sx_enum_batch b;
    const sx_region* regions;
    const uint8_t* seq4;
    const char* ref;
    const uint32_t* key_ins_off;
    const char* key_ins;
};

K7A_HD uint8_t k7a_code(const uint8_t* seq4, const uint64_t read_byte, const uint32_t i) // bam_seq::get_code
{
    const uint8_t v(seq4[read_byte + (i >> 1)]);
    return (i & 1) ? (uint8_t)(v & 15) : (uint8_t)(v >> 4);
}
K7A_HD char k7a_char(const uint8_t c) // get_bam_seq_char, htsapi/bam_seq.hh:50-71
{
    return c == 0 ? '=' : c == 1 ? 'A' : c == 2 ? 'C' : c == 4 ? 'G' : c == 8 ? 'T' : 'N';
}
K7A_HD uint8_t k7a_ref_code(const k7a_view& v, const sx_region& g, const int32_t pos) // rc_segment_bam_seq::get_code: get_bam_seq_code(ref.get_base(pos))
{
    const int64_t i((int64_t)pos - g.ref_begin);
    const char c((i >= 0 && i < (int64_t)g.ref_len) ? v.ref[g.ref_off + (uint64_t)i] : 'N');
    return c == '=' ? 0 : c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 15;
}

// window index of IndelKey(pos, type, del, <ins_len read bases from read offset ro | the single char mm>), or SX_NO_KEY
K7A_HD uint16_t k7a_find(const k7a_view& v, const uint32_t k0, const uint32_t n_win, const int32_t pos, const unsigned type, const uint32_t del, const uint32_t ins_len,
                         const uint64_t read_byte, const uint32_t ro, const char mm)
{
    const sx_indel_key* win(v.b.keys + k0);
    uint32_t lo(0), hi(n_win);
    while (lo < hi)
    {
        const uint32_t mid((lo + hi) / 2);
        if (win[mid].pos < pos) lo = mid + 1;
        else hi = mid;
    }
    for (uint32_t k = lo; k < n_win && win[k].pos == pos; ++k)
    {
        if (win[k].type != type || win[k].del_len != del || win[k].ins_len != ins_len) continue;
        const char* seq(v.key_ins + v.key_ins_off[k0 + k]);
        bool same(true);
        for (uint32_t x = 0; x < ins_len && same; ++x) same = (seq[x] == (type == SX_INDEL_TYPE_MISMATCH ? mm : k7a_char(k7a_code(v.seq4, read_byte, ro + x))));
        if (same) return (uint16_t)k;
    }
    return SX_NO_KEY;
}

K7A_HD void k7a_add(uint16_t* keys, uint32_t& n, const uint16_t w) // std::set insert; a full array turns the surplus into SX_NO_KEY
{
    uint32_t j(0);
    while (j < n && keys[j] < w) ++j;
    if (j < n && keys[j] == w) return;
    if (n >= K7A_MAX_KEYS)
    {
        keys[K7A_MAX_KEYS - 1] = SX_NO_KEY;
        return;
    }
    for (uint32_t i = n; i > j; --i) keys[i] = keys[i - 1];
    keys[j] = w;
    ++n;
}

// keys[K7A_MAX_KEYS] <- the read's window indices (ascending, SX_NO_KEY last); returns their number
K7A_HD uint32_t k7a_read(const k7a_view& v, const uint32_t region, const uint32_t r, const uint64_t read_byte, uint16_t* keys, uint16_t& lead, uint16_t& trail)
{
    const sx_region g(v.regions[region]);
    const uint32_t k0(v.b.region_key_off[region]), n_win(v.b.region_key_off[region + 1] - k0);
    const uint32_t s0(v.b.in_seg_off[r]), aps(v.b.in_seg_off[r + 1] - s0);
    const sx_aln_seg* path(v.b.in_segs + s0);
    uint32_t first(aps), last(aps); // get_match_edge_segments
    for (uint32_t i = 0; i < aps; ++i)
    {
        const unsigned t(path[i].kind);
        if (t == SX_AP_MATCH || t == SX_AP_SEQ_MATCH || t == SX_AP_SEQ_MISMATCH)
        {
            if (first == aps) first = i;
            last = i;
        }
    }
    uint32_t n(0), ro(0);
    int32_t ref_pos(v.b.in_pos[r]);
    lead = trail = SX_NO_KEY;
    bool hasLead(false), hasTrail(false);
    uint32_t i(0);
    while (i < aps)
    {
        const unsigned t(path[i].kind);
        const uint32_t len(path[i].len);
        uint32_t j(i + 1);
        const bool edge(i < first || i > last);
        // is_segment_swap_start (align_path.cpp:868-895): a run of adjacent insert / delete segments holding both kinds
        uint32_t insLen(0), delLen(0), q(i);
        for (; q < aps && (path[q].kind == SX_AP_INSERT || path[q].kind == SX_AP_DELETE); ++q) (path[q].kind == SX_AP_INSERT ? insLen : delLen) += path[q].len;
        const bool swap(insLen && delLen);
        if (edge)
        {
            if (t == SX_AP_INSERT || t == SX_AP_DELETE) // ignore all edge segments except INSERT / DELETE (:89-104); the key is the edge key
            {
                const uint16_t w(k7a_find(v, k0, n_win, ref_pos, SX_INDEL_TYPE_INDEL, t == SX_AP_DELETE ? len : 0, t == SX_AP_INSERT ? len : 0, read_byte, ro, 0));
                // getCandidateAlignment (:1495-1518) sets the edge key anew for every edge insert / delete segment -- the last one wins --
                // and getAlignmentIndels inserts that final key (:94-103), once
                if (i < first)
                {
                    lead = w;
                    hasLead = true;
                }
                else
                {
                    trail = w;
                    hasTrail = true;
                }
            }
        }
        else if (swap)
        {
            j = q;
            const uint32_t m(insLen > delLen ? insLen : delLen);
            k7a_add(keys, n, m <= v.b.opts.max_indel_size ? k7a_find(v, k0, n_win, ref_pos, SX_INDEL_TYPE_INDEL, delLen, insLen, read_byte, ro, 0) : (uint16_t)SX_NO_KEY);
        }
        else if (t == SX_AP_INSERT || t == SX_AP_DELETE)
        {
            k7a_add(keys, n,
                    len <= v.b.opts.max_indel_size ? k7a_find(v, k0, n_win, ref_pos, SX_INDEL_TYPE_INDEL, t == SX_AP_DELETE ? len : 0, t == SX_AP_INSERT ? len : 0, read_byte, ro, 0)
                                                   : (uint16_t)SX_NO_KEY); // (a breakend pair in the reference: not part of this build)
        }
        else if (t == SX_AP_MATCH || t == SX_AP_SEQ_MATCH || t == SX_AP_SEQ_MISMATCH)
        {
            for (uint32_t x = 0; x < len; ++x)
            {
                const uint8_t sbase(k7a_code(v.seq4, read_byte, ro + x));
                if (sbase == 0 || sbase == 15) continue;
                const int32_t rp(ref_pos + (int32_t)x);
                if (sbase == k7a_ref_code(v, g, rp)) continue;
                const uint16_t w(k7a_find(v, k0, n_win, rp, SX_INDEL_TYPE_MISMATCH, 1, 1, read_byte, 0, k7a_char(sbase)));
                if (w != SX_NO_KEY) k7a_add(keys, n, w); // a mismatch that is no window entry is dropped (starling_read_align.cpp:1865)
            }
        }
        for (uint32_t s = i; s < j; ++s) // increment_path
        {
            const unsigned k(path[s].kind);
            if (k == SX_AP_MATCH || k == SX_AP_INSERT || k == SX_AP_SOFT_CLIP || k == SX_AP_SEQ_MATCH || k == SX_AP_SEQ_MISMATCH) ro += path[s].len;
            if (k == SX_AP_MATCH || k == SX_AP_DELETE || k == SX_AP_SKIP || k == SX_AP_SEQ_MATCH || k == SX_AP_SEQ_MISMATCH) ref_pos += (int32_t)path[s].len;
        }
        i = j;
    }
    if (hasLead) k7a_add(keys, n, lead);
    if (hasTrail) k7a_add(keys, n, trail);
    return n;
}
