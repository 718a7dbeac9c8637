// flat_batch_normalize.h -- TEST INFRASTRUCTURE (oracle/): turns regions [r0, r1) of an sx_align_batch in any of the compact wire
// formats (SX_FMT_ALN8, SX_FMT_SEG2, SX_FMT_BASEQ, SX_FMT_REF4, qual_bits == 2; include/strelka_b200.h) into the wide form the oracle and the reference
// harness walk.  All indices stay ABSOLUTE: the returned batch's arrays are shifted pointers into range-sized vectors, so only
// regions [r0, r1] / their reads, alignments and segments may be touched through it.
#pragma once

#include "../include/strelka_b200.h"

#include <cstdint>
#include <vector>

struct sx_norm_batch
{
    sx_align_batch b;
    std::vector<sx_region> regions;
    std::vector<uint8_t> qual, seq;
    std::vector<char> ref;
    std::vector<sx_aln> alns;
    std::vector<sx_aln_seg> segs;
};

static inline const sx_align_batch* sx_normalize_range(const sx_align_batch* in, uint32_t r0, uint32_t r1, sx_norm_batch& n)
{
    if (in->format == 0 && in->qual_bits != 2) return in;
    n.b = *in;
    n.b.format = 0;
    n.regions.assign(in->regions + r0, in->regions + r1 + 1);
    n.b.regions = n.regions.data() - r0;
    const sx_region& first(in->regions[r0]);
    const sx_region& last(in->regions[r1]);
    if (in->format & SX_FMT_ALN8)
    {
        const sx_aln8* a8(reinterpret_cast<const sx_aln8*>(in->alns));
        n.alns.resize(last.aln_begin - first.aln_begin + 1);
        for (uint32_t ri = r0; ri < r1; ++ri)
        {
            const sx_region& reg(in->regions[ri]);
            for (uint32_t a = reg.aln_begin; a < in->regions[ri + 1].aln_begin; ++a)
            {
                sx_aln& o(n.alns[a - first.aln_begin]);
                o.read = reg.read_begin + a8[a].read;
                o.ref_pos = reg.ref_begin + a8[a].ref_pos;
                o.seg_off = reg.seg_begin + a8[a].seg_off;
                o.ins_off = reg.ins_begin + a8[a].ins_off;
            }
        }
        sx_aln& s(n.alns.back()); // the sentinel of the range: closes the last alignment's segment list
        s.read = last.read_begin;
        s.ref_pos = 0;
        s.seg_off = last.seg_begin;
        s.ins_off = last.ins_begin;
        n.b.alns = n.alns.data() - first.aln_begin;
    }
    if (in->format & SX_FMT_SEG2)
    {
        const sx_aln_seg2* s2(reinterpret_cast<const sx_aln_seg2*>(in->segs));
        n.segs.resize(last.seg_begin - first.seg_begin + 1);
        for (uint32_t s = first.seg_begin; s < last.seg_begin; ++s)
        {
            sx_aln_seg& o(n.segs[s - first.seg_begin]);
            o.len = s2[s] & 0xfffu;
            o.kind = (s2[s] >> 12) & 7u;
            o.flags = s2[s] >> 15;
        }
        n.b.segs = n.segs.data() - first.seg_begin;
    }
    if (in->format & SX_FMT_REF4)
    {
        // packed BAM codes -> ASCII, window by window, at new 16-byte aligned offsets
        static const char CODE2CHAR[16] = {'N', 'A', 'C', 'N', 'G', 'N', 'N', 'N', 'T', 'N', 'N', 'N', 'N', 'N', 'N', 'N'};
        uint64_t ro(0);
        for (uint32_t ri = r0; ri < r1; ++ri)
        {
            const sx_region& reg(in->regions[ri]);
            n.regions[ri - r0].ref_off = ro;
            for (uint32_t i = 0; i < reg.ref_len; ++i)
                n.ref.push_back(CODE2CHAR[(reinterpret_cast<const uint8_t*>(in->ref)[reg.ref_off + (i >> 1)] >> ((~i & 1) << 2)) & 15]);
            ro += reg.ref_len;
            while (ro & 15)
            {
                n.ref.push_back('N');
                ++ro;
            }
        }
        n.regions[r1 - r0].ref_off = ro;
        n.ref.resize(n.ref.size() + 64);
        n.b.ref = n.ref.data();
        n.b.ref_bytes = ro;
    }
    if (in->format & SX_FMT_BASEQ)
    {
        // (base << 2 | quality code) nibbles + exceptions -> BAM nibbles in a seq4 copy of the range, qualities to one byte per base
        static const uint8_t BASE2CODE[4] = {1, 2, 4, 8};
        const uint64_t s0(first.seq_off), s1(last.seq_off);
        n.seq.assign(in->seq4 + s0, in->seq4 + s1);
        n.seq.resize(n.seq.size() + 64);
        uint64_t qo(0);
        for (uint32_t ri = r0; ri < r1; ++ri)
        {
            const sx_region& reg(in->regions[ri]);
            n.regions[ri - r0].qual_off = qo;
            uint64_t so(0);
            for (uint32_t r = reg.read_begin; r < in->regions[ri + 1].read_begin; ++r)
            {
                const uint32_t len(in->read_len[r]);
                for (uint32_t i = 0; i < len; ++i)
                {
                    const uint64_t p(2 * so + i);
                    uint8_t& byte(n.seq[reg.seq_off - s0 + (p >> 1)]);
                    const unsigned sh((~p & 1) << 2);
                    const unsigned nib((byte >> sh) & 15);
                    n.qual.push_back(in->qual_dict[nib & 3]);
                    byte = static_cast<uint8_t>((byte & ~(15u << sh)) | (BASE2CODE[nib >> 2] << sh));
                }
                so += (len + 1) / 2;
                qo += len;
            }
            while (qo & 15)
            {
                n.qual.push_back(0);
                ++qo;
            }
            for (uint32_t e = in->exc_off[ri]; e < in->exc_off[ri + 1]; ++e)
            {
                const uint64_t p(in->exc[e] & 0xffffffu);
                uint8_t& byte(n.seq[reg.seq_off - s0 + (p >> 1)]);
                const unsigned sh((~p & 1) << 2);
                byte = static_cast<uint8_t>((byte & ~(15u << sh)) | ((in->exc[e] >> 24) << sh));
            }
        }
        n.regions[r1 - r0].qual_off = qo;
        n.qual.resize(n.qual.size() + 64);
        n.b.seq4 = n.seq.data() - s0;
        n.b.qual = n.qual.data();
        n.b.qual_bits = 8;
        n.b.qual_bytes = qo;
    }
    else if (in->qual_bits == 2)
    {
        uint64_t qo(0);
        for (uint32_t ri = r0; ri < r1; ++ri)
        {
            const sx_region& reg(in->regions[ri]);
            n.regions[ri - r0].qual_off = qo;
            uint64_t so(0); // packed-byte offset of the read inside the region's seq4 slice
            for (uint32_t r = reg.read_begin; r < in->regions[ri + 1].read_begin; ++r)
            {
                const uint32_t len(in->read_len[r]);
                for (uint32_t i = 0; i < len; ++i)
                {
                    const uint64_t p(2 * so + i);
                    const uint8_t code((in->qual[reg.qual_off + (p >> 2)] >> (6 - 2 * (p & 3))) & 3);
                    n.qual.push_back(in->qual_dict[code]);
                }
                so += (len + 1) / 2;
                qo += len;
            }
            while (qo & 15)
            {
                n.qual.push_back(0);
                ++qo;
            }
        }
        n.regions[r1 - r0].qual_off = qo;
        n.qual.resize(n.qual.size() + 64);
        n.b.qual = n.qual.data();
        n.b.qual_bits = 8;
        n.b.qual_bytes = qo;
    }
    return &n.b;
}
