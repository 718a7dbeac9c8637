// strelka_b200.hh -- C++ host side above the C ABI (include/strelka_b200.h).
//
// The reference is C++, so this is the layer a reference developer programs against: the same nouns and argument meaning as the
// reference's own interface for this path (paths relative to /root/reference/src/c++/lib/), batch-oriented where the reference is
// per-object, and free of reference headers so it builds stand-alone:
//
//   sx::path_segment / sx::path_t         ALIGNPATH::path_segment, path_t                 blt_util/align_path.hh:171
//   sx::IndelKey                          IndelKey{pos,type,deletionLength,insertSequence} starling_common/IndelKey.hh:39-193
//   sx::CandidateAlignment                CandidateAlignment{al, indels, leading/trailing} starling_common/CandidateAlignment.hh:36-83
//   sx::ReadAlignBatch::scoreCandidateAlignments   the loop at starling_read_align.cpp:1564-1571 over scoreCandidateAlignment
//   sx::IndelScoreBatch::scoreIndels      score_indels (starling_read_align_score_indels.cpp:454) + the arg-max of scoreCandidateAlignments
//   sx::ReadPathScores                    ReadPathScores                                   starling_common/IndelData.hh:64-116
//   sx::AlignmentScores<int>, sx::GlobalAligner<int>::align, sx::AlignmentResult<int>   alignment/{AlignmentScores,GlobalAligner}.hh
//   sx::Context                           RAII sx_ctx; failures throw sx::Exception (the reference throws blt_exception)
//
// Header-only; link with -lstrelka_b200.
#pragma once

#include "strelka_b200.h"

#include <algorithm>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace sx
{
typedef int32_t pos_t;

struct Exception : public std::runtime_error
{
    Exception(int c, const std::string& m) : std::runtime_error(m), code(c) {}
    int code;
};

class Context
{
public:
    explicit Context(int cudaDevice = 0, const sx_params* params = nullptr)
    {
        sx_params p;
        if (params) p = *params;
        else sx_default_params(&p);
        const int rc(sx_create(cudaDevice, &p, &_ctx));
        if (rc != SX_OK) throw Exception(rc, sx_last_error(nullptr));
    }
    ~Context() { sx_destroy(_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    sx_ctx* get() const { return _ctx; }
    void check(int rc) const
    {
        if (rc != SX_OK) throw Exception(rc, sx_last_error(_ctx));
    }

private:
    sx_ctx* _ctx = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------------------
// alignment paths and indel keys
// ---------------------------------------------------------------------------------------------------------------------------
namespace ALIGNPATH
{
enum align_t { NONE, MATCH, INSERT, DELETE, SKIP, SOFT_CLIP, HARD_CLIP, PAD, SEQ_MATCH, SEQ_MISMATCH }; // blt_util/align_path.hh:41-55
}
struct path_segment
{
    path_segment(ALIGNPATH::align_t t = ALIGNPATH::NONE, unsigned l = 0) : type(t), length(l) {}
    ALIGNPATH::align_t type;
    unsigned length;
};
typedef std::vector<path_segment> path_t;

inline ALIGNPATH::align_t cigar_code_to_segment_type(char c)
{
    using namespace ALIGNPATH;
    switch (c)
    {
    case 'M': return MATCH;
    case 'I': return INSERT;
    case 'D': return DELETE;
    case 'N': return SKIP;
    case 'S': return SOFT_CLIP;
    case 'H': return HARD_CLIP;
    case 'P': return PAD;
    case '=': return SEQ_MATCH;
    case 'X': return SEQ_MISMATCH;
    default: return NONE;
    }
}
inline void cigar_to_apath(const char* cigar, path_t& apath)
{
    apath.clear();
    unsigned n(0);
    for (const char* p(cigar); *p; ++p)
    {
        if (*p >= '0' && *p <= '9') n = n * 10 + (*p - '0');
        else
        {
            apath.push_back(path_segment(cigar_code_to_segment_type(*p), n));
            n = 0;
        }
    }
}
inline std::string apath_to_cigar(const path_t& apath)
{
    static const char* codes = "?MIDNSHP=X";
    std::string s;
    for (const path_segment& ps : apath) s += std::to_string(ps.length) + codes[ps.type];
    return s;
}

namespace INDEL
{
enum index_t { NONE, INDEL, MISMATCH, BP_LEFT, BP_RIGHT }; // starling_common/indel_core.hh:57-66
}
struct IndelKey
{
    IndelKey(pos_t p = 0, INDEL::index_t t = INDEL::NONE, unsigned l = 0, const char* is = "") : pos(p), type(t), deletionLength(l), insertSequence(is) {}
    unsigned insert_length() const { return insertSequence.size(); }
    unsigned delete_length() const { return deletionLength; }
    bool isMismatch() const { return type == INDEL::MISMATCH; }
    bool operator<(const IndelKey& rhs) const // IndelKey.hh:53-76
    {
        if (pos != rhs.pos) return pos < rhs.pos;
        if (type != rhs.type) return type < rhs.type;
        if (type == INDEL::NONE || type == INDEL::BP_LEFT || type == INDEL::BP_RIGHT) return false;
        if (insert_length() != rhs.insert_length()) return insert_length() < rhs.insert_length();
        if (delete_length() != rhs.delete_length()) return delete_length() < rhs.delete_length();
        return insertSequence < rhs.insertSequence;
    }
    bool operator==(const IndelKey& rhs) const { return pos == rhs.pos && type == rhs.type && deletionLength == rhs.deletionLength && insertSequence == rhs.insertSequence; }
    pos_t pos;
    INDEL::index_t type;
    unsigned deletionLength;
    std::string insertSequence;
};

struct alignment
{
    path_t path;
    pos_t pos = 0;
    bool is_fwd_strand = true;
};

struct CandidateAlignment
{
    alignment al;
    std::vector<IndelKey> indels; ///< getIndels(), ordered by position like indel_set_t
    IndelKey leading_indel_key;
    IndelKey trailing_indel_key;
};

// ---------------------------------------------------------------------------------------------------------------------------
// K1: batch of regions / reads / candidate alignments -> ln P(read | path)
// ---------------------------------------------------------------------------------------------------------------------------
class ReadAlignBatch
{
public:
    /// isCandidate(key): what IndelBuffer::isCandidateIndel would answer (starling_read_align_score.cpp:473-475);
    /// insertSeqOf(key): getInsertSeq (score.cpp:229-256) -- the key's own sequence unless it is a breakpoint
    typedef std::function<bool(const IndelKey&)> candidate_fn;

    void beginRegion(const std::string& ref, pos_t refBegin)
    {
        closeRegion();
        pad();
        sx_region r;
        std::memset(&r, 0, sizeof(r));
        r.seq_off = _seq4.size();
        r.qual_off = _qual.size();
        r.ref_off = _ref.size();
        r.read_begin = _readLen.size();
        r.aln_begin = _alns.size();
        r.seg_begin = _segs.size();
        r.ins_begin = _ins.size();
        r.ref_begin = refBegin;
        r.ref_len = ref.size();
        _regions.push_back(r);
        _ref.insert(_ref.end(), ref.begin(), ref.end());
        _open = true;
    }

    /// read given as BAM 4-bit codes (bam_seq::get_code) and qualities; returns the read's index in the batch
    unsigned addRead(const uint8_t* codes, const uint8_t* qual, unsigned len)
    {
        require(_open, "addRead outside a region");
        _readLen.push_back(len);
        for (unsigned i(0); i < len; i += 2) _seq4.push_back(static_cast<uint8_t>(((codes[i] & 15) << 4) | ((i + 1 < len) ? (codes[i + 1] & 15) : 0)));
        _qual.insert(_qual.end(), qual, qual + len);
        return _readLen.size() - 1;
    }
    unsigned addRead(const std::string& bases, const uint8_t* qual)
    {
        std::vector<uint8_t> codes(bases.size());
        for (size_t i(0); i < bases.size(); ++i) codes[i] = get_bam_seq_code(bases[i]);
        return addRead(codes.data(), qual, bases.size());
    }

    /// flattening of one CandidateAlignment: the segment walk of scoreCandidateAlignment (score.cpp:289-499) with every host-side
    /// lookup resolved (getMatchingIndelKey :172-224, leading-edge insert tail :334-338/:394-398, candidacy :473-475)
    void addCandidateAlignment(unsigned readIndex, const CandidateAlignment& cal, const candidate_fn& isCandidate)
    {
        using namespace ALIGNPATH;
        require(_open, "addCandidateAlignment outside a region");
        require(_alns.empty() || _alns.back().read <= readIndex, "alignments must be added in read order");
        const path_t& path(cal.al.path);
        const unsigned aps(path.size());
        std::pair<unsigned, unsigned> ends(aps, aps); // get_match_edge_segments, blt_util/align_path.cpp:736-752
        {
            bool isFirst(false);
            for (unsigned i(0); i < aps; ++i)
                if (isAlignMatch(path[i].type))
                {
                    if (!isFirst) ends.first = i;
                    isFirst = true;
                    ends.second = i;
                }
        }
        sx_aln a{readIndex, cal.al.pos, static_cast<uint32_t>(_segs.size()), static_cast<uint32_t>(_ins.size())};
        pos_t ref_head_pos(cal.al.pos);
        unsigned path_index(0);
        while (path_index < aps)
        {
            const path_segment& ps(path[path_index]);
            unsigned n_seg(1);
            // is_segment_swap_start, blt_util/align_path.cpp:868-895
            unsigned j(path_index), insLen(0), delLen(0);
            for (; j < aps && (path[j].type == INSERT || path[j].type == DELETE); ++j) (path[j].type == INSERT ? insLen : delLen) += path[j].length;
            const bool isSwap(insLen && delLen);
            if (isSwap || ps.type == SEQ_MISMATCH)
            {
                unsigned del, ins;
                if (ps.type == SEQ_MISMATCH) del = ins = ps.length;
                else
                {
                    del = delLen;
                    ins = insLen;
                    n_seg = j - path_index;
                }
                const IndelKey& k(matchingKey(cal, ref_head_pos, del, ins, ends, path_index));
                pushInsert(k, ps.length, ins, path_index < ends.first, isCandidate(k) ? 0 : SX_SEGF_NONCANDIDATE);
                pushSeg(del, SX_SEG_REFSKIP, 0);
                ref_head_pos += del;
            }
            else if (isAlignMatch(ps.type))
            {
                pushSeg(ps.length, SX_SEG_MATCH, 0);
                ref_head_pos += ps.length;
            }
            else if (ps.type == INSERT)
            {
                const IndelKey& k(matchingKey(cal, ref_head_pos, 0, ps.length, ends, path_index));
                pushInsert(k, ps.length, ps.length, path_index < ends.first, isCandidate(k) ? 0 : SX_SEGF_NONCANDIDATE);
            }
            else if (ps.type == DELETE)
            {
                const IndelKey& k(matchingKey(cal, ref_head_pos, ps.length, 0, ends, path_index));
                pushSeg(ps.length, SX_SEG_REFSKIP, isCandidate(k) ? 0 : SX_SEGF_NONCANDIDATE);
                ref_head_pos += ps.length;
            }
            else if (ps.type == SKIP)
            {
                pushSeg(ps.length, SX_SEG_REFSKIP, 0);
                ref_head_pos += ps.length;
            }
            else if (ps.type == SOFT_CLIP) pushSeg(ps.length, SX_SEG_SOFTCLIP, 0);
            else if (ps.type == HARD_CLIP) pushSeg(ps.length, SX_SEG_HARDCLIP, 0);
            else throw Exception(SX_ERR_ARG, "Can't handle cigar code"); // score.cpp:461-466
            path_index += n_seg;
        }
        _alns.push_back(a);
    }

    /// the loop of starling_read_align.cpp:1564-1571 for every alignment in the batch
    void scoreCandidateAlignments(const Context& ctx, std::vector<double>& candAlignmentScores)
    {
        const sx_align_batch b(view());
        candAlignmentScores.assign(b.n_alns, 0.);
        ctx.check(sx_score_alignments(ctx.get(), &b, candAlignmentScores.data()));
    }

    /// closes the batch (sentinels, slack) and returns the ABI view; valid until the next mutation
    sx_align_batch view()
    {
        closeRegion();
        pad();
        _regionsOut = _regions;
        sx_region s;
        std::memset(&s, 0, sizeof(s));
        s.seq_off = _seq4.size();
        s.qual_off = _qual.size();
        s.ref_off = _ref.size();
        s.read_begin = _readLen.size();
        s.aln_begin = _alns.size();
        s.seg_begin = _segs.size();
        s.ins_begin = _ins.size();
        _regionsOut.push_back(s);
        _alnsOut = _alns;
        _alnsOut.push_back(sx_aln{static_cast<uint32_t>(_readLen.size()), 0, static_cast<uint32_t>(_segs.size()), static_cast<uint32_t>(_ins.size())});
        _seq4Out = _seq4;
        _qualOut = _qual;
        _refOut = _ref;
        _insOut = _ins;
        _segsOut = _segs;
        _seq4Out.resize(_seq4.size() + SX_POOL_SLACK);
        _qualOut.resize(_qual.size() + SX_POOL_SLACK);
        _refOut.resize(_ref.size() + SX_POOL_SLACK);
        _insOut.resize(_ins.size() + SX_POOL_SLACK);
        _segsOut.resize(_segs.size() + 16, sx_aln_seg{0, SX_SEG_HARDCLIP, 0});
        sx_align_batch b;
        std::memset(&b, 0, sizeof(b));
        b.n_regions = _regions.size();
        b.n_reads = _readLen.size();
        b.n_alns = _alns.size();
        b.n_segs = _segs.size();
        b.regions = _regionsOut.data();
        b.read_len = _readLen.data();
        b.seq4 = _seq4Out.data();
        b.qual = _qualOut.data();
        b.ref = _refOut.data();
        b.alns = _alnsOut.data();
        b.segs = _segsOut.data();
        b.ins = _insOut.data();
        b.seq4_bytes = _seq4.size();
        b.qual_bytes = _qual.size();
        b.ref_bytes = _ref.size();
        b.ins_bytes = _ins.size();
        if (_compact) compactView(b);
        return b;
    }

    /// false: always send the wide structs and one quality byte per base (default: the most compact formats that fit)
    void setCompactWireFormats(bool on) { _compact = on; }

    static uint8_t get_bam_seq_code(char c) // htsapi/bam_seq.hh:98-118
    {
        switch (c)
        {
        case '=': return 0;
        case 'A': return 1;
        case 'C': return 2;
        case 'G': return 4;
        case 'T': return 8;
        default: return 15;
        }
    }

private:
    static bool isAlignMatch(ALIGNPATH::align_t t) { return t == ALIGNPATH::MATCH || t == ALIGNPATH::SEQ_MATCH || t == ALIGNPATH::SEQ_MISMATCH; }
    static void require(bool ok, const char* msg)
    {
        if (!ok) throw Exception(SX_ERR_ARG, msg);
    }
    void closeRegion() { _open = false; }
    void pad()
    {
        // staging rule of include/strelka_b200.h: every region slice starts 16-byte aligned; pad segments are no-op hard clips
        while (_seq4.size() & 15) _seq4.push_back(0);
        while (_qual.size() & 15) _qual.push_back(0);
        while (_ref.size() & 15) _ref.push_back(0);
        while (_ins.size() & 15) _ins.push_back(0);
        while (_segs.size() & 7) _segs.push_back(sx_aln_seg{0, SX_SEG_HARDCLIP, 0}); // 8: also valid for the 2-byte segment format
    }
    void pushSeg(unsigned len, uint8_t kind, uint8_t flags) { _segs.push_back(sx_aln_seg{static_cast<uint16_t>(len), kind, flags}); }
    void pushInsert(const IndelKey& k, unsigned firstSegLen, unsigned insertLength, bool isLeadingEdge, uint8_t flags)
    {
        const std::string& seq(k.insertSequence);
        int head(0);
        if (isLeadingEdge) head = static_cast<int>(seq.size()) - static_cast<int>(firstSegLen);
        for (unsigned i(0); i < insertLength; ++i)
        {
            const int p(head + static_cast<int>(i));
            _ins.push_back((p >= 0 && p < static_cast<int>(seq.size())) ? seq[p] : 'N'); // string_bam_seq::get_char out of range
        }
        pushSeg(insertLength, SX_SEG_INSERT, flags);
    }
    static const IndelKey& matchingKey(const CandidateAlignment& cal, pos_t ref_head_pos, unsigned del, unsigned ins, const std::pair<unsigned, unsigned>& ends,
                                       unsigned path_index)
    {
        if (path_index < ends.first) return cal.leading_indel_key;
        if (path_index > ends.second) return cal.trailing_indel_key;
        for (const IndelKey& k : cal.indels)
            if (k.pos == ref_head_pos && (k.type == INDEL::INDEL || k.isMismatch()) && k.delete_length() == del && k.insert_length() == ins) return k;
        throw Exception(SX_ERR_ARG, "candidate alignment does not contain the indel its path implies"); // assert(isFound), score.cpp:222
    }

    /// Compact wire formats (include/strelka_b200.h: sx_aln8, sx_aln_seg2, 4-/2-bit quality codes).  Lossless; each is used only
    /// when every value of the batch fits it.  The host entry points are PCIe-bound, so bytes are throughput.
    void compactView(sx_align_batch& b)
    {
        // ---- qualities: dictionary of the distinct values
        bool seen[256] = {false};
        for (size_t ri(0); ri < _regions.size(); ++ri)
        {
            size_t qo(_regions[ri].qual_off);
            const size_t r1(ri + 1 < _regions.size() ? _regions[ri + 1].read_begin : _readLen.size());
            for (size_t r(_regions[ri].read_begin); r < r1; ++r)
                for (unsigned i(0); i < _readLen[r]; ++i) seen[_qual[qo++]] = true;
        }
        unsigned nq(0);
        uint8_t code[256] = {0};
        for (unsigned q(0); q < 256; ++q)
            if (seen[q])
            {
                if (nq < 16) b.qual_dict[nq] = static_cast<uint8_t>(q);
                code[q] = static_cast<uint8_t>(nq++);
            }
        const unsigned qbits(nq <= 4 ? 2 : nq <= 16 ? 4 : 8);
        bool allCallable(true); // SX_FMT_BASEQ needs every quality <= 70 (a larger one must reach the kernel's range check)
        for (unsigned q(71); q < 256; ++q) allCallable = allCallable && !seen[q];
        if (qbits == 2 && allCallable)
        {
            // base and quality code in one nibble; the bases that are not A/C/G/T go to the exception list
            _excOff.assign(_regions.size() + 1, 0);
            _exc.clear();
            for (size_t ri(0); ri < _regions.size(); ++ri)
            {
                _excOff[ri] = static_cast<uint32_t>(_exc.size());
                size_t qo(_regions[ri].qual_off), so(_regions[ri].seq_off);
                const size_t r1(ri + 1 < _regions.size() ? _regions[ri + 1].read_begin : _readLen.size());
                for (size_t r(_regions[ri].read_begin); r < r1; ++r)
                {
                    const unsigned len(_readLen[r]);
                    for (unsigned i(0); i < len; ++i)
                    {
                        uint8_t& byte(_seq4Out[so + (i >> 1)]);
                        const unsigned sh((~i & 1) << 2), bam((byte >> sh) & 15), qc(code[_qual[qo++]]);
                        unsigned base(0);
                        if (bam == 1) base = 0;
                        else if (bam == 2) base = 1;
                        else if (bam == 4) base = 2;
                        else if (bam == 8) base = 3;
                        else _exc.push_back(SX_EXC(2 * (so - _regions[ri].seq_off) + i, bam));
                        byte = static_cast<uint8_t>((byte & ~(15u << sh)) | (((base << 2) | qc) << sh));
                    }
                    so += (len + 1) / 2;
                }
                _regionsOut[ri].qual_off = 0;
            }
            _excOff[_regions.size()] = static_cast<uint32_t>(_exc.size());
            _exc.push_back(0);
            _regionsOut.back().qual_off = 0;
            b.qual_bytes = 0;
            b.qual_bits = 2;
            b.exc_off = _excOff.data();
            b.exc = _exc.data();
            b.format |= SX_FMT_BASEQ;
        }
        else if (qbits != 8)
        {
            _qualOut.clear();
            for (size_t ri(0); ri < _regions.size(); ++ri)
            {
                while (_qualOut.size() & 15) _qualOut.push_back(0);
                size_t qo(_regions[ri].qual_off);
                _regionsOut[ri].qual_off = _qualOut.size();
                const size_t r1(ri + 1 < _regions.size() ? _regions[ri + 1].read_begin : _readLen.size());
                std::vector<uint8_t> codes; // one code per nibble position of the region's seq4 slice
                for (size_t r(_regions[ri].read_begin); r < r1; ++r)
                {
                    const unsigned len(_readLen[r]);
                    for (unsigned i(0); i < len; ++i) codes.push_back(code[_qual[qo++]]);
                    if (len & 1) codes.push_back(0);
                }
                if (qbits == 4)
                    for (size_t i(0); i < codes.size(); i += 2) _qualOut.push_back(static_cast<uint8_t>((codes[i] << 4) | codes[i + 1]));
                else
                {
                    while (codes.size() & 3) codes.push_back(0);
                    for (size_t i(0); i < codes.size(); i += 4)
                        _qualOut.push_back(static_cast<uint8_t>((codes[i] << 6) | (codes[i + 1] << 4) | (codes[i + 2] << 2) | codes[i + 3]));
                }
            }
            while (_qualOut.size() & 15) _qualOut.push_back(0);
            _regionsOut.back().qual_off = _qualOut.size();
            b.qual_bytes = _qualOut.size();
            _qualOut.resize(_qualOut.size() + SX_POOL_SLACK);
            b.qual = _qualOut.data();
            b.qual_bits = qbits;
        }
        else std::memset(b.qual_dict, 0, sizeof(b.qual_dict));
        // ---- alignment headers: region-relative 16-bit fields
        bool fits(true);
        for (size_t ri(0); ri < _regions.size() && fits; ++ri)
        {
            const sx_region& reg(_regions[ri]);
            const size_t a1(ri + 1 < _regions.size() ? _regions[ri + 1].aln_begin : _alns.size());
            for (size_t a(reg.aln_begin); a < a1; ++a)
            {
                const int64_t rp(static_cast<int64_t>(_alns[a].ref_pos) - reg.ref_begin);
                if (_alns[a].read - reg.read_begin > 65535u || rp < -32768 || rp > 32767 || _alns[a].seg_off - reg.seg_begin > 65535u || _alns[a].ins_off - reg.ins_begin > 65535u)
                {
                    fits = false;
                    break;
                }
            }
        }
        if (fits)
        {
            _alns8Out.assign(_alns.size() + 3, sx_aln8{0, 0, 0, 0}); // + slack: the kernels stage a slice from a 16-byte boundary
            for (size_t ri(0); ri < _regions.size(); ++ri)
            {
                const sx_region& reg(_regions[ri]);
                const size_t a1(ri + 1 < _regions.size() ? _regions[ri + 1].aln_begin : _alns.size());
                for (size_t a(reg.aln_begin); a < a1; ++a)
                    _alns8Out[a] = sx_aln8{static_cast<uint16_t>(_alns[a].read - reg.read_begin), static_cast<int16_t>(_alns[a].ref_pos - reg.ref_begin),
                                           static_cast<uint16_t>(_alns[a].seg_off - reg.seg_begin), static_cast<uint16_t>(_alns[a].ins_off - reg.ins_begin)};
            }
            b.alns = reinterpret_cast<const sx_aln*>(_alns8Out.data());
            b.format |= SX_FMT_ALN8;
        }
        // ---- reference windows as BAM 4-bit codes
        {
            std::vector<char> packed;
            for (size_t ri(0); ri < _regions.size(); ++ri)
            {
                while (packed.size() & 15) packed.push_back(0);
                const size_t o(_regions[ri].ref_off), n(_regions[ri].ref_len);
                _regionsOut[ri].ref_off = packed.size();
                for (size_t i(0); i < n; i += 2)
                    packed.push_back(static_cast<char>((get_bam_seq_code(_ref[o + i] == '=' ? 'N' : _ref[o + i]) << 4) |
                                                       (i + 1 < n ? get_bam_seq_code(_ref[o + i + 1] == '=' ? 'N' : _ref[o + i + 1]) : 0)));
            }
            while (packed.size() & 15) packed.push_back(0);
            _regionsOut.back().ref_off = packed.size();
            b.ref_bytes = packed.size();
            packed.resize(packed.size() + SX_POOL_SLACK);
            _refOut.swap(packed);
            b.ref = _refOut.data();
            b.format |= SX_FMT_REF4;
        }
        // ---- segments: 12-bit lengths
        fits = true;
        for (const sx_aln_seg& sg : _segs) fits = fits && sg.len <= 4095;
        if (fits)
        {
            _segs2Out.assign(_segs.size() + 16, static_cast<sx_aln_seg2>(SX_SEG_HARDCLIP << 12));
            for (size_t i(0); i < _segs.size(); ++i) _segs2Out[i] = static_cast<sx_aln_seg2>(_segs[i].len | (_segs[i].kind << 12) | ((_segs[i].flags & 1) << 15));
            b.segs = reinterpret_cast<const sx_aln_seg*>(_segs2Out.data());
            b.format |= SX_FMT_SEG2;
        }
    }

    bool _open = false;
    bool _compact = true;
    std::vector<sx_aln8> _alns8Out;
    std::vector<uint32_t> _excOff, _exc;
    std::vector<sx_aln_seg2> _segs2Out;
    std::vector<sx_region> _regions, _regionsOut;
    std::vector<uint16_t> _readLen;
    std::vector<uint8_t> _seq4, _qual, _seq4Out, _qualOut;
    std::vector<char> _ref, _ins, _refOut, _insOut;
    std::vector<sx_aln> _alns, _alnsOut;
    std::vector<sx_aln_seg> _segs, _segsOut;
};

// ---------------------------------------------------------------------------------------------------------------------------
// K6: score_indels over a batch of reads (starling_read_align_score_indels.cpp:454-1079 + starling_read_align.cpp:1573-1593)
// ---------------------------------------------------------------------------------------------------------------------------
struct IndelBufferEntry ///< what score_indels reads of one IndelBuffer entry
{
    IndelKey key;
    bool isCandidate = true;            ///< indelBuffer.isCandidateIndel(key)
    double refToIndelLogProb = 0;       ///< getErrorRates().refToIndelErrorProb.getLogValue()
    double indelToRefLogProb = 0;       ///< getErrorRates().indelToRefErrorProb.getLogValue()
};

struct ReadPathScores ///< starling_common/IndelData.hh:64-116
{
    float ref = 0, indel = 0;
    uint16_t nonAmbiguousBasesInRead = 0, read_length = 0;
    bool is_tier1_read = true, is_fwd_strand = true;
    int16_t read_pos = 0, distanceFromClosestReadEdge = 0;
    std::vector<std::pair<IndelKey, float>> alt_indel;
};

class IndelScoreBatch
{
public:
    struct Result
    {
        unsigned read;        ///< index returned by addRead
        IndelKey key;         ///< the evaluated indel
        bool isSuboverlap;    ///< true: only suboverlap_tier{1,2}_read_ids.insert(read); false: read_path_lnp[read] = scores
        ReadPathScores scores;
    };

    /// window: every IndelBuffer entry a rangeIterator() over any alignment of the region's reads can visit, in IndelKey order
    void beginRegion(const std::vector<IndelBufferEntry>& window)
    {
        closeRegion();
        for (size_t i(1); i < window.size(); ++i) require(window[i - 1].key < window[i].key, "window is not in IndelKey order");
        require(window.size() <= 65535, "more than 65535 window entries");
        _regionKeyOff.push_back(_keys.size());
        _regionReadOff.push_back(_readLen.size());
        std::vector<std::string> interned(1, "");
        for (const IndelBufferEntry& e : window)
        {
            require(e.key.type == INDEL::INDEL || e.key.type == INDEL::MISMATCH, "breakend entries are not supported");
            sx_indel_key k;
            std::memset(&k, 0, sizeof(k));
            k.pos = e.key.pos;
            k.del_len = e.key.delete_length();
            k.ins_len = e.key.insert_length();
            size_t id(0);
            for (; id < interned.size(); ++id)
                if (interned[id] == e.key.insertSequence) break;
            if (id == interned.size()) interned.push_back(e.key.insertSequence);
            k.ins_id = id;
            k.type = e.key.isMismatch() ? SX_INDEL_TYPE_MISMATCH : SX_INDEL_TYPE_INDEL;
            k.flags = e.isCandidate ? SX_IKF_CANDIDATE : 0;
            k.ref_to_indel_lnp = e.refToIndelLogProb;
            k.indel_to_ref_lnp = e.indelToRefLogProb;
            _keys.push_back(k);
            _keyObjects.push_back(e.key);
        }
        _open = true;
    }

    /// a tier1 or tier2 read segment (rseg.read_size(), its non-'N' base count, strand of its candidate alignments, tier, incomplete search)
    unsigned addRead(unsigned readSize, unsigned nonAmbiguousBases, bool isFwdStrand, bool isTier1, bool isIncompleteSearch = false)
    {
        require(_open, "addRead outside a region");
        _alnOff.push_back(_alnPos.size());
        _recOff.push_back(_slots);
        _slots += _keys.size() - _regionKeyOff.back();
        _readLen.push_back(readSize);
        _nonAmbig.push_back(nonAmbiguousBases);
        _readFlags.push_back((isFwdStrand ? SX_SIF_FWD : 0) | (isTier1 ? SX_SIF_TIER1 : 0) | (isIncompleteSearch ? SX_SIF_INCOMPLETE : 0));
        return _readLen.size() - 1;
    }

    /// the next candidate alignment of the last added read, in std::set<CandidateAlignment> order (== the order their scores come in)
    void addCandidateAlignment(const CandidateAlignment& cal)
    {
        using namespace ALIGNPATH;
        require(_open && !_readLen.empty() && _regionReadOff.back() < _readLen.size(), "addCandidateAlignment before addRead");
        _alnPos.push_back(cal.al.pos);
        _alnSegOff.push_back(_segs.size());
        _alnKeyOff.push_back(_alnKeys.size());
        for (const path_segment& ps : cal.al.path)
        {
            uint8_t kind(0);
            switch (ps.type)
            {
            case MATCH: case SEQ_MATCH: case SEQ_MISMATCH: kind = SX_SEG_MATCH; break;
            case INSERT: kind = SX_SEG_INSERT; break;
            case DELETE: kind = SX_SEG_DELETE; break;
            case SOFT_CLIP: kind = SX_SEG_SOFTCLIP; break;
            case HARD_CLIP: kind = SX_SEG_HARDCLIP; break;
            default: throw Exception(SX_ERR_UNSUPPORTED, "score_indels: segment type outside its domain (get_alignment_indel_bp_overlap asserts)");
            }
            require(ps.length <= 65535, "path segment longer than 65535");
            _segs.push_back(sx_aln_seg{static_cast<uint16_t>(ps.length), kind, 0});
        }
        const size_t k0(_regionKeyOff.back());
        std::vector<uint16_t> idx;
        for (const IndelKey& key : cal.indels)
        {
            size_t lo(k0), hi(_keyObjects.size()); // binary search in the region's window
            while (lo < hi)
            {
                const size_t mid((lo + hi) / 2);
                if (_keyObjects[mid] < key) lo = mid + 1;
                else hi = mid;
            }
            require(lo < _keyObjects.size() && _keyObjects[lo] == key, "an alignment's indel is not in the region's window");
            idx.push_back(lo - k0);
        }
        std::sort(idx.begin(), idx.end());
        _alnKeys.insert(_alnKeys.end(), idx.begin(), idx.end());
    }

    size_t alignmentCount() const { return _alnPos.size(); }

    /// candAlignmentScores[alignmentCount()]: the K1 scores in the order the alignments were added.
    /// maxAlignment[read] = maxCandAlignmentPtr of scoreCandidateAlignments as an alignment index (UINT32_MAX: read without alignments)
    void scoreIndels(const Context& ctx, const std::vector<double>& candAlignmentScores, const sx_score_indels_opts* opts, std::vector<Result>& results,
                     std::vector<uint32_t>& maxAlignment)
    {
        closeRegion();
        require(candAlignmentScores.size() == _alnPos.size(), "one score per candidate alignment");
        results.clear();
        const size_t nReads(_readLen.size());
        maxAlignment.assign(nReads, UINT32_MAX);
        if (nReads == 0) return;
        std::vector<uint32_t> alnOff(_alnOff), recOff(_recOff), alnSegOff(_alnSegOff), alnKeyOff(_alnKeyOff), regionReadOff(_regionReadOff), regionKeyOff(_regionKeyOff);
        alnOff.push_back(_alnPos.size());
        recOff.push_back(_slots);
        alnSegOff.push_back(_segs.size());
        alnKeyOff.push_back(_alnKeys.size());
        regionReadOff.push_back(nReads);
        regionKeyOff.push_back(_keys.size());
        std::vector<sx_indel_key> keys(_keys);
        keys.resize(keys.size() + 1);
        std::vector<sx_aln_seg> segs(_segs);
        segs.resize(segs.size() + 4);
        std::vector<uint16_t> alnKeys(_alnKeys);
        alnKeys.resize(alnKeys.size() + 4);
        std::vector<int32_t> alnPos(_alnPos);
        alnPos.push_back(0);
        std::vector<double> lnp(candAlignmentScores);
        lnp.push_back(0);
        sx_score_indels_batch b;
        std::memset(&b, 0, sizeof(b));
        b.n_regions = regionReadOff.size() - 1;
        b.n_reads = nReads;
        b.n_alns = _alnPos.size();
        b.n_keys = _keys.size();
        b.region_read_off = regionReadOff.data();
        b.region_key_off = regionKeyOff.data();
        b.keys = keys.data();
        b.aln_off = alnOff.data();
        b.aln_pos = alnPos.data();
        b.aln_seg_off = alnSegOff.data();
        b.segs = segs.data();
        b.aln_key_off = alnKeyOff.data();
        b.aln_keys = alnKeys.data();
        b.read_len = _readLen.data();
        b.non_ambig = _nonAmbig.data();
        b.read_flags = _readFlags.data();
        b.rec_off = recOff.data();
        if (opts) b.opts = *opts;
        else sx_default_score_indels_opts(&b.opts);
        std::vector<sx_read_indel_score> recs(_slots + 1);
        std::vector<uint32_t> nRec(nReads), evalAln(nReads);
        sx_score_indels_out out{recs.data(), nRec.data(), maxAlignment.data(), evalAln.data()};
        ctx.check(sx_score_indels(ctx.get(), &b, lnp.data(), &out));
        for (size_t g(0); g + 1 < regionReadOff.size(); ++g)
            for (size_t r(regionReadOff[g]); r < regionReadOff[g + 1]; ++r)
                for (uint32_t j(0); j < nRec[r]; ++j)
                {
                    const sx_read_indel_score& q(recs[recOff[r] + j]);
                    Result res;
                    res.read = r;
                    res.key = _keyObjects[regionKeyOff[g] + q.key];
                    res.isSuboverlap = (q.flags & SX_RIS_SUBOVERLAP) != 0;
                    if (q.flags & SX_RIS_SCORED)
                    {
                        ReadPathScores& s(res.scores);
                        s.ref = q.ref_lnp;
                        s.indel = q.indel_lnp;
                        s.nonAmbiguousBasesInRead = _nonAmbig[r];
                        s.read_length = _readLen[r];
                        s.is_tier1_read = (_readFlags[r] & SX_SIF_TIER1) != 0;
                        s.is_fwd_strand = (_readFlags[r] & SX_SIF_FWD) != 0;
                        s.read_pos = q.read_pos;
                        s.distanceFromClosestReadEdge = q.dist_from_edge;
                        for (unsigned i(0); i < q.n_alt; ++i) s.alt_indel.push_back(std::make_pair(_keyObjects[regionKeyOff[g] + q.alt_key[i]], q.alt_lnp[i]));
                    }
                    results.push_back(res);
                }
    }

private:
    static void require(bool ok, const char* what)
    {
        if (!ok) throw Exception(SX_ERR_ARG, std::string("IndelScoreBatch: ") + what);
    }
    void closeRegion() { _open = false; }

    bool _open = false;
    uint32_t _slots = 0;
    std::vector<uint32_t> _regionReadOff, _regionKeyOff, _alnOff, _alnSegOff, _alnKeyOff, _recOff;
    std::vector<sx_indel_key> _keys;
    std::vector<IndelKey> _keyObjects;
    std::vector<int32_t> _alnPos;
    std::vector<sx_aln_seg> _segs;
    std::vector<uint16_t> _alnKeys, _readLen, _nonAmbig;
    std::vector<uint8_t> _readFlags;
};

// ---------------------------------------------------------------------------------------------------------------------------
// K7: getCandidateAlignments over a batch of reads (starling_read_align.cpp:1816-1994 + candidate_alignment_search :857-1277)
// ---------------------------------------------------------------------------------------------------------------------------
struct IndelSearchEntry ///< what the alignment search reads of one IndelBuffer entry (IndelData.hh:270-273,374,382,387)
{
    IndelKey key;
    bool isCandidate = true;             ///< indelBuffer.isCandidateIndel(key, data)
    bool notDiscoveredFromReads = false; ///< data.status.notDiscoveredFromReads
    bool isForcedOutput = false;         ///< data.isForcedOutput
    int activeRegionId = -1;             ///< data.activeRegionId
    int8_t haplotypeId[SX_ENUM_MAX_SAMPLES] = {0, 0, 0, 0};             ///< data.getSampleData(s).haplotypeId
    bool isHaplotypingBypassed[SX_ENUM_MAX_SAMPLES] = {false, false, false, false};
};

class AlignmentSearchBatch
{
public:
    struct ReadResult
    {
        std::vector<CandidateAlignment> alignments; ///< the std::set<CandidateAlignment>, in its iteration order
        bool originSkip = false, maxToggleDepth = false; ///< mca_warnings; either one = is_incomplete_search (:2100)
        bool threw = false;      ///< the reference throws blt_exception for this read
        bool overLimit = false;  ///< outside this build's per-read capacity: run the reference's own getCandidateAlignments on it
    };

    /// window: every IndelBuffer entry a rangeIterator() over any candidate alignment of the region's reads can visit, IndelKey order;
    /// ref / refBegin: the reference bases getAlignmentIndels compares read bases with (reference_contig_segment: 'N' outside)
    void beginRegion(const std::vector<IndelSearchEntry>& window, const std::string& ref, pos_t refBegin, pos_t realignBegin, pos_t realignEnd)
    {
        for (size_t i(1); i < window.size(); ++i) require(window[i - 1].key < window[i].key, "window is not in IndelKey order");
        require(window.size() < 65535, "more than 65534 window entries");
        _regionKeyOff.push_back(_keys.size());
        _regionReadOff.push_back(_readLen.size());
        _realignBegin.push_back(realignBegin);
        _realignEnd.push_back(realignEnd);
        _ref = ref;
        _refBegin = refBegin;
        for (const IndelSearchEntry& e : window)
        {
            require(e.key.type == INDEL::INDEL || e.key.type == INDEL::MISMATCH, "breakend entries are not supported");
            sx_indel_key k;
            std::memset(&k, 0, sizeof(k));
            k.pos = e.key.pos;
            k.del_len = e.key.delete_length();
            k.ins_len = e.key.insert_length();
            k.type = e.key.isMismatch() ? SX_INDEL_TYPE_MISMATCH : SX_INDEL_TYPE_INDEL;
            k.flags = (e.isCandidate ? SX_IKF_CANDIDATE : 0) | (e.notDiscoveredFromReads ? SX_IKF_NOT_DISCOVERED : 0) | (e.isForcedOutput ? SX_IKF_FORCED_OUTPUT : 0);
            _keys.push_back(k);
            sx_key_hap h;
            std::memset(&h, 0, sizeof(h));
            h.active_region_id = e.activeRegionId;
            for (unsigned s(0); s < SX_ENUM_MAX_SAMPLES; ++s)
            {
                h.haplotype_id[s] = e.haplotypeId[s];
                if (e.isHaplotypingBypassed[s]) h.bypass_mask |= (1u << s);
            }
            _hap.push_back(h);
            _anyHap = _anyHap || e.activeRegionId >= 0;
            _keyObjects.push_back(e.key);
        }
        _open = true;
    }

    /// one read segment: its bases (ASCII), the NORMALIZED input alignment realignAndScoreRead passes on (:2049-2057) and the
    /// window entries whose tier1 / tier2 / submap / noise read-id sets hold this read (is_usable_indel :289-305).
    /// Throws what the reference throws when the alignment holds an indel the window lacks (:1866-1872).
    unsigned addRead(const std::string& readBases, const alignment& normalizedInputAlignment, const std::vector<IndelKey>& observedKeys, unsigned maxIndelSize = 49)
    {
        using namespace ALIGNPATH;
        require(_open, "addRead outside a region");
        const path_t& path(normalizedInputAlignment.path);
        _inPos.push_back(normalizedInputAlignment.pos);
        _inSegOff.push_back(_inSegs.size());
        _inKeyOff.push_back(_inKeys.size());
        _useKeyOff.push_back(_useKeys.size());
        for (const path_segment& ps : path)
        {
            require(ps.length <= 65535 && ps.type != NONE, "bad path segment");
            _inSegs.push_back(sx_aln_seg{static_cast<uint16_t>(ps.length), static_cast<uint8_t>(ps.type), 0});
        }
        // getAlignmentIndels(cal, ref, rseg, maxIndelSize, includeMismatches = true), CandidateAlignment.cpp:58-173, and the edge keys
        // of getCandidateAlignment, starling_read_align.cpp:1481-1522
        size_t first(path.size()), last(path.size());
        for (size_t i(0); i < path.size(); ++i)
            if (path[i].type == MATCH || path[i].type == SEQ_MATCH || path[i].type == SEQ_MISMATCH)
            {
                if (first == path.size()) first = i;
                last = i;
            }
        std::vector<uint16_t> keys;
        uint16_t lead(SX_NO_KEY), trail(SX_NO_KEY);
        bool hasLead(false), hasTrail(false);
        unsigned readOff(0);
        pos_t refPos(normalizedInputAlignment.pos);
        for (size_t i(0); i < path.size();)
        {
            const path_segment& ps(path[i]);
            const bool edge(i < first || i > last);
            // a swap = a run of adjacent insert / delete segments holding BOTH kinds (is_segment_swap_start, align_path.cpp:868-895)
            size_t j(i);
            unsigned insLen(0), delLen(0);
            for (; j < path.size() && (path[j].type == INSERT || path[j].type == DELETE); ++j) (path[j].type == INSERT ? insLen : delLen) += path[j].length;
            const bool isSwap(insLen && delLen);
            size_t step(1);
            if (edge)
            {
                if (ps.type == INSERT || ps.type == DELETE) // the edge key is set anew for every edge indel segment (:1495-1518); it is inserted once
                {
                    const IndelKey k(refPos, INDEL::INDEL, ps.type == DELETE ? ps.length : 0, ps.type == INSERT ? readBases.substr(readOff, ps.length).c_str() : "");
                    const uint16_t w(indexOf(k, true));
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
            else if (isSwap)
            {
                step = j - i;
                if (std::max(insLen, delLen) > maxIndelSize) throw Exception(SX_ERR_UNSUPPORTED, "AlignmentSearchBatch: indel above maxIndelSize (breakend keys are not supported)");
                keys.push_back(indexOf(IndelKey(refPos, INDEL::INDEL, delLen, readBases.substr(readOff, insLen).c_str()), true));
            }
            else if (ps.type == INSERT || ps.type == DELETE)
            {
                if (ps.length > maxIndelSize) throw Exception(SX_ERR_UNSUPPORTED, "AlignmentSearchBatch: indel above maxIndelSize (breakend keys are not supported)");
                keys.push_back(indexOf(IndelKey(refPos, INDEL::INDEL, ps.type == DELETE ? ps.length : 0, ps.type == INSERT ? readBases.substr(readOff, ps.length).c_str() : ""), true));
            }
            else if (ps.type == MATCH || ps.type == SEQ_MATCH || ps.type == SEQ_MISMATCH)
            {
                for (unsigned b(0); b < ps.length; ++b)
                {
                    char base(readBases[readOff + b]);
                    if (base == '=' || base == 'N') continue; // BAM_BASE::REF, BAM_BASE::ANY
                    if (base != 'A' && base != 'C' && base != 'G' && base != 'T') base = 'N'; // every other code reads back as 'N' and never equals the reference code
                    const pos_t rp(refPos + static_cast<pos_t>(b));
                    const char refBase((rp >= _refBegin && rp < _refBegin + static_cast<pos_t>(_ref.size())) ? _ref[rp - _refBegin] : 'N');
                    if (base == refBase) continue;
                    const char ins[2] = {base, 0};
                    const uint16_t w(indexOf(IndelKey(rp, INDEL::MISMATCH, 1, ins), false));
                    if (w != SX_NO_KEY) keys.push_back(w); // a mismatch that is no window entry is dropped (:1865)
                }
            }
            for (size_t s(i); s < i + step; ++s)
            {
                const ALIGNPATH::align_t t(path[s].type);
                if (t == MATCH || t == INSERT || t == SOFT_CLIP || t == SEQ_MATCH || t == SEQ_MISMATCH) readOff += path[s].length;
                if (t == MATCH || t == DELETE || t == SKIP || t == SEQ_MATCH || t == SEQ_MISMATCH) refPos += path[s].length;
            }
            i += step;
        }
        if (hasLead) keys.push_back(lead);
        if (hasTrail) keys.push_back(trail);
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        _inKeys.insert(_inKeys.end(), keys.begin(), keys.end());
        std::vector<uint16_t> use;
        for (const IndelKey& k : observedKeys)
        {
            const uint16_t w(indexOf(k, false));
            if (w != SX_NO_KEY) use.push_back(w);
        }
        std::sort(use.begin(), use.end());
        use.erase(std::unique(use.begin(), use.end()), use.end());
        _useKeys.insert(_useKeys.end(), use.begin(), use.end());
        _lead.push_back(lead);
        _trail.push_back(trail);
        _readLen.push_back(readBases.size());
        _fwd.push_back(normalizedInputAlignment.is_fwd_strand);
        return _readLen.size() - 1;
    }

    /// results[read]: what getCandidateAlignments would have put into cal_set (and the warn flags)
    void enumerate(const Context& ctx, const sx_enum_opts* opts, std::vector<ReadResult>& results)
    {
        _open = false;
        const size_t nReads(_readLen.size());
        results.assign(nReads, ReadResult());
        if (nReads == 0) return;
        std::vector<uint32_t> regionReadOff(_regionReadOff), regionKeyOff(_regionKeyOff), inSegOff(_inSegOff), inKeyOff(_inKeyOff), useKeyOff(_useKeyOff);
        regionReadOff.push_back(nReads);
        regionKeyOff.push_back(_keys.size());
        inSegOff.push_back(_inSegs.size());
        inKeyOff.push_back(_inKeys.size());
        useKeyOff.push_back(_useKeys.size());
        std::vector<sx_indel_key> keys(_keys);
        keys.resize(keys.size() + 1);
        std::vector<sx_key_hap> hap(_hap);
        hap.resize(hap.size() + 1);
        std::vector<sx_aln_seg> inSegs(_inSegs);
        inSegs.resize(inSegs.size() + 4);
        std::vector<uint16_t> inKeys(_inKeys), useKeys(_useKeys);
        inKeys.resize(inKeys.size() + 4);
        useKeys.resize(useKeys.size() + 4);
        sx_enum_batch b;
        std::memset(&b, 0, sizeof(b));
        b.n_regions = regionReadOff.size() - 1;
        b.n_reads = nReads;
        b.n_keys = _keys.size();
        b.region_read_off = regionReadOff.data();
        b.region_key_off = regionKeyOff.data();
        b.keys = keys.data();
        b.key_hap = _anyHap ? hap.data() : nullptr;
        b.realign_begin = _realignBegin.data();
        b.realign_end = _realignEnd.data();
        b.in_pos = _inPos.data();
        b.in_seg_off = inSegOff.data();
        b.in_segs = inSegs.data();
        b.in_key_off = inKeyOff.data();
        b.in_keys = inKeys.data();
        b.use_key_off = useKeyOff.data();
        b.use_keys = useKeys.data();
        b.in_lead_key = _lead.data();
        b.in_trail_key = _trail.data();
        b.read_len = _readLen.data();
        if (opts) b.opts = *opts;
        else sx_default_enum_opts(&b.opts);
        const uint32_t perRead(b.opts.max_alns_per_read ? b.opts.max_alns_per_read : 64u);
        sx_enum_out o;
        std::memset(&o, 0, sizeof(o));
        uint32_t totals[3] = {0, 0, 0};
        std::vector<uint32_t> alnOff(nReads + 1), alnSegOff, alnKeyOff;
        std::vector<uint8_t> status(nReads);
        std::vector<int32_t> alnPos;
        std::vector<sx_aln_seg> segs;
        std::vector<uint16_t> alnKeys, alnLead, alnTrail;
        // first guess: a few alignments per read; on SX_ERR_CAPACITY the library says what is needed
        o.cap_alns = std::min<uint64_t>(nReads * std::min<uint32_t>(perRead, 16u), 0x7fffffffu);
        o.cap_segs = o.cap_alns * 8;
        o.cap_keys = o.cap_alns * 4;
        for (int attempt(0); attempt < 2; ++attempt)
        {
            alnPos.resize(o.cap_alns + 1);
            alnSegOff.resize(o.cap_alns + 2);
            alnKeyOff.resize(o.cap_alns + 2);
            segs.resize(o.cap_segs + 1);
            alnKeys.resize(o.cap_keys + 1);
            alnLead.resize(o.cap_alns + 1);
            alnTrail.resize(o.cap_alns + 1);
            o.totals = totals;
            o.aln_off = alnOff.data();
            o.status = status.data();
            o.aln_pos = alnPos.data();
            o.aln_seg_off = alnSegOff.data();
            o.segs = segs.data();
            o.aln_key_off = alnKeyOff.data();
            o.aln_keys = alnKeys.data();
            o.aln_lead_key = alnLead.data();
            o.aln_trail_key = alnTrail.data();
            const int rc(sx_enumerate_alignments(ctx.get(), &b, &o));
            if (rc == SX_ERR_CAPACITY && attempt == 0)
            {
                o.cap_alns = totals[0];
                o.cap_segs = totals[1];
                o.cap_keys = totals[2];
                continue;
            }
            ctx.check(rc);
            break;
        }
        // keep the enumeration as the library wrote it: chooseRealignments() (K9) reads the same arrays
        _eRegionReadOff = regionReadOff;
        _eRegionKeyOff = regionKeyOff;
        _eAlnOff = alnOff;
        _eAlnPos.assign(alnPos.begin(), alnPos.begin() + totals[0]);
        _eAlnSegOff.assign(alnSegOff.begin(), alnSegOff.begin() + totals[0] + 1);
        _eSegs.assign(segs.begin(), segs.begin() + totals[1]);
        _eAlnKeyOff.assign(alnKeyOff.begin(), alnKeyOff.begin() + totals[0] + 1);
        _eAlnKeys.assign(alnKeys.begin(), alnKeys.begin() + totals[2]);
        for (size_t g(0); g + 1 < regionReadOff.size(); ++g)
            for (size_t r(regionReadOff[g]); r < regionReadOff[g + 1]; ++r)
            {
                ReadResult& res(results[r]);
                res.originSkip = (status[r] & SX_ENUM_ST_ORIGIN_SKIP) != 0;
                res.maxToggleDepth = (status[r] & SX_ENUM_ST_MAX_TOGGLE) != 0;
                res.threw = (status[r] & SX_ENUM_ST_EXCEPTION) != 0;
                res.overLimit = (status[r] & SX_ENUM_ST_LIMIT) != 0;
                for (uint32_t a(alnOff[r]); a < alnOff[r + 1]; ++a)
                {
                    CandidateAlignment cal;
                    cal.al.pos = alnPos[a];
                    cal.al.is_fwd_strand = _fwd[r];
                    for (uint32_t s(alnSegOff[a]); s < alnSegOff[a + 1]; ++s) cal.al.path.push_back(path_segment(static_cast<ALIGNPATH::align_t>(segs[s].kind), segs[s].len));
                    for (uint32_t k(alnKeyOff[a]); k < alnKeyOff[a + 1]; ++k) cal.indels.push_back(_keyObjects[regionKeyOff[g] + alnKeys[k]]);
                    if (alnLead[a] != SX_NO_KEY) cal.leading_indel_key = _keyObjects[regionKeyOff[g] + alnLead[a]];
                    if (alnTrail[a] != SX_NO_KEY) cal.trailing_indel_key = _keyObjects[regionKeyOff[g] + alnTrail[a]];
                    res.alignments.push_back(cal);
                }
            }
    }

    struct Realignment ///< what scoreCandidateAlignments leaves in the read segment (:1739-1740)
    {
        bool is_realigned = false;
        alignment realignment;
        uint32_t bestAlignment = UINT32_MAX; ///< smooth_cal_ptr as an index into the read's alignments of enumerate()
        bool unsupported = false;            ///< pinned read / malformed path: stays with the caller
    };

    /// K9, after enumerate(): candAlignmentScores = the K1 scores of every alignment in result order (read by read, alignment by
    /// alignment).  The tail of scoreCandidateAlignments (:1573-1741) for unpinned reads: smooth pool, preferred alignment,
    /// finishRealignment with the pool clipper.
    void chooseRealignments(const Context& ctx, const std::vector<double>& candAlignmentScores, bool isSmoothedAlignments, double smoothedLnpRange,
                            std::vector<Realignment>& out)
    {
        const size_t nReads(_readLen.size()), nAlns(_eAlnPos.size());
        require(candAlignmentScores.size() == nAlns, "one score per enumerated alignment");
        out.assign(nReads, Realignment());
        if (nReads == 0) return;
        std::vector<sx_indel_key> keys(_keys);
        keys.resize(keys.size() + 1);
        std::vector<int32_t> alnPos(_eAlnPos);
        alnPos.push_back(0);
        std::vector<sx_aln_seg> segs(_eSegs);
        segs.resize(segs.size() + 4);
        std::vector<uint16_t> alnKeys(_eAlnKeys);
        alnKeys.resize(alnKeys.size() + 4);
        std::vector<double> lnp(candAlignmentScores);
        lnp.push_back(0);
        sx_realign_batch b;
        std::memset(&b, 0, sizeof(b));
        b.n_regions = _eRegionReadOff.size() - 1;
        b.n_reads = nReads;
        b.n_alns = nAlns;
        b.region_read_off = _eRegionReadOff.data();
        b.region_key_off = _eRegionKeyOff.data();
        b.keys = keys.data();
        b.aln_off = _eAlnOff.data();
        b.aln_pos = alnPos.data();
        b.aln_seg_off = _eAlnSegOff.data();
        b.segs = segs.data();
        b.aln_key_off = _eAlnKeyOff.data();
        b.aln_keys = alnKeys.data();
        b.read_len = _readLen.data();
        b.is_smoothed_alignments = isSmoothedAlignments ? 1 : 0;
        b.smoothed_lnp_range = smoothedLnpRange;
        sx_realign_out o;
        std::memset(&o, 0, sizeof(o));
        uint32_t total(0);
        std::vector<uint32_t> segOff(nReads + 1), best(nReads);
        std::vector<int32_t> pos(nReads);
        std::vector<uint16_t> nSeg(nReads);
        std::vector<uint8_t> status(nReads);
        o.cap_segs = _eSegs.size() + 2 * nReads + 64;
        std::vector<sx_aln_seg> outSegs(o.cap_segs + 1);
        o.totals = &total;
        o.seg_off = segOff.data();
        o.pos = pos.data();
        o.n_seg = nSeg.data();
        o.status = status.data();
        o.best_aln = best.data();
        o.segs = outSegs.data();
        ctx.check(sx_choose_realignment(ctx.get(), &b, lnp.data(), &o));
        for (size_t r(0); r < nReads; ++r)
        {
            Realignment& res(out[r]);
            res.unsupported = (status[r] & (SX_REALIGN_ST_UNSUPPORTED | SX_REALIGN_ST_LIMIT | SX_REALIGN_ST_BADPATH)) != 0;
            if (!(status[r] & SX_REALIGN_ST_REALIGNED)) continue;
            res.is_realigned = true;
            res.realignment.pos = pos[r];
            res.realignment.is_fwd_strand = _fwd[r];
            for (uint32_t i(0); i < nSeg[r]; ++i) res.realignment.path.push_back(path_segment(static_cast<ALIGNPATH::align_t>(outSegs[segOff[r] + i].kind), outSegs[segOff[r] + i].len));
            res.bestAlignment = best[r] - _eAlnOff[r];
        }
    }

private:
    static void require(bool ok, const char* what)
    {
        if (!ok) throw Exception(SX_ERR_ARG, std::string("AlignmentSearchBatch: ") + what);
    }
    /// window index of `key` in the open region; absent: SX_NO_KEY, or (must = true) the reference's exception
    uint16_t indexOf(const IndelKey& key, bool must) const
    {
        size_t lo(_regionKeyOff.back()), hi(_keyObjects.size());
        const size_t k0(lo);
        while (lo < hi)
        {
            const size_t mid((lo + hi) / 2);
            if (_keyObjects[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        if (lo < _keyObjects.size() && _keyObjects[lo] == key) return static_cast<uint16_t>(lo - k0);
        if (must) throw Exception(SX_ERR_ARG, "Exemplar alignment contains indel not found in the overlap indel set"); // starling_read_align.cpp:1866-1872
        return SX_NO_KEY;
    }

    bool _open = false, _anyHap = false;
    std::string _ref;
    pos_t _refBegin = 0;
    std::vector<uint32_t> _regionReadOff, _regionKeyOff, _inSegOff, _inKeyOff, _useKeyOff;
    std::vector<int32_t> _realignBegin, _realignEnd, _inPos;
    std::vector<sx_indel_key> _keys;
    std::vector<sx_key_hap> _hap;
    std::vector<IndelKey> _keyObjects;
    std::vector<sx_aln_seg> _inSegs;
    std::vector<uint16_t> _inKeys, _useKeys, _lead, _trail, _readLen;
    std::vector<bool> _fwd;
    // the last enumeration, as the library wrote it
    std::vector<uint32_t> _eRegionReadOff, _eRegionKeyOff, _eAlnOff, _eAlnSegOff, _eAlnKeyOff;
    std::vector<int32_t> _eAlnPos;
    std::vector<sx_aln_seg> _eSegs;
    std::vector<uint16_t> _eAlnKeys;
};

// ---------------------------------------------------------------------------------------------------------------------------
// K3: GlobalAligner
// ---------------------------------------------------------------------------------------------------------------------------
template <typename ScoreType> struct AlignmentScores // alignment/AlignmentScores.hh:24-53
{
    AlignmentScores(ScoreType initMatch, ScoreType initMismatch, ScoreType initOpen, ScoreType initExtend, ScoreType initOffEdge, ScoreType initInsertDelete = 0,
                    bool initIsAllowEdgeInsertion = false, bool initIsRequireEdgeDeletion = false)
        : match(initMatch), mismatch(initMismatch), open(initOpen), extend(initExtend), offEdge(initOffEdge), insertDelete(initInsertDelete),
          isAllowEdgeInsertion(initIsAllowEdgeInsertion), isRequireEdgeDeletion(initIsRequireEdgeDeletion)
    {
    }
    const ScoreType match, mismatch, open, extend, offEdge, insertDelete;
    const bool isAllowEdgeInsertion, isRequireEdgeDeletion;
};

struct Alignment // alignment/Alignment.hh:30-52
{
    pos_t beginPos = 0;
    path_t apath;
};
template <typename ScoreType> struct AlignmentResult // alignment/SingleRefAlignerShared.hh:33-50
{
    ScoreType score = 0;
    Alignment align;
};

template <typename ScoreType> class GlobalAligner
{
public:
    GlobalAligner(const Context& ctx, const AlignmentScores<ScoreType>& scores) : _ctx(ctx)
    {
        _sc.match = scores.match;
        _sc.mismatch = scores.mismatch;
        _sc.open = scores.open;
        _sc.extend = scores.extend;
        _sc.offEdge = scores.offEdge;
        _sc.insertDelete = scores.insertDelete;
        _sc.isAllowEdgeInsertion = scores.isAllowEdgeInsertion;
        _sc.isRequireEdgeDeletion = scores.isRequireEdgeDeletion;
    }

    /// GlobalAligner<ScoreType>::align(queryBegin, queryEnd, refBegin, refEnd, result)  alignment/GlobalAlignerImpl.hh:36
    template <typename SymIter> void align(SymIter queryBegin, SymIter queryEnd, SymIter refBegin, SymIter refEnd, AlignmentResult<ScoreType>& result) const
    {
        std::vector<std::pair<std::string, std::string>> one(1, std::make_pair(std::string(queryBegin, queryEnd), std::string(refBegin, refEnd)));
        std::vector<AlignmentResult<ScoreType>> res;
        alignBatch(one, res);
        result = res[0];
    }

    /// all (haplotype, reference segment) pairs of an active region in one launch
    void alignBatch(const std::vector<std::pair<std::string, std::string>>& queryRef, std::vector<AlignmentResult<ScoreType>>& results) const
    {
        const uint32_t n(queryRef.size());
        std::string q, r;
        std::vector<uint32_t> qo(n + 1, 0), ro(n + 1, 0);
        uint32_t maxOps(8);
        for (uint32_t i(0); i < n; ++i)
        {
            q += queryRef[i].first;
            r += queryRef[i].second;
            qo[i + 1] = q.size();
            ro[i + 1] = r.size();
            maxOps = std::max<uint32_t>(maxOps, queryRef[i].first.size() + queryRef[i].second.size() + 2);
        }
        q.resize(q.size() + 16);
        r.resize(r.size() + 16);
        sx_ga_batch b{n, q.data(), r.data(), qo.data(), ro.data(), maxOps};
        std::vector<sx_ga_result> res(n);
        std::vector<uint32_t> cig(static_cast<size_t>(n) * maxOps);
        _ctx.check(sx_global_align(_ctx.get(), &_sc, &b, res.data(), cig.data()));
        results.assign(n, AlignmentResult<ScoreType>());
        for (uint32_t i(0); i < n; ++i)
        {
            if (res[i].status != 0) throw Exception(SX_ERR_ARG, "sx_global_align: problem too large for the kernel's shared-memory tile");
            results[i].score = static_cast<ScoreType>(res[i].score);
            results[i].align.beginPos = res[i].beginPos;
            for (uint32_t k(0); k < res[i].n_ops; ++k)
            {
                const uint32_t op(cig[static_cast<size_t>(i) * maxOps + k]);
                results[i].align.apath.push_back(path_segment(cigar_code_to_segment_type("MIDNSHP=X"[op & 15]), op >> 4));
            }
        }
    }

private:
    const Context& _ctx;
    sx_ga_scores _sc;
};

} // namespace sx
