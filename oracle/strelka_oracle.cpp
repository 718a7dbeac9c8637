// strelka_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see strelka_oracle.h).
//
// Plain scalar C++ restatement of the reference algorithms, written to make the SAME libm / libstdc++ calls
// (std::log/std::pow on float -> logf/powf, std::sort, std::exp/std::log on double) in the SAME order as the
// reference, so that its results are bit-identical to the reference's on the same host.  Every function cites
// the reference lines it follows (paths relative to /root/reference/src/c++/lib/).
//
// Compile WITHOUT -ffast-math / -march=native (no FMA contraction: the reference's release build is plain
// x86-64 SSE2, src/cmake/cxxConfigure.cmake:438): see oracle/Makefile.

#include "strelka_oracle.h"
#include "flat_batch_normalize.h"

#include "../strelka_b200/csrc/sx_libm_mirror.h" // only for the ox_*_restated test exports below

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace
{

typedef float blt_float_t; // blt_util/blt_types.hh:27

// ------------------------------------------------------------------------------------------------
// qphred cache: blt_util/qscore_cache.cpp:33-49, blt_util/math_util.hh:33-47
// ------------------------------------------------------------------------------------------------
const int MAX_QSCORE = 70;

// boost::math::log1p<double> (boost 1.58 math/special_functions/log1p.hpp) resolves to ::log1p for double
// when BOOST_HAS_LOG1P is defined (glibc): the reference therefore calls libm's log1p here.
double log1p_switch(const double x)
{
    static const double smallx_thresh(0.01);
    if (std::abs(x) < smallx_thresh) return ::log1p(x);
    return std::log(1 + x);
}

struct qphred_cache
{
    qphred_cache()
    {
        static const double q2lnp(-std::log(10.) / 10.);
        for (int i(0); i <= MAX_QSCORE; ++i)
        {
            q2p[i] = std::pow(10., -static_cast<double>(i) / 10.);
            q2lncompe[i] = log1p_switch(-q2p[i]);
            q2lne[i] = static_cast<double>(i) * q2lnp;
        }
    }
    double q2p[MAX_QSCORE + 1], q2lncompe[MAX_QSCORE + 1], q2lne[MAX_QSCORE + 1];
};

const qphred_cache& qc()
{
    static const qphred_cache c;
    return c;
}

// ------------------------------------------------------------------------------------------------
// a1  scoreCandidateAlignment  starling_common/starling_read_align_score.cpp:108-170,260-499
// ------------------------------------------------------------------------------------------------
inline uint8_t bam_code_of_char(const char c) // htsapi/bam_seq.hh:98-118
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

} // namespace

extern "C" int ox_score_alignments_range(const sx_align_batch* b, uint32_t r0, uint32_t r1, double* lnp_out)
{
    sx_norm_batch wide; // compact wire formats are widened first (same values, wider fields)
    b = sx_normalize_range(b, r0, r1, wide);
    static const double lnthird(-std::log(3.));                           // score.cpp:118,152
    static const double unalignedBasecallLogLikelihood(std::log(0.25));   // score.cpp:453
    static const double nonCandidateIndelPenalty(std::log(1e-5));         // score.cpp:483
    const qphred_cache& q(qc());

    for (uint32_t ri = r0; ri < r1; ++ri)
    {
        const sx_region& reg(b->regions[ri]);
        const sx_region& nxt(b->regions[ri + 1]);
        // per-read offsets inside the region: reads are packed back to back
        std::vector<uint64_t> seqOff, qualOff;
        {
            uint64_t so(reg.seq_off), qo(reg.qual_off);
            for (uint32_t r = reg.read_begin; r < nxt.read_begin; ++r)
            {
                seqOff.push_back(so);
                qualOff.push_back(qo);
                so += (b->read_len[r] + 1) / 2;
                qo += (b->qual_bits == 4) ? (b->read_len[r] + 1) / 2 : b->read_len[r];
            }
        }
        for (uint32_t a = reg.aln_begin; a < nxt.aln_begin; ++a)
        {
            const sx_aln& al(b->alns[a]);
            const uint32_t rl(al.read - reg.read_begin);
            const uint8_t* seq(b->seq4 + seqOff[rl]);
            const uint8_t* qual(b->qual + qualOff[rl]);
            const char* ins(b->ins + al.ins_off);

            double alignmentLogProb(0.);
            unsigned read_offset(0);
            int32_t ref_head_pos(al.ref_pos);
            for (uint32_t s = al.seg_off; s < b->alns[a + 1].seg_off; ++s)
            {
                const sx_aln_seg& ps(b->segs[s]);
                if (ps.kind == SX_SEG_MATCH || ps.kind == SX_SEG_INSERT)
                {
                    // scoreMatchSegment / scoreInsertSegment (identical loops, different "ref" sequence)
                    for (unsigned i(0); i < ps.len; ++i)
                    {
                        const unsigned readPos(read_offset + i);
                        const uint8_t sbase((seq[readPos >> 1] >> ((~readPos & 1) << 2)) & 0xf); // bam_seq::get_code
                        if (sbase == 15) continue; // BAM_BASE::ANY
                        // qualities: one byte per base, or (qual_bits == 4) dictionary-coded nibbles, high nibble first
                        const uint8_t qscore((b->qual_bits == 4) ? b->qual_dict[(qual[readPos >> 1] >> ((~readPos & 1) << 2)) & 0xf] : qual[readPos]);
                        bool is_ref(sbase == 0); // BAM_BASE::REF
                        if (!is_ref)
                        {
                            uint8_t rcode;
                            if (ps.kind == SX_SEG_MATCH)
                            {
                                const int64_t rp((int64_t)ref_head_pos + i - reg.ref_begin);
                                // reference_contig_segment::get_base: 'N' outside the held segment
                                const char rc((rp < 0 || rp >= (int64_t)reg.ref_len) ? 'N' : b->ref[reg.ref_off + rp]);
                                rcode = bam_code_of_char(rc);
                            }
                            else
                            {
                                rcode = bam_code_of_char(ins[i]);
                            }
                            is_ref = (sbase == rcode);
                        }
                        alignmentLogProb += (is_ref ? q.q2lncompe[qscore] : q.q2lne[qscore] + lnthird);
                    }
                    read_offset += ps.len;
                    if (ps.kind == SX_SEG_MATCH) ref_head_pos += ps.len;
                    else ins += ps.len;
                }
                else if (ps.kind == SX_SEG_REFSKIP)
                {
                    ref_head_pos += ps.len;
                }
                else if (ps.kind == SX_SEG_SOFTCLIP)
                {
                    alignmentLogProb += (ps.len * unalignedBasecallLogLikelihood);
                    read_offset += ps.len;
                }
                else if (ps.kind == SX_SEG_HARDCLIP)
                {
                }
                else
                {
                    return SX_ERR_ARG;
                }
                if (ps.flags & SX_SEGF_NONCANDIDATE) alignmentLogProb += nonCandidateIndelPenalty;
            }
            lnp_out[a] = alignmentLogProb;
        }
    }
    return SX_OK;
}

extern "C" int ox_score_alignments(const sx_align_batch* b, double* lnp_out)
{
    return ox_score_alignments_range(b, 0, b->n_regions, lnp_out);
}

// ------------------------------------------------------------------------------------------------
// a5  GlobalAligner<int>::align   alignment/GlobalAlignerImpl.hh:36-228
//     backTraceAlignment          alignment/SingleRefAlignerSharedImpl.hh:80-170
//     apath_add_seqmatch          blt_util/align_path_impl.hh:36-86
// ------------------------------------------------------------------------------------------------
namespace
{
enum { ST_MATCH = 0, ST_DELETE = 1, ST_INSERT = 2 }; // AlignState, alignment/Alignment.hh

struct ScoreVal
{
    int match, del, ins;
};
struct PtrVal
{
    uint8_t match, del, ins;
    uint8_t get(int st) const { return st == ST_MATCH ? match : (st == ST_DELETE ? del : ins); }
};

inline uint8_t max3(int& max, const int v0, const int v1, const int v2) // alignment/AlignerBase.hh:71-92
{
    max = v0;
    uint8_t ptr = 0;
    if (v1 > v0)
    {
        max = v1;
        ptr = 1;
    }
    if (v2 > max)
    {
        max = v2;
        ptr = 2;
    }
    return ptr;
}

struct BackTrace // alignment/AlignerUtil.hh:47-80
{
    int max = 0;
    int state = ST_MATCH;
    unsigned queryBegin = 0, refBegin = 0;
    bool isInit = false;
};

inline void updateBacktrace(const int thisMax, const unsigned refIndex, const unsigned queryIndex, BackTrace& bt, const int state = ST_MATCH)
{
    if ((!bt.isInit) || (thisMax > bt.max))
    {
        bt.max = thisMax;
        bt.refBegin = refIndex;
        bt.queryBegin = queryIndex;
        bt.isInit = true;
        bt.state = state;
    }
}

enum { CIG_M = 0, CIG_I = 1, CIG_D = 2, CIG_S = 4, CIG_EQ = 7, CIG_X = 8 };

struct Seg
{
    int type;
    unsigned length;
};

void align_one(const sx_ga_scores& sc, const char* query, const unsigned querySize, const char* ref, const unsigned refSize,
               sx_ga_result& res, uint32_t* cigar, const uint32_t max_ops)
{
    static const int badVal(-10000);
    std::vector<ScoreVal> score1(querySize + 1), score2(querySize + 1);
    std::vector<PtrVal> ptrMat((size_t)(querySize + 1) * (refSize + 1));
    auto PM = [&](unsigned qi, unsigned ri) -> PtrVal& { return ptrMat[(size_t)qi * (refSize + 1) + ri]; };
    std::vector<ScoreVal>* thisSV(&score1);
    std::vector<ScoreVal>* prevSV(&score2);

    for (unsigned queryIndex(0); queryIndex <= querySize; queryIndex++)
    {
        PtrVal& headPtr(PM(queryIndex, 0));
        ScoreVal& val((*thisSV)[queryIndex]);
        headPtr.match = ST_MATCH;
        val.match = queryIndex * sc.offEdge;
        headPtr.del = ST_MATCH;
        val.del = badVal;
        if (!sc.isAllowEdgeInsertion)
        {
            headPtr.ins = ST_MATCH;
            val.ins = badVal;
        }
        else
        {
            headPtr.ins = ST_INSERT;
            val.ins = sc.open + (queryIndex * sc.extend);
        }
    }

    BackTrace btrace;
    for (unsigned refIndex(0); refIndex < refSize; ++refIndex)
    {
        std::swap(thisSV, prevSV);
        {
            PtrVal& headPtr(PM(0, refIndex + 1));
            ScoreVal& val((*thisSV)[0]);
            if (!sc.isRequireEdgeDeletion)
            {
                headPtr.match = ST_MATCH;
                val.match = 0;
                headPtr.del = ST_MATCH;
                val.del = badVal;
            }
            else
            {
                headPtr.match = ST_MATCH;
                val.match = badVal;
                headPtr.del = ST_DELETE;
                val.del = sc.open + ((refIndex + 1) * sc.extend);
            }
            headPtr.ins = ST_MATCH;
            val.ins = badVal;
        }
        for (unsigned queryIndex(0); queryIndex < querySize; ++queryIndex)
        {
            ScoreVal& headScore((*thisSV)[queryIndex + 1]);
            PtrVal& headPtr(PM(queryIndex + 1, refIndex + 1));
            {
                const ScoreVal& sval((*prevSV)[queryIndex]);
                headPtr.match = max3(headScore.match, sval.match, sval.del, sval.ins);
                headScore.match += ((query[queryIndex] == ref[refIndex]) ? sc.match : sc.mismatch);
            }
            {
                const ScoreVal& sval((*prevSV)[queryIndex + 1]);
                headPtr.del = max3(headScore.del, sval.match + sc.open, sval.del, sval.ins + sc.insertDelete);
                headScore.del += sc.extend;
                if (0 == refIndex) headScore.del = badVal;
            }
            {
                const ScoreVal& sval((*thisSV)[queryIndex]);
                headPtr.ins = max3(headScore.ins, sval.match + sc.open, badVal, sval.ins);
                headScore.ins += sc.extend;
                if (0 == queryIndex) headScore.ins = badVal;
            }
        }
        if (!sc.isRequireEdgeDeletion)
        {
            const ScoreVal& sval((*thisSV)[querySize]);
            updateBacktrace(sval.match, refIndex + 1, querySize, btrace);
        }
    }
    if (sc.isRequireEdgeDeletion)
    {
        const ScoreVal& sval((*thisSV)[querySize]);
        updateBacktrace(sval.match, refSize, querySize, btrace, ST_MATCH);
        updateBacktrace(sval.del, refSize, querySize, btrace, ST_DELETE);
    }
    if (sc.isAllowEdgeInsertion)
    {
        const ScoreVal& sval((*thisSV)[querySize]);
        updateBacktrace(sval.ins, refSize, querySize, btrace, ST_INSERT);
    }
    for (unsigned queryIndex(0); queryIndex < querySize; queryIndex++)
    {
        const ScoreVal& sval((*thisSV)[queryIndex]);
        const int thisMax(sval.match + (int)(querySize - queryIndex) * sc.offEdge);
        updateBacktrace(thisMax, refSize, queryIndex, btrace);
    }

    // backTraceAlignment
    res.score = btrace.max;
    std::vector<Seg> apath;
    Seg ps{-1, 0};
    auto updatePath = [&](int atype) {
        if (ps.type == atype) return;
        if (ps.type != -1) apath.push_back(ps);
        ps.type = atype;
        ps.length = 0;
    };
    if (btrace.queryBegin < querySize)
    {
        ps.type = CIG_S;
        ps.length = (querySize - btrace.queryBegin);
    }
    while (true)
    {
        const int nextState(PM(btrace.queryBegin, btrace.refBegin).get(btrace.state));
        if (btrace.state == ST_MATCH)
        {
            if ((btrace.queryBegin < 1) || (btrace.refBegin < 1)) break;
            updatePath(CIG_M);
            btrace.queryBegin--;
            btrace.refBegin--;
        }
        else if (btrace.state == ST_DELETE)
        {
            if (btrace.refBegin < 1) break;
            updatePath(CIG_D);
            btrace.refBegin--;
        }
        else
        {
            if (btrace.queryBegin < 1) break;
            updatePath(CIG_I);
            btrace.queryBegin--;
        }
        btrace.state = nextState;
        ps.length++;
    }
    if (ps.type != -1) apath.push_back(ps);
    if (btrace.queryBegin != 0)
    {
        ps.type = CIG_S;
        ps.length = btrace.queryBegin;
        apath.push_back(ps);
    }
    res.beginPos = btrace.refBegin;
    std::reverse(apath.begin(), apath.end());

    // apath_add_seqmatch: M -> runs of '=' / 'X'; an 'N' on either side is a mismatch
    std::vector<Seg> apath2;
    {
        unsigned qi(0), rix(res.beginPos);
        for (const Seg& s : apath)
        {
            if (s.type == CIG_M)
            {
                for (unsigned k(0); k < s.length; ++k)
                {
                    bool isSeqMatch(query[qi] == ref[rix]);
                    if ((query[qi] == 'N') || (ref[rix] == 'N')) isSeqMatch = false;
                    const int t(isSeqMatch ? CIG_EQ : CIG_X);
                    if (!apath2.empty() && apath2.back().type == t) apath2.back().length++;
                    else apath2.push_back(Seg{t, 1});
                    ++qi;
                    ++rix;
                }
            }
            else
            {
                apath2.push_back(s);
                if (s.type == CIG_I || s.type == CIG_S) qi += s.length;
                if (s.type == CIG_D) rix += s.length;
            }
        }
    }
    res.n_ops = (uint32_t)apath2.size();
    res.status = (res.n_ops > max_ops) ? 1 : 0;
    for (uint32_t i(0); i < res.n_ops && i < max_ops; ++i) cigar[i] = (apath2[i].length << 4) | (uint32_t)apath2[i].type;
}
} // namespace

extern "C" int ox_global_align(const sx_ga_scores* s, const sx_ga_batch* b, sx_ga_result* res, uint32_t* cigar)
{
    for (uint32_t i = 0; i < b->n; ++i)
    {
        const unsigned qs(b->query_off[i + 1] - b->query_off[i]);
        const unsigned rs(b->ref_off[i + 1] - b->ref_off[i]);
        if (qs == 0 || rs == 0) return SX_ERR_ARG; // asserts at GlobalAlignerImpl.hh:47-48
        align_one(*s, b->query + b->query_off[i], qs, b->ref + b->ref_off[i], rs, res[i], cigar + (size_t)i * b->max_ops, b->max_ops);
    }
    return SX_OK;
}

// ------------------------------------------------------------------------------------------------
// pileup helpers: base_call bit layout  blt_common/snp_pos_info.hh:109-118
// ------------------------------------------------------------------------------------------------
namespace
{
struct base_call
{
    explicit base_call(uint16_t v) : raw(v) {}
    unsigned get_qscore() const { return raw & 63; }
    unsigned base_id() const { return (raw >> 6) & 15; }
    bool is_fwd_strand() const { return (raw >> 10) & 1; }
    bool is_neighbor_mismatch() const { return (raw >> 11) & 1; }
    bool is_call_filter() const { return (raw >> 12) & 1; }
    bool is_tier_specific_call_filter() const { return (raw >> 13) & 1; }
    double error_prob() const { return qc().q2p[get_qscore()]; }
    double ln_error_prob() const { return qc().q2lne[get_qscore()]; }
    double ln_comp_error_prob() const { return qc().q2lncompe[get_qscore()]; }
    uint16_t raw;
};

inline unsigned base_to_id(const char c) // blt_util/seq_util.hh
{
    switch (c)
    {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return 4;
    }
}

// CleanPileupFilter  starling_common/PileupCleaner.cpp:30-64
void clean_pileup(const sx_pileup_batch* b, const uint32_t site, const bool is_include_tier2, std::vector<base_call>& calls)
{
    calls.clear();
    for (uint32_t i = b->site_off[site]; i < b->site_off[site + 1]; ++i)
    {
        const base_call bc(b->calls[i]);
        if (bc.is_call_filter())
        {
            if (!(is_include_tier2 && bc.is_tier_specific_call_filter())) continue;
        }
        calls.push_back(bc);
    }
    if (is_include_tier2 && b->t2_off)
    {
        for (uint32_t i = b->t2_off[site]; i < b->t2_off[site + 1]; ++i)
        {
            const base_call bc(b->t2_calls[i]);
            if (bc.is_call_filter()) continue;
            calls.push_back(bc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// a8  adjust_joint_eprob   blt_common/adjust_joint_eprob.cpp:60-243
// ------------------------------------------------------------------------------------------------
blt_float_t get_dependent_eprob(const unsigned qscore, const blt_float_t vexp)
{
    static const blt_float_t dep_converge_prob(0.75);
    const blt_float_t eprob(qc().q2p[qscore]);
    const blt_float_t val(std::pow(eprob, vexp));
    const blt_float_t frac((1 - val) / (1 - eprob));
    return std::max(eprob, frac * val + (1 - frac) * dep_converge_prob);
}

struct dependent_prob_cache // blt_common/adjust_joint_eprob.hh
{
    dependent_prob_cache() : is_init(MAX_QSCORE + 1, false), val(MAX_QSCORE + 1) {}
    blt_float_t get_dependent_val(const unsigned qscore, const blt_float_t vexp)
    {
        if (!is_init[qscore])
        {
            val[qscore] = get_dependent_eprob(qscore, vexp);
            is_init[qscore] = true;
        }
        return val[qscore];
    }
    std::vector<bool> is_init;
    std::vector<blt_float_t> val;
};

struct sort_icall_by_eprob
{
    explicit sort_icall_by_eprob(const std::vector<base_call>& c) : calls(c) {}
    bool operator()(const unsigned& a, const unsigned& b) const { return (calls[a].get_qscore() > calls[b].get_qscore()); }
    const std::vector<base_call>& calls;
};

void adjust_icalls_eprob(const sx_params& opt, dependent_prob_cache& dpc, std::vector<unsigned>& ic, const std::vector<base_call>& calls,
                         std::vector<float>& dependent_eprob)
{
    const unsigned ic_size(ic.size());
    blt_float_t vexp_frac;
    {
        static const blt_float_t lnran(std::log(0.75));
        blt_float_t num(0);
        blt_float_t den(0);
        for (unsigned i(0); i < ic_size; ++i)
        {
            const base_call& bi(calls[ic[i]]);
            const blt_float_t weight(lnran - bi.ln_error_prob());
            den += weight;
            if (bi.is_neighbor_mismatch()) num += weight;
        }
        blt_float_t mismatch_frac(0);
        if (ic_size && (den > 0.)) mismatch_frac = (num / den);
        vexp_frac = (1 - mismatch_frac) * opt.bsnp_ssd_no_mismatch + mismatch_frac * opt.bsnp_ssd_one_mismatch;
    }
    const bool is_limit_vexp(opt.is_min_vexp);
    const blt_float_t min_vexp(opt.min_vexp);
    bool is_min_vexp(false);

    std::sort(ic.begin(), ic.end(), sort_icall_by_eprob(calls));
    blt_float_t vexp(1.);
    for (unsigned i(0); i < ic_size; ++i)
    {
        const base_call& bi(calls[ic[i]]);
        if (!is_min_vexp)
        {
            dependent_eprob[ic[i]] = static_cast<float>(get_dependent_eprob(bi.get_qscore(), vexp));
            blt_float_t next_vexp(vexp);
            next_vexp *= (1 - vexp_frac);
            if (is_limit_vexp)
            {
                is_min_vexp = (next_vexp <= min_vexp);
                vexp = std::max(min_vexp, next_vexp);
            }
            else
            {
                vexp = next_vexp;
            }
        }
        else
        {
            dependent_eprob[ic[i]] = static_cast<float>(dpc.get_dependent_val(bi.get_qscore(), vexp));
        }
    }
}

void adjust_joint_eprob(const sx_params& opt, dependent_prob_cache& dpc, const std::vector<base_call>& calls, std::vector<float>& dependent_eprob)
{
    const unsigned n_calls(calls.size());
    dependent_eprob.clear();
    for (unsigned i(0); i < n_calls; ++i) dependent_eprob.push_back(static_cast<float>(calls[i].error_prob()));

    // blt_options::is_dependent_eprob  blt_common/blt_shared.hh:76-81
    if (!(opt.is_bsnp_diploid && (opt.bsnp_ssd_no_mismatch > 0. || opt.bsnp_ssd_one_mismatch > 0))) return;

    static const unsigned group_size(8);
    std::vector<unsigned> icalls[group_size];
    for (unsigned i(0); i < n_calls; ++i)
    {
        const base_call& b(calls[i]);
        if (b.is_call_filter()) continue;
        if (b.get_qscore() < 3) continue;
        const unsigned group_index((b.is_fwd_strand()) + (2 * b.base_id()));
        if (group_index >= group_size) continue; // base_id ANY never reaches the pileup (snp_util.hh:43 assert)
        icalls[group_index].push_back(i);
    }
    for (unsigned i(0); i < group_size; ++i) adjust_icalls_eprob(opt, dpc, icalls[i], calls, dependent_eprob);
}

// ------------------------------------------------------------------------------------------------
// a6-a7  pprob_digt_caller   blt_common/position_snp_call_pprob_digt.cpp
// ------------------------------------------------------------------------------------------------
const unsigned N_BASE = 4;
const unsigned DIGT_SIZE = 10;

inline bool digt_is_het(const unsigned idx) { return idx >= N_BASE; }

inline double digt_expect(const int base_id, const int gt) // blt_util/digt.hh:94-113
{
    static const double ex[DIGT_SIZE][N_BASE] = {{1.0, 0.0, 0.0, 0.0}, {0.0, 1.0, 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0}, {0.0, 0.0, 0.0, 1.0},
                                                 {0.5, 0.5, 0.0, 0.0}, {0.5, 0.0, 0.5, 0.0}, {0.5, 0.0, 0.0, 0.5}, {0.0, 0.5, 0.5, 0.0},
                                                 {0.0, 0.5, 0.0, 0.5}, {0.0, 0.0, 0.5, 0.5}};
    return ex[gt][base_id];
}

inline unsigned digt_expect2(const int base_id, const int gt) // blt_util/digt.hh:119-140
{
    static const unsigned ex[DIGT_SIZE][N_BASE] = {{2, 0, 0, 0}, {0, 2, 0, 0}, {0, 0, 2, 0}, {0, 0, 0, 2}, {1, 1, 0, 0},
                                                   {1, 0, 1, 0}, {1, 0, 0, 1}, {0, 1, 1, 0}, {0, 1, 0, 1}, {0, 0, 1, 1}};
    return ex[gt][base_id];
}

const blt_float_t one_third(1. / 3.);
const blt_float_t log_one_third(std::log(one_third));
const blt_float_t one_half(1. / 2.);
const blt_float_t log_one_half(std::log(one_half));

struct prior_set
{
    prior_set()
    {
        for (unsigned i(0); i < DIGT_SIZE; ++i) genome[i] = poly[i] = 0;
    }
    blt_float_t genome[DIGT_SIZE];
    blt_float_t poly[DIGT_SIZE];
};
typedef std::array<prior_set, N_BASE + 1> prior_group;

void get_genomic_prior(const unsigned ref_gt, const blt_float_t theta, blt_float_t* const prior) // :50-72
{
    blt_float_t prior_sum(0.);
    for (unsigned gt(0); gt < DIGT_SIZE; ++gt)
    {
        if (gt == ref_gt) continue;
        prior[gt] = (theta * one_third);
        if (digt_is_het(gt))
        {
            if (digt_expect(ref_gt, gt) <= 0.) prior[gt] *= theta;
        }
        else
        {
            prior[gt] *= .5;
        }
        prior_sum += prior[gt];
    }
    prior[ref_gt] = (1. - prior_sum);
}

void get_haploid_genomic_prior(const unsigned ref_gt, const blt_float_t theta, blt_float_t* const prior) // :76-97
{
    blt_float_t prior_sum(0.);
    for (unsigned gt(0); gt < DIGT_SIZE; ++gt)
    {
        if (gt == ref_gt) continue;
        if (digt_is_het(gt)) prior[gt] = 0;
        else prior[gt] = (theta * one_third);
        prior_sum += prior[gt];
    }
    prior[ref_gt] = (1. - prior_sum);
}

void get_poly_prior(const unsigned ref_gt, const blt_float_t theta, blt_float_t* const prior) // :101-135
{
    const blt_float_t ctheta(1. - theta);
    for (unsigned gt(0); gt < DIGT_SIZE; ++gt)
    {
        if (gt == ref_gt)
        {
            prior[gt] = 0.25 * (ctheta);
        }
        else if (digt_is_het(gt))
        {
            if (digt_expect(ref_gt, gt) <= 0.) prior[gt] = theta * one_third;
            else prior[gt] = 0.5 * one_third * ctheta;
        }
        else
        {
            prior[gt] = 0.25 * one_third * ctheta;
        }
    }
}

void get_haploid_poly_prior(const unsigned ref_gt, const blt_float_t, blt_float_t* const prior) // :139-163
{
    for (unsigned gt(0); gt < DIGT_SIZE; ++gt)
    {
        if (gt == ref_gt) prior[gt] = 0.5;
        else if (digt_is_het(gt)) prior[gt] = 0;
        else prior[gt] = 0.5 * one_third;
    }
}

void finish_prior(prior_group& prior) // :205-228
{
    prior_set& nps(prior[N_BASE]);
    for (unsigned i(0); i < N_BASE; ++i)
    {
        prior_set& ps(prior[i]);
        for (unsigned gt(0); gt < DIGT_SIZE; ++gt)
        {
            nps.genome[gt] += ps.genome[gt];
            nps.poly[gt] += ps.poly[gt];
        }
    }
    auto norm_gt = [](blt_float_t* const x) {
        blt_float_t sum(0);
        for (unsigned gt(0); gt < DIGT_SIZE; ++gt) sum += x[gt];
        sum = 1. / sum;
        for (unsigned gt(0); gt < DIGT_SIZE; ++gt) x[gt] *= sum;
    };
    norm_gt(nps.genome);
    norm_gt(nps.poly);
    for (unsigned i(0); i < (N_BASE + 1); ++i)
    {
        prior_set& ps(prior[i]);
        for (unsigned gt(0); gt < DIGT_SIZE; ++gt)
        {
            ps.genome[gt] = std::log(ps.genome[gt]);
            ps.poly[gt] = std::log(ps.poly[gt]);
        }
    }
}

struct pprob_digt_caller
{
    explicit pprob_digt_caller(const blt_float_t theta) // :232-248
    {
        for (unsigned i(0); i < N_BASE; ++i)
        {
            get_genomic_prior(i, theta, lnprior[i].genome);
            get_poly_prior(i, theta, lnprior[i].poly);
            get_haploid_genomic_prior(i, theta, lnprior_haploid[i].genome);
            get_haploid_poly_prior(i, theta, lnprior_haploid[i].poly);
        }
        finish_prior(lnprior);
        finish_prior(lnprior_haploid);
    }
    prior_group lnprior;
    prior_group lnprior_haploid;
};

void get_diploid_gt_lhood(const std::vector<base_call>& calls, const std::vector<float>& de, const unsigned ref_gt, blt_float_t* const lhood,
                          const bool is_strand_specific = false, const bool is_ss_fwd = false) // :326-385
{
    for (unsigned gt(0); gt < DIGT_SIZE; ++gt) lhood[gt] = 0.;
    const unsigned n_calls(calls.size());
    for (unsigned i(0); i < n_calls; ++i)
    {
        const base_call& bc(calls[i]);
        const blt_float_t eprob(de[i]);
        const blt_float_t ceprob(1. - bc.error_prob());
        const blt_float_t lnce(bc.ln_comp_error_prob());
        blt_float_t val[3];
        val[0] = std::log(eprob) + log_one_third;
        val[1] = std::log((ceprob) + ((1. - ceprob) * one_third)) + log_one_half;
        val[2] = lnce;
        const bool is_force_ref(is_strand_specific && (is_ss_fwd != bc.is_fwd_strand()));
        const uint8_t obs_id(bc.base_id());
        for (unsigned gt(0); gt < DIGT_SIZE; ++gt) lhood[gt] += val[digt_expect2(obs_id, (is_force_ref ? ref_gt : gt))];
    }
}

// blt_util/qscore.hh:40-72
template <typename FloatType> FloatType error_prob_to_phred(const FloatType prob)
{
    static const FloatType minlog10(static_cast<FloatType>(std::numeric_limits<FloatType>::min_exponent10));
    return -10. * std::max(minlog10, std::log10(prob));
}
template <typename FloatType> FloatType ln_error_prob_to_phred(const FloatType lnProb)
{
    static const FloatType minlog10(static_cast<FloatType>(std::numeric_limits<FloatType>::min_exponent10));
    static const FloatType ln10(std::log(static_cast<FloatType>(10)));
    return -10. * std::max(minlog10, lnProb / ln10);
}
template <typename FloatType> int error_prob_to_qphred(const FloatType prob) { return static_cast<int>(std::floor(error_prob_to_phred(prob) + 0.5)); }
template <typename FloatType> int ln_error_prob_to_qphred(const FloatType lnProb) { return static_cast<int>(std::floor(ln_error_prob_to_phred(lnProb) + 0.5)); }

// blt_util/prob_util.hh:179-237
template <typename It> double prob_comp(It begin, const It end, const unsigned cgt)
{
    unsigned i(0);
    double val(0.);
    for (; begin != end; ++begin, ++i)
    {
        if (i == cgt) continue;
        val = val + *begin;
    }
    return val;
}
template <typename It> void normalizeLogDistro(const It pbegin, const It pend, unsigned& maxElementIndex)
{
    maxElementIndex = 0;
    if (pbegin == pend) return;
    double max(*pbegin);
    unsigned i(1);
    for (It p(pbegin + 1); p != pend; ++p, ++i)
    {
        if (*p > max)
        {
            max = *p;
            maxElementIndex = i;
        }
    }
    double sum(0.);
    for (It p(pbegin); p != pend; ++p)
    {
        *p = std::exp(*p - max);
        sum += *p;
    }
    sum = 1. / sum;
    for (It p(pbegin); p != pend; ++p) *p *= sum;
}

void calculate_result_set(const blt_float_t* lhood, const blt_float_t* lnprior, const unsigned ref_gt, sx_digt_result_set& rs) // :412-433
{
    std::array<double, DIGT_SIZE> pprob;
    for (unsigned gt(0); gt < DIGT_SIZE; ++gt) pprob[gt] = lhood[gt] + lnprior[gt];
    unsigned max_gt(0);
    normalizeLogDistro(pprob.begin(), pprob.end(), max_gt);
    rs.max_gt = max_gt;
    rs.ref_pprob = pprob[ref_gt];
    rs.snp_qphred = error_prob_to_qphred(pprob[ref_gt]);
    rs.max_gt_qphred = error_prob_to_qphred(prob_comp(pprob.begin(), pprob.end(), rs.max_gt));
}

void reset_digt(sx_digt_result& d)
{
    std::memset(&d, 0, sizeof(d));
}

} // namespace

extern "C" int ox_dependent_eprob(const sx_params* p, const sx_pileup_batch* b, uint32_t* out_off, float* de)
{
    dependent_prob_cache dpc;
    std::vector<base_call> calls;
    std::vector<float> dep;
    uint32_t off(0);
    for (uint32_t s = 0; s < b->n_sites; ++s)
    {
        clean_pileup(b, s, false, calls);
        adjust_joint_eprob(*p, dpc, calls, dep);
        out_off[s] = off;
        for (float v : dep) de[off++] = v;
    }
    out_off[b->n_sites] = off;
    return SX_OK;
}

extern "C" int ox_site_gl_germline_range(const sx_params* p, const sx_pileup_batch* b, int is_always_test, uint32_t s0, uint32_t s1, sx_digt_result* out)
{
    if (p->hetVariantFrequencyExtension > 0) return SX_ERR_UNSUPPORTED;
    const pprob_digt_caller caller(p->bsnp_diploid_theta);
    dependent_prob_cache dpc; // PileupCleaner::_dpcache: one per processor, lives across sites
    std::vector<base_call> calls;
    std::vector<float> de;
    for (uint32_t s = s0; s < s1; ++s)
    {
        sx_digt_result& dgt(out[s]);
        reset_digt(dgt);
        clean_pileup(b, s, false, calls);        // CleanPileupFilter
        dgt.n_used_calls = calls.size();
        adjust_joint_eprob(*p, dpc, calls, de);  // CleanPileupErrorProb

        // position_snp_call_pprob_digt :471-539
        const char ref_base(b->ref_base[s]);
        const int ploidy(b->ploidy ? b->ploidy[s] : 2);
        if (ref_base == 'N') continue;
        dgt.ref_gt = base_to_id(ref_base);
        if (!is_always_test)
        {
            bool allref(true); // is_spi_allref  blt_common/snp_util.hh:34-47
            for (const base_call& bc : calls)
                if (dgt.ref_gt != bc.base_id())
                {
                    allref = false;
                    break;
                }
            if (allref) continue;
        }
        dgt.is_computed = 1;
        const bool is_haploid(ploidy == 1);
        blt_float_t lhood[DIGT_SIZE];
        get_diploid_gt_lhood(calls, de, dgt.ref_gt, lhood);
        {
            unsigned gtcount(DIGT_SIZE);
            if (is_haploid) gtcount = N_BASE;
            unsigned maxIndex(0);
            for (unsigned gt(1); gt < gtcount; ++gt)
                if (lhood[gt] > lhood[maxIndex]) maxIndex = gt;
            for (unsigned gt(0); gt < gtcount; ++gt) dgt.phredLoghood[gt] = ln_error_prob_to_qphred(lhood[gt] - lhood[maxIndex]);
        }
        for (unsigned gt(0); gt < DIGT_SIZE; ++gt) dgt.lhood[gt] = lhood[gt];
        const prior_group& pg(is_haploid ? caller.lnprior_haploid : caller.lnprior);
        calculate_result_set(lhood, pg[dgt.ref_gt].genome, dgt.ref_gt, dgt.genome);
        calculate_result_set(lhood, pg[dgt.ref_gt].poly, dgt.ref_gt, dgt.poly);
        if (dgt.genome.snp_qphred != 0) // is_snp()
        {
            blt_float_t lhood_fwd[DIGT_SIZE];
            get_diploid_gt_lhood(calls, de, dgt.ref_gt, lhood_fwd, true, true);
            blt_float_t lhood_rev[DIGT_SIZE];
            get_diploid_gt_lhood(calls, de, dgt.ref_gt, lhood_rev, true, false);
            const unsigned tgt(dgt.genome.max_gt);
            dgt.strand_bias = std::max(lhood_fwd[tgt], lhood_rev[tgt]) - lhood[tgt];
        }
        else
        {
            dgt.strand_bias = 0;
        }
    }
    return SX_OK;
}

extern "C" int ox_site_gl_germline(const sx_params* p, const sx_pileup_batch* b, int is_always_test, sx_digt_result* out)
{
    return ox_site_gl_germline_range(p, b, is_always_test, 0, b->n_sites, out);
}

// ------------------------------------------------------------------------------------------------
// a10-a11 somatic SNV strand grid
//   applications/strelka/position_somatic_snv_strand_grid_lhood_cached.cpp:41-234
//   applications/strelka/position_somatic_snv_strand_grid.cpp:42-363
//   applications/strelka/qscore_calculator.cpp:33-209
//   applications/strelka/strelka_digt_states.{hh,cpp}
// ------------------------------------------------------------------------------------------------
namespace
{
enum { SD_REF = 0, SD_HOM = 1, SD_HET = 2, SD_SIZE = 3 };
enum { HET_RES = 9, HET_COUNT = HET_RES * 2 + 1, HOM_SIZE = 2, PRESTRAND_SIZE = HOM_SIZE + HET_COUNT, STRAND_STATE_SIZE = HET_RES, GRID_SIZE = PRESTRAND_SIZE + STRAND_STATE_SIZE };
const blt_float_t RATIO_INCREMENT = 0.5f / static_cast<blt_float_t>(HET_RES + 1);

blt_float_t get_fraction_from_index(int index) // strelka_digt_states.cpp:34-41
{
    if (index == SD_REF) return 0.f;
    if (index == SD_HOM) return 1.f;
    if (index == SD_HET) return 0.5f;
    if (index < SD_SIZE + HET_RES) return RATIO_INCREMENT * (index - SD_SIZE + 1);
    return RATIO_INCREMENT * (index - SD_SIZE + 2);
}

const blt_float_t s_ln_one_third(std::log(one_third));
const blt_float_t s_ln_one_half(std::log(one_half));

// the reference memoizes these per (qscore, ratio index) in function-static het_ratio_cache objects; the cached value is a pure
// function of its key, so memoization does not change results and is dropped here.
void get_diploid_gt_lhood_cached_simple(const std::vector<base_call>& calls, const unsigned ref_gt, blt_float_t* const lhood) // :41-85
{
    for (unsigned gt(0); gt < SD_SIZE; ++gt) lhood[gt] = 0.;
    for (const base_call& bc : calls)
    {
        blt_float_t val[3];
        const blt_float_t eprob(bc.error_prob());
        const blt_float_t ceprob(1 - eprob);
        const blt_float_t lne(bc.ln_error_prob());
        const blt_float_t lnce(bc.ln_comp_error_prob());
        val[0] = lne + s_ln_one_third;
        val[1] = std::log((ceprob) + ((eprob)*one_third)) + s_ln_one_half;
        val[2] = lnce;
        if (bc.base_id() == ref_gt)
        {
            lhood[SD_REF] += val[2];
            lhood[SD_HET] += val[1];
            lhood[SD_HOM] += val[0];
        }
        else
        {
            lhood[SD_REF] += val[0];
            lhood[SD_HET] += val[1];
            lhood[SD_HOM] += val[2];
        }
    }
}

void get_high_low_het_ratio_lhood_cached(const std::vector<base_call>& calls, const unsigned ref_gt, const blt_float_t het_ratio, blt_float_t* lhood_high,
                                         blt_float_t* lhood_low) // :87-131
{
    const blt_float_t chet_ratio(1. - het_ratio);
    for (const base_call& bc : calls)
    {
        const blt_float_t eprob(bc.error_prob());
        const blt_float_t ceprob(1 - eprob);
        blt_float_t val[2];
        val[0] = std::log((ceprob)*het_ratio + ((eprob)*one_third) * chet_ratio);
        val[1] = std::log((ceprob)*chet_ratio + ((eprob)*one_third) * het_ratio);
        if (bc.base_id() == ref_gt)
        {
            *lhood_high += val[0];
            *lhood_low += val[1];
        }
        else
        {
            *lhood_high += val[1];
            *lhood_low += val[0];
        }
    }
}

void get_diploid_het_grid_lhood_cached(const std::vector<base_call>& calls, const unsigned ref_gt, const unsigned hetResolution, blt_float_t* const lhood) // :133-153
{
    const unsigned totalHetRatios(hetResolution * 2);
    for (unsigned gt(0); gt < totalHetRatios; ++gt) lhood[gt] = 0.;
    for (unsigned hetIndex(0); hetIndex < hetResolution; ++hetIndex)
    {
        const blt_float_t het_ratio((hetIndex + 1) * RATIO_INCREMENT);
        get_high_low_het_ratio_lhood_cached(calls, ref_gt, het_ratio, lhood + (totalHetRatios - (hetIndex + 1)), lhood + hetIndex);
    }
}

blt_float_t getLogSum(blt_float_t x1, blt_float_t x2) // blt_util/logSumUtil.hh:33-41 with FloatType = float
{
    if (x1 < x2) std::swap(x1, x2);
    // log1p_switch<float>: boost::math::log1p(float) promotes to double internally (policy promote_float) and calls ::log1p; else std::log(1+x) float
    const blt_float_t x(std::exp(x2 - x1));
    static const blt_float_t smallx_thresh(0.01);
    blt_float_t l;
    if (std::abs(x) < smallx_thresh) l = static_cast<blt_float_t>(::log1p(static_cast<double>(x)));
    else l = std::log(1 + x);
    return x1 + l;
}

void get_strand_ratio_lhood_spi(const std::vector<base_call>& calls, const unsigned ref_gt, const blt_float_t het_ratio, blt_float_t* lhood) // :164-234
{
    const blt_float_t chet_ratio(1. - het_ratio);
    blt_float_t lhood_fwd = 0;
    blt_float_t lhood_rev = 0;
    for (const base_call& bc : calls)
    {
        blt_float_t val[2];
        const blt_float_t eprob(bc.error_prob());
        const blt_float_t ceprob(1. - eprob);
        val[0] = (std::log((ceprob)*chet_ratio + ((eprob)*one_third) * het_ratio));
        val[1] = (std::log((ceprob)*het_ratio + ((eprob)*one_third) * chet_ratio));
        if (bc.base_id() == ref_gt)
        {
            const blt_float_t val_off_strand(bc.ln_comp_error_prob());
            const blt_float_t val_fwd(bc.is_fwd_strand() ? val[0] : val_off_strand);
            const blt_float_t val_rev(bc.is_fwd_strand() ? val_off_strand : val[0]);
            lhood_fwd += val_fwd;
            lhood_rev += val_rev;
        }
        else
        {
            const blt_float_t val_off_strand(bc.ln_error_prob() + s_ln_one_third);
            const blt_float_t val_fwd(bc.is_fwd_strand() ? val[1] : val_off_strand);
            const blt_float_t val_rev(bc.is_fwd_strand() ? val_off_strand : val[1]);
            lhood_fwd += val_fwd;
            lhood_rev += val_rev;
        }
    }
    *lhood = getLogSum(lhood_fwd, lhood_rev) + s_ln_one_half;
}

struct snv_result_set
{
    unsigned ntype = 0, max_gt = 0;
    int qphred = 0, from_ntype_qphred = 0;
    unsigned normal_alt_id = 0, tumor_alt_id = 0;
    float strandBias = 0;
};

void calculate_result_set_grid(const blt_float_t contam_tolerance, const blt_float_t logSharedErrorRate, const blt_float_t logSharedErrorRateComplement,
                               const blt_float_t* normal_lhood, const blt_float_t* tumor_lhood, const blt_float_t* germlineGenotypeLogPrior,
                               const blt_float_t lnmatch, const blt_float_t lnmismatch, snv_result_set& rs) // qscore_calculator.cpp:47-209
{
    static const blt_float_t neg_inf = -std::numeric_limits<float>::infinity();
    static const blt_float_t ln_one_half(std::log(1. / 2.));
    static const blt_float_t log_error_mod = -std::log(static_cast<double>(PRESTRAND_SIZE - 1));

    double log_post_prob[SD_SIZE][2];
    double max_log_prob = neg_inf;
    rs.max_gt = 0;
    for (unsigned ngt(0); ngt < SD_SIZE; ++ngt)
    {
        for (unsigned tgt(0); tgt < 2; ++tgt)
        {
            double max_log_sum = neg_inf;
            double log_sum[PRESTRAND_SIZE * PRESTRAND_SIZE];
            int index = 0;
            for (unsigned tumor_freq_index(0); tumor_freq_index < PRESTRAND_SIZE; ++tumor_freq_index)
            {
                blt_float_t tumor_freq = get_fraction_from_index(tumor_freq_index);
                bool consider_norm_contam = contam_tolerance * tumor_freq >= RATIO_INCREMENT;
                for (unsigned normal_freq_index(0); normal_freq_index < PRESTRAND_SIZE; ++normal_freq_index)
                {
                    double lprior_freq;
                    if (tgt == 0)
                    {
                        if (normal_freq_index != tumor_freq_index) continue;
                        lprior_freq = (normal_freq_index == ngt) ? logSharedErrorRateComplement : logSharedErrorRate + log_error_mod;
                    }
                    else
                    {
                        if (normal_freq_index == tumor_freq_index) continue;
                        if (ngt != SD_REF)
                        {
                            if (normal_freq_index != ngt) continue;
                            lprior_freq = log_error_mod;
                        }
                        else
                        {
                            if (!consider_norm_contam)
                            {
                                if (normal_freq_index == 0) lprior_freq = log_error_mod;
                                else continue;
                            }
                            else
                            {
                                if ((normal_freq_index == ngt) || (normal_freq_index == SD_SIZE)) lprior_freq = log_error_mod + ln_one_half;
                                else continue;
                            }
                        }
                    }
                    double lsum = lprior_freq + normal_lhood[normal_freq_index] + tumor_lhood[tumor_freq_index];
                    log_sum[index++] = lsum;
                    if (lsum > max_log_sum) max_log_sum = lsum;
                }
            }
            double sum = 0.0;
            for (int i(0); i < index; ++i) sum += std::exp(log_sum[i] - max_log_sum);
            double log_genotype_prior = germlineGenotypeLogPrior[ngt] + ((tgt == 0) ? lnmatch : lnmismatch);
            log_post_prob[ngt][tgt] = log_genotype_prior + max_log_sum + std::log(sum);
            if (log_post_prob[ngt][tgt] > max_log_prob)
            {
                max_log_prob = log_post_prob[ngt][tgt];
                rs.max_gt = ngt * 2 + tgt; // DDIGT::get_state
            }
        }
    }
    double sum_prob = 0.0;
    for (unsigned ngt(0); ngt < SD_SIZE; ++ngt)
        for (unsigned tgt(0); tgt < 2; ++tgt)
        {
            double prob = std::exp(log_post_prob[ngt][tgt] - max_log_prob);
            sum_prob += prob;
        }
    double log_sum_prob = std::log(sum_prob);
    double min_not_somfrom_sum(INFINITY);
    double nonsom_prob = 0.0;
    double post_prob[SD_SIZE][2];
    for (unsigned ngt(0); ngt < SD_SIZE; ++ngt)
    {
        double som_prob_given_ngt(0);
        for (unsigned tgt(0); tgt < 2; ++tgt)
        {
            post_prob[ngt][tgt] = std::exp(log_post_prob[ngt][tgt] - max_log_prob - log_sum_prob);
            if (tgt == 0) nonsom_prob += post_prob[ngt][tgt];
            else som_prob_given_ngt += post_prob[ngt][tgt];
        }
        double err_som_and_ngt = 1.0 - som_prob_given_ngt;
        if (err_som_and_ngt < min_not_somfrom_sum)
        {
            min_not_somfrom_sum = err_som_and_ngt;
            rs.from_ntype_qphred = error_prob_to_qphred(err_som_and_ngt);
            rs.ntype = ngt;
        }
    }
    rs.qphred = error_prob_to_qphred(nonsom_prob);
}

unsigned get_most_frequent_alt_id(const std::vector<base_call>& calls, const unsigned ref_gt) // snp_pos_info.hh:175-198
{
    unsigned alt_count[5] = {};
    for (const base_call& tbc : calls)
    {
        const uint8_t obs_id(tbc.base_id());
        if (obs_id == ref_gt || obs_id == 4) continue;
        ++alt_count[obs_id];
    }
    unsigned alt_id = ref_gt;
    unsigned max_count = 0;
    for (unsigned base_id(0); base_id < 5; ++base_id)
    {
        if (alt_count[base_id] > max_count)
        {
            if (base_id == ref_gt) continue;
            max_count = alt_count[base_id];
            alt_id = base_id;
        }
    }
    return alt_id;
}

} // namespace

extern "C" int ox_site_gl_somatic_range(const sx_params* p, const sx_pileup_batch* normal, const sx_pileup_batch* tumor, const uint8_t* is_forced_output,
                                        uint32_t s0, uint32_t s1, sx_ssnv_result* out)
{
    // somatic_snv_caller_strand_grid ctor  position_somatic_snv_strand_grid.cpp:42-55
    const blt_float_t contam_tolerance(p->ssnv_contam_tolerance);
    const blt_float_t ln_csse_rate(log1p_switch(-p->shared_site_error_rate));
    const blt_float_t ln_som_match(log1p_switch(-p->somatic_snv_rate));
    const blt_float_t ln_som_mismatch(std::log(p->somatic_snv_rate));
    blt_float_t germlineGenotypeLogPrior[SD_SIZE];
    {
        const double theta(p->bsnp_diploid_theta); // qscore_calculator.cpp:33-43
        germlineGenotypeLogPrior[SD_REF] = (blt_float_t)log1p_switch(-(3. * theta) / 2.);
        germlineGenotypeLogPrior[SD_HOM] = (blt_float_t)std::log(theta / 2.);
        germlineGenotypeLogPrior[SD_HET] = (blt_float_t)std::log(theta);
    }
    const blt_float_t strand_sse_rate(p->shared_site_error_rate * p->shared_site_error_strand_bias_fraction);
    const blt_float_t nostrand_sse_rate(p->shared_site_error_rate - strand_sse_rate);
    const blt_float_t ln_sse_rate(std::log(nostrand_sse_rate));

    const bool is_tier2(normal->t2_off != nullptr && tumor->t2_off != nullptr);
    std::vector<base_call> ncalls[2], tcalls[2];
    for (uint32_t s = s0; s < s1; ++s)
    {
        sx_ssnv_result& sgt(out[s]);
        std::memset(&sgt, 0, sizeof(sgt));
        const bool forced(is_forced_output ? is_forced_output[s] != 0 : false);
        const char ref_base(normal->ref_base[s]);
        if (ref_base == 'N') continue;
        sgt.ref_gt = base_to_id(ref_base);
        for (int t = 0; t < (is_tier2 ? 2 : 1); ++t)
        {
            clean_pileup(normal, s, t == 1, ncalls[t]);
            clean_pileup(tumor, s, t == 1, tcalls[t]);
        }
        if (!forced)
        {
            auto allref = [&](const std::vector<base_call>& c) {
                for (const base_call& bc : c)
                    if (bc.base_id() != sgt.ref_gt) return false;
                return true;
            };
            if (allref(ncalls[0]) && allref(tcalls[0])) continue;
        }
        blt_float_t normal_lhood[2][GRID_SIZE];
        blt_float_t tumor_lhood[2][GRID_SIZE];
        snv_result_set tier_rs[2];
        for (unsigned i(0); i < 2; ++i)
        {
            for (unsigned k(0); k < GRID_SIZE; ++k) normal_lhood[i][k] = tumor_lhood[i][k] = 0;
            const bool is_include_tier2(i == 1);
            if (is_include_tier2)
            {
                if (!is_tier2) continue;
                if (tier_rs[0].qphred == 0)
                {
                    tier_rs[1] = tier_rs[0];
                    for (unsigned k(0); k < GRID_SIZE; ++k)
                    {
                        normal_lhood[1][k] = normal_lhood[0][k];
                        tumor_lhood[1][k] = tumor_lhood[0][k];
                    }
                    continue;
                }
            }
            get_diploid_gt_lhood_cached_simple(ncalls[i], sgt.ref_gt, normal_lhood[i]);
            get_diploid_gt_lhood_cached_simple(tcalls[i], sgt.ref_gt, tumor_lhood[i]);
            get_diploid_het_grid_lhood_cached(ncalls[i], sgt.ref_gt, HET_RES, normal_lhood[i] + SD_SIZE);
            get_diploid_het_grid_lhood_cached(tcalls[i], sgt.ref_gt, HET_RES, tumor_lhood[i] + SD_SIZE);
            for (unsigned k(0); k < HET_RES; ++k) // get_diploid_strand_grid_lhood_spi  strand_grid.cpp:61-81
            {
                const blt_float_t het_ratio((k + 1) * RATIO_INCREMENT);
                get_strand_ratio_lhood_spi(tcalls[i], sgt.ref_gt, het_ratio, tumor_lhood[i] + PRESTRAND_SIZE + k);
            }
            calculate_result_set_grid(contam_tolerance, ln_sse_rate, ln_csse_rate, normal_lhood[i], tumor_lhood[i], germlineGenotypeLogPrior, ln_som_match,
                                      ln_som_mismatch, tier_rs[i]);
            // wrapper strand_grid.cpp:157-226: early return leaves strandBias at its default when qphred==0 and not forced
            if (forced || tier_rs[i].qphred != 0)
            {
                const blt_float_t symm_lhood(*std::max_element(tumor_lhood[i] + SD_SIZE, tumor_lhood[i] + PRESTRAND_SIZE));
                const blt_float_t strand_lhood(*std::max_element(tumor_lhood[i] + PRESTRAND_SIZE, tumor_lhood[i] + GRID_SIZE));
                tier_rs[i].strandBias = std::max(0.f, (strand_lhood - symm_lhood));
            }
            tier_rs[i].normal_alt_id = get_most_frequent_alt_id(ncalls[i], sgt.ref_gt);
            tier_rs[i].tumor_alt_id = get_most_frequent_alt_id(tcalls[i], sgt.ref_gt);
        }
        if (!forced)
        {
            if ((tier_rs[0].qphred == 0) || (is_tier2 && (tier_rs[1].qphred == 0))) continue;
        }
        sgt.is_computed = 1;
        sgt.snv_tier = 0;
        sgt.snv_from_ntype_tier = 0;
        if (is_tier2)
        {
            if (tier_rs[0].qphred > tier_rs[1].qphred) sgt.snv_tier = 1;
            if (tier_rs[0].from_ntype_qphred > tier_rs[1].from_ntype_qphred) sgt.snv_from_ntype_tier = 1;
        }
        snv_result_set rs(tier_rs[sgt.snv_from_ntype_tier]);
        if (is_tier2 && (tier_rs[0].ntype != tier_rs[1].ntype))
        {
            rs.ntype = 3; // NTYPE::CONFLICT
            rs.from_ntype_qphred = 0;
        }
        // else: SOMATIC_DIGT {REF,HOM,HET} maps onto NTYPE {REF,HOM,HET} value-for-value (somatic_call_shared.hh)
        rs.qphred = tier_rs[sgt.snv_tier].qphred;
        sgt.ntype = rs.ntype;
        sgt.max_gt = rs.max_gt;
        sgt.qphred = rs.qphred;
        sgt.from_ntype_qphred = rs.from_ntype_qphred;
        sgt.normal_alt_id = rs.normal_alt_id;
        sgt.tumor_alt_id = rs.tumor_alt_id;
        sgt.strandBias = rs.strandBias;
        for (unsigned k(0); k < GRID_SIZE; ++k)
        {
            sgt.normal_lhood[k] = normal_lhood[sgt.snv_from_ntype_tier][k];
            sgt.tumor_lhood[k] = tumor_lhood[sgt.snv_from_ntype_tier][k];
        }
    }
    return SX_OK;
}

extern "C" int ox_site_gl_somatic(const sx_params* p, const sx_pileup_batch* normal, const sx_pileup_batch* tumor, const uint8_t* is_forced_output,
                                  sx_ssnv_result* out)
{
    return ox_site_gl_somatic_range(p, normal, tumor, is_forced_output, 0, normal->n_sites, out);
}

// ------------------------------------------------------------------------------------------------
// test exports: the restatements the DEVICE code mirrors, next to the real thing
// ------------------------------------------------------------------------------------------------
extern "C" float ox_logf_restated(float x) { return sx_logf(x); }
extern "C" float ox_powf_restated(float x, float y) { return sx_powf(x, y); }

// the double-precision mirrors (sx_libm_mirror_d.h) against the live libm, argument by argument: returns the number of arguments whose results
// differ in any bit (kind 0 exp, 1 log10, 2 log, 3 log1p); *first_bad = the first such argument
#include "../strelka_b200/csrc/sx_libm_mirror_d.h"
// sx_expf against the live expf on every stride-th float bit pattern of [start, end]: the number of mismatching arguments
extern "C" uint64_t ox_expf_mirror_check(uint32_t start, uint32_t end, uint32_t stride, uint32_t* first_bad)
{
    uint64_t bad(0);
    for (uint64_t u = start; u <= (uint64_t)end; u += stride)
    {
        const uint32_t v((uint32_t)u);
        float x;
        std::memcpy(&x, &v, 4);
        const float a(std::exp(x)), b(sx_expf(x));
        uint32_t ua, ub;
        std::memcpy(&ua, &a, 4);
        std::memcpy(&ub, &b, 4);
        if (ua != ub && !(a != a && b != b))
        {
            if (!bad && first_bad) *first_bad = v;
            ++bad;
        }
    }
    return bad;
}

extern "C" uint64_t ox_libm_d_mirror_check(int kind, const double* xs, uint64_t n, double* first_bad)
{
    uint64_t bad(0);
    for (uint64_t i = 0; i < n; ++i)
    {
        const double x(xs[i]);
        double a, b;
        if (kind == 0)
        {
            a = std::exp(x);
            b = sx_exp(x);
        }
        else if (kind == 1)
        {
            a = std::log10(x);
            b = sx_log10(x);
        }
        else if (kind == 2)
        {
            a = std::log(x);
            b = sx_log(x);
        }
        else
        {
            a = std::log1p(x);
            b = sx_log1p(x);
        }
        uint64_t ua, ub;
        std::memcpy(&ua, &a, 8);
        std::memcpy(&ub, &b, 8);
        if (ua != ub && !(a != a && b != b))
        {
            if (!bad && first_bad) *first_bad = x;
            ++bad;
        }
    }
    return bad;
}

namespace
{
#include "../strelka_b200/csrc/sx_stdsort_mirror.h"
}

extern "C" void ox_sort_restated(uint32_t* idx, uint32_t n, const uint8_t* key_by_idx)
{
    sx_stdsort_desc(idx, n, key_by_idx);
}

extern "C" void ox_sort_std(uint32_t* idx, uint32_t n, const uint8_t* key_by_idx)
{
    std::sort(idx, idx + n, [key_by_idx](const uint32_t& a, const uint32_t& b) { return key_by_idx[a] > key_by_idx[b]; });
}

// ------------------------------------------------------------------------------------------------
// a12  indel genotype likelihoods
//   getVariantAlleleGroupGenotypeLhoodsForSample       starling_common/AlleleGroupGenotype.cpp:184-258
//   updateGenotypeLogLhoodFromAlleleLogLhood           :34-111
//   updateSupportingReadStats                          :122-152
//   integrateOutMappingStatus                          starling_common/readMappingAdjustmentUtil.hh:28-56
//   get_het_observed_allele_ratio                      starling_common/starling_indel_call_pprob_digt.cpp:40-71
//   getLogSum<double>                                  blt_util/logSumUtil.hh:33-41
// ------------------------------------------------------------------------------------------------
namespace
{
double getLogSumD(double x1, double x2)
{
    if (x1 < x2) std::swap(x1, x2);
    return x1 + log1p_switch(std::exp(x2 - x1));
}

void get_het_observed_allele_ratio(const unsigned read_length, const unsigned min_overlap, const unsigned del_len, const unsigned ins_len,
                                   const double het_allele_ratio, double& log_ref_prob, double& log_indel_prob)
{
    const unsigned base_expect((read_length + 1) < (2 * min_overlap) ? 0 : (read_length + 1) - (2 * min_overlap));
    const double ref_path_expect(base_expect + std::min(del_len, base_expect));
    const double indel_path_expect(base_expect + std::min(ins_len, base_expect));
    const double ref_path_term((1 - het_allele_ratio) * ref_path_expect);
    const double indel_path_term(het_allele_ratio * indel_path_expect);
    const double total_path_term(ref_path_term + indel_path_term);
    if (total_path_term > 0)
    {
        const double indel_prob(indel_path_term / total_path_term);
        log_ref_prob = std::log(1. - indel_prob);
        log_indel_prob = std::log(indel_prob);
    }
}
} // namespace

extern "C" int ox_indel_gl(const sx_params* p, const sx_indel_batch* b, sx_indel_result* out)
{
    const double randomBaseMatchLogProb(std::log(p->randomBaseMatchProb));  // starling_base_shared.cpp:44
    const double correctMappingLogPrior(std::log(1.7e-10));                 // :66
    const unsigned min_read_bp_flank(p->min_read_bp_flank);
    const double readSupportThreshold(p->readConfidentSupportThreshold);
    auto integrateOutMappingStatus = [&](const uint16_t nonAmbiguousBasesInRead, const double correctMappingLogLikelihood) {
        return getLogSumD((correctMappingLogLikelihood + correctMappingLogPrior), randomBaseMatchLogProb * nonAmbiguousBasesInRead);
    };
    for (uint32_t l = 0; l < b->n_loci; ++l)
    {
        sx_indel_result& o(out[l]);
        std::memset(&o, 0, sizeof(o));
        const unsigned nonRefAlleleCount(b->allele_off[l + 1] - b->allele_off[l]);
        const unsigned fullAlleleCount(nonRefAlleleCount + 1);
        if (nonRefAlleleCount < 1 || nonRefAlleleCount > SX_INDEL_MAX_ALLELES) return SX_ERR_ARG;
        const unsigned callerPloidy(b->ploidy[l]);
        if (callerPloidy != 1 && callerPloidy != 2) return SX_ERR_ARG;
        const unsigned genotypeCount(callerPloidy == 1 ? fullAlleleCount : (fullAlleleCount * (fullAlleleCount + 1)) / 2);
        o.n_gt = genotypeCount;
        const uint16_t* del(b->allele_del_len + b->allele_off[l]);
        const uint16_t* ins(b->allele_ins_len + b->allele_off[l]);
        const uint32_t r0(b->read_off[l]), r1(b->read_off[l + 1]);
        for (uint32_t r = r0; r < r1; ++r)
        {
            const float* lnp(b->allele_lnp + b->lnp_off[l] + (size_t)(r - r0) * fullAlleleCount);
            std::vector<double> alleleLogLhood(fullAlleleCount);
            for (unsigned a(0); a < fullAlleleCount; ++a) alleleLogLhood[a] = lnp[a];
            const unsigned read_length(b->read_length[r]);
            const uint16_t nonAmbig(b->non_ambig[r]);
            if (callerPloidy == 1)
            {
                for (unsigned allele0Index(0); allele0Index < fullAlleleCount; ++allele0Index)
                    o.gt_lhood[allele0Index] += integrateOutMappingStatus(nonAmbig, alleleLogLhood[allele0Index]);
            }
            else
            {
                for (unsigned allele1Index(0); allele1Index < fullAlleleCount; ++allele1Index)
                {
                    for (unsigned allele0Index(0); allele0Index <= allele1Index; ++allele0Index)
                    {
                        const bool isHet(allele0Index != allele1Index);
                        const unsigned genotypeIndex(allele0Index + (allele1Index * (allele1Index + 1) / 2));
                        double rawLogLhood(0);
                        if (isHet)
                        {
                            static const double hetAlleleRatio(0.5);
                            static const double loghalf(std::log(0.5));
                            double logHetAllele0Prior(loghalf);
                            double logHetAllele1Prior(loghalf);
                            get_het_observed_allele_ratio(read_length, min_read_bp_flank, del[allele1Index - 1], ins[allele1Index - 1], hetAlleleRatio, logHetAllele0Prior,
                                                          logHetAllele1Prior);
                            if (allele0Index > 0)
                            {
                                double logRefPrior(loghalf);
                                logHetAllele0Prior = loghalf;
                                get_het_observed_allele_ratio(read_length, min_read_bp_flank, del[allele0Index - 1], ins[allele0Index - 1], hetAlleleRatio, logRefPrior,
                                                              logHetAllele0Prior);
                                const double normalizeHetRatio(getLogSumD(logHetAllele0Prior, logHetAllele1Prior));
                                logHetAllele0Prior -= normalizeHetRatio;
                                logHetAllele1Prior -= normalizeHetRatio;
                            }
                            rawLogLhood = getLogSumD(alleleLogLhood[allele0Index] + logHetAllele0Prior, alleleLogLhood[allele1Index] + logHetAllele1Prior);
                        }
                        else
                        {
                            rawLogLhood = alleleLogLhood[allele0Index];
                        }
                        o.gt_lhood[genotypeIndex] += integrateOutMappingStatus(nonAmbig, rawLogLhood);
                    }
                }
            }
            // updateSupportingReadStats
            for (double& alleleHood : alleleLogLhood) alleleHood = integrateOutMappingStatus(nonAmbig, alleleHood);
            unsigned maxIndex(0);
            normalizeLogDistro(alleleLogLhood.begin(), alleleLogLhood.end(), maxIndex);
            bool isConfidentAlleleFound(false);
            const unsigned strand(b->is_fwd[r] ? 1 : 0);
            for (unsigned alleleIndex(0); alleleIndex < fullAlleleCount; ++alleleIndex)
            {
                if (alleleLogLhood[alleleIndex] < readSupportThreshold) continue;
                o.support[strand][alleleIndex]++;
                isConfidentAlleleFound = true;
                break;
            }
            if (!isConfidentAlleleFound) o.support[strand][SX_INDEL_MAX_ALLELES + 1]++;
        }
    }
    return SX_OK;
}

// ------------------------------------------------------------------------------------------------
// f1  pileup_read_segment  starling_common/starling_pos_processor_base.cpp:1127-1421
//     create_mismatch_filter_map  starling_common/starling_read_util.cpp:52-217 (ddata + the map)
//     getReadAmbiguousEndLength   htsapi/bam_seq_read_util.cpp:29-54
//     qphred_cache::mappedq       blt_util/qscore_cache.cpp:39-48, blt_util/qscore.hh:107-113
// Per-position columns are std::vector push_backs in read order (pos_basecall_buffer.hh:118-130).
// ------------------------------------------------------------------------------------------------
namespace
{
struct mapped_q_table
{
    mapped_q_table()
    {
        for (int i(0); i <= 70; ++i)
            for (int j(0); j <= 90; ++j)
            {
                const double be(std::pow(10., -static_cast<double>(i) / 10.)); // phred_to_error_prob
                const double me(std::pow(10., -static_cast<double>(j) / 10.));
                q[j][i] = static_cast<uint8_t>(error_prob_to_qphred(((1. - me) * be) + (me * 0.75)));
            }
    }
    uint8_t q[91][71];
};

inline char char_of_bam_code(const uint8_t c) // bam_seq::get_char, htsapi/bam_seq.hh:60-80
{
    switch (c)
    {
    case 0: return '=';
    case 1: return 'A';
    case 2: return 'C';
    case 4: return 'G';
    case 8: return 'T';
    default: return 'N';
    }
}
inline bool is_match_kind(const uint8_t k) { return k == SX_SEG_MATCH; }
} // namespace

static int pileup_reads_impl(const sx_pileup_reads_batch* b, std::vector<uint16_t>* tier1 /*[n_sites]*/, std::vector<uint16_t>* tier2, uint32_t* n_spandel,
                            uint32_t* n_submapped)
{
    static const mapped_q_table mq;
    const sx_pileup_opts& opt(b->opts);
    const int64_t n_sites(static_cast<int64_t>(b->report_end) - b->report_begin);
    for (int64_t i = 0; i < n_sites; ++i)
    {
        tier1[i].clear();
        tier2[i].clear();
        n_spandel[i] = 0;
        n_submapped[i] = 0;
    }
    auto ref_char = [&](const int64_t pos) -> char { // reference_contig_segment::get_base
        const int64_t i(pos - b->ref_begin);
        return (i >= 0 && i < static_cast<int64_t>(b->ref_len)) ? b->ref[i] : 'N';
    };
    auto is_cand_snv = [&](const int64_t pos, const char readChar) -> bool {
        const int id(readChar == 'A' ? 0 : readChar == 'C' ? 1 : readChar == 'G' ? 2 : readChar == 'T' ? 3 : 4);
        if (id == 4) return false;
        const int64_t rel(pos - b->report_begin);
        if (rel < 0 || rel >= (static_cast<int64_t>(1) << 30)) return false;
        const uint32_t key((static_cast<uint32_t>(rel) << 2) | id);
        return std::binary_search(b->cand_snv, b->cand_snv + b->n_cand_snv, key);
    };
    struct rmi_t
    {
        int delta;
        bool is_mismatch, mismatch_filter_map, tier2_mismatch_filter_map;
        int mismatch_count, mismatch_count_ns;
    };
    std::vector<rmi_t> rmi;
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        const sx_pileup_read& rd(b->reads[r]);
        if (rd.flags & SX_PRF_SKIP) continue; // buffered, not piled up (the early returns of pileup_read_segment, :1147 / :1171)
        const unsigned read_size(rd.len);
        const uint8_t* seq(b->seq4 + rd.seq_off);
        const uint8_t* qual(b->qual + rd.qual_off);
        const sx_aln_seg* path(b->segs + rd.seg_off);
        const unsigned as(b->reads[r + 1].seg_off - rd.seg_off);
        const bool is_fwd(rd.flags & SX_PRF_FWD);
        auto code_at = [&](const unsigned i) -> uint8_t { return (seq[i >> 1] >> ((~i & 1) << 2)) & 0xf; };

        static const uint8_t min_adjust_mapq(5);
        const uint8_t mapq(rd.mapq);
        const uint8_t adjustedMapq(std::max(min_adjust_mapq, mapq));
        const bool is_mapq_adjust(opt.isBasecallQualAdjustedForMapq && (adjustedMapq <= 80));
        unsigned read_ref_mapped_size(0);
        for (unsigned i(0); i < as; ++i)
            if (path[i].kind == SX_SEG_MATCH || path[i].kind == SX_SEG_DELETE || path[i].kind == SX_SEG_SKIP) read_ref_mapped_size += path[i].len;
        // exact begin and end report range filters (:1189-1194)
        if (rd.pos >= b->report_end) continue;
        if (static_cast<int64_t>(rd.pos) + read_ref_mapped_size <= b->report_begin) continue;

        unsigned ambig(0); // getReadAmbiguousEndLength
        if (is_fwd)
        {
            unsigned read_end(read_size);
            while ((read_end > 0) && (char_of_bam_code(code_at(read_end - 1)) == 'N')) read_end--;
            ambig = read_size - read_end;
        }
        else
        {
            while ((ambig < read_size) && (char_of_bam_code(code_at(ambig)) == 'N')) ambig++;
        }
        unsigned read_begin(0), read_end(read_size);
        if (ambig > 0)
        {
            if (is_fwd) read_end -= ambig;
            else read_begin += ambig;
        }
        if (opt.minDistanceFromReadEdge > 0)
        {
            read_begin += opt.minDistanceFromReadEdge;
            if (opt.minDistanceFromReadEdge <= read_end) read_end -= opt.minDistanceFromReadEdge;
            else read_end = 0;
            if (read_end <= read_begin) continue;
        }
        bool is_neighbor_mismatch(false);
        const bool is_submapped(!(rd.flags & SX_PRF_TIER1OR2));
        const bool is_tier1(rd.flags & SX_PRF_TIER1);
        const bool isDensity(opt.mismatchDensityFilterFlankSize > 0);

        // get_match_edge_segments
        std::pair<unsigned, unsigned> ends(as, as);
        {
            bool isFirst(false);
            for (unsigned i(0); i < as; ++i)
                if (is_match_kind(path[i].kind))
                {
                    if (!isFirst) ends.first = i;
                    isFirst = true;
                    ends.second = i;
                }
        }
        if ((!is_submapped) && isDensity)
        {
            // create_mismatch_filter_map
            rmi.assign(read_size + 1, rmi_t{0, false, false, false, 0, 0});
            const unsigned fs(opt.mismatchDensityFilterFlankSize), fs2(fs * 2);
            const unsigned delta_size(std::max(1 + fs2, read_size) - fs2);
            auto inc = [&](const unsigned start_pos, const unsigned length) {
                rmi[std::max(fs2, start_pos) - fs2].delta += 1;
                if ((start_pos + length) < delta_size) rmi[start_pos + length].delta -= 1;
            };
            int64_t ref_head_pos(rd.pos);
            unsigned read_head_pos(0);
            for (unsigned i(0); i < as; ++i)
            {
                const sx_aln_seg& ps(path[i]);
                const bool is_edge_segment((i < ends.first) || (i > ends.second));
                if (ps.kind == SX_SEG_INSERT)
                {
                    if (!is_edge_segment) inc(read_head_pos, ps.len);
                    read_head_pos += ps.len;
                }
                else if (ps.kind == SX_SEG_DELETE)
                {
                    if (!is_edge_segment) inc(read_head_pos, 0);
                    ref_head_pos += ps.len;
                }
                else if (is_match_kind(ps.kind))
                {
                    for (unsigned j(0); j < ps.len; ++j)
                    {
                        const unsigned read_pos(read_head_pos + j);
                        if ((read_pos < read_begin) || (read_pos >= read_end)) continue;
                        const int64_t ref_pos(ref_head_pos + j);
                        const char readChar(char_of_bam_code(code_at(read_pos)));
                        if (readChar != ref_char(ref_pos))
                        {
                            if (!is_cand_snv(ref_pos, readChar))
                            {
                                rmi[read_pos].is_mismatch = true;
                                inc(read_pos, 1);
                            }
                        }
                    }
                    read_head_pos += ps.len;
                    ref_head_pos += ps.len;
                }
                else if (ps.kind == SX_SEG_SOFTCLIP) read_head_pos += ps.len;
                else if (ps.kind == SX_SEG_HARDCLIP) {}
                else if (ps.kind == SX_SEG_SKIP) return SX_ERR_ARG; // "Can't handle cigar code" (the map has no SKIP branch)
                else return SX_ERR_ARG;
            }
            for (unsigned i(1); i < delta_size; ++i) rmi[i].delta += rmi[i - 1].delta; // ddata::total
            const int max_pass(static_cast<int>(opt.mismatchDensityFilterMaxMismatchCount));
            std::vector<int> del(read_size);
            for (unsigned i(0); i < read_size; ++i) del[i] = rmi[std::min(delta_size - 1, std::max(fs, i) - fs)].delta;
            for (unsigned i(0); i < read_size; ++i)
            {
                rmi[i].mismatch_count = del[i];
                rmi[i].mismatch_count_ns = del[i] - rmi[i].is_mismatch;
                rmi[i].mismatch_filter_map = (max_pass < del[i]);
            }
            if (opt.useTier2Evidence)
            {
                const int max_pass2(opt.tier2MismatchDensityFilterMaxMismatchCount);
                for (unsigned i(0); i < read_size; ++i) rmi[i].tier2_mismatch_filter_map = (max_pass2 < rmi[i].mismatch_count);
            }
        }

        int64_t ref_head_pos(rd.pos);
        unsigned read_head_pos(0);
        for (unsigned i(0); i < as; ++i)
        {
            const sx_aln_seg& ps(path[i]);
            if (is_match_kind(ps.kind))
            {
                for (unsigned j(0); j < ps.len; ++j)
                {
                    const unsigned read_pos(read_head_pos + j);
                    if ((read_pos < read_begin) || (read_pos >= read_end)) continue;
                    const int64_t ref_pos(ref_head_pos + j);
                    if (ref_pos < b->report_begin || ref_pos >= b->report_end) continue; // is_pos_reportable
                    const uint8_t call_code(code_at(read_pos));
                    uint8_t call_id; // bam_seq_code_to_id with ref = ANY
                    switch (call_code)
                    {
                    case 1: call_id = 0; break;
                    case 2: call_id = 1; break;
                    case 4: call_id = 2; break;
                    case 8: call_id = 3; break;
                    case 0:
                    case 15: call_id = 4; break;
                    default: return SX_ERR_ARG; // base_error
                    }
                    uint8_t qscore(b->qual_bits == 4 ? b->qual_dict[(qual[read_pos >> 1] >> ((~read_pos & 1) << 2)) & 15] : qual[read_pos]);
                    if (is_mapq_adjust)
                    {
                        if (qscore > 70) return SX_ERR_RANGE; // qscore_check in get_mapped_qscore_imp
                        qscore = mq.q[std::min<int>(adjustedMapq, 90)][qscore];
                    }
                    bool current_call_filter(true), is_tier_specific_filter(false);
                    if (!is_submapped)
                    {
                        bool is_call_filter((call_code == 15) || (qscore < opt.minBasecallErrorPhredProb));
                        bool is_tier2_call_filter(is_call_filter);
                        if ((!is_call_filter) && isDensity)
                        {
                            is_call_filter = rmi[read_pos].mismatch_filter_map;
                            if (opt.useTier2Evidence) is_tier2_call_filter = rmi[read_pos].tier2_mismatch_filter_map;
                            else is_tier2_call_filter = is_call_filter;
                        }
                        current_call_filter = (is_tier1 ? is_call_filter : is_tier2_call_filter);
                        is_tier_specific_filter = (is_tier1 && is_call_filter && (!is_tier2_call_filter));
                        if (isDensity) is_neighbor_mismatch = (rmi[read_pos].mismatch_count_ns > 0);
                    }
                    const int64_t site(ref_pos - b->report_begin);
                    if (is_submapped)
                    {
                        n_submapped[site]++;
                        continue;
                    }
                    // base_call ctor (snp_pos_info.hh:51-79): the quality is clipped to 6 bits (63) BEFORE its qscore_check, which
                    // therefore cannot fire
                    const uint16_t bc(static_cast<uint16_t>((std::min<unsigned>(qscore, 63u)) | (call_id << 6) | ((is_fwd ? 1u : 0u) << 10) |
                                                            ((is_neighbor_mismatch ? 1u : 0u) << 11) | ((current_call_filter ? 1u : 0u) << 12) |
                                                            ((is_tier_specific_filter ? 1u : 0u) << 13)));
                    (is_tier1 ? tier1 : tier2)[site].push_back(bc);
                }
            }
            else if (ps.kind == SX_SEG_DELETE)
            {
                const bool is_edge_deletion((i < ends.first) || (i > ends.second));
                const bool is_pinned_deletion(((i < ends.first) && (rd.flags & SX_PRF_PIN_FIRST)) || ((i > ends.second) && (rd.flags & SX_PRF_PIN_SECOND)));
                if ((!is_edge_deletion) || is_pinned_deletion)
                {
                    for (unsigned j(0); j < ps.len; ++j)
                    {
                        const int64_t ref_pos(ref_head_pos + j);
                        if (ref_pos < b->report_begin || ref_pos >= b->report_end) continue;
                        const int64_t site(ref_pos - b->report_begin);
                        if (is_submapped) n_submapped[site]++;
                        else n_spandel[site]++;
                    }
                }
            }
            if (ps.kind == SX_SEG_MATCH || ps.kind == SX_SEG_INSERT || ps.kind == SX_SEG_SOFTCLIP) read_head_pos += ps.len; // is_segment_type_read_length
            if (ps.kind == SX_SEG_MATCH || ps.kind == SX_SEG_DELETE || ps.kind == SX_SEG_SKIP) ref_head_pos += ps.len;      // is_segment_type_ref_length
        }
    }
    return SX_OK;
}

// flat C interface: CSR columns into caller buffers (SX_ERR_NOMEM if a capacity is too small)
extern "C" int ox_pileup_reads(const sx_pileup_reads_batch* b, uint32_t* site_off, uint16_t* calls, uint64_t calls_cap, uint32_t* t2_off, uint16_t* t2_calls,
                               uint64_t t2_cap, uint32_t* n_spandel, uint32_t* n_submapped)
{
    const size_t n_sites(static_cast<size_t>(static_cast<int64_t>(b->report_end) - b->report_begin));
    std::vector<std::vector<uint16_t>> t1(n_sites), t2(n_sites);
    const int rc(pileup_reads_impl(b, t1.data(), t2.data(), n_spandel, n_submapped));
    if (rc) return rc;
    uint64_t n1(0), n2(0);
    for (size_t i = 0; i < n_sites; ++i)
    {
        site_off[i] = static_cast<uint32_t>(n1);
        t2_off[i] = static_cast<uint32_t>(n2);
        if (n1 + t1[i].size() > calls_cap || n2 + t2[i].size() > t2_cap) return SX_ERR_NOMEM;
        for (const uint16_t c : t1[i]) calls[n1++] = c;
        for (const uint16_t c : t2[i]) t2_calls[n2++] = c;
    }
    site_off[n_sites] = static_cast<uint32_t>(n1);
    t2_off[n_sites] = static_cast<uint32_t>(n2);
    return SX_OK;
}
