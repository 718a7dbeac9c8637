// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A C-ABI shim that drives the UNMODIFIED reference functions (compiled from /root/reference by oracle/build_ref.sh)
// on the inputs our tests generate.  It contains no algorithm of its own: it rebuilds the reference's native objects
// (read_segment, CandidateAlignment, IndelBuffer, snp_pos_info, ...) from flat arrays and calls
//   scoreCandidateAlignment                         starling_common/starling_read_align_score.cpp:260
//   GlobalAligner<int>::align                       alignment/GlobalAlignerImpl.hh:36
//   PileupCleaner::CleanPileupFilter/ErrorProb      starling_common/PileupCleaner.cpp:30,69
//   pprob_digt_caller::position_snp_call_pprob_digt blt_common/position_snp_call_pprob_digt.cpp:471
//   somatic_snv_caller_strand_grid::position_somatic_snv_call  applications/strelka/position_somatic_snv_strand_grid.cpp:228
// It is linked into oracle/_ref/libstrelka_ref.so, which the tests use to pin oracle/strelka_oracle.cpp and which
// bench.py may use as the "reference" CPU baseline.  Never shipped, never on the product path.

#include "../include/strelka_b200.h"
#include "flat_batch_normalize.h"

#include "alignment/GlobalAligner.hh"
#include "applications/strelka/position_somatic_snv_strand_grid.hh"
#include "applications/strelka/strelka_shared.hh"
#include "blt_common/adjust_joint_eprob.hh"
#include "blt_common/position_snp_call_pprob_digt.hh"
#include "blt_common/snp_pos_info.hh"
#include "blt_util/align_path.hh"
#include "blt_util/blt_exception.hh"
#include "blt_util/logSumUtil.hh"
#include "htsapi/align_path_bam_util.hh"
#include "htsapi/bam_record.hh"
#include "starling_common/CandidateAlignment.hh"
#include "starling_common/IndelBuffer.hh"
#include "starling_common/PileupCleaner.hh"
#include "starling_common/starling_read.hh"
#include "starling_common/starling_read_align_score.hh"
#include "test/starling_base_options_test.hh"

#include <chrono>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <vector>

namespace
{

void set_err(char* err, int errlen, const char* msg)
{
    if (err && errlen > 0)
    {
        std::strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

ALIGNPATH::align_t type_of_char(const char c)
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
    case '=': return SEQ_MATCH;
    case 'X': return SEQ_MISMATCH;
    default: return NONE;
    }
}

/// starling_base_options with the germline caller's pileup settings (applications/starling/starling_shared.hh:32-61)
struct harness_options final : public starling_base_options
{
    harness_options() {}
    const AlignmentFileOptions& getAlignmentFileOptions() const override
    {
        static AlignmentFileOptions alignFileOpt;
        if (alignFileOpt.alignmentFilenames.empty()) alignFileOpt.alignmentFilenames.push_back("sample.bam");
        return alignFileOpt;
    }
    bool is_bsnp_diploid() const override { return isBsnpDiploid; }
    bool isBsnpDiploid = true;
};

void apply_params(const sx_params& p, blt_options& opt)
{
    opt.bsnp_diploid_theta = p.bsnp_diploid_theta;
    opt.bsnp_ssd_no_mismatch = p.bsnp_ssd_no_mismatch;
    opt.bsnp_ssd_one_mismatch = p.bsnp_ssd_one_mismatch;
    opt.is_min_vexp = (p.is_min_vexp != 0);
    opt.min_vexp = p.min_vexp;
    opt.hetVariantFrequencyExtension = p.hetVariantFrequencyExtension;
}

base_call make_call(const uint16_t raw)
{
    return base_call((raw >> 6) & 15, raw & 63, (raw >> 10) & 1, 0, 0, (raw >> 12) & 1, (raw >> 11) & 1, (raw >> 13) & 1);
}

void fill_pileup(const sx_pileup_batch* b, const uint32_t s, snp_pos_info& pi)
{
    pi.clear();
    pi.set_ref_base(b->ref_base[s]);
    for (uint32_t i = b->site_off[s]; i < b->site_off[s + 1]; ++i) pi.calls.push_back(make_call(b->calls[i]));
    if (b->t2_off)
        for (uint32_t i = b->t2_off[s]; i < b->t2_off[s + 1]; ++i) pi.tier2_calls.push_back(make_call(b->t2_calls[i]));
}

} // namespace

extern "C" int ref_harness_version() { return 1; }

// ---------------------------------------------------------------------------------------------------------------
// scoreCandidateAlignment on one region
// ---------------------------------------------------------------------------------------------------------------
extern "C" int ref_score_alignments(const char* ref_seq, int ref_len, int ref_offset, int n_reads, const uint8_t* read_codes, const uint8_t* read_quals,
                                    const int* read_off, int n_alns, const int* aln_read, const int* aln_pos, const int* aln_path_off,
                                    const char* path_type, const int* path_len, const int* aln_indel_off, const int* indel_pos, const int* indel_type,
                                    const int* indel_del_len, const int* indel_ins_off, const char* ins_pool, const uint8_t* indel_is_candidate,
                                    const int* aln_leading, const int* aln_trailing, double* out_lnp, char* err, int errlen)
{
    try
    {
        harness_options opt;
        opt.is_candidate_indel_signal_test = false; // candidacy == ">= 1 tier1 supporting read" (IndelBuffer.cpp:196-210): directly controllable
        starling_base_deriv_options dopt(opt);
        reference_contig_segment ref;
        ref.seq() = std::string(ref_seq, ref_len);
        ref.set_offset(ref_offset);

        IndelBuffer indelBuffer(opt, dopt, ref);
        depth_buffer db, db2;
        indelBuffer.registerSample(db, db2, false);
        indelBuffer.finalizeSamples();

        const int n_indels(aln_indel_off[n_alns]);
        std::vector<IndelKey> keys;
        for (int i = 0; i < n_indels; ++i)
        {
            const std::string ins(ins_pool + indel_ins_off[i], ins_pool + indel_ins_off[i + 1]);
            keys.emplace_back(indel_pos[i], static_cast<INDEL::index_t>(indel_type[i]), indel_del_len[i], ins.c_str());
            IndelObservation obs;
            obs.key = keys.back();
            obs.data.id = 1 + i;
            obs.data.iat = indel_is_candidate[i] ? INDEL_ALIGN_TYPE::GENOME_TIER1_READ : INDEL_ALIGN_TYPE::GENOME_SUBMAP_READ;
            indelBuffer.addIndelObservation(0, obs);
        }

        // reads
        std::vector<std::unique_ptr<bam_record>> bams;
        std::vector<std::unique_ptr<starling_read>> sreads;
        for (int r = 0; r < n_reads; ++r)
        {
            const int len(read_off[r + 1] - read_off[r]);
            std::unique_ptr<bam_record> br(new bam_record);
            br->set_qname("R");
            const std::string dummy(len, 'A');
            br->set_readqual(dummy.c_str(), read_quals + read_off[r]);
            // write the raw 4-bit codes (set_readqual maps ASCII through get_bam_seq_code and cannot express every nibble)
            uint8_t* p(bam_get_seq(br->get_data()));
            std::memset(p, 0, (len + 1) / 2);
            for (int i = 0; i < len; ++i) p[i / 2] |= (read_codes[read_off[r] + i] & 0xf) << 4 * (1 - i % 2);
            alignment al;
            al.pos = 0;
            al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::MATCH, len));
            br->get_data()->core.pos = al.pos;
            edit_bam_cigar(al.path, *(br->get_data()));
            sreads.emplace_back(new starling_read(*br, al, MAPLEVEL::UNKNOWN, r));
            bams.push_back(std::move(br));
        }

        for (int a = 0; a < n_alns; ++a)
        {
            CandidateAlignment cal;
            cal.al.pos = aln_pos[a];
            cal.al.is_fwd_strand = true;
            for (int s = aln_path_off[a]; s < aln_path_off[a + 1]; ++s)
                cal.al.path.push_back(ALIGNPATH::path_segment(type_of_char(path_type[s]), path_len[s]));
            indel_set_t iset;
            for (int i = aln_indel_off[a]; i < aln_indel_off[a + 1]; ++i)
            {
                const int li(i - aln_indel_off[a]);
                if (li == aln_leading[a]) cal.leading_indel_key = keys[i];
                if (li == aln_trailing[a]) cal.trailing_indel_key = keys[i];
                // cal.getIndels() holds the edge keys too (addKeysToCandidateAlignment, starling_read_align.cpp:801-802) -- and
                // getMatchingIndelKey may need them for an INTERIOR gap: make_start_pos_alignment labels a deletion that follows a
                // mismatch segment without a match in between as leading_indel_key (:552-555)
                iset.insert(keys[i]);
            }
            cal.setIndels(iset);
            const read_segment& rseg(sreads[aln_read[a]]->get_full_segment());
            out_lnp[a] = scoreCandidateAlignment(opt, indelBuffer, rseg, cal, ref);
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err(err, errlen, e.what());
        return 1;
    }
    catch (...)
    {
        set_err(err, errlen, "unknown exception");
        return 2;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GlobalAligner<ScoreType>::align
// ---------------------------------------------------------------------------------------------------------------
template <typename ScoreType>
static void ga_run(const sx_ga_scores* s, const sx_ga_batch* b, sx_ga_result* res, uint32_t* cigar)
{
    const AlignmentScores<ScoreType> scores(s->match, s->mismatch, s->open, s->extend, s->offEdge, s->insertDelete, s->isAllowEdgeInsertion != 0,
                                            s->isRequireEdgeDeletion != 0);
    const GlobalAligner<ScoreType> aligner(scores);
    for (uint32_t i = 0; i < b->n; ++i)
    {
        const std::string q(b->query + b->query_off[i], b->query + b->query_off[i + 1]);
        const std::string r(b->ref + b->ref_off[i], b->ref + b->ref_off[i + 1]);
        AlignmentResult<ScoreType> result;
        aligner.align(q.begin(), q.end(), r.begin(), r.end(), result);
        res[i].score = result.score;
        res[i].beginPos = result.align.beginPos;
        const ALIGNPATH::path_t& ap(result.align.apath);
        res[i].n_ops = ap.size();
        res[i].status = (ap.size() > b->max_ops) ? 1 : 0;
        for (uint32_t k = 0; k < ap.size() && k < b->max_ops; ++k)
        {
            static const char* codes = "MIDNSHP=X";
            const char c(segment_type_to_cigar_code(ap[k].type));
            const char* f(std::strchr(codes, c));
            cigar[(size_t)i * b->max_ops + k] = (ap[k].length << 4) | (uint32_t)(f ? (f - codes) : 15);
        }
    }
}

extern "C" int ref_global_align(const sx_ga_scores* s, const sx_ga_batch* b, int use_short_scores, sx_ga_result* res, uint32_t* cigar, char* err, int errlen)
{
    try
    {
        if (use_short_scores) ga_run<short>(s, b, res, cigar); // the score type of alignment/test/GlobalAlignerTest.cpp:43
        else ga_run<int>(s, b, res, cigar);                    // the production type (ActiveRegionDetector.hh:171)
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// germline site model
// ---------------------------------------------------------------------------------------------------------------
extern "C" int ref_dependent_eprob(const sx_params* p, const sx_pileup_batch* b, uint32_t* out_off, float* de, char* err, int errlen)
{
    try
    {
        harness_options opt;
        opt.isBsnpDiploid = (p->is_bsnp_diploid != 0);
        apply_params(*p, opt);
        PileupCleaner cleaner(opt);
        snp_pos_info pi;
        CleanedPileup cpi;
        uint32_t off(0);
        for (uint32_t s = 0; s < b->n_sites; ++s)
        {
            fill_pileup(b, s, pi);
            cleaner.CleanPileupFilter(pi, false, cpi);
            cleaner.CleanPileupErrorProb(cpi);
            out_off[s] = off;
            for (float v : static_cast<const CleanedPileup&>(cpi).dependentErrorProb()) de[off++] = v;
        }
        out_off[b->n_sites] = off;
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// secs (may be NULL): accumulates the time inside the reference's own calls per site -- CleanPileupFilter + CleanPileupErrorProb +
// position_snp_call_pprob_digt, what computeSampleDiploidSiteGenotype's path spends -- leaving out this harness's snp_pos_info filling and
// its second get_diploid_gt_lhood call (made only to export the float likelihoods)
extern "C" int ref_site_gl_germline_timed(const sx_params* p, const sx_pileup_batch* b, int is_always_test, sx_digt_result* out, double* secs, char* err, int errlen);
extern "C" int ref_site_gl_germline(const sx_params* p, const sx_pileup_batch* b, int is_always_test, sx_digt_result* out, char* err, int errlen)
{
    return ref_site_gl_germline_timed(p, b, is_always_test, out, nullptr, err, errlen);
}
extern "C" int ref_site_gl_germline_timed(const sx_params* p, const sx_pileup_batch* b, int is_always_test, sx_digt_result* out, double* secs, char* err, int errlen)
{
    try
    {
        harness_options opt;
        opt.isBsnpDiploid = (p->is_bsnp_diploid != 0);
        apply_params(*p, opt);
        PileupCleaner cleaner(opt);
        const pprob_digt_caller caller(opt.bsnp_diploid_theta);
        snp_pos_info pi;
        CleanedPileup cpi;
        for (uint32_t s = 0; s < b->n_sites; ++s)
        {
            sx_digt_result& o(out[s]);
            std::memset(&o, 0, sizeof(o));
            fill_pileup(b, s, pi);
            const auto t0(std::chrono::steady_clock::now());
            cleaner.CleanPileupFilter(pi, false, cpi);
            cleaner.CleanPileupErrorProb(cpi);
            o.n_used_calls = cpi.usedBasecallCount();
            diploid_genotype dgt;
            dgt.ploidy = b->ploidy ? b->ploidy[s] : 2;
            caller.position_snp_call_pprob_digt(opt, cpi.getExtendedPosInfo(), dgt, is_always_test != 0);
            if (secs) *secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            o.ref_gt = dgt.ref_gt;
            o.strand_bias = dgt.strand_bias;
            const diploid_genotype::result_set* rsin[2] = {&dgt.genome, &dgt.poly};
            sx_digt_result_set* rsout[2] = {&o.genome, &o.poly};
            for (int k = 0; k < 2; ++k)
            {
                rsout[k]->max_gt = rsin[k]->max_gt;
                rsout[k]->ref_pprob = rsin[k]->ref_pprob;
                rsout[k]->snp_qphred = rsin[k]->snp_qphred;
                rsout[k]->max_gt_qphred = rsin[k]->max_gt_qphred;
            }
            for (unsigned gt = 0; gt < 10; ++gt) o.phredLoghood[gt] = dgt.phredLoghood[gt];
            // the reference does not export lhood[]; recompute it through its public static for the float comparison
            bool computed(b->ref_base[s] != 'N');
            if (computed && !is_always_test)
            {
                computed = false;
                for (const base_call& bc : static_cast<const CleanedPileup&>(cpi).cleanedPileup().calls)
                    if (bc.base_id != dgt.ref_gt) computed = true;
            }
            o.is_computed = computed;
            if (computed)
            {
                blt_float_t lhood[DIGT::SIZE];
                pprob_digt_caller::get_diploid_gt_lhood(opt, cpi.getExtendedPosInfo(), false, 0, lhood);
                for (unsigned gt = 0; gt < 10; ++gt) o.lhood[gt] = lhood[gt];
            }
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// somatic site model
// ---------------------------------------------------------------------------------------------------------------
extern "C" int ref_site_gl_somatic(const sx_params* p, const sx_pileup_batch* normal, const sx_pileup_batch* tumor, const uint8_t* is_forced_output,
                                   sx_ssnv_result* out, char* err, int errlen)
{
    try
    {
        strelka_options opt;
        apply_params(*p, opt);
        opt.somatic_snv_rate = p->somatic_snv_rate;
        opt.shared_site_error_rate = p->shared_site_error_rate;
        opt.shared_site_error_strand_bias_fraction = p->shared_site_error_strand_bias_fraction;
        opt.ssnv_contam_tolerance = p->ssnv_contam_tolerance;
        const somatic_snv_caller_strand_grid caller(opt);
        PileupCleaner cleaner(opt);
        const bool is_tier2(normal->t2_off != nullptr && tumor->t2_off != nullptr);
        snp_pos_info npi, tpi;
        CleanedPileup ncpi[2], tcpi[2];
        for (uint32_t s = 0; s < normal->n_sites; ++s)
        {
            sx_ssnv_result& o(out[s]);
            std::memset(&o, 0, sizeof(o));
            fill_pileup(normal, s, npi);
            fill_pileup(tumor, s, tpi);
            // strelka_pos_processor.cpp:183-189
            cleaner.CleanPileupFilter(npi, false, ncpi[0]);
            cleaner.CleanPileupFilter(tpi, false, tcpi[0]);
            if (is_tier2)
            {
                cleaner.CleanPileupFilter(npi, true, ncpi[1]);
                cleaner.CleanPileupFilter(tpi, true, tcpi[1]);
            }
            for (int t = 0; t < (is_tier2 ? 2 : 1); ++t)
            {
                cleaner.CleanPileupErrorProb(ncpi[t]);
                cleaner.CleanPileupErrorProb(tcpi[t]);
            }
            somatic_snv_genotype_grid sgt;
            sgt.is_forced_output = is_forced_output ? (is_forced_output[s] != 0) : false;
            const extended_pos_info* nt2(is_tier2 ? &ncpi[1].getExtendedPosInfo() : nullptr);
            const extended_pos_info* tt2(is_tier2 ? &tcpi[1].getExtendedPosInfo() : nullptr);
            caller.position_somatic_snv_call(ncpi[0].getExtendedPosInfo(), tcpi[0].getExtendedPosInfo(), nt2, tt2, false, sgt);
            o.ref_gt = sgt.ref_gt;
            o.is_computed = (sgt.is_forced_output || sgt.is_output()) ? 1 : 0;
            o.snv_tier = sgt.snv_tier;
            o.snv_from_ntype_tier = sgt.snv_from_ntype_tier;
            o.ntype = sgt.rs.ntype;
            o.max_gt = sgt.rs.max_gt;
            o.qphred = sgt.rs.qphred;
            o.from_ntype_qphred = sgt.rs.from_ntype_qphred;
            o.normal_alt_id = sgt.rs.normal_alt_id;
            o.tumor_alt_id = sgt.rs.tumor_alt_id;
            o.strandBias = sgt.rs.strandBias;
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// small numerics probes (blt_util/logSumUtil.hh) used to pin the float log-sum in the strand-state likelihood
extern "C" float ref_getLogSum_float(float a, float b) { return getLogSum(a, b); }
extern "C" double ref_getLogSum_double(double a, double b) { return getLogSum(a, b); }

// ---------------------------------------------------------------------------------------------------------------
// scoreCandidateAlignment over regions [r0, r1) of a FLATTENED batch (bench.py --impl reference).
// The reference objects (IndelBuffer, CandidateAlignment, read_segment) are rebuilt from the flattened description -- the inverse of
// the host-side flattening -- and only the scoreCandidateAlignment calls themselves are timed (score_seconds), so the reported CPU
// number is the reference's own scoring loop (starling_read_align.cpp:1564-1571), not this shim's object construction.
// Supported segment kinds: MATCH, INSERT (internal, leading/trailing edge, or followed by a REFSKIP = swap), REFSKIP (as DELETE),
// SOFTCLIP, HARDCLIP; zero-length segments are dropped.
// ---------------------------------------------------------------------------------------------------------------
#include <chrono>

extern "C" int ref_score_flat_batch(const sx_align_batch* b, uint32_t r0, uint32_t r1, double* out_lnp, double* score_seconds, char* err, int errlen)
{
    try
    {
        sx_norm_batch wide; // compact wire formats are widened first (same values, wider fields)
        b = sx_normalize_range(b, r0, r1, wide);
        harness_options opt;
        opt.is_candidate_indel_signal_test = false;
        starling_base_deriv_options dopt(opt);
        double secs(0);
        for (uint32_t ri = r0; ri < r1; ++ri)
        {
            const sx_region& reg(b->regions[ri]);
            const sx_region& nxt(b->regions[ri + 1]);
            reference_contig_segment ref;
            ref.seq() = std::string(b->ref + reg.ref_off, reg.ref_len);
            ref.set_offset(reg.ref_begin);
            IndelBuffer indelBuffer(opt, dopt, ref);
            depth_buffer db, db2;
            indelBuffer.registerSample(db, db2, false);
            indelBuffer.finalizeSamples();

            std::vector<std::unique_ptr<bam_record>> bams;
            std::vector<std::unique_ptr<starling_read>> sreads;
            uint64_t so(reg.seq_off), qo(reg.qual_off);
            for (uint32_t r = reg.read_begin; r < nxt.read_begin; ++r)
            {
                const int len(b->read_len[r]);
                std::unique_ptr<bam_record> br(new bam_record);
                br->set_qname("R");
                const std::string dummy(len, 'A');
                if (b->qual_bits == 4)
                {
                    std::vector<uint8_t> q8(len);
                    for (int i = 0; i < len; ++i) q8[i] = b->qual_dict[((b->qual + qo)[i >> 1] >> ((~i & 1) << 2)) & 0xf];
                    br->set_readqual(dummy.c_str(), q8.data());
                }
                else br->set_readqual(dummy.c_str(), b->qual + qo);
                std::memcpy(bam_get_seq(br->get_data()), b->seq4 + so, (len + 1) / 2);
                alignment al;
                al.pos = 0;
                al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::MATCH, len));
                br->get_data()->core.pos = al.pos;
                edit_bam_cigar(al.path, *(br->get_data()));
                sreads.emplace_back(new starling_read(*br, al, MAPLEVEL::UNKNOWN, r));
                bams.push_back(std::move(br));
                so += (len + 1) / 2;
                qo += (b->qual_bits == 4) ? (len + 1) / 2 : len;
            }
            std::vector<CandidateAlignment> cals(nxt.aln_begin - reg.aln_begin);
            for (uint32_t a = reg.aln_begin; a < nxt.aln_begin; ++a)
            {
                const sx_aln& A(b->alns[a]);
                CandidateAlignment& cal(cals[a - reg.aln_begin]);
                cal.al.pos = A.ref_pos;
                cal.al.is_fwd_strand = true;
                indel_set_t iset;
                pos_t ref_head(A.ref_pos);
                const char* ins(b->ins + A.ins_off);
                bool seenMatch(false);
                const uint32_t s1(b->alns[a + 1].seg_off);
                // last MATCH segment, to tell trailing-edge insertions
                int lastMatch(-1);
                for (uint32_t s = A.seg_off; s < s1; ++s)
                    if (b->segs[s].kind == SX_SEG_MATCH && b->segs[s].len) lastMatch = (int)s;
                for (uint32_t s = A.seg_off; s < s1; ++s)
                {
                    const sx_aln_seg& sg(b->segs[s]);
                    if (sg.len == 0) continue;
                    const bool cand(!(sg.flags & SX_SEGF_NONCANDIDATE));
                    auto observe = [&](const IndelKey& k, const bool isCand) {
                        IndelObservation obs;
                        obs.key = k;
                        obs.data.id = 1 + a;
                        obs.data.iat = isCand ? INDEL_ALIGN_TYPE::GENOME_TIER1_READ : INDEL_ALIGN_TYPE::GENOME_SUBMAP_READ;
                        indelBuffer.addIndelObservation(0, obs);
                    };
                    if (sg.kind == SX_SEG_MATCH)
                    {
                        cal.al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::MATCH, sg.len));
                        ref_head += sg.len;
                        seenMatch = true;
                    }
                    else if (sg.kind == SX_SEG_INSERT)
                    {
                        const std::string seq(ins, ins + sg.len);
                        ins += sg.len;
                        unsigned del(0);
                        if (s + 1 < s1 && b->segs[s + 1].kind == SX_SEG_REFSKIP && b->segs[s + 1].len && seenMatch && (int)s < lastMatch) del = b->segs[s + 1].len;
                        const IndelKey k(ref_head, INDEL::INDEL, del, seq.c_str());
                        observe(k, cand && del == 0);
                        cal.al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::INSERT, sg.len));
                        if (!seenMatch) cal.leading_indel_key = k;
                        else if ((int)s > lastMatch) cal.trailing_indel_key = k;
                        else iset.insert(k);
                        if (del)
                        {
                            cal.al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::DELETE, del));
                            ref_head += del;
                            ++s;
                        }
                    }
                    else if (sg.kind == SX_SEG_REFSKIP)
                    {
                        const IndelKey k(ref_head, INDEL::INDEL, sg.len, "");
                        observe(k, cand);
                        iset.insert(k);
                        cal.al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::DELETE, sg.len));
                        ref_head += sg.len;
                    }
                    else if (sg.kind == SX_SEG_SOFTCLIP) cal.al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::SOFT_CLIP, sg.len));
                    else if (sg.kind == SX_SEG_HARDCLIP) cal.al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::HARD_CLIP, sg.len));
                }
                cal.setIndels(iset);
            }
            const auto t0(std::chrono::steady_clock::now());
            for (uint32_t a = reg.aln_begin; a < nxt.aln_begin; ++a)
            {
                const read_segment& rseg(sreads[b->alns[a].read - reg.read_begin]->get_full_segment());
                out_lnp[a] = scoreCandidateAlignment(opt, indelBuffer, rseg, cals[a - reg.aln_begin], ref);
            }
            secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (score_seconds) *score_seconds = secs;
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err(err, errlen, e.what());
        return 1;
    }
    catch (...)
    {
        set_err(err, errlen, "unknown exception");
        return 2;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// indel genotype likelihoods: getVariantAlleleGroupGenotypeLhoodsForSample (starling_common/AlleleGroupGenotype.cpp:184)
// The allele group is rebuilt in an IndelBuffer; every read gets a ReadPathScores entry under every allele whose `ref` is the
// read's reference-path score and whose `indel` is that allele's score, which is what getAlleleLogLhoodFromRead turns back into
// the flat per-read vector {ref, alt1, ...} the batch carries.
// ---------------------------------------------------------------------------------------------------------------
#include "starling_common/AlleleGroupGenotype.hh"
#include "starling_common/OrthogonalVariantAlleleCandidateGroup.hh"

extern "C" int ref_indel_gl(const sx_params* p, const sx_indel_batch* b, sx_indel_result* out, char* err, int errlen)
{
    try
    {
        harness_options opt;
        opt.is_candidate_indel_signal_test = false;
        opt.randomBaseMatchProb = p->randomBaseMatchProb;
        opt.default_min_read_bp_flank = p->min_read_bp_flank;
        {
            char buf[64];
            snprintf(buf, sizeof(buf), "%.17g", p->readConfidentSupportThreshold);
            opt.readConfidentSupportThreshold.update(buf);
        }
        starling_base_deriv_options dopt(opt);
        starling_sample_options sample_opt(opt);
        for (uint32_t l = 0; l < b->n_loci; ++l)
        {
            sx_indel_result& o(out[l]);
            std::memset(&o, 0, sizeof(o));
            const unsigned A(b->allele_off[l + 1] - b->allele_off[l]);
            reference_contig_segment ref;
            ref.seq() = std::string(2000, 'A');
            IndelBuffer indelBuffer(opt, dopt, ref);
            depth_buffer db, db2;
            indelBuffer.registerSample(db, db2, false);
            indelBuffer.finalizeSamples();
            std::vector<IndelKey> keys;
            for (unsigned a = 0; a < A; ++a)
            {
                const unsigned dl(b->allele_del_len[b->allele_off[l] + a]), il(b->allele_ins_len[b->allele_off[l] + a]);
                // distinct keys: vary the insert sequence per allele index so equal-length alleles do not collide
                std::string ins(il, "ACGT"[a & 3]);
                keys.emplace_back(500 + (int)a, INDEL::INDEL, dl, ins.c_str());
                IndelObservation obs;
                obs.key = keys.back();
                obs.data.id = 1;
                obs.data.iat = INDEL_ALIGN_TYPE::GENOME_TIER1_READ;
                indelBuffer.addIndelObservation(0, obs);
            }
            const uint32_t r0(b->read_off[l]), r1(b->read_off[l + 1]);
            OrthogonalVariantAlleleCandidateGroup group, contrast;
            for (unsigned a = 0; a < A; ++a)
            {
                IndelBuffer::iterator it(indelBuffer.getIndelIter(keys[a]));
                IndelSampleData& isd(getIndelData(it).getSampleData(0));
                for (uint32_t r = r0; r < r1; ++r)
                {
                    const float* lnp(b->allele_lnp + b->lnp_off[l] + (size_t)(r - r0) * (A + 1));
                    isd.read_path_lnp[r - r0] = ReadPathScores(lnp[0], lnp[1 + a], b->non_ambig[r], b->read_length[r], true, b->is_fwd[r] != 0, 0, 0);
                }
                group.addVariantAllele(indelBuffer.getIndelIter(keys[a]));
            }
            std::vector<double> gl;
            LocusSupportingReadStats stats;
            getVariantAlleleGroupGenotypeLhoodsForSample(opt, dopt, sample_opt, b->ploidy[l], 0, group, contrast, gl, stats);
            o.n_gt = gl.size();
            for (size_t g = 0; g < gl.size() && g < SX_INDEL_MAX_GT; ++g) o.gt_lhood[g] = gl[g];
            for (int s = 0; s < 2; ++s)
            {
                const SupportingReadCountGroup& c(stats.getCounts(s == 1));
                for (unsigned a = 0; a <= A; ++a) o.support[s][a] = c.confidentAlleleCount(a);
                o.support[s][SX_INDEL_MAX_ALLELES + 1] = c.nonConfidentCount;
            }
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err(err, errlen, e.what());
        return 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// starling_pos_processor_base::pileup_read_segment (starling_pos_processor_base.cpp:1127-1421) driven for every read of a
// flattened sx_pileup_reads_batch, on a real (minimal) pos processor; the per-position buffers are then read back.
// pileup_read_segment is private: it is reached through the explicit-instantiation idiom (no source is modified, no macro
// redefines keywords).
// ---------------------------------------------------------------------------------------------------------------
#include "appstats/RunStatsManager.hh"
#include "starling_common/starling_pos_processor_base.hh"
#include "starling_common/starling_streams_base.hh"

namespace
{
struct HarnessPosProcessor : public starling_pos_processor_base
{
    using starling_pos_processor_base::starling_pos_processor_base;
    void process_pos_variants_impl(const pos_t, const bool) override {}
    void resetRegionForTest(const std::string& chrom, const known_pos_range2& range) { resetRegionBase(chrom, range); }
    CandidateSnvBuffer& snvBuffer() { return getCandidateSnvBuffer(); }
};

template <typename Tag, typename Tag::type M> struct MemberPtrOf
{
    friend typename Tag::type memberPtr(Tag) { return M; }
};
struct PileupReadSegmentTag
{
    typedef void (starling_pos_processor_base::*type)(const read_segment&, const unsigned);
    friend type memberPtr(PileupReadSegmentTag);
};
template struct MemberPtrOf<PileupReadSegmentTag, &starling_pos_processor_base::pileup_read_segment>;
} // namespace

// secs (may be NULL): accumulates the time inside pileup_read_segment itself (the bam_record / starling_read objects around it are this harness's)
extern "C" int ref_pileup_reads_timed(const sx_pileup_reads_batch* b, uint32_t* site_off, uint16_t* calls, uint64_t calls_cap, uint32_t* t2_off, uint16_t* t2_calls,
                                      uint64_t t2_cap, uint32_t* n_spandel, uint32_t* n_submapped, double* secs, char* err, int errlen);
extern "C" int ref_pileup_reads(const sx_pileup_reads_batch* b, uint32_t* site_off, uint16_t* calls, uint64_t calls_cap, uint32_t* t2_off, uint16_t* t2_calls,
                                uint64_t t2_cap, uint32_t* n_spandel, uint32_t* n_submapped, char* err, int errlen)
{
    return ref_pileup_reads_timed(b, site_off, calls, calls_cap, t2_off, t2_calls, t2_cap, n_spandel, n_submapped, nullptr, err, errlen);
}
extern "C" int ref_pileup_reads_timed(const sx_pileup_reads_batch* b, uint32_t* site_off, uint16_t* calls, uint64_t calls_cap, uint32_t* t2_off, uint16_t* t2_calls,
                                      uint64_t t2_cap, uint32_t* n_spandel, uint32_t* n_submapped, double* secs, char* err, int errlen)
{
    try
    {
        harness_options opt;
        opt.is_candidate_indel_signal_test = false;
        opt.isBasecallQualAdjustedForMapq = (b->opts.isBasecallQualAdjustedForMapq != 0);
        opt.minBasecallErrorPhredProb = b->opts.minBasecallErrorPhredProb;
        opt.mismatchDensityFilterFlankSize = b->opts.mismatchDensityFilterFlankSize;
        opt.mismatchDensityFilterMaxMismatchCount = b->opts.mismatchDensityFilterMaxMismatchCount;
        opt.useTier2Evidence = (b->opts.useTier2Evidence != 0);
        opt.tier2.mismatchDensityFilterMaxMismatchCount = b->opts.tier2MismatchDensityFilterMaxMismatchCount;
        opt.minDistanceFromReadEdge = b->opts.minDistanceFromReadEdge;
        starling_base_deriv_options dopt(opt);
        reference_contig_segment ref;
        ref.seq() = std::string(b->ref, b->ref_len);
        ref.set_offset(b->ref_begin);
        starling_streams_base streams(1);
        RunStatsManager stats("");
        HarnessPosProcessor proc(opt, dopt, ref, streams, 1, stats);
        proc.resetRegionForTest("chrT", known_pos_range2(b->report_begin, b->report_end));
        for (uint32_t i = 0; i < b->n_cand_snv; ++i)
        {
            static const char BASE[4] = {'A', 'C', 'G', 'T'};
            proc.snvBuffer().addCandidateSnv(0, b->report_begin + static_cast<pos_t>(b->cand_snv[i] >> 2), BASE[b->cand_snv[i] & 3], 1, 0.5f);
        }
        const auto pileup(memberPtr(PileupReadSegmentTag()));
        std::vector<std::unique_ptr<bam_record>> bams;
        std::vector<std::unique_ptr<starling_read>> sreads;
        for (uint32_t r = 0; r < b->n_reads; ++r)
        {
            const sx_pileup_read& rd(b->reads[r]);
            if (rd.flags & SX_PRF_SKIP) continue;
            const int len(rd.len);
            std::unique_ptr<bam_record> br(new bam_record);
            br->set_qname("R");
            const std::string dummy(len, 'A');
            if (b->qual_bits == 4) // dictionary-coded qualities, two per byte: widened for the reference's bam_record
            {
                std::vector<uint8_t> wide(len);
                for (int i = 0; i < len; ++i) wide[i] = b->qual_dict[(b->qual[rd.qual_off + (i >> 1)] >> ((~i & 1) << 2)) & 15];
                br->set_readqual(dummy.c_str(), wide.data());
            }
            else
                br->set_readqual(dummy.c_str(), b->qual + rd.qual_off);
            std::memcpy(bam_get_seq(br->get_data()), b->seq4 + rd.seq_off, (len + 1) / 2);
            alignment al;
            al.pos = rd.pos;
            al.is_fwd_strand = (rd.flags & SX_PRF_FWD) != 0;
            for (uint32_t s = rd.seg_off; s < b->reads[r + 1].seg_off; ++s)
            {
                const sx_aln_seg& sg(b->segs[s]);
                if (sg.len == 0 && sg.kind == SX_SEG_HARDCLIP) continue; // a pad of K9's slot layout: not a segment of the path
                ALIGNPATH::align_t t(ALIGNPATH::NONE);
                switch (sg.kind)
                {
                case SX_SEG_MATCH: t = ALIGNPATH::MATCH; break;
                case SX_SEG_INSERT: t = ALIGNPATH::INSERT; break;
                case SX_SEG_DELETE: t = ALIGNPATH::DELETE; break;
                case SX_SEG_SKIP: t = ALIGNPATH::SKIP; break;
                case SX_SEG_SOFTCLIP: t = ALIGNPATH::SOFT_CLIP; break;
                case SX_SEG_HARDCLIP: t = ALIGNPATH::HARD_CLIP; break;
                default: throw blt_exception("ref_pileup_reads: unknown segment kind");
                }
                al.path.push_back(ALIGNPATH::path_segment(t, sg.len));
            }
            br->get_data()->core.pos = al.pos;
            br->get_data()->core.qual = rd.mapq;
            edit_bam_cigar(al.path, *(br->get_data()));
            const MAPLEVEL::index_t lev((rd.flags & SX_PRF_TIER1) ? MAPLEVEL::TIER1_MAPPED : (rd.flags & SX_PRF_TIER1OR2) ? MAPLEVEL::TIER2_MAPPED : MAPLEVEL::SUB_MAPPED);
            sreads.emplace_back(new starling_read(*br, al, lev, r));
            bams.push_back(std::move(br));
            const auto t0(std::chrono::steady_clock::now());
            (proc.*pileup)(sreads.back()->get_full_segment(), 0);
            if (secs) *secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        const pos_basecall_buffer& buf(proc.sample(0).basecallBuffer);
        const uint32_t n_sites(static_cast<uint32_t>(b->report_end - b->report_begin));
        uint64_t n1(0), n2(0);
        for (uint32_t i = 0; i < n_sites; ++i)
        {
            const snp_pos_info& pi(buf.get_pos(b->report_begin + static_cast<pos_t>(i)));
            site_off[i] = static_cast<uint32_t>(n1);
            t2_off[i] = static_cast<uint32_t>(n2);
            for (const base_call& bc : pi.calls)
            {
                if (n1 >= calls_cap) throw blt_exception("ref_pileup_reads: calls capacity");
                std::memcpy(&calls[n1], &bc, 2);
                calls[n1++] &= 0x3fffu; // the bit-field struct has 14 used bits; the two padding bits are indeterminate
            }
            for (const base_call& bc : pi.tier2_calls)
            {
                if (n2 >= t2_cap) throw blt_exception("ref_pileup_reads: tier2 capacity");
                std::memcpy(&t2_calls[n2], &bc, 2);
                t2_calls[n2++] &= 0x3fffu;
            }
            n_spandel[i] = pi.spanningDeletionReadCount;
            n_submapped[i] = pi.submappedReadCount;
        }
        site_off[n_sites] = static_cast<uint32_t>(n1);
        t2_off[n_sites] = static_cast<uint32_t>(n2);
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err(err, errlen, e.what());
        return 1;
    }
}
