// ref_harness_score_indels.cpp -- TEST INFRASTRUCTURE ONLY (second translation unit of oracle/_ref/libstrelka_ref.so).
//
// Drives the UNMODIFIED reference score_indels (starling_common/starling_read_align_score_indels.cpp:454) on a flattened
// sx_score_indels_batch: the IndelBuffer, the read segments and the std::set<CandidateAlignment> of every read are rebuilt from the
// flat arrays, score_indels is called with the given alignment scores, and what it wrote into the IndelBuffer (read_path_lnp,
// suboverlap_tier{1,2}_read_ids) is read back as sx_read_indel_score records.
//
// The maximum-alignment choice that precedes score_indels (scoreCandidateAlignments, starling_read_align.cpp:1573-1593) is a loop
// around file-static functions; to reach the reference's own tie-break isFirstCandidateAlignmentPreferred (:1352) this unit
// includes that source file (nothing is copied or modified; the archive member of the same file is then simply not pulled at
// link time).  The six-line loop itself has to be spelled here because the function that contains it also computes the scores.
// Indel error rates are injected through IndelSampleData's private _errorRates with the explicit-instantiation idiom.

#include "../include/strelka_b200.h"

#include "starling_common/starling_read_align.cpp" // NOLINT: see above

#include "htsapi/align_path_bam_util.hh"
#include "htsapi/bam_record.hh"
#include "starling_common/starling_read.hh"
#include "starling_common/starling_read_align_score_indels.hh"
#include "test/starling_base_options_test.hh"

#include <chrono>
#include <cstring>
#include <iterator>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace
{

void set_err2(char* err, int errlen, const char* msg)
{
    if (err && errlen > 0)
    {
        std::strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

struct harness_options2 final : public starling_base_options
{
    const AlignmentFileOptions& getAlignmentFileOptions() const override
    {
        static AlignmentFileOptions alignFileOpt;
        if (alignFileOpt.alignmentFilenames.empty()) alignFileOpt.alignmentFilenames.push_back("sample.bam");
        return alignFileOpt;
    }
    bool is_bsnp_diploid() const override { return true; }
};

template <typename Tag, typename Tag::type M> struct MemberPtrOf2
{
    friend typename Tag::type memberPtr(Tag) { return M; }
};
struct ErrorRatesTag
{
    typedef IndelErrorRates IndelSampleData::*type;
    friend type memberPtr(ErrorRatesTag);
};
template struct MemberPtrOf2<ErrorRatesTag, &IndelSampleData::_errorRates>;

IndelKey key_of(const sx_indel_key& k, const char* ins_pool, const uint32_t* ins_off, const uint32_t global_index)
{
    const std::string ins(ins_pool + ins_off[global_index], ins_pool + ins_off[global_index + 1]);
    if (ins.size() != k.ins_len) throw blt_exception("ref_score_indels: insert sequence length does not match ins_len");
    return IndelKey(k.pos, (k.type == SX_INDEL_TYPE_MISMATCH) ? INDEL::MISMATCH : INDEL::INDEL, k.del_len, ins.c_str());
}

void path_of(const sx_score_indels_batch* b, const uint32_t a, ALIGNPATH::path_t& path)
{
    for (uint32_t s = b->aln_seg_off[a]; s < b->aln_seg_off[a + 1]; ++s)
    {
        const sx_aln_seg& sg(b->segs[s]);
        ALIGNPATH::align_t t(ALIGNPATH::NONE);
        switch (sg.kind)
        {
        case SX_SEG_MATCH: t = ALIGNPATH::MATCH; break;
        case SX_SEG_INSERT: t = ALIGNPATH::INSERT; break;
        case SX_SEG_DELETE: t = ALIGNPATH::DELETE; break;
        case SX_SEG_SOFTCLIP: t = ALIGNPATH::SOFT_CLIP; break;
        case SX_SEG_HARDCLIP: t = ALIGNPATH::HARD_CLIP; break;
        default: throw blt_exception("ref_score_indels: segment kind outside score_indels' domain");
        }
        path.push_back(ALIGNPATH::path_segment(t, sg.len));
    }
}

void cal_of(const sx_score_indels_batch* b, const uint32_t a, const bool fwd, const std::vector<IndelKey>& winKeys, CandidateAlignment& cal)
{
    cal.al.pos = b->aln_pos[a];
    cal.al.is_fwd_strand = fwd;
    path_of(b, a, cal.al.path);
    indel_set_t iset;
    for (uint32_t i = b->aln_key_off[a]; i < b->aln_key_off[a + 1]; ++i) iset.insert(winKeys.at(b->aln_keys[i]));
    cal.setIndels(iset);
}

} // namespace

// perm[a0 + rank] = index of the alignment that std::set<CandidateAlignment> iterates at position `rank` among read r's alignments
extern "C" int ref_candidate_alignment_order(const sx_score_indels_batch* b, const char* ins_pool, const uint32_t* ins_off, uint32_t* perm, char* err, int errlen)
{
    try
    {
        for (uint32_t region = 0; region < b->n_regions; ++region)
        {
            const uint32_t k0(b->region_key_off[region]), k1(b->region_key_off[region + 1]);
            std::vector<IndelKey> winKeys;
            for (uint32_t k = k0; k < k1; ++k) winKeys.push_back(key_of(b->keys[k], ins_pool, ins_off, k));
            for (uint32_t r = b->region_read_off[region]; r < b->region_read_off[region + 1]; ++r)
            {
                const uint32_t a0(b->aln_off[r]), a1(b->aln_off[r + 1]);
                std::map<CandidateAlignment, uint32_t> order;
                for (uint32_t a = a0; a < a1; ++a)
                {
                    CandidateAlignment cal;
                    cal_of(b, a, (b->read_flags[r] & SX_SIF_FWD) != 0, winKeys, cal);
                    if (!order.insert(std::make_pair(cal, a)).second) throw blt_exception("ref_candidate_alignment_order: duplicate candidate alignment");
                }
                uint32_t rank(a0);
                for (const auto& kv : order) perm[rank++] = kv.second;
            }
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err2(err, errlen, e.what());
        return 1;
    }
}

// allow_unordered = 0: the flat alignment order must be the std::set's (tests).  1: the alignments of a read are put into set order
// here and exact duplicates are dropped, as the reference's container would (bench.py: the synthetic K1 workload lists one
// alignment per haplotype in haplotype order); max_aln still reports the flat index.
extern "C" int ref_score_indels_ex(const sx_score_indels_batch* b, const double* lnp, const char* ins_pool, const uint32_t* ins_off, sx_read_indel_score* recs,
                                   uint32_t* n_rec, uint32_t* max_aln, int allow_unordered, char* err, int errlen);

extern "C" int ref_score_indels(const sx_score_indels_batch* b, const double* lnp, const char* ins_pool, const uint32_t* ins_off, sx_read_indel_score* recs,
                                uint32_t* n_rec, uint32_t* max_aln, char* err, int errlen)
{
    return ref_score_indels_ex(b, lnp, ins_pool, ins_off, recs, n_rec, max_aln, 0, err, errlen);
}

extern "C" int ref_score_indels_ex(const sx_score_indels_batch* b, const double* lnp, const char* ins_pool, const uint32_t* ins_off, sx_read_indel_score* recs,
                                   uint32_t* n_rec, uint32_t* max_aln, int allow_unordered, char* err, int errlen)
{
    try
    {
        harness_options2 opt;
        opt.is_candidate_indel_signal_test = false;
        opt.maxIndelSize = b->opts.max_indel_size;
        opt.upstream_oligo_size = b->opts.upstream_oligo_size;
        opt.is_smoothed_alignments = (b->opts.is_smoothed_alignments != 0);
        opt.smoothed_lnp_range = b->opts.smoothed_lnp_range;
        starling_base_deriv_options dopt(opt);
        starling_sample_options sample_opt(opt);
        sample_opt.min_read_bp_flank = b->opts.min_read_bp_flank;
        const auto errorRatesPtr(memberPtr(ErrorRatesTag()));

        for (uint32_t region = 0; region < b->n_regions; ++region)
        {
            const uint32_t k0(b->region_key_off[region]), k1(b->region_key_off[region + 1]);
            int32_t lo(k0 < k1 ? b->keys[k0].pos : 0), hi(lo + 1);
            for (uint32_t k = k0; k < k1; ++k)
            {
                lo = std::min(lo, b->keys[k].pos);
                hi = std::max(hi, b->keys[k].pos + (int32_t)b->keys[k].del_len);
            }
            reference_contig_segment ref;
            {
                std::string seq;
                static const char cyc[4] = {'A', 'C', 'G', 'T'};
                for (int32_t p = lo - 64; p < hi + 64; ++p) seq.push_back(cyc[(p & 0x7fffffff) % 4]);
                ref.seq() = seq;
                ref.set_offset(lo - 64);
            }
            IndelBuffer indelBuffer(opt, dopt, ref);
            depth_buffer db, db2;
            indelBuffer.registerSample(db, db2, false);
            indelBuffer.finalizeSamples();

            std::vector<IndelKey> winKeys;
            std::map<IndelKey, uint32_t> indexOfKey;
            for (uint32_t k = k0; k < k1; ++k)
            {
                const IndelKey ik(key_of(b->keys[k], ins_pool, ins_off, k));
                if (!winKeys.empty() && !(winKeys.back() < ik)) throw blt_exception("ref_score_indels: window is not in IndelKey order");
                winKeys.push_back(ik);
                indexOfKey[ik] = k - k0;
                IndelObservation obs;
                obs.key = ik;
                obs.data.id = 1000000 + k;
                obs.data.iat = INDEL_ALIGN_TYPE::GENOME_TIER1_READ;
                indelBuffer.addIndelObservation(0, obs);
            }
            for (uint32_t k = k0; k < k1; ++k)
            {
                IndelData* idp(indelBuffer.getIndelDataPtr(winKeys[k - k0]));
                if (idp == nullptr) throw blt_exception("ref_score_indels: key not in the IndelBuffer");
                idp->status.is_candidate_indel = (b->keys[k].flags & SX_IKF_CANDIDATE) != 0;
                idp->status.is_candidate_indel_cached = true;
                IndelErrorRates& rates(idp->getSampleData(0).*errorRatesPtr);
                rates.refToIndelErrorProb.updateLogValue(b->keys[k].ref_to_indel_lnp);
                rates.indelToRefErrorProb.updateLogValue(b->keys[k].indel_to_ref_lnp);
            }

            std::vector<std::unique_ptr<bam_record>> bams;
            std::vector<std::unique_ptr<starling_read>> sreads;
            for (uint32_t r = b->region_read_off[region]; r < b->region_read_off[region + 1]; ++r)
            {
                n_rec[r] = 0;
                max_aln[r] = UINT32_MAX;
                const uint32_t a0(b->aln_off[r]), a1(b->aln_off[r + 1]);
                if (a0 == a1) continue;
                if (b->full_len || b->full_off) throw blt_exception("ref_score_indels: only full read segments are rebuilt");
                const bool fwd((b->read_flags[r] & SX_SIF_FWD) != 0);
                const int len(b->read_len[r]);
                std::unique_ptr<bam_record> br(new bam_record);
                br->set_qname("R");
                std::string seq(len, 'A');
                if ((int)b->non_ambig[r] > len) throw blt_exception("ref_score_indels: non_ambig exceeds the read length");
                for (int i = 0; i < len - (int)b->non_ambig[r]; ++i) seq[i] = 'N';
                const std::vector<uint8_t> qual(len, 30);
                br->set_readqual(seq.c_str(), qual.data());
                alignment al;
                al.pos = b->aln_pos[a0];
                al.is_fwd_strand = fwd;
                al.path.push_back(ALIGNPATH::path_segment(ALIGNPATH::MATCH, len));
                br->get_data()->core.pos = al.pos;
                edit_bam_cigar(al.path, *(br->get_data()));
                if (!fwd) br->get_data()->core.flag |= 0x10;
                sreads.emplace_back(new starling_read(*br, al, (b->read_flags[r] & SX_SIF_TIER1) ? MAPLEVEL::TIER1_MAPPED : MAPLEVEL::TIER2_MAPPED, r));
                bams.push_back(std::move(br));
                const read_segment& rseg(sreads.back()->get_full_segment());

                std::set<CandidateAlignment> cals;
                std::vector<double> scores;
                std::vector<uint32_t> flatIndex;
                if (allow_unordered)
                {
                    std::map<CandidateAlignment, uint32_t> first;
                    for (uint32_t a = a0; a < a1; ++a)
                    {
                        CandidateAlignment cal;
                        cal_of(b, a, fwd, winKeys, cal);
                        first.insert(std::make_pair(cal, a));
                    }
                    for (const auto& kv : first)
                    {
                        cals.insert(cals.end(), kv.first);
                        scores.push_back(lnp[kv.second]);
                        flatIndex.push_back(kv.second);
                    }
                }
                else
                    for (uint32_t a = a0; a < a1; ++a)
                    {
                        CandidateAlignment cal;
                        cal_of(b, a, fwd, winKeys, cal);
                        const auto ins(cals.insert(cal));
                        if (!ins.second) throw blt_exception("ref_score_indels: duplicate candidate alignment");
                        // the flat order must already be the set's order: a new element must land at the end
                        if (std::next(ins.first) != cals.end()) throw blt_exception("ref_score_indels: alignments are not in std::set<CandidateAlignment> order");
                        scores.push_back(lnp[a]);
                        flatIndex.push_back(a);
                    }

                // starling_read_align.cpp:1573-1593 with the reference's own tie-break
                double maxScore(0);
                const CandidateAlignment* maxPtr(nullptr);
                uint32_t maxIndex(0), index(0);
                for (const CandidateAlignment& ical : cals)
                {
                    const double path_lnp(scores[index]);
                    const uint32_t thisIndex(index++);
                    if (nullptr != maxPtr)
                    {
                        if (path_lnp < maxScore) continue;
                        if ((path_lnp <= maxScore) && isFirstCandidateAlignmentPreferred(indelBuffer, *maxPtr, ical)) continue;
                    }
                    maxScore = path_lnp;
                    maxPtr = &ical;
                    maxIndex = thisIndex;
                }
                max_aln[r] = flatIndex[maxIndex];

                score_indels(opt, dopt, sample_opt, rseg, indelBuffer, 0, cals, (b->read_flags[r] & SX_SIF_INCOMPLETE) != 0, scores, maxScore, maxPtr);

                // read back this read's entries, in key order
                sx_read_indel_score* out(recs + b->rec_off[r]);
                const uint32_t cap(b->rec_off[r + 1] - b->rec_off[r]);
                uint32_t n(0);
                for (uint32_t k = k0; k < k1; ++k)
                {
                    const IndelSampleData& sd(indelBuffer.getIndelDataPtr(winKeys[k - k0])->getSampleData(0));
                    const auto it(sd.read_path_lnp.find(r));
                    const bool sub(sd.suboverlap_tier1_read_ids.count(r) || sd.suboverlap_tier2_read_ids.count(r));
                    if (it == sd.read_path_lnp.end() && !sub) continue;
                    if (n >= cap) throw blt_exception("ref_score_indels: record capacity");
                    sx_read_indel_score rec;
                    std::memset(&rec, 0, sizeof(rec));
                    rec.key = (uint16_t)(k - k0);
                    if (sub)
                    {
                        const bool t1(sd.suboverlap_tier1_read_ids.count(r) != 0);
                        if (t1 != ((b->read_flags[r] & SX_SIF_TIER1) != 0)) throw blt_exception("ref_score_indels: suboverlap tier differs from the read's tier");
                        rec.flags |= SX_RIS_SUBOVERLAP;
                    }
                    if (it != sd.read_path_lnp.end())
                    {
                        const ReadPathScores& rps(it->second);
                        if (rps.nonAmbiguousBasesInRead != b->non_ambig[r] || rps.read_length != b->read_len[r] || rps.is_fwd_strand != fwd ||
                            rps.is_tier1_read != ((b->read_flags[r] & SX_SIF_TIER1) != 0))
                            throw blt_exception("ref_score_indels: per-read ReadPathScores fields differ from the batch");
                        rec.flags |= SX_RIS_SCORED;
                        rec.ref_lnp = rps.ref;
                        rec.indel_lnp = rps.indel;
                        rec.read_pos = rps.read_pos;
                        rec.dist_from_edge = rps.distanceFromClosestReadEdge;
                        rec.n_alt = (uint8_t)rps.alt_indel.size();
                        for (unsigned i = 0; i < rps.alt_indel.size() && i < 2; ++i)
                        {
                            rec.alt_key[i] = (uint16_t)indexOfKey.at(rps.alt_indel[i].first);
                            rec.alt_lnp[i] = rps.alt_indel[i].second;
                        }
                    }
                    out[n++] = rec;
                }
                n_rec[r] = n;
            }
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        set_err2(err, errlen, e.what());
        return 1;
    }
    catch (...)
    {
        set_err2(err, errlen, "unknown exception");
        return 2;
    }
}

#include "ref_harness_enumerate.inc" // K7: the reference getCandidateAlignments (same translation unit: the function is file-static)
