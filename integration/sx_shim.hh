// integration/sx_shim.hh -- the C++ shim a reference maintainer adds at the reference's own call sites to route the hot path through
// libstrelka_b200.so (include/strelka_b200.h).  Included ONLY by the patched copies of the reference's translation units that
// integration/build_patched.py makes (the reference tree itself is never modified, nor copied into this repository).
//
// One sx_ctx per process: the reference runs one single-threaded process per genome segment.  ABI errors become blt_exception, the
// reference's own convention (e.g. starling_pos_processor_base.cpp:755-760).  Setting SX_SHIM_OFF=1 in the environment keeps the
// reference's CPU functions (A/B runs of the same binary).
#pragma once

#include "strelka_b200.h"
#include "strelka_b200.hh"

#include "alignment/GlobalAligner.hh"
#include "blt_common/blt_shared.hh"
#include "blt_common/position_snp_call_pprob_digt.hh"
#include "blt_common/snp_pos_info.hh"
#include "blt_util/blt_exception.hh"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace sx_shim
{
static_assert(sizeof(base_call) == 2, "base_call must be the 16-bit word the ABI carries (blt_common/snp_pos_info.hh:109-118)");

inline bool enabled()
{
    static const bool on(std::getenv("SX_SHIM_OFF") == nullptr);
    return on;
}

struct Counters
{
    unsigned long long site_gl_germline = 0, site_gl_somatic = 0, global_align = 0;
    ~Counters()
    {
        if (std::getenv("SX_SHIM_REPORT"))
            std::fprintf(stderr, "sx_shim: %llu germline site calls, %llu somatic site calls, %llu haplotype alignments through libstrelka_b200.so\n", site_gl_germline,
                         site_gl_somatic, global_align);
    }
};
inline Counters& counters()
{
    static Counters c;
    return c;
}

inline sx_ctx* context(const sx_params& p)
{
    struct Holder
    {
        sx_ctx* ctx = nullptr;
        explicit Holder(const sx_params& q)
        {
            const char* dev(std::getenv("SX_SHIM_DEVICE"));
            if (sx_create(dev ? std::atoi(dev) : 0, &q, &ctx) != SX_OK) throw blt_exception(sx_last_error(nullptr));
        }
        ~Holder() { sx_destroy(ctx); }
    };
    static Holder h(p);
    return h.ctx;
}

inline void check(sx_ctx* ctx, const int rc)
{
    if (rc != SX_OK) throw blt_exception(sx_last_error(ctx));
}

inline sx_params params_of(const blt_options& opt)
{
    sx_params p;
    sx_default_params(&p);
    p.bsnp_diploid_theta = opt.bsnp_diploid_theta;
    p.bsnp_ssd_no_mismatch = opt.bsnp_ssd_no_mismatch;
    p.bsnp_ssd_one_mismatch = opt.bsnp_ssd_one_mismatch;
    p.is_min_vexp = opt.is_min_vexp;
    p.min_vexp = opt.min_vexp;
    p.is_bsnp_diploid = opt.is_bsnp_diploid();
    p.hetVariantFrequencyExtension = opt.hetVariantFrequencyExtension;
    return p;
}

inline const uint16_t* words(const std::vector<base_call>& v) { return reinterpret_cast<const uint16_t*>(v.data()); }

/// computeSampleDiploidSiteGenotype (applications/starling/starling_pos_processor.cpp:254-267): CleanPileupFilter + CleanPileupErrorProb +
/// position_snp_call_pprob_digt of one sample's position, from the RAW pile-up column
inline void site_gl_germline(const blt_options& opt, const snp_pos_info& raw, const bool is_always_test, diploid_genotype& dgt)
{
    sx_ctx* ctx(context(params_of(opt)));
    const uint32_t off[2] = {0u, static_cast<uint32_t>(raw.calls.size())};
    const uint16_t none(0);
    const char ref_base(raw.get_ref_base());
    const uint8_t ploidy(static_cast<uint8_t>(dgt.ploidy));
    sx_pileup_batch b;
    std::memset(&b, 0, sizeof(b));
    b.n_sites = 1;
    b.site_off = off;
    b.calls = raw.calls.empty() ? &none : words(raw.calls);
    b.ref_base = &ref_base;
    b.ploidy = &ploidy;
    sx_digt_result r;
    check(ctx, sx_site_gl_germline(ctx, &b, is_always_test ? 1 : 0, &r));
    ++counters().site_gl_germline;
    const int keep_ploidy(dgt.ploidy);
    dgt.reset();
    dgt.ploidy = keep_ploidy;
    dgt.ref_gt = r.ref_gt;
    if (!r.is_computed) return; // the reference returned before touching the rest (position_snp_call_pprob_digt.cpp:484-492)
    dgt.strand_bias = r.strand_bias;
    const sx_digt_result_set* in[2] = {&r.genome, &r.poly};
    diploid_genotype::result_set* out[2] = {&dgt.genome, &dgt.poly};
    for (int k = 0; k < 2; ++k)
    {
        out[k]->max_gt = in[k]->max_gt;
        out[k]->ref_pprob = in[k]->ref_pprob;
        out[k]->snp_qphred = in[k]->snp_qphred;
        out[k]->max_gt_qphred = in[k]->max_gt_qphred;
    }
    for (unsigned gt = 0; gt < 10; ++gt) dgt.phredLoghood[gt] = r.phredLoghood[gt];
}

/// _aligner.align(hap.begin, hap.end, ref.begin, ref.end, result) (starling_common/ActiveRegionProcessor.cpp:591)
inline void global_align(const AlignmentScores<int>& sc, const std::string& query, const std::string& ref, AlignmentResult<int>& result)
{
    sx_params p;
    sx_default_params(&p);
    sx_ctx* ctx(context(p));
    sx_ga_scores s = {sc.match, sc.mismatch, sc.open, sc.extend, sc.offEdge, sc.insertDelete, sc.isAllowEdgeInsertion ? 1 : 0, sc.isRequireEdgeDeletion ? 1 : 0};
    std::string q(query), r(ref);
    const uint32_t qo[2] = {0u, static_cast<uint32_t>(q.size())}, ro[2] = {0u, static_cast<uint32_t>(r.size())};
    const uint32_t maxOps(static_cast<uint32_t>(q.size() + r.size() + 2));
    q.resize(q.size() + 16);
    r.resize(r.size() + 16);
    sx_ga_batch b = {1u, q.data(), r.data(), qo, ro, maxOps};
    sx_ga_result res;
    std::vector<uint32_t> cig(maxOps);
    check(ctx, sx_global_align(ctx, &s, &b, &res, cig.data()));
    ++counters().global_align;
    if (res.status != 0) throw blt_exception("sx_global_align: problem too large for the kernel");
    result.score = res.score;
    result.align.beginPos = res.beginPos;
    result.align.apath.clear();
    static const ALIGNPATH::align_t kind[9] = {ALIGNPATH::MATCH, ALIGNPATH::INSERT, ALIGNPATH::DELETE, ALIGNPATH::SKIP, ALIGNPATH::SOFT_CLIP, ALIGNPATH::HARD_CLIP,
                                               ALIGNPATH::PAD, ALIGNPATH::SEQ_MATCH, ALIGNPATH::SEQ_MISMATCH};
    for (uint32_t k = 0; k < res.n_ops; ++k) result.align.apath.push_back(ALIGNPATH::path_segment(kind[cig[k] & 15u], cig[k] >> 4));
}

#ifdef SX_SHIM_SOMATIC
/// sscaller_strand_grid().position_somatic_snv_call(nepi, tepi, nepi_t2, tepi_t2, isComputeNonSomatic = false, sgtg)
/// (applications/strelka/strelka_pos_processor.cpp:213-219), from the RAW tumor / normal columns (tier1 + tier2 calls)
inline void site_gl_somatic(const strelka_options& opt, const snp_pos_info& normal, const snp_pos_info& tumor, somatic_snv_genotype_grid& sgtg)
{
    sx_params p(params_of(opt));
    p.somatic_snv_rate = opt.somatic_snv_rate;
    p.shared_site_error_rate = opt.shared_site_error_rate;
    p.shared_site_error_strand_bias_fraction = opt.shared_site_error_strand_bias_fraction;
    p.ssnv_contam_tolerance = opt.ssnv_contam_tolerance;
    sx_ctx* ctx(context(p));
    const uint16_t none(0);
    const char ref_base(normal.get_ref_base());
    const snp_pos_info* pi[2] = {&normal, &tumor};
    uint32_t off[2][2], t2off[2][2];
    sx_pileup_batch b[2];
    for (int s = 0; s < 2; ++s)
    {
        off[s][0] = t2off[s][0] = 0;
        off[s][1] = static_cast<uint32_t>(pi[s]->calls.size());
        t2off[s][1] = static_cast<uint32_t>(pi[s]->tier2_calls.size());
        std::memset(&b[s], 0, sizeof(b[s]));
        b[s].n_sites = 1;
        b[s].site_off = off[s];
        b[s].calls = pi[s]->calls.empty() ? &none : words(pi[s]->calls);
        if (opt.useTier2Evidence)
        {
            b[s].t2_off = t2off[s];
            b[s].t2_calls = pi[s]->tier2_calls.empty() ? &none : words(pi[s]->tier2_calls);
        }
        b[s].ref_base = &ref_base;
    }
    const uint8_t forced(sgtg.is_forced_output ? 1 : 0);
    sx_ssnv_result r;
    check(ctx, sx_site_gl_somatic(ctx, &b[0], &b[1], &forced, &r));
    ++counters().site_gl_somatic;
    sgtg.ref_gt = r.ref_gt;
    if (!r.is_computed) return;
    sgtg.snv_tier = r.snv_tier != 0;
    sgtg.snv_from_ntype_tier = r.snv_from_ntype_tier != 0;
    sgtg.rs.ntype = r.ntype;
    sgtg.rs.max_gt = r.max_gt;
    sgtg.rs.qphred = r.qphred;
    sgtg.rs.from_ntype_qphred = r.from_ntype_qphred;
    sgtg.rs.normal_alt_id = r.normal_alt_id;
    sgtg.rs.tumor_alt_id = r.tumor_alt_id;
    sgtg.rs.strandBias = r.strandBias;
}
#endif
} // namespace sx_shim
