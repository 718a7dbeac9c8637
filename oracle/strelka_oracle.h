/*
 * strelka_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the Strelka2 hot path, operating on the same flattened
 * batches as the product's C ABI (include/strelka_b200.h).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or the reported CPU baseline.  The product (strelka_b200/) never links or calls it.
 *
 * Parity status: PINNED.  Every function here is checked in tests/ against the reference's own
 * code compiled from /root/reference (oracle/_ref/libstrelka_ref.so, built by oracle/build_ref.sh)
 * and against the reference's unit-test golden vectors (alignment/test/GlobalAlignerTest.cpp).
 */
#ifndef STRELKA_ORACLE_H
#define STRELKA_ORACLE_H

#include "strelka_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* a1: starling_common/starling_read_align_score.cpp:260-499 */
int ox_score_alignments(const sx_align_batch* b, double* lnp_out);
/* same, regions [r0, r1) only; lnp_out is still indexed by global alignment index */
int ox_score_alignments_range(const sx_align_batch* b, uint32_t r0, uint32_t r1, double* lnp_out);

/* a5: alignment/GlobalAlignerImpl.hh:36-228 + SingleRefAlignerSharedImpl.hh:80-170 */
int ox_global_align(const sx_ga_scores* s, const sx_ga_batch* b, sx_ga_result* res, uint32_t* cigar);

/* a6-a9: PileupCleaner.cpp:30-75, adjust_joint_eprob.cpp:60-243, position_snp_call_pprob_digt.cpp:326-539 */
int ox_site_gl_germline(const sx_params* p, const sx_pileup_batch* b, int is_always_test, sx_digt_result* out);
int ox_site_gl_germline_range(const sx_params* p, const sx_pileup_batch* b, int is_always_test, uint32_t s0, uint32_t s1, sx_digt_result* out);
int ox_dependent_eprob(const sx_params* p, const sx_pileup_batch* b, uint32_t* out_off, float* de);

/* a10-a11: position_somatic_snv_strand_grid*.cpp, qscore_calculator.cpp */
int ox_site_gl_somatic(const sx_params* p, const sx_pileup_batch* normal, const sx_pileup_batch* tumor,
                       const uint8_t* is_forced_output, sx_ssnv_result* out);
int ox_site_gl_somatic_range(const sx_params* p, const sx_pileup_batch* normal, const sx_pileup_batch* tumor,
                             const uint8_t* is_forced_output, uint32_t s0, uint32_t s1, sx_ssnv_result* out);

/* f2: the arg-max epilogue of scoreCandidateAlignments (starling_read_align.cpp:1573-1593) + score_indels
 * (starling_read_align_score_indels.cpp:454-1079); oracle/score_indels_oracle.cpp */
void ox_default_score_indels_opts(sx_score_indels_opts* o);
int ox_score_indels(const sx_score_indels_batch* b, const double* lnp, sx_read_indel_score* recs, uint32_t* n_rec, uint32_t* max_aln, uint32_t* eval_aln);

/* restatements of the host-libm single-precision routines the device mirrors (glibc 2.39 x86_64 FMA ifunc variants),
 * exported so tests can compare them with the real logf/powf exhaustively */
float ox_logf_restated(float x);
float ox_powf_restated(float x, float y);
/* restatement of libstdc++'s std::sort (bits/stl_algo.h) on (key desc) index arrays, for comparison with the real one */
void ox_sort_restated(uint32_t* idx, uint32_t n, const uint8_t* key_by_idx);
void ox_sort_std(uint32_t* idx, uint32_t n, const uint8_t* key_by_idx);

#ifdef __cplusplus
}
#endif
#endif
