// tests/cpp/k6_core_host.cpp -- TEST ONLY.  Compiles the device body of K6 (strelka_b200/csrc/k6_core.cuh, __host__ __device__)
// with g++ and runs it read by read on the CPU, so that the container without a GPU can check the kernel's logic against the
// oracle.  Not part of the product: libstrelka_b200.so has no host execution path.
#include "k6_core.cuh"

#include <algorithm>
#include <vector>

extern "C" int k6core_run(const sx_score_indels_batch* b, const double* lnp, sx_read_indel_score* recs, uint32_t* n_rec, uint32_t* max_aln, uint32_t* eval_aln,
                          uint32_t* status_out)
{
    uint32_t maxA(1), maxE(1);
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        maxA = std::max(maxA, b->aln_off[r + 1] - b->aln_off[r]);
        maxE = std::max(maxE, std::min<uint32_t>(K6_MAX_EVAL, b->rec_off[r + 1] - b->rec_off[r]));
    }
    std::vector<uint32_t> ord(maxA);
    std::vector<double> smooth(maxA), present(maxE), absent(maxE), alt((size_t)maxE * maxE);
    std::vector<uint8_t> filt(maxA), has(maxE), pair((size_t)maxE * maxE);
    std::vector<uint16_t> ev(maxE);
    k6_scratch S;
    S.ord = {ord.data(), 1};
    S.smooth = {smooth.data(), 1};
    S.filt = {filt.data(), 1};
    S.ev = {ev.data(), 1};
    S.present = {present.data(), 1};
    S.absent = {absent.data(), 1};
    S.has = {has.data(), 1};
    S.alt = {alt.data(), 1};
    S.pair = {pair.data(), 1};
    S.maxA = maxA;
    S.maxE = maxE;
    k6_view v;
    v.b = *b;
    v.lnp = lnp;
    v.recs = recs;
    v.n_rec = n_rec;
    v.max_aln = max_aln;
    v.eval_aln = eval_aln;
    uint32_t status(0);
    for (uint32_t region = 0; region < b->n_regions; ++region)
        for (uint32_t r = b->region_read_off[region]; r < b->region_read_off[region + 1]; ++r) status |= k6_score_read(v, region, r, S);
    *status_out = status;
    return 0;
}
