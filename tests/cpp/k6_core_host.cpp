// tests/cpp/k6_core_host.cpp -- TEST ONLY.  Compiles the device body of K6 (strelka_b200/csrc/k6_core.cuh, __host__ __device__)
// with g++ and runs it read by read on the CPU, so that the container without a GPU can check the kernel's logic against the
// oracle.  Not part of the product: libstrelka_b200.so has no host execution path.
#include "k6_core.cuh"

#include <algorithm>
#include <vector>

extern "C" int k6core_run(const sx_score_indels_batch* b, const double* lnp, sx_read_indel_score* recs, uint32_t* n_rec, uint32_t* max_aln, uint32_t* eval_aln,
                          uint32_t* status_out, uint32_t block_reads, uint32_t smem_cap, uint32_t* staged_blocks)
{
    *staged_blocks = 0;
    uint32_t maxA(1), maxE(1);
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        maxA = std::max(maxA, b->aln_off[r + 1] - b->aln_off[r]);
        maxE = std::max(maxE, std::min<uint32_t>(K6_MAX_EVAL, b->rec_off[r + 1] - b->rec_off[r]));
    }
    std::vector<uint32_t> ord(maxA);
    std::vector<double> smooth(maxA);
    std::vector<float> present(maxE), absent(maxE), alt((size_t)maxE * maxE);
    std::vector<uint16_t> slot(maxE);
    std::vector<uint8_t> filt(maxA), has(maxE), pair((size_t)maxE * maxE);
    std::vector<uint16_t> ev(maxE);
    k6_scratch S;
    S.ord = {ord.data(), 1};
    S.smooth = {smooth.data(), 1};
    S.filt = {filt.data(), 1};
    S.ev = {ev.data(), 1};
    S.slot = {slot.data(), 1};
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
    if (block_reads == 0) // the plain body, read by read
    {
        for (uint32_t region = 0; region < b->n_regions; ++region)
            for (uint32_t r = b->region_read_off[region]; r < b->region_read_off[region + 1]; ++r) status |= k6_score_read(v, region, r, S);
    }
    else // what a thread block does: plan, stage the block's slices ("threads" t = 0..nt-1 in turn), run the body on the rebased view
    {
        std::vector<uint64_t> smem;
        for (uint32_t r0 = 0; r0 < b->n_reads; r0 += block_reads)
        {
            const uint32_t r1(std::min(b->n_reads, r0 + block_reads));
            const k6_block_plan p(k6_plan_block(v.b, r0, r1));
            const bool staged(p.bytes <= smem_cap);
            k6_view lv(v);
            if (staged)
            {
                smem.assign(p.bytes / 8 + 2, 0xdeadbeefdeadbeefull);
                unsigned char* sm(reinterpret_cast<unsigned char*>(smem.data()));
                for (uint32_t t = 0; t < block_reads; ++t) k6_stage(v, p, sm, t, block_reads);
                lv = k6_rebased(v, p, sm);
                ++*staged_blocks;
            }
            for (uint32_t r = r0; r < r1; ++r) status |= k6_score_read_in_block(lv, p, r, S);
        }
    }
    *status_out = status;
    return 0;
}
