// tests/cpp/k9_core_host.cpp -- TEST ONLY.  Compiles the device body of K9 (strelka_b200/csrc/k9_core.cuh, __host__ __device__) with g++
// and runs it the way the kernels of k9_realign.cu do (slots per read, scan, choose), with a poisoned per-thread map.  Not part of the product.
#include "k9_core.cuh"

#include <vector>

extern "C" int k9core_run(const sx_realign_batch* b, const double* lnp, sx_realign_out* o)
{
    k9_view v;
    v.b = *b;
    v.lnp = lnp;
    std::vector<uint32_t> read_region(b->n_reads);
    for (uint32_t g = 0; g < b->n_regions; ++g)
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r) read_region[r] = g;
    uint32_t total(0);
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        o->seg_off[r] = total;
        total += k9_slots(*b, r);
    }
    o->seg_off[b->n_reads] = total;
    o->totals[0] = total;
    if (total > o->cap_segs) return SX_ERR_CAPACITY;
    std::vector<uint8_t> type(K9_MAX_READ);
    std::vector<int32_t> pos(K9_MAX_READ);
    k9_scratch S = {type.data(), pos.data()};
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        for (uint32_t i = 0; i < K9_MAX_READ; ++i) // never read before written
        {
            type[i] = 0x7B;
            pos[i] = -12345678;
        }
        const uint32_t s0(o->seg_off[r]), s1(o->seg_off[r + 1]);
        int32_t p;
        uint16_t ns;
        uint32_t best;
        const uint32_t st(k9_read(v, read_region[r], r, S, o->segs + s0, s1 - s0, p, ns, best));
        if (!(st & SX_REALIGN_ST_REALIGNED)) k9_fallback(*b, r, o->segs + s0, s1 - s0, p, ns);
        o->pos[r] = p;
        o->n_seg[r] = ns;
        o->status[r] = (uint8_t)st;
        o->best_aln[r] = best;
    }
    return 0;
}
