// tests/cpp/test_k7_mirror_cpu.cpp -- TEST ONLY, runs without a GPU.  The K7 part of the C++ host mirror
// (sx::AlignmentSearchBatch: flattening reference-shaped objects into an sx_enum_batch, restating getAlignmentIndels, decoding the
// CSR result back into CandidateAlignments) checked against the reference's getCandidateAlignments (tests/golden/k7_cases.tsv) with
// the four library entry points it needs answered here: sx_create / sx_destroy / sx_last_error are stubs and
// sx_enumerate_alignments runs the device body of K7 (strelka_b200/csrc/k7_core.cuh, __host__ __device__) read by read the way the
// kernels do.  On the GPU box tests/cpp/test_k7_mirror.cpp runs the same check through the real library.
#include "k7_core.cuh"
#include "k9_core.cuh"

#include "k7_mirror_check.hh"

#include <vector>

extern "C" int sx_create(int, const sx_params*, sx_ctx** out)
{
    static int dummy;
    *out = reinterpret_cast<sx_ctx*>(&dummy);
    return SX_OK;
}
extern "C" void sx_destroy(sx_ctx*) {}
extern "C" const char* sx_last_error(const sx_ctx*) { return "host stand-in"; }

extern "C" int sx_enumerate_alignments(sx_ctx*, const sx_enum_batch* b, sx_enum_out* o)
{
    const uint32_t maxA(b->opts.max_alns_per_read ? b->opts.max_alns_per_read : 64u);
    std::vector<unsigned char> arena(k7_scratch_bytes(maxA) + 64);
    k7_scratch S(k7_scratch_at(arena.data(), maxA));
    k7_view v;
    v.b = *b;
    std::vector<uint32_t> region(b->n_reads), ca(b->n_reads), cs(b->n_reads), ck(b->n_reads);
    for (uint32_t g = 0; g < b->n_regions; ++g)
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r) region[r] = g;
    uint32_t ta(0), ts(0), tk(0);
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        const uint32_t st(k7_enumerate_read(v, region[r], r, S));
        uint32_t na, ns, nk;
        k7_count(S, st, na, ns, nk);
        o->status[r] = (uint8_t)st;
        ca[r] = ta;
        cs[r] = ts;
        ck[r] = tk;
        o->aln_off[r] = ta;
        ta += na;
        ts += ns;
        tk += nk;
    }
    o->aln_off[b->n_reads] = ta;
    o->totals[0] = ta;
    o->totals[1] = ts;
    o->totals[2] = tk;
    if (ta > o->cap_alns || ts > o->cap_segs || tk > o->cap_keys) return SX_ERR_CAPACITY;
    o->aln_seg_off[ta] = ts;
    o->aln_key_off[ta] = tk;
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        if (o->aln_off[r + 1] == o->aln_off[r]) continue;
        k7_enumerate_read(v, region[r], r, S);
        k7_write(S, *o, ca[r], cs[r], ck[r]);
    }
    return SX_OK;
}

extern "C" int sx_choose_realignment(sx_ctx*, const sx_realign_batch* b, const double* lnp, sx_realign_out* o)
{
    k9_view v;
    v.b = *b;
    v.lnp = lnp;
    std::vector<uint32_t> region(b->n_reads);
    for (uint32_t g = 0; g < b->n_regions; ++g)
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r) region[r] = g;
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
        int32_t p;
        uint16_t ns;
        uint32_t best;
        o->status[r] = (uint8_t)k9_read(v, region[r], r, S, o->segs + o->seg_off[r], o->seg_off[r + 1] - o->seg_off[r], p, ns, best);
        o->pos[r] = p;
        o->n_seg[r] = ns;
        o->best_aln[r] = best;
    }
    return SX_OK;
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    int checks(0), failures(0);
    try
    {
        sx::Context ctx(0);
        k7_mirror_check(ctx, argv[1], checks, failures);
    }
    catch (const std::exception& e)
    {
        std::cerr << "EXCEPTION: " << e.what() << "\n";
        return 3;
    }
    std::cout << "k7 host mirror (CPU stand-in): " << checks << " checks, " << failures << " failures\n";
    return failures ? 1 : 0;
}
