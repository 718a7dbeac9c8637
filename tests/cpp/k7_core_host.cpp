// tests/cpp/k7_core_host.cpp -- TEST ONLY.  Compiles the device body of K7 (strelka_b200/csrc/k7_core.cuh, __host__ __device__)
// with g++ and runs it read by read on the CPU the way the kernels of k7_enumerate.cu do (count pass, exclusive scan, write pass),
// so that the container without a GPU can check the kernel's logic against the reference's getCandidateAlignments.  Not part of
// the product: libstrelka_b200.so has no host execution path.
#include "k7_core.cuh"

#include <algorithm>
#include <vector>

extern "C" int k7core_run(const sx_enum_batch* b, sx_enum_out* o, uint32_t maxA)
{
    if (maxA == 0) maxA = 64;
    // the device arena is never cleared: poison it, so that any read-before-write of the scratch shows up here as a mismatch
    std::vector<unsigned char> arena(k7_scratch_bytes(maxA) + 64, 0xCD);
    k7_scratch S(k7_scratch_at(arena.data(), maxA));
    k7_view v;
    v.b = *b;
    std::vector<uint32_t> read_region(b->n_reads), ca(b->n_reads + 1), cs(b->n_reads + 1), ck(b->n_reads + 1);
    for (uint32_t g = 0; g < b->n_regions; ++g)
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r) read_region[r] = g;
    uint32_t ta(0), ts(0), tk(0);
    for (uint32_t r = 0; r < b->n_reads; ++r) // k7_count_kernel + the scan
    {
        const uint32_t st(k7_enumerate_read(v, read_region[r], r, S));
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
    for (uint32_t r = 0; r < b->n_reads; ++r) // k7_write_kernel
    {
        if (o->status[r] & (SX_ENUM_ST_EXCEPTION | SX_ENUM_ST_LIMIT)) continue;
        if (o->aln_off[r + 1] == o->aln_off[r]) continue;
        std::fill(arena.begin(), arena.end(), (unsigned char)(0x5A + (r & 0x3f))); // a different thread's leftovers
        k7_enumerate_read(v, read_region[r], r, S);
        k7_write(S, *o, ca[r], cs[r], ck[r]);
    }
    return 0;
}

// hooks for the known-answer vectors of the reference's own unit test (tests/golden/read_align_unit_goldens.json): the whole window
// is the indel set
extern "C" int k7core_make_start_pos(const sx_indel_key* win, uint32_t n_win, int32_t ref_start, int32_t read_start, uint32_t read_length, int32_t* pos, uint16_t* lead,
                                     uint16_t* trail, sx_aln_seg* segs, uint32_t* n_seg)
{
    uint16_t indels[K7_MAX_INDELS];
    for (uint32_t i = 0; i < n_win; ++i) indels[i] = (uint16_t)i;
    k7_path cal;
    const uint32_t st(k7_make_start_pos(win, ref_start, read_start, read_length, indels, n_win, cal));
    if (st) return (int)st;
    *pos = cal.pos;
    *lead = cal.lead;
    *trail = cal.trail;
    *n_seg = cal.n_seg;
    for (uint32_t i = 0; i < cal.n_seg; ++i) segs[i] = cal.seg[i];
    return 0;
}

extern "C" int k7core_end_pin_start_pos(const sx_indel_key* win, uint32_t n_win, uint32_t read_length, int32_t ref_end, int32_t read_end, int32_t* ref_start,
                                        int32_t* read_start)
{
    uint16_t indels[K7_MAX_INDELS];
    for (uint32_t i = 0; i < n_win; ++i) indels[i] = (uint16_t)i;
    return (int)k7_end_pin_start_pos(win, indels, n_win, read_length, ref_end, read_end, *ref_start, *read_start);
}
