// tests/cpp/k7_core_host.cpp -- TEST ONLY.  Compiles the device body of K7 (strelka_b200/csrc/k7_core.cuh, __host__ __device__)
// with g++ and runs it read by read on the CPU the way the kernels of k7_enumerate.cu do (count pass, exclusive scan, write pass),
// so that the container without a GPU can check the kernel's logic against the reference's getCandidateAlignments.  Not part of
// the product: libstrelka_b200.so has no host execution path.
#include "k7_core.cuh"

#include <algorithm>
#include <vector>


// what k7_frames_needed (k7_enumerate.cu) computes on the device: a search holds at most as many indels as its region's window has entries
static uint32_t frames_needed(const sx_enum_batch* b)
{
    uint32_t m(0);
    for (uint32_t g = 0; g < b->n_regions; ++g) m = std::max(m, b->region_key_off[g + 1] - b->region_key_off[g]);
    return std::min<uint32_t>(K7_MAX_INDELS, m) + 1u;
}

extern "C" int k7core_run(const sx_enum_batch* b, sx_enum_out* o, uint32_t maxA)
{
    if (maxA == 0) maxA = 64;
    // the device arena is never cleared: poison it, so that any read-before-write of the scratch shows up here as a mismatch
    const uint32_t maxF(frames_needed(b));
    std::vector<unsigned char> arena(k7_scratch_bytes(maxA, maxF) + 64, 0xCD);
    k7_scratch S(k7_scratch_at(arena.data(), maxA, maxF));
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
    for (uint32_t i = 0; i < cal.n_seg; ++i) segs[i] = sx_aln_seg{(uint16_t)k7_sl(cal.seg[i]), (uint8_t)k7_sk(cal.seg[i]), (uint8_t)(cal.seg[i] >> 24)};
    return 0;
}

extern "C" int k7core_end_pin_start_pos(const sx_indel_key* win, uint32_t n_win, uint32_t read_length, int32_t ref_end, int32_t read_end, int32_t* ref_start,
                                        int32_t* read_start)
{
    uint16_t indels[K7_MAX_INDELS];
    for (uint32_t i = 0; i < n_win; ++i) indels[i] = (uint16_t)i;
    return (int)k7_end_pin_start_pos(win, indels, n_win, read_length, ref_end, read_end, *ref_start, *read_start);
}

// The SX_ENUM_F_FAST plan on the CPU: tier 1 (small scratch: 16 alignments, 12 frames; K7_ST_RETRY when it is full), tier 2 (arena) for
// the marked reads, every read's alignments appended to a log -- here in REVERSE read order, to show that the result does not depend on
// who got which piece of the log --, scan, gather.
extern "C" int k7core_run_fast(const sx_enum_batch* b, sx_enum_out* o, uint32_t maxA, uint32_t* n_retried)
{
    if (maxA == 0) maxA = 64;
    const uint32_t LA(16), LF(12);
    const uint32_t maxF(frames_needed(b));
    std::vector<unsigned char> small(k7_scratch_bytes(LA, LF) + 64, 0xCD), arena(k7_scratch_bytes(maxA, maxF) + 64, 0xCD);
    k7_view v;
    v.b = *b;
    const uint32_t n(b->n_reads);
    std::vector<uint32_t> read_region(n), ca(n + 1), cs(n + 1), ck(n + 1), blob_off(n, UINT32_MAX);
    std::vector<uint8_t> tier(n, 0);
    for (uint32_t g = 0; g < b->n_regions; ++g)
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r) read_region[r] = g;
    std::vector<uint32_t> log;
    *n_retried = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (uint32_t rr = 0; rr < n; ++rr)
        {
            const uint32_t r(n - 1 - rr);
            if (pass == 1 && !tier[r]) continue;
            std::fill(small.begin(), small.end(), (unsigned char)(0x3C + (r & 0x1f)));
            k7_scratch S(pass == 0 ? k7_scratch_at(small.data(), LA, LF, K7_ST_RETRY) : k7_scratch_at(arena.data(), maxA, maxF));
            const uint32_t st(k7_enumerate_read(v, read_region[r], r, S));
            if (pass == 0 && (st & K7_ST_RETRY))
            {
                tier[r] = 1;
                ++*n_retried;
                ca[r] = cs[r] = ck[r] = 0;
                continue;
            }
            uint32_t na, ns, nk;
            k7_count(S, st, na, ns, nk);
            o->status[r] = (uint8_t)st;
            ca[r] = na;
            cs[r] = ns;
            ck[r] = nk;
            if (na)
            {
                blob_off[r] = (uint32_t)log.size();
                log.resize(log.size() + k7_blob_words(S), 0xDEADBEEFu);
                k7_blob_write(S, log.data() + blob_off[r]);
            }
        }
    uint32_t ta(0), ts(0), tk(0);
    for (uint32_t r = 0; r < n; ++r) // the scan
    {
        const uint32_t na(ca[r]), ns(cs[r]), nk(ck[r]);
        ca[r] = ta;
        cs[r] = ts;
        ck[r] = tk;
        o->aln_off[r] = ta;
        ta += na;
        ts += ns;
        tk += nk;
    }
    ca[n] = ta;
    o->aln_off[n] = ta;
    o->totals[0] = ta;
    o->totals[1] = ts;
    o->totals[2] = tk;
    if (ta > o->cap_alns || ts > o->cap_segs || tk > o->cap_keys) return SX_ERR_CAPACITY;
    o->aln_seg_off[ta] = ts;
    o->aln_key_off[ta] = tk;
    for (uint32_t r = 0; r < n; ++r) // k7_gather_kernel
    {
        const uint32_t na(ca[r + 1] - ca[r]);
        if (na == 0 || blob_off[r] == UINT32_MAX) continue;
        k7_blob_gather(log.data() + blob_off[r], na, *o, ca[r], cs[r], ck[r]);
    }
    return 0;
}
