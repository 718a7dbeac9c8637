// tests/cpp/k7a_core_host.cpp -- TEST ONLY.  Compiles the device body of K7a (strelka_b200/csrc/k7a_core.cuh, __host__ __device__) with
// g++ and runs it the way the kernels of k7a_prepare.cu do (read offsets, count, scan, write), so that the GPU-less container can check
// it against the reference's getAlignmentIndels.  Not part of the product.
#include "k7a_core.cuh"

#include <vector>

extern "C" int k7acore_run(const sx_enum_batch* b, const sx_region* regions, const uint8_t* seq4, const char* ref, const uint32_t* key_ins_off, const char* key_ins,
                           sx_prep_out* o)
{
    k7a_view v;
    v.b = *b;
    v.regions = regions;
    v.seq4 = seq4;
    v.ref = ref;
    v.key_ins_off = key_ins_off;
    v.key_ins = key_ins;
    std::vector<uint64_t> read_byte(b->n_reads);
    std::vector<uint32_t> read_region(b->n_reads), cnt(b->n_reads + 1);
    for (uint32_t g = 0; g < b->n_regions; ++g)
    {
        uint64_t at(regions[g].seq_off);
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r)
        {
            read_byte[r] = at;
            read_region[r] = g;
            at += (b->read_len[r] + 1u) / 2u;
        }
    }
    uint32_t total(0);
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        uint16_t keys[K7A_MAX_KEYS], lead, trail;
        const uint32_t n(k7a_read(v, read_region[r], r, read_byte[r], keys, lead, trail));
        o->in_lead_key[r] = lead;
        o->in_trail_key[r] = trail;
        o->in_key_off[r] = total;
        cnt[r] = total;
        total += n;
    }
    o->in_key_off[b->n_reads] = total;
    o->totals[0] = total;
    if (total > o->cap_keys) return SX_ERR_CAPACITY;
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        uint16_t keys[K7A_MAX_KEYS], lead, trail;
        for (uint32_t i = 0; i < K7A_MAX_KEYS; ++i) keys[i] = 0xABCD; // never read before written
        const uint32_t n(k7a_read(v, read_region[r], r, read_byte[r], keys, lead, trail));
        for (uint32_t i = 0; i < n; ++i) o->in_keys[cnt[r] + i] = keys[i];
    }
    return 0;
}

// K7g realign_gates: the per-read body run read by read (one kernel, no scan)
extern "C" int k7gcore_run(const sx_gate_batch* b, sx_gate_out* o)
{
    for (uint32_t g = 0; g < b->n_regions; ++g)
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r)
        {
            int32_t pos;
            o->gate[r] = (uint8_t)k7g_read(*b, g, r, pos, o->in_segs + b->seg_off[r]);
            o->in_pos[r] = pos;
        }
    return 0;
}
