// tests/cpp/k8_core_host.cpp -- TEST ONLY.  Compiles the device body of K7b (strelka_b200/csrc/k8_core.cuh, __host__ __device__) with
// g++ and runs it the way the kernels of k8_link.cu do (sizes, per-region offsets with K1's padding rule, region scan, write, pads),
// so that the GPU-less container can check the link between K7 and K1.  Not part of the product.
#include "k8_core.cuh"

#include <vector>

extern "C" int k8core_run(const sx_enum_batch* b, const sx_enum_out* e, uint32_t n_alns, const uint32_t* key_ins_off, const char* key_ins, sx_link_out* o)
{
    k8_view v;
    v.b = *b;
    v.e = *e;
    v.key_ins_off = key_ins_off;
    v.key_ins = key_ins;
    std::vector<uint32_t> aln_read(n_alns), read_region(b->n_reads), seg_n(n_alns), ins_n(n_alns), reg_seg(b->n_regions + 1), reg_ins(b->n_regions + 1);
    for (uint32_t g = 0; g < b->n_regions; ++g)
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r) read_region[r] = g;
    for (uint32_t r = 0; r < b->n_reads; ++r)
        for (uint32_t a = e->aln_off[r]; a < e->aln_off[r + 1]; ++a) aln_read[a] = r;
    uint32_t st(0);
    for (uint32_t a = 0; a < n_alns; ++a) st |= k8_walk(v, read_region[aln_read[a]], a, seg_n[a], ins_n[a], nullptr, nullptr);
    if (st) return -(int)st;
    uint32_t ts(0), ti(0);
    for (uint32_t g = 0; g < b->n_regions; ++g) // k8_region_kernel + scan + k8_finish_kernel
    {
        const uint32_t a0(e->aln_off[b->region_read_off[g]]), a1(e->aln_off[b->region_read_off[g + 1]]);
        uint32_t s(0), n(0);
        for (uint32_t a = a0; a < a1; ++a)
        {
            const uint32_t ds(seg_n[a]), dn(ins_n[a]);
            seg_n[a] = s;
            ins_n[a] = n;
            s += ds;
            n += dn;
        }
        reg_seg[g] = ts;
        reg_ins[g] = ti;
        o->regions[g].aln_begin = a0;
        o->regions[g].seg_begin = ts;
        o->regions[g].ins_begin = ti;
        ts += (s + 7u) & ~7u;
        ti += (n + 15u) & ~15u;
    }
    reg_seg[b->n_regions] = ts;
    reg_ins[b->n_regions] = ti;
    o->totals[0] = ts;
    o->totals[1] = ti;
    o->regions[b->n_regions].aln_begin = n_alns;
    o->regions[b->n_regions].seg_begin = ts;
    o->regions[b->n_regions].ins_begin = ti;
    o->regions[b->n_regions].read_begin = b->n_reads;
    if (ts > o->cap_segs || ti > o->cap_ins) return SX_ERR_CAPACITY;
    o->alns[n_alns] = sx_aln{b->n_reads, 0, ts, ti};
    for (uint32_t a = 0; a < n_alns; ++a) // k8_write_kernel
    {
        const uint32_t r(aln_read[a]), g(read_region[r]);
        const uint32_t s(reg_seg[g] + seg_n[a]), i(reg_ins[g] + ins_n[a]);
        o->alns[a] = sx_aln{r, e->aln_pos[a], s, i};
        uint32_t ns, ni;
        k8_walk(v, g, a, ns, ni, o->segs + s, o->ins + i);
        if (o->k6_segs)
            for (uint32_t q = e->aln_seg_off[a]; q < e->aln_seg_off[a + 1]; ++q) o->k6_segs[q] = sx_aln_seg{e->segs[q].len, k8_k6_kind(e->segs[q].kind), 0};
    }
    for (uint32_t g = 0; g < b->n_regions; ++g) // k8_pad_kernel
    {
        const uint32_t a0(e->aln_off[b->region_read_off[g]]), a1(e->aln_off[b->region_read_off[g + 1]]);
        uint32_t s(reg_seg[g]), i(reg_ins[g]);
        if (a1 > a0)
        {
            uint32_t ns, ni;
            k8_walk(v, g, a1 - 1, ns, ni, nullptr, nullptr);
            s += seg_n[a1 - 1] + ns;
            i += ins_n[a1 - 1] + ni;
        }
        for (; s < reg_seg[g + 1]; ++s) o->segs[s] = sx_aln_seg{0, SX_SEG_HARDCLIP, 0};
        for (; i < reg_ins[g + 1]; ++i) o->ins[i] = 0;
    }
    return 0;
}
