// tools/synth.cpp -- BENCH/TEST TOOLING (not part of the product ABI): seeded, multi-threaded generator of the synthetic
// workloads of BASELINE.json, written straight into the flattened batch layout of include/strelka_b200.h.
//
//   cfg2  "synthetic 30x germline pileup, 150 bp reads, 1M candidate loci, 4 haplotypes/locus":
//         per locus one region = a 416-base reference window, `depth` reads of `read_len` bases sampled from the locus's diploid
//         genotype over {ref, 3 alt indel alleles} with phred-distributed base errors, each read scored against all 4 haplotype
//         paths (K1); one germline pileup column (K2a); for half of the loci 3 haplotype-vs-reference DP problems (K3).
//   cfg3  somatic 60x/30x pileups (K2b).
//   cfg5  300x depth, 32 haplotypes/locus.
// Qualities are i.i.d. {Q11: 3 %, Q25: 7 %, Q37: 90 %}; alt alleles are insertions/deletions of length
// Geom(0.4) capped at 20.  Every locus has its own counter-based RNG stream, so output is independent of the thread count.
#include "strelka_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
struct rng_t // splitmix64 stream per (seed, locus)
{
    uint64_t s;
    explicit rng_t(uint64_t seed, uint64_t stream) : s(seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull) { next(); }
    inline uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    inline double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

const char BASES[4] = {'A', 'C', 'G', 'T'};
inline uint8_t code_of(char c) { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 15; }
inline uint32_t pad16(uint64_t x) { return (uint32_t)((x + 15) & ~15ull); }

struct allele
{
    bool is_ins;
    uint32_t len;
    char seq[32];
};

struct k1_cfg
{
    uint32_t n_loci, depth, read_len, n_haps, ref_len;
    uint64_t seed;
    uint32_t qual_bits; // 8: one byte per base; 4: dictionary-coded nibbles; 2: 2-bit codes per seq4 nibble position (dictionary {11, 25, 37})
    uint32_t rpr;       // reads per region: a locus deeper than this is cut into several regions sharing the reference window
    uint32_t fmt;       // SX_FMT_ALN8 | SX_FMT_SEG2: write the compact alignment-header / segment wire formats
};

inline uint64_t qual_slice_bytes(const k1_cfg& c, uint32_t n_reads_region)
{
    const uint64_t packed = (c.read_len + 1) / 2;
    return c.qual_bits == 2 ? (n_reads_region * packed + 1) / 2 : c.qual_bits == 4 ? n_reads_region * packed : (uint64_t)n_reads_region * c.read_len;
}

inline void put_aln(const k1_cfg& c, void* alns, const sx_region* reg, uint32_t aidx, uint32_t ridx, uint32_t start, uint32_t seg_abs, uint32_t ins_abs)
{
    if (c.fmt & SX_FMT_ALN8)
    {
        sx_aln8& A = static_cast<sx_aln8*>(alns)[reg->aln_begin + aidx];
        A.read = (uint16_t)ridx;
        A.ref_pos = (int16_t)start;
        A.seg_off = (uint16_t)(seg_abs - reg->seg_begin);
        A.ins_off = (uint16_t)(ins_abs - reg->ins_begin);
    }
    else
    {
        sx_aln& A = static_cast<sx_aln*>(alns)[reg->aln_begin + aidx];
        A.read = reg->read_begin + ridx;
        A.ref_pos = reg->ref_begin + (int32_t)start;
        A.seg_off = seg_abs;
        A.ins_off = ins_abs;
    }
}

inline void put_seg(const k1_cfg& c, void* segs, uint32_t s, uint32_t len, uint32_t kind)
{
    if (c.fmt & SX_FMT_SEG2) static_cast<sx_aln_seg2*>(segs)[s] = (sx_aln_seg2)(len | (kind << 12));
    else static_cast<sx_aln_seg*>(segs)[s] = sx_aln_seg{(uint16_t)len, (uint8_t)kind, 0};
}

inline uint32_t geom_len(rng_t& r)
{
    uint32_t n = 1;
    while (n < 20 && r.unit() > 0.4) ++n;
    return n;
}

// everything that depends only on the locus: alleles, window
struct locus_desc
{
    allele alt[31];
    uint32_t n_alt;
};

inline void make_locus(const k1_cfg& c, uint32_t l, rng_t& r, locus_desc& d, char* ref /*ref_len*/)
{
    for (uint32_t i = 0; i < c.ref_len; ++i) ref[i] = BASES[r.below(4)];
    d.n_alt = c.n_haps - 1;
    for (uint32_t a = 0; a < d.n_alt; ++a)
    {
        d.alt[a].is_ins = r.below(2) != 0;
        d.alt[a].len = geom_len(r);
        for (uint32_t i = 0; i < d.alt[a].len; ++i) d.alt[a].seq[i] = BASES[r.below(4)];
    }
}

// per-region sizes are a pure function of the config except for the insert pool (sum over reads and insertion alleles of the
// observed insert length); pass 1 computes it
struct region_sizes
{
    uint32_t ins_bytes;
};

inline uint32_t pick_qual(rng_t& r)
{
    const double u = r.unit();
    return u < 0.03 ? 11u : u < 0.10 ? 25u : 37u;
}

const double QERR[3] = {0.07943282347242814, 0.0031622776601683794, 0.00019952623149688788}; // 10^-1.1, 10^-2.5, 10^-3.7

void gen_region(const k1_cfg& c, uint32_t rg, bool fill, region_sizes& sz, const sx_region* reg, uint16_t* read_len, uint8_t* seq4, uint8_t* qual, char* ref_pool,
                void* alns, void* segs, char* ins)
{
    const uint32_t chunks = (c.depth + c.rpr - 1) / c.rpr;
    const uint32_t l = rg / chunks, chunk = rg % chunks;
    const uint32_t k_begin = chunk * c.rpr, k_end = std::min(c.depth, k_begin + c.rpr);
    rng_t r(c.seed, l);                      // structure stream of the locus: window, alleles, genotype, read placement
    locus_desc d;
    std::vector<char> refv(c.ref_len);
    make_locus(c, l, r, d, refv.data());
    const uint32_t locus = c.ref_len / 2; // window-relative (the window's contig coordinate is reg->ref_begin, set by synth_k1_plan)
    if (fill)
    {
        if (c.fmt & SX_FMT_REF4) // BAM 4-bit codes, two bases per byte, high nibble first
        {
            uint8_t* rp = reinterpret_cast<uint8_t*>(ref_pool) + reg->ref_off;
            memset(rp, 0, (c.ref_len + 1) / 2);
            for (uint32_t i = 0; i < c.ref_len; ++i) rp[i >> 1] |= code_of(refv[i]) << ((~i & 1) << 2);
        }
        else memcpy(ref_pool + reg->ref_off, refv.data(), c.ref_len);
    }
    // genotype: two haplotype indices
    const uint32_t g0 = r.below(c.n_haps), g1 = r.below(3) ? r.below(c.n_haps) : g0;
    uint32_t ins_off = fill ? reg->ins_begin : 0;
    uint32_t ins_used = 0;
    const uint32_t segs_per_read = 1 + 3 * d.n_alt;
    const uint32_t packed = (c.read_len + 1) / 2;
    std::vector<char> rd(c.read_len);
    if (fill && c.qual_bits == 2 && !(c.fmt & SX_FMT_BASEQ)) memset(qual + reg->qual_off, 0, qual_slice_bytes(c, k_end - k_begin));
    for (uint32_t k = 0; k < k_begin; ++k) // reads of earlier regions of this locus: consume their structure draws
    {
        r.below(c.read_len - 20);
        r.below(2);
    }
    for (uint32_t k = k_begin; k < k_end; ++k)
    {
        rng_t rb(c.seed ^ 0x5DEECE66Dull, (uint64_t)l * c.depth + k); // base-level stream of this read: qualities and sequencing errors
        const uint32_t left = 10 + r.below(c.read_len - 20);     // read bases before the locus
        const uint32_t start = locus - left;                      // window-relative start
        const uint32_t h = r.below(2) ? g0 : g1;
        // source sequence
        if (fill)
        {
            uint32_t n = 0;
            for (uint32_t i = 0; i < left; ++i) rd[n++] = refv[start + i];
            uint32_t rp = locus;
            if (h > 0)
            {
                const allele& a = d.alt[h - 1];
                if (a.is_ins)
                    for (uint32_t i = 0; i < a.len && n < c.read_len; ++i) rd[n++] = a.seq[i];
                else
                    rp += a.len;
            }
            while (n < c.read_len) rd[n++] = refv[rp++];
        }
        const uint32_t ridx = k - k_begin;
        if (fill)
        {
            read_len[reg->read_begin + ridx] = (uint16_t)c.read_len;
            uint8_t* sq = seq4 + reg->seq_off + (uint64_t)ridx * packed;
            uint8_t* ql = qual + reg->qual_off + (c.qual_bits == 2 ? 0 : (uint64_t)ridx * (c.qual_bits == 4 ? packed : c.read_len));
            memset(sq, 0, packed);
            if (c.qual_bits == 4) memset(ql, 0, packed);
            for (uint32_t i = 0; i < c.read_len; ++i)
            {
                const uint32_t q = pick_qual(rb);
                char b = rd[i];
                if (rb.unit() < QERR[q == 11 ? 0 : q == 25 ? 1 : 2]) b = BASES[rb.below(4)];
                const uint32_t qcode = q == 11 ? 0 : q == 25 ? 1 : 2;
                if (c.fmt & SX_FMT_BASEQ)
                {
                    // the quality code rides in the base nibble (below); no quality pool
                }
                else if (c.qual_bits == 2)
                {
                    const uint64_t pos = (uint64_t)ridx * packed * 2 + i; // nibble position in the region's seq4 slice
                    ql[pos >> 2] |= qcode << (6 - 2 * (pos & 3));
                }
                else if (c.qual_bits == 4) ql[i >> 1] |= qcode << ((~i & 1) << 2);
                else ql[i] = (uint8_t)q;
                if (c.fmt & SX_FMT_BASEQ) sq[i >> 1] |= (((b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : 3u) << 2) | qcode) << ((~i & 1) << 2);
                else sq[i >> 1] |= code_of(b) << ((~i & 1) << 2);
            }
        }
        // alignments: one per haplotype
        uint32_t sbase = fill ? reg->seg_begin + ridx * segs_per_read : 0;
        for (uint32_t hh = 0; hh < c.n_haps; ++hh)
        {
            const uint32_t aidx = ridx * c.n_haps + hh;
            if (fill) put_aln(c, alns, reg, aidx, ridx, start, sbase, ins_off);
            if (hh == 0)
            {
                if (fill) put_seg(c, segs, sbase, c.read_len, SX_SEG_MATCH);
                sbase += 1;
            }
            else
            {
                const allele& a = d.alt[hh - 1];
                if (a.is_ins)
                {
                    const uint32_t n = std::min(a.len, c.read_len - left);
                    if (fill)
                    {
                        put_seg(c, segs, sbase + 0, left, SX_SEG_MATCH);
                        put_seg(c, segs, sbase + 1, n, SX_SEG_INSERT);
                        put_seg(c, segs, sbase + 2, c.read_len - left - n, SX_SEG_MATCH);
                        memcpy(ins + ins_off, a.seq, n);
                    }
                    ins_off += n;
                    ins_used += n;
                }
                else
                {
                    if (fill)
                    {
                        put_seg(c, segs, sbase + 0, left, SX_SEG_MATCH);
                        put_seg(c, segs, sbase + 1, a.len, SX_SEG_REFSKIP);
                        put_seg(c, segs, sbase + 2, c.read_len - left, SX_SEG_MATCH);
                    }
                }
                sbase += 3;
            }
        }
    }
    sz.ins_bytes = ins_used;
}

template <typename F> void parallel_for(uint32_t n, int threads, F f)
{
    threads = std::max(1, threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([=] {
            const uint32_t a = (uint32_t)((uint64_t)n * t / threads), b = (uint32_t)((uint64_t)n * (t + 1) / threads);
            for (uint32_t i = a; i < b; ++i) f(i);
        });
    for (auto& x : th) x.join();
}
} // namespace

extern "C" {

struct synth_k1_sizes
{
    uint64_t n_regions, n_reads, n_alns, n_segs, seq4_bytes, qual_bytes, ref_bytes, ins_bytes, cells;
};

// pass 1: region table (needs the per-region insert bytes) + totals.  regions must hold n_loci+1 entries.
int synth_k1_plan(uint32_t n_loci, uint32_t depth, uint32_t read_len, uint32_t n_haps, uint64_t seed, int threads, uint32_t qual_bits, uint32_t reads_per_region,
                  uint32_t fmt, sx_region* regions, synth_k1_sizes* out)
{
    if (n_haps < 1 || n_haps > 32 || read_len < 40 || read_len > 1000) return -1;
    const uint32_t rpr = reads_per_region ? std::min(reads_per_region, depth) : depth;
    const uint32_t chunks = (depth + rpr - 1) / rpr;
    const uint32_t n_regions = n_loci * chunks;
    k1_cfg c{n_loci, depth, read_len, n_haps, 416, seed, qual_bits, rpr, fmt};
    std::vector<uint32_t> ins(n_regions);
    parallel_for(n_regions, threads, [&](uint32_t rg) {
        region_sizes sz;
        gen_region(c, rg, false, sz, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        ins[rg] = sz.ins_bytes;
    });
    const uint32_t packed = (read_len + 1) / 2;
    uint64_t seq = 0, qual = 0, ref = 0, insb = 0, seg = 0, reads = 0;
    for (uint32_t rg = 0; rg <= n_regions; ++rg)
    {
        sx_region& R = regions[rg];
        const uint32_t l = rg / chunks, chunk = rg % chunks;
        const uint32_t nr = rg < n_regions ? std::min(depth, (chunk + 1) * rpr) - chunk * rpr : 0;
        R.seq_off = seq;
        R.qual_off = qual;
        R.ref_off = ref;
        R.read_begin = (uint32_t)reads;
        R.aln_begin = (uint32_t)(reads * n_haps);
        R.seg_begin = (uint32_t)seg;
        R.ins_begin = (uint32_t)insb;
        R.ref_begin = 1000 + (int32_t)(l % 2000000u) * 1000;
        R.ref_len = rg < n_regions ? c.ref_len : 0;
        if (rg == n_regions) break;
        reads += nr;
        seq += pad16((uint64_t)nr * packed);
        qual += (fmt & SX_FMT_BASEQ) ? 0 : pad16(qual_slice_bytes(c, nr));
        ref += pad16((fmt & SX_FMT_REF4) ? (c.ref_len + 1) / 2 : c.ref_len);
        insb += pad16(ins[rg]);
        seg += (fmt & SX_FMT_SEG2) ? (((uint64_t)nr * (1 + 3 * (n_haps - 1)) + 7u) & ~7ull) : (((uint64_t)nr * (1 + 3 * (n_haps - 1)) + 3u) & ~3ull);
    }
    if (seg > 0xffffffffull || insb > 0xffffffffull) return -2;
    out->n_regions = n_regions;
    out->n_reads = (uint64_t)n_loci * depth;
    out->n_alns = (uint64_t)n_loci * depth * n_haps;
    out->n_segs = seg;
    out->seq4_bytes = seq;
    out->qual_bytes = qual;
    out->ref_bytes = ref;
    out->ins_bytes = insb;
    out->cells = (uint64_t)n_loci * depth * n_haps * read_len; // upper bound; exact count via sx_align_batch_cells
    return 0;
}

// pass 2: fill caller-allocated pools (sizes from synth_k1_plan, plus SX_POOL_SLACK; alns has n_alns+1 entries, segs n_segs+16)
int synth_k1_fill(uint32_t n_loci, uint32_t depth, uint32_t read_len, uint32_t n_haps, uint64_t seed, int threads, uint32_t qual_bits, uint32_t reads_per_region,
                  uint32_t fmt, const sx_region* regions,
                  uint16_t* read_lens, uint8_t* seq4, uint8_t* qual, char* ref, void* alns, void* segs, char* ins)
{
    const uint32_t rpr = reads_per_region ? std::min(reads_per_region, depth) : depth;
    const uint32_t n_regions = n_loci * ((depth + rpr - 1) / rpr);
    k1_cfg c{n_loci, depth, read_len, n_haps, 416, seed, qual_bits, rpr, fmt};
    const uint32_t n_segs = regions[n_regions].seg_begin;
    for (uint32_t i = 0; i < n_segs + 16; ++i) put_seg(c, segs, i, 0, SX_SEG_HARDCLIP);
    parallel_for(n_regions, threads, [&](uint32_t rg) {
        region_sizes sz;
        gen_region(c, rg, true, sz, &regions[rg], read_lens, seq4, qual, ref, alns, segs, ins);
    });
    if (fmt & SX_FMT_ALN8)
    {
        static_cast<sx_aln8*>(alns)[(uint64_t)n_loci * depth * n_haps] = sx_aln8{0, 0, 0, 0}; // unused sentinel
        return 0;
    }
    sx_aln& S = static_cast<sx_aln*>(alns)[(uint64_t)n_loci * depth * n_haps];
    S.read = n_loci * depth;
    S.ref_pos = 0;
    S.seg_off = n_segs;
    S.ins_off = regions[n_regions].ins_begin;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// K6 view of the same cfg2/cfg5 workload: per region the IndelBuffer window = the distinct alt alleles of its locus (IndelKey
// order), per read the same n_haps candidate alignments as the K1 batch (same RNG draws, same order), each with its path in K4
// kinds and the window index of its indel.  An insertion that runs off the read end is an edge insertion: not in the indel set.
// ---------------------------------------------------------------------------------------------------------------------------
struct synth_k6_sizes
{
    uint64_t n_regions, n_reads, n_alns, n_keys, n_segs, n_aln_keys, n_slots;
};
// the insert sequences behind ins_id are also available (tests / the reference harness rebuild IndelKeys from them): key_ins[32 * key]

namespace
{
struct k6_counts
{
    uint32_t keys, segs, akeys;
};

struct k6_out
{
    uint32_t *region_read_off, *region_key_off, *aln_off, *aln_seg_off, *aln_key_off, *rec_off;
    sx_indel_key* keys;
    int32_t* aln_pos;
    sx_aln_seg* segs;
    uint16_t *aln_keys, *read_len, *non_ambig;
    uint8_t* read_flags;
    char* key_ins; // optional: 32 bytes per key, the insert sequence (zero padded)
};

void gen_region_k6(const k1_cfg& c, uint32_t rg, bool fill, k6_counts& cnt, const k6_counts* base /*prefix of this region*/, const k6_out* o)
{
    const uint32_t chunks = (c.depth + c.rpr - 1) / c.rpr;
    const uint32_t l = rg / chunks, chunk = rg % chunks;
    const uint32_t k_begin = chunk * c.rpr, k_end = std::min(c.depth, k_begin + c.rpr);
    rng_t r(c.seed, l);
    locus_desc d;
    std::vector<char> refv(c.ref_len);
    make_locus(c, l, r, d, refv.data());
    const uint32_t locus = c.ref_len / 2;
    const int32_t ref_begin = 1000 + (int32_t)(l % 2000000u) * 1000;
    r.below(c.n_haps); // g0
    if (r.below(3)) r.below(c.n_haps); // g1
    // window: distinct alleles in IndelKey order (same pos and type: insert length, delete length, insert sequence)
    uint32_t order[31], n_win = 0, win_of[31];
    auto less = [&](uint32_t x, uint32_t y) {
        const allele &a = d.alt[x], &b = d.alt[y];
        const uint32_t ai = a.is_ins ? a.len : 0, bi = b.is_ins ? b.len : 0, ad = a.is_ins ? 0 : a.len, bd = b.is_ins ? 0 : b.len;
        if (ai != bi) return ai < bi;
        if (ad != bd) return ad < bd;
        return a.is_ins && memcmp(a.seq, b.seq, a.len) < 0;
    };
    for (uint32_t a = 0; a < d.n_alt; ++a) order[a] = a;
    std::sort(order, order + d.n_alt, less);
    for (uint32_t i = 0; i < d.n_alt; ++i)
    {
        if (i > 0 && !less(order[i - 1], order[i])) win_of[order[i]] = n_win - 1; // equal to its predecessor
        else win_of[order[i]] = n_win++;
    }
    cnt.keys = n_win;
    const uint32_t first_read = l * c.depth + k_begin; // == region's read_begin in the K1 batch
    if (fill)
    {
        o->region_read_off[rg] = first_read;
        o->region_key_off[rg] = base->keys;
        uint32_t w = 0;
        for (uint32_t i = 0; i < d.n_alt; ++i)
        {
            if (win_of[order[i]] != w) continue;
            const allele& a = d.alt[order[i]];
            sx_indel_key& k = o->keys[base->keys + w];
            memset(&k, 0, sizeof(k));
            k.pos = ref_begin + (int32_t)locus;
            k.del_len = a.is_ins ? 0 : (uint16_t)a.len;
            k.ins_len = a.is_ins ? (uint16_t)a.len : 0;
            k.ins_id = a.is_ins ? w + 1 : 0;
            k.type = SX_INDEL_TYPE_INDEL;
            k.flags = SX_IKF_CANDIDATE;
            k.ref_to_indel_lnp = -9.903487552536127; // ln 5e-5
            k.indel_to_ref_lnp = -9.903487552536127;
            if (o->key_ins)
            {
                memset(o->key_ins + 32ull * (base->keys + w), 0, 32);
                if (a.is_ins) memcpy(o->key_ins + 32ull * (base->keys + w), a.seq, a.len);
            }
            ++w;
        }
    }
    for (uint32_t k = 0; k < k_begin; ++k)
    {
        r.below(c.read_len - 20);
        r.below(2);
    }
    uint32_t segs = 0, akeys = 0;
    for (uint32_t k = k_begin; k < k_end; ++k)
    {
        const uint32_t left = 10 + r.below(c.read_len - 20);
        const uint32_t start = locus - left;
        r.below(2);
        const uint32_t read = first_read + (k - k_begin);
        if (fill)
        {
            o->aln_off[read] = read * c.n_haps;
            o->read_len[read] = (uint16_t)c.read_len;
            o->non_ambig[read] = (uint16_t)c.read_len;
            o->read_flags[read] = SX_SIF_FWD | SX_SIF_TIER1;
            o->rec_off[read] = 0; // per-read slot counts first; prefix-summed by synth_k6_fill
        }
        for (uint32_t hh = 0; hh < c.n_haps; ++hh)
        {
            const uint32_t a_idx = read * c.n_haps + hh;
            sx_aln_seg* sg = fill ? o->segs + base->segs + segs : nullptr;
            if (fill)
            {
                o->aln_pos[a_idx] = ref_begin + (int32_t)start;
                o->aln_seg_off[a_idx] = base->segs + segs;
                o->aln_key_off[a_idx] = base->akeys + akeys;
            }
            if (hh == 0)
            {
                if (fill) sg[0] = sx_aln_seg{(uint16_t)c.read_len, SX_SEG_MATCH, 0};
                segs += 1;
                continue;
            }
            const allele& a = d.alt[hh - 1];
            bool in_set = true;
            if (a.is_ins)
            {
                const uint32_t n = std::min(a.len, c.read_len - left), rest = c.read_len - left - n;
                if (fill)
                {
                    sg[0] = sx_aln_seg{(uint16_t)left, SX_SEG_MATCH, 0};
                    sg[1] = sx_aln_seg{(uint16_t)n, SX_SEG_INSERT, 0};
                    if (rest) sg[2] = sx_aln_seg{(uint16_t)rest, SX_SEG_MATCH, 0};
                }
                segs += rest ? 3 : 2;
                in_set = rest != 0;
            }
            else
            {
                if (fill)
                {
                    sg[0] = sx_aln_seg{(uint16_t)left, SX_SEG_MATCH, 0};
                    sg[1] = sx_aln_seg{(uint16_t)a.len, SX_SEG_DELETE, 0};
                    sg[2] = sx_aln_seg{(uint16_t)(c.read_len - left), SX_SEG_MATCH, 0};
                }
                segs += 3;
            }
            if (in_set)
            {
                if (fill) o->aln_keys[base->akeys + akeys] = (uint16_t)win_of[hh - 1];
                akeys += 1;
            }
        }
    }
    cnt.segs = segs;
    cnt.akeys = akeys;
}
} // namespace

// pass 1: per-region counts (counts[3 * n_regions]: keys, segs, alignment keys) and totals
int synth_k6_plan(uint32_t n_loci, uint32_t depth, uint32_t read_len, uint32_t n_haps, uint64_t seed, int threads, uint32_t reads_per_region, uint32_t* counts,
                  synth_k6_sizes* out)
{
    if (n_haps < 1 || n_haps > 32 || read_len < 40 || read_len > 1000) return -1;
    const uint32_t rpr = reads_per_region ? std::min(reads_per_region, depth) : depth;
    const uint32_t n_regions = n_loci * ((depth + rpr - 1) / rpr);
    k1_cfg c{n_loci, depth, read_len, n_haps, 416, seed, 8, rpr, 0};
    parallel_for(n_regions, threads, [&](uint32_t rg) {
        k6_counts cnt;
        gen_region_k6(c, rg, false, cnt, nullptr, nullptr);
        counts[3 * rg] = cnt.keys;
        counts[3 * rg + 1] = cnt.segs;
        counts[3 * rg + 2] = cnt.akeys;
    });
    memset(out, 0, sizeof(*out));
    out->n_regions = n_regions;
    out->n_reads = (uint64_t)n_loci * depth;
    out->n_alns = out->n_reads * n_haps;
    const uint32_t chunks = (depth + rpr - 1) / rpr;
    for (uint32_t rg = 0; rg < n_regions; ++rg)
    {
        const uint32_t chunk = rg % chunks, nr = std::min(depth, (chunk + 1) * rpr) - chunk * rpr;
        out->n_keys += counts[3 * rg];
        out->n_segs += counts[3 * rg + 1];
        out->n_aln_keys += counts[3 * rg + 2];
        out->n_slots += (uint64_t)nr * counts[3 * rg];
    }
    return 0;
}

// pass 2: fill caller-allocated arrays ([n+1] offset arrays, keys[n_keys], segs[n_segs], aln_keys[n_aln_keys], per-read arrays)
int synth_k6_fill(uint32_t n_loci, uint32_t depth, uint32_t read_len, uint32_t n_haps, uint64_t seed, int threads, uint32_t reads_per_region, const uint32_t* counts,
                  uint32_t* region_read_off, uint32_t* region_key_off, sx_indel_key* keys, uint32_t* aln_off, int32_t* aln_pos, uint32_t* aln_seg_off, sx_aln_seg* segs,
                  uint32_t* aln_key_off, uint16_t* aln_keys, uint16_t* read_lens, uint16_t* non_ambig, uint8_t* read_flags, uint32_t* rec_off, char* key_ins /*optional, 32 bytes per key*/)
{
    const uint32_t rpr = reads_per_region ? std::min(reads_per_region, depth) : depth;
    const uint32_t chunks = (depth + rpr - 1) / rpr;
    const uint32_t n_regions = n_loci * chunks;
    k1_cfg c{n_loci, depth, read_len, n_haps, 416, seed, 8, rpr, 0};
    std::vector<k6_counts> base(n_regions + 1);
    k6_counts run{0, 0, 0};
    for (uint32_t rg = 0; rg < n_regions; ++rg)
    {
        base[rg] = run;
        run.keys += counts[3 * rg];
        run.segs += counts[3 * rg + 1];
        run.akeys += counts[3 * rg + 2];
    }
    base[n_regions] = run;
    k6_out o{region_read_off, region_key_off, aln_off, aln_seg_off, aln_key_off, rec_off, keys, aln_pos, segs, aln_keys, read_lens, non_ambig, read_flags, key_ins};
    parallel_for(n_regions, threads, [&](uint32_t rg) {
        k6_counts cnt;
        gen_region_k6(c, rg, true, cnt, &base[rg], &o);
    });
    const uint64_t n_reads = (uint64_t)n_loci * depth, n_alns = n_reads * n_haps;
    region_read_off[n_regions] = (uint32_t)n_reads;
    region_key_off[n_regions] = run.keys;
    aln_off[n_reads] = (uint32_t)n_alns;
    aln_seg_off[n_alns] = run.segs;
    aln_key_off[n_alns] = run.akeys;
    uint64_t slots = 0; // a read may evaluate every entry of its region's window
    for (uint32_t rg = 0; rg < n_regions; ++rg)
    {
        const uint32_t chunk = rg % chunks, nr = std::min(depth, (chunk + 1) * rpr) - chunk * rpr;
        for (uint32_t i = 0; i < nr; ++i)
        {
            rec_off[region_read_off[rg] + i] = (uint32_t)slots;
            slots += counts[3 * rg];
        }
    }
    rec_off[n_reads] = (uint32_t)slots;
    return slots > 0xffffffffull ? -2 : 0;
}

// pileup columns: Poisson-ish depth (binomial approximation via sum of uniforms is avoided: exact inverse-CDF Poisson), a
// site is hom-ref (80 %), het (13 %) or hom-alt (7 %) for germline; for the tumour sample `vaf` gives the alt fraction.
static uint32_t poisson(rng_t& r, double mean)
{
    // Knuth for small means, normal approximation above 60
    if (mean > 60)
    {
        double u = 0;
        for (int i = 0; i < 12; ++i) u += r.unit();
        const double v = mean + (u - 6.0) * std::sqrt(mean);
        return v < 0 ? 0u : (uint32_t)(v + 0.5);
    }
    const double L = std::exp(-mean);
    uint32_t k = 0;
    double p = 1.0;
    do
    {
        ++k;
        p *= r.unit();
    } while (p > L);
    return k - 1;
}

// mode 0: germline site; mode 1: somatic normal; mode 2: somatic tumour.  site_off must hold n_sites+1 entries; pass calls=NULL to size.
uint64_t synth_pileups(uint32_t n_sites, double depth, int mode, uint64_t seed, int threads, uint32_t* site_off, uint16_t* calls, char* ref_base)
{
    std::vector<uint32_t> cnt(n_sites);
    parallel_for(n_sites, threads, [&](uint32_t s) {
        rng_t r(seed ^ 0xABCDEFull, s);
        cnt[s] = poisson(r, depth);
    });
    uint64_t off = 0;
    for (uint32_t s = 0; s < n_sites; ++s)
    {
        site_off[s] = (uint32_t)off;
        off += cnt[s];
    }
    site_off[n_sites] = (uint32_t)off;
    if (!calls) return off;
    parallel_for(n_sites, threads, [&](uint32_t s) {
        rng_t r(seed ^ 0xABCDEFull, s);
        const uint32_t n = poisson(r, depth);
        rng_t g(seed ^ 0x13579Bull, s); // genotype stream shared by the normal/tumour samples of a site
        const uint32_t ref_id = g.below(4), alt_id = (ref_id + 1 + g.below(3)) & 3;
        const double u = g.unit();
        double af;
        if (mode == 0) af = u < 0.80 ? 0.0 : u < 0.93 ? 0.5 : 1.0;
        else if (mode == 1) af = u < 0.97 ? 0.0 : 0.5;                       // normal: mostly hom-ref, some germline hets
        else
        {
            const double v = g.unit();
            af = u < 0.97 ? (v < 0.6 ? 0.0 : v < 0.7 ? 0.05 : v < 0.8 ? 0.1 : v < 0.9 ? 0.2 : 0.4) : 0.5; // tumour: somatic VAF mix on hom-ref normals
        }
        ref_base[s] = BASES[ref_id];
        uint16_t* c = calls + site_off[s];
        for (uint32_t i = 0; i < n; ++i)
        {
            const uint32_t q = pick_qual(r);
            uint32_t b = (r.unit() < af) ? alt_id : ref_id;
            if (r.unit() < QERR[q == 11 ? 0 : q == 25 ? 1 : 2]) b = r.below(4);
            const uint32_t fwd = r.below(2), nbr = r.unit() < 0.05, filt = r.unit() < 0.03;
            c[i] = SX_CALL(q, b, fwd, nbr, filt, 0);
        }
    });
    return off;
}

// haplotype-vs-reference DP problems: reference segment of 30..100 bases, query = segment with 1..3 edits (SNV / short indel)
uint64_t synth_ga(uint32_t n, uint64_t seed, int threads, uint32_t* q_off, uint32_t* r_off, char* query, char* ref)
{
    std::vector<uint32_t> ql(n), rl(n);
    auto gen = [&](uint32_t i, char* q, char* rf, uint32_t& Q, uint32_t& R) {
        rng_t r(seed ^ 0x777ull, i);
        R = 30 + r.below(71);
        char rb[128], qb[192];
        for (uint32_t k = 0; k < R; ++k) rb[k] = BASES[r.below(4)];
        uint32_t n_q = 0;
        const uint32_t n_ed = 1 + r.below(3);
        uint32_t ed_pos[3], ed_kind[3], ed_len[3];
        for (uint32_t e = 0; e < n_ed; ++e)
        {
            ed_pos[e] = 5 + r.below(R - 10);
            ed_kind[e] = r.below(3);
            ed_len[e] = 1 + r.below(8);
        }
        for (uint32_t k = 0; k < R; ++k)
        {
            bool skip = false;
            for (uint32_t e = 0; e < n_ed; ++e)
            {
                if (ed_pos[e] == k)
                {
                    if (ed_kind[e] == 0) { qb[n_q++] = BASES[r.below(4)]; skip = true; }
                    else if (ed_kind[e] == 1) { for (uint32_t z = 0; z < ed_len[e] && n_q < 180; ++z) qb[n_q++] = BASES[r.below(4)]; }
                }
                if (ed_kind[e] == 2 && k >= ed_pos[e] && k < ed_pos[e] + ed_len[e]) skip = true;
            }
            if (!skip && n_q < 190) qb[n_q++] = rb[k];
        }
        if (n_q == 0) qb[n_q++] = 'A';
        Q = n_q;
        if (q) memcpy(q, qb, Q);
        if (rf) memcpy(rf, rb, R);
    };
    parallel_for(n, threads, [&](uint32_t i) { gen(i, nullptr, nullptr, ql[i], rl[i]); });
    uint64_t qo = 0, ro = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        q_off[i] = (uint32_t)qo;
        r_off[i] = (uint32_t)ro;
        qo += ql[i];
        ro += rl[i];
    }
    q_off[n] = (uint32_t)qo;
    r_off[n] = (uint32_t)ro;
    if (!query) return qo | (ro << 32);
    parallel_for(n, threads, [&](uint32_t i) {
        uint32_t Q, R;
        gen(i, query + q_off[i], ref + r_off[i], Q, R);
    });
    return qo | (ro << 32);
}
}
