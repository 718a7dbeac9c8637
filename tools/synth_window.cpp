// tools/synth_window.cpp -- BENCH/TEST TOOLING (not part of the product ABI): the whole-path workload of BASELINE.json's cfg2 as ONE
// description per window of positions -- what sx_process_window_dev (and the reference's align_pos / pileup_pos_reads / per-site
// genotyping) takes: a contig segment tiled by candidate loci, reads at 30x in read-buffer order with the MAPPER's alignments, the
// candidate indels around every locus.  SURVEY.md 8(d) cfg2: loci spaced 300 bp, each with up to 3 candidate alleles (indel lengths
// Geom(0.4) capped at 20: a deletion and an insertion at the locus and a second deletion one base on -- overlapping alternatives),
// genotype het : hom 2 : 1, an SNV within 10 bp of half of the loci, reads of 150 bp at uniform starts (60 per 300 bp cell = 30x) sampled
// from the two haplotypes, 0.5 % base errors, qualities {Q11 3 %, Q25 7 %, Q37 90 %}, MAPQ 60, both strands.
//
// The mapper is imitated the way a seed-and-extend mapper answers: a read that carries the locus's indel with at least 20 bases on both
// sides is aligned with the indel in its CIGAR; closer to an end the short side is left as mismatches (the alignment is anchored on the
// long side, so an indel near the read's start shifts its position) -- the cases the realigner exists for.
//
// Every cell has its own counter-based RNG stream (output independent of the thread count).  Two passes: sizes, then fill.
#include "strelka_b200.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
struct rng_t // splitmix64 stream per (seed, cell)
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
inline uint8_t code_of(char c) { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : 8; }
inline uint64_t pad16(uint64_t x) { return (x + 15) & ~15ull; }
inline uint32_t geom(rng_t& g) // Geom(0.4) capped at 20
{
    uint32_t n(1);
    while (n < 20 && g.unit() >= 0.4) ++n;
    return n;
}
// reference base at contig position p: a hash of the position, so neighbouring cells agree without sharing state
inline char ref_base(uint64_t seed, int64_t p)
{
    uint64_t z = (uint64_t)p * 0x9E3779B97F4A7C15ull + seed * 0xD6E8FEB86659FD93ull;
    z = (z ^ (z >> 32)) * 0xD6E8FEB86659FD93ull;
    z ^= z >> 29;
    return BASES[z & 3];
}

struct cfg_t
{
    uint32_t n_cells, reads_per_cell, read_len, cell_len;
    uint64_t seed;
    int32_t contig_begin; // contig position of the first cell; the reference array starts REF_LEAD bases before it
    uint32_t qual_bits;   // 8 or 4 (dictionary {11, 25, 37})
};
constexpr int32_t REF_LEAD = 64, REF_TAIL = 320; // reference bases carried before the first / after the last cell
constexpr uint32_t CENTRE = 225;                 // offset of the locus inside its cell: a read overlaps at most its own cell's alleles

struct read_t
{
    int32_t start;    // true start (first reference base of the haplotype the read was sampled from, in reference coordinates before the locus)
    int32_t map_pos;  // the mapper's position
    uint8_t n_seg;
    sx_aln_seg seg[3];
    uint8_t fwd;
    char bases[160];
    uint8_t qcode[160];
};

struct cell_t
{
    int32_t c; // locus position
    uint32_t d0, i1, d2;
    char ins[24];
    int allele;   // -1 none, 0 del(c, d0), 1 ins(c, i1), 2 del(c+1, d2)
    bool hom;
    int32_t snv_pos; // or -1
    char snv_base;
};

// light: only what the sizes need (positions, segment counts) -- every read has its own RNG stream, so skipping the bases changes nothing else
void make_cell(const cfg_t& C, uint32_t cell, cell_t& L, std::vector<read_t>& reads, const bool light = false)
{
    rng_t g(C.seed, cell);
    const int32_t cs(C.contig_begin + (int32_t)(cell * C.cell_len));
    L.c = cs + (int32_t)CENTRE;
    L.d0 = geom(g);
    L.i1 = geom(g);
    L.d2 = geom(g);
    for (uint32_t i = 0; i < L.i1; ++i) L.ins[i] = BASES[g.below(4)];
    // an inserted sequence must not begin with the reference base that follows it (it would be the same haplotype as a shifted insertion)
    if (L.ins[L.i1 - 1] == ref_base(C.seed, L.c - 1)) L.ins[L.i1 - 1] = BASES[(g.below(3) + 1 + (uint32_t)(std::find(BASES, BASES + 4, ref_base(C.seed, L.c - 1)) - BASES)) & 3];
    const double u(g.unit());
    L.allele = u < 0.1 ? -1 : (int)g.below(3);
    L.hom = g.unit() < (1.0 / 3.0);
    L.snv_pos = -1;
    if (g.unit() < 0.5)
    {
        L.snv_pos = L.c - 1 - (int32_t)g.below(10);
        const char rb(ref_base(C.seed, L.snv_pos));
        L.snv_base = BASES[((uint32_t)(std::find(BASES, BASES + 4, rb) - BASES) + 1 + g.below(3)) & 3];
    }
    const uint32_t R(C.read_len);
    reads.resize(C.reads_per_cell);
    for (uint32_t k = 0; k < C.reads_per_cell; ++k)
    {
        read_t& rd(reads[k]);
        rng_t g(C.seed ^ 0x5bd1e9955bd1e995ull, (uint64_t)cell * 256u + k + 1u); // (shadows the cell's stream: per-read draws from here on)
        const int32_t s(cs + (int32_t)g.below(C.cell_len));
        const bool hapA(g.unit() < 0.5);
        const bool alt(L.allele >= 0 && (hapA || L.hom));  // hapA carries the alt indel allele, hapB too when homozygous
        const bool snv(L.snv_pos >= 0 && hapA);            // the SNV is heterozygous, on hapA
        rd.fwd = g.unit() < 0.5;
        // haplotype walk: reference position p, emitting R bases
        const int32_t ev_pos(L.allele == 2 ? L.c + 1 : L.c);
        int32_t s_eff(s);
        bool done_event(!alt);
        if (alt && L.allele != 1 && s == ev_pos) // a read that begins right behind the deleted bases: its first base is the one after them
        {
            s_eff = s + (int32_t)(L.allele == 0 ? L.d0 : L.d2);
            done_event = true;
        }
        rd.start = s_eff;
        int32_t p(s_eff);
        uint32_t n(0);
        int32_t indel_read_off(-1); // read offset at which the indel event sits (bases before it)
        while (n < R)
        {
            if (!done_event && p == ev_pos)
            {
                indel_read_off = (int32_t)n;
                done_event = true;
                if (L.allele == 1)
                {
                    for (uint32_t i = 0; i < L.i1 && n < R; ++i) rd.bases[n++] = L.ins[i];
                    continue;
                }
                p += (int32_t)(L.allele == 0 ? L.d0 : L.d2);
                continue;
            }
            rd.bases[n++] = light ? 'A' : ((snv && p == L.snv_pos) ? L.snv_base : ref_base(C.seed, p));
            ++p;
        }
        // sequencing errors and qualities.  The inserted bases of a read whose CIGAR will carry the insertion stay clean: an erroneous insert
        // would be a private indel of that read, which the reference's IndelBuffer would hold as a non-candidate observation (insert_indel,
        // starling_pos_processor_base.cpp:397) -- the windows here hold the candidate alleles only
        const bool cigar_ins(alt && L.allele == 1 && indel_read_off >= 20 && (int32_t)R - indel_read_off - (int32_t)std::min<uint32_t>(L.i1, R - (uint32_t)indel_read_off) >= 20);
        for (uint32_t i = 0; i < R && !light; ++i)
        {
            const uint64_t x(g.next()); // one draw per base: 16 bits for the quality, 16 for the error, 2 for the erroneous base
            const uint32_t q((uint32_t)(x & 0xFFFFu)), e((uint32_t)((x >> 16) & 0xFFFFu));
            rd.qcode[i] = q < 1966u ? 0 : q < 6554u ? 1 : 2; // 3 % / 7 % / 90 %
            if (e < 328u && !(cigar_ins && (int32_t)i >= indel_read_off && i < (uint32_t)indel_read_off + L.i1)) rd.bases[i] = BASES[(x >> 32) & 3]; // 0.5 %
        }
        // the mapper's answer
        rd.map_pos = s_eff;
        rd.n_seg = 1;
        rd.seg[0] = sx_aln_seg{(uint16_t)R, SX_AP_MATCH, 0};
        if (alt && indel_read_off >= 0)
        {
            const uint32_t left((uint32_t)indel_read_off);
            const uint32_t ev_len(L.allele == 1 ? std::min<uint32_t>(L.i1, R - left) : 0u);
            const uint32_t right(R - left - ev_len);
            const uint32_t dlen(L.allele == 0 ? L.d0 : L.allele == 2 ? L.d2 : 0u);
            if (left >= 20 && right >= 20)
            {
                rd.n_seg = 3;
                rd.seg[0] = sx_aln_seg{(uint16_t)left, SX_AP_MATCH, 0};
                rd.seg[1] = L.allele == 1 ? sx_aln_seg{(uint16_t)ev_len, SX_AP_INSERT, 0} : sx_aln_seg{(uint16_t)dlen, SX_AP_DELETE, 0};
                rd.seg[2] = sx_aln_seg{(uint16_t)right, SX_AP_MATCH, 0};
            }
            else if (left < right) // anchored on the long right side: the whole read as a match, shifted by the event
                rd.map_pos = L.allele == 1 ? s_eff - (int32_t)ev_len : s_eff + (int32_t)dlen;
            // else: anchored on the left, the position stands
        }
    }
    std::stable_sort(reads.begin(), reads.end(), [](const read_t& a, const read_t& b) { return a.map_pos < b.map_pos; });
}

template <typename F> void parallel_cells(uint32_t n, uint32_t threads, F f)
{
    threads = std::max(1u, std::min(threads, n));
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < threads; ++t)
        th.emplace_back([=]() {
            const uint64_t a((uint64_t)n * t / threads), b((uint64_t)n * (t + 1) / threads);
            for (uint64_t i = a; i < b; ++i) f((uint32_t)i);
        });
    for (auto& x : th) x.join();
}
} // namespace

struct synth_window_sizes
{
    uint64_t n_regions, n_reads, n_keys, n_raw_segs, seq4_bytes, qual_bytes, ref_bytes, key_ins_bytes, n_sites;
};

// pass 1: seg_count[n_cells] <- raw segments of each cell's reads; sizes
extern "C" int synth_window_plan(uint32_t n_cells, uint32_t reads_per_cell, uint32_t read_len, uint32_t cell_len, uint64_t seed, int32_t contig_begin, uint32_t qual_bits,
                                 uint32_t threads, uint32_t* seg_count, uint32_t* ins_count, synth_window_sizes* out)
{
    if (read_len > 160 || cell_len < 2 * CENTRE / 3 || (qual_bits != 8 && qual_bits != 4) || (contig_begin - REF_LEAD) % 16 != 0) return 1;
    const cfg_t C{n_cells, reads_per_cell, read_len, cell_len, seed, contig_begin, qual_bits};
    parallel_cells(n_cells, threads, [&](uint32_t cell) {
        cell_t L;
        std::vector<read_t> reads;
        make_cell(C, cell, L, reads, true);
        uint32_t n(0);
        for (const read_t& r : reads) n += r.n_seg;
        seg_count[cell] = n;
        ins_count[cell] = L.i1;
    });
    uint64_t ns(0), ni(0);
    for (uint32_t i = 0; i < n_cells; ++i)
    {
        ns += seg_count[i];
        ni += ins_count[i];
    }
    const uint64_t packed((read_len + 1) / 2);
    out->n_regions = n_cells;
    out->n_reads = (uint64_t)n_cells * reads_per_cell;
    out->n_keys = 3ull * n_cells;
    out->n_raw_segs = ns;
    out->seq4_bytes = (uint64_t)n_cells * pad16(reads_per_cell * packed);
    out->qual_bytes = (uint64_t)n_cells * pad16(qual_bits == 4 ? reads_per_cell * packed : (uint64_t)reads_per_cell * read_len);
    out->ref_bytes = (uint64_t)REF_LEAD + (uint64_t)n_cells * cell_len + REF_TAIL;
    out->key_ins_bytes = ni;
    out->n_sites = (uint64_t)n_cells * cell_len;
    return 0;
}

// pass 2: every array of the window (see include/strelka_b200.h sx_window_batch); ASCII read bases / one-byte qualities are written too
// when read_ascii / qual_wide are given (what the reference harness takes)
extern "C" int synth_window_fill(uint32_t n_cells, uint32_t reads_per_cell, uint32_t read_len, uint32_t cell_len, uint64_t seed, int32_t contig_begin, uint32_t qual_bits,
                                 uint32_t threads, const uint32_t* seg_count, const uint32_t* ins_count, uint32_t* region_read_off, uint32_t* region_key_off,
                                 sx_indel_key* keys, uint32_t* key_ins_off, char* key_ins, int32_t* realign_begin, int32_t* realign_end, int32_t* raw_pos,
                                 uint32_t* raw_seg_off, sx_aln_seg* raw_segs, uint16_t* read_len_out, uint8_t* read_flags, uint8_t* mapq, uint32_t* use_key_off,
                                 uint32_t* rec_off, sx_region* regions, uint8_t* seq4, uint8_t* qual, char* ref, char* read_ascii, uint8_t* qual_wide,
                                 double ref_to_indel_lnp, double indel_to_ref_lnp)
{
    const cfg_t C{n_cells, reads_per_cell, read_len, cell_len, seed, contig_begin, qual_bits};
    static const uint8_t QD[3] = {11, 25, 37};
    const uint64_t packed((read_len + 1) / 2);
    const uint64_t seq_stride(pad16(reads_per_cell * packed)), qual_stride(pad16(qual_bits == 4 ? reads_per_cell * packed : (uint64_t)reads_per_cell * read_len));
    const int32_t ref_begin(contig_begin - REF_LEAD);
    const uint64_t ref_bytes((uint64_t)REF_LEAD + (uint64_t)n_cells * cell_len + REF_TAIL);
    // prefix sums of the per-cell counts
    std::vector<uint64_t> seg_base(n_cells + 1, 0), ins_base(n_cells + 1, 0);
    for (uint32_t i = 0; i < n_cells; ++i)
    {
        seg_base[i + 1] = seg_base[i] + seg_count[i];
        ins_base[i + 1] = ins_base[i] + ins_count[i];
    }
    parallel_cells((uint32_t)((ref_bytes + 4095) / 4096), threads, [&](uint32_t blk) {
        const uint64_t a((uint64_t)blk * 4096), b(std::min<uint64_t>(ref_bytes, a + 4096));
        for (uint64_t i = a; i < b; ++i) ref[i] = ref_base(seed, (int64_t)ref_begin + (int64_t)i);
    });
    parallel_cells(n_cells, threads, [&](uint32_t cell) {
        cell_t L;
        std::vector<read_t> reads;
        make_cell(C, cell, L, reads);
        const int32_t cs(contig_begin + (int32_t)(cell * cell_len));
        const uint32_t r0(cell * reads_per_cell), k0(3 * cell);
        region_read_off[cell] = r0;
        region_key_off[cell] = k0;
        // window entries in IndelKey order (IndelKey.hh:53-76: position, type, INSERT length, delete length, insert sequence): at the locus the
        // deletion (insert length 0) sorts before the insertion
        sx_indel_key K[3];
        memset(K, 0, sizeof(K));
        K[0].pos = L.c;
        K[0].del_len = (uint16_t)L.d0;
        K[1].pos = L.c;
        K[1].ins_len = (uint16_t)L.i1;
        K[1].ins_id = 1;
        K[2].pos = L.c + 1;
        K[2].del_len = (uint16_t)L.d2;
        for (int k = 0; k < 3; ++k)
        {
            K[k].type = SX_INDEL_TYPE_INDEL;
            K[k].flags = SX_IKF_CANDIDATE;
            K[k].ref_to_indel_lnp = ref_to_indel_lnp;
            K[k].indel_to_ref_lnp = indel_to_ref_lnp;
            keys[k0 + k] = K[k];
        }
        key_ins_off[k0] = key_ins_off[k0 + 1] = (uint32_t)ins_base[cell];
        key_ins_off[k0 + 2] = (uint32_t)(ins_base[cell] + L.i1);
        memcpy(key_ins + ins_base[cell], L.ins, L.i1);
        realign_begin[cell] = std::max(0, cs - 400);
        realign_end[cell] = cs + (int32_t)cell_len + 600;
        // region record: read / quality slices, reference window as a view of the contig array (16-byte aligned start)
        sx_region& reg(regions[cell]);
        memset(&reg, 0, sizeof(reg));
        reg.seq_off = (uint64_t)cell * seq_stride;
        reg.qual_off = (uint64_t)cell * qual_stride;
        const int32_t wb(ref_begin + (int32_t)(((uint32_t)(cs - 48 - ref_begin)) & ~15u));
        reg.ref_off = (uint64_t)(wb - ref_begin);
        reg.ref_begin = wb;
        reg.ref_len = (uint32_t)std::min<int64_t>((int64_t)cs + cell_len + read_len + 96 - wb, (int64_t)ref_bytes - (int64_t)reg.ref_off);
        reg.read_begin = r0;
        uint8_t* sq(seq4 + reg.seq_off);
        uint8_t* ql(qual + reg.qual_off);
        memset(sq, 0, seq_stride);
        memset(ql, 0, qual_stride);
        uint64_t so(seg_base[cell]);
        for (uint32_t k = 0; k < reads_per_cell; ++k)
        {
            const read_t& rd(reads[k]);
            const uint32_t r(r0 + k);
            raw_pos[r] = rd.map_pos;
            raw_seg_off[r] = (uint32_t)so;
            for (uint32_t s = 0; s < rd.n_seg; ++s) raw_segs[so++] = rd.seg[s];
            read_len_out[r] = (uint16_t)read_len;
            read_flags[r] = (uint8_t)((rd.fwd ? SX_PRF_FWD : 0u) | SX_PRF_TIER1 | SX_PRF_TIER1OR2);
            mapq[r] = 60;
            use_key_off[r] = 0;
            rec_off[r] = 3u * r;
            for (uint32_t i = 0; i < read_len; ++i)
            {
                sq[k * packed + (i >> 1)] |= (uint8_t)(code_of(rd.bases[i]) << ((~i & 1u) << 2));
                if (qual_bits == 4) ql[k * packed + (i >> 1)] |= (uint8_t)(rd.qcode[i] << ((~i & 1u) << 2));
                else ql[(uint64_t)k * read_len + i] = QD[rd.qcode[i]];
            }
            if (read_ascii) memcpy(read_ascii + (uint64_t)r * read_len, rd.bases, read_len);
            if (qual_wide)
                for (uint32_t i = 0; i < read_len; ++i) qual_wide[(uint64_t)r * read_len + i] = QD[rd.qcode[i]];
        }
    });
    const uint32_t n_reads(n_cells * reads_per_cell);
    region_read_off[n_cells] = n_reads;
    region_key_off[n_cells] = 3 * n_cells;
    key_ins_off[3 * n_cells] = (uint32_t)ins_base[n_cells];
    raw_seg_off[n_reads] = (uint32_t)seg_base[n_cells];
    use_key_off[n_reads] = 0;
    rec_off[n_reads] = 3u * n_reads;
    sx_region& end(regions[n_cells]);
    memset(&end, 0, sizeof(end));
    end.seq_off = (uint64_t)n_cells * seq_stride;
    end.qual_off = (uint64_t)n_cells * qual_stride;
    end.ref_off = ref_bytes & ~15ull;
    end.read_begin = n_reads;
    return 0;
}
