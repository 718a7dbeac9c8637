// realign_oracle.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/liboracle.so).  CPU restatement of what K9 choose_realignment
// (include/strelka_b200.h) replaces, in the reference's own shape (a pool vector, a per-read-position map, a rebuilt alignment):
//   scoreCandidateAlignments, tail          starling_common/starling_read_align.cpp:1573-1741 (unpinned reads, no soft-clip retention test)
//   isFirstCandidateAlignmentPreferred      :1352-1377 with getExtraPathInfo :1293-1318, getCandidateIndelCount :1322-1334
//   finishRealignment                       :1411-1450
//   getClippedAlignmentFromTopAlignmentPool starling_common/starling_read_align_clipper.cpp:340-424
//   get_alignment_ref_map :96-146, mark_ref_map_conflicts :150-225, soft_clip_alignment :255-338, extend_or_add_sc :229-243
// Parity status: PINNED -- tests/test_oracle_vs_reference.py compares it with the reference's own scoreCandidateAlignments driven on rebuilt
// objects (oracle/ref_harness_enumerate.inc: ref_choose_realignment), and tests/golden/realign_ref.npz freezes the reference's output.
// Only tests/, smoke() and bench.py's CPU arm may call this.
#include "../include/strelka_b200.h"

#include <algorithm>
#include <stdexcept>
#include <vector>

namespace
{
struct Seg
{
    int type;
    unsigned length;
};
struct Aln
{
    int pos = 0;
    std::vector<Seg> path;
    bool empty() const { return path.empty(); }
};

bool isAlignMatch(int t) { return t == SX_AP_MATCH || t == SX_AP_SEQ_MATCH || t == SX_AP_SEQ_MISMATCH; }
bool isReadLength(int t) { return isAlignMatch(t) || t == SX_AP_INSERT || t == SX_AP_SOFT_CLIP; }

struct PathInfo
{
    unsigned indelCount = 0, totalDeletionSize = 0, totalInsertionSize = 0, sumSegmentPos = 0;
};

PathInfo pathInfo(const std::vector<Seg>& p)
{
    PathInfo e;
    unsigned read_pos(0);
    for (const Seg& s : p)
    {
        if (!isAlignMatch(s.type)) e.indelCount++;
        if (s.type == SX_AP_DELETE || s.type == SX_AP_INSERT)
        {
            (s.type == SX_AP_DELETE ? e.totalDeletionSize : e.totalInsertionSize) += s.length;
            e.sumSegmentPos += read_pos;
        }
        if (isReadLength(s.type)) read_pos += s.length;
    }
    return e;
}

struct Cand
{
    Aln al;
    unsigned candidateIndels;
    double score;
};

bool firstPreferred(const Cand& c1, const Cand& c2)
{
    const PathInfo e1(pathInfo(c1.al.path)), e2(pathInfo(c2.al.path));
    if (e2.indelCount < e1.indelCount) return false;
    if (e2.indelCount > e1.indelCount) return true;
    if (c2.candidateIndels > c1.candidateIndels) return false;
    if (c2.candidateIndels < c1.candidateIndels) return true;
    if (e2.totalInsertionSize < e1.totalInsertionSize) return false;
    if (e2.totalInsertionSize > e1.totalInsertionSize) return true;
    if (e2.totalDeletionSize < e1.totalDeletionSize) return false;
    if (e2.totalDeletionSize > e1.totalDeletionSize) return true;
    return e2.sumSegmentPos >= e1.sumSegmentPos;
}

enum MapType { NONE, MATCH, INSERT, SOFT_CLIP, CONFLICT };
struct RefMap
{
    MapType type;
    int pos;
};

struct BadPath
{
};

void refMapOf(const Aln& al, std::vector<RefMap>& m)
{
    m.clear();
    int ref_head(al.pos);
    for (const Seg& s : al.path)
    {
        if (isAlignMatch(s.type))
        {
            for (unsigned j = 0; j < s.length; ++j) m.push_back(RefMap{MATCH, ref_head + (int)j});
            ref_head += (int)s.length;
        }
        else if (s.type == SX_AP_INSERT) m.insert(m.end(), s.length, RefMap{INSERT, 0});
        else if (s.type == SX_AP_DELETE || s.type == SX_AP_SKIP) ref_head += (int)s.length;
        else if (s.type == SX_AP_SOFT_CLIP) m.insert(m.end(), s.length, RefMap{SOFT_CLIP, 0});
        else if (s.type != SX_AP_HARD_CLIP) throw BadPath();
    }
}

void markConflicts(const Aln& al, std::vector<RefMap>& m)
{
    int ref_head(al.pos);
    size_t read_head(0);
    for (const Seg& s : al.path)
    {
        if (isAlignMatch(s.type) || s.type == SX_AP_INSERT || s.type == SX_AP_SOFT_CLIP)
        {
            const MapType want(isAlignMatch(s.type) ? MATCH : (s.type == SX_AP_INSERT ? INSERT : SOFT_CLIP));
            for (unsigned j = 0; j < s.length; ++j)
            {
                if (read_head + j >= m.size()) throw BadPath();
                RefMap& rm(m[read_head + j]);
                if (rm.type == CONFLICT) continue;
                if (rm.type != want || (want == MATCH && rm.pos != ref_head + (int)j)) rm.type = CONFLICT;
            }
            read_head += s.length;
            if (isAlignMatch(s.type)) ref_head += (int)s.length;
        }
        else if (s.type == SX_AP_DELETE || s.type == SX_AP_SKIP) ref_head += (int)s.length;
        else if (s.type != SX_AP_HARD_CLIP) throw BadPath();
    }
}

void extendOrAddSoftClip(Aln& al, unsigned length)
{
    if (!al.path.empty() && al.path.back().type == SX_AP_SOFT_CLIP) al.path.back().length += length;
    else al.path.push_back(Seg{SX_AP_SOFT_CLIP, length});
}

void softClipAlignment(Aln& al, unsigned leading, unsigned trailing)
{
    unsigned read_head(0);
    Aln out;
    out.pos = al.pos;
    for (const Seg& s : al.path)
    {
        if (isAlignMatch(s.type) || s.type == SX_AP_INSERT)
        {
            if (leading > read_head)
            {
                const unsigned clip(std::min(s.length, leading - read_head));
                extendOrAddSoftClip(out, clip);
                if (isAlignMatch(s.type)) out.pos += (int)clip;
                if (clip < s.length) out.path.push_back(Seg{s.type, s.length - clip});
            }
            else if (trailing < read_head + s.length)
            {
                const unsigned clip(std::min(s.length, read_head + s.length - trailing));
                if (clip < s.length) out.path.push_back(Seg{s.type, s.length - clip});
                extendOrAddSoftClip(out, clip);
            }
            else out.path.push_back(s);
            read_head += s.length;
        }
        else if (s.type == SX_AP_DELETE || s.type == SX_AP_SKIP)
        {
            if (leading >= read_head) out.pos += (int)s.length;
            else if (trailing <= read_head)
            {
            }
            else out.path.push_back(s);
        }
        else if (s.type == SX_AP_SOFT_CLIP)
        {
            extendOrAddSoftClip(out, s.length);
            read_head += s.length;
        }
        else if (s.type == SX_AP_HARD_CLIP) out.path.push_back(s);
        else throw BadPath();
    }
    al = out;
}

Aln clippedFromPool(const std::vector<const Cand*>& pool, size_t best)
{
    Aln clipped(pool[best]->al);
    if (pool.size() == 1) return clipped;
    std::vector<RefMap> m;
    refMapOf(clipped, m);
    for (size_t i = 0; i < pool.size(); ++i)
        if (i != best) markConflicts(pool[i]->al, m);
    const unsigned read_size(m.size());
    unsigned leading(0), trailing(read_size);
    while (leading < read_size && m[leading].type != MATCH) ++leading;
    while (leading > 0 && m[leading - 1].type != CONFLICT && m[leading - 1].type != SOFT_CLIP) --leading;
    while (trailing > 0 && m[trailing - 1].type != MATCH) --trailing;
    while (trailing < read_size && m[trailing].type != CONFLICT && m[trailing].type != SOFT_CLIP) ++trailing;
    if (leading >= trailing) return Aln(); // "clear"
    if (leading != 0 || trailing != read_size) softClipAlignment(clipped, leading, trailing);
    return clipped;
}

uint8_t outKind(const sx_realign_batch& b, int t)
{
    if (!b.k4_kinds) return (uint8_t)t;
    if (isAlignMatch(t)) return SX_SEG_MATCH;
    switch (t)
    {
    case SX_AP_INSERT: return SX_SEG_INSERT;
    case SX_AP_DELETE: return SX_SEG_DELETE;
    case SX_AP_SKIP: return SX_SEG_SKIP;
    case SX_AP_SOFT_CLIP: return SX_SEG_SOFTCLIP;
    default: return SX_SEG_HARDCLIP;
    }
}
} // namespace

extern "C" int ox_choose_realignment(const sx_realign_batch* b, const double* lnp, sx_realign_out* o)
{
    uint32_t total(0);
    for (uint32_t r = 0; r < b->n_reads; ++r)
    {
        o->seg_off[r] = total;
        uint32_t longest(0);
        for (uint32_t a = b->aln_off[r]; a < b->aln_off[r + 1]; ++a) longest = std::max(longest, b->aln_seg_off[a + 1] - b->aln_seg_off[a]);
        uint32_t need(longest ? longest + 2 : 0);
        if (b->raw_seg_off) need = std::max(need, b->raw_seg_off[r + 1] - b->raw_seg_off[r]);
        total += need;
    }
    o->seg_off[b->n_reads] = total;
    o->totals[0] = total;
    if (total > o->cap_segs) return SX_ERR_CAPACITY;
    for (uint32_t g = 0; g < b->n_regions; ++g)
        for (uint32_t r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r)
        {
            const uint32_t s0(o->seg_off[r]), s1(o->seg_off[r + 1]);
            for (uint32_t i = s0; i < s1; ++i) o->segs[i] = sx_aln_seg{0, outKind(*b, SX_AP_HARD_CLIP), 0};
            o->pos[r] = 0;
            o->n_seg[r] = 0;
            o->status[r] = 0;
            o->best_aln[r] = UINT32_MAX;
            if (b->raw_seg_off) // getBestAlignment of a read that keeps the mapper's alignment (starling_read_segment.hh:134-138); overwritten below by a realignment
            {
                const uint32_t q0(b->raw_seg_off[r]), nq(b->raw_seg_off[r + 1] - q0);
                for (uint32_t i = 0; i < nq; ++i) o->segs[s0 + i] = sx_aln_seg{b->raw_segs[q0 + i].len, outKind(*b, b->raw_segs[q0 + i].kind), 0};
                o->pos[r] = b->raw_pos[r];
                o->n_seg[r] = (uint16_t)nq;
            }
            const uint32_t a0(b->aln_off[r]), a1(b->aln_off[r + 1]);
            if (a0 == a1) continue;
            if (b->pin_flags && b->pin_flags[r])
            {
                o->status[r] = SX_REALIGN_ST_UNSUPPORTED;
                continue;
            }
            std::vector<Cand> cands(a1 - a0);
            for (uint32_t a = a0; a < a1; ++a)
            {
                Cand& c(cands[a - a0]);
                c.al.pos = b->aln_pos[a];
                for (uint32_t s = b->aln_seg_off[a]; s < b->aln_seg_off[a + 1]; ++s) c.al.path.push_back(Seg{b->segs[s].kind, b->segs[s].len});
                c.candidateIndels = 0;
                for (uint32_t q = b->aln_key_off[a]; q < b->aln_key_off[a + 1]; ++q)
                    c.candidateIndels += (b->keys[b->region_key_off[g] + b->aln_keys[q]].flags & SX_IKF_CANDIDATE) ? 1u : 0u;
                c.score = lnp[a];
            }
            // the maximum, :1573-1593
            const Cand* maxPtr(nullptr);
            double maxScore(0);
            for (const Cand& c : cands)
            {
                if (maxPtr)
                {
                    if (c.score < maxScore) continue;
                    if (c.score <= maxScore && firstPreferred(*maxPtr, c)) continue;
                }
                maxScore = c.score;
                maxPtr = &c;
            }
            // the smooth pool, :1659-1683
            const double range(b->is_smoothed_alignments ? b->smoothed_lnp_range : 0.);
            std::vector<const Cand*> pool;
            const Cand* smooth(nullptr);
            for (const Cand& c : cands)
            {
                if (c.score + range < maxScore) continue;
                pool.push_back(&c);
                if (!smooth || !firstPreferred(*smooth, c)) smooth = &c;
            }
            try
            {
                // finishRealignment, :1411-1450
                Aln realignment(smooth->al);
                if (pool.size() > 1)
                {
                    const size_t best(std::find(pool.begin(), pool.end(), smooth) - pool.begin());
                    realignment = clippedFromPool(pool, best);
                    if (realignment.empty()) realignment = smooth->al;
                }
                if (realignment.path.size() > s1 - s0) throw BadPath();
                for (size_t i = 0; i < realignment.path.size(); ++i)
                    o->segs[s0 + i] = sx_aln_seg{(uint16_t)realignment.path[i].length, outKind(*b, realignment.path[i].type), 0};
                for (uint32_t i = s0 + (uint32_t)realignment.path.size(); i < s1; ++i) o->segs[i] = sx_aln_seg{0, outKind(*b, SX_AP_HARD_CLIP), 0};
                o->pos[r] = realignment.pos;
                o->n_seg[r] = (uint16_t)realignment.path.size();
                o->status[r] = SX_REALIGN_ST_REALIGNED;
                o->best_aln[r] = a0 + (uint32_t)(smooth - cands.data());
            }
            catch (const BadPath&)
            {
                o->status[r] = SX_REALIGN_ST_BADPATH;
            }
        }
    return 0;
}
