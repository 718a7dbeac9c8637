// oracle/score_indels_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see strelka_oracle.h).
//
// CPU restatement of the arg-max epilogue of scoreCandidateAlignments and of score_indels on the flat K6 batch
// (include/strelka_b200.h, "K6 score_indels"), written in the reference's own shape -- ordered maps and sets keyed by the
// indel -- so that it can be read side by side with
//   starling_common/starling_read_align.cpp:1295-1377, 1573-1593
//   starling_common/starling_read_align_score_indels.cpp:60-1079
//   starling_common/alignment_util.cpp:45-55, 222-302;  starling_common/indel_util.cpp:29-45
//   starling_common/IndelBuffer.cpp:76-91;  starling_common/IndelData.cpp:42-68
// A window index stands for the IndelKey it describes: a region's window is in IndelKey order, so ordered containers of indices
// iterate exactly like the reference's containers of keys.
// Parity status: PINNED by tests/test_oracle_vs_reference.py against the reference's score_indels (oracle/ref_harness_score_indels.cpp).

#include "strelka_oracle.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <utility>
#include <vector>

namespace
{

typedef std::set<uint32_t> key_set;

struct flat_alignment
{
    int32_t pos;
    bool fwd;
    const sx_aln_seg* seg;
    uint32_t n_seg;
    key_set indels; // cal.getIndels()
};

bool is_read_kind(const uint8_t k) { return k == SX_SEG_MATCH || k == SX_SEG_INSERT || k == SX_SEG_SOFTCLIP; }
bool is_ref_kind(const uint8_t k) { return k == SX_SEG_MATCH || k == SX_SEG_DELETE || k == SX_SEG_SKIP; }

// indel_util.cpp:29-45 with IndelKey::open_pos_range for the complete types
bool is_conflict(const sx_indel_key& a, const sx_indel_key& b)
{
    const bool eitherMismatch(a.type == SX_INDEL_TYPE_MISMATCH || b.type == SX_INDEL_TYPE_MISMATCH);
    const int64_t b1(a.pos), b2(b.pos);
    int64_t e1(b1 + a.del_len), e2(b2 + b.del_len);
    if (!eitherMismatch)
    {
        e1++;
        e2++;
    }
    return (e2 > b1) && (b2 < e1);
}

// starling_read_align_score_indels.cpp:131-234; false: a segment kind outside its domain
bool bp_overlap(const unsigned oligo, const flat_alignment& al, const sx_indel_key& ik, int& left, int& right)
{
    int32_t read_head(0), ref_head(al.pos);
    bool is_left(false), is_right(false);
    int32_t left_read(0), right_read(0);
    const int32_t ik_right(ik.pos + (int32_t)ik.del_len);
    for (uint32_t s = 0; s < al.n_seg; ++s)
    {
        const sx_aln_seg& ps(al.seg[s]);
        int32_t next_read(read_head), next_ref(ref_head);
        if (ps.kind == SX_SEG_MATCH)
        {
            next_read += ps.len;
            next_ref += ps.len;
        }
        else if (ps.kind == SX_SEG_INSERT) next_read += ps.len;
        else if (ps.kind == SX_SEG_DELETE) next_ref += ps.len;
        else if (ps.kind == SX_SEG_SOFTCLIP || ps.kind == SX_SEG_HARDCLIP) {}
        else return false;
        if (!is_left && ik.pos <= next_ref)
        {
            left_read = read_head + (ik.pos - ref_head);
            is_left = true;
        }
        if (!is_right && ik_right < next_ref)
        {
            right_read = read_head + (ik_right - ref_head);
            is_right = true;
        }
        read_head = next_read;
        ref_head = next_ref;
    }
    int left_ext(0), right_ext(0);
    if (al.fwd)
    {
        if (left_read > 0) left_ext = oligo;
    }
    else
    {
        if ((read_head - right_read) > 0) right_ext = oligo;
    }
    left = 0;
    if (is_left) left = std::max(0, std::min(left_read + left_ext, read_head - left_read));
    right = 0;
    if (is_right) right = std::max(0, std::min(right_read, (read_head - right_read) + right_ext));
    return true;
}

// alignment_util.cpp:45-55 (apath_insert_lead_size / _trail_size / apath_ref_length, align_path.cpp:160-345)
void soft_clip_range(const flat_alignment& al, int32_t& begin, int32_t& end)
{
    int32_t lead(0), trail(0), asize(0);
    for (uint32_t s = 0; s < al.n_seg; ++s)
    {
        const uint8_t k(al.seg[s].kind);
        if (k == SX_SEG_HARDCLIP || k == SX_SEG_SOFTCLIP) continue;
        if (k != SX_SEG_INSERT) break;
        lead += al.seg[s].len;
    }
    for (uint32_t s = al.n_seg; s-- > 0;)
    {
        const uint8_t k(al.seg[s].kind);
        if (k == SX_SEG_HARDCLIP || k == SX_SEG_SOFTCLIP) continue;
        if (k != SX_SEG_INSERT) break;
        trail += al.seg[s].len;
    }
    for (uint32_t s = 0; s < al.n_seg; ++s)
        if (is_ref_kind(al.seg[s].kind)) asize += al.seg[s].len;
    begin = al.pos - lead;
    end = al.pos + asize + trail;
}

// alignment_util.cpp:222-302
int32_t lowest_fwd_read_pos(const flat_alignment& al, const bool fwd, const int32_t range_begin, const int32_t range_end)
{
    int32_t target(fwd ? range_begin : range_end - 1);
    target -= al.pos;
    int32_t readOffset(-1);
    if (target >= 0)
    {
        int32_t ref_offset(0), read_offset(0);
        for (uint32_t s = 0; s < al.n_seg; ++s)
        {
            const sx_aln_seg& ps(al.seg[s]);
            if (is_read_kind(ps.kind)) read_offset += ps.len;
            if (!is_ref_kind(ps.kind)) continue;
            ref_offset += ps.len;
            if (ref_offset <= target) continue;
            if (!is_read_kind(ps.kind)) break;
            readOffset = read_offset - (ref_offset - target);
            break;
        }
    }
    if (readOffset < 0) return readOffset;
    if (fwd) return readOffset;
    int32_t readLength(0);
    for (uint32_t s = 0; s < al.n_seg; ++s)
        if (is_read_kind(al.seg[s].kind)) readLength += al.seg[s].len;
    return readLength - (readOffset + 1);
}

// starling_read_align.cpp:1280-1377
struct path_stats
{
    unsigned indelCount = 0, totalDeletionSize = 0, totalInsertionSize = 0, sumSegmentPos = 0;
};

path_stats stats_of(const flat_alignment& al)
{
    path_stats e;
    unsigned read_pos(0);
    for (uint32_t s = 0; s < al.n_seg; ++s)
    {
        const sx_aln_seg& ps(al.seg[s]);
        if (ps.kind != SX_SEG_MATCH) e.indelCount++;
        if (ps.kind == SX_SEG_DELETE)
        {
            e.totalDeletionSize += ps.len;
            e.sumSegmentPos += read_pos;
        }
        if (ps.kind == SX_SEG_INSERT)
        {
            e.totalInsertionSize += ps.len;
            e.sumSegmentPos += read_pos;
        }
        if (is_read_kind(ps.kind)) read_pos += ps.len;
    }
    return e;
}

unsigned candidate_count(const sx_indel_key* win, const flat_alignment& al)
{
    unsigned n(0);
    for (const uint32_t k : al.indels)
        if (win[k].flags & SX_IKF_CANDIDATE) n++;
    return n;
}

bool is_first_preferred(const sx_indel_key* win, const flat_alignment& c1, const flat_alignment& c2)
{
    const path_stats e1(stats_of(c1)), e2(stats_of(c2));
    if (e2.indelCount < e1.indelCount) return false;
    if (e2.indelCount > e1.indelCount) return true;
    const unsigned cic1(candidate_count(win, c1)), cic2(candidate_count(win, c2));
    if (cic2 > cic1) return false;
    if (cic2 < cic1) return true;
    if (e2.totalInsertionSize < e1.totalInsertionSize) return false;
    if (e2.totalInsertionSize > e1.totalInsertionSize) return true;
    if (e2.totalDeletionSize < e1.totalDeletionSize) return false;
    if (e2.totalDeletionSize > e1.totalDeletionSize) return true;
    return (e2.sumSegmentPos >= e1.sumSegmentPos);
}

// score_indels.cpp:247-276
bool is_equiv(const sx_indel_key* win, const flat_alignment& c1, const flat_alignment& c2, std::set<std::pair<uint32_t, uint32_t>>& pairs)
{
    pairs.clear();
    if (c1.indels.size() != c2.indels.size()) return false;
    key_set::const_iterator i1(c1.indels.begin()), i2(c2.indels.begin());
    for (; i1 != c1.indels.end(); ++i1, ++i2)
    {
        if (*i1 == *i2) continue;
        const sx_indel_key &a(win[*i1]), &b(win[*i2]);
        if (a.type != b.type) return false;
        if (a.del_len != b.del_len) return false;
        if (a.ins_len != b.ins_len || a.ins_id != b.ins_id) return false;
        pairs.insert(std::make_pair(*i1, *i2));
    }
    return true;
}

// score_indels.cpp:285-300
bool is_first_dominant(const sx_indel_key* win, const uint32_t k1, const uint32_t k2)
{
    const bool ic1(win[k1].flags & SX_IKF_CANDIDATE), ic2(win[k2].flags & SX_IKF_CANDIDATE);
    if (ic2 && !ic1) return false;
    if (ic2 == ic1) return (win[k1].pos <= win[k2].pos);
    return true;
}

// key of iks_map_t (:42-54): (active indel, (is X present, X))
typedef std::pair<uint32_t, std::pair<bool, uint32_t>> status_key;
typedef std::map<status_key, double> status_map;

void tick(status_map& m, const uint32_t call, const bool present, const uint32_t x, const double lnp)
{
    const status_key k(call, std::make_pair(present, x));
    const status_map::const_iterator j(m.find(k));
    if (j != m.end() && j->second >= lnp) return;
    m[k] = lnp;
}

} // namespace

extern "C" void ox_default_score_indels_opts(sx_score_indels_opts* o)
{
    o->max_indel_size = 49;
    o->upstream_oligo_size = 0;
    o->min_read_bp_flank = 5;
    o->is_smoothed_alignments = 1;
    o->smoothed_lnp_range = 2.302585092994046; // std::log(10.)
}

extern "C" int ox_score_indels(const sx_score_indels_batch* b, const double* lnp, sx_read_indel_score* recs, uint32_t* n_rec, uint32_t* max_aln, uint32_t* eval_aln)
{
    const sx_score_indels_opts& opt(b->opts);
    for (uint32_t region = 0; region < b->n_regions; ++region)
    {
        const sx_indel_key* win(b->keys + b->region_key_off[region]);
        const uint32_t n_win(b->region_key_off[region + 1] - b->region_key_off[region]);
        if (n_win > 65535) return SX_ERR_RANGE;
        for (uint32_t k = 0; k < n_win; ++k)
            if (win[k].type > SX_INDEL_TYPE_MISMATCH) return SX_ERR_UNSUPPORTED;
        for (uint32_t r = b->region_read_off[region]; r < b->region_read_off[region + 1]; ++r)
        {
            n_rec[r] = 0;
            max_aln[r] = eval_aln[r] = UINT32_MAX;
            const uint32_t a0(b->aln_off[r]), n_cal(b->aln_off[r + 1] - a0);
            if (n_cal == 0) continue;
            const uint8_t rflags(b->read_flags[r]);
            const bool fwd(rflags & SX_SIF_FWD), is_tier1(rflags & SX_SIF_TIER1), is_incomplete(rflags & SX_SIF_INCOMPLETE);
            std::vector<flat_alignment> cals(n_cal);
            std::vector<double> scores(n_cal);
            for (uint32_t c = 0; c < n_cal; ++c)
            {
                flat_alignment& al(cals[c]);
                al.pos = b->aln_pos[a0 + c];
                al.fwd = fwd;
                al.seg = b->segs + b->aln_seg_off[a0 + c];
                al.n_seg = b->aln_seg_off[a0 + c + 1] - b->aln_seg_off[a0 + c];
                for (uint32_t s = 0; s < al.n_seg; ++s)
                    if (al.seg[s].kind == SX_SEG_SKIP || al.seg[s].kind == SX_SEG_REFSKIP || al.seg[s].kind > SX_SEG_SKIP) return SX_ERR_UNSUPPORTED;
                for (uint32_t i = b->aln_key_off[a0 + c]; i < b->aln_key_off[a0 + c + 1]; ++i)
                {
                    if (b->aln_keys[i] >= n_win) return SX_ERR_ARG;
                    al.indels.insert(b->aln_keys[i]);
                }
                scores[c] = lnp[a0 + c];
            }

            // ---- scoreCandidateAlignments, starling_read_align.cpp:1573-1593
            double maxScore(0);
            int maxCal(-1);
            for (uint32_t c = 0; c < n_cal; ++c)
            {
                const double path_lnp(scores[c]);
                if (maxCal >= 0)
                {
                    if (path_lnp < maxScore) continue;
                    if ((path_lnp <= maxScore) && is_first_preferred(win, cals[maxCal], cals[c])) continue;
                }
                maxScore = path_lnp;
                maxCal = (int)c;
            }
            max_aln[r] = a0 + maxCal;

            // ---- late_indel_normalization_filter, :281-450.  Its nonnorm_indels argument is taken by value (:308), so the set
            // score_indels owns stays empty; what survives the call is isFilter[] and the re-chosen maximum.
            std::vector<bool> isFilter(n_cal, false);
            {
                const double equiv_range(opt.is_smoothed_alignments ? opt.smoothed_lnp_range : 0.);
                std::vector<std::pair<double, unsigned>> sorted;
                for (uint32_t c = 0; c < n_cal; ++c) sorted.push_back(std::make_pair(scores[c], c));
                std::sort(sorted.rbegin(), sorted.rend());
                std::vector<double> smooth(scores);
                bool any_excluded(false);
                std::set<std::pair<uint32_t, uint32_t>> pairs;
                for (uint32_t i1 = 0; i1 < n_cal; ++i1)
                {
                    const unsigned s1(sorted[i1].second);
                    if (isFilter[s1]) continue;
                    for (uint32_t i2 = i1 + 1; i2 < n_cal; ++i2)
                    {
                        const unsigned s2(sorted[i2].second);
                        if (isFilter[s2]) continue;
                        if (smooth[s2] + equiv_range < smooth[s1]) break;
                        if (!is_equiv(win, cals[s1], cals[s2], pairs)) continue;
                        if (pairs.empty()) continue;
                        bool s1_removed(false), removed(false);
                        for (const auto& pr : pairs)
                        {
                            if (is_first_dominant(win, pr.first, pr.second))
                            {
                                if (!removed)
                                {
                                    isFilter[s2] = true;
                                    any_excluded = true;
                                    smooth[s1] = std::max(smooth[s1], smooth[s2]);
                                }
                            }
                            else
                            {
                                if (!removed)
                                {
                                    isFilter[s1] = true;
                                    any_excluded = true;
                                    smooth[s2] = std::max(smooth[s1], smooth[s2]);
                                    s1_removed = true;
                                }
                            }
                            removed = true;
                        }
                        if (s1_removed) break;
                    }
                }
                if (any_excluded)
                {
                    for (uint32_t i = 0; i < n_cal; ++i)
                    {
                        const unsigned s(sorted[i].second);
                        if (isFilter[s]) continue;
                        maxScore = scores[s];
                        maxCal = (int)s;
                        break;
                    }
                }
            }
            eval_aln[r] = a0 + maxCal;
            const flat_alignment& maxAl(cals[maxCal]);

            sx_read_indel_score* out(recs + b->rec_off[r]);
            const uint32_t out_cap(b->rec_off[r + 1] - b->rec_off[r]);
            uint32_t n_out(0);
            std::vector<sx_read_indel_score> sub_recs; // records are emitted in key order: merged at the end

            // ---- (2a) the indels this read evaluates, :520-656
            key_set toEvaluate;
            {
                int32_t rb, re;
                soft_clip_range(maxAl, rb, re);
                // IndelBuffer::rangeIterator(rb, re): [lower_bound(IndelKey(rb - maxIndelSize)), lower_bound(IndelKey(re))), then skip
                // the leading entries whose right_pos() < rb  (an IndelKey of type NONE sorts before every entry at its position)
                uint32_t k_end(0);
                while (k_end < n_win && win[k_end].pos < re) k_end++;
                uint32_t k(0);
                while (k < k_end && (int64_t)win[k].pos < (int64_t)rb - (int64_t)opt.max_indel_size) k++;
                for (; k != k_end; ++k)
                    if (win[k].pos + (int32_t)win[k].del_len >= rb) break;
                for (; k != k_end; ++k)
                {
                    const sx_indel_key& ik(win[k]);
                    if (ik.type == SX_INDEL_TYPE_MISMATCH) continue;
                    if (!(ik.flags & SX_IKF_CANDIDATE)) continue;
                    const bool inMax(maxAl.indels.count(k) != 0);
                    int best(-1);
                    if (inMax) best = maxCal;
                    else
                    {
                        double bestScore(0);
                        for (uint32_t c = 0; c < n_cal; ++c)
                        {
                            if ((int)c == maxCal) continue;
                            if (isFilter[c]) continue;
                            if (cals[c].indels.count(k) == 0) continue;
                            if (best < 0 || scores[c] > bestScore)
                            {
                                bestScore = scores[c];
                                best = (int)c;
                            }
                        }
                    }
                    if (best < 0) continue;
                    int lo, ro;
                    if (!bp_overlap(opt.upstream_oligo_size, cals[best], ik, lo, ro)) return SX_ERR_UNSUPPORTED;
                    const int bpo(std::max(lo, ro));
                    if (bpo < opt.min_read_bp_flank)
                    {
                        if (bpo > 0)
                        {
                            sx_read_indel_score rec;
                            std::memset(&rec, 0, sizeof(rec));
                            rec.key = (uint16_t)k;
                            rec.flags = SX_RIS_SUBOVERLAP;
                            sub_recs.push_back(rec);
                        }
                        continue;
                    }
                    toEvaluate.insert(k);
                }
            }

            // ---- which evaluated indels conflict with each other, :665-686
            std::map<uint32_t, key_set> orthogonal;
            for (key_set::const_iterator i(toEvaluate.begin()); i != toEvaluate.end(); ++i)
            {
                key_set::const_iterator j(i);
                for (++j; j != toEvaluate.end(); ++j)
                    if (is_conflict(win[*i], win[*j]))
                    {
                        orthogonal[*i].insert(*j);
                        orthogonal[*j].insert(*i);
                    }
            }

            // ---- (2b) best score of every (indel, state), :688-849
            status_map info;
            for (uint32_t c = 0; c < n_cal; ++c)
            {
                if (isFilter[c]) continue;
                const double score(scores[c]);
                const key_set& inCal(cals[c].indels);
                key_set nonCandidateOrthogonal;
                for (const uint32_t e : toEvaluate)
                {
                    const sx_indel_key& ek(win[e]);
                    if (inCal.count(e) != 0)
                    {
                        tick(info, e, true, e, score);
                        tick(info, e, false, e, score + ek.ref_to_indel_lnp);
                        for (const uint32_t o : orthogonal[e])
                        {
                            tick(info, o, false, o, score + ek.ref_to_indel_lnp);
                            tick(info, o, true, e, score);
                        }
                    }
                    else
                    {
                        // which_interfering_indel, :100-118
                        int interfering(-1);
                        for (const uint32_t cur : inCal)
                        {
                            if (win[cur].type == SX_INDEL_TYPE_MISMATCH) continue;
                            if (is_conflict(win[cur], ek))
                            {
                                interfering = (int)cur;
                                break;
                            }
                        }
                        if (interfering >= 0 && toEvaluate.count((uint32_t)interfering) == 0) nonCandidateOrthogonal.insert((uint32_t)interfering);
                        if (interfering < 0)
                        {
                            tick(info, e, false, e, score);
                            tick(info, e, true, e, score + ek.indel_to_ref_lnp);
                        }
                        else tick(info, e, true, e, score + ek.indel_to_ref_lnp);
                    }
                }
                for (const uint32_t nc : nonCandidateOrthogonal)
                    for (const uint32_t e : toEvaluate)
                    {
                        if (!is_conflict(win[nc], win[e])) continue;
                        tick(info, e, false, e, score + win[nc].ref_to_indel_lnp);
                    }
            }

            // ---- (3) one ReadPathScores per evaluated indel, :852-1075
            const unsigned read_length(b->read_len[r]);
            const unsigned fullReadLength(b->full_len ? b->full_len[r] : read_length);
            const unsigned fullReadOffset(b->full_off ? b->full_off[r] : 0);
            std::vector<sx_read_indel_score> scored;
            for (const uint32_t e : toEvaluate)
            {
                const sx_indel_key& ek(win[e]);
                const bool inMax(maxAl.indels.count(e) != 0);
                double indelScore(maxScore);
                if (!inMax)
                {
                    const status_map::const_iterator it(info.find(status_key(e, std::make_pair(true, e))));
                    if (it == info.end()) continue; // incomplete search, or the reference's safe-mode warning: both skip the indel
                    indelScore = it->second;
                }
                double refScore(0);
                {
                    const status_map::const_iterator it(info.find(status_key(e, std::make_pair(false, e))));
                    if (it == info.end()) continue;
                    refScore = it->second;
                }
                (void)is_incomplete; // both outcomes of the is_incomplete_search tests skip the indel (:905-930, :945-968)
                const int32_t right_pos(ek.pos + (int32_t)ek.del_len);
                const int32_t readPos(lowest_fwd_read_pos(maxAl, fwd, ek.pos - 1, right_pos + 1));
                int32_t dist((int32_t)fullReadLength);
                {
                    const int32_t revReadPos(lowest_fwd_read_pos(maxAl, !fwd, ek.pos - 1, right_pos + 1));
                    if (readPos >= 0) dist = readPos + (int32_t)fullReadOffset;
                    if (revReadPos >= 0)
                    {
                        const int32_t fullRev(revReadPos + (int32_t)(fullReadLength - (fullReadOffset + read_length)));
                        if (fullRev < dist) dist = fullRev;
                    }
                }
                sx_read_indel_score rec;
                std::memset(&rec, 0, sizeof(rec));
                rec.key = (uint16_t)e;
                rec.flags = SX_RIS_SCORED;
                rec.ref_lnp = (float)refScore;
                rec.indel_lnp = (float)indelScore;
                rec.read_pos = (int16_t)readPos;
                rec.dist_from_edge = (int16_t)dist;
                // alternate alleles, :1022-1062 with ReadPathScores::insertAlt
                std::vector<std::pair<uint32_t, float>> alt;
                for (const uint32_t o : orthogonal[e])
                {
                    const status_map::const_iterator it(info.find(status_key(e, std::make_pair(true, o))));
                    if (it == info.end()) continue;
                    const float a((float)it->second);
                    if (alt.size() < 2) alt.push_back(std::make_pair(o, a));
                    else
                    {
                        unsigned min_index(alt.size());
                        float mn(a);
                        for (unsigned i = 0; i < alt.size(); ++i)
                            if (alt[i].second < mn)
                            {
                                mn = alt[i].second;
                                min_index = i;
                            }
                        if (min_index < alt.size()) alt[min_index] = std::make_pair(o, a);
                    }
                }
                rec.n_alt = (uint8_t)alt.size();
                for (unsigned i = 0; i < alt.size(); ++i)
                {
                    rec.alt_key[i] = (uint16_t)alt[i].first;
                    rec.alt_lnp[i] = alt[i].second;
                }
                scored.push_back(rec);
            }
            (void)is_tier1; // the tier of a record is the read's (SX_SIF_TIER1)

            // records of a read in key order (a key is either scored or a suboverlap mark, never both)
            std::vector<sx_read_indel_score> all(scored);
            all.insert(all.end(), sub_recs.begin(), sub_recs.end());
            std::sort(all.begin(), all.end(), [](const sx_read_indel_score& x, const sx_read_indel_score& y) { return x.key < y.key; });
            if (all.size() > out_cap) return SX_ERR_NOMEM;
            for (const auto& rec : all) out[n_out++] = rec;
            n_rec[r] = n_out;
        }
    }
    return SX_OK;
}
