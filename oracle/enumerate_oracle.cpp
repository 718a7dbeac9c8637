// enumerate_oracle.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/liboracle.so).  CPU restatement of the reference's candidate
// alignment enumeration for K7 (include/strelka_b200.h "K7 enumerate_alignments"):
//   getCandidateAlignments          starling_common/starling_read_align.cpp:1816-1994
//   candidate_alignment_search      :857-1277
//   make_start_pos_alignment        :393-584      get_end_pin_start_pos   :593-719
//   add_indels_in_range             :311-375      sort_remove_only_indels_last :724-749
//   addKeysToCandidateAlignment     :789-804      getCurIndelHaplotypeIds :811-849
//   HaplotypeStatus                 :56-179
// Unlike the device body (strelka_b200/csrc/k7_core.cuh: explicit frame stack, bit masks, sorted index array) this restatement
// keeps the reference's own shape -- a recursion whose arguments are ordered containers passed by value, results collected in a
// std::set -- with one substitution: an IndelKey is the index of its entry in the region's window, which is in IndelKey order, so
// integer order is key order.  Parity status: PINNED -- tests/test_oracle_vs_reference.py compares it with the reference's own
// getCandidateAlignments (oracle/_ref, oracle/ref_harness_enumerate.inc) alignment by alignment, and tests/golden/enumerate_ref.npz
// freezes the reference's output.  Only tests/, smoke() and bench.py's CPU arm may call this.
#include "../include/strelka_b200.h"

#include <algorithm>
#include <map>
#include <set>
#include <stdexcept>
#include <tuple>
#include <utility>
#include <vector>

namespace
{
typedef int Key; // window index; NONE = -1 (IndelKey() sorts before every real key: position 0, type NONE)

struct Seg
{
    int type; // SX_AP_* == ALIGNPATH::align_t
    unsigned length;
    bool operator<(const Seg& o) const { return std::tie(type, length) < std::tie(o.type, o.length); } // align_path.hh:190-196
    bool operator==(const Seg& o) const { return type == o.type && length == o.length; }
};
typedef std::vector<Seg> Path;

struct Cal // CandidateAlignment.hh:36-83 (every alignment of a read has the read's strand)
{
    int pos = 0;
    Path path;
    std::set<Key> indels;
    Key leading = -1, trailing = -1;
    bool operator<(const Cal& o) const
    {
        if (pos != o.pos) return pos < o.pos;
        if (path.size() != o.path.size()) return path.size() < o.path.size(); // alignment.hh:80-83: size first, then elementwise
        if (path != o.path) return path < o.path;
        if (indels != o.indels) return indels < o.indels;
        if (leading != o.leading) return leading < o.leading;
        return trailing < o.trailing;
    }
};

struct Info // starling_align_indel_info, :48-53
{
    bool is_present = false, is_remove_only = false, isInOriginalAlignment = false;
};
typedef std::map<Key, Info> StatusMap;

struct HapStatus // :134-179
{
    std::vector<int> constraints;
    bool isAnyIndelOn = false;
    explicit HapStatus(unsigned n) : constraints(n, 3) {}
};
typedef std::map<int, HapStatus> HapMap;

struct Thrown // the reference throws blt_exception (or trips an assert) here
{
};

struct LimitHit // a per-read capacity of the device build (k7_core.cuh) would be exceeded here; only raised when `limits` is on
{
};
bool g_limits = false;   // (test infrastructure: single-threaded)
unsigned g_maxA = 64, g_nClipSegs = 0;

struct Warn
{
    bool origin_skip = false, max_toggle_depth = false;
};

struct Ctx
{
    const sx_indel_key* win;
    const sx_key_hap* hap;
    unsigned n_win;
    std::set<Key> usable; // non-candidate entries this read is an observation of
    int realign_begin, realign_end;
    const sx_enum_opts* opt;
};

bool isMismatch(const sx_indel_key& k) { return k.type == SX_INDEL_TYPE_MISMATCH; }
int rightPos(const sx_indel_key& k) { return k.pos + (int)k.del_len; }
bool isPrimitiveDeletion(const sx_indel_key& k) { return k.type == SX_INDEL_TYPE_INDEL && k.ins_len == 0 && k.del_len > 0; }

bool refLen(int t) { return t == SX_AP_MATCH || t == SX_AP_DELETE || t == SX_AP_SKIP || t == SX_AP_SEQ_MATCH || t == SX_AP_SEQ_MISMATCH; }
bool readLen(int t) { return t == SX_AP_MATCH || t == SX_AP_INSERT || t == SX_AP_SOFT_CLIP || t == SX_AP_SEQ_MATCH || t == SX_AP_SEQ_MISMATCH; }
bool unalignedEdge(int t) { return t == SX_AP_INSERT || t == SX_AP_HARD_CLIP || t == SX_AP_SOFT_CLIP; }
bool isClip(int t) { return t == SX_AP_SOFT_CLIP || t == SX_AP_HARD_CLIP; }

unsigned pathRefLength(const Path& p)
{
    unsigned v(0);
    for (const Seg& s : p)
        if (refLen(s.type)) v += s.length;
    return v;
}

unsigned unalignedPrefix(const Path& p)
{
    unsigned v(0);
    for (const Seg& s : p)
    {
        if (!unalignedEdge(s.type)) break;
        if (readLen(s.type)) v += s.length;
    }
    return v;
}

unsigned unalignedSuffix(const Path& p)
{
    unsigned v(0);
    for (auto it = p.rbegin(); it != p.rend(); ++it)
    {
        if (!unalignedEdge(it->type)) break;
        if (readLen(it->type)) v += it->length;
    }
    return v;
}

// get_soft_clip_alignment_range, alignment_util.cpp:45-55
std::pair<int, int> softClipRange(const Cal& c)
{
    unsigned lead(0), trail(0);
    for (const Seg& s : c.path)
    {
        if (isClip(s.type)) continue;
        if (s.type != SX_AP_INSERT) break;
        lead += s.length;
    }
    for (auto it = c.path.rbegin(); it != c.path.rend(); ++it)
    {
        if (isClip(it->type)) continue;
        if (it->type != SX_AP_INSERT) break;
        trail += it->length;
    }
    return std::make_pair(c.pos - (int)lead, c.pos + (int)pathRefLength(c.path) + (int)trail);
}

// indel_util.cpp:29-76
bool conflict(const sx_indel_key& a, const sx_indel_key& b)
{
    const long margin((isMismatch(a) || isMismatch(b)) ? 0 : 1);
    return ((long)b.pos + b.del_len + margin > (long)a.pos) && ((long)b.pos < (long)a.pos + a.del_len + margin);
}
bool openIntersect(int b, int e, int p) { return p > b && p < e; } // known_pos_range(b,e).is_range_intersect(pos_range(p,p))
bool bpIntersect(int b, int e, const sx_indel_key& k)
{
    if (isMismatch(k)) return k.pos >= b && k.pos < e;
    if (openIntersect(b, e, k.pos)) return true;
    return rightPos(k) != k.pos && openIntersect(b, e, rightPos(k));
}
bool closedAdjacent(int b, int e, int p) { return p + 1 > b && p - 1 < e; }
bool bpAdjacent(int b, int e, const sx_indel_key& k)
{
    if (closedAdjacent(b, e, k.pos)) return true;
    return rightPos(k) != k.pos && closedAdjacent(b, e, rightPos(k));
}

// :311-375 over IndelBuffer::rangeIterator (IndelBuffer.cpp:76-91)
void addIndelsInRange(const Ctx& C, int b, int e, StatusMap& status, std::vector<Key>& order)
{
    unsigned k(0);
    while (k < C.n_win && (long)C.win[k].pos < (long)b - (long)C.opt->max_indel_size) ++k;
    while (k < C.n_win && C.win[k].pos < e && rightPos(C.win[k]) < b) ++k;
    for (; k < C.n_win && C.win[k].pos < e; ++k)
    {
        const sx_indel_key& ik(C.win[k]);
        if (ik.type > SX_INDEL_TYPE_MISMATCH) throw std::runtime_error("breakend window entries are not part of this build");
        if (!bpAdjacent(b, e, ik)) continue;
        const bool removeOnly(!bpIntersect(b, e, ik));
        auto it(status.find((Key)k));
        if (it != status.end())
        {
            if (!removeOnly) it->second.is_remove_only = false;
        }
        else if ((ik.flags & SX_IKF_CANDIDATE) || C.usable.count((Key)k))
        {
            if (g_limits && status.size() >= 64) throw LimitHit();
            Info info;
            info.is_remove_only = removeOnly;
            status[(Key)k] = info;
            order.push_back((Key)k);
        }
    }
}

// :724-749
void removeOnlyLast(const StatusMap& status, std::vector<Key>& order, size_t from)
{
    std::stable_partition(order.begin() + from, order.end(), [&](Key k) {
        const Info& i(status.at(k));
        return i.is_present || !i.is_remove_only;
    });
}

void push(Path& p, int type, unsigned len)
{
    p.push_back(Seg{type, len});
    if (g_limits && p.size() > 32) throw LimitHit();
}

// :393-584
Cal makeStartPosAlignment(const Ctx& C, int ref_start, int read_start, unsigned read_length, const std::set<Key>& indels)
{
    if (read_length == 0 || ref_start < 0 || read_start < 0) throw Thrown();
    const bool leadingRead(read_start != 0);
    Cal cal;
    cal.pos = ref_start;
    int ref_head(ref_start), read_head(read_start);
    bool prevMismatch(false);
    for (const Key w : indels)
    {
        const sx_indel_key& ik(C.win[w]);
        const bool mm(isMismatch(ik));
        if (rightPos(ik) < ref_start) continue;
        if (rightPos(ik) == ref_start && (mm || !leadingRead)) continue;
        const bool firstIntersecting(cal.path.empty());
        if (leadingRead && firstIntersecting)
        {
            if (ik.pos != ref_start) throw Thrown();
            if (ik.ins_len == 0 || (int)ik.ins_len < read_start) throw Thrown();
            push(cal.path, SX_AP_INSERT, (unsigned)read_start);
            if (ik.del_len)
            {
                push(cal.path, SX_AP_DELETE, ik.del_len);
                ref_head += ik.del_len;
            }
            cal.leading = w;
            prevMismatch = mm;
            continue;
        }
        if (firstIntersecting && read_start != 0) throw Thrown();
        const bool edgeDelete(isPrimitiveDeletion(ik) && ik.pos == ref_start);
        const int gap(ik.pos - ref_head);
        if (gap < ((prevMismatch || mm) ? 0 : 1) && !edgeDelete) throw Thrown();
        if (gap < 0) throw Thrown();
        if (!(firstIntersecting || gap > 0 || mm || prevMismatch)) throw Thrown();
        const unsigned reach((unsigned)read_head + (unsigned)gap);
        if (reach > read_length || (reach == read_length && !isPrimitiveDeletion(ik))) break;
        if (gap > 0)
        {
            push(cal.path, SX_AP_MATCH, (unsigned)gap);
            ref_head += gap;
            read_head += gap;
        }
        if (mm)
        {
            push(cal.path, SX_AP_SEQ_MISMATCH, ik.del_len);
            ref_head += ik.del_len;
            read_head += ik.del_len;
            if (read_head >= (int)read_length) break;
        }
        else
        {
            if (ik.del_len)
            {
                push(cal.path, SX_AP_DELETE, ik.del_len);
                ref_head += ik.del_len;
            }
            if (ik.ins_len)
            {
                const unsigned room(read_length - (unsigned)read_head);
                const unsigned take(std::min<unsigned>(ik.ins_len, room));
                push(cal.path, SX_AP_INSERT, take);
                read_head += (int)take;
                if (ik.ins_len >= room)
                {
                    cal.trailing = w;
                    break;
                }
            }
            else if (gap == 0) cal.leading = w;
            else if (read_head == (int)read_length) cal.trailing = w;
        }
        prevMismatch = mm;
    }
    if (read_head > (int)read_length) throw Thrown();
    if (read_head < (int)read_length) push(cal.path, SX_AP_MATCH, read_length - (unsigned)read_head);
    return cal;
}

// :593-719
void endPinStartPos(const Ctx& C, const std::set<Key>& indels, unsigned read_length, int ref_end, int read_end, int& ref_start, int& read_start)
{
    if (read_length == 0 || ref_end <= 0 || read_end <= 0) throw Thrown();
    ref_start = ref_end;
    read_start = read_end;
    const bool trailingRead(read_end != (int)read_length);
    bool first(true), prevMismatch(false);
    for (auto it = indels.rbegin(); it != indels.rend(); ++it)
    {
        const sx_indel_key& ik(C.win[*it]);
        const bool mm(isMismatch(ik));
        if (ik.pos > ref_end) continue;
        if (ik.pos == ref_end && (mm || !trailingRead)) continue;
        if (!mm && rightPos(ik) == ref_end) // trailing-edge insertion / deletion
        {
            if (!(first && ref_start == ref_end)) throw Thrown();
            if (ik.ins_len > 0 && ik.ins_len < read_length - (unsigned)read_end) throw Thrown();
            ref_start -= (int)ik.del_len;
        }
        else
        {
            if (first && read_end != (int)read_length) throw Thrown();
            const int gap(ref_start - rightPos(ik));
            if (gap < ((prevMismatch || mm) ? 0 : 1)) throw Thrown();
            const int step(std::min(gap, read_start));
            ref_start -= step;
            read_start -= step;
            if (read_start == 0) return;
            ref_start -= (int)ik.del_len;
            if (mm)
            {
                read_start -= (int)ik.del_len;
                if (read_start == 0) return;
            }
            else if (ik.ins_len > 0)
            {
                if ((int)ik.ins_len >= read_start) return;
                read_start -= (int)ik.ins_len;
            }
        }
        first = false;
        prevMismatch = mm;
    }
    if (read_start < 0) throw Thrown();
    ref_start -= read_start;
    read_start = 0;
}

// :66-130
int updatedConstraints(int hc, int id, bool on, bool anyOn)
{
    if (hc < 0) return hc;
    if (id < 0 && on) return -1;
    if (id <= 0) return hc;
    const int want(on ? id : 3 - id);
    if (want == 0) return anyOn ? -1 : 0;
    if (want == 3) return hc > 0 ? hc : -1;
    if (hc == 3 || hc == want) return want;
    return anyOn ? -1 : 0;
}

bool updateHap(HapStatus& h, const std::vector<int>& ids, bool on)
{
    h.isAnyIndelOn = h.isAnyIndelOn || on;
    bool valid(false);
    for (size_t s = 0; s < h.constraints.size(); ++s)
    {
        h.constraints[s] = updatedConstraints(h.constraints[s], ids[s], on, h.isAnyIndelOn);
        if (h.constraints[s] >= 0) valid = true;
    }
    return valid;
}

std::set<Key> presentSet(const StatusMap& status)
{
    std::set<Key> s;
    for (const auto& kv : status)
        if (kv.second.is_present) s.insert(kv.first);
    return s;
}

// :857-1277; every container by value, as there
void search(const Ctx& C, unsigned read_length, std::set<Cal>& out, Warn& warn, StatusMap status, HapMap hapMap, std::vector<Key> order, unsigned depth,
            unsigned indelToggleDepth, unsigned totalToggleDepth, std::pair<int, int> read_range, int maxToggle, const Cal& cal)
{
    bool newIndels(indelToggleDepth == 0);
    {
        const size_t before(status.size());
        const std::pair<int, int> pr(softClipRange(cal));
        if (!(pr.first >= C.realign_begin && pr.second <= C.realign_end)) return;
        if (pr.first < read_range.first)
        {
            addIndelsInRange(C, pr.first, read_range.first + 1, status, order);
            read_range.first = pr.first;
        }
        if (pr.second > read_range.second)
        {
            addIndelsInRange(C, read_range.second - 1, pr.second, status, order);
            read_range.second = pr.second;
        }
        if (!newIndels) newIndels = (before != status.size());
        if (newIndels) removeOnlyLast(status, order, before);
    }
    if (depth == order.size())
    {
        Cal done(cal);
        const int b(cal.pos), e(cal.pos + (int)pathRefLength(cal.path));
        for (const auto& kv : status)
            if (kv.second.is_present && bpIntersect(b, e, C.win[kv.first])) done.indels.insert(kv.first);
        if (cal.leading >= 0) done.indels.insert(cal.leading);
        if (cal.trailing >= 0) done.indels.insert(cal.trailing);
        if (g_limits && (done.indels.size() > 24 || done.path.size() + g_nClipSegs > 32)) throw LimitHit();
        // (with limits on, alignments outside the realignment range are dropped here rather than after the search, as the device does,
        // so that the capacity applies to the same set)
        if (g_limits && !(b >= C.realign_begin && e <= C.realign_end)) return;
        out.insert(done);
        if (g_limits && out.size() > g_maxA) throw LimitHit();
        return;
    }
    if (newIndels)
    {
        const double maxIndels(read_length * C.opt->max_candidate_indel_density);
        maxToggle = (status.size() > maxIndels) ? 1 : C.opt->max_read_indel_toggle;
        const int limit(status.size() >= C.opt->n_max_toggle ? 1 : (int)C.opt->max_toggle[status.size()]);
        maxToggle = std::min(maxToggle, limit);
    }
    if ((int)indelToggleDepth > maxToggle)
    {
        warn.max_toggle_depth = true;
        return;
    }
    const Key curKey(order[depth]);
    const sx_indel_key& cur(C.win[curKey]);
    bool conflicting(false), haveUndiscovered(false);
    for (unsigned i = 0; i < depth; ++i)
    {
        if (!status[order[i]].is_present) continue;
        if (conflict(C.win[order[i]], cur)) conflicting = true;
        if (C.win[order[i]].flags & SX_IKF_NOT_DISCOVERED) haveUndiscovered = true;
    }
    const unsigned nSamples(C.opt->n_samples);
    const bool on(status[curKey].is_present);
    const int ar(C.hap ? C.hap[curKey].active_region_id : -1);
    const bool inAr(ar >= 0);
    if (inAr && !hapMap.count(ar))
    {
        if (g_limits && hapMap.size() >= 4) throw LimitHit();
        hapMap.insert(std::make_pair(ar, HapStatus(nSamples)));
    }
    const bool curUndiscovered((cur.flags & SX_IKF_NOT_DISCOVERED) != 0);
    std::vector<int> ids(nSamples, 0);
    if (inAr) // :811-849
        for (unsigned s = 0; s < nSamples; ++s)
        {
            int id(C.hap[curKey].haplotype_id[s]);
            if (id == 0)
            {
                bool ok(!C.opt->is_haplotyping_enabled || ((C.hap[curKey].bypass_mask >> s) & 1) || (cur.flags & SX_IKF_FORCED_OUTPUT));
                if (!ok && s == C.opt->sample_id && status[curKey].isInOriginalAlignment) ok = true;
                if (isMismatch(cur) && s != C.opt->sample_id) ok = false;
                id = ok ? 0 : -1;
            }
            ids[s] = id;
        }
    { // 1) the indel as it is
        HapMap next(hapMap);
        bool valid;
        if (!conflicting && inAr) valid = updateHap(next.at(ar), ids, on);
        else valid = !isMismatch(cur) || !on;
        if (on && haveUndiscovered && curUndiscovered) valid = false;
        if (!valid && totalToggleDepth == 0) valid = true;
        if (valid) search(C, read_length, out, warn, status, next, order, depth + 1, indelToggleDepth, totalToggleDepth, read_range, maxToggle, cal);
    }
    HapMap next(hapMap);
    bool valid;
    if (!conflicting && inAr) valid = updateHap(next.at(ar), ids, !on);
    else valid = !isMismatch(cur) || on;
    if (!on && haveUndiscovered && curUndiscovered) valid = false;
    if (!valid) return;
    if (!on && (status[curKey].is_remove_only || conflicting)) return;
    const unsigned inc(isMismatch(cur) ? 0 : 1);
    if ((int)(indelToggleDepth + inc) > maxToggle)
    {
        warn.max_toggle_depth = true;
        return;
    }
    status[curKey].is_present = !on;
    const std::set<Key> current(presentSet(status));
    { // 2) toggled, start position pinned
        const int ref_start(cal.pos);
        bool pinOk(true);
        if (!isMismatch(cur)) pinOk = !((cur.pos <= ref_start && ref_start < rightPos(cur)) || (on && curKey == cal.leading));
        if (pinOk)
        {
            const Cal start(makeStartPosAlignment(C, ref_start, (int)unalignedPrefix(cal.path), read_length, current));
            search(C, read_length, out, warn, status, next, order, depth + 1, indelToggleDepth + inc, totalToggleDepth + 1, read_range, maxToggle, start);
        }
    }
    if (isMismatch(cur) || cur.del_len == cur.ins_len) return;
    { // 3) toggled, end position pinned
        const int ref_end(cal.pos + (int)pathRefLength(cal.path));
        if ((cur.pos <= ref_end - 1 && ref_end - 1 < rightPos(cur)) || (on && curKey == cal.trailing)) return;
        int ref_start(0), read_start(0);
        endPinStartPos(C, current, read_length, ref_end, (int)read_length - (int)unalignedSuffix(cal.path), ref_start, read_start);
        if (ref_start < 0)
        {
            warn.origin_skip = true;
            return;
        }
        const Cal start(makeStartPosAlignment(C, ref_start, read_start, read_length, current));
        search(C, read_length, out, warn, status, next, order, depth + 1, indelToggleDepth + inc, totalToggleDepth + 1, read_range, maxToggle, start);
    }
}

// :1816-1994
void candidateAlignments(const Ctx& C, const sx_enum_batch& b, unsigned r, Warn& warn, std::set<Cal>& result)
{
    const unsigned read_length(b.read_len[r]);
    Cal cal;
    cal.pos = b.in_pos[r];
    g_nClipSegs = 0;
    for (unsigned s = b.in_seg_off[r]; s < b.in_seg_off[r + 1]; ++s) push(cal.path, b.in_segs[s].kind, b.in_segs[s].len);
    cal.leading = b.in_lead_key[r] == SX_NO_KEY ? -1 : (Key)b.in_lead_key[r];
    cal.trailing = b.in_trail_key[r] == SX_NO_KEY ? -1 : (Key)b.in_trail_key[r];
    StatusMap status;
    std::vector<Key> order;
    const std::pair<int, int> exemplar(softClipRange(cal));
    addIndelsInRange(C, exemplar.first, exemplar.second, status, order);
    {
        std::set<Key> validIndels;
        bool recompute(false);
        for (unsigned i = b.in_key_off[r]; i < b.in_key_off[r + 1]; ++i)
        {
            if (b.in_keys[i] == SX_NO_KEY) throw Thrown(); // an indel of the alignment that is no window entry (:1866-1872)
            const Key w(b.in_keys[i]);
            const bool mm(isMismatch(C.win[w]));
            auto it(status.find(w));
            if (it == status.end())
            {
                if (mm) continue;
                throw Thrown();
            }
            if (mm) recompute = true;
            it->second.is_present = true;
            it->second.isInOriginalAlignment = true;
            validIndels.insert(w);
        }
        if (recompute) cal = makeStartPosAlignment(C, cal.pos, (int)unalignedPrefix(cal.path), read_length, validIndels);
    }
    order.clear();
    for (const auto& kv : status)
        if (kv.second.is_present) order.push_back(kv.first);
    for (const auto& kv : status)
        if (!kv.second.is_present) order.push_back(kv.first);
    removeOnlyLast(status, order, 0);

    unsigned search_length(read_length), hc_lead(0), hc_trail(0), sc_lead(0), sc_trail(0);
    const bool clipped(!cal.path.empty() && (isClip(cal.path.front().type) || (cal.path.size() > 1 && isClip(cal.path.back().type))));
    if (clipped)
    {
        Path core;
        bool lead(true);
        for (const Seg& s : cal.path)
        {
            if (s.type == SX_AP_HARD_CLIP) (lead ? hc_lead : hc_trail) += s.length;
            else if (s.type == SX_AP_SOFT_CLIP) (lead ? sc_lead : sc_trail) += s.length;
            else
            {
                lead = false;
                if (hc_trail || sc_trail) throw Thrown();
                core.push_back(s);
            }
        }
        cal.path = core;
        if (search_length < sc_lead + sc_trail) throw Thrown();
        search_length -= sc_lead + sc_trail;
    }
    g_nClipSegs = (hc_lead ? 1 : 0) + (sc_lead ? 1 : 0) + (sc_trail ? 1 : 0) + (hc_trail ? 1 : 0);
    std::set<Cal> found;
    search(C, search_length, found, warn, status, HapMap(), order, 0, 0, 0, exemplar, C.opt->max_read_indel_toggle, cal);
    for (Cal c : found)
    {
        if (clipped)
        {
            Path p;
            if (hc_lead) push(p, SX_AP_HARD_CLIP, hc_lead);
            if (sc_lead) push(p, SX_AP_SOFT_CLIP, sc_lead);
            p.insert(p.end(), c.path.begin(), c.path.end());
            if (sc_trail) push(p, SX_AP_SOFT_CLIP, sc_trail);
            if (hc_trail) push(p, SX_AP_HARD_CLIP, hc_trail);
            c.path = p;
        }
        if (c.pos >= C.realign_begin && c.pos + (int)pathRefLength(c.path) <= C.realign_end) result.insert(c);
    }
}
} // namespace

// `limits` != 0: the per-read capacities of the device build (opts.max_alns_per_read (0 = 64) alignments in range, 64 indels in the
// search, 32 path segments, 24 keys per alignment, 4 active regions) are applied at the points of the search where the device
// applies them, so that SX_ENUM_ST_LIMIT lands on the same reads; 0: no limits (the reference has none).  A read that failed
// (SX_ENUM_ST_EXCEPTION / SX_ENUM_ST_LIMIT) reports that bit alone.
extern "C" int ox_enumerate_alignments(const sx_enum_batch* b, sx_enum_out* o, int limits)
{
    try
    {
        unsigned nA(0), nS(0), nK(0);
        o->aln_off[0] = 0;
        g_limits = (limits != 0);
        g_maxA = b->opts.max_alns_per_read ? b->opts.max_alns_per_read : 64u;
        // first pass: results; the arrays are filled as far as the capacities go, the totals always
        for (unsigned g = 0; g < b->n_regions; ++g)
        {
            Ctx C;
            const unsigned k0(b->region_key_off[g]);
            C.win = b->keys + k0;
            C.hap = b->key_hap ? b->key_hap + k0 : nullptr;
            C.n_win = b->region_key_off[g + 1] - k0;
            C.realign_begin = b->realign_begin[g];
            C.realign_end = b->realign_end[g];
            C.opt = &b->opts;
            for (unsigned r = b->region_read_off[g]; r < b->region_read_off[g + 1]; ++r)
            {
                C.usable.clear();
                for (unsigned i = b->use_key_off[r]; i < b->use_key_off[r + 1]; ++i) C.usable.insert((Key)b->use_keys[i]);
                Warn warn;
                std::set<Cal> cals;
                unsigned status(0);
                try
                {
                    if (!b->gate || (b->gate[r] & SX_GATE_REALIGN)) candidateAlignments(C, *b, r, warn, cals);
                }
                catch (const Thrown&)
                {
                    status = SX_ENUM_ST_EXCEPTION;
                    cals.clear();
                }
                catch (const LimitHit&)
                {
                    status = SX_ENUM_ST_LIMIT;
                    cals.clear();
                }
                if (status == 0)
                {
                    if (warn.origin_skip) status |= SX_ENUM_ST_ORIGIN_SKIP;
                    if (warn.max_toggle_depth) status |= SX_ENUM_ST_MAX_TOGGLE;
                }
                o->status[r] = (uint8_t)status;
                for (const Cal& c : cals)
                {
                    if (nA < o->cap_alns)
                    {
                        o->aln_pos[nA] = c.pos;
                        o->aln_seg_off[nA] = nS;
                        o->aln_key_off[nA] = nK;
                        o->aln_lead_key[nA] = c.leading < 0 ? (uint16_t)SX_NO_KEY : (uint16_t)c.leading;
                        o->aln_trail_key[nA] = c.trailing < 0 ? (uint16_t)SX_NO_KEY : (uint16_t)c.trailing;
                    }
                    for (const Seg& s : c.path)
                    {
                        if (nS < o->cap_segs)
                        {
                            o->segs[nS].kind = (uint8_t)s.type;
                            o->segs[nS].len = (uint16_t)s.length;
                            o->segs[nS].flags = 0;
                        }
                        ++nS;
                    }
                    for (const Key k : c.indels)
                    {
                        if (nK < o->cap_keys) o->aln_keys[nK] = (uint16_t)k;
                        ++nK;
                    }
                    ++nA;
                }
                o->aln_off[r + 1] = nA;
            }
        }
        o->totals[0] = nA;
        o->totals[1] = nS;
        o->totals[2] = nK;
        if (nA > o->cap_alns || nS > o->cap_segs || nK > o->cap_keys) return SX_ERR_CAPACITY;
        o->aln_seg_off[nA] = nS;
        o->aln_key_off[nA] = nK;
        return 0;
    }
    catch (const std::exception&)
    {
        return SX_ERR_ARG;
    }
}
