// k7_mirror_check.hh -- TEST ONLY: the K7 part of the C++ host-mirror test, shared by tests/cpp/test_k7_mirror.cpp (GPU: the real
// library) and tests/cpp/test_k7_mirror_cpu.cpp (no GPU: the same builder / decoder code of sx::AlignmentSearchBatch, with
// sx_enumerate_alignments answered by the device body compiled for the host).
//
// sx::AlignmentSearchBatch against the candidate alignments the REFERENCE's getCandidateAlignments returned for the same regions
// (tests/golden/k7_cases.tsv, written by tests/golden/make_k7_cpp_fixture.py from oracle/_ref/libstrelka_ref.so).
#pragma once

#include "strelka_b200.hh"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

inline std::vector<std::string> k7_split(const std::string& s, char d)
{
    std::vector<std::string> out;
    std::string cur;
    std::istringstream is(s);
    while (std::getline(is, cur, d)) out.push_back(cur);
    return out;
}

inline void k7_mirror_check(const sx::Context& ctx, const std::string& dir, int& checks, int& failures)
{
    std::ifstream in(dir + "/k7_cases.tsv");
    std::string line;
    struct Want
    {
        int status;
        std::vector<std::vector<std::string>> alns;
        std::string realignPos, realignCigar; // the reference's rseg.realignment ("" = not realigned)
    };
    sx::AlignmentSearchBatch batch;
    std::vector<Want> want;
    std::vector<std::vector<sx::IndelKey>> windowOfRead;
    std::vector<sx::IndelSearchEntry> window;
    std::vector<std::string> region; // pending REGION line: opened when its first READ arrives (the window is complete then)
    sx_enum_opts opts;
    sx_default_enum_opts(&opts);
    unsigned k7checks(0), k7alns(0);
    auto keyString = [](const sx::IndelKey& k, const std::vector<sx::IndelKey>& win) {
        for (size_t i(0); i < win.size(); ++i)
            if (win[i] == k) return std::to_string(i);
        return std::string("?");
    };
    auto flush = [&]() {
        std::vector<sx::AlignmentSearchBatch::ReadResult> res;
        batch.enumerate(ctx, &opts, res);
        // K9 through the mirror: the reference's own scores (frozen as bit patterns) -> the reference's realignments
        std::vector<double> scores;
        bool scorable(true);
        for (size_t r(0); r < res.size(); ++r)
        {
            if (res[r].alignments.size() != want[r].alns.size()) scorable = false;
            for (const std::vector<std::string>& w : want[r].alns)
            {
                const unsigned long long bits(std::strtoull(w[6].c_str(), nullptr, 16));
                double d;
                std::memcpy(&d, &bits, 8);
                scores.push_back(d);
            }
        }
        std::vector<sx::AlignmentSearchBatch::Realignment> realn;
        if (scorable) batch.chooseRealignments(ctx, scores, true, 2.302585092994046, realn);
        for (size_t r(0); r < res.size(); ++r)
        {
            ++checks;
            ++k7checks;
            const int st((res[r].originSkip ? 1 : 0) | (res[r].maxToggleDepth ? 2 : 0) | (res[r].threw ? 4 : 0) | (res[r].overLimit ? 8 : 0));
            bool ok(st == want[r].status && res[r].alignments.size() == want[r].alns.size());
            for (size_t a(0); ok && a < want[r].alns.size(); ++a)
            {
                const sx::CandidateAlignment& cal(res[r].alignments[a]);
                const std::vector<std::string>& w(want[r].alns[a]); // ALN pos cigar keys lead trail
                std::string keys;
                for (const sx::IndelKey& k : cal.indels) keys += (keys.empty() ? "" : ";") + keyString(k, windowOfRead[r]);
                if (keys.empty()) keys = "-";
                const std::string lead(cal.leading_indel_key.type == sx::INDEL::NONE ? "65535" : keyString(cal.leading_indel_key, windowOfRead[r]));
                const std::string trail(cal.trailing_indel_key.type == sx::INDEL::NONE ? "65535" : keyString(cal.trailing_indel_key, windowOfRead[r]));
                ok = cal.al.pos == atoi(w[1].c_str()) && sx::apath_to_cigar(cal.al.path) == w[2] && keys == w[3] && lead == w[4] && trail == w[5];
                ++k7alns;
            }
            if (ok && scorable)
            {
                ++checks;
                const bool wantRealigned(!want[r].realignCigar.empty());
                if (realn[r].is_realigned != wantRealigned ||
                    (wantRealigned && (realn[r].realignment.pos != atoi(want[r].realignPos.c_str()) || sx::apath_to_cigar(realn[r].realignment.path) != want[r].realignCigar)))
                {
                    ++failures;
                    std::cerr << "FAIL k9 read " << r << ": realignment " << realn[r].realignment.pos << " " << sx::apath_to_cigar(realn[r].realignment.path) << ", want "
                              << want[r].realignPos << " " << want[r].realignCigar << "\n";
                }
            }
            if (!ok)
            {
                ++failures;
                std::cerr << "FAIL k7 read " << r << ": status " << st << " (want " << want[r].status << "), " << res[r].alignments.size() << " alignments (want "
                          << want[r].alns.size() << ")\n";
            }
        }
        batch = sx::AlignmentSearchBatch();
        want.clear();
        windowOfRead.clear();
    };
    while (std::getline(in, line))
    {
        const std::vector<std::string> f(k7_split(line, '\t'));
        if (f.empty()) continue;
        if (f[0] == "BATCH")
        {
            sx_default_enum_opts(&opts);
            opts.max_alns_per_read = 6000;
            opts.n_samples = atoi(f[1].c_str());
            opts.sample_id = atoi(f[2].c_str());
            opts.is_haplotyping_enabled = atoi(f[3].c_str());
            opts.max_read_indel_toggle = atoi(f[4].c_str());
        }
        else if (f[0] == "REGION")
        {
            region = f;
            window.clear();
        }
        else if (f[0] == "KEY")
        {
            sx::IndelSearchEntry e;
            e.key = sx::IndelKey(atoi(f[1].c_str()), atoi(f[2].c_str()) == 2 ? sx::INDEL::MISMATCH : sx::INDEL::INDEL, atoi(f[3].c_str()), f[4] == "-" ? "" : f[4].c_str());
            e.isCandidate = f[5] == "1";
            e.notDiscoveredFromReads = f[6] == "1";
            e.isForcedOutput = f[7] == "1";
            e.activeRegionId = atoi(f[8].c_str());
            const std::vector<std::string> h(k7_split(f[9], ','));
            for (unsigned s2(0); s2 < 4; ++s2)
            {
                e.haplotypeId[s2] = static_cast<int8_t>(atoi(h[s2].c_str()));
                e.isHaplotypingBypassed[s2] = ((atoi(f[10].c_str()) >> s2) & 1) != 0;
            }
            window.push_back(e);
        }
        else if (f[0] == "READ")
        {
            if (!region.empty())
            {
                batch.beginRegion(window, region[1], atoi(region[2].c_str()), atoi(region[3].c_str()), atoi(region[4].c_str()));
                region.clear();
            }
            sx::alignment al;
            al.pos = atoi(f[2].c_str());
            sx::cigar_to_apath(f[3].c_str(), al.path);
            std::vector<sx::IndelKey> observed, winKeys;
            for (const sx::IndelSearchEntry& e : window) winKeys.push_back(e.key);
            if (f[4] != "-")
                for (const std::string& i : k7_split(f[4], ';')) observed.push_back(winKeys[atoi(i.c_str())]);
            batch.addRead(f[1], al, observed);
            want.push_back(Want{atoi(f[5].c_str()), {}, "", ""});
            windowOfRead.push_back(winKeys);
        }
        else if (f[0] == "ALN") want.back().alns.push_back(f);
        else if (f[0] == "REALIGN")
        {
            want.back().realignPos = f[1];
            want.back().realignCigar = f[2];
        }
        else if (f[0] == "END") flush();
    }
    if (k7checks < 100 || k7alns < 1000)
    {
        ++failures;
        std::cerr << "FAIL k7: fixture too small (" << k7checks << " reads, " << k7alns << " alignments)\n";
    }
}
