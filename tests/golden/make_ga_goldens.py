#!/usr/bin/env python
"""Extracts the known-answer cases of the reference's own GlobalAligner unit test
(/root/reference/src/c++/lib/alignment/test/GlobalAlignerTest.cpp) into tests/golden/global_aligner_goldens.json.
Run in the build container (needs /root/reference); the JSON is committed because the GPU box has no reference tree."""
import json
import os
import re
import sys

SRC = "/root/reference/src/c++/lib/alignment/test/GlobalAlignerTest.cpp"
txt = open(SRC).read()
cases = []
for m in re.finditer(r"BOOST_AUTO_TEST_CASE\(\s*(\w+)\s*\)\s*\{(.*?)\n\}", txt, re.S):
    name, body = m.group(1), m.group(2)
    seq = re.search(r'seq\(\s*"([^"]*)"\s*\)', body)
    ref = re.search(r'ref\(\s*"([^"]*)"\s*\)', body)
    call = re.search(r"testAlign\(\s*seq\s*,\s*ref\s*((?:,[^)]*)?)\)", body)
    cig = re.search(r'apath_to_cigar\(result\.align\.apath\)\s*,\s*"([^"]*)"', body)
    beg = re.search(r"result\.align\.beginPos\s*,\s*(-?\d+)", body)
    sco = re.search(r"result\.score\s*,\s*(-?\d+)", body)
    if not (seq and ref and call and cig and beg):
        print("skipping", name, file=sys.stderr)
        continue
    # testAlign(seq, ref, offEdgeScore=-4, insertDeleteScore=0, isAllowEdgeInsertion=false, isRequireEdgeDeletion=false)
    args = [a.strip() for a in call.group(1).split(",") if a.strip()]
    d = [-4, 0, False, False]
    for i, a in enumerate(args):
        d[i] = {"true": True, "false": False}.get(a, None) if a in ("true", "false") else int(a)
    # AlignmentScores<short>(2, -4, -5, -1, offEdge, insertDelete, allowEdgeIns, requireEdgeDel)  (GlobalAlignerTest.cpp:44)
    case = {"name": name, "query": seq.group(1), "ref": ref.group(1), "scores": [2, -4, -5, -1, d[0], d[1], d[2], d[3]],
            "cigar": cig.group(1), "beginPos": int(beg.group(1))}
    if sco:
        case["score"] = int(sco.group(1))
    cases.append(case)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "global_aligner_goldens.json")
json.dump({"source": SRC, "cases": cases}, open(out, "w"), indent=1)
print(len(cases), "cases ->", out)
