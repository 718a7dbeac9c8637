#!/usr/bin/env python
"""Text fixtures for the C++ host-mirror test (tests/cpp/test_host_mirror.cpp):
   global_aligner_goldens.tsv  -- the reference's GlobalAlignerTest.cpp cases (from global_aligner_goldens.json)
   k1_cases.tsv                -- reference-shaped candidate alignments with the REFERENCE's scoreCandidateAlignment result
                                  (oracle/_ref/libstrelka_ref.so), doubles as hex bit patterns
Run in the build container."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reflib  # noqa: E402
import specgen  # noqa: E402

g = json.load(open(os.path.join(HERE, "global_aligner_goldens.json")))
with open(os.path.join(HERE, "global_aligner_goldens.tsv"), "w") as f:
    for c in g["cases"]:
        s = c["scores"]
        f.write("\t".join([c["name"], c["query"], c["ref"]] + [str(int(x)) for x in s] + [c["cigar"], str(c["beginPos"]), str(c.get("score", "NA"))]) + "\n")

INV = {1: "A", 2: "C", 4: "G", 8: "T", 15: "N", 0: "="}
rng = np.random.default_rng(2024)
with open(os.path.join(HERE, "k1_cases.tsv"), "w") as f:
    for _ in range(25):
        r = specgen.random_region(rng, n_reads=int(rng.integers(1, 5)))
        # the C++ API takes read bases as letters: restrict to nibbles that have one
        for codes, _q in r.reads:
            codes[~np.isin(codes, list(INV))] = 15
        lnp = reflib.ref_score_region(r)
        f.write(f"REGION\t{r.ref}\t{r.ref_begin}\n")
        for codes, q in r.reads:
            f.write("READ\t" + "".join(INV[int(c)] for c in codes) + "\t" + ",".join(str(int(x)) for x in q) + "\n")
        for a, cal in enumerate(r.alns):
            cig = "".join(f"{l}{t}" for t, l in cal.path)
            keys = ";".join(f"{k.pos}:{k.type}:{k.delete_length}:{k.insert_seq or '-'}:{int(k.is_candidate)}" for k in cal.indels) or "-"
            f.write(f"ALN\t{cal.read}\t{cal.pos}\t{cig}\t{keys}\t{cal.leading}\t{cal.trailing}\t{np.float64(lnp[a]).view(np.uint64):016x}\n")
print("cpp fixtures written")
