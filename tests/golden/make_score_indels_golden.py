#!/usr/bin/env python
"""Freezes what the REFERENCE's own score_indels (oracle/_ref/libstrelka_ref.so, oracle/ref_harness_score_indels.cpp) writes on
seeded K6 batches into tests/golden/score_indels_ref.npz.  Run in the build container; the fixture is committed because the GPU
box has no reference tree.  The inputs are regenerated from the seed by tests/specgen.py (score_indels_case)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reflib  # noqa: E402
import specgen  # noqa: E402

out = {}
for case in range(specgen.SCORE_INDELS_GOLDEN_CASES):
    sb, lnp = specgen.score_indels_case(case)
    assert np.array_equal(reflib.ref_candidate_alignment_order(sb), np.arange(sb.n_alns))
    recs, n_rec, max_aln = reflib.ref_score_indels(sb, lnp)
    out[f"recs{case}"] = np.frombuffer(recs.tobytes(), dtype=np.uint8)
    out[f"n_rec{case}"] = n_rec
    out[f"max_aln{case}"] = max_aln
np.savez_compressed(os.path.join(HERE, "score_indels_ref.npz"), **out)
print("score_indels golden written:", sum(int(out[f"n_rec{c}"].sum()) for c in range(specgen.SCORE_INDELS_GOLDEN_CASES)), "records")
