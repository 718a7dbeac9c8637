"""Freezes the reference's own gate functions of realignAndScoreRead (is_realignable, check_for_candidate_indel_overlap,
normalizeInputAlignmentIndels, matchify_edge_soft_clip; oracle/ref_harness_enumerate.inc: ref_realign_gates) on seeded mapper-style
alignments into tests/golden/gates_ref.npz, for the boxes that have no /root/reference.  Run in the build container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reflib  # noqa: E402
import specgen  # noqa: E402
from strelka_b200 import batch as B  # noqa: E402

out, n = {}, 0
for case in range(specgen.GATES_GOLDEN_CASES):
    eb = specgen.enum_edge_case(case) if case % 2 else specgen.enum_case(case)
    gate, res = reflib.ref_realign_gates(B.GateBatch(eb, specgen.raw_alignments_for(eb, case)))
    out[f"gate{case}"] = gate
    out[f"pos{case}"] = np.array([r[0] if r else 0 for r in res], np.int64)
    out[f"cigar{case}"] = np.array([r[1] if r else "" for r in res])
    n += len(res)
np.savez_compressed(os.path.join(HERE, "gates_ref.npz"), **out)
print(specgen.GATES_GOLDEN_CASES, "batches,", n, "reads")
