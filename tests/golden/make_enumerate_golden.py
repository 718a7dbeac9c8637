"""Freezes the reference's own getCandidateAlignments (oracle/_ref/libstrelka_ref.so, oracle/ref_harness_enumerate.inc) on the first
specgen.ENUM_GOLDEN_CASES seeded K7 batches into tests/golden/enumerate_ref.npz, for the boxes that have no /root/reference.
Run in the build container after oracle/build_ref.sh."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reflib  # noqa: E402
import specgen  # noqa: E402

NAMES = ("aln_off", "status", "aln_pos", "aln_seg_off", "segs", "aln_key_off", "aln_keys", "lead", "trail")


def main():
    out, total = {}, 0
    for case in range(specgen.ENUM_GOLDEN_CASES):
        eb = specgen.enum_case(case)
        ref = reflib.ref_enumerate_alignments(eb, cap_alns=eb.n_reads * 6000 + 64)
        for name, arr in zip(NAMES, ref.trimmed()):
            out[f"{name}{case}"] = arr
        total += int(ref.totals[0])
    np.savez_compressed(os.path.join(HERE, "enumerate_ref.npz"), **out)
    print(specgen.ENUM_GOLDEN_CASES, "cases,", total, "alignments")


if __name__ == "__main__":
    main()
