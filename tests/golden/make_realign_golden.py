"""Freezes the reference's own scoreCandidateAlignments (oracle/_ref/libstrelka_ref.so, oracle/ref_harness_enumerate.inc) -- its scores and
the realignment it writes into rseg.realignment -- for seeded K7 batches into tests/golden/realign_ref.npz, for the boxes that have no
/root/reference.  Qualities: specgen.realign_quals(case).  Run in the build container after oracle/build_ref.sh."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reflib  # noqa: E402
import specgen  # noqa: E402


def main():
    out, n = {}, 0
    for name, case in specgen.REALIGN_GOLDEN_CASES:
        eb = specgen.realign_case_batch(name, case)
        enum = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)
        for tag, (smooth, rng_) in specgen.REALIGN_MODES.items():
            lnp, res = reflib.ref_choose_realignment(eb, enum, specgen.realign_quals(eb, case), is_smoothed=smooth, smoothed_range=rng_)
            out[f"lnp_{name}{case}_{tag}"] = lnp
            out[f"pos_{name}{case}_{tag}"] = np.array([r[0] if r else -1 for r in res], np.int64)
            out[f"cigar_{name}{case}_{tag}"] = np.array([r[1] if r else "" for r in res])
            n += len(res)
    np.savez_compressed(os.path.join(HERE, "realign_ref.npz"), **out)
    print(len(specgen.REALIGN_GOLDEN_CASES), "batches,", n, "read evaluations")


if __name__ == "__main__":
    main()
