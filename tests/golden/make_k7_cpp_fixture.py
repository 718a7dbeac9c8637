#!/usr/bin/env python
"""Text fixture for the K7 part of the C++ host-mirror test (tests/cpp/test_host_mirror.cpp): k7_cases.tsv -- seeded regions
(window, reads with bases and normalized input alignments) with the candidate alignments the REFERENCE's getCandidateAlignments
returns for them (oracle/_ref/libstrelka_ref.so through oracle/ref_harness_enumerate.inc).  Run in the build container."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reflib  # noqa: E402
import specgen  # noqa: E402
from strelka_b200 import batch as B  # noqa: E402

CASES = (0, 1, 2, 8, 11)  # plain, clustered, phased two-sample (both samples), hard-clipped phased

with open(os.path.join(HERE, "k7_cases.tsv"), "w") as f:
    n_alns = 0
    for case in CASES:
        eb = specgen.enum_case(case)
        ref = reflib.ref_enumerate_alignments(eb, cap_alns=eb.n_reads * 6000 + 64)
        # K9: the reference's own scores (qualities 30 everywhere) and the realignment it chooses, default smoothing
        import numpy as np

        lnp, realn = reflib.ref_choose_realignment(eb, ref, np.full(int(eb.read_off[eb.n_reads]) + 1, 30, np.uint8))
        o = eb.opts
        f.write(f"BATCH\t{o.n_samples}\t{o.sample_id}\t{o.is_haplotyping_enabled}\t{o.max_read_indel_toggle}\n")
        for g in range(eb.n_regions):
            r0, r1 = int(eb.ref_off[g]), int(eb.ref_off[g + 1])
            f.write(f"REGION\t{bytes(eb.ref_pool[r0:r1]).decode()}\t{int(eb.ref_begin[g])}\t{int(eb.realign_begin[g])}\t{int(eb.realign_end[g])}\n")
            for k in range(int(eb.region_key_off[g]), int(eb.region_key_off[g + 1])):
                key, hap = eb.keys[k], eb.key_hap[k]
                ins = bytes(eb.ins_pool[int(eb.ins_off[k]) : int(eb.ins_off[k + 1])]).decode() or "-"
                fl = int(key["flags"])
                f.write(f"KEY\t{int(key['pos'])}\t{int(key['type']) + 1}\t{int(key['del_len'])}\t{ins}\t{fl & 1}\t{(fl >> 1) & 1}\t{(fl >> 2) & 1}\t"
                        f"{int(hap['active_region_id'])}\t{','.join(str(int(x)) for x in hap['haplotype_id'])}\t{int(hap['bypass_mask'])}\n")
            for r in range(int(eb.region_read_off[g]), int(eb.region_read_off[g + 1])):
                seq = bytes(eb.read_pool[int(eb.read_off[r]) : int(eb.read_off[r + 1])]).decode()
                cig = "".join(f"{int(s['len'])}{B.AP_CHAR[int(s['kind'])]}" for s in eb.in_segs[int(eb.in_seg_off[r]) : int(eb.in_seg_off[r + 1])])
                use = ";".join(str(int(x)) for x in eb.use_keys[int(eb.use_key_off[r]) : int(eb.use_key_off[r + 1])]) or "-"
                f.write(f"READ\t{seq}\t{int(eb.in_pos[r])}\t{cig}\t{use}\t{int(ref.status[r])}\n")
                for i, (pos, cigar, keys, lead, trail) in enumerate(ref.alignments_of(r)):
                    score = np.float64(lnp[int(ref.aln_off[r]) + i]).view(np.uint64)
                    f.write(f"ALN\t{pos}\t{cigar}\t{';'.join(map(str, keys)) or '-'}\t{lead}\t{trail}\t{int(score):016x}\n")
                    n_alns += 1
                if realn[r] is not None:
                    f.write(f"REALIGN\t{realn[r][0]}\t{realn[r][1]}\n")
        f.write("END\n")
print("k7_cases.tsv:", len(CASES), "batches,", n_alns, "alignments")
