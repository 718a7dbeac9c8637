"""K7 enumerate_alignments without a GPU: the device body (strelka_b200/csrc/k7_core.cuh, __host__
__device__) compiled for the host and run the way the kernels run it, against
  * the known-answer vectors of the reference's own unit test (starling_common/test/starling_read_align_test.cpp),
  * the reference's getCandidateAlignments itself (oracle/_ref/libstrelka_ref.so) on seeded batches, when it is built here,
  * the frozen outputs of the same function (tests/golden/enumerate_ref.npz).
The GPU parity tests (tests/test_zz_gpu_enumerate.py) run the CUDA kernels against the same checkers."""
import json
import os
import re

import numpy as np
import pytest

import reflib
import specgen
from strelka_b200 import _abi as A
from strelka_b200 import batch as B

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not reflib.have_ref(), reason="oracle/_ref/libstrelka_ref.so not built (needs /root/reference)")
GOLD_NAMES = ("aln_off", "status", "aln_pos", "aln_seg_off", "segs", "aln_key_off", "aln_keys", "lead", "trail")


def _keys(specs):
    specs = sorted(specs, key=lambda k: k.order())
    arr = np.zeros(len(specs) + 1, dtype=A.INDEL_KEY_DT)
    for i, k in enumerate(specs):
        arr[i] = (k.pos, k.del_len, len(k.ins), 1 if k.ins else 0, A.SX_INDEL_TYPE_INDEL, A.SX_IKF_CANDIDATE, 0, 0.0, 0.0)
    return specs, arr


def _cigar(segs, n):
    return "".join(f"{int(s['len'])}{B.AP_CHAR[int(s['kind'])]}" for s in segs[:n])


def test_reference_unit_test_vectors():
    """test_make_start_pos_alignment and test_end_pin_start_pos of the reference's starling_read_align_test.cpp, through the device
    functions k7_make_start_pos / k7_end_pin_start_pos."""
    gold = json.load(open(os.path.join(HERE, "golden", "read_align_unit_goldens.json")))
    lib = reflib.k7core_enumerate(None)
    fixed = B.WindowKeySpec(**gold["fixed_key"])
    n_start = n_end = 0
    for case in gold["make_start_pos_alignment"]:
        specs, win = _keys([fixed, B.WindowKeySpec(**case["key"])])
        me = [i for i, k in enumerate(specs) if k is not fixed][0]
        pos, lead, trail, n_seg = np.zeros(1, np.int32), np.zeros(1, np.uint16), np.zeros(1, np.uint16), np.zeros(1, np.uint32)
        segs = np.zeros(40, dtype=A.ALN_SEG_DT)
        rc = lib.k7core_make_start_pos(A.ptr(win), 2, gold["ref_start"], case["read_start"], gold["read_length"], A.ptr(pos), A.ptr(lead), A.ptr(trail), A.ptr(segs),
                                       A.ptr(n_seg))
        assert rc == 0
        assert _cigar(segs, int(n_seg[0])) == case["path"], case
        if "pos" in case:
            assert int(pos[0]) == case["pos"]
        for side, got in (("leading", int(lead[0])), ("trailing", int(trail[0]))):
            if side in case:
                assert got == (me if case[side] else A.SX_NO_KEY), (case, side, got)
        n_start += 1
    for case in gold["get_end_pin_start_pos"]:
        _specs, win = _keys([fixed, B.WindowKeySpec(**case["key"])])
        ref_start, read_start = np.zeros(1, np.int32), np.zeros(1, np.int32)
        rc = lib.k7core_end_pin_start_pos(A.ptr(win), 2, gold["read_length"], gold["ref_end"], case["read_end"], A.ptr(ref_start), A.ptr(read_start))
        if case.get("throws"):
            assert rc == A.SX_ENUM_ST_EXCEPTION
        else:
            assert rc == 0 and (int(ref_start[0]), int(read_start[0])) == (case["ref_start"], case["read_start"]), case
        n_end += 1
    assert n_start == 12 and n_end == 15


def test_goldens_are_the_reference_unit_test():
    """the committed JSON still is what tests/golden/make_read_align_goldens.py extracts (only checkable where the reference is)."""
    src = "/root/reference/src/c++/lib/starling_common/test/starling_read_align_test.cpp"
    if not os.path.exists(src):
        pytest.skip("no /root/reference here")
    text = open(src).read()
    gold = json.load(open(os.path.join(HERE, "golden", "read_align_unit_goldens.json")))
    assert len(re.findall(r"path_compare\(\"", text)) == len(gold["make_start_pos_alignment"])
    for case in gold["make_start_pos_alignment"]:
        assert f'path_compare("{case["path"]}"' in text


def _same(a: B.EnumOut, b: B.EnumOut):
    for x, y in zip(a.trimmed(), b.trimmed()):
        assert x.tobytes() == y.tobytes()


@needs_ref
def test_max_toggle_table_is_the_references():
    """sx_default_enum_opts computes starling_align_limit's table itself (k7_enumerate.cu); the reference's for 5000 candidates."""
    o = A.default_enum_opts()
    want = reflib.ref_max_toggle_table(5000)
    got = np.array([o.max_toggle[i] if i < o.n_max_toggle else 1 for i in range(100)], np.uint8)
    assert np.array_equal(want, got) and o.n_max_toggle < 100


@needs_ref
def test_device_body_against_the_reference():
    """every alignment, in std::set order, with its keys, edge keys and the warn / exception status: identical to the reference's
    getCandidateAlignments on 48 seeded batches (plain, clustered, phased two-sample, dense, tight toggle budgets, hard clips)."""
    total, statuses = 0, set()
    for case in range(48):
        eb = specgen.enum_case(case)
        want = reflib.ref_enumerate_alignments(eb, cap_alns=eb.n_reads * 6000 + 64)
        rc, got = reflib.k7core_enumerate(eb, max_alns=6000, cap_alns=eb.n_reads * 6000 + 64)
        assert rc == 0
        _same(want, got)
        total += int(want.totals[0])
        statuses |= set(int(s) for s in want.status[: eb.n_reads])
    assert total > 15000
    assert {0, A.SX_ENUM_ST_MAX_TOGGLE, A.SX_ENUM_ST_EXCEPTION} <= statuses


def test_device_body_against_the_frozen_reference_output():
    gold = np.load(os.path.join(HERE, "golden", "enumerate_ref.npz"))
    total = 0
    for case in range(specgen.ENUM_GOLDEN_CASES):
        eb = specgen.enum_case(case)
        rc, got = reflib.k7core_enumerate(eb, max_alns=6000, cap_alns=eb.n_reads * 6000 + 64)
        assert rc == 0
        for name, arr in zip(GOLD_NAMES, got.trimmed()):
            assert arr.tobytes() == gold[f"{name}{case}"].tobytes(), (case, name)
        total += int(got.totals[0])
    assert total > 3000


def test_device_body_against_the_oracle_with_the_device_limits():
    """default per-read capacity (64 alignments, 64 indels, 32 segments, 24 keys): the oracle applies the same limits at the same
    points of the search, so SX_ENUM_ST_LIMIT lands on the same reads and everything else is identical (this is the comparison the
    GPU tests make); the oracle itself is pinned against the reference in tests/test_oracle_vs_reference.py and here on the goldens."""
    gold = np.load(os.path.join(HERE, "golden", "enumerate_ref.npz"))
    total = limited = 0
    for case in range(150):
        eb = specgen.enum_case(case)
        cap = eb.n_reads * 64 + 64
        want = reflib.ox_enumerate_alignments(eb, cap_alns=cap)
        rc, got = reflib.k7core_enumerate(eb, cap_alns=cap)
        assert rc == 0 and want.rc == 0
        _same(want, got)
        total += int(want.totals[0])
        limited += int((want.status[: eb.n_reads] & A.SX_ENUM_ST_LIMIT != 0).sum())
        if case < specgen.ENUM_GOLDEN_CASES:
            free = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 6000 + 64, limits=False)
            for name, arr in zip(GOLD_NAMES, free.trimmed()):
                assert arr.tobytes() == gold[f"{name}{case}"].tobytes(), (case, name)
    assert total > 20000 and limited > 50


def test_limits_and_capacity():
    """a read that needs more alignment slots than max_alns_per_read is flagged SX_ENUM_ST_LIMIT and contributes nothing (the other
    reads are unaffected); output arrays that are too small make the call fail with the needed sizes in totals[]."""
    eb = specgen.enum_case(3)
    rc, full = reflib.k7core_enumerate(eb, max_alns=6000, cap_alns=eb.n_reads * 6000 + 64)
    assert rc == 0
    n = np.diff(full.aln_off[: eb.n_reads + 1].astype(np.int64))
    cut = int(np.sort(n)[len(n) // 2])
    assert 0 < cut < n.max()
    rc, lim = reflib.k7core_enumerate(eb, max_alns=cut, cap_alns=eb.n_reads * 6000 + 64)
    assert rc == 0
    for r in range(eb.n_reads):
        if n[r] > cut:
            assert lim.status[r] & A.SX_ENUM_ST_LIMIT and lim.aln_off[r + 1] == lim.aln_off[r]
        else:
            assert lim.status[r] == full.status[r] and lim.alignments_of(r) == full.alignments_of(r)
    rc, small = reflib.k7core_enumerate(eb, max_alns=6000, cap_alns=int(full.totals[0]) - 1, cap_segs=1 << 20, cap_keys=1 << 20)
    assert rc == A.SX_ERR_CAPACITY and list(small.totals[:3]) == list(full.totals[:3])


def test_host_builder_restates_getAlignmentIndels():
    """batch.alignment_indels (the host side of K7's input): indels, swaps, edge keys and window mismatches of an input alignment."""
    ref, rb = "ACGTACGTACGTACGTACGTACGTACGTACGT", 100
    win = sorted([B.WindowKeySpec(104, 2, ""), B.WindowKeySpec(110, 0, "GG"), B.WindowKeySpec(116, 1, "TT"), B.WindowKeySpec(102, 1, "A", mismatch=True),
                  B.WindowKeySpec(121, 1, "T", mismatch=True)], key=lambda k: k.order())
    idx = {k.order(): i for i, k in enumerate(win)}
    #        100..103 M4 | D2 (104,105) | 106..109 M4 | I2 GG | 110..115 M6 | D1 (116) I2 TT swap | 117..122 M6
    seq = "ACAT" + "GTAC" + "GG" + "GTACGT" + "TT" + "CGTACG"
    r = B.EnumReadSpec(seq, 100, [("M", 4), ("D", 2), ("M", 4), ("I", 2), ("M", 6), ("D", 1), ("I", 2), ("M", 6)])
    keys, lead, trail = B.alignment_indels(r, ref, rb, win)
    want = sorted([idx[(104, 1, 0, 2, "")], idx[(110, 1, 2, 0, "GG")], idx[(116, 1, 2, 1, "TT")], idx[(102, 2, 1, 1, "A")]])
    assert keys == want and lead == A.SX_NO_KEY and trail == A.SX_NO_KEY
    with pytest.raises(KeyError):  # an indel of the alignment that the window does not hold: the reference throws (:1866-1872)
        B.alignment_indels(B.EnumReadSpec("ACGTAAACGT", 100, [("M", 4), ("I", 2), ("M", 4)]), ref, rb, win)


def test_cpp_host_mirror_builder_on_the_cpu(tmp_path):
    """sx::AlignmentSearchBatch (strelka_b200/host/strelka_b200.hh): reference-shaped objects in, CandidateAlignments out, against
    the candidate alignments the reference returned (tests/golden/k7_cases.tsv).  Without a GPU the library call in the middle is
    answered by the device body compiled for the host (tests/cpp/test_k7_mirror_cpu.cpp); tests/test_zz_gpu_enumerate.py::
    test_cpp_host_mirror_k7 runs the same check through libstrelka_b200.so."""
    import subprocess

    root = os.path.dirname(HERE)
    exe = str(tmp_path / "test_k7_mirror_cpu")
    lib = os.path.join(root, "strelka_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "strelka_b200", "host"), "-I" + lib,
                           os.path.join(HERE, "cpp", "test_k7_mirror_cpu.cpp"), "-o", exe, "-L" + lib, "-lstrelka_b200", "-Wl,-rpath," + lib])
    out = subprocess.run([exe, os.path.join(HERE, "golden")], capture_output=True, text=True)
    assert out.returncode == 0 and "0 failures" in out.stdout, out.stdout + out.stderr


def test_bench_workload_of_the_k7_leg():
    """bench.py's k7_enumerate leg: its numpy-built cfg2-shaped workload is a valid K7 batch, the device body and the oracle agree on
    it, and (where the reference library is built) both agree with the reference's getCandidateAlignments -- which is also how the
    leg itself spot-checks the GPU result."""
    import sys

    sys.path.insert(0, os.path.dirname(HERE))
    import bench

    eb = bench.make_enum_workload(300, 30, 150, 7)
    cap = eb.n_reads * 64 + 64
    want = reflib.ox_enumerate_alignments(eb, cap_alns=cap)
    rc, got = reflib.k7core_enumerate(eb, cap_alns=cap)
    assert rc == 0 and want.rc == 0
    _same(want, got)
    assert int(want.totals[0]) > 5 * eb.n_reads and not want.status[: eb.n_reads].any()
    assert eb.algorithmic_bytes(*[int(x) for x in want.totals[:3]]) > 200 * eb.n_reads
    if reflib.have_ref():
        _same(reflib.ref_enumerate_alignments(eb, cap_alns=cap), want)
        sub = bench.enum_subbatch(eb, 100)
        assert sub.n_regions == 100 and sub.n_reads == 3000 and sub.n_keys == 300


# ------------------------------------------------------------------------------------------------------------------------------
# K7b link_alignments: K7's output -> K1's alignment description
# ------------------------------------------------------------------------------------------------------------------------------
def _link_case(case):
    eb = specgen.enum_case(case)
    out = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)
    rng = np.random.default_rng(case)
    regions = B.regions_from_enumeration(eb, out, lambda r, n: rng.choice([11, 25, 37], n).astype(np.uint8))
    return eb, out, regions, B.build_align_batch(regions)


def _same_alignment_description(want: B.AlignBatch, got: B.AlignBatch):
    """per alignment: header, segments (no-op pads aside) and insert bytes."""
    assert want.n_alns == got.n_alns
    pad = lambda s: s[~((s["kind"] == A.SX_SEG_HARDCLIP) & (s["len"] == 0))]  # noqa: E731
    for i in range(want.n_alns):
        ws = pad(want.segs[int(want.alns["seg_off"][i]) : int(want.alns["seg_off"][i + 1])])
        gs = pad(got.segs[int(got.alns["seg_off"][i]) : int(got.alns["seg_off"][i + 1])])
        n_ins = int(ws["len"][ws["kind"] == A.SX_SEG_INSERT].sum())
        wi = bytes(want.ins[int(want.alns["ins_off"][i]) : int(want.alns["ins_off"][i]) + n_ins])
        gi = bytes(got.ins[int(got.alns["ins_off"][i]) : int(got.alns["ins_off"][i]) + n_ins])
        assert ws.tobytes() == gs.tobytes() and wi == gi, i
        assert want.alns["read"][i] == got.alns["read"][i] and want.alns["ref_pos"][i] == got.alns["ref_pos"][i]
    # K1's staging rule: every region's first segment index a multiple of 8, its first insert byte a multiple of 16
    assert not (got.regions["seg_begin"] % 8).any() and not (got.regions["ins_begin"] % 16).any()
    assert np.array_equal(got.regions["aln_begin"], want.regions["aln_begin"])


def test_link_device_body_against_the_host_flattening_and_the_reference():
    """K7b's device body (k8_core.cuh compiled for the host, run like the kernels of k8_link.cu) on K7's output: the K1 alignment
    description it writes equals what the host flattening (batch.flatten_alignment = the segment walk of scoreCandidateAlignment)
    writes for the same alignments; scoring both with the K1 oracle gives identical doubles, and -- where the reference library is
    built -- the reference's own scoreCandidateAlignment on the CandidateAlignment objects gives the same bits: enumerator -> link ->
    scorer is one pinned chain."""
    n = 0
    for case in range(36):
        eb, out, regions, want = _link_case(case)
        rc, lo = reflib.k8core_link(eb, out, want.regions)
        assert rc == 0
        got = lo.align_batch(want)
        _same_alignment_description(want, got)
        specgen.score_indels_batch_from_enumeration(eb, out, k6_segs=lo.k6_segs)  # asserts the K6-kind copy of K7's segments
        s_got = reflib.ox_score(got)
        assert np.array_equal(s_got.view(np.uint64), reflib.ox_score(want).view(np.uint64))
        if reflib.have_ref():
            s_ref = np.concatenate([reflib.ref_score_region(r) for r in regions])
            assert np.array_equal(s_ref.view(np.uint64), s_got.view(np.uint64)), case
        n += got.n_alns
    assert n > 5000


def test_link_capacity_and_bad_key():
    eb, out, regions, want = _link_case(1)
    rc, full = reflib.k8core_link(eb, out, want.regions)
    assert rc == 0
    rc, small = reflib.k8core_link(eb, out, want.regions, cap_segs=int(full.totals[0]) - 1)
    assert rc == A.SX_ERR_CAPACITY and list(small.totals) == list(full.totals)
    # a path gap that matches no key of its alignment (here: every window deletion made one base longer than the paths say): the
    # reference's assert(isFound) (score.cpp:222) -> an error status, not garbage
    assert (out.segs[: int(out.totals[1])]["kind"] == A.SX_AP_DELETE).any()
    eb.keys["del_len"][eb.keys["del_len"] > 0] += 1
    rc, _ = reflib.k8core_link(eb, out, want.regions)
    assert rc == -1


def test_fast_launch_plan_on_the_cpu():
    """SX_ENUM_F_FAST as the kernels run it (tests/cpp/k7_core_host.cpp::k7core_run_fast): the small scratch tier with K7_ST_RETRY, the
    arena tier for the marked reads, ONE search per read, blobs appended to a log in reverse read order, scan, gather -- identical to
    the oracle, whichever tier served a read."""
    total = retried = 0
    for case in range(60):
        eb = specgen.enum_case(case)
        cap = eb.n_reads * 64 + 64
        want = reflib.ox_enumerate_alignments(eb, cap_alns=cap)
        rc, got = reflib.k7core_enumerate(eb, cap_alns=cap, fast=True)
        assert rc == 0
        _same(want, got)
        total += int(want.totals[0])
        retried += got.n_retried
    assert total > 8000 and retried > 100  # both tiers did real work
    rc, small = reflib.k7core_enumerate(specgen.enum_case(3), cap_alns=5, fast=True)
    assert rc == A.SX_ERR_CAPACITY


# ------------------------------------------------------------------------------------------------------------------------------
# K7a alignment_indels: which window entries an input alignment already contains (getAlignmentIndels + the edge keys)
# ------------------------------------------------------------------------------------------------------------------------------
def _prep_inputs(eb):
    return (eb.in_key_off[: eb.n_reads + 1], eb.in_keys[: int(eb.in_key_off[eb.n_reads])], eb.in_lead_key[: eb.n_reads], eb.in_trail_key[: eb.n_reads])


def test_alignment_indels_device_body_host_builder_and_reference_agree():
    """K7a's device body (k7a_core.cuh compiled for the host, on K1's packed read / reference pools), the host builder
    (batch.alignment_indels) and -- where the reference library is built -- the reference's own getCandidateAlignment +
    getAlignmentIndels(includeMismatches) (oracle/ref_harness_enumerate.inc): the same keys, edge keys and SX_NO_KEY marks on tidy
    inputs and on the awkward ones (edge insertions / deletions, private indels, insert+delete runs, '=' and IUPAC bases); and K7 on
    those inputs is the reference's getCandidateAlignments, blt_exception included."""
    n_keys = n_nokey = n_edge = n_exc = 0
    for case in range(40):
        eb = specgen.enum_edge_case(case) if case % 2 else specgen.enum_case(case)
        rc, got = reflib.k7acore_prepare(eb, B.read_pools_of(eb))
        assert rc == 0
        for x, y in zip(got.trimmed(), _prep_inputs(eb)):
            assert np.array_equal(x, y), case
        if reflib.have_ref():
            for x, y in zip(got.trimmed(), reflib.ref_alignment_indels(eb).trimmed()):
                assert np.array_equal(x, y), case
            cap = eb.n_reads * 6000 + 64
            want = reflib.ref_enumerate_alignments(eb, cap_alns=cap)
            rc, en = reflib.k7core_enumerate(eb, max_alns=6000, cap_alns=cap)
            assert rc == 0
            _same(want, en)
            n_exc += int((want.status[: eb.n_reads] & A.SX_ENUM_ST_EXCEPTION != 0).sum())
        keys = got.trimmed()[1]
        n_keys += len(keys)
        n_nokey += int((keys == A.SX_NO_KEY).sum())
        n_edge += int((got.in_lead_key[: eb.n_reads] != A.SX_NO_KEY).sum() + (got.in_trail_key[: eb.n_reads] != A.SX_NO_KEY).sum())
    assert n_keys > 800 and n_nokey > 100 and n_edge > 50
    assert n_exc > 100 or not reflib.have_ref()
    rc, small = reflib.k7acore_prepare(eb, B.read_pools_of(eb), cap_keys=1)
    assert rc == A.SX_ERR_CAPACITY and small.totals[0] == got.totals[0]


# ------------------------------------------------------------------------------------------------------------------------------
# K9 choose_realignment: from the scores to rseg.realignment (the tail of scoreCandidateAlignments)
# ------------------------------------------------------------------------------------------------------------------------------
def _realign_expect(gold, name, case, tag, n_reads):
    pos, cig = gold[f"pos_{name}{case}_{tag}"], gold[f"cigar_{name}{case}_{tag}"]
    return [(int(pos[r]), str(cig[r])) if str(cig[r]) else None for r in range(n_reads)]


def test_choose_realignment_device_body_against_the_frozen_reference_output():
    """K9's device body (k9_core.cuh compiled for the host, poisoned per-thread map) fed with the REFERENCE's scores: the realignment of
    every read -- chosen alignment, soft-clipped ends where the smooth pool disagrees -- equals what the reference's
    scoreCandidateAlignments wrote into rseg.realignment (tests/golden/realign_ref.npz), for the default smoothing range, smoothing off and
    two wide ranges that put many alignments into the pool."""
    gold = np.load(os.path.join(HERE, "golden", "realign_ref.npz"))
    n = clipped = 0
    for name, case in specgen.REALIGN_GOLDEN_CASES:
        eb = specgen.realign_case_batch(name, case)
        out = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)
        for tag, (smooth, rng_) in specgen.REALIGN_MODES.items():
            lnp = gold[f"lnp_{name}{case}_{tag}"]
            rb = B.RealignBatch(eb, out, is_smoothed=smooth, smoothed_lnp_range=rng_)
            rc, got = reflib.k9core_choose(rb, np.concatenate([lnp, [0.0]]))
            assert rc == 0
            want = _realign_expect(gold, name, case, tag, eb.n_reads)
            for r in range(eb.n_reads):
                g = got.realignment_of(r)
                assert g == want[r], (name, case, tag, r)
                n += 1
                clipped += bool(g and "S" in g[1])
    assert n > 1000 and clipped > 50


@needs_ref
def test_choose_realignment_device_body_against_the_reference():
    """... and against the reference itself on further batches; its scores are also what the K1 oracle computes for the linked batch."""
    n = 0
    for case in range(20, 50):
        eb = specgen.enum_case(case)
        out = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)
        quals = specgen.realign_quals(eb, case)
        for smooth, rng_ in ((True, 2.302585092994046), (True, 60.0)):
            lnp, want = reflib.ref_choose_realignment(eb, out, quals, is_smoothed=smooth, smoothed_range=rng_)
            rc, got = reflib.k9core_choose(B.RealignBatch(eb, out, is_smoothed=smooth, smoothed_lnp_range=rng_), np.concatenate([lnp, [0.0]]))
            assert rc == 0
            for r in range(eb.n_reads):
                assert got.realignment_of(r) == want[r], (case, r)
                n += 1
        if case % 10 == 0:  # the reference's scores == the K1 oracle's on the batch K7b links
            regions = B.regions_from_enumeration(eb, out, lambda r, k: quals[int(eb.read_off[r]) : int(eb.read_off[r]) + k])
            assert np.array_equal(reflib.ox_score(B.build_align_batch(regions)).view(np.uint64), lnp.view(np.uint64))
    assert n > 1000
    # K4's segment kinds on request; capacity error
    rb = B.RealignBatch(eb, out, k4_kinds=True)
    rc, k4 = reflib.k9core_choose(rb, np.concatenate([lnp, [0.0]]))
    assert rc == 0 and set(np.unique(k4.segs["kind"][: int(k4.totals[0])])) <= {0, 1, 3, 4, 5, 6}
    rc, small = reflib.k9core_choose(rb, np.concatenate([lnp, [0.0]]), cap_segs=3)
    assert rc == A.SX_ERR_CAPACITY and small.totals[0] == k4.totals[0]


def test_choose_realignment_device_body_against_the_oracle():
    """every output array of K9's device body == oracle/realign_oracle.cpp's (the comparison the GPU tests make), on scores with many
    exact ties (a coarse grid) as well as on distinct ones; both segment-kind conventions."""
    n = 0
    for case in range(30):
        eb = specgen.enum_edge_case(case) if case % 3 == 0 else specgen.enum_case(case)
        out = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)
        rng = np.random.default_rng(case)
        nA = int(out.totals[0])
        for lnp in (-rng.random(nA + 1) * 40.0, -rng.integers(0, 4, nA + 1) * 2.0):
            for k4 in (False, True):
                rb = B.RealignBatch(eb, out, k4_kinds=k4)
                want = reflib.ox_choose_realignment(rb, lnp)
                rc, got = reflib.k9core_choose(rb, lnp)
                assert rc == 0 and want.rc == 0
                for nm in ("seg_off", "pos", "n_seg", "status", "best_aln"):
                    assert np.array_equal(getattr(want, nm)[: eb.n_reads], getattr(got, nm)[: eb.n_reads]), (case, nm)
                assert want.segs[: int(want.totals[0])].tobytes() == got.segs[: int(got.totals[0])].tobytes()
                n += eb.n_reads
    assert n > 2000


# ------------------------------------------------------------------------------------------------------------------------------
# K7g realign_gates: the front of realignAndScoreRead
# ------------------------------------------------------------------------------------------------------------------------------
def test_realign_gates_device_body_against_the_reference():
    """k7g_read (k7a_core.cuh compiled for the host) against the reference's own is_realignable / check_for_candidate_indel_overlap /
    normalizeInputAlignmentIndels / matchify_edge_soft_clip on mapper-style alignments (soft clips, edge insertions and deletions,
    over-long deletions, reads outside the realignment range or away from every candidate): the same reads pass, with the same
    normalized alignment."""
    if not reflib.have_ref():
        pytest.skip("oracle/_ref/libstrelka_ref.so not built (needs /root/reference)")
    n = 0
    seen = set()
    for case in range(40):
        eb = specgen.enum_edge_case(case) if case % 2 else specgen.enum_case(case)
        gb = B.GateBatch(eb, specgen.raw_alignments_for(eb, case))
        gate, want = reflib.ref_realign_gates(gb)
        rc, got = reflib.k7gcore_gates(gb)
        assert rc == 0
        for r in range(eb.n_reads):
            assert int(got.gate[r]) == int(gate[r]) and got.alignment_of(r) == want[r], (case, r)
            seen.add(int(gate[r]))
            n += 1
    assert n > 700 and seen >= {0, A.SX_GATE_REALIGN, A.SX_GATE_REALIGN | A.SX_GATE_SOFT_CLIPPED}


def test_search_ignores_zero_length_hard_clip_pads():
    """K7g writes a normalized path into the slots of the raw one and fills the rest with zero-length HARD_CLIP segments; K7a and K7 must
    not see them: the enumeration of a batch with padded input paths == that of the clean batch (device body and oracle)."""
    for case in (0, 1, 2, 5):
        rng = np.random.default_rng(case)
        regions = [specgen.random_enum_region(rng, n_reads=5, cluster=bool(case & 1), n_keys=(2, 7), hap=(case == 2)) for _ in range(4)]
        clean = B.EnumBatch(regions)
        padded = B.EnumBatch([(ref, rb, rr, win, [B.EnumReadSpec(r.seq, r.pos, list(r.path) + [("H", 0)] * int(rng.integers(1, 3)), r.use_keys) for r in reads])
                              for ref, rb, rr, win, reads in regions])
        assert np.array_equal(clean.in_keys, padded.in_keys)
        cap = clean.n_reads * 64 + 64
        want = reflib.ox_enumerate_alignments(clean, cap_alns=cap)
        _same(want, reflib.ox_enumerate_alignments(padded, cap_alns=cap))
        rc, got = reflib.k7core_enumerate(padded, cap_alns=cap)
        assert rc == 0
        _same(want, got)
        rc, prep = reflib.k7acore_prepare(padded, B.read_pools_of(padded))
        assert rc == 0 and np.array_equal(prep.trimmed()[1], padded.in_keys[: int(padded.in_key_off[padded.n_reads])])


def test_realign_gates_device_body_against_the_frozen_reference_output():
    gold = np.load(os.path.join(HERE, "golden", "gates_ref.npz"))
    n = 0
    for case in range(specgen.GATES_GOLDEN_CASES):
        eb = specgen.enum_edge_case(case) if case % 2 else specgen.enum_case(case)
        rc, got = reflib.k7gcore_gates(B.GateBatch(eb, specgen.raw_alignments_for(eb, case)))
        assert rc == 0 and np.array_equal(got.gate[: eb.n_reads], gold[f"gate{case}"])
        for r in range(eb.n_reads):
            want = (int(gold[f"pos{case}"][r]), str(gold[f"cigar{case}"][r])) if int(gold[f"gate{case}"][r]) & A.SX_GATE_REALIGN else None
            assert got.alignment_of(r) == want, (case, r)
            n += 1
    assert n > 200


def test_gated_out_reads_get_no_alignments():
    """sx_enum_batch.gate (K7g's output): a read whose SX_GATE_REALIGN bit is clear contributes no keys (K7a) and no alignments (K7, status 0);
    every other read is unaffected -- device bodies, both launch plans, and the oracle."""
    eb = specgen.enum_case(1)
    cap = eb.n_reads * 64 + 64
    full = reflib.ox_enumerate_alignments(eb, cap_alns=cap)
    rng = np.random.default_rng(5)
    gate = (rng.random(eb.n_reads + 1) < 0.6).astype(np.uint8) * A.SX_GATE_REALIGN | (rng.random(eb.n_reads + 1) < 0.3).astype(np.uint8) * A.SX_GATE_SOFT_CLIPPED
    eb.set_gate(gate)
    want = reflib.ox_enumerate_alignments(eb, cap_alns=cap)
    for fast in (False, True):
        rc, got = reflib.k7core_enumerate(eb, cap_alns=cap, fast=fast)
        assert rc == 0
        _same(want, got)
    off = 0
    for r in range(eb.n_reads):
        if gate[r] & A.SX_GATE_REALIGN:
            assert want.alignments_of(r) == full.alignments_of(r) and want.status[r] == full.status[r]
        else:
            assert want.aln_off[r + 1] == want.aln_off[r] and want.status[r] == 0
            off += 1
    assert 0 < off < eb.n_reads
    rc, prep = reflib.k7acore_prepare(eb, B.read_pools_of(eb))
    assert rc == 0
    for r in range(eb.n_reads):
        n = int(prep.in_key_off[r + 1]) - int(prep.in_key_off[r])
        assert n == (int(eb.in_key_off[r + 1]) - int(eb.in_key_off[r]) if gate[r] & A.SX_GATE_REALIGN else 0)
    eb.set_gate(None)
