"""Pins oracle/strelka_oracle.cpp (the travelling CPU checker) against the reference's own code
(oracle/_ref/libstrelka_ref.so = /root/reference compiled by oracle/build_ref.sh).  CPU only."""
import numpy as np
import pytest

import reflib
import specgen
from strelka_b200 import _abi as A
from strelka_b200 import batch as B

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("seed", range(8))
def test_score_alignments_bit_exact(seed):
    rng = np.random.default_rng(seed)
    regions = [specgen.random_region(rng, n_reads=int(rng.integers(1, 8))) for _ in range(12)]
    regions.append(specgen.simple_region(rng, n_reads=10))
    batch = B.build_align_batch(regions)
    got = reflib.ox_score(batch)
    want = np.concatenate([reflib.ref_score_region(r) for r in regions])
    assert got.shape == want.shape
    # bit-exact: compare the raw IEEE-754 words
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert np.all(want <= 0)


def _ga_scores(match, mismatch, open_, extend, off_edge, ins_del=0, allow_edge_ins=False, require_edge_del=False):
    return A.SxGaScores(match, mismatch, open_, extend, off_edge, ins_del, int(allow_edge_ins), int(require_edge_del))


@pytest.mark.parametrize("flags", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("seed", range(3))
def test_global_align_bit_exact(seed, flags):
    rng = np.random.default_rng(100 + seed)
    qs, rs = specgen.random_ga_problems(rng, 150, n_frac=0.01)
    gb = B.GaBatch(qs, rs, max_ops=400)
    for sc in (_ga_scores(1, -4, -5, -1, -100, -5, *flags), _ga_scores(2, -4, -5, -1, -1, 0, *flags)):
        r_res, r_cig = reflib.ref_global_align(sc, gb)
        o_res, o_cig = reflib.ox_global_align(sc, gb)
        assert np.array_equal(r_res, o_res)
        assert np.array_equal(r_cig, o_cig)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("always", [True, False])
def test_site_gl_germline(seed, always):
    rng = np.random.default_rng(200 + seed)
    pb = specgen.random_pileups(rng, 400, depth=[8.0, 30.0, 60.0, 120.0][seed])
    p = A.default_params()
    want = reflib.ref_germline(p, pb, always)
    got = reflib.ox_germline(p, pb, always)
    for f in ("ref_gt", "is_computed", "n_used_calls", "phredLoghood", "strand_bias"):
        assert np.array_equal(want[f], got[f]), f
    for rs in ("genome", "poly"):
        for f in ("max_gt", "snp_qphred", "max_gt_qphred", "ref_pprob"):
            assert np.array_equal(want[rs][f], got[rs][f]), (rs, f)
    assert np.array_equal(want["lhood"].view(np.uint32), got["lhood"].view(np.uint32))
    ro, rde = reflib.ref_dependent_eprob(p, pb)
    oo, ode = reflib.ox_dependent_eprob(p, pb)
    assert np.array_equal(ro, oo) and np.array_equal(rde.view(np.uint32), ode.view(np.uint32))


def test_site_gl_germline_haploid_and_nodep():
    rng = np.random.default_rng(7)
    pb0 = specgen.random_pileups(rng, 300, depth=25.0)
    pl = rng.integers(1, 3, pb0.n_sites).astype(np.uint8)
    pb = B.PileupBatch(pb0.site_off, pb0.calls, pb0.ref_base, pl)
    for p in (A.default_params(), A.SxParams(0.001, 0.0, 0.0, 0, 1, 0.0, 0.0, 1e-4, 5e-10, 0.0, 0.15, 0, 0)):
        want = reflib.ref_germline(p, pb, True)
        got = reflib.ox_germline(p, pb, True)
        assert want.tobytes() == got.tobytes()


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("tier2", [False, True])
def test_site_gl_somatic(seed, tier2):
    rng = np.random.default_rng(300 + seed)
    n = 300
    npb = specgen.random_pileups(rng, n, depth=30.0, with_tier2=tier2, alt_frac_choices=(0.0, 0.0, 0.0, 0.0, 0.02, 0.5))
    tpb0 = specgen.random_pileups(rng, n, depth=60.0, with_tier2=tier2, alt_frac_choices=(0.0, 0.0, 0.05, 0.1, 0.2, 0.4))
    # same reference base at a site in both samples
    tpb = B.PileupBatch(tpb0.site_off, tpb0.calls, npb.ref_base, None, tpb0.t2_off, tpb0.t2_calls)
    forced = (rng.random(n) < 0.2).astype(np.uint8)
    p = A.default_params()
    want = reflib.ref_somatic(p, npb, tpb, forced)
    got = reflib.ox_somatic(p, npb, tpb, forced)
    assert np.array_equal(want["is_computed"], got["is_computed"])
    m = want["is_computed"] == 1
    assert m.sum() > 10
    for f in ("ref_gt", "snv_tier", "snv_from_ntype_tier", "ntype", "max_gt", "qphred", "from_ntype_qphred", "normal_alt_id", "tumor_alt_id"):
        assert np.array_equal(want[f][m], got[f][m]), f
    assert np.array_equal(want["strandBias"][m].view(np.uint32), got["strandBias"][m].view(np.uint32))


def test_flat_batch_reference_scorer_matches_oracle_on_bench_workload():
    """bench.py --impl reference scores the synthetic workload with the reference's own scoreCandidateAlignment, rebuilt from the
    flattened batch; it must agree bit for bit with the oracle on that batch (so both arms of the bench compute the same thing)."""
    import ctypes as C
    import sys, os

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    synth = bench.load_synth()
    ab, pb, gb = bench.make_workload(synth, bench.HostAlloc(None, False), 300, 30, 150, 4, 7, 4)
    want = reflib.ox_score(ab)
    got = np.zeros(ab.n_alns, np.float64)
    secs = C.c_double(0)
    err = C.create_string_buffer(512)
    rc = reflib.ref().ref_score_flat_batch(C.byref(ab.c), C.c_uint32(0), C.c_uint32(ab.n_regions), C.c_void_p(got.ctypes.data), C.byref(secs), err, 512)
    assert rc == 0, err.value
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert secs.value > 0
    # the other two kernels' inputs from the same generator
    p = A.default_params()
    assert reflib.ref_germline(p, pb, True).tobytes() == reflib.ox_germline(p, pb, True).tobytes()
    sc = A.SxGaScores(1, -4, -5, -1, -100, -5, 1, 1)
    r1, c1 = reflib.ref_global_align(sc, gb)
    r2, c2 = reflib.ox_global_align(sc, gb)
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2)


@pytest.mark.parametrize("seed", range(3))
def test_compact_wire_formats_are_lossless(seed):
    """sx_aln8 / sx_aln_seg2 / 2-bit qualities carry the same batch: the oracle and the reference's own scorer return the same doubles
    on the compact batch as on the wide one (full and partial region ranges)."""
    import ctypes as C

    rng = np.random.default_rng(900 + seed)
    regions = [specgen.random_region(rng, n_reads=int(rng.integers(1, 10))) for _ in range(25)]
    for r in regions:  # <= 4 distinct qualities so that the 2-bit format applies
        r.reads = [(codes, np.array([11, 25, 37, 2], np.uint8)[np.asarray(q) % 4]) for codes, q in r.reads]
    regions += [specgen.simple_region(rng, n_reads=int(rng.integers(3, 30))) for _ in range(10)]
    wide = B.build_align_batch(regions)
    want = reflib.ox_score(wide)
    for qb, compact in ((8, True), (4, True), (2, False), (2, True)):
        cb = B.build_align_batch(regions, qual_bits=qb, compact=compact)
        assert cb.fmt == (0 if not compact else 15 if qb == 2 else 11)  # ALN8 | SEG2 | REF4 (| BASEQ with the 2-bit dictionary)
        assert cb.cells() == wide.cells()
        assert np.array_equal(reflib.ox_score(cb).view(np.uint64), want.view(np.uint64)), (qb, compact)
        got = np.zeros(cb.n_alns, np.float64)
        secs = C.c_double(0)
        err = C.create_string_buffer(512)
        lo, hi = 26, cb.n_regions - 1  # a partial range of the cfg2-shaped regions (the flat-batch shim rebuilds only those shapes)
        rc = reflib.ref().ref_score_flat_batch(C.byref(cb.c), C.c_uint32(lo), C.c_uint32(hi), C.c_void_p(got.ctypes.data), C.byref(secs), err, 512)
        assert rc == 0, err.value
        a0, a1 = int(cb.regions["aln_begin"][lo]), int(cb.regions["aln_begin"][hi])
        assert np.array_equal(got[a0:a1].view(np.uint64), want[a0:a1].view(np.uint64)), (qb, compact)


@pytest.mark.parametrize("seed", range(3))
def test_indel_genotype_likelihoods(seed):
    rng = np.random.default_rng(400 + seed)
    ib = B.IndelBatch(specgen.random_indel_loci(rng, 150))
    p = A.default_params()
    want = reflib.ref_indel_gl(p, ib)
    got = reflib.ox_indel_gl(p, ib)
    assert np.array_equal(want["n_gt"], got["n_gt"])
    assert np.array_equal(want["support"], got["support"])
    assert np.array_equal(want["gt_lhood"].view(np.uint64), got["gt_lhood"].view(np.uint64))


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("mode", ["germline", "somatic", "nofilter", "edge"])
def test_pileup_reads_columns_match_the_reference(seed, mode):
    """f1: the restated pileup loop (oracle) against starling_pos_processor_base::pileup_read_segment itself, driven read by read on
    a real pos processor: every position's tier1 / tier2 base_call column in order, spanning-deletion and sub-mapped counts."""
    rng = np.random.default_rng(4000 + seed)
    reads, ref, ref_begin, cand = specgen.random_pileup_reads(rng, n_reads=250)
    opts = A.default_pileup_opts()
    if mode == "somatic":
        opts = A.SxPileupOpts(1, 0, 20, 3, 1, 10, 0, 0)
    elif mode == "nofilter":
        opts = A.SxPileupOpts(0, 17, 0, 0, 0, 10, 0, 0)
    elif mode == "edge":
        opts = A.SxPileupOpts(1, 17, 3, 1, 1, 2, 5, 0)
    lo, hi = ref_begin + 100, ref_begin + len(ref) - 150  # reads hang off both ends of the report range
    pb = B.PileupReadsBatch(reads, ref, ref_begin, lo, hi, cand, opts)
    want = reflib.ref_pileup_reads(pb)
    got = reflib.ox_pileup_reads(pb)
    assert int(want[0][-1]) > 1000 and int(want[4].sum()) > 0 and int(want[5].sum()) > 0
    for w, g, name in zip(want, got, ("site_off", "calls", "t2_off", "t2_calls", "n_spandel", "n_submapped")):
        assert np.array_equal(w, g), name


@pytest.mark.parametrize("seed", range(3))
def test_pileup_reads_in_read_buffer_order_match_the_reference(seed):
    """The reference piles reads up in read-buffer order while each contributes through its best alignment (realignments move starts):
    the oracle on a batch in buffer order with unsorted alignments and dictionary-coded qualities == pileup_read_segment driven in that order."""
    rng = np.random.default_rng(4100 + seed)
    reads, ref, ref_begin, cand = specgen.random_pileup_reads(rng, n_reads=300)
    bpos = [int(r.pos) + (int(rng.integers(-25, 26)) if rng.random() < 0.4 else 0) for r in reads]
    order = np.argsort(np.asarray(bpos), kind="stable")
    reads, bpos = [reads[i] for i in order], [bpos[i] for i in order]
    qd = sorted({int(q) for r in reads for q in r.quals})
    if len(qd) > 16:
        qd = qd[:: (len(qd) + 15) // 16][:16]
        for r in reads:
            r.quals = [min(qd, key=lambda v: abs(v - int(q))) for q in r.quals]
    lo, hi = ref_begin + 100, ref_begin + len(ref) - 150
    plain = B.PileupReadsBatch(reads, ref, ref_begin, lo, hi, cand, buffer_pos=bpos)
    packed = B.PileupReadsBatch(reads, ref, ref_begin, lo, hi, cand, buffer_pos=bpos, qual_dict=qd)
    want = reflib.ref_pileup_reads(plain)
    for pb in (plain, packed):
        for w, g, name in zip(want, reflib.ox_pileup_reads(pb), ("site_off", "calls", "t2_off", "t2_calls", "n_spandel", "n_submapped")):
            assert np.array_equal(w, g), name
    for w, g in zip(want, reflib.ref_pileup_reads(packed)):
        assert np.array_equal(w, g)


def test_pileup_bench_workload_matches_the_reference():
    """The K4 leg of bench.py: the oracle and the reference's pileup_read_segment agree on the synthetic read set it times."""
    import ctypes as C
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    w = bench.make_pileup_reads_workload(8000, 30, 150, 3)
    hb = A.SxPileupReadsBatch(w["n_reads"], w["n_segs"], A.ptr(w["reads"]), A.ptr(w["seq4"]), A.ptr(w["qual"]), A.ptr(w["segs"]), A.ptr(w["ref"]), 0, w["ref_len"],
                              w["report_begin"], w["report_end"], None, 0, w["max_ref_span"], 150, 0, A.default_pileup_opts())

    class _PB:  # the two fields the reflib wrappers need beside the ABI struct
        c, n_sites, total_bases = hb, w["report_end"] - w["report_begin"], w["bases"]

    want, got = reflib.ref_pileup_reads(_PB), reflib.ox_pileup_reads(_PB)
    assert int(want[0][-1]) > 0.9 * w["bases"] * 8000 / (8000 + 150) and int(want[4].sum()) > 0
    for a, b, name in zip(want, got, ("site_off", "calls", "t2_off", "t2_calls", "n_spandel", "n_submapped")):
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("case", range(40))
def test_score_indels_matches_the_reference(case):
    """K6 oracle vs the reference's own score_indels (and its isFirstCandidateAlignmentPreferred tie-break) on rebuilt IndelBuffer /
    read_segment / std::set<CandidateAlignment> objects: records (ReadPathScores incl. alternate alleles, suboverlap marks) byte for
    byte, and the chosen maximum alignment.  Option sets cycle through oligo anchors, smoothing off, small maxIndelSize / large flank."""
    sb, lnp = specgen.score_indels_case(case)
    # the generator's Python restatement of CandidateAlignment::operator< must agree with the reference's std::set
    assert np.array_equal(reflib.ref_candidate_alignment_order(sb), np.arange(sb.n_alns))
    o_recs, o_n, o_max, _ = reflib.ox_score_indels(sb, lnp)
    r_recs, r_n, r_max = reflib.ref_score_indels(sb, lnp)
    assert np.array_equal(o_n, r_n)
    assert np.array_equal(o_max, r_max)
    assert o_recs.tobytes() == r_recs.tobytes()


def test_score_indels_on_k1_scores_matches_the_reference():
    """K1 oracle scores of haplotype-shaped loci fed to K6: oracle vs reference (real score ties between equivalent alignments)."""
    rng = np.random.default_rng(77)
    regions = [specgen.simple_region(rng, n_reads=12) for _ in range(12)]
    # the reference keeps a std::set: drop duplicate alignments of a read (two identical haplotypes) and order them its way
    for rg in regions:
        seen, keep = set(), []
        for cal in rg.alns:
            key = (cal.read, cal.pos, tuple(cal.path), tuple((k.pos, k.delete_length, k.insert_seq) for k in cal.indels))
            if key not in seen and cal.trailing < 0 and cal.leading < 0:
                seen.add(key)
                keep.append(cal)
        rg.alns = keep
    sb = B.score_indels_batch_from_regions(regions)
    perm = reflib.ref_candidate_alignment_order(sb)
    ab = B.build_align_batch(regions)
    assert ab.n_alns == sb.n_alns
    lnp = reflib.ox_score(ab)
    # reorder each read's alignments (and their scores) into std::set order
    flat = [a for rg in regions for a in rg.alns]
    by_region, pos = [], 0
    for rg in regions:
        n = len(rg.alns)
        rg.alns = [flat[int(i)] for i in perm[pos : pos + n]]
        pos += n
    sb = B.score_indels_batch_from_regions(regions)
    lnp = np.concatenate([lnp[perm], [0.0]])
    assert np.array_equal(reflib.ref_candidate_alignment_order(sb), np.arange(sb.n_alns))
    o_recs, o_n, o_max, _ = reflib.ox_score_indels(sb, lnp)
    r_recs, r_n, r_max = reflib.ref_score_indels(sb, lnp)
    assert np.array_equal(o_n, r_n) and np.array_equal(o_max, r_max) and o_recs.tobytes() == r_recs.tobytes()
    assert int((o_recs["flags"] & A.SX_RIS_SCORED).sum()) > 100


def test_score_indels_bench_workload_matches_the_reference():
    """The K6 leg of bench.py: its synthetic batch describes the same alignments as the K1 batch of the same loci (so K1's lnp[a] is
    K6's score of alignment a), and on K1's scores the oracle and the reference's score_indels produce the same records.  The
    synthetic workload lists one alignment per haplotype in haplotype order; the harness puts them into the reference's std::set
    (ordering them, dropping exact duplicates), which does not change any record."""
    import ctypes as C
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    class _Alloc:
        def array(self, nbytes, dt):
            return np.zeros(nbytes // np.dtype(dt).itemsize + 1, dt)

    synth = bench.load_synth()
    for n_loci, depth, n_haps, rpr in ((400, 30, 4, 0), (12, 300, 32, 16)):  # cfg2- and cfg5-shaped
        sb = bench.make_score_indels_workload(synth, n_loci, depth, 150, n_haps, 5, 4, rpr)
        ab, _pb, _gb = bench.make_workload(synth, _Alloc(), n_loci, depth, 150, n_haps, 5, 4, 8, rpr, 0)
        assert sb.n_alns == ab.n_alns and sb.n_reads == ab.n_reads and sb.n_regions == ab.n_regions
        assert np.array_equal(ab.alns["ref_pos"][: ab.n_alns], sb.aln_pos[: sb.n_alns])
        assert np.array_equal(ab.regions["read_begin"][: ab.n_regions + 1], sb.region_read_off)
        lnp = np.concatenate([reflib.ox_score(ab), [0.0]])
        o_recs, o_n, _o_max, _ = reflib.ox_score_indels(sb, lnp)
        out = B.ScoreIndelsOut(sb)
        err = C.create_string_buffer(512)
        fn = reflib.ref().ref_score_indels_ex
        fn.argtypes = [C.POINTER(A.SxScoreIndelsBatch)] + [C.c_void_p] * 6 + [C.c_int, C.c_char_p, C.c_int]
        rc = fn(C.byref(sb.c), A.ptr(lnp), A.ptr(sb.ins_pool), A.ptr(sb.ins_off), A.ptr(out.recs), A.ptr(out.n_rec), A.ptr(out.max_aln), 1, err, 512)
        assert rc == 0, err.value
        r_recs, r_n, _r_max, _ = out.compact()
        assert np.array_equal(o_n, r_n) and o_recs.tobytes() == r_recs.tobytes()
        assert int((o_recs["flags"] & A.SX_RIS_SCORED).sum()) > 0.5 * sb.n_reads


# ------------------------------------------------------------------------------------------------------------------------------
# K7 enumerate_alignments: oracle/enumerate_oracle.cpp against the reference's getCandidateAlignments (oracle/ref_harness_enumerate.inc)
# ------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("chunk", range(12))
def test_enumerate_alignments_matches_the_reference(chunk):
    """every candidate alignment in std::set order -- position, path, indel keys, leading / trailing edge keys -- and the per-read
    status (warn flags, blt_exception): 25 seeded batches per chunk (plain / clustered / phased two-sample / dense / tight toggle
    budget / hard-clipped windows)."""
    total = 0
    for case in range(25 * chunk, 25 * (chunk + 1)):
        eb = specgen.enum_case(case)
        cap = eb.n_reads * 6000 + 64
        want = reflib.ref_enumerate_alignments(eb, cap_alns=cap)
        got = reflib.ox_enumerate_alignments(eb, cap_alns=cap, limits=False)
        assert got.rc == 0
        for x, y in zip(want.trimmed(), got.trimmed()):
            assert x.tobytes() == y.tobytes(), case
        total += int(want.totals[0])
    assert total > 4000


@pytest.mark.parametrize("chunk", range(4))
def test_choose_realignment_matches_the_reference(chunk):
    """oracle/realign_oracle.cpp (K9's checker) against the reference's own scoreCandidateAlignments: rseg.realignment of every read, for
    the default smoothing range, smoothing off and two wide ranges; tidy and awkward batches."""
    n = clipped = 0
    for case in range(20 * chunk, 20 * (chunk + 1)):
        eb = specgen.enum_edge_case(case) if case % 3 == 0 else specgen.enum_case(case)
        out = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)
        quals = specgen.realign_quals(eb, case)
        for smooth, rng_ in ((True, 2.302585092994046), (False, 1.0), (True, 30.0), (True, 500.0)):
            lnp, want = reflib.ref_choose_realignment(eb, out, quals, is_smoothed=smooth, smoothed_range=rng_)
            ox = reflib.ox_choose_realignment(B.RealignBatch(eb, out, is_smoothed=smooth, smoothed_lnp_range=rng_), np.concatenate([lnp, [0.0]]))
            assert ox.rc == 0
            for r in range(eb.n_reads):
                g = ox.realignment_of(r)
                assert g == want[r], (case, r)
                n += 1
                clipped += bool(g and "S" in g[1])
    assert n > 1000 and clipped > 20
