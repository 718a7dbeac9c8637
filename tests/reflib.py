"""Loaders for the two CPU checkers (TEST INFRASTRUCTURE):
   oracle/liboracle.so          -- our restatement (travels as source, built by __graft_entry__.build())
   oracle/_ref/libstrelka_ref.so -- the reference's own code behind a C shim (built here from /root/reference)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from strelka_b200 import _abi as A
from strelka_b200 import batch as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libstrelka_ref.so")

_P = C.c_void_p
_oracle = None
_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
        lib = C.CDLL(ORACLE_SO)
        lib.ox_score_alignments.argtypes = [C.POINTER(A.SxAlignBatch), _P]
        lib.ox_score_alignments_range.argtypes = [C.POINTER(A.SxAlignBatch), C.c_uint32, C.c_uint32, _P]
        lib.ox_global_align.argtypes = [C.POINTER(A.SxGaScores), C.POINTER(A.SxGaBatch), _P, _P]
        lib.ox_site_gl_germline.argtypes = [C.POINTER(A.SxParams), C.POINTER(A.SxPileupBatch), C.c_int, _P]
        lib.ox_site_gl_germline_range.argtypes = [C.POINTER(A.SxParams), C.POINTER(A.SxPileupBatch), C.c_int, C.c_uint32, C.c_uint32, _P]
        lib.ox_dependent_eprob.argtypes = [C.POINTER(A.SxParams), C.POINTER(A.SxPileupBatch), _P, _P]
        lib.ox_site_gl_somatic.argtypes = [C.POINTER(A.SxParams), C.POINTER(A.SxPileupBatch), C.POINTER(A.SxPileupBatch), _P, _P]
        lib.ox_site_gl_somatic_range.argtypes = [C.POINTER(A.SxParams), C.POINTER(A.SxPileupBatch), C.POINTER(A.SxPileupBatch), _P, C.c_uint32, C.c_uint32, _P]
        lib.ox_logf_restated.argtypes = [C.c_float]
        lib.ox_logf_restated.restype = C.c_float
        lib.ox_powf_restated.argtypes = [C.c_float, C.c_float]
        lib.ox_powf_restated.restype = C.c_float
        lib.ox_sort_restated.argtypes = [_P, C.c_uint32, _P]
        lib.ox_sort_std.argtypes = [_P, C.c_uint32, _P]
        _oracle = lib
    return _oracle


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.ref_getLogSum_float.argtypes = [C.c_float, C.c_float]
        lib.ref_getLogSum_float.restype = C.c_float
        _ref = lib
    return _ref


# ---------------------------------------------------------------------------------------------------------------------
# oracle wrappers
# ---------------------------------------------------------------------------------------------------------------------
def ox_score(batch: B.AlignBatch) -> np.ndarray:
    out = np.zeros(batch.n_alns, np.float64)
    rc = oracle().ox_score_alignments(C.byref(batch.c), A.ptr(out))
    assert rc == 0, rc
    return out


def ox_global_align(scores: A.SxGaScores, gb: B.GaBatch):
    res = np.zeros(gb.n, A.GA_RESULT_DT)
    cig = np.zeros((gb.n, gb.max_ops), np.uint32)
    rc = oracle().ox_global_align(C.byref(scores), C.byref(gb.c), A.ptr(res), A.ptr(cig))
    assert rc == 0, rc
    return res, cig


def ox_germline(params: A.SxParams, pb: B.PileupBatch, is_always_test: bool) -> np.ndarray:
    out = np.zeros(pb.n_sites, A.DIGT_RESULT_DT)
    rc = oracle().ox_site_gl_germline(C.byref(params), C.byref(pb.c), int(is_always_test), A.ptr(out))
    assert rc == 0, rc
    return out


def ox_dependent_eprob(params: A.SxParams, pb: B.PileupBatch):
    off = np.zeros(pb.n_sites + 1, np.uint32)
    de = np.zeros(max(1, pb.n_calls), np.float32)
    rc = oracle().ox_dependent_eprob(C.byref(params), C.byref(pb.c), A.ptr(off), A.ptr(de))
    assert rc == 0, rc
    return off, de[: off[-1]]


def ox_somatic(params: A.SxParams, npb: B.PileupBatch, tpb: B.PileupBatch, forced=None) -> np.ndarray:
    out = np.zeros(npb.n_sites, A.SSNV_RESULT_DT)
    f = None if forced is None else np.ascontiguousarray(forced, np.uint8)
    rc = oracle().ox_site_gl_somatic(C.byref(params), C.byref(npb.c), C.byref(tpb.c), A.ptr(f), A.ptr(out))
    assert rc == 0, rc
    return out


# ---------------------------------------------------------------------------------------------------------------------
# reference wrappers
# ---------------------------------------------------------------------------------------------------------------------
def _err():
    return C.create_string_buffer(1024)


def ref_score_region(r: B.RegionSpec) -> np.ndarray:
    """scoreCandidateAlignment (reference code) on every candidate alignment of one region."""
    read_off = np.zeros(len(r.reads) + 1, np.int32)
    read_off[1:] = np.cumsum([len(c) for c, _ in r.reads])
    codes = np.concatenate([np.asarray(c, np.uint8) for c, _ in r.reads]) if r.reads else np.zeros(0, np.uint8)
    quals = np.concatenate([np.asarray(q, np.uint8) for _, q in r.reads]) if r.reads else np.zeros(0, np.uint8)
    n = len(r.alns)
    aln_read = np.array([a.read for a in r.alns], np.int32)
    aln_pos = np.array([a.pos for a in r.alns], np.int32)
    path_off = np.zeros(n + 1, np.int32)
    path_off[1:] = np.cumsum([len(a.path) for a in r.alns])
    path_type = "".join(t for a in r.alns for t, _ in a.path).encode()
    path_len = np.array([l for a in r.alns for _, l in a.path], np.int32)
    indel_off = np.zeros(n + 1, np.int32)
    indel_off[1:] = np.cumsum([len(a.indels) for a in r.alns])
    keys = [k for a in r.alns for k in a.indels]
    ipos = np.array([k.pos for k in keys], np.int32)
    ityp = np.array([k.type for k in keys], np.int32)
    idel = np.array([k.delete_length for k in keys], np.int32)
    iins_off = np.zeros(len(keys) + 1, np.int32)
    iins_off[1:] = np.cumsum([len(k.insert_seq) for k in keys])
    ins_pool = "".join(k.insert_seq for k in keys).encode()
    icand = np.array([1 if k.is_candidate else 0 for k in keys], np.uint8)
    lead = np.array([a.leading for a in r.alns], np.int32)
    trail = np.array([a.trailing for a in r.alns], np.int32)
    out = np.zeros(n, np.float64)
    err = _err()
    rc = ref().ref_score_alignments(
        r.ref.encode(), len(r.ref), r.ref_begin, len(r.reads), _P(A.ptr(codes)), _P(A.ptr(quals)), _P(A.ptr(read_off)), n,
        _P(A.ptr(aln_read)), _P(A.ptr(aln_pos)), _P(A.ptr(path_off)), path_type, _P(A.ptr(path_len)), _P(A.ptr(indel_off)), _P(A.ptr(ipos)),
        _P(A.ptr(ityp)), _P(A.ptr(idel)), _P(A.ptr(iins_off)), ins_pool, _P(A.ptr(icand)), _P(A.ptr(lead)), _P(A.ptr(trail)), _P(A.ptr(out)), err, 1024,
    )
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return out


def ref_global_align(scores: A.SxGaScores, gb: B.GaBatch, use_short: bool = False):
    res = np.zeros(gb.n, A.GA_RESULT_DT)
    cig = np.zeros((gb.n, gb.max_ops), np.uint32)
    err = _err()
    rc = ref().ref_global_align(C.byref(scores), C.byref(gb.c), int(use_short), _P(A.ptr(res)), _P(A.ptr(cig)), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return res, cig


def ref_germline(params: A.SxParams, pb: B.PileupBatch, is_always_test: bool) -> np.ndarray:
    out = np.zeros(pb.n_sites, A.DIGT_RESULT_DT)
    err = _err()
    rc = ref().ref_site_gl_germline(C.byref(params), C.byref(pb.c), int(is_always_test), _P(A.ptr(out)), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return out


def ref_dependent_eprob(params: A.SxParams, pb: B.PileupBatch):
    off = np.zeros(pb.n_sites + 1, np.uint32)
    de = np.zeros(max(1, pb.n_calls), np.float32)
    err = _err()
    rc = ref().ref_dependent_eprob(C.byref(params), C.byref(pb.c), _P(A.ptr(off)), _P(A.ptr(de)), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return off, de[: off[-1]]


def ref_somatic(params: A.SxParams, npb: B.PileupBatch, tpb: B.PileupBatch, forced=None) -> np.ndarray:
    out = np.zeros(npb.n_sites, A.SSNV_RESULT_DT)
    f = None if forced is None else np.ascontiguousarray(forced, np.uint8)
    err = _err()
    rc = ref().ref_site_gl_somatic(C.byref(params), C.byref(npb.c), C.byref(tpb.c), _P(A.ptr(f)), _P(A.ptr(out)), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return out


def ox_indel_gl(params: A.SxParams, ib: B.IndelBatch) -> np.ndarray:
    out = np.zeros(ib.n_loci, A.INDEL_RESULT_DT)
    lib = oracle()
    lib.ox_indel_gl.argtypes = [C.POINTER(A.SxParams), C.POINTER(A.SxIndelBatch), _P]
    rc = lib.ox_indel_gl(C.byref(params), C.byref(ib.c), A.ptr(out))
    assert rc == 0, rc
    return out


def ref_indel_gl(params: A.SxParams, ib: B.IndelBatch) -> np.ndarray:
    out = np.zeros(ib.n_loci, A.INDEL_RESULT_DT)
    err = _err()
    rc = ref().ref_indel_gl(C.byref(params), C.byref(ib.c), _P(A.ptr(out)), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return out


def ox_pileup_reads(pb: "B.PileupReadsBatch"):
    out = B.PileupColumns(pb)
    lib = oracle()
    lib.ox_pileup_reads.argtypes = [C.POINTER(A.SxPileupReadsBatch), _P, _P, C.c_uint64, _P, _P, C.c_uint64, _P, _P]
    rc = lib.ox_pileup_reads(C.byref(pb.c), A.ptr(out.site_off), A.ptr(out.calls), out.calls.size, A.ptr(out.t2_off), A.ptr(out.t2_calls), out.t2_calls.size,
                             A.ptr(out.n_spandel), A.ptr(out.n_submapped))
    assert rc == 0, rc
    return out.trimmed()


def ref_pileup_reads(pb: "B.PileupReadsBatch"):
    """The reference's own starling_pos_processor_base::pileup_read_segment, driven read by read (oracle/ref_harness.cpp)."""
    out = B.PileupColumns(pb)
    err = _err()
    fn = ref().ref_pileup_reads
    fn.argtypes = [C.POINTER(A.SxPileupReadsBatch), _P, _P, C.c_uint64, _P, _P, C.c_uint64, _P, _P, C.c_char_p, C.c_int]
    rc = fn(C.byref(pb.c), A.ptr(out.site_off), A.ptr(out.calls), out.calls.size, A.ptr(out.t2_off), A.ptr(out.t2_calls), out.t2_calls.size,
            A.ptr(out.n_spandel), A.ptr(out.n_submapped), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return out.trimmed()


def ox_score_indels(sb: "B.ScoreIndelsBatch", lnp: np.ndarray):
    out = B.ScoreIndelsOut(sb)
    lib = oracle()
    lib.ox_score_indels.argtypes = [C.POINTER(A.SxScoreIndelsBatch), _P, _P, _P, _P, _P]
    rc = lib.ox_score_indels(C.byref(sb.c), A.ptr(lnp), A.ptr(out.recs), A.ptr(out.n_rec), A.ptr(out.max_aln), A.ptr(out.eval_aln))
    assert rc == 0, rc
    return out.compact()


def ref_candidate_alignment_order(sb: "B.ScoreIndelsBatch") -> np.ndarray:
    """Iteration order of the reference's std::set<CandidateAlignment> over each read's alignments (as alignment indices)."""
    perm = np.zeros(sb.n_alns + 1, np.uint32)
    err = _err()
    fn = ref().ref_candidate_alignment_order
    fn.argtypes = [C.POINTER(A.SxScoreIndelsBatch), _P, _P, _P, C.c_char_p, C.c_int]
    rc = fn(C.byref(sb.c), A.ptr(sb.ins_pool), A.ptr(sb.ins_off), A.ptr(perm), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return perm[: sb.n_alns]


def ref_score_indels(sb: "B.ScoreIndelsBatch", lnp: np.ndarray):
    """The reference's own score_indels on rebuilt IndelBuffer / read_segment / CandidateAlignment objects
    (oracle/ref_harness_score_indels.cpp); returns (records, n_rec, max_aln)."""
    out = B.ScoreIndelsOut(sb)
    err = _err()
    fn = ref().ref_score_indels
    fn.argtypes = [C.POINTER(A.SxScoreIndelsBatch), _P, _P, _P, _P, _P, _P, C.c_char_p, C.c_int]
    rc = fn(C.byref(sb.c), A.ptr(lnp), A.ptr(sb.ins_pool), A.ptr(sb.ins_off), A.ptr(out.recs), A.ptr(out.n_rec), A.ptr(out.max_aln), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    recs, n_rec, max_aln, _ = out.compact()
    return recs, n_rec, max_aln


def ref_max_toggle_table(max_alignment_count: int) -> np.ndarray:
    """starling_align_limit(max_alignment_count).get_max_toggle(i), i < 100, from the reference."""
    out = np.zeros(100, np.uint8)
    fn = ref().ref_max_toggle_table
    fn.argtypes = [C.c_uint, _P]
    fn(max_alignment_count, A.ptr(out))
    return out


def ref_enumerate_alignments(eb: "B.EnumBatch", cap_alns=None) -> "B.EnumOut":
    """The reference's own getCandidateAlignments on rebuilt IndelBuffer / read_segment objects (oracle/ref_harness_enumerate.inc)."""
    out = B.EnumOut(eb, cap_alns)
    err = _err()
    fn = ref().ref_enumerate_alignments
    fn.argtypes = [C.POINTER(A.SxEnumBatch), _P, _P, _P, _P, _P, _P, _P, C.POINTER(A.SxEnumOut), C.c_char_p, C.c_int]
    rc = fn(C.byref(eb.c), A.ptr(eb.ins_pool), A.ptr(eb.ins_off), A.ptr(eb.ref_pool), A.ptr(eb.ref_off), A.ptr(eb.ref_begin), A.ptr(eb.read_pool), A.ptr(eb.read_off),
            C.byref(out.c), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return out


_k7core = None


def k7core_enumerate(eb: "B.EnumBatch", max_alns: int = 0, cap_alns=None, cap_segs=None, cap_keys=None, fast: bool = False):
    """strelka_b200/csrc/k7_core.cuh compiled for the host (tests/cpp/k7_core_host.cpp): the device body run read by read on the CPU.
    Returns (rc, EnumOut)."""
    global _k7core
    if _k7core is None:
        import tempfile

        so = os.path.join(tempfile.mkdtemp(prefix="k7core"), "libk7core.so")
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "strelka_b200", "csrc"),
                               os.path.join(ROOT, "tests", "cpp", "k7_core_host.cpp"), "-o", so])
        _k7core = C.CDLL(so)
        _k7core.k7core_run.argtypes = [C.POINTER(A.SxEnumBatch), C.POINTER(A.SxEnumOut), C.c_uint32]
        _k7core.k7core_run_fast.argtypes = [C.POINTER(A.SxEnumBatch), C.POINTER(A.SxEnumOut), C.c_uint32, _P]
        _k7core.k7core_make_start_pos.argtypes = [_P, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32, _P, _P, _P, _P, _P]
        _k7core.k7core_end_pin_start_pos.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, _P, _P]
    if eb is None:
        return _k7core
    out = B.EnumOut(eb, cap_alns, cap_segs, cap_keys)
    if fast:  # the SX_ENUM_F_FAST launch plan (two scratch tiers, one search, log + gather)
        retried = np.zeros(1, np.uint32)
        rc = _k7core.k7core_run_fast(C.byref(eb.c), C.byref(out.c), max_alns, A.ptr(retried))
        out.n_retried = int(retried[0])
        return rc, out
    rc = _k7core.k7core_run(C.byref(eb.c), C.byref(out.c), max_alns)
    return rc, out


def ox_enumerate_alignments(eb: "B.EnumBatch", cap_alns=None, cap_segs=None, cap_keys=None, limits: bool = True) -> "B.EnumOut":
    """oracle/enumerate_oracle.cpp: the CPU restatement of getCandidateAlignments; limits=True applies the device build's per-read
    capacities (SX_ENUM_ST_LIMIT), False is the reference's unlimited behaviour."""
    out = B.EnumOut(eb, cap_alns, cap_segs, cap_keys)
    lib = oracle()
    lib.ox_enumerate_alignments.argtypes = [C.POINTER(A.SxEnumBatch), C.POINTER(A.SxEnumOut), C.c_int]
    out.rc = lib.ox_enumerate_alignments(C.byref(eb.c), C.byref(out.c), 1 if limits else 0)
    assert out.rc in (0, A.SX_ERR_CAPACITY), out.rc
    return out


_k8core = None


def k8core_link(eb: "B.EnumBatch", out: "B.EnumOut", regions: np.ndarray, cap_segs=None, cap_ins=None):
    """strelka_b200/csrc/k8_core.cuh compiled for the host (tests/cpp/k8_core_host.cpp): K7b's device body on the CPU.  (rc, LinkOut)"""
    global _k8core
    if _k8core is None:
        import tempfile

        so = os.path.join(tempfile.mkdtemp(prefix="k8core"), "libk8core.so")
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "strelka_b200", "csrc"),
                               os.path.join(ROOT, "tests", "cpp", "k8_core_host.cpp"), "-o", so])
        _k8core = C.CDLL(so)
        _k8core.k8core_run.argtypes = [C.POINTER(A.SxEnumBatch), C.POINTER(A.SxEnumOut), C.c_uint32, _P, _P, C.POINTER(A.SxLinkOut)]
    n_alns = int(out.totals[0])
    lo = B.LinkOut(regions, n_alns, cap_segs if cap_segs is not None else 2 * int(out.totals[1]) + 8 * eb.n_regions + 64,
                   cap_ins if cap_ins is not None else 64 * n_alns + 16 * eb.n_regions + 64, n_enum_segs=int(out.totals[1]))
    rc = _k8core.k8core_run(C.byref(eb.c), C.byref(out.c), n_alns, A.ptr(eb.ins_off), A.ptr(eb.ins_pool), C.byref(lo.c))
    return rc, lo


def ref_alignment_indels(eb: "B.EnumBatch"):
    """The reference's getCandidateAlignment + getAlignmentIndels(includeMismatches) for every read, as window indices
    (oracle/ref_harness_enumerate.inc): (in_key_off, in_keys, lead, trail)."""
    po = B.PrepOut(eb, 64 * eb.n_reads + 64)
    err = _err()
    fn = ref().ref_alignment_indels
    fn.argtypes = [C.POINTER(A.SxEnumBatch), _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_uint32, _P, _P, C.c_char_p, C.c_int]
    rc = fn(C.byref(eb.c), A.ptr(eb.ins_pool), A.ptr(eb.ins_off), A.ptr(eb.ref_pool), A.ptr(eb.ref_off), A.ptr(eb.ref_begin), A.ptr(eb.read_pool), A.ptr(eb.read_off),
            A.ptr(po.in_key_off), A.ptr(po.in_keys), po.cap_keys, A.ptr(po.in_lead_key), A.ptr(po.in_trail_key), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    po.totals[0] = po.in_key_off[eb.n_reads]
    return po


_k7acore = None


def k7acore_prepare(eb: "B.EnumBatch", pools: "B.AlignBatch", cap_keys=None):
    """strelka_b200/csrc/k7a_core.cuh compiled for the host (tests/cpp/k7a_core_host.cpp): K7a's device body on the CPU.  (rc, PrepOut)"""
    global _k7acore
    if _k7acore is None:
        import tempfile

        so = os.path.join(tempfile.mkdtemp(prefix="k7acore"), "libk7acore.so")
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "strelka_b200", "csrc"),
                               os.path.join(ROOT, "tests", "cpp", "k7a_core_host.cpp"), "-o", so])
        _k7acore = C.CDLL(so)
        _k7acore.k7acore_run.argtypes = [C.POINTER(A.SxEnumBatch), _P, _P, _P, _P, _P, C.POINTER(A.SxPrepOut)]
    po = B.PrepOut(eb, cap_keys)
    rc = _k7acore.k7acore_run(C.byref(eb.c), A.ptr(pools.regions), A.ptr(pools.seq4), A.ptr(pools.ref), A.ptr(eb.ins_off), A.ptr(eb.ins_pool), C.byref(po.c))
    return rc, po


def ref_choose_realignment(eb: "B.EnumBatch", out: "B.EnumOut", quals: np.ndarray, is_smoothed=True, smoothed_range=2.302585092994046, max_segs=48):
    """The reference's own scoreCandidateAlignments on rebuilt objects (oracle/ref_harness_enumerate.inc): (lnp[n_alns], realignment per
    read as (pos, cigar) or None)."""
    n, nA = eb.n_reads, int(out.totals[0])
    lnp = np.zeros(nA + 1, np.float64)
    pos, nseg, real = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.uint16), np.zeros(n + 1, np.uint8)
    segs = np.zeros((n + 1) * max_segs, dtype=A.ALN_SEG_DT)
    err = _err()
    fn = ref().ref_choose_realignment
    fn.argtypes = [C.POINTER(A.SxEnumBatch), C.POINTER(A.SxEnumOut)] + [_P] * 8 + [C.c_int, C.c_double] + [_P] * 4 + [C.c_uint32, _P, C.c_char_p, C.c_int]
    rc = fn(C.byref(eb.c), C.byref(out.c), A.ptr(eb.ins_pool), A.ptr(eb.ins_off), A.ptr(eb.ref_pool), A.ptr(eb.ref_off), A.ptr(eb.ref_begin), A.ptr(eb.read_pool),
            A.ptr(eb.read_off), A.ptr(quals), 1 if is_smoothed else 0, smoothed_range, A.ptr(lnp), A.ptr(pos), A.ptr(nseg), A.ptr(segs), max_segs, A.ptr(real), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    res = []
    for r in range(n):
        if not real[r]:
            res.append(None)
            continue
        row = segs[r * max_segs : r * max_segs + int(nseg[r])]
        res.append((int(pos[r]), "".join(f"{int(s['len'])}{B.AP_CHAR[int(s['kind'])]}" for s in row)))
    return lnp[:nA], res


_k9core = None


def k9core_choose(rb: "B.RealignBatch", lnp: np.ndarray, cap_segs=None):
    """strelka_b200/csrc/k9_core.cuh compiled for the host (tests/cpp/k9_core_host.cpp).  (rc, RealignOut)"""
    global _k9core
    if _k9core is None:
        import tempfile

        so = os.path.join(tempfile.mkdtemp(prefix="k9core"), "libk9core.so")
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "strelka_b200", "csrc"),
                               os.path.join(ROOT, "tests", "cpp", "k9_core_host.cpp"), "-o", so])
        _k9core = C.CDLL(so)
        _k9core.k9core_run.argtypes = [C.POINTER(A.SxRealignBatch), _P, C.POINTER(A.SxRealignOut)]
    lnp = np.ascontiguousarray(lnp, dtype=np.float64)
    ro = B.RealignOut(rb, cap_segs)
    rc = _k9core.k9core_run(C.byref(rb.c), A.ptr(lnp), C.byref(ro.c))
    return rc, ro


def ox_choose_realignment(rb: "B.RealignBatch", lnp: np.ndarray, cap_segs=None) -> "B.RealignOut":
    """oracle/realign_oracle.cpp: the CPU restatement of the tail of scoreCandidateAlignments + finishRealignment + the pool clipper."""
    lnp = np.ascontiguousarray(lnp, dtype=np.float64)
    ro = B.RealignOut(rb, cap_segs)
    lib = oracle()
    lib.ox_choose_realignment.argtypes = [C.POINTER(A.SxRealignBatch), _P, C.POINTER(A.SxRealignOut)]
    ro.rc = lib.ox_choose_realignment(C.byref(rb.c), A.ptr(lnp), C.byref(ro.c))
    return ro


def ref_realign_gates(gb: "B.GateBatch", max_segs=64):
    """The reference's own is_realignable / check_for_candidate_indel_overlap / normalizeInputAlignmentIndels / matchify_edge_soft_clip per read
    (oracle/ref_harness_enumerate.inc): (gate[n_reads], [(pos, cigar) or None])."""
    eb = gb.eb
    n = eb.n_reads
    gate, pos, nseg = np.zeros(n + 1, np.uint8), np.zeros(n + 1, np.int32), np.zeros(n + 1, np.uint16)
    segs = np.zeros((n + 1) * max_segs, dtype=A.ALN_SEG_DT)
    err = _err()
    fn = ref().ref_realign_gates
    fn.argtypes = [C.POINTER(A.SxGateBatch)] + [_P] * 8 + [C.c_uint32, C.c_char_p, C.c_int]
    rc = fn(C.byref(gb.c), A.ptr(eb.ins_pool), A.ptr(eb.ins_off), A.ptr(eb.read_pool), A.ptr(eb.read_off), A.ptr(gate), A.ptr(pos), A.ptr(nseg), A.ptr(segs), max_segs, err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    res = []
    for r in range(n):
        if not (int(gate[r]) & A.SX_GATE_REALIGN):
            res.append(None)
            continue
        row = segs[r * max_segs : r * max_segs + int(nseg[r])]
        res.append((int(pos[r]), "".join(f"{int(s['len'])}{B.AP_CHAR[int(s['kind'])]}" for s in row)))
    return gate[:n], res


def ref_realign_and_score_read(gb: "B.GateBatch", quals: np.ndarray, retain_soft=False, is_smoothed=True, smoothed_range=2.302585092994046, max_segs=64,
                               full_window=True, read_flags=None, rec_off=None):
    """The reference's whole realignAndScoreRead per read (oracle/ref_harness_enumerate.inc), mapper alignments in: (status[n_reads] --
    0 not realigned, 1 realigned, 2 the reference threw --, [(pos, cigar) or None]).  full_window: the IndelBuffer is rebuilt with everything
    the batch says about its window (observation sets from use_keys, notDiscoveredFromReads, forced output, phasing rows, search options), as
    ref_enumerate_alignments does -- what the chain sees; False: candidacy only (the harness of round 1).  rec_off: also return what
    score_indels left in the indel buffer, (recs, n_rec) with read r's records at recs[rec_off[r] : rec_off[r] + n_rec[r]]."""
    eb = gb.eb
    n = eb.n_reads
    status, pos, nseg = np.zeros(n + 1, np.uint8), np.zeros(n + 1, np.int32), np.zeros(n + 1, np.uint16)
    segs = np.zeros((n + 1) * max_segs, dtype=A.ALN_SEG_DT)
    err = _err()
    fn = ref().ref_realign_and_score_read_ex
    fn.argtypes = [C.POINTER(A.SxGateBatch), C.POINTER(A.SxEnumBatch)] + [_P] * 10 + [C.c_int, C.c_int, C.c_double] + [_P] * 4 + [C.c_uint32] + [_P] * 4 + [C.c_char_p, C.c_int]
    recs = n_rec = None
    if rec_off is not None:
        recs, n_rec = np.zeros(int(rec_off[n]) + 1, A.READ_INDEL_SCORE_DT), np.zeros(n + 1, np.uint32)
    secs = C.c_double(0.0)
    rc = fn(C.byref(gb.c), C.byref(eb.c) if full_window else None, A.ptr(eb.ins_pool), A.ptr(eb.ins_off), A.ptr(eb.ref_pool), A.ptr(eb.ref_off), None, A.ptr(eb.ref_begin),
            A.ptr(eb.read_pool), A.ptr(eb.read_off), A.ptr(quals), A.ptr(read_flags), 1 if retain_soft else 0, 1 if is_smoothed else 0, smoothed_range, A.ptr(status),
            A.ptr(pos), A.ptr(nseg), A.ptr(segs), max_segs, A.ptr(rec_off), A.ptr(recs), A.ptr(n_rec), C.addressof(secs), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    res = []
    for r in range(n):
        if status[r] != 1:
            res.append(None)
            continue
        row = segs[r * max_segs : r * max_segs + int(nseg[r])]
        res.append((int(pos[r]), "".join(f"{int(s['len'])}{B.AP_CHAR[int(s['kind'])]}" for s in row)))
    if rec_off is not None:
        return status[:n], res, recs, n_rec[:n]
    return status[:n], res


def k7gcore_gates(gb: "B.GateBatch"):
    """k7g_read (strelka_b200/csrc/k7a_core.cuh) compiled for the host."""
    k7acore_prepare(gb.eb, B.read_pools_of(gb.eb)) if _k7acore is None else None
    _k7acore.k7gcore_run.argtypes = [C.POINTER(A.SxGateBatch), C.POINTER(A.SxGateOut)]
    go = B.GateOut(gb)
    rc = _k7acore.k7gcore_run(C.byref(gb.c), C.byref(go.c))
    return rc, go
