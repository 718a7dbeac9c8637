"""GPU parity of K7's second launch plan (sx_enum_opts.flags = SX_ENUM_F_FAST: local-memory scratch tier + arena tier, one search
per read, log + gather).  Written from the first ncu capture of the original plan after the round's GPU minutes were spent, so it has
only run as host-compiled device code: this file sorts last and carries a timeout for the same reason as test_zz_gpu_enumerate.py."""
import numpy as np
import pytest

import reflib
import specgen
from strelka_b200 import _abi as A
from strelka_b200 import batch as B

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]
GOLD_NAMES = ("aln_off", "status", "aln_pos", "aln_seg_off", "segs", "aln_key_off", "aln_keys", "lead", "trail")


@pytest.fixture(scope="module")
def ctx():
    from strelka_b200.api import Context

    c = Context(0)
    yield c
    c.close()


def _fast(eb, max_alns=None):
    eb.opts.flags = 0  # the first launch plan (the fast plan is the default and runs in every other test)
    if max_alns:
        eb.opts.max_alns_per_read = max_alns
    eb.c.opts = eb.opts
    return eb


def _same(want, got):
    for name, x, y in zip(GOLD_NAMES, want.trimmed(), got.trimmed()):
        assert x.tobytes() == y.tobytes(), name


@pytest.mark.parametrize("case", range(24))
def test_k7_fast_plan(ctx, case):
    eb = _fast(specgen.enum_case(case), 6000 if case % 2 else None)
    cap = eb.n_reads * (6000 if case % 2 else 64) + 64
    got = ctx.enumerate_alignments(eb, cap_alns=cap)
    assert ctx.timing().launches == 6  # the two-pass plan: count, 3 scan kernels, write (+ the frame count)
    _same(reflib.ox_enumerate_alignments(eb, cap_alns=cap), got)


def test_k7_fast_plan_many_regions_device_resident(ctx):
    from strelka_b200.api import DevEnumBatch

    rng = np.random.default_rng(78)
    regions = [specgen.random_enum_region(rng, n_reads=int(rng.integers(1, 9)), cluster=bool(i % 3 == 0), n_keys=(1, 6)) for i in range(2500)]
    eb = _fast(B.EnumBatch(regions))
    want = reflib.ox_enumerate_alignments(eb, cap_alns=eb.n_reads * 64 + 64)
    db = DevEnumBatch(ctx, eb, cap_alns=eb.n_reads * 64 + 64)
    ctx.enumerate_alignments_dev(db)
    _same(want, db.download())
    ctx.enumerate_alignments_dev(db)  # a second run on the same context: the log cursor is reset, the result identical
    _same(want, db.download())


def test_k7_fast_plan_capacity_error(ctx):
    from strelka_b200.api import SxError

    eb = _fast(specgen.enum_case(1))
    full = ctx.enumerate_alignments(eb)
    with pytest.raises(SxError) as e:
        ctx.enumerate_alignments(eb, cap_alns=int(full.totals[0]) - 1)
    assert e.value.code == A.SX_ERR_CAPACITY
    _same(full, ctx.enumerate_alignments(eb))


@pytest.mark.parametrize("fast", [False, True])
def test_k7_full_size_properties(ctx, fast):
    """K7 at BASELINE.json's cfg2 size (1M candidate loci x 30 reads = 30M reads -> ~3.2e8 candidate alignments), both launch plans, through
    size-independent properties: a consistent CSR, every read answered, each read's alignments in std::set order on the leading key (position),
    the first 2000 loci identical to the CPU oracle array by array, and a second run identical to the first."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from strelka_b200.api import DevEnumBatch

    eb = bench.make_enum_workload(1_000_000, 30, 150, 3)
    if fast:
        _fast(eb)
    n = eb.n_reads
    db = DevEnumBatch(ctx, eb, cap_alns=n * 16, cap_segs=n * 64, cap_keys=n * 32)
    ctx.enumerate_alignments_dev(db)
    nA, nS, nK = (int(x) for x in db.obufs["totals"].download(np.uint32, 3))
    aln_off = db.obufs["aln_off"].download(np.uint32, n + 1)
    status = db.obufs["status"].download(np.uint8, n)
    assert not status.any() and int(aln_off[-1]) == nA and nA > 8 * n
    per_read = np.diff(aln_off.astype(np.int64))
    assert per_read.min() >= 1 and per_read.max() <= 64
    seg_off = db.obufs["aln_seg_off"].download(np.uint32, nA + 1)
    key_off = db.obufs["aln_key_off"].download(np.uint32, nA + 1)
    assert int(seg_off[-1]) == nS and int(key_off[-1]) == nK and (np.diff(seg_off.astype(np.int64)) >= 1).all() and (np.diff(key_off.astype(np.int64)) >= 0).all()
    pos = db.obufs["aln_pos"].download(np.int32, nA)
    inside = np.ones(nA - 1, bool)
    inside[aln_off[1:-1].astype(np.int64) - 1] = False  # the steps from one read's last alignment to the next read's first
    assert (np.diff(pos.astype(np.int64))[inside] >= 0).all()
    checksum = (int(pos.astype(np.int64).sum()), int(seg_off.astype(np.int64).sum()), int(key_off.astype(np.int64).sum()))
    del pos, seg_off, key_off
    # the first loci against the oracle, array by array
    m = 2000
    sub = bench.enum_subbatch(eb, m)
    n_m = int(eb.region_read_off[m])
    want = B.EnumOut(eb, cap_alns=n_m * 64 + 64)
    import ctypes as C

    ox = reflib.oracle()
    ox.ox_enumerate_alignments.argtypes = [C.POINTER(A.SxEnumBatch), C.POINTER(A.SxEnumOut), C.c_int]
    assert ox.ox_enumerate_alignments(C.byref(sub), C.byref(want.c), 1) == 0
    a_m, s_m, k_m = (int(x) for x in want.totals[:3])
    assert np.array_equal(aln_off[: n_m + 1], want.aln_off[: n_m + 1])
    for name, dtype, count in (("aln_pos", np.int32, a_m), ("aln_seg_off", np.uint32, a_m + 1), ("segs", A.ALN_SEG_DT, s_m), ("aln_key_off", np.uint32, a_m + 1),
                               ("aln_keys", np.uint16, k_m), ("aln_lead_key", np.uint16, a_m), ("aln_trail_key", np.uint16, a_m)):
        assert db.obufs[name].download(dtype, count).tobytes() == getattr(want, name)[:count].tobytes(), name
    # idempotence
    ctx.enumerate_alignments_dev(db)
    assert [int(x) for x in db.obufs["totals"].download(np.uint32, 3)] == [nA, nS, nK]
    again = (int(db.obufs["aln_pos"].download(np.int32, nA).astype(np.int64).sum()), int(db.obufs["aln_seg_off"].download(np.uint32, nA + 1).astype(np.int64).sum()),
             int(db.obufs["aln_key_off"].download(np.uint32, nA + 1).astype(np.int64).sum()))
    assert again == checksum
    for d in list(db.bufs.values()) + list(db.obufs.values()):
        d.free()
