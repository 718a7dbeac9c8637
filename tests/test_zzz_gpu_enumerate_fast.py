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
    eb.opts.flags = A.SX_ENUM_F_FAST
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
    assert ctx.timing().launches == 7
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
