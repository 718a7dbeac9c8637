"""GPU parity at BASELINE.json sizes, through sampling and size-independent properties (the oracle cannot finish 1.8e10 cell
updates in seconds):
  cfg2  1M candidate loci x 30 reads x 150 bp x 4 haplotype paths (+ 1M pileup columns, 1.5M DP matrices)
  cfg3  500k somatic sites, normal 30x / tumor 60x
Properties: a random sample of regions/sites agrees bit-for-bit with the oracle; the host-buffer path (chunk-pipelined) and the
device-resident path give identical bytes; results do not depend on batch composition (a sub-batch gives the same values)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import reflib  # noqa: E402
from strelka_b200 import _abi as A  # noqa: E402
from strelka_b200 import batch as B  # noqa: E402

pytestmark = pytest.mark.gpu

N_LOCI = int(os.environ.get("SX_FULLSIZE_LOCI", "1000000"))


@pytest.fixture(scope="module")
def ctx():
    from strelka_b200.api import Context

    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def workload(ctx):
    import bench

    synth = bench.load_synth()
    alloc = bench.HostAlloc(ctx.lib, True)
    ab, pb, gb = bench.make_workload(synth, alloc, N_LOCI, 30, 150, 4, 11, os.cpu_count() or 8)
    yield ab, pb, gb
    alloc.free()


def test_cfg2_k1_full_size(ctx, workload):
    from strelka_b200.api import DevAlignBatch

    ab, _, _ = workload
    host = ctx.score_alignments(ab)                      # chunk-pipelined host path
    db = DevAlignBatch(ctx, ab)
    ctx.score_alignments_dev(db)                         # resident path, single launch
    dev = db.out.download(np.float64, ab.n_alns)
    assert np.array_equal(host.view(np.uint64), dev.view(np.uint64))
    assert np.all(np.isfinite(host)) and np.all(host < 0)
    # sampled regions against the oracle, bit for bit
    rng = np.random.default_rng(0)
    want = np.zeros(ab.n_alns, np.float64)
    ox = reflib.oracle()
    for r in rng.integers(0, ab.n_regions, 400):
        ox.ox_score_alignments_range(C.byref(ab.c), C.c_uint32(int(r)), C.c_uint32(int(r) + 1), want.ctypes.data)
        a0, a1 = int(ab.regions["aln_begin"][r]), int(ab.regions["aln_begin"][r + 1])
        assert np.array_equal(host[a0:a1].view(np.uint64), want[a0:a1].view(np.uint64)), r
    # a read's best haplotype path is never worse than its reference path
    per_read = host.reshape(-1, 4)
    assert np.all(per_read.max(axis=1) >= per_read[:, 0])


def test_cfg2_k2a_full_size(ctx, workload):
    _, pb, _ = workload
    got = ctx.site_gl_germline(pb, True)
    assert np.all(got["is_computed"] == 1)
    # PLs: the best genotype has PL 0, every PL is a non-negative integer bounded by the float floor
    assert np.all(got["phredLoghood"].min(axis=1) == 0)
    assert got["phredLoghood"].max() <= 370
    # sub-batch independence + oracle on a sample
    rng = np.random.default_rng(1)
    idx = np.sort(rng.choice(pb.n_sites, 3000, replace=False))
    sub_off = np.zeros(len(idx) + 1, np.uint32)
    sub_off[1:] = np.cumsum(pb.site_off[idx + 1] - pb.site_off[idx])
    sub_calls = np.concatenate([pb.calls[pb.site_off[i]:pb.site_off[i + 1]] for i in idx])
    sub = B.PileupBatch(sub_off, sub_calls, pb.ref_base[idx])
    want = reflib.ox_germline(A.default_params(), sub, True)
    g = got[idx]
    assert np.array_equal(g["phredLoghood"], want["phredLoghood"])
    assert np.array_equal(g["lhood"].view(np.uint32), want["lhood"].view(np.uint32))
    for rs in ("genome", "poly"):
        for f in ("max_gt", "snp_qphred", "max_gt_qphred"):
            assert np.array_equal(g[rs][f], want[rs][f])
    again = ctx.site_gl_germline(sub, True)
    assert again.tobytes() == g.tobytes()


def test_cfg2_k3_full_size(ctx, workload):
    _, _, gb = workload
    sc = ctx.active_region_scores()
    res, cig = ctx.global_align(sc, gb)
    assert np.all(res["status"] == 0)
    # every path explains the whole query and (isRequireEdgeDeletion) the whole reference
    qlen = np.diff(gb.query_off).astype(np.int64)
    rlen = np.diff(gb.ref_off).astype(np.int64)
    ops_len = (cig >> 4).astype(np.int64)
    ops_code = cig & 15
    valid = np.arange(cig.shape[1])[None, :] < res["n_ops"][:, None]
    q_cons = (ops_len * (valid & np.isin(ops_code, [1, 4, 7, 8]))).sum(axis=1)   # I S = X
    r_cons = (ops_len * (valid & np.isin(ops_code, [2, 7, 8]))).sum(axis=1)      # D = X
    assert np.array_equal(q_cons, qlen)
    assert np.array_equal(r_cons + res["beginPos"], rlen)
    rng = np.random.default_rng(2)
    idx = np.sort(rng.choice(gb.n, 3000, replace=False))
    qs = [bytes(gb.query[gb.query_off[i]:gb.query_off[i + 1]]).decode() for i in idx]
    rs = [bytes(gb.ref[gb.ref_off[i]:gb.ref_off[i + 1]]).decode() for i in idx]
    sub = B.GaBatch(qs, rs, max_ops=gb.max_ops)
    o_res, o_cig = reflib.ox_global_align(sc, sub)
    assert np.array_equal(o_res, res[idx])
    assert np.array_equal(o_cig, cig[idx])


def test_cfg3_somatic_full_size(ctx):
    import bench

    synth = bench.load_synth()
    n = int(os.environ.get("SX_FULLSIZE_SOMATIC_SITES", "500000"))
    thr = os.cpu_count() or 8

    def pile(depth, mode):
        off = np.zeros(n + 1, np.uint32)
        nc = synth.synth_pileups(n, C.c_double(depth), mode, C.c_uint64(5), thr, C.c_void_p(off.ctypes.data), None, None)
        calls = np.zeros(nc + 8, np.uint16)
        ref = np.zeros(n, np.uint8)
        synth.synth_pileups(n, C.c_double(depth), mode, C.c_uint64(5), thr, C.c_void_p(off.ctypes.data), C.c_void_p(calls.ctypes.data), C.c_void_p(ref.ctypes.data))
        return off, calls, ref

    noff, ncalls, ref = pile(30.0, 1)
    toff, tcalls, _ = pile(60.0, 2)
    npb = B.PileupBatch(noff, ncalls, ref)
    tpb = B.PileupBatch(toff, tcalls, ref)
    got = ctx.site_gl_somatic(npb, tpb)
    m = got["is_computed"] == 1
    assert m.sum() > n // 100
    rng = np.random.default_rng(3)
    idx = np.sort(rng.choice(n, 3000, replace=False))

    def sub(pb):
        so = np.zeros(len(idx) + 1, np.uint32)
        so[1:] = np.cumsum(pb.site_off[idx + 1] - pb.site_off[idx])
        return B.PileupBatch(so, np.concatenate([pb.calls[pb.site_off[i]:pb.site_off[i + 1]] for i in idx]), pb.ref_base[idx])

    want = reflib.ox_somatic(A.default_params(), sub(npb), sub(tpb))
    g = got[idx]
    assert np.array_equal(g["is_computed"], want["is_computed"])
    mm = want["is_computed"] == 1
    for f in ("ntype", "max_gt", "qphred", "from_ntype_qphred", "normal_alt_id", "tumor_alt_id"):
        assert np.array_equal(g[f][mm], want[f][mm]), f
    assert np.array_equal(g["normal_lhood"][mm][:, :21].view(np.uint32), want["normal_lhood"][mm][:, :21].view(np.uint32))
    assert np.array_equal(g["tumor_lhood"][mm][:, :21].view(np.uint32), want["tumor_lhood"][mm][:, :21].view(np.uint32))


def test_cfg5_high_depth_amplicon(ctx):
    """BASELINE.json configs[4]: 300x depth, 32 haplotype paths per read, 10k loci (1.44e10 cell updates).  Loci are cut into
    regions of 32 reads that share the reference window; pileup columns of ~300 calls take K2a's global-scratch path."""
    import bench

    synth = bench.load_synth()
    alloc = bench.HostAlloc(ctx.lib, True)
    n_loci = int(os.environ.get("SX_FULLSIZE_CFG5_LOCI", "10000"))
    ab, pb, gb = bench.make_workload(synth, alloc, n_loci, 300, 150, 32, 21, os.cpu_count() or 8, 4, 32)
    assert ab.n_regions == n_loci * 10 and ab.n_alns == n_loci * 300 * 32
    got = ctx.score_alignments(ab)
    rng = np.random.default_rng(4)
    want = np.zeros(ab.n_alns, np.float64)
    ox = reflib.oracle()
    for r in rng.integers(0, ab.n_regions, 60):
        ox.ox_score_alignments_range(C.byref(ab.c), C.c_uint32(int(r)), C.c_uint32(int(r) + 1), want.ctypes.data)
        a0, a1 = int(ab.regions["aln_begin"][r]), int(ab.regions["aln_begin"][r + 1])
        assert np.array_equal(got[a0:a1].view(np.uint64), want[a0:a1].view(np.uint64)), r
    gl = ctx.site_gl_germline(pb, True)
    idx = np.sort(rng.choice(pb.n_sites, 300, replace=False))
    sub_off = np.zeros(len(idx) + 1, np.uint32)
    sub_off[1:] = np.cumsum(pb.site_off[idx + 1] - pb.site_off[idx])
    sub = B.PileupBatch(sub_off, np.concatenate([pb.calls[pb.site_off[i]:pb.site_off[i + 1]] for i in idx]), pb.ref_base[idx])
    w = reflib.ox_germline(A.default_params(), sub, True)
    assert np.array_equal(gl[idx]["phredLoghood"], w["phredLoghood"])
    assert np.array_equal(gl[idx]["lhood"].view(np.uint32), w["lhood"].view(np.uint32))
    alloc.free()
