#!/usr/bin/env python
"""The site-model kernels that are not part of the germline whole-path step, each measured alone (bench.py starts these as processes of their
own; also the ncu targets):
    python tools/site_legs.py k2b [hbm_peak_gbs [n_sites]]   K2b somatic strand-grid SNV model on BASELINE.json's cfg3 (500k tumor/normal sites, 60x / 30x)
    python tools/site_legs.py k5  [hbm_peak_gbs [n_loci]]    K5 indel allele-group genotype likelihoods (1M loci, 1-4 alt alleles, ~30 reads)
Inputs and outputs resident in HBM, CUDA-event kernel times, algorithmic bytes / time against the HBM peak, and the reference's own
function on one host core over a prefix as the CPU figure (the reference's site caches are function-local statics: one process = one core)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import window_workload as WW  # noqa: E402
from strelka_b200 import _abi as A  # noqa: E402
from strelka_b200 import batch as B  # noqa: E402
from strelka_b200.api import Context, DeviceArray, DevPileupBatch  # noqa: E402


def _ref():
    p = os.path.join(ROOT, "oracle", "_ref", "libstrelka_ref.so")
    return C.CDLL(p) if os.path.exists(p) else None


def k2b_leg(ctx, peak, n=500_000, reps=5, cpu_sites=20000):
    synth = WW.load_synth()
    thr = max(1, len(os.sched_getaffinity(0)))

    def pile(depth, mode):
        off = np.zeros(n + 1, np.uint32)
        nc = synth.synth_pileups(n, C.c_double(depth), mode, C.c_uint64(5), thr, C.c_void_p(off.ctypes.data), None, None)
        calls, ref = np.zeros(nc + 8, np.uint16), np.zeros(n, np.uint8)
        synth.synth_pileups(n, C.c_double(depth), mode, C.c_uint64(5), thr, C.c_void_p(off.ctypes.data), C.c_void_p(calls.ctypes.data), C.c_void_p(ref.ctypes.data))
        return off, calls, ref

    noff, ncalls, ref = pile(30.0, 1)
    toff, tcalls, _ = pile(60.0, 2)
    npb, tpb = B.PileupBatch(noff, ncalls, ref), B.PileupBatch(toff, tcalls, ref)
    dn, dt = DevPileupBatch(ctx, npb), DevPileupBatch(ctx, tpb)
    out = DeviceArray(ctx, n * A.SSNV_RESULT_DT.itemsize)
    ms = []
    for i in range(reps + 2):
        ctx.site_gl_somatic_dev(dn, dt, None, out)
        if i >= 2:
            ms.append(ctx.timing().kernel_ms)
    t = float(np.mean(ms)) * 1e-3
    alg = (int(noff[-1]) + int(toff[-1])) * 2 + n * (4 + 4 + 1) + n * A.SSNV_RESULT_DT.itemsize
    leg = {"what": f"K2b site_gl_somatic: {n} tumor/normal sites (cfg3: 60x / 30x), resident in HBM", "ms": 1e3 * t, "sites_per_s": n / t,
           "roofline": {"bound": "hbm", "achieved": alg / t / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / t / 1e9 / peak, "algorithmic_bytes": int(alg)}}
    rf = _ref()
    if rf is not None:
        m = min(cpu_sites, n)
        sub_n = A.SxPileupBatch(m, A.ptr(noff), A.ptr(ncalls), None, None, A.ptr(ref), None)
        sub_t = A.SxPileupBatch(m, A.ptr(toff), A.ptr(tcalls), None, None, A.ptr(ref), None)
        res = np.zeros(m, A.SSNV_RESULT_DT)
        err = C.create_string_buffer(512)
        p = A.default_params()
        t0 = time.perf_counter()
        rc = rf.ref_site_gl_somatic(C.byref(p), C.byref(sub_n), C.byref(sub_t), None, C.c_void_p(res.ctypes.data), err, 512)
        dt_cpu = time.perf_counter() - t0
        if rc == 0:
            got = out.download(A.SSNV_RESULT_DT, m)
            same = bool(np.array_equal(got["qphred"], res["qphred"]) and np.array_equal(got["from_ntype_qphred"], res["from_ntype_qphred"]) and np.array_equal(got["ntype"], res["ntype"]))
            leg["cpu_reference"] = {"sites_per_s": m / dt_cpu, "cores": 1, "kind": "reference", "matches_gpu": same,
                                    "sample": f"first {m} sites through the reference's position_somatic_snv_call (incl. the shim's pileup construction)"}
    return leg


def k5_leg(ctx, peak, n=1_000_000, reps=5, cpu_loci=20000):
    """numpy-built allele groups: A in 1..3 alt alleles, 30 reads each with one strongly supported allele"""
    rng = np.random.default_rng(11)
    Acnt = rng.choice([1, 1, 2, 3], size=n).astype(np.int64)
    depth = np.full(n, 30, np.int64)
    read_off = np.concatenate([[0], np.cumsum(depth)]).astype(np.uint32)
    lnp_off = np.concatenate([[0], np.cumsum(depth * (Acnt + 1))]).astype(np.uint32)
    allele_off = np.concatenate([[0], np.cumsum(Acnt)]).astype(np.uint32)
    nr, nl, na = int(read_off[-1]), int(lnp_off[-1]), int(allele_off[-1])
    lnp = rng.normal(-60.0, 15.0, nl).astype(np.float32)
    # one supported allele per read
    row0 = np.repeat(lnp_off[:-1].astype(np.int64), depth) + np.tile(np.arange(30), n) * np.repeat(Acnt + 1, depth)
    supp = (rng.random(nr) * np.repeat(Acnt + 1, depth)).astype(np.int64)
    lnp[row0 + supp] = rng.normal(-8.0, 4.0, nr).astype(np.float32)
    lnp = np.minimum(lnp, np.float32(-0.01))
    arr = {"read_off": read_off, "lnp_off": lnp_off, "allele_off": allele_off, "ploidy": np.full(n, 2, np.uint8), "allele_del_len": rng.integers(0, 20, na).astype(np.uint16),
           "allele_ins_len": rng.integers(0, 20, na).astype(np.uint16), "allele_lnp": lnp, "read_length": np.full(nr, 150, np.uint16),
           "non_ambig": np.full(nr, 150, np.uint16), "is_fwd": (rng.random(nr) < 0.5).astype(np.uint8)}
    order = ("read_off", "lnp_off", "allele_off", "ploidy", "allele_del_len", "allele_ins_len", "allele_lnp", "read_length", "non_ambig", "is_fwd")
    dev = {k: DeviceArray(ctx, arr[k].nbytes + 64).upload(arr[k]) for k in order}
    bc = A.SxIndelBatch(n, *[dev[k].ptr for k in order])
    out = DeviceArray(ctx, n * A.INDEL_RESULT_DT.itemsize)
    ms = []
    for i in range(reps + 2):
        ctx._chk(ctx.lib.sx_indel_gl_dev(ctx.h, C.byref(bc), out.ptr))
        if i >= 2:
            ms.append(ctx.timing().kernel_ms)
    t = float(np.mean(ms)) * 1e-3
    alg = sum(arr[k].nbytes for k in order) + n * A.INDEL_RESULT_DT.itemsize
    leg = {"what": f"K5 indel_gl: {n} indel loci (1-3 alt alleles, 30 reads each), resident in HBM", "ms": 1e3 * t, "loci_per_s": n / t,
           "roofline": {"bound": "hbm", "achieved": alg / t / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / t / 1e9 / peak, "algorithmic_bytes": int(alg)}}
    rf = _ref()
    if rf is not None:
        m = min(cpu_loci, n)
        hb = A.SxIndelBatch(m, *[A.ptr(arr[k]) for k in order])
        res = np.zeros(m, A.INDEL_RESULT_DT)
        err = C.create_string_buffer(512)
        p = A.default_params()
        t0 = time.perf_counter()
        rc = rf.ref_indel_gl(C.byref(p), C.byref(hb), C.c_void_p(res.ctypes.data), err, 512)
        dt_cpu = time.perf_counter() - t0
        if rc == 0:
            got = out.download(A.INDEL_RESULT_DT, m)
            same = bool(np.array_equal(got["support"], res["support"]) and np.array_equal(np.ascontiguousarray(got["gt_lhood"]).view(np.uint64), np.ascontiguousarray(res["gt_lhood"]).view(np.uint64)))
            leg["cpu_reference"] = {"loci_per_s": m / dt_cpu, "cores": 1, "kind": "reference", "matches_gpu": same,
                                    "sample": f"first {m} loci through the reference's getVariantAlleleGroupGenotypeLhoodsForSample (incl. the shim's object construction)"}
        else:
            leg["cpu_reference"] = {"error": err.value.decode(errors="replace")}
    return leg


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "k2b"
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6572.2
    ctx = Context(0)
    if which == "k2b":
        print(json.dumps(k2b_leg(ctx, peak, int(sys.argv[3]) if len(sys.argv) > 3 else 500_000)))
    else:
        print(json.dumps(k5_leg(ctx, peak, int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000)))
