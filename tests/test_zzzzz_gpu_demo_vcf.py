"""GPU: the shim compiled for real (integration/build_patched.py: patched copies of three reference translation units linked against
libstrelka_b200.so) on the reference's bundled demo -- SURVEY.md 8(c) gates 4 and 6, VERDICT row g1:
    strelka2_sx   (position_somatic_snv_call -> sx_site_gl_somatic)                 bodies == src/demo/expectedResults/somatic.{snvs,indels}.vcf.gz
    starling2_sx  (position_snp_call_pprob_digt -> sx_site_gl_germline, haplotype GlobalAligner -> sx_global_align)
                                                                                   bodies == the unmodified starling2's variants / genome VCFs
The binaries and the demo inputs are built / staged here where /root/reference exists (oracle/_ref/bin, oracle/_ref/demo: git-ignored, they travel
to the GPU box); the frozen outputs of the unmodified binaries are tests/golden/demo_vcf_bodies.json."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integration"))
import run_demo as R  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "demo_vcf_bodies.json")))


def _need(binary):
    if not os.path.exists(os.path.join(R.BIN, binary)) or not os.path.isdir(R.DEMO):
        pytest.skip("oracle/_ref/bin or oracle/_ref/demo missing: run integration/build_patched.py where the reference tree exists")


def _calls(stderr, what):
    for ln in stderr.splitlines():
        if ln.startswith("sx_shim:"):
            parts = ln.split()
            return {"germline": int(parts[1]), "somatic": int(parts[5]), "align": int(parts[9])}[what]
    return -1


def test_somatic_demo_vcfs_are_the_expected_results(tmp_path):
    _need("strelka2_sx")
    got, err = R.run("somatic", "strelka2_sx", str(tmp_path), env={"SX_SHIM_REPORT": "1"})
    assert _calls(err, "somatic") > 1000, err[-500:]  # the GPU path really ran: one call per position of the 5 kb contig with coverage
    exp = {"snvs.vcf": R.body(os.path.join(R.DEMO, "expected", "somatic.snvs.vcf.gz")), "indels.vcf": R.body(os.path.join(R.DEMO, "expected", "somatic.indels.vcf.gz"))}
    assert len(exp["snvs.vcf"]) == 17 and len(exp["indels.vcf"]) == 2
    for f in exp:
        assert got[f] == exp[f], f
        assert got[f] == GOLD["somatic"][f], f


def test_germline_demo_vcfs_equal_the_unmodified_binary(tmp_path):
    _need("starling2_sx")
    got, err = R.run("germline", "starling2_sx", str(tmp_path), env={"SX_SHIM_REPORT": "1"})
    assert _calls(err, "germline") > 5000, err[-500:]  # two samples x every position
    for f, want in GOLD["germline"].items():
        assert got[f] == want, f
    if os.path.exists(os.path.join(R.BIN, "starling2")):  # and against the unmodified binary run on this very box
        ref, _ = R.run("germline", "starling2", str(tmp_path / "ref"))
        for f in ref:
            assert got[f] == ref[f], f
