#!/usr/bin/env python
"""Freezes outputs of the REFERENCE's own code (oracle/_ref/libstrelka_ref.so, built from /root/reference by
oracle/build_ref.sh) on seeded inputs into tests/golden/*.npz.  Run in the build container; the fixtures are committed
because the GPU box has no reference tree.  The inputs are regenerated from the seed by tests/specgen.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reflib  # noqa: E402
import specgen  # noqa: E402
from strelka_b200 import _abi as A  # noqa: E402
from strelka_b200 import batch as B  # noqa: E402

seed, n_regions = 4242, 60
rng = np.random.default_rng(seed)
regions = [specgen.random_region(rng, n_reads=int(rng.integers(1, 8))) for _ in range(n_regions)]
lnp = np.concatenate([reflib.ref_score_region(r) for r in regions])
np.savez_compressed(os.path.join(HERE, "k1_fixture.npz"), seed=seed, n_regions=n_regions, lnp_bits=lnp.view(np.uint64))

seed, n_sites = 777, 1500
rng = np.random.default_rng(seed)
pb = specgen.random_pileups(rng, n_sites, depth=30.0)
g = reflib.ref_germline(A.default_params(), pb, True)
npb = specgen.random_pileups(rng, n_sites, depth=30.0, alt_frac_choices=(0.0, 0.0, 0.0, 0.0, 0.02, 0.5))
tpb0 = specgen.random_pileups(rng, n_sites, depth=60.0, alt_frac_choices=(0.0, 0.0, 0.05, 0.1, 0.2, 0.4))
tpb = B.PileupBatch(tpb0.site_off, tpb0.calls, npb.ref_base)
s = reflib.ref_somatic(A.default_params(), npb, tpb)
np.savez_compressed(
    os.path.join(HERE, "k2_fixture.npz"), seed=seed, n_sites=n_sites, germ_pl=g["phredLoghood"], germ_lhood_bits=g["lhood"].view(np.uint32),
    germ_snp_q=g["genome"]["snp_qphred"], germ_max_gt=g["genome"]["max_gt"], som_computed=s["is_computed"], som_qss=s["qphred"],
    som_qss_nt=s["from_ntype_qphred"], som_ntype=s["ntype"],
)
print("fixtures written")
