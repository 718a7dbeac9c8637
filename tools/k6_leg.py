#!/usr/bin/env python
"""Runs only the K6 score_indels leg of bench.py (for profiling): python tools/k6_leg.py [n_loci]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from strelka_b200.api import Context  # noqa: E402

ctx = Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
print(json.dumps(bench.k6_score_indels_leg(ctx, bench.load_synth(), 6572.2, n, 30, 150, 4, 0, os.cpu_count() or 8, 0, None, reps=3, cpu_regions=2000)))
