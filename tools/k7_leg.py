#!/usr/bin/env python
"""Runs only the K7 enumerate_alignments leg of bench.py (for profiling): python tools/k7_leg.py [n_loci]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from strelka_b200.api import Context  # noqa: E402

ctx = Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
print(json.dumps(bench.k7_enumerate_leg(ctx, 6572.2, n_loci=n)))
