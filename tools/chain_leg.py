#!/usr/bin/env python
"""Runs only the realign_chain leg of bench.py (bench.py starts it as a process of its own; also an ncu target):
python tools/chain_leg.py [n_loci [depth [read_len [hbm_peak_gbs [original|fast]]]]]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from strelka_b200.api import Context  # noqa: E402

ctx = Context(0)
arg = lambda i, d, t: t(sys.argv[i]) if len(sys.argv) > i else d  # noqa: E731
print(json.dumps(bench.realign_chain_leg(ctx, arg(4, 6572.2, float), n_loci=arg(1, 100_000, int), depth=arg(2, 30, int), read_len=arg(3, 150, int),
                                         fast=arg(5, "original", str) == "fast")))
