"""CPU: the C-ABI library loads without a GPU, exports every symbol include/strelka_b200.h declares, fails loudly (no CPU
fallback) and its POD layouts match the numpy/ctypes mirrors."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from strelka_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = A.load()
    hdr = open(os.path.join(ROOT, "include", "strelka_b200.h")).read()
    declared = set(re.findall(r"\b(sx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sx_ctx"}
    bound = {s[0] for s in A.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name)
    assert lib.sx_abi_version() == 2


def test_pod_layouts():
    assert A.ALN_SEG_DT.itemsize == 4 and A.ALN_DT.itemsize == 16 and A.REGION_DT.itemsize == 48
    assert A.GA_RESULT_DT.itemsize == 16 and C.sizeof(A.SxGaScores) == 32
    assert A.DIGT_RS_DT.itemsize == 24 and A.DIGT_RESULT_DT.itemsize == 152
    assert A.SSNV_RESULT_DT.itemsize == 288
    assert C.sizeof(A.SxParams) == 104
    assert A.INDEL_RESULT_DT.itemsize == 152
    p = A.SxParams()
    A.load().sx_default_params(C.byref(p))
    d = A.default_params()
    assert bytes(p) == bytes(d)


def test_no_cpu_fallback():
    """Without a CUDA device sx_create must fail with SX_ERR_CUDA; with one it must succeed.  Either way nothing computes on the CPU."""
    lib = A.load()
    h = C.c_void_p()
    p = A.default_params()
    rc = lib.sx_create(0, C.byref(p), C.byref(h))
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = os.path.exists("/dev/nvidia0")
    if has_gpu:
        assert rc == 0
        lib.sx_destroy(h)
    else:
        assert rc == A.SX_ERR_CUDA
        assert b"no CPU fallback" in lib.sx_last_error(None)
    p.hetVariantFrequencyExtension = 0.45
    rc = lib.sx_create(0, C.byref(p), C.byref(h))
    assert rc in (A.SX_ERR_UNSUPPORTED, A.SX_ERR_CUDA)


def test_product_never_references_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/: the product package must not."""
    pkg = os.path.join(ROOT, "strelka_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cpp", "Makefile")):
                s = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in s and "strelka_oracle" not in s and "libstrelka_ref" not in s, os.path.join(dp, f)
