"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/kxpu.h
declares, and refuses to work without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_built_and_exports_every_declared_symbol():
    import kxpu_b200 as K
    from kxpu_b200.binding import ABI_SYMBOLS
    L = K.load_library()
    hdr = open(os.path.join(ROOT, "include", "kxpu.h")).read()
    declared = set(re.findall(r"\b(kxpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(ABI_SYMBOLS), declared ^ set(ABI_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_struct_layouts_match_header():
    from kxpu_b200.binding import CDIDEV_DTYPE, DEVREC_DTYPE
    assert DEVREC_DTYPE.itemsize == 64 and CDIDEV_DTYPE.itemsize == 32
    assert DEVREC_DTYPE.fields["iommu_group"][1] == 48 and DEVREC_DTYPE.fields["flags"][1] == 54
    assert CDIDEV_DTYPE.fields["index"][1] == 24


def test_strerror_and_invalid_args():
    import kxpu_b200 as K
    L = K.load_library()
    assert L.kxpu_strerror(0) == b"ok"
    assert b"GPU" in L.kxpu_strerror(-3)
    assert L.kxpu_ctx_create(0, None) == -1


def test_no_gpu_means_hard_failure():
    """On a box without a B200 the product must fail loudly, never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    import kxpu_b200 as K
    with pytest.raises(K.KxpuError) as e:
        K.Kxpu(0)
    assert e.value.status == -3


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "kata-xpu-device-plugin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".go")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "kxo_" not in src and "libkxpu_oracle" not in src and "from oracle" not in src, f


def test_workloads_are_deterministic(workloads, oracle_rows):
    a = workloads.cfg2_queries(oracle_rows["key"])
    b = workloads.cfg2_queries(oracle_rows["key"])
    assert np.array_equal(a, b) and len(a) == 1024
    pres = set(int(k) for k in oracle_rows["key"])
    assert sum(int(k) in pres for k in a) == 768
    r = workloads.cfg3_records(oracle_rows["key"], n=4096)
    assert r["bdf"][0] == b"0000:00:00.0" and r["bdf"][9] == b"0000:00:01.1" and r["bdf"][4095] == b"0000:0f:1f.7"
    assert list(r["bdf"]) == sorted(r["bdf"])  # filepath.Walk order is lexical
