"""CPU: the C-ABI library loads and exports exactly what include/opl_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "opl_b200.h")


def _declared():
    with open(HEADER) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(opl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from learning_b200 import _native as nat
    declared = _declared()
    assert len(declared) >= 35
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    # and the Python binding covers the same surface
    assert sorted(nat.SIGNATURES) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", nat.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (opl_[a-z0-9_]+)", out)))
    assert exported == declared


def test_header_cites_reference_call_sites():
    with open(HEADER) as fh:
        text = fh.read()
    for cite in ("mlp.py:134-163", "optimizer.py:234-338", "optimizer.py:414-508",
                 "linesearch.py:392-399", "error.py:46-116"):
        assert cite in text


def test_version_and_error_plumbing():
    from learning_b200 import _native as nat
    assert nat.lib.opl_version() >= 100
    assert nat.device_count() >= 0
    # argument validation happens before any CUDA call
    out = ctypes.c_void_p()
    rc = nat.lib.opl_mlp_create(ctypes.byref(out), 1, None, None, None, 0, 1, 0, None)
    assert rc == nat.OPL_E_ARG and "opl_mlp_create" in nat.last_error()
    with pytest.raises(ValueError):
        nat.check(rc)
    rc = nat.lib.opl_bfgs_create(ctypes.byref(out), 0, 0, 0, None)
    assert rc == nat.OPL_E_ARG


def test_no_cpu_fallback_without_device():
    """On a host without a GPU the product path must fail loudly, not compute on the CPU."""
    from learning_b200 import _native as nat
    if nat.device_count() > 0:
        pytest.skip("a CUDA device is present")
    import learning_b200 as L
    widths = (ctypes.c_int64 * 3)(4, 2, 3)
    tids = (ctypes.c_int32 * 2)(2, 4)
    out = ctypes.c_void_p()
    rc = nat.lib.opl_mlp_create(ctypes.byref(out), 3, widths, tids, None, 1, 8, 0, None)
    assert rc == nat.OPL_E_CUDA
    with pytest.raises(RuntimeError):
        L.MLP((2, 2, 2)).activate([0.0, 0.0])
    with pytest.raises(RuntimeError):
        L.SoftmaxTransfer()([1.0, 2.0])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "optimal-py-learning_b200")
    for base, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(base, name)) as fh:
                    body = fh.read()
                assert "oracle" not in body.replace("no oracle", ""), os.path.join(base, name)
