"""The C-ABI library loads and exports every symbol include/mce_b200.h declares (no compute without a GPU)."""
import ctypes as ct
import os
import re
import subprocess

import numpy as np
import pytest

from harness import ROOT


@pytest.fixture(scope="module")
def lib():
    from cauchyfriendly_b200 import build
    build.build()
    from cauchyfriendly_b200 import _capi
    return _capi.load()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mce_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mce_[a-z_0-9]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 15
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "cauchyfriendly_b200", "libmce_b200.so")]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    from cauchyfriendly_b200 import _capi
    assert sorted(_capi.SYMBOLS) == syms


def test_library_targets_sm100a_and_holds_no_cpu_backend(lib):
    so = os.path.join(ROOT, "cauchyfriendly_b200", "libmce_b200.so")
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-lelf", so]).decode()
    assert "sm_100a" in out
    syms = subprocess.check_output(["nm", "-DC", so]).decode()
    assert "EmuBackend" not in syms and "mceo_" not in syms      # neither the test emulation nor the oracle is linked in


def test_bad_arguments_fail_loudly(lib):
    from cauchyfriendly_b200 import _capi
    o = _capi.MceOptions()
    lib.mce_default_options(ct.byref(o))
    assert o.device == -1 and list(o.tr_search_order) == list(range(12))
    d = 3
    A0 = np.eye(d).ravel().copy(); p0 = np.ones(d); b0 = np.zeros(d); rp = np.ones(d) * 1.5; bp = np.zeros(64)
    dp = lambda a: a.ctypes.data_as(ct.POINTER(ct.c_double))
    # 40 steps of one process-noise column would need 42 hyperplanes: beyond the reference's cap of 31 (est:231-235)
    h = lib.mce_create(d, 0, 1, 1, 40, dp(A0), dp(p0), dp(b0), dp(rp), dp(bp), ct.byref(o))
    assert not h and b"31" in lib.mce_last_error()
    h = lib.mce_create(9, 0, 1, 1, 3, dp(np.eye(9).ravel().copy()), dp(np.ones(9)), dp(np.zeros(9)), dp(np.ones(9)), dp(bp), ct.byref(o))
    assert not h


def test_create_without_gpu_reports_no_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cauchyfriendly_b200 import _capi
    o = _capi.MceOptions()
    lib.mce_default_options(ct.byref(o))
    d = 3
    dp = lambda a: a.ctypes.data_as(ct.POINTER(ct.c_double))
    h = lib.mce_create(d, 0, 1, 1, 5, dp(np.eye(d).ravel().copy()), dp(np.ones(d)), dp(np.zeros(d)), dp(np.ones(d) * 1.5), dp(np.zeros(64)), ct.byref(o))
    assert not h
    assert b"no CPU fallback" in lib.mce_last_error()
