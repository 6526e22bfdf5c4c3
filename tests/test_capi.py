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


def test_ctypes_mirrors_match_the_c_structs(tmp_path):
    """The Python mirrors of the C ABI structs (cauchyfriendly_b200/_capi.py) have the size and the field offsets the C compiler gives include/mce_b200.h:
    a field added to one side only would silently shift everything behind it."""
    import ctypes as ct
    import subprocess
    from cauchyfriendly_b200 import _capi
    src = tmp_path / "abi.c"
    src.write_text("""#include <stdio.h>
#include <stddef.h>
#include "mce_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\\n", sizeof(mce_options), offsetof(mce_options, lean_group_kernel), offsetof(mce_options, fast_moments_min_slots), offsetof(mce_options, reserved));
  printf("%zu %zu %zu\\n", sizeof(mce_moments), offsetof(mce_moments, g_scale_factor), offsetof(mce_moments, skip_post_mu));
  printf("%zu %zu %zu\\n", sizeof(mce_step_stats), offsetof(mce_step_stats, ev_mu_ms), offsetof(mce_step_stats, gtable_lean_launches));
  printf("%zu\\n", sizeof(mce_alltoallv_args));
  return 0;
}
""")
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe])
    rows = [[int(v) for v in line.split()] for line in subprocess.check_output([exe]).decode().splitlines()]
    O, M, S = _capi.MceOptions, _capi.MceMoments, _capi.MceStepStats
    assert rows[0] == [ct.sizeof(O), O.lean_group_kernel.offset, O.fast_moments_min_slots.offset, O.reserved.offset]
    assert rows[1] == [ct.sizeof(M), M.g_scale_factor.offset, M.skip_post_mu.offset]
    assert rows[2] == [ct.sizeof(S), S.ev_mu_ms.offset, S.gtable_lean_launches.offset]
    assert rows[3] == [ct.sizeof(_capi.MceAllToAllV)]
