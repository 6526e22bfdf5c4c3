#!/usr/bin/env python
"""A short pass over every kernel family for compute-sanitizer (memcheck / racecheck / initcheck): a few steps of a TP scenario and of
the LEO7 window, the small transforms, and the 1-D / 2-D marginal cpdf grids.  Usage: compute-sanitizer --tool racecheck python tools/sanitize_pass.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import SHIFT_EXPLICIT, Session, _dp, load_product  # noqa: E402
from mceio import read_scenario  # noqa: E402

# --lib <path>: an alternative build of the library (build/libmce_b200_racecheck.so: -DMCE_RACECHECK, see mce_kern_group2.h)
if "--lib" in sys.argv:
    import ctypes as ct
    from cauchyfriendly_b200._capi import bind
    lib = bind(ct.CDLL(sys.argv[sys.argv.index("--lib") + 1]))
else:
    lib = load_product()
for name, steps in (("lti3", 7), ("leo7", 6), ("lti4_2pnoise", 4)):
    sc = read_scenario(os.path.join(ROOT, "tests", "golden", name + ".mces"))
    s = Session(lib, sc)
    for k in range(steps):
        r = sc.rec[k]
        s.step(r)
        if r.shift_kind == SHIFT_EXPLICIT:
            s.shift_b(r.delta, -1.0)
    nu = np.full(sc.d, 0.7)
    xy = np.zeros((64, 2)); xyz = np.zeros((81, 3))
    assert lib.mce_marginal_1d_grid(s.h, 0, _dp(nu), -1.0, 1.0, 0.04, _dp(xy), 64) == 51
    assert lib.mce_marginal_2d_grid(s.h, 0, 1, _dp(nu), -1.0, 1.0, 0.25, -1.0, 1.0, 0.25, _dp(xyz), 81, None, None) == 81
    T = np.eye(sc.d) + 0.05
    assert lib.mce_deterministic_time_prop(s.h, _dp(np.ascontiguousarray(T)), None, None) == 0
    s.step(sc.rec[steps])
    print(name, "ok: Nt =", s.moments().Nt, flush=True)
    s.close()
