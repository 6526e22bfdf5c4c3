#!/usr/bin/env python
"""compute-sanitizer pass over the kernels and variants added in round 2: the lean G-table kernel (forced, with split groups), G_SCALE_FACTOR from the exact scan on
every step, fast_moments (tree sums + scan) on every step, and the tiled scan (KSumTileSums / KSumTileMaps / KSumScan) on 40 000 addends with binade crossings.
Usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_pass2.py [--lib build/libmce_b200_racecheck.so]"""
import ctypes as ct
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import SHIFT_EXPLICIT, Session, load_product  # noqa: E402
from mceio import read_scenario  # noqa: E402

if "--lib" in sys.argv:
    from cauchyfriendly_b200._capi import bind
    lib = bind(ct.CDLL(sys.argv[sys.argv.index("--lib") + 1]))
else:
    lib = load_product()
dp = ct.POINTER(ct.c_double)
for name, steps, kw in (("lti3", 7, dict(lean=True, split=3)), ("leo7", 6, dict(lean=True, early_scale=1)), ("lti3", 7, dict(fast_moments=2)), ("lti4_2pnoise", 4, dict(early_scale=1))):
    sc = read_scenario(os.path.join(ROOT, "tests", "golden", name + ".mces"))
    s = Session(lib, sc, **kw)
    for k in range(steps):
        r = sc.rec[k]
        s.step(r)
        if r.shift_kind == SHIFT_EXPLICIT:
            s.shift_b(r.delta, -1.0)
    print(name, kw, "ok: Nt =", s.moments().Nt, flush=True)
    s.close()
sc = read_scenario(os.path.join(ROOT, "tests", "golden", "lti3.mces"))
s = Session(lib, sc)
rng = np.random.default_rng(1)
a = np.exp(np.linspace(-20, 2, 40000)) * rng.choice([1.0, 1.0, -0.5], 40000)
g = np.zeros((len(a), 2)); g[:, 0] = a
out = np.zeros(3)
assert lib.mce_debug_sum_scan(s.h, len(a), g.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
want = np.add.accumulate(a)[-1]
print("tiled scan ok: exact", out[0].tobytes() == np.float64(want).tobytes(), "restarts", int(out[1]), "tiles from summaries", int(out[2]), flush=True)
s.close()
