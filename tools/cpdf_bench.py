#!/usr/bin/env python
"""Times the device 1-D marginal cpdf (mce_marginal_1d_grid) next to the unmodified reference's grid evaluation
(oracle/_ref/ref_cpdf_cpu1, one thread) on the same scenario, step and grid.  Prints one JSON line."""
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import SHIFT_EXPLICIT, Session, _dp, load_product  # noqa: E402
from mceio import read_scenario  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "leo7"
step = int(sys.argv[2]) if len(sys.argv) > 2 else 11
lo, hi, res = (float(v) for v in (sys.argv[3:6] if len(sys.argv) > 5 else ("-0.2", "0.2", "0.0001")))
scen = os.path.join(ROOT, "tests", "golden", name + ".mces")
sc = read_scenario(scen)
lib = load_product()
s = Session(lib, sc)
for k in range(step):
    r = sc.rec[k]
    s.step(r)
    if r.shift_kind == SHIFT_EXPLICIT:
        s.shift_b(r.delta, -1.0)
nt = s.moments().Nt
n = lib.mce_cpdf_grid_count(lo, hi, res)
nu = np.array([0.25 + 1.5 * v - int(1.5 * v) for v in sc.root_point[: sc.d]])
xy = np.zeros((n, 2))
dev_ms, wall_ms = [], []
for rep in range(3):
    for idx in range(sc.d):
        t0 = time.perf_counter()
        assert lib.mce_marginal_1d_grid(s.h, idx, _dp(nu), lo, hi, res, _dp(xy), n) == n
        wall_ms.append((time.perf_counter() - t0) * 1e3)
        dev_ms.append(lib.mce_cpdf_last_ms(s.h))
s.close()
dev, wall = float(np.median(dev_ms)), float(np.median(wall_ms))
out = {"scenario": name, "step": step, "terms": int(nt), "grid_points": int(n), "device_ms_per_state": dev, "e2e_ms_per_state": wall,
       "term_point_evals_per_s": nt * n / (wall * 1e-3)}
ref = os.path.join(ROOT, "oracle", "_ref", "ref_cpdf_cpu1")
if os.path.exists(ref) and "--no-ref" not in sys.argv:
    # the reference on a coarser grid of the same span (its cost is linear in the number of points), one thread
    coarse = res * 50
    txt = subprocess.run([ref, scen, "/tmp/_cpdf_ref.mced", repr(lo), repr(hi), repr(coarse), str(step), "--time"], capture_output=True, text=True).stdout
    ms = [int(m.group(3)) for m in re.finditer(r"step (\d+) idx (\d+): \d+ terms x (?:\d+) points in (\d+) ms", txt)]
    pts = [int(m.group(1)) for m in re.finditer(r"terms x (\d+) points", txt)]
    if ms:
        out["reference_cpu_ms_per_state"] = float(np.median(ms))
        out["reference_grid_points"] = pts[0]
        out["reference_term_point_evals_per_s"] = nt * pts[0] / (np.median(ms) * 1e-3)
        out["speedup_e2e"] = out["term_point_evals_per_s"] / out["reference_term_point_evals_per_s"]
if "--2d" in sys.argv:
    # 2-D marginal of the pair (0, 1) on an n2 x n2 grid over the same span; the reference on a 21 x 21 grid, one thread
    import ctypes as ct
    s = Session(lib, sc)
    for k in range(step):
        r = sc.rec[k]
        s.step(r)
        if r.shift_kind == SHIFT_EXPLICIT:
            s.shift_b(r.delta, -1.0)
    n2 = 201
    r2 = (hi - lo) / (n2 - 1)
    cnt = lib.mce_cpdf_grid_count(lo, hi, r2) ** 2
    xyz = np.zeros((cnt, 3))
    ms2, wall2 = [], []
    for rep in range(3):
        t0 = time.perf_counter()
        assert lib.mce_marginal_2d_grid(s.h, 0, 1, _dp(nu), lo, hi, r2, lo, hi, r2, _dp(xyz), cnt, None, None) == cnt
        wall2.append((time.perf_counter() - t0) * 1e3)
        ms2.append(lib.mce_cpdf_last_ms(s.h))
    s.close()
    out["cpdf2d"] = {"pair": [0, 1], "grid_points": int(cnt), "device_ms": float(np.median(ms2)), "e2e_ms": float(np.median(wall2)),
                     "term_point_evals_per_s": nt * cnt / (np.median(wall2) * 1e-3)}
    if os.path.exists(ref) and "--no-ref" not in sys.argv:
        rr = (hi - lo) / 20
        txt = subprocess.run([ref, scen, "/tmp/_cpdf_ref2.mced", repr(lo), repr(hi), repr((hi - lo) / 4), str(step), "--time", "--2d",
                              repr(lo), repr(hi), repr(rr), repr(lo), repr(hi), repr(rr)], capture_output=True, text=True).stdout
        m2 = re.search(r"pair 0,1: \d+ terms x (\d+) points in (\d+) ms", txt)
        if m2:
            rp, rms = int(m2.group(1)), int(m2.group(2))
            out["cpdf2d"].update({"reference_grid_points": rp, "reference_cpu_ms": rms, "reference_term_point_evals_per_s": nt * rp / (rms * 1e-3),
                                  "speedup_e2e": out["cpdf2d"]["term_point_evals_per_s"] / (nt * rp / (rms * 1e-3))})
print(json.dumps(out))
