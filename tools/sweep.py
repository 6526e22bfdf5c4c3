#!/usr/bin/env python
"""Synthetic LTI term-count sweep: n = 2..8 states, one scalar Cauchy measurement per step, windows as
deep as the term cap / max_shape <= 16 allow.  Prints one line per (n, step) and a summary per n:
child terms/s over the whole window on one GPU.  Usage: python tools/sweep.py [term_cap] [n_lo] [n_hi]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_scenarios import lti  # noqa: E402
from harness import Session, load_product  # noqa: E402

cap = float(sys.argv[1]) if len(sys.argv) > 1 else 3e6
n_lo = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n_hi = int(sys.argv[3]) if len(sys.argv) > 3 else 8
lib = load_product()


def cells_half(m, d):
    from math import comb
    return sum(comb(m - 1, i) for i in range(d)) if m >= d else 2 ** (m - 1)


summary = []
for n in range(n_lo, n_hi + 1):
    steps = 17 - n                                   # max_shape = steps - 1 + n <= 16 ...
    while cells_half(steps - 1 + n, n) > 5000:       # ... and the largest table must fit the group kernel's shared memory
        steps -= 1
    rng = np.random.RandomState(1000 + n)
    Phi = rng.uniform(-1, 1, (n, n)); Gam = rng.uniform(-1, 1, n); H = rng.uniform(-1, 1, n)
    Phi *= 0.95 / np.max(np.abs(np.linalg.eigvals(Phi)))
    x = np.zeros(n); zs = []
    for _ in range(steps):
        x = Phi @ x + Gam * 0.1 * rng.standard_cauchy(); zs.append(H @ x + 0.2 * rng.standard_cauchy())
    _, sc = lti("syn%d_deep" % n, Phi, Gam, H, [0.1], [0.2], np.eye(n), np.full(n, .1), np.zeros(n), zs, steps, seed=100 + n)
    phases = bool(os.environ.get("SWEEP_PHASES"))
    s = Session(lib, sc, phase_timing=phases)
    for rep in range(2):                             # pass 0 sizes the buffers
        lib.mce_reset(s.h)
        child = 0; ms = 0.0; prev = 1; last = 0; rows = []
        for k, r in enumerate(sc.rec):
            t0 = time.perf_counter()
            s.step(r)
            dt = (time.perf_counter() - t0) * 1e3
            st = s.stats()
            child += st.terms_after_muc - prev if k > 0 else st.terms_after_muc
            prev = st.survivors if st.survivors else prev
            ms += dt; last = k + 1
            rows.append((k + 1, st.parents, st.terms_after_muc, st.survivors, dt, st.split_groups))
            if phases and rep == 1:
                print("   phases: tp %.2f mu %.2f mom %.2f regroup %.2f ftr %.2f gtable %.2f compact %.2f" % (st.ms_tp, st.ms_mu, st.ms_moments, st.ms_regroup, st.ms_ftr, st.ms_gtable, st.ms_compact))
            if st.terms_after_muc * 3.5 > cap and k + 1 < len(sc.rec):      # the next step would exceed the cap
                break
    s.close()
    for row in rows:
        print("n=%d step %2d: parents %9d  terms after MUC %10d  survivors %9d  %8.2f ms  split groups %d" % ((n,) + row), flush=True)
    summary.append({"n": n, "steps_run": last, "child_terms": int(child), "ms": ms, "child_terms_per_s": child / (ms * 1e-3),
                    "largest_step_terms": int(rows[-1][2]), "largest_step_ms": rows[-1][4]})
    print("n=%d: %d steps, %d child terms in %.1f ms -> %.2f M child terms/s" % (n, last, child, ms, child / ms / 1e3), flush=True)
print(json.dumps({"sweep": summary, "term_cap": cap}))
