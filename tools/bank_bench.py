#!/usr/bin/env python
"""Steady-state step time of the sliding-window bank (3-state LTI system, W windows) with sequential and with concurrent
stepping of the windows on one GPU."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cauchyfriendly_b200.windows import SlidingWindowBank  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 9
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
Phi = np.array([[1.4, -0.6, -1.0], [-0.2, 1.0, 0.5], [0.6, -0.6, -0.2]])
Gamma = np.array([.1, .3, -.2]); H = np.array([[1.0, .5, .2]])
rng = np.random.RandomState(3)
x = np.zeros(3); zs = []
for _ in range(steps):
    x = Phi @ x + Gamma * 0.1 * rng.standard_cauchy(); zs.append(H[0] @ x + 0.2 * rng.standard_cauchy())
for concurrent in (False, True):
    bank = SlidingWindowBank(W, np.eye(3), [.1, .08, .05], np.zeros(3), Phi, None, Gamma, [.1], H, [.2], seed=5, concurrent=concurrent)
    ts = []
    for z in zs:
        t0 = time.perf_counter(); bank.step([z]); ts.append(time.perf_counter() - t0)
    bank.shutdown()
    ss = np.array(ts[2 * W:])
    print("W=%d concurrent=%s: steady-state bank step %.2f ms (median), %.2f ms (mean), %.1f Hz" % (W, concurrent, 1e3 * np.median(ss), 1e3 * ss.mean(), 1.0 / ss.mean()))
