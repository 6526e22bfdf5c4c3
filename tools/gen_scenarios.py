#!/usr/bin/env python
"""Generates the LTI scenario files under tests/golden/ (inputs only).

Each scenario pins the inputs of one of the reference's own example programs
(/root/reference/src/cauchy_estimator.cpp, file:line cited per config) or a synthetic LTI system
(SURVEY.md section 8d config 5).  root_point / b_pert -- which the reference draws with libc rand()
(cauchy_estimator.hpp:125-128, cell_enumeration.hpp:467-470) -- are drawn here from a seeded
numpy generator and recorded in the scenario, so that every implementation sees the same values.

Usage: python tools/gen_scenarios.py [outdir]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from mceio import Scenario, StepRecord, write_scenario  # noqa: E402


def _rand_vectors(seed, d, max_shape):
    rng = np.random.RandomState(seed)
    root_point = 1.0 + (rng.randint(0, 2**31 - 1, d) + 1.0) / 2.0**31  # 1 + U(0,1], as est:128
    b_pert = 2.0 * (rng.randint(0, 2**31 - 1, max_shape) + 1.0) / 2.0**31 - 1.0  # 2u-1, as ce:470
    return root_point, b_pert


def lti(name, Phi, Gamma, H, beta, gamma, A0, p0, b0, zs, steps, p=1, order=None, seed=7):
    Phi = np.asarray(Phi, float)
    d = Phi.shape[0]
    Gamma = np.asarray(Gamma, float).reshape(d, -1) if Gamma is not None else np.zeros((d, 0))
    pncc = Gamma.shape[1]
    H = np.asarray(H, float).reshape(p, d)
    gamma = np.asarray(gamma, float).reshape(p)
    beta = np.asarray(beta, float).reshape(pncc)
    max_shape = (steps - 1) * pncc + d if d > 1 else d + pncc
    root_point, b_pert = _rand_vectors(seed, d, max_shape)
    s = Scenario(d, 0, pncc, p, steps, order or list(range(12)), root_point, b_pert,
                 np.asarray(A0, float).reshape(d, d), np.asarray(p0, float), np.asarray(b0, float))
    for i, z in enumerate(zs):
        s.rec.append(StepRecord(float(z), float(gamma[i % p]), Phi, Gamma, beta, H[i % p]))
    return name, s


def all_scenarios():
    out = []
    # 3-state LTI, src/cauchy_estimator.cpp:91-119 (steps=11 with 10 initialisers -> 11th z is 0.0)
    Phi3 = [[1.4, -0.6, -1.0], [-0.2, 1.0, 0.5], [0.6, -0.6, -0.2]]
    zs3 = [-1.2172011200334241, -0.35943271347277583, -0.52353301003957098, 0.5855389648301792,
           -0.8048243525901404, 0.34053610027255954, 1.0580483915838776, -0.55152999529515989,
           -0.72879029737003309, -0.82415138330170357, 0.0]
    out.append(lti("lti3", Phi3, [.1, .3, -.2], [1.0, .5, .2], [.1], [.2], np.eye(3), [.10, .08, .05],
                   np.zeros(3), zs3, 11))
    # 2-state LTI, src/cauchy_estimator.cpp:50-88
    zs2 = [0.0338, 0.2049, -2.3543, -0.6042, -0.2662, 0.1307, -0.2250, 0.1951, -0.2191, 0.0996]
    out.append(lti("lti2", [[.9, .1], [-.2, 1.1]], [1, .3], [1.0, 1.0], [.1], [.2], np.eye(2), [.10, .05],
                   np.zeros(2), zs2, 10))
    # 4-state LTI, src/cauchy_estimator.cpp:122-158
    Phi4 = [[1.4, -.6, -1.0, 0], [-.2, 1.0, .5, 0], [.6, -.6, -.2, 0], [0, 0, 0, .5]]
    zs4 = [-0.26300165310514712, -0.98289343232730964, -0.93317363235517392, -0.81311530427193779,
           -0.24140673945883995, 0.013971096637110103, -0.4842328985975715, -0.1607056967588112]
    out.append(lti("lti4", Phi4, [.1, .3, -.2, .4], [2.0, .5, .2, -.1], [.1], [.2], np.eye(4),
                   [.1, .08, .05, .2], np.zeros(4), zs4, 8))
    # 3-state, 3 measurements per step with H = I rows (H-orthogonal hyperplanes), src/cauchy_estimator.cpp:161-207
    zs33 = [0.10943250903225685, 0.32131358116921616, -0.39352816664526724,
            0.76258687662854907, -0.25344840215960657, 0.1578820974338809,
            0.52543601367678883, -0.67309502187832315, -0.37267005411252474,
            2.7335350536863903, -0.3754139600950176, 0.6986657326616188,
            0.52558307773279223, 0.82802377147093642, 0.98211422248186553]
    out.append(lti("lti3_3msmts", Phi3, [.1, .3, -.2], np.eye(3), [.1], [.2, .15, .10], np.eye(3),
                   [.1, .08, .05], np.zeros(3), zs33, 5, p=3))
    # 4-state, two process noises, src/cauchy_estimator.cpp:253-300 (steps=7 with 5 initialisers)
    Phi42 = [[1.4, -.6, -1.0, 0], [-.2, 1.0, .5, 0], [.6, -.6, -.2, 0], [0, 0, 0, 1.0]]
    Gam42 = [[.1, 0], [.3, 0], [.2, 0], [0, -1.0]]
    zs42 = [-5.3335189550166655, -4.4110988021211845, -3.6610012492599329,
            -2.5741683288219699, -6.5109959475268671, 0.0, 0.0]
    out.append(lti("lti4_2pnoise", Phi42, Gam42, [0.4165285461783826, -.60, -1.0, 1.0], [.1, .001], [.2],
                   np.eye(4), [.4, .5, .6, .7], np.zeros(4), zs42, 7))
    # 4-state, two measurements per step, src/cauchy_estimator.cpp:303-347
    zs44 = [-0.2630016531051471, 1.5804720565253951, -0.8123538549377811, 0.4001811238098553,
            -0.8607383321320500, -1.1124356889621634, -0.1026529815581601, -0.6794624240892977,
            -3.9121237378676339, -1.3582285870633606, -0.5498081141794045, 1.1925540511414112]
    out.append(lti("lti4_2msmts", Phi4, [.1, .3, -.2, .4], [[2.0, .5, .2, -.1], [.4, -.7, 1.3, -1.5]], [.1],
                   [.2, .15], np.eye(4), [.1, .08, .05, .2], np.zeros(4), zs44, 6, p=2))
    # Synthetic LTI sweep: n = 2..8 states, seeded numpy MT19937 streams.
    for n, steps in ((2, 12), (3, 10), (4, 8), (5, 7), (6, 6), (7, 6), (8, 5)):
        rng = np.random.RandomState(1000 + n)
        Phi = rng.uniform(-1, 1, (n, n))
        Gam = rng.uniform(-1, 1, n)
        H = rng.uniform(-1, 1, n)
        Phi *= 0.95 / np.max(np.abs(np.linalg.eigvals(Phi)))
        beta, gamma = 0.1, 0.2
        x = np.zeros(n)
        zs = []
        for _ in range(steps):
            x = Phi @ x + Gam * beta * rng.standard_cauchy()
            zs.append(H @ x + gamma * rng.standard_cauchy())
        out.append(lti("syn%d" % n, Phi, Gam, H, [beta], [gamma], np.eye(n), np.full(n, .1), np.zeros(n), zs,
                       steps, seed=100 + n))
    return out


def homing3():
    """Homing-missile-shaped window (BASELINE.json configs[1], src/homing_missile.cpp:378-447, 520-682): 3 states (relative
    position, relative velocity, target acceleration), one process noise, ONE CONTROL INPUT (the pursuer's acceleration command:
    cmcc = 1, B and u passed to step()), a radar bearing linearised about the running estimate -- so H and gamma change every
    step -- and the extended estimator's re-centring after every step (finalize_extended_moments, est:1358).  The closed loop
    of the example (simulator, guidance law, time(NULL) seed) is replaced by a seeded open-loop recording of the same shape."""
    n, steps, dt, tau, Vc, tf = 3, 8, 0.1, 2.0, 300.0, 10.0
    rng = np.random.RandomState(2024)
    Phi = np.array([[1.0, dt, dt * dt / 2.0], [0.0, 1.0, dt], [0.0, 0.0, np.exp(-dt / tau)]])
    Gamma = np.array([[0.0], [0.0], [1.0]])
    B = np.array([[-dt * dt / 2.0], [-dt], [0.0]])
    beta = np.array([0.05])
    max_shape = (steps - 1) * 1 + n
    root_point, b_pert = _rand_vectors(311, n, max_shape)
    s = Scenario(n, 1, 1, 1, steps, list(range(12)), root_point, b_pert, np.eye(n), np.array([0.6, 0.4, 0.3]), np.zeros(n))
    x = np.array([0.5, -0.2, 0.3])
    for k in range(steps):
        t_k = (k + 1) * dt
        u = np.array([3.0 * x[0] / ((tf - t_k) ** 2) + 0.4 * rng.standard_normal()])       # proportional-navigation-like command
        x = Phi @ x + (B @ u) + Gamma[:, 0] * beta[0] * rng.standard_cauchy()
        rng_h = 1.0 / (Vc * (tf - t_k + 1e-6))                                               # d atan(y / (Vc (tf - t))) / dy at y ~ 0
        H = np.array([rng_h * Vc, 0.0, 0.0]) * (1.0 + 0.05 * k)                              # rescaled so that H is O(1) and varies per step
        gamma = 0.08 * (1.0 + 0.1 * k)
        z = float(H @ x + gamma * rng.standard_cauchy())
        s.rec.append(StepRecord(z, float(gamma), Phi, Gamma, beta, H, B=B, u=u, shift_kind=1))
    return "homing3", s


# Deep-DECLARED windows: the same recordings with a larger declared `steps`, so that max_shape = d + (steps-1) pncc
# (est:97) exceeds 16 and the engine takes its max_shape > 16 kernels (sort/hash G-table and DCE-TP variants; the bitmap
# kernels serve max_shape <= 16 only) while the replayed prefix stays cheap for the CPU reference.  The reference supports
# up to 31 hyperplanes (est:231-235).  name -> (source scenario, declared steps, records kept)
DEEP = {"lti3_deep": ("lti3", 20, 9), "lti4_2pnoise_deep": ("lti4_2pnoise", 8, 6), "lti3_3msmts_deep": ("lti3_3msmts", 16, 12),
        # 7 states with 16 hyperplanes declared (tables of up to 9 949 cells: accepted since the lean group kernel exists), first 8 MUs replayed
        "leo7_deep16": ("leo7", 10, 8)}


def deepen(outdir, name):
    from mceio import read_scenario
    src, steps, nrec = DEEP[name]
    s = read_scenario(os.path.join(outdir, src + ".mces"))
    s.steps = steps
    rng = np.random.RandomState(9000 + steps)
    extra = 2.0 * (rng.randint(0, 2**31 - 1, s.max_shape) + 1.0) / 2.0**31 - 1.0
    s.b_pert = np.concatenate([s.b_pert, extra])[: s.max_shape]
    s.rec = s.rec[:nrec]
    return name, s


if __name__ == "__main__":
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    made = all_scenarios() + [homing3()]
    for name, s in made:
        path = os.path.join(outdir, name + ".mces")
        write_scenario(path, s)
        print("wrote", path, "d=%d steps=%d records=%d" % (s.d, s.steps, len(s.rec)))
    for name in sorted(DEEP):
        if not os.path.exists(os.path.join(outdir, DEEP[name][0] + ".mces")):
            continue        # leo5.mces is recorded by oracle/_ref/ref_gen_leo5 (tools/make_golden.sh runs this script again after it)
        name, s = deepen(outdir, name)
        path = os.path.join(outdir, name + ".mces")
        write_scenario(path, s)
        print("wrote", path, "d=%d steps=%d max_shape=%d records=%d" % (s.d, s.steps, s.max_shape, len(s.rec)))
