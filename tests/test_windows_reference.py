"""The sliding-window bank against the REFERENCE's SlidingWindowManager on the CPU-only box: same check as
tests/test_gpu_windows.py::test_window_bank_matches_the_reference_window_manager, with the estimators running the sequential
emulation of the kernel bodies (tests/emu, test infrastructure).  Pins the host-side bank logic -- window selection by the C++
manager's rule, Speyer re-initialisation through the restated eigen-solver, log layout -- to the reference's own log files."""
import functools
import os

import numpy as np

from harness import ROOT, load_emu

def _load_best(path):
    """`<win_idx>:<values>` lines of the reference's bank logs (cauchy_windows.hpp:1378-1400)."""
    idx, rows = [], []
    for line in open(path):
        a, b = line.split(":")
        idx.append(int(a)); rows.append([float(v) for v in b.split()])
    return np.array(idx), np.array(rows)


def test_window_bank_over_emulated_kernels_matches_the_reference_window_manager(tmp_path):
    """The Python bank against the REFERENCE's SlidingWindowManager: tests/golden/winbank_cpu1/ holds the log files the
    unmodified NUM_CPUS=1 reference wrote for the inputs of src/window_manager.cpp (srand(11), 201 measurements, 8 windows;
    tests/dropin/winbank_dropin.cpp) together with the measurement sequence.  Replaying the measurements through
    SlidingWindowBank must select the same window at every step and write the same means / covariances / normalisation
    factors into the same log layout -- digit for digit (16 decimals): the estimators are bit-exact, the bank's Speyer
    initialisation restates the reference's own symmetric eigen-solver operation for operation (windows.py::sym_eig), and
    the windows get the root_point / b_pert the reference's window processes drew."""
    from cauchyfriendly_b200.estimator import CauchyEstimator as _Est
    CauchyEstimator = functools.partial(_Est, _lib=load_emu())
    from cauchyfriendly_b200.windows import SlidingWindowBank
    gold = os.path.join(ROOT, "tests", "golden", "winbank_cpu1")
    zs = np.loadtxt(os.path.join(gold, "msmts.txt"))
    Phi = np.array([[1.4, -0.6, -1.0], [-0.2, 1.0, 0.5], [0.6, -0.6, -0.2]])
    Gamma = np.array([.1, .3, -.2]); H = np.array([[1.0, .5, .2]])
    rb = np.loadtxt(os.path.join(gold, "root_point_b_pert.txt"))       # the vectors every window of the reference run drew with rand()
    bank = SlidingWindowBank(8, np.eye(3), [.10, .08, .05], np.zeros(3), Phi, None, Gamma, [.1], H, [.2], estimator_cls=CauchyEstimator,
                             est_kwargs=dict(root_point=rb[:3], b_pert=rb[3:]), log_dir=str(tmp_path / "logs"), log_windows=False, selection="cpp")
    for z in zs:
        bank.step([z])
    bank.shutdown()
    worst = {}
    for name in ("cond_means.txt", "cond_covars.txt", "norm_factors.txt", "cerr_cond_means.txt", "cerr_cond_covars.txt", "cerr_norm_factors.txt"):
        gi, gv = _load_best(os.path.join(gold, name))
        oi, ov = _load_best(str(tmp_path / "logs" / name))
        assert gv.shape == ov.shape and gi.shape == oi.shape, name
        assert np.array_equal(gi, oi), "%s: a different window was selected at steps %s" % (name, np.nonzero(gi != oi)[0][:10])
        rel = np.max(np.abs(gv - ov), axis=1) / np.max(np.abs(gv), axis=1)
        worst[name] = float(rel.max())
        assert np.array_equal(gv, ov), "%s: worst row deviates by %.3e (step %d)" % (name, rel.max(), int(rel.argmax()))
    gi, gv = _load_best(os.path.join(gold, "numeric_error_codes.txt"))
    oi, ov = _load_best(str(tmp_path / "logs" / "numeric_error_codes.txt"))
    assert np.array_equal(gv, ov)
    print("window bank vs reference manager, worst relative row deviation:", worst)
